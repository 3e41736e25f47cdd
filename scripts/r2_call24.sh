#!/bin/bash
# Round 2, call 25 (8 GPUs): C2 at N=8 (weak value + e2e, strong block, parity_check incl. the fused one-call step) and C5 at N=8
# (BASELINE configs[4]) with the exchanges fused into the compute kernels; the 8-rank multi-GPU tests.
T=${1:-r2y}
N=${2:-8}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 150 python bench.py --gpus $N --steps 100 --warmup 10 --no_alt --no_cpu_baseline --no_probes > gpurun_out/bench_${T}_${N}gpu.json 2> gpurun_out/bench_${T}_${N}gpu.err; stamp "bench C2 N=$N rc=$?"
timeout 90 python bench.py --gpus $N --workload C5 --steps 100 --warmup 10 --no_alt --no_cpu_baseline --no_probes > gpurun_out/bench_${T}_${N}gpu_C5.json 2> gpurun_out/bench_${T}_${N}gpu_C5.err; stamp "bench C5 N=$N rc=$?"
timeout 120 python -m pytest tests/test_multi_gpu.py -m gpu -q -x -k "8" > gpurun_out/pytest_multi_$T.log 2>&1; stamp "8-rank tests rc=$?"
tail -4 gpurun_out/pytest_multi_$T.log
python - <<PY
import json
for n in ["bench_${T}_${N}gpu", "bench_${T}_${N}gpu_C5"]:
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n) if l.startswith("{")][-1])
        print(n, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 4), d["clocks"])
        print("   strong", {k: v for k, v in (d.get("strong") or {}).items() if k != "note"})
        print("   parity", d.get("parity_check"))
    except Exception as e:
        print(n, "ERR", e)
        print(open("gpurun_out/%s.err" % n).read()[-2000:])
PY
