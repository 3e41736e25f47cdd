#!/bin/bash
# Round 2, 8-GPU call: C2 at N=8 (weak value + e2e, strong block, parity_check vs unsharded and vs the reference, e2e
# breakdown) and C5 at N=8 (BASELINE configs[4]).
T=${1:-r2f}
N=${2:-8}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 400 python bench.py --gpus $N --steps 100 --warmup 10 --e2e_breakdown --no_alt > gpurun_out/bench_${T}_${N}gpu.json 2> gpurun_out/bench_${T}_${N}gpu.err; stamp "bench C2 N=$N rc=$?"
timeout 300 python bench.py --gpus $N --workload C5 --steps 100 --warmup 10 --no_alt --e2e_breakdown > gpurun_out/bench_${T}_${N}gpu_C5.json 2> gpurun_out/bench_${T}_${N}gpu_C5.err; stamp "bench C5 N=$N rc=$?"
python - <<PY
import json
for n in ["bench_${T}_${N}gpu", "bench_${T}_${N}gpu_C5"]:
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n) if l.startswith("{")][-1])
        print(n, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 4), d["clocks"])
        print("   strong", {k: v for k, v in (d.get("strong") or {}).items() if k != "note"})
        print("   parity", d.get("parity_check"))
        print("   breakdown", d.get("e2e_breakdown"))
        print("   config", d["config"].get("host_cores_per_rank"), d.get("collectives"))
    except Exception as e:
        print(n, "ERR", e)
        print(open("gpurun_out/%s.err" % n).read()[-3000:])
PY
nproc; lscpu | grep -E "Model name|Socket|NUMA" | head -5
