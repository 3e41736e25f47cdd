#!/bin/bash
# Round 2, call 30 (1 GPU): final C2 line (cpu_baseline, TF32 alt, probes) + reference arm + C3 / C5 / C1 lines with three live slots
# and the list sampler; sampler / facade tests.
T=${1:-r2D}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
b() { local name=$1; shift; timeout 200 python bench.py "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b C2 --steps 200 --warmup 20
timeout 100 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2>/dev/null; stamp "reference arm rc=$?"
b C3 --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
b C5 --workload C5 --steps 200 --warmup 20 --no_cpu_baseline --no_alt
b C1 --workload C1 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
timeout 300 python -m pytest tests/test_cpp_facade.py tests/test_gpu_parity.py -m gpu -q -n 4 --maxfail=5 > gpurun_out/pytest_$T.log 2>&1; stamp "tests rc=$?"
tail -3 gpurun_out/pytest_$T.log
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")) + ["gpurun_out/bench_ref_$T.json"]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-10s %10d %.4f e2e %10d %.4f | %s %s" % (f.split("bench_")[1][:-5], d["value"], d["ms_per_step"], d["e2e"].get("value", 0), d["e2e"].get("ms_per_step", 0), (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons")))
    except Exception as e:
        print(f, "ERR", e)
PY
