#!/bin/bash
# Round 2, call 11 (1 GPU): overlapped timeline of the fused C2 step (nvsm_set_profiling 2) and A/B of the grad_transform
# GEMM on a side stream under the word update (NVSM_GT_SIDE), with the no-overlap control.
T=${1:-r2k}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -x -n 4 > gpurun_out/pytest_$T.log 2>&1; stamp "pytest rc=$?"
tail -3 gpurun_out/pytest_$T.log
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" --timeline gpurun_out/timeline_${T}_$name.md > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b base
NVSM_GT_SIDE=1 b gtside
NVSM_NO_OVERLAP=1 b noverlap
b base2
NVSM_GT_SIDE=1 b gtside2
NVSM_GT_SIDE=1 b gtside_C3 --workload C3
b base_C3 --workload C3
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-14s %.4f e2e %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
for n in base gtside; do echo "== $n"; awk 'NR>4' gpurun_out/timeline_${T}_$n.md | tail -16; done
