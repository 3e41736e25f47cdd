#!/bin/bash
# Round 2, call 21 (1 GPU): score_ring_kernel with one sweep over the staged rows (chain per group of four rows, Gp from
# the rows in registers): parity suites, C2 / C3 / C5 / C1 bench with phases.
T=${1:-r2u}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference.py tests/test_gpu_loss_curve.py -m gpu -q -n 4 --maxfail=10 > gpurun_out/pytest_$T.log 2>&1; stamp "parity suites rc=$?"
tail -4 gpurun_out/pytest_$T.log
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b C2
b C2b
b C3 --workload C3
b C5 --workload C5
b C1 --workload C1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-8s %.4f e2e %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
        print("      phases", {k: v for k, v in d["roofline"]["phase_ms"].items() if v > 0})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
