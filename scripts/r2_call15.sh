#!/bin/bash
# Round 2, call 15 (1 GPU): loss written to the pinned ring by the score kernels' last block (no D2H copy in the stream),
# accumulator zeroing folded into the statistics kernel, programmatic dependent launches (NVSM_PDL=1) A/B; whole suite.
T=${1:-r2o}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 500 python -m pytest tests -m gpu -q -n 4 --maxfail=10 -k "not loss_curve_matches_oracle" > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -6 gpurun_out/pytest_$T.log
NVSM_PDL=1 timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference.py -m gpu -q -n 4 --maxfail=10 > gpurun_out/pytest_pdl_$T.log 2>&1; stamp "suite (PDL) rc=$?"
tail -4 gpurun_out/pytest_pdl_$T.log
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" --timeline gpurun_out/timeline_${T}_$name.md > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b base
NVSM_PDL=1 b pdl
b base2
NVSM_PDL=1 b pdl2
NVSM_PDL=1 b pdl_C3 --workload C3
b base_C3 --workload C3
NVSM_PDL=1 b pdl_C5 --workload C5
b base_C5 --workload C5
NVSM_PDL=1 b pdl_C1 --workload C1
b base_C1 --workload C1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-14s %.4f e2e %.4f clocks %s %s cost %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons"), d.get("final_cost")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
for n in pdl; do echo "== $n"; awk 'NR>4' gpurun_out/timeline_${T}_$n.md | tail -11; done
