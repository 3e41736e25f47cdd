#!/bin/bash
# Round 2, call 24 (1 GPU): L2 prefetch of theta / m / v at the top of a row in adam_full_pull_kernel + ids loaded one n-gram ahead in
# gather_mean_lanes_kernel (libnvsm_b200.so) vs the previous build (libnvsm_b200_prev.so), same box, interleaved; parity suites.
T=${1:-r2x}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
OLD=$PWD/cunvsm_b200/libnvsm_b200_prev.so
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -n 4 --maxfail=10 > gpurun_out/pytest_$T.log 2>&1; stamp "parity suites rc=$?"
tail -3 gpurun_out/pytest_$T.log
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b new1
NVSM_LIB_PATH=$OLD b old1
b new2
NVSM_LIB_PATH=$OLD b old2
b new_C3 --workload C3
NVSM_LIB_PATH=$OLD b old_C3 --workload C3
b new_C5 --workload C5
NVSM_LIB_PATH=$OLD b old_C5 --workload C5
b new_C1 --workload C1
NVSM_LIB_PATH=$OLD b old_C1 --workload C1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        ph = d["roofline"]["phase_ms"]
        print("%-8s %.4f e2e %.4f gather %.4f ent %.4f words %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], ph["gather_mean"], ph["update_entities"], ph["update_words"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
