#!/bin/bash
# Round 2, calls 23 and 32 (1 GPU): score_ring_kernel variants, new (libnvsm_b200.so) vs the previous build (libnvsm_b200_prev.so) on one box;
# lane r = row r stores) vs the two-pass kernel (libnvsm_b200_prev.so) on one box; parity suites; ncu of the new kernel.
T=${1:-r2w}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
OLD=$PWD/cunvsm_b200/libnvsm_b200_prev.so
timeout 500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py tests/test_gpu_reference.py -m gpu -q -n 4 --maxfail=10 > gpurun_out/pytest_$T.log 2>&1; stamp "parity suites rc=$?"
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
timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:score_ring -o gpurun_out/prof_${T}_score_new -f \
    python scripts/profile_step.py --workload C2 > gpurun_out/prof_${T}_score_new.log 2>&1; stamp "ncu score rc=$?"
python profiles/summarize_ncu.py full gpurun_out/prof_${T}_score_new.ncu-rep gpurun_out/prof_${T}_score_new_full.md
ncu -i gpurun_out/prof_${T}_score_new.ncu-rep --page details --csv > gpurun_out/prof_${T}_score_new_details.csv 2>/dev/null
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-8s %.4f e2e %.4f score %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["phase_ms"]["score_loss_bwd"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
cat gpurun_out/prof_${T}_score_new_full.md
grep -E "Issued Instructions|Issue Slots Busy" gpurun_out/prof_${T}_score_new_details.csv | cut -d, -f12-16
