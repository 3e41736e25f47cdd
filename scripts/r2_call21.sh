#!/bin/bash
# Round 2, call 22 (1 GPU): score_ring_kernel one-sweep (libnvsm_b200.so) vs two-pass (libnvsm_b200_twopass.so, NVSM_LIB_PATH) on the
# same box, interleaved; ncu --set full of the score kernel for both.
T=${1:-r2v}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
OLD=$PWD/cunvsm_b200/libnvsm_b200_twopass.so
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b new1
NVSM_LIB_PATH=$OLD b old1
b new2
NVSM_LIB_PATH=$OLD b old2
b new_C3 --workload C3
NVSM_LIB_PATH=$OLD b old_C3 --workload C3
b new_C5 --workload C5
NVSM_LIB_PATH=$OLD b old_C5 --workload C5
for v in new old; do
  if [ $v = old ]; then export NVSM_LIB_PATH=$OLD; else unset NVSM_LIB_PATH; fi
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:score_ring -o gpurun_out/prof_${T}_score_$v -f \
      python scripts/profile_step.py --workload C2 > gpurun_out/prof_${T}_score_$v.log 2>&1; stamp "ncu score $v rc=$?"
  python profiles/summarize_ncu.py full gpurun_out/prof_${T}_score_$v.ncu-rep gpurun_out/prof_${T}_score_${v}_full.md
  ncu -i gpurun_out/prof_${T}_score_$v.ncu-rep --page details --csv > gpurun_out/prof_${T}_score_${v}_details.csv 2>/dev/null
done
unset NVSM_LIB_PATH
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-8s %.4f e2e %.4f score %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["roofline"]["phase_ms"]["score_loss_bwd"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
cat gpurun_out/prof_${T}_score_new_full.md gpurun_out/prof_${T}_score_old_full.md
