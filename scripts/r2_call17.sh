#!/bin/bash
# Round 2, call 17 (1 GPU): PDL site masks, second pass (bn_backward site gated by the asynchronous entity update).
T=${1:-r2q}
mkdir -p gpurun_out
S=$(date +%s)
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; }
i=0
for mask in 0 33 35 51 55 63 7 0 63 55 7; do i=$((i+1)); NVSM_PDL=$mask b C2_${i}_m${mask}; done
for mask in 0 63 0 63; do i=$((i+1)); NVSM_PDL=$mask b C3_${i}_m${mask} --workload C3; done
for mask in 0 63 0 63; do i=$((i+1)); NVSM_PDL=$mask b C1_${i}_m${mask} --workload C1; done
for mask in 0 63; do i=$((i+1)); NVSM_PDL=$mask b C5_${i}_m${mask} --workload C5; done
echo "[$(( $(date +%s) - S )) s] done"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-22s %.4f e2e %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
