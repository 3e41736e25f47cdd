#!/bin/bash
# Round 2, call 16 (1 GPU): which programmatic-dependent-launch sites pay (NVSM_PDL bit mask: 1 gemm_fwd, 2 stats, 4 score,
# 8 bn_backward, 16 gemm_gT, 32 gemm_gP), C2 / C3 / C1.
T=${1:-r2p}
mkdir -p gpurun_out
S=$(date +%s)
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; }
for mask in 0 1 2 4 8 16 32 7 56 63 0; do NVSM_PDL=$mask b C2_m${mask}_$RANDOM; done
for mask in 0 7 15 63; do NVSM_PDL=$mask b C3_m${mask}_$RANDOM --workload C3; done
for mask in 0 7 15 63; do NVSM_PDL=$mask b C1_m${mask}_$RANDOM --workload C1; done
for mask in 0 7 15 63; do NVSM_PDL=$mask b C5_m${mask}_$RANDOM --workload C5; done
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
