#!/bin/bash
# Round 2, call 14 (1 GPU): where the bucket build (auxiliary stream) starts relative to the forward pass
# (NVSM_BUCKETS_AT 0 / 1 / 2) and main-stream priority (NVSM_MAIN_PRIO=-1), C2 and C3.
T=${1:-r2n}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" --timeline gpurun_out/timeline_${T}_$name.md > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b base
NVSM_BUCKETS_AT=1 b at1
NVSM_BUCKETS_AT=2 b at2
NVSM_MAIN_PRIO=-1 b prio
NVSM_MAIN_PRIO=-1 NVSM_BUCKETS_AT=2 b prio_at2
b base2
NVSM_BUCKETS_AT=2 b at2_2
NVSM_BUCKETS_AT=2 b at2_C3 --workload C3
b base_C3 --workload C3
NVSM_BUCKETS_AT=2 b at2_C5 --workload C5
b base_C5 --workload C5
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-14s %.4f e2e %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
for n in at2; do echo "== $n"; awk 'NR>4' gpurun_out/timeline_${T}_$n.md | tail -11; done
