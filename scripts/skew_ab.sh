#!/bin/bash
# A/B of the heavy-row split of the pull updates under skewed ids (one GPU box).  Usage: bash scripts/skew_ab.sh <tag> [quick]
T=${1:-r1e}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 300 python -m pytest tests -m gpu -q -n 4 -k "not loss_curve" --maxfail=8 > gpurun_out/pytest_skew_$T.log 2>&1; stamp "suite rc=$?"
tail -8 gpurun_out/pytest_skew_$T.log
B="--steps 200 --warmup 20 --no_cpu_baseline --no_alt"
python bench.py $B > gpurun_out/ab_${T}_C2.json 2>/dev/null; stamp C2
NVSM_NO_HEAVY=1 python bench.py $B > gpurun_out/ab_${T}_C2_noheavy.json 2>/dev/null; stamp C2-noheavy
python bench.py $B --zipf_words 1.0 > gpurun_out/ab_${T}_C2_zipf.json 2>/dev/null; stamp C2zipf
python bench.py $B --workload C5 > gpurun_out/ab_${T}_C5.json 2>/dev/null; stamp C5
python bench.py $B --workload C3 > gpurun_out/ab_${T}_C3.json 2>/dev/null; stamp C3
python bench.py $B --workload C3 --zipf_words 1.0 > gpurun_out/ab_${T}_C3_zipf.json 2>/dev/null; stamp C3zipf
python - <<PY
import json, glob
for n in sorted(glob.glob("gpurun_out/ab_${T}_*.json")):
    try:
        d = json.load(open(n))
        print(n.split("ab_${T}_")[1], round(d["value"]), "ms", round(d["ms_per_step"], 4), "e2e_ms", round(d["e2e"]["ms_per_step"], 4), d["roofline"]["phase_ms"])
    except Exception as e:
        print(n, "ERR", e)
PY
