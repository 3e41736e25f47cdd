#!/bin/bash
# Round 2, call 26 (1 GPU): whole -m gpu suite with the new tests (timeline mode, PDL parity in a subprocess), knobs read once at
# create; C2 line.
T=${1:-r2A}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 500 python -m pytest tests -m gpu -q -n 4 --maxfail=10 -k "not loss_curve_matches_oracle" > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -6 gpurun_out/pytest_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; stamp "smoke rc=$?"
tail -2 gpurun_out/smoke_$T.log
timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes > gpurun_out/bench_${T}_C2.json 2> gpurun_out/bench_${T}_C2.err; stamp "bench rc=$?"
python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/bench_${T}_C2.json") if l.startswith("{")][-1])
print(d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"])
PY
