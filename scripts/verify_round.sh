#!/bin/bash
# One GPU box, budget-aware verification of the tree: new / changed paths first, the rest of the -m gpu suite on 4
# workers (the 1000-step loss curves are a separate call: LONG=1), smoke, the C2 bench line, the C5 (Zipf negatives)
# line and the ncu launch list of a C2 step.   Usage: bash scripts/verify_round.sh <tag>   -> gpurun_out/*_<tag>*
T=${1:-r1d}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
if [ "${LONG:-0}" = "1" ]; then
  timeout 420 python -m pytest tests -m gpu -q -n 4 -k "loss_curve" > gpurun_out/pytest_long_$T.log 2>&1; stamp "long curves rc=$?"
  tail -3 gpurun_out/pytest_long_$T.log
  exit 0
fi
timeout 200 python -m pytest tests/test_negative_sampler.py tests/test_cpp_facade.py tests/test_gpu_parity.py -m gpu -q \
    -k "sampler or cli or data_source or out_of_range or facade" --maxfail=5 > gpurun_out/pytest_new_$T.log 2>&1; stamp "new tests rc=$?"
tail -15 gpurun_out/pytest_new_$T.log
timeout 330 python -m pytest tests -m gpu -q -n 4 -k "not loss_curve" --maxfail=10 --durations=12 > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -25 gpurun_out/pytest_$T.log
timeout 90 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; stamp "smoke rc=$?"
tail -3 gpurun_out/smoke_$T.log
timeout 150 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; stamp "bench C2 rc=$?"
timeout 90 python bench.py --workload C5 --steps 100 --warmup 10 --no_cpu_baseline --no_alt > gpurun_out/bench_${T}_C5.json 2> gpurun_out/bench_${T}_C5.err; stamp "bench C5 rc=$?"
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_alt > gpurun_out/launches_$T.log 2>&1; stamp "ncu launch list rc=$?"
python - <<PY
import json
for n in ["bench_$T", "bench_${T}_C5"]:
    try:
        d = json.load(open("gpurun_out/%s.json" % n))
        print(n, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"], d["roofline"]["phase_ms"])
    except Exception as e:
        print(n, "ERR", e)
PY
