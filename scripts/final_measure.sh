#!/bin/bash
# One GPU box: benches of every BASELINE config (ours + reference arm), ncu launch list and full capture of one C2 step.
# Usage: bash scripts/final_measure.sh <tag>   -> gpurun_out/*_<tag>*
T=${1:-r1c}
python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err
python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$T.json 2>/dev/null
for w in C1 C3 C5; do
  python bench.py --workload $w --steps 100 --warmup 10 --no_cpu_baseline --no_alt > gpurun_out/bench_${T}_$w.json 2>/dev/null
  python bench.py --impl reference --workload $w --steps 10 --warmup 3 > gpurun_out/bench_ref_${T}_$w.json 2>/dev/null
done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_alt > gpurun_out/launches_$T.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -s 140 -c 30 -o gpurun_out/prof_${T}_step -f \
    python bench.py --steps 3 --warmup 3 --no_cpu_baseline --no_alt > gpurun_out/prof_${T}_step.log 2>&1
python - <<PY
import json
for n in ["bench_$T","bench_ref_$T","bench_${T}_C1","bench_ref_${T}_C1","bench_${T}_C3","bench_ref_${T}_C3","bench_${T}_C5","bench_ref_${T}_C5"]:
    try:
        d=json.load(open("gpurun_out/%s.json"%n)); print(n, round(d["value"]), round(d["ms_per_step"],4), d.get("e2e",{}).get("value"))
    except Exception as e: print(n, "ERR", e)
PY
