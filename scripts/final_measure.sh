#!/bin/bash
# One GPU box, the numbers and ncu evidence of a round: the C2 bench line (with cpu_baseline and the TF32 alt), the
# reference arm, the ncu launch list of `bench.py --steps 2` and one `--set full` capture of a steady-state step
# (bucket build + forward + backward + updates).   Usage: bash scripts/final_measure.sh <tag>   -> gpurun_out/*_<tag>*
T=${1:-r1i}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 120 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; stamp "bench rc=$?"
timeout 60 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref_$T.json 2>/dev/null; stamp "reference arm rc=$?"
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_alt > gpurun_out/launches_$T.log 2>&1; stamp "launch list rc=$?"
timeout 150 ncu --set full --clock-control none --import-source on -s 57 -c 34 -o gpurun_out/prof_${T}_step -f \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_alt > gpurun_out/prof_${T}_step.log 2>&1; stamp "full capture rc=$?"
python - <<PY
import json
for n in ["bench_$T", "bench_ref_$T"]:
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n) if l.startswith("{")][-1])
        print(n, round(d["value"]), round(d["ms_per_step"], 4), d.get("e2e", {}).get("value"), d.get("clocks"), d.get("cpu_baseline"))
    except Exception as e:
        print(n, "ERR", e)
PY
ls -la gpurun_out/prof_${T}_step.ncu-rep 2>/dev/null
