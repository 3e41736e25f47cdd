#!/bin/bash
# Round 2, call 8 (1 GPU): --check_gradients CLI tests, C5 / C3 / C1 with the lane-parallel empty-row scan, fresh ncu --set full
# captures of one C2 / C3 / C5 step (traffic.json) and the ncu launch list of `bench.py --steps 2`.
T=${1:-r2h}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 300 python -m pytest tests/test_cpp_facade.py tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -n 4 --maxfail=10 -k "not loss_curve" > gpurun_out/pytest_$T.log 2>&1; stamp "pytest rc=$?"
tail -6 gpurun_out/pytest_$T.log
b() { local name=$1; shift; timeout 150 python bench.py "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b C5 --workload C5 --steps 200 --warmup 20 --no_cpu_baseline --no_alt
b C3 --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
b C1 --workload C1 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
for W in C2 C3 C5; do
  timeout 240 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_${T}_$W -f \
      python scripts/profile_step.py --workload $W > gpurun_out/prof_${T}_$W.log 2>&1; stamp "ncu full $W rc=$?"
  ncu -i gpurun_out/prof_${T}_$W.ncu-rep --page raw --csv > gpurun_out/prof_${T}_$W.csv 2>/dev/null
  python profiles/summarize_ncu.py full gpurun_out/prof_${T}_$W.ncu-rep gpurun_out/prof_${T}_${W}_kernels_full.md
  rm -f gpurun_out/prof_${T}_$W.ncu-rep
done
timeout 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_alt --no_probes > gpurun_out/launches_$T.log 2>&1; stamp "ncu launch list rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        r = d["roofline"]
        print("%-14s %10d %.4f e2e %.4f | %s frac %.3f" % (f.split("bench_${T}_")[1][:-5], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"], r["kernel"], r["frac"]))
        print("     ", {k: v for k, v in r["phase_ms"].items() if v > 0})
    except Exception as e:
        print(f, "ERR", e)
PY
