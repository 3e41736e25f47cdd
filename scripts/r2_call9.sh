#!/bin/bash
# Round 2, call 9 (1 GPU): --check_gradients CLI tests, C5 / C3 after the round-robin scan, ncu --set full of C3 / C5.
T=${1:-r2i}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 300 python -m pytest tests/test_cpp_facade.py tests/test_gpu_parity.py -m gpu -q -n 4 --maxfail=10 > gpurun_out/pytest_$T.log 2>&1; stamp "pytest rc=$?"
tail -6 gpurun_out/pytest_$T.log
grep -E "Parameter .* of|Gradient check" gpurun_out/pytest_$T.log | head -20
b() { local name=$1; shift; timeout 150 python bench.py "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b C5 --workload C5 --steps 200 --warmup 20 --no_cpu_baseline --no_alt
b C3 --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
b C5u --workload C5 --zipf_negatives 0 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
for W in C3 C5; do
  timeout 240 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_${T}_$W -f \
      python scripts/profile_step.py --workload $W > gpurun_out/prof_${T}_$W.log 2>&1; stamp "ncu full $W rc=$?"
  ncu -i gpurun_out/prof_${T}_$W.ncu-rep --page raw --csv > gpurun_out/prof_${T}_$W.csv 2>/dev/null
  python profiles/summarize_ncu.py full gpurun_out/prof_${T}_$W.ncu-rep gpurun_out/prof_${T}_${W}_kernels_full.md
  rm -f gpurun_out/prof_${T}_$W.ncu-rep
done
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
