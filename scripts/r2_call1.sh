#!/bin/bash
# Round 2, first 1-GPU call: the whole -m gpu suite incl. the new full-size parity tests, smoke, bench lines of C2 / C3 /
# C5 with the new roofline, ncu --set full captures of one C3 and one C5 step.
T=${1:-r2a}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q -s --maxfail=10 > gpurun_out/pytest_fullsize_$T.log 2>&1; stamp "full-size parity rc=$?"
grep -E "^C[235] |passed|failed|Error|assert" gpurun_out/pytest_fullsize_$T.log | tail -40
timeout 400 python -m pytest tests -m gpu -q -n 4 --ignore=tests/test_gpu_fullsize.py --maxfail=10 --durations=8 > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -22 gpurun_out/pytest_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; stamp "smoke rc=$?"
tail -4 gpurun_out/smoke_$T.log
timeout 200 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_$T.json 2> gpurun_out/bench_$T.err; stamp "bench C2 rc=$?"
timeout 100 python bench.py --steps 20 --warmup 5 --no_alt --no_cpu_baseline > gpurun_out/bench_${T}_20steps.json 2> gpurun_out/bench_${T}_20steps.err; stamp "bench C2 (driver flags) rc=$?"
timeout 150 python bench.py --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt > gpurun_out/bench_${T}_C3.json 2> gpurun_out/bench_${T}_C3.err; stamp "bench C3 rc=$?"
timeout 100 python bench.py --workload C5 --steps 100 --warmup 10 --no_cpu_baseline --no_alt > gpurun_out/bench_${T}_C5.json 2> gpurun_out/bench_${T}_C5.err; stamp "bench C5 rc=$?"
for W in C3 C5 C2; do
  timeout 240 ncu --set full --clock-control none --profile-from-start off -o gpurun_out/prof_${T}_$W -f \
      python scripts/profile_step.py --workload $W > gpurun_out/prof_${T}_$W.log 2>&1; stamp "ncu full $W rc=$?"
  ncu -i gpurun_out/prof_${T}_$W.ncu-rep --page raw --csv > gpurun_out/prof_${T}_$W.csv 2>/dev/null
  python profiles/summarize_ncu.py full gpurun_out/prof_${T}_$W.ncu-rep gpurun_out/prof_${T}_${W}_kernels_full.md
  [ "$W" != "C2" ] && rm -f gpurun_out/prof_${T}_$W.ncu-rep
done
python - <<PY
import json
for n in ["bench_$T", "bench_${T}_20steps", "bench_${T}_C3", "bench_${T}_C5"]:
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n) if l.startswith("{")][-1])
        r = d["roofline"]
        print(n, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d["clocks"])
        print("   roofline", r["kernel"], round(r["frac"], 3), "l2", r["l2"], "copy", r["stream_copy_gbs_here"])
        print("   per_kernel", r["per_kernel"])
        print("   cpu", d.get("cpu_baseline"))
    except Exception as e:
        print(n, "ERR", e)
PY
ls -la gpurun_out/ | tail -30; du -sh gpurun_out
