#!/bin/bash
# Round 2, final 1-GPU call: whole -m gpu suite, smoke, C2 (cpu_baseline + TF32 alt) / C3 / C5 / C1 lines, the reference arm, the ncu
# launch list of `bench.py --steps 2`, one `--set full` capture of a steady-state step of C2, C3 and C5.
T=${1:-r2z}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 400 python -m pytest tests -m gpu -q -n 4 --maxfail=10 -k "not loss_curve_matches_oracle" > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -4 gpurun_out/pytest_$T.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_$T.log 2>&1; stamp "smoke rc=$?"
tail -2 gpurun_out/smoke_$T.log
b() { local name=$1; shift; timeout 200 python bench.py "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b C2 --steps 200 --warmup 20 --timeline gpurun_out/timeline_${T}_C2.md
timeout 100 python bench.py --impl reference --steps 20 --warmup 5 > gpurun_out/bench_ref_$T.json 2>/dev/null; stamp "reference arm rc=$?"
b C3 --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
b C5 --workload C5 --steps 200 --warmup 20 --no_cpu_baseline --no_alt
b C1 --workload C1 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
b default
timeout 60 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$T.csv \
    python bench.py --steps 2 --warmup 3 --no_cpu_baseline --no_alt --no_probes > gpurun_out/launches_$T.log 2>&1; stamp "launch list rc=$?"
for W in C2 C3 C5; do
  timeout 200 ncu --set full --clock-control none --import-source on --profile-from-start off -o gpurun_out/prof_${T}_$W -f \
      python scripts/profile_step.py --workload $W > gpurun_out/prof_${T}_$W.log 2>&1; stamp "ncu full $W rc=$?"
  ncu -i gpurun_out/prof_${T}_$W.ncu-rep --page raw --csv > gpurun_out/prof_${T}_$W.csv 2>/dev/null
  python profiles/summarize_ncu.py full gpurun_out/prof_${T}_$W.ncu-rep gpurun_out/prof_${T}_${W}_kernels_full.md
  [ "$W" != "C2" ] && rm -f gpurun_out/prof_${T}_$W.ncu-rep
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")) + ["gpurun_out/bench_ref_$T.json"]:
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        r = d.get("roofline")
        print("%-14s %10d %.4f e2e %.4f | clocks %s" % (f.split("bench_")[1][:-5], d["value"], d["ms_per_step"], d["e2e"].get("ms_per_step", 0), d.get("clocks")))
        if r:
            print("      roofline", r["kernel"], "hbm frac %.3f" % r["frac"], "dram_frac", r["dram_frac"], "l2", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r["l2"].items() if k in ("achieved", "peak", "frac")})
    except Exception as e:
        print(f, "ERR", e)
PY
