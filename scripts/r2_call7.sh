#!/bin/bash
# Round 2, call 7 (1 GPU): NULL-weights parity, the no-op decay skip (C3 / C5), fused-statistics A/B interleaved, suite.
T=${1:-r2g}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 400 python -m pytest tests -m gpu -q -n 4 --maxfail=10 -k "not loss_curve_matches_oracle" > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -6 gpurun_out/pytest_$T.log
b() { # name env -- args
  local name=$1 e=$2; shift 2
  env $e timeout 150 python bench.py "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"
}
Q="--steps 300 --warmup 30 --no_cpu_baseline --no_alt --no_probes"
b C2_fuse_1 X=1 $Q; b C2_nofuse_1 NVSM_NO_FUSED_STATS=1 $Q; b C2_fuse_2 X=1 $Q; b C2_nofuse_2 NVSM_NO_FUSED_STATS=1 $Q
b C2_fuse_3 X=1 $Q; b C2_nofuse_3 NVSM_NO_FUSED_STATS=1 $Q
b C3 X=1 --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
b C5 X=1 --workload C5 --steps 200 --warmup 20 --no_cpu_baseline --no_alt
b C1 X=1 --workload C1 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        r = d["roofline"]
        print("%-14s %10d %.4f e2e %.4f h2d %d | %s frac %.3f l2 %s" % (f.split("bench_${T}_")[1][:-5], d["value"], d["ms_per_step"], d["e2e"]["ms_per_step"],
              d["e2e"]["h2d_bytes_per_step"], r["kernel"], r["frac"], r["l2"]["frac"]))
        if "fuse" not in f: print("     ", {k: v for k, v in r["phase_ms"].items() if v > 0})
    except Exception as e:
        print(f, "ERR", e)
PY
