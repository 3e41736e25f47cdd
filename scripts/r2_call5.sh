#!/bin/bash
# Round 2, call 5 (1 GPU): eight epilogue warps in the tcgen05 GEMMs -- GEMM tests, microbenchmarks, C2 / C3 bench, class-surface test.
T=${1:-r2e}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 120 ./cpp/classes_test > gpurun_out/classes_test_$T.log 2>&1; stamp "classes_test rc=$?"
grep -v "^ok  " gpurun_out/classes_test_$T.log | tail -8
timeout 300 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_parity.py tests/test_cpp_facade.py -m gpu -q -x -n 4 > gpurun_out/pytest_$T.log 2>&1; stamp "pytest rc=$?"
tail -4 gpurun_out/pytest_$T.log
timeout 120 python scripts/bench_gemm.py > gpurun_out/gemm_$T.log 2>&1; stamp "gemm bench rc=$?"
cat gpurun_out/gemm_$T.log
for v in C2:"--steps 200 --warmup 20 --no_cpu_baseline" C2_nofuse:"--steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes" C3:"--workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt --no_probes"; do
  name=${v%%:*}; args=${v#*:}
  if [ "$name" = "C2_nofuse" ]; then export NVSM_NO_FUSED_STATS=1; else unset NVSM_NO_FUSED_STATS; fi
  timeout 150 python bench.py $args > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        r = d["roofline"]
        print(f.split("bench_${T}_")[1][:-5], round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), d.get("alt_single_pass_tf32"))
        print("     ", {k: v for k, v in r["phase_ms"].items() if v > 0})
    except Exception as e:
        print(f, "ERR", e)
PY
