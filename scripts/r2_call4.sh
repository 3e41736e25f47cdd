#!/bin/bash
# Round 2, call 4 (1 GPU): the class-surface test, the facade tests, GEMM microbenchmarks (fused statistics on / off).
T=${1:-r2d}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 120 ./cpp/classes_test > gpurun_out/classes_test_$T.log 2>&1; stamp "classes_test rc=$?"
grep -v "^ok  " gpurun_out/classes_test_$T.log | tail -30
timeout 300 python -m pytest tests/test_cpp_facade.py tests/test_gpu_parity.py tests/test_gpu_gemm.py -m gpu -q -x -n 4 > gpurun_out/pytest_$T.log 2>&1; stamp "pytest rc=$?"
tail -5 gpurun_out/pytest_$T.log
timeout 120 python scripts/bench_gemm.py > gpurun_out/gemm_$T.log 2>&1; stamp "gemm bench rc=$?"
cat gpurun_out/gemm_$T.log
NVSM_TC_2CTA=0 timeout 120 python scripts/bench_gemm.py > gpurun_out/gemm_1cta_$T.log 2>&1; stamp "gemm bench (1 CTA) rc=$?"
cat gpurun_out/gemm_1cta_$T.log
