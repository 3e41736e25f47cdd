#!/bin/bash
# Round 2, 2-GPU call: soak of the all-gather parity worker with diagnostics (hard_tanh vs tanh in the Adam combinations),
# the multi-GPU tests, bench at N=2 with the strong-scaling block, parity_check and the e2e breakdown.
T=${1:-r2b}
N=${2:-2}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
NVSM_TEST_GEMM_MODE=0 NVSM_TEST_SPARSE_MODE=1 NVSM_TEST_SOAK=${SOAK:-12} timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29700 tests/dist_worker.py > gpurun_out/soak_$T.log 2>&1; stamp "soak rc=$?"
grep -E "SOAK|Error|error" gpurun_out/soak_$T.log | tail -8
timeout 400 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi_$T.log 2>&1; stamp "multi-gpu tests rc=$?"
tail -5 gpurun_out/pytest_multi_$T.log
timeout 300 python bench.py --gpus $N --steps 100 --warmup 10 --e2e_breakdown > gpurun_out/bench_${T}_${N}gpu.json 2> gpurun_out/bench_${T}_${N}gpu.err; stamp "bench C2 N=$N rc=$?"
timeout 200 python bench.py --gpus $N --workload C5 --steps 100 --warmup 10 --no_alt > gpurun_out/bench_${T}_${N}gpu_C5.json 2> gpurun_out/bench_${T}_${N}gpu_C5.err; stamp "bench C5 N=$N rc=$?"
python - <<PY
import json
for n in ["bench_${T}_${N}gpu", "bench_${T}_${N}gpu_C5"]:
    try:
        d = json.loads([l for l in open("gpurun_out/%s.json" % n) if l.startswith("{")][-1])
        print(n, round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 4), d["clocks"])
        print("   strong", d.get("strong"))
        print("   parity", d.get("parity_check"))
        print("   breakdown", d.get("e2e_breakdown"))
    except Exception as e:
        print(n, "ERR", e)
        print(open("gpurun_out/%s.err" % n).read()[-3000:])
PY
