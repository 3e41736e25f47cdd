#!/bin/bash
# Round 2, third call (1 GPU): control experiment for the all-gather soak (world = 1: two identical single-GPU runs, no
# communication), the -m gpu suite on the new kernels (fused BN statistics in the GEMM epilogue, pipelined pull kernels),
# and A/B bench lines.
T=${1:-r2c}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 500 python -m pytest tests -m gpu -q -n 4 --maxfail=10 --durations=5 > gpurun_out/pytest_$T.log 2>&1; stamp "suite rc=$?"
tail -12 gpurun_out/pytest_$T.log
NVSM_TEST_GEMM_MODE=0 NVSM_TEST_SPARSE_MODE=1 NVSM_TEST_SOAK=${SOAK:-8} timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 \
    --master-addr 127.0.0.1 --master-port 29701 tests/dist_worker.py > gpurun_out/soak_control_$T.log 2>&1; stamp "control soak rc=$?"
grep -E "SOAK|Error|error" gpurun_out/soak_control_$T.log | tail -4
mv gpurun_out/allgather_soak_w1.json gpurun_out/allgather_soak_control_$T.json 2>/dev/null
run() { # name, env..., -- bench args
  local name=$1; shift
  local envs=()
  while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" timeout 150 python bench.py "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"
}
run C2 X=1 -- --steps 200 --warmup 20
run C2_nofuse NVSM_NO_FUSED_STATS=1 -- --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
run C2_adampipe NVSM_ADAM_PIPE=1 -- --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
run C3 X=1 -- --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
run C3_nopipe NVSM_SGD_PIPE=0 -- --workload C3 --steps 100 --warmup 10 --no_cpu_baseline --no_alt --no_probes
run C5 X=1 -- --workload C5 --steps 100 --warmup 10 --no_cpu_baseline --no_alt
run C5_pipe NVSM_SGD_PIPE=1 -- --workload C5 --steps 100 --warmup 10 --no_cpu_baseline --no_alt --no_probes
run C1 X=1 -- --workload C1 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        r = d["roofline"]
        print(f.split("bench_${T}_")[1][:-5], round(d["value"]), round(d["ms_per_step"], 4), "e2e", round(d["e2e"]["value"]), "|", r["kernel"], "frac", round(r["frac"], 3),
              "l2", {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r["l2"].items() if k in ("achieved", "peak", "frac", "gather_probe_gbs", "read_probe_gbs")})
        print("     ", {k: v for k, v in r["phase_ms"].items() if v > 0})
    except Exception as e:
        print(f, "ERR", e)
PY
