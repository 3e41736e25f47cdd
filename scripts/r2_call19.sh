#!/bin/bash
# Round 2, calls 19-20 (2 GPUs): fused grad_transform exchange with the block-level flag wait (call 20: push kernel on the main stream) vs ncclAllReduce (NVSM_NO_FUSED_GT=1).
T=${1:-r2s}
N=${2:-2}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
[ -n "$SKIP_TESTS" ] || timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi_$T.log 2>&1; stamp "multi-gpu tests rc=$?"
tail -3 gpurun_out/pytest_multi_$T.log
b() { local name=$1; shift; timeout 300 python bench.py --gpus $N --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b gtfused
NVSM_NO_FUSED_GT=1 b gtnccl --no_parity_check
b gtfused2 --no_parity_check
NVSM_NO_FUSED_GT=1 b gtnccl2 --no_parity_check
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-12s %.4f e2e %.4f" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"]), "strong", d.get("strong") and (round(d["strong"]["ms_per_step"], 4), round(d["strong"]["e2e"]["ms_per_step"], 4)), d["clocks"].get("sm_mhz"))
        if d.get("parity_check"): print("      parity", d["parity_check"]["ok"], d["parity_check"]["max_rel_err"], d["parity_check"]["peer_exchange_error"])
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-1500:])
PY
