#!/bin/bash
# Round 2, call 13 (2 GPUs): exchanges fused into compute kernels (col_stats_reduce_finalize_kernel<true>, peer_sums_tail)
# vs the stand-alone one-block all-reduce launches (NVSM_NO_FUSED_XCHG=1): multi-GPU tests, bench at N=2 with parity_check;
# device sampler with 16-candidate chunks (bit-exactness tests).
T=${1:-r2m}
N=${2:-2}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 500 python -m pytest tests/test_multi_gpu.py -m gpu -q -x > gpurun_out/pytest_multi_$T.log 2>&1; stamp "multi-gpu tests (fused exchange) rc=$?"
tail -5 gpurun_out/pytest_multi_$T.log
timeout 300 python -m pytest tests/test_negative_sampler.py tests/test_gpu_parity.py tests/test_cpp_facade.py -m gpu -q -x -n 4 -k "sampler or sampled or cli or facade" > gpurun_out/pytest_sampler_$T.log 2>&1; stamp "sampler tests rc=$?"
tail -3 gpurun_out/pytest_sampler_$T.log
b() { local name=$1; shift; timeout 300 python bench.py --gpus $N --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b fused
NVSM_NO_FUSED_XCHG=1 b unfused
b fused2 --no_parity_check
NVSM_NO_FUSED_XCHG=1 b unfused2 --no_parity_check
timeout 200 python bench.py --gpus 1 --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes > gpurun_out/bench_${T}_1gpu.json 2> gpurun_out/bench_${T}_1gpu.err; stamp "bench 1 GPU rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-12s %.4f e2e %.4f" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"]), "strong", d.get("strong") and (round(d["strong"]["ms_per_step"], 4), round(d["strong"]["e2e"]["ms_per_step"], 4)),
              "parity", d.get("parity_check") and (d["parity_check"]["ok"], d["parity_check"]["max_rel_err"], d["parity_check"]["peer_exchange_error"]), d["clocks"].get("sm_mhz"))
        print("      phases", {k: v for k, v in d["roofline"]["phase_ms"].items() if v > 0})
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-1500:])
PY
