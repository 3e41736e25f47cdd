#!/bin/bash
# Round 2, call 28 (1 GPU): rare-rejection path of the device sampler (two launches instead of six) vs the scan path
# (NVSM_SAMPLER_SCAN=1): bit-exactness tests, C2 / C3 e2e A/B, e2e timeline.
T=${1:-r2C}
mkdir -p gpurun_out
S=$(date +%s)
stamp() { echo "[$(( $(date +%s) - S )) s] $*"; }
timeout 400 python -m pytest tests/test_gpu_parity.py tests/test_negative_sampler.py tests/test_gpu_reference.py tests/test_cpp_facade.py -m gpu -q -n 4 --maxfail=10 -k "sampler or sampled or cli or reference_init or ids" > gpurun_out/pytest_$T.log 2>&1; stamp "sampler tests rc=$?"
tail -3 gpurun_out/pytest_$T.log
b() { local name=$1; shift; timeout 200 python bench.py --steps 200 --warmup 20 --no_cpu_baseline --no_alt --no_probes "$@" > gpurun_out/bench_${T}_$name.json 2> gpurun_out/bench_${T}_$name.err; stamp "bench $name rc=$?"; }
b list1 --timeline gpurun_out/timeline_${T}.md
NVSM_SAMPLER_SCAN=1 b scan1
b list2
NVSM_SAMPLER_SCAN=1 b scan2
b list_C3 --workload C3
NVSM_SAMPLER_SCAN=1 b scan_C3 --workload C3
b list_C1 --workload C1
NVSM_SAMPLER_SCAN=1 b scan_C1 --workload C1
python - <<PY
import json, glob
for f in sorted(glob.glob("gpurun_out/bench_${T}_*.json")):
    try:
        d = json.loads([l for l in open(f) if l.startswith("{")][-1])
        print("%-8s %.4f e2e %.4f clocks %s %s" % (f.split("bench_${T}_")[1][:-5], d["ms_per_step"], d["e2e"]["ms_per_step"], d["clocks"].get("sm_mhz"), d["clocks"].get("reasons")))
    except Exception as e:
        print(f, "ERR", e, open(f.replace(".json", ".err")).read()[-800:])
PY
sed -n "/host-fed/,\$p" gpurun_out/timeline_${T}.md | grep -E "h2d|gather|gemm_fwd" | tail -9
