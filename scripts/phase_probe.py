"""Print per-phase device times of the C2 step (or WORKLOAD=...) for the current environment knobs.
Usage (GPU box): NVSM_SCORE_S=1 python scripts/phase_probe.py [steps]"""
import json
import os
import subprocess
import sys

steps = sys.argv[1] if len(sys.argv) > 1 else "40"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--steps", steps, "--warmup", "5", "--no_cpu_baseline",
                      "--no_alt", "--workload", os.environ.get("WORKLOAD", "C2")], capture_output=True, text=True)
try:
    d = json.loads(out.stdout.strip().splitlines()[-1])
    ph = d["roofline"]["phase_ms"]
    print("ms/step %.4f  e2e %.4f | " % (d["ms_per_step"], d["e2e"]["ms_per_step"]) +
          " ".join("%s=%.1f" % (k[:12], 1e3 * v) for k, v in ph.items() if v > 0))
except Exception as e:
    print("FAILED", e, out.stdout[-500:], out.stderr[-1500:])
