"""Host-timed phases of the reference CUDA step (oracle/_ref) on C2, with a device sync after each phase.
Run on the GPU box: python scripts/ref_phase_timing.py [steps]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
from oracle import binding as O  # noqa: E402
from oracle import ref_binding as R  # noqa: E402

w = dict(bench.WORKLOADS[os.environ.get("WORKLOAD", "C2")])
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 10
B = w["B"]
um = {"sgd": (R.SGD, 0), "adagrad": (R.ADAGRAD, 0), "full_adam": (R.ADAM, R.DENSE_UPDATE_DENSE_VARIANCE)}[w["update_method"]]
m = R.Model(w["V"], w["D"], w["dw"], w["dd"], batch_size=B, window_size=w["n"], num_random_entities=w["z"],
            nonlinearity=R.HARD_TANH if w["nonlinearity"] == "hard_tanh" else R.TANH, batch_normalization=w["bn"],
            clip_sigmoid=True, bias_negative_samples=w["bias_neg"], update_method=um[0], adam_mode=um[1],
            regularization_lambda=w["lam"], seed=1, dtype=np.float32)
fw, iw = np.ones((B, w["n"]), np.float32), np.ones(B, np.float32)
batches = [m.new_batch().fill(f, l, fw, iw) for f, l in bench.make_batches(w, B, 1234, 4)]
acc = dict(sampler_host_only=0.0, forward=0.0, get_cost=0.0, gradients=0.0, update=0.0)
for it in range(steps + 3):
    b = batches[it % 4]
    t = [time.perf_counter()]
    m.forward(b); m.synchronize(); t.append(time.perf_counter())
    m.get_cost(); t.append(time.perf_counter())
    m.compute_gradients(); m.synchronize(); t.append(time.perf_counter())
    m.update(w["lr"], m.scaled_lambda()); m.synchronize(); t.append(time.perf_counter())
    t0 = time.perf_counter(); O.generate_labels(np.zeros(B, np.int64), w["z"], w["D"], 1); ts = time.perf_counter() - t0
    if it >= 3:
        for k, d in zip(("forward", "get_cost", "gradients", "update"), np.diff(t)):
            acc[k] += d
        acc["sampler_host_only"] += ts
print({k: round(1e3 * v / steps, 3) for k, v in acc.items()}, "ms per step;",
      "sum", round(1e3 * sum(v for k, v in acc.items() if k != "sampler_host_only") / steps, 3))
