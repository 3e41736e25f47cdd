#!/usr/bin/env python
"""One steady-state step of a bench workload between cudaProfilerStart / Stop, for
    ncu --set full --clock-control none --import-source on --profile-from-start off -o <out> python scripts/profile_step.py --workload C3
(the capture then holds exactly the launches of one step: bucket build, forward, backward, updates)."""
import argparse
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cunvsm_b200 as nv  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--workload", default="C2")
ap.add_argument("--gemm_mode", type=int, default=2)
ap.add_argument("--warmup", type=int, default=4)
ap.add_argument("--steps", type=int, default=1)
ap.add_argument("--e2e", action="store_true", help="profile the host-fed call (nvsm_step_sampled) instead of a staged step")
args = ap.parse_args()
w = dict(bench.WORKLOADS[args.workload])
method, mode = nv.UPDATE_METHODS[w["update_method"]]
desc = nv.ModelDesc(word_repr_size=w["dw"], entity_repr_size=w["dd"], batch_normalization=w["bn"],
                    nonlinearity=nv.NONLINEARITIES[w["nonlinearity"]], clip_sigmoid=True, bias_negative_samples=w["bias_neg"])
tc = nv.TrainConfig(batch_size=w["B"], window_size=w["n"], num_random_entities=w["z"], regularization_lambda=w["lam"],
                    learning_rate=w["lr"], update_method=method, adam_mode=mode)
m = nv.Model(w["V"], w["D"], desc, tc, gemm_mode=args.gemm_mode, num_batch_slots=4)
rng = nv.RNG(1)
m.initialize(rng)
if w.get("neg_zipf", 0.0) > 0.0:
    m.set_negative_distribution(nv.zipf_cdf(w["D"], w["neg_zipf"]))
batches = []
for k, (f, labels) in enumerate(bench.make_batches(w, w["B"], 1234, 4)):
    b = nv.Batch(w["B"], w["n"]).fill(f, labels)
    m.stage_batch(k, b, m.generate_labels(labels, rng))
    batches.append(b)
m.sampler_seed(rng)
step = (lambda it: m.step_sampled(batches[it % 4], w["lr"])) if args.e2e else (lambda it: m.train_step_staged(it % 4, w["lr"]))
for it in range(args.warmup):
    step(it)
m.synchronize()
torch.cuda.profiler.start()
for it in range(args.steps):
    step(args.warmup + it)
m.synchronize()
torch.cuda.profiler.stop()
print("profiled %d step(s) of %s, cost %.5f" % (args.steps, args.workload, m.last_cost()))
