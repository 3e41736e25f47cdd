#!/usr/bin/env python
"""bench.py — n-grams/sec of the NVSM training step (BASELINE.json metric) on N B200s.

A step = forward + NCE loss + backward + parameter update of one batch (the body of the
reference's iterate_data, cpp/main.cu:405-431). Default workload is BASELINE.json configs[1]:
NVSM hard_tanh + batch-norm, batch 51200, |V|=|D|=50k, d_w=300, d_d=256, n=10, z=10, full_adam
(the canonical optimiser, scripts/functions.sh:394), lambda=1e-2, lr=1e-3, synthetic uniform
ids, Glorot-initialised tables from minstd_rand0(seed 1), negatives from the reference's host
sampler.

  value : steps on batches already resident in HBM (10 staged batches, cycled), CUDA events.
  e2e   : the same steps through nvsm_step_sampled with pinned HOST buffers (H2D of word ids, weights
          and positive labels every step inside the timed region, negatives drawn by the bit-exact
          device sampler inside the step, loss read back every step, one step lagged).
  --impl reference : the reference's own step. The reference has no CPU path: its step is a CUDA program,
          so this arm runs the UNMODIFIED reference sources (oracle/_ref, compiled by oracle/ref_shim/Makefile
          over a reconstruction of the un-vendored device_matrix; cuBLAS + cuDNN) on GPU 0, full batch, host
          sampler and host<->device copies included. `--reference_kind cpu` (or a missing oracle/_ref) falls
          back to the CPU restatement (oracle port, OpenMP) on a bounded sample.

N > 1 (torchrun): one process per GPU; every rank runs its own shard of n-gram rows
(per-GPU batch fixed => weak scaling), batch-norm statistics and the dense projection
gradients are all-reduced with NCCL inside the step, sparse tables are updated per rank.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # BASELINE.json configs[1] — the configuration the metric is quoted on.
    "C2": dict(name="NVSM hard_tanh+BN B=51200 V=50k D=50k d_w=300 d_d=256 n=10 z=10",
               V=50000, D=50000, dw=300, dd=256, n=10, z=10, B=51200, nonlinearity="hard_tanh", bn=True,
               update_method="full_adam", lr=1e-3, lam=1e-2, bias_neg=False),
    # configs[2] — larger tables, Adagrad sparse updates.
    "C3": dict(name="NVSM hard_tanh+BN B=51200 V=200k D=500k d_w=300 d_d=256 n=10 z=16 adagrad",
               V=200000, D=500000, dw=300, dd=256, n=10, z=16, B=51200, nonlinearity="hard_tanh", bn=True,
               update_method="adagrad", lr=1e-2, lam=1e-2, bias_neg=False),
    # configs[4] — LSE with bias_negative_samples, large document table (per-GPU batch 4096).
    "C5": dict(name="LSE tanh bias_negative_samples B=4096 V=100k D=1M d=128 n=10 z=32 sgd, Zipf(1.0) negatives",
               V=100000, D=1000000, dw=128, dd=128, n=10, z=32, B=4096, nonlinearity="tanh", bn=False,
               update_method="sgd", lr=1e-2, lam=1e-2, bias_neg=True, neg_zipf=1.0),
    # configs[0] — the reference's own small LSE case.
    "C1": dict(name="LSE tanh B=4096 V=2k D=200 d=64 n=10 z=4", V=2000, D=200, dw=64, dd=64, n=10, z=4, B=4096,
               nonlinearity="tanh", bn=False, update_method="sgd", lr=1e-2, lam=1e-2, bias_neg=False),
}
NUM_BATCHES = 10


def algorithmic_bytes_per_ngram(w):
    """SURVEY.md §8d, per REFERENCE: gather W and E rows once, read-modify-write the same rows once, ids as
    delivered (8 B), weights (4 B). Every table row is referenced ~10x per batch and the tables fit the 126 MB L2, so
    these are bytes L2 serves: they are reported against the measured L2 gather peak (roofline.l2), never against
    the HBM peak."""
    n, R, dw, dd = w["n"], w["z"] + 1, w["dw"], w["dd"]
    per_phase = {
        "gather_mean": 4 * n * dw + 8 * n + 4 * n,
        "score_loss_bwd": 4 * R * dd + 8 * R + 4,
        "update_entities": 2 * 4 * R * dd,
        "update_words": 2 * 4 * n * dw,
    }
    if w["update_method"] == "full_adam":  # fused dense pass: theta, m, v read + written
        per_phase["update_entities"] += 24.0 * w["D"] * dd / w["B"]
        per_phase["update_words"] += 24.0 * w["V"] * dw / w["B"]
    return per_phase


def expected_unique(num_rows, references):
    """Expected number of distinct rows hit by `references` uniform draws over `num_rows`."""
    return num_rows * (1.0 - np.exp(-float(references) / num_rows))


def compulsory_bytes_per_launch(w, B, gemm_mode, unique_words=None, unique_entities=None, world=1):
    """Bytes that MUST cross HBM per launch of each gather-type kernel of the design that is built:
    every distinct table row once (re-references are L2 hits by construction), the per-step tensors the kernel
    streams, its index / weight arrays, and -- for the pull updates -- one read-modify-write of every state row the
    optimiser touches. This is the numerator of roofline.frac (against the measured HBM copy peak)."""
    n, R, dw, dd, V, D = w["n"], w["z"] + 1, w["dw"], w["dd"], w["V"], w["D"]
    uw = expected_unique(V, B * n) if unique_words is None else unique_words
    ue = expected_unique(D, B * R) if unique_entities is None else unique_entities
    tc = gemm_mode != 0 and dw % 4 == 0 and dd % 32 == 0
    ldp = (dw + 31) // 32 * 32 if tc else dw
    lo = 2 if (tc and gemm_mode == 2) else 1
    method = w["update_method"]
    # the dense decay of SGD / Adagrad multiplies by the float32 factor 1 - (lambda / B) * lr; when that rounds to exactly
    # 1.0f the pass is a bit-exact no-op and the library skips rows without references (C3, C5)
    lam_s = np.float32(w["lam"]) / np.float32(B * max(1, world))
    lam = w["lam"] > 0 and np.float32(1.0 - float(np.float32(lam_s * np.float32(w["lr"])))) != np.float32(1.0)
    pull = method == "full_adam" or (method in ("sgd", "adagrad") and max(V, D) >= 8192)
    out = {
        # W rows once, ids + weights, P (+ P_lo) written
        "gather_mean": 4 * uw * dw + 12 * B * n + 4 * B * ldp * lo,
        # Z read, E rows once, ids, instance weights; Gp, (Y), probs, mult written
        "score_loss_bwd": 4 * B * dd + 4 * ue * dd + 8 * B * R + 4 * B + 4 * B * dd + (4 * B * dd if pull else 0) + 8 * B * R,
    }
    def table(rows, dim, touched, refs, src_bytes):
        if method == "full_adam":
            state = 24 * rows * dim                      # theta, m, v: read + write, every row (decay)
        elif method in ("sgd", "adagrad"):
            state = 8 * (rows if lam else touched) * dim  # dense decay when lambda > 0 (cpp/storage.cu:65-67)
            if method == "adagrad":
                state += 8 * touched
        else:                                            # sparse / dense-update Adam: m dense, theta dense or touched
            state = 16 * rows * dim + 8 * rows
        return state + src_bytes + 4 * refs + 4 * rows   # + reference list + bucket offsets
    out["update_entities"] = table(D, dd, ue, B * R, 4 * B * dd + 4 * B * R)       # Y rows + multipliers
    out["update_words"] = table(V, dw, uw, B * n, 4 * B * dw + 4 * B * n)          # grad_phrase rows + word weights
    return out


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """SM clock / throttle reasons of one GPU, polled from a thread through NVML (nvidia_ml_py) every ~2 ms from before
    the warm-up until after the last timed loop; `window()` brackets the timed regions so that the summary is
    computed over samples taken DURING them (a 20-step timed region lasts ~13 ms, far below nvidia-smi's own polling
    period, which is why round 1 reported "no samples"). nvidia-smi -lms is the fallback when NVML cannot be loaded."""
    REASONS = (("hw_slowdown", 0x8), ("hw_thermal_slowdown", 0x40), ("sw_thermal_slowdown", 0x20), ("sw_power_cap", 0x4))

    def __init__(self, gpu_index):
        self.gpu, self.samples, self.windows, self._stop = gpu_index, [], [], False
        self.thread, self.how, self.smi = None, None, None

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.gpu)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM))
            reasons_fn = getattr(pynvml, "nvmlDeviceGetCurrentClocksEventReasons", None) or pynvml.nvmlDeviceGetCurrentClocksThrottleReasons

            def poll():
                while not self._stop:
                    try:
                        self.samples.append((time.perf_counter(), float(pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)),
                                             int(reasons_fn(h))))
                    except Exception:
                        pass
                    time.sleep(0.002)
            self.how = "nvml"
            self.thread = threading.Thread(target=poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            pass
        try:
            q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
                 "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
            self.smi = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + q, "--format=csv,noheader,nounits",
                                         "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)

            def pump():
                for line in self.smi.stdout:
                    parts = [x.strip() for x in line.split(",")]
                    try:
                        mask = sum(bit for (_, bit), val in zip(self.REASONS, parts[2:6]) if val.lower().startswith("active"))
                        self.max_mhz = float(parts[1])
                        self.samples.append((time.perf_counter(), float(parts[0]), mask))
                    except Exception:
                        continue
            self.how = "nvidia-smi -lms 20"
            self.thread = threading.Thread(target=pump, daemon=True)
            self.thread.start()
        except Exception:
            self.smi = None

    def window(self, t0, t1):
        self.windows.append((t0, t1))

    def stop(self):
        self._stop = True
        if self.smi:
            self.smi.terminate()
        if self.thread:
            self.thread.join(timeout=2)
        if not self.how:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no sampler (NVML and nvidia-smi unavailable)"]}
        timed = [x for x in self.samples if any(a <= x[0] <= b for a, b in self.windows)]
        # a timed region shorter than the polling period: fall back to the samples between the first warm-up step
        # and the end of the last timed loop (the GPU runs the same steps back to back there)
        scope = "timed regions"
        if len(timed) < 3 and self.windows:
            lo = min(a for a, _ in self.windows) - getattr(self, "lead_s", 0.0)
            timed = [x for x in self.samples if lo <= x[0] <= max(b for _, b in self.windows)]
            scope = "warm-up + timed regions"
        if not timed:
            return {"sm_mhz": None, "sm_max_mhz": getattr(self, "max_mhz", None), "reasons": ["no samples"], "how": self.how}
        mask = 0
        for x in timed:
            mask |= x[2]
        return {"sm_mhz": float(np.median([x[1] for x in timed])), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(nm for nm, bit in self.REASONS if mask & bit), "samples": len(timed),
                "samples_total": len(self.samples), "scope": scope, "how": self.how}


def zipf_ids(rng, cdf, size):
    return np.minimum(np.searchsorted(cdf, rng.random(size), side="right"), cdf.size - 1).astype(np.int64)


def make_batches(w, B, seed, num):
    """Uniform word ids and positive document ids; word ids ~ Zipf(s) when the workload carries
    word_zipf = s (the C3 gather/scatter sweep variant, --zipf_words)."""
    rng = np.random.default_rng(seed)
    out = []
    word_cdf = None
    if w.get("word_zipf", 0.0) > 0.0:
        import cunvsm_b200 as nv
        word_cdf = nv.zipf_cdf(w["V"], w["word_zipf"])
    for _ in range(num):
        if word_cdf is not None:
            f = zipf_ids(rng, word_cdf, (B, w["n"]))
        else:
            f = rng.integers(0, w["V"], size=(B, w["n"]), dtype=np.int64)
        labels = rng.integers(0, w["D"], size=B, dtype=np.int64)
        out.append((f, labels))
    return out


# -------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the oracle port on the host cores
# -------------------------------------------------------------------------------------------
def cpu_port_run(w, steps, warmup, sample_B):
    from oracle import binding as O
    native = True
    try:
        O.build(native=True)
    except Exception:
        native = False
    um = {"sgd": (O.SGD, 0), "adagrad": (O.ADAGRAD, 0), "sparse_adam": (O.ADAM, O.SPARSE),
          "dense_adam": (O.ADAM, O.DENSE_UPDATE), "full_adam": (O.ADAM, O.DENSE_UPDATE_DENSE_VARIANCE)}[w["update_method"]]
    m = O.Model(w["V"], w["D"], w["dw"], w["dd"], nonlinearity=O.HARD_TANH if w["nonlinearity"] == "hard_tanh" else O.TANH,
                batch_normalization=w["bn"], clip_sigmoid=True, bias_negative_samples=w["bias_neg"],
                update_method=um[0], adam_mode=um[1], num_random_entities=w["z"], regularization_lambda=w["lam"],
                dtype=np.float32, native=native)
    state = m.initialize(1)
    batches = make_batches(w, sample_B, 1234, max(2, min(NUM_BATCHES, steps + warmup)))
    fw = np.ones((sample_B, w["n"]), np.float32)
    iw = np.ones(sample_B, np.float32)
    cores = O.lib(native).oracle_num_threads()
    neg_cdf = None
    if w.get("neg_zipf", 0.0) > 0.0:
        import cunvsm_b200 as nv
        neg_cdf = nv.zipf_cdf(w["D"], w["neg_zipf"])
    times = []
    for it in range(warmup + steps):
        f, labels = batches[it % len(batches)]
        t0 = time.perf_counter()
        if neg_cdf is not None:
            ids, state = O.generate_labels_cdf(labels, w["z"], neg_cdf, state)
        else:
            ids, state = O.generate_labels(labels, w["z"], w["D"], state)   # serial host sampler, as the reference
        m.compute_cost(f, fw, ids, iw, w["n"])
        m.compute_gradients()
        m.update(w["lr"], m.scaled_lambda())
        dt = time.perf_counter() - t0
        if it >= warmup:
            times.append(dt)
    sec = float(np.sum(times))
    return dict(value=sample_B * steps / sec, ms_per_step=1e3 * sec / steps, cores=cores, native=native)


def reference_cuda_run(w, steps, warmup):
    """The UNMODIFIED reference step (oracle/_ref, built from /root/reference/cpp by oracle/ref_shim/Makefile over a
    reconstructed device_matrix) on GPU 0: full batch, pinned host Batch -> H2D, host sampler, compute_cost,
    compute_gradients, update, blocking get_cost -- the body of iterate_data (cpp/main.cu:405-444)."""
    from oracle import ref_binding as R
    um = {"sgd": (R.SGD, 0), "adagrad": (R.ADAGRAD, 0), "sparse_adam": (R.ADAM, R.SPARSE),
          "dense_adam": (R.ADAM, R.DENSE_UPDATE), "full_adam": (R.ADAM, R.DENSE_UPDATE_DENSE_VARIANCE)}[w["update_method"]]
    B = w["B"]
    m = R.Model(w["V"], w["D"], w["dw"], w["dd"], batch_size=B, window_size=w["n"], num_random_entities=w["z"],
                nonlinearity=R.HARD_TANH if w["nonlinearity"] == "hard_tanh" else R.TANH, batch_normalization=w["bn"],
                clip_sigmoid=True, bias_negative_samples=w["bias_neg"], update_method=um[0], adam_mode=um[1],
                regularization_lambda=w["lam"], seed=1, dtype=np.float32)
    fw, iw = np.ones((B, w["n"]), np.float32), np.ones(B, np.float32)
    batches = [m.new_batch().fill(f, labels, fw, iw) for f, labels in make_batches(w, B, 1234, NUM_BATCHES)]
    cost = None
    for it in range(warmup):
        cost = m.step(batches[it % NUM_BATCHES], w["lr"])
    m.synchronize()
    t0 = time.perf_counter()
    for it in range(steps):
        cost = m.step(batches[it % NUM_BATCHES], w["lr"])   # ends in the reference's blocking get_cost()
    m.synchronize()
    sec = time.perf_counter() - t0
    return dict(value=B * steps / sec, ms_per_step=1e3 * sec / steps, final_cost=cost)


def run_reference(args, w, rank):
    if rank != 0:
        return
    use_cuda_ref = False
    if args.reference_kind != "cpu":
        try:
            import torch
            from oracle import ref_binding as R
            use_cuda_ref = R.available(np.float32) and torch.cuda.is_available()
        except Exception:
            use_cuda_ref = False
    base = {"impl": "reference", "metric": "n-grams/sec", "unit": "n-grams/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic"}
    if use_cuda_ref:
        r = reference_cuda_run(w, args.steps, max(args.warmup, 3))
        B = w["B"]
        h2d = int(B * w["n"] * 8 + B * w["n"] * 4 + B * 4 + B * (w["z"] + 1) * 8)
        sample = ("full %d-n-gram batch per step on GPU 0: the reference has no CPU implementation, its step IS a CUDA "
                  "program; one host thread samples negatives and issues it (cpp/main.cu:405-444)" % B)
        line = dict(base, value=r["value"], ms_per_step=r["ms_per_step"],
                    config={"workload": w["name"], "update_method": w["update_method"], "per_gpu_batch": B,
                            "global_batch": B,
                            "reference_build": "unmodified /root/reference/cpp/*.cu, nvcc -O3 -use_fast_math float32 NDEBUG (the reference's "
                                               "flags except -default-stream per-thread, CMakeLists.txt:73: one host thread issues "
                                               "the step here, so the legacy default stream orders the same launches identically), "
                                               "cuBLAS SGEMM + cuDNN batch-norm, over oracle/ref_shim's reconstruction of "
                                               "the un-vendored device_matrix (one Thrust/CUDA kernel per op, size-bucketed caching "
                                               "pool for cnmem)"},
                    cpu_baseline={"value": r["value"], "unit": "n-grams/s", "cores": 1, "kind": "reference", "sample": sample},
                    e2e={"value": r["value"], "unit": "n-grams/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 4},
                    final_cost=r["final_cost"],
                    note="single GPU only: the reference has no multi-GPU path; at N>1 this is still one GPU"
                         + ("; negatives are uniform: UniformLabelGenerator is the reference's only generator"
                            if w.get("neg_zipf", 0.0) > 0.0 else ""))
        print(json.dumps(line), flush=True)
        return
    sample_B = min(args.cpu_sample, w["B"]) if args.cpu_sample > 0 else min(5120, w["B"])
    r = cpu_port_run(w, args.steps, max(args.warmup, 1), sample_B)
    sample = ("%d n-grams/step (1/%d of the %d batch), full-size tables, float32, sampler+forward+backward+update; "
              "%s build" % (sample_B, max(1, w["B"] // sample_B), w["B"], "-march=native" if r["native"] else "portable"))
    line = dict(base, value=r["value"], ms_per_step=r["ms_per_step"],
                config={"workload": w["name"], "update_method": w["update_method"], "sample_batch": sample_B},
                cpu_baseline={"value": r["value"], "unit": "n-grams/s", "cores": r["cores"], "kind": "port", "sample": sample},
                e2e={"value": r["value"], "unit": "n-grams/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                note="CPU restatement (oracle port) of the reference step on the host cores; used when oracle/_ref (the "
                     "reference's CUDA step) is not built or no GPU is visible.")
    print(json.dumps(line), flush=True)


# -------------------------------------------------------------------------------------------
# our arm
# -------------------------------------------------------------------------------------------
def pin_rank_to_cores(local_rank, world):
    """N processes on one node: give every rank its own slice of the host cores (launch thread, NCCL proxy and clock
    sampler stay off each other's cores)."""
    try:
        cores = sorted(os.sched_getaffinity(0))
        per = len(cores) // world
        if world > 1 and per >= 1:
            os.sched_setaffinity(0, set(cores[local_rank * per:(local_rank + 1) * per]))
            return per
    except Exception:
        pass
    return None


def rel_err(a, b):
    a = np.asarray(a, np.float64).ravel(); b = np.asarray(b, np.float64).ravel()
    return float(np.abs(a - b).max() / max(np.abs(b).max(), 1e-30))


def parity_check(args, w, nv, sharding, dist, rank, world, local_rank):
    """N > 1, outside every timed region: ONE step of the workload's global batch (w["B"] rows) sharded over the
    `world` ranks, compared on rank 0 with (a) the same library on one GPU over the whole batch and (b) the
    reference itself (oracle/_ref: the unmodified cpp/*.cu, float32) -- loss, grad_transform, grad_bias, the
    batch-norm statistics, and T / b after the update. Same seed, same bit-exact negatives everywhere."""
    Bg = w["B"]
    if Bg % world:
        return {"ok": None, "skipped": "global batch %d does not split over %d ranks" % (Bg, world)}
    method, mode = nv.UPDATE_METHODS[w["update_method"]]
    desc = nv.ModelDesc(word_repr_size=w["dw"], entity_repr_size=w["dd"], batch_normalization=w["bn"],
                        nonlinearity=nv.NONLINEARITIES[w["nonlinearity"]], clip_sigmoid=True, bias_negative_samples=w["bias_neg"])
    mk = lambda bs: nv.TrainConfig(batch_size=bs, window_size=w["n"], num_random_entities=w["z"], regularization_lambda=w["lam"],
                                   learning_rate=w["lr"], update_method=method, adam_mode=mode)
    f, labels = make_batches(w, Bg, 4321, 1)[0]          # identical on every rank
    fw, iw = np.ones((Bg, w["n"]), np.float32), np.ones(Bg, np.float32)
    pm = nv.Model(w["V"], w["D"], desc, mk(Bg // world), device=local_rank, gemm_mode=args.gemm_mode)
    prng = nv.RNG(1)
    pm.initialize(prng)
    sampler_state = prng.state
    ids = pm.generate_labels(labels, nv.RNG(sampler_state))      # uniform negatives: the reference's own generator
    sharding.init_model_comm(pm, dist, rank, world, peer_exchange=not args.no_peer)
    sf, sfw, sl, sw, sids = sharding.shard_batch(f, fw, labels, iw, ids, rank, world)
    names = ["grad_transform", "grad_bias"] + (["bn_mean", "bn_invstd"] if w["bn"] else [])

    def one_step(model, batch, entity_ids):
        res = model.compute_cost(batch, entity_ids=entity_ids)
        out = {"loss": res.get_cost()}
        model.compute_gradients(res)
        for nm in names:
            out[nm] = model.get_tensor(nm)
        model.update(None, w["lr"], res.scaled_regularization_lambda())
        out["T"], out["b"] = model.get_tensor(nv.TRANSFORM), model.get_tensor(nv.BIAS)
        return out

    got = one_step(pm, nv.Batch(Bg // world, w["n"]).fill(sf, sl, sfw, sw), sids)
    peer_err = pm.comm_peer_status()[1]
    pm.close()

    def fused_step(model, batch, entity_ids):
        """the one-call step bench.py times (N > 1: the critical-path reductions inside compute kernels over NVLink inboxes)"""
        model.train_step(batch, entity_ids, w["lr"])
        return {"fused_loss": model.last_cost(), "fused_T": model.get_tensor(nv.TRANSFORM), "fused_b": model.get_tensor(nv.BIAS)}

    pm = nv.Model(w["V"], w["D"], desc, mk(Bg // world), device=local_rank, gemm_mode=args.gemm_mode)
    pm.initialize(nv.RNG(1))
    sharding.init_model_comm(pm, dist, rank, world, peer_exchange=not args.no_peer)
    got.update(fused_step(pm, nv.Batch(Bg // world, w["n"]).fill(sf, sl, sfw, sw), sids))
    peer_err = max(peer_err, pm.comm_peer_status()[1])
    pm.close()
    report = None
    if rank == 0:
        tol = 2e-4 if args.gemm_mode != 1 else 2e-2
        um = nv.Model(w["V"], w["D"], desc, mk(Bg), device=local_rank, gemm_mode=args.gemm_mode)
        um.initialize(nv.RNG(1))
        want = one_step(um, nv.Batch(Bg, w["n"]).fill(f, labels, fw, iw), ids)
        um.close()
        um = nv.Model(w["V"], w["D"], desc, mk(Bg), device=local_rank, gemm_mode=args.gemm_mode)
        um.initialize(nv.RNG(1))
        want.update(fused_step(um, nv.Batch(Bg, w["n"]).fill(f, labels, fw, iw), ids))
        um.close()
        vs_un = {k: (abs(got[k] - want[k]) / abs(want[k]) if k.endswith("loss") else rel_err(got[k], want[k])) for k in got}
        report = {"global_batch": Bg, "ranks": world, "tolerance": tol, "vs_unsharded": vs_un,
                  "what": "one sharded step vs one GPU on the concatenated batch and vs oracle/_ref (reference, float32): "
                          "loss relative; tensors max|a-b| / max|b|; T and b after the update; fused_* = the same step through the "
                          "one-call API the bench times"}
        worst = max(vs_un.values())
        try:
            from oracle import ref_binding as R
            if R.available(np.float32) and not args.no_ref_check:
                um_ = {"sgd": (R.SGD, 0), "adagrad": (R.ADAGRAD, 0), "sparse_adam": (R.ADAM, R.SPARSE),
                       "dense_adam": (R.ADAM, R.DENSE_UPDATE), "full_adam": (R.ADAM, R.DENSE_UPDATE_DENSE_VARIANCE)}[w["update_method"]]
                rm = R.Model(w["V"], w["D"], w["dw"], w["dd"], batch_size=Bg, window_size=w["n"], num_random_entities=w["z"],
                             nonlinearity=R.HARD_TANH if w["nonlinearity"] == "hard_tanh" else R.TANH, batch_normalization=w["bn"],
                             clip_sigmoid=True, bias_negative_samples=w["bias_neg"], update_method=um_[0], adam_mode=um_[1],
                             regularization_lambda=w["lam"], seed=1, dtype=np.float32)
                assert rm.rng_state == sampler_state, "reference and library consumed the engine differently in initialize"
                rm.forward(rm.new_batch().fill(f, labels, fw, iw))
                ids_equal = bool((rm.entity_ids() == ids).all())
                ref = {"loss": rm.get_cost()}
                rm.compute_gradients()
                ref["grad_transform"], ref["grad_bias"] = rm.get("grad_transform"), rm.get("grad_bias")
                rm.update(w["lr"], rm.scaled_lambda())
                ref["T"], ref["b"] = rm.get("transform"), rm.get("bias")
                ref["fused_loss"], ref["fused_T"], ref["fused_b"] = ref["loss"], ref["T"], ref["b"]   # same step, one call
                vs_ref = {k: (abs(got[k] - ref[k]) / abs(ref[k]) if k.endswith("loss") else rel_err(got[k], ref[k])) for k in ref}
                vs_ref["sampled_ids_bit_exact"] = ids_equal
                report["vs_reference"] = vs_ref
                worst = max(worst, max(v for k, v in vs_ref.items() if k != "sampled_ids_bit_exact"))
                if not ids_equal:
                    worst = float("inf")
                del rm
            else:
                report["vs_reference"] = "oracle/_ref not built" if not R.available(np.float32) else "skipped (--no_ref_check)"
        except Exception as e:   # the checker must not take the bench line down
            report["vs_reference"] = "failed: %r" % (e,)
            worst = float("inf")
        report["max_rel_err"] = worst
        report["peer_exchange_error"] = peer_err
        report["ok"] = bool(worst <= tol and peer_err == 0)
    return report


def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import cunvsm_b200 as nv
    from cunvsm_b200 import sharding

    cores_per_rank = pin_rank_to_cores(local_rank, world)
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    B = w["B"]  # per GPU in the headline numbers (weak scaling); the strong block below shards w["B"] itself
    method, mode = nv.UPDATE_METHODS[w["update_method"]]
    desc = nv.ModelDesc(word_repr_size=w["dw"], entity_repr_size=w["dd"], batch_normalization=w["bn"],
                        nonlinearity=nv.NONLINEARITIES[w["nonlinearity"]], clip_sigmoid=True,
                        bias_negative_samples=w["bias_neg"])
    tc = nv.TrainConfig(batch_size=B, window_size=w["n"], num_random_entities=w["z"],
                        regularization_lambda=w["lam"], learning_rate=w["lr"], update_method=method, adam_mode=mode)
    stream = torch.cuda.Stream(device=local_rank, priority=env_int("NVSM_MAIN_PRIO", 0))   # (-1: above the library's side streams)
    sparse_mode = 1 if args.sparse_sync == "allgather" else 0

    def make_model(gemm_mode):
        m = nv.Model(w["V"], w["D"], desc, tc, device=local_rank, gemm_mode=gemm_mode, num_batch_slots=NUM_BATCHES)
        m.set_stream(stream.cuda_stream)
        r = nv.RNG(1)
        m.initialize(r)   # identical on every rank: same seed, same engine
        sharding.init_model_comm(m, dist, rank, world, sparse_mode=sparse_mode, peer_exchange=not args.no_peer)
        if w.get("neg_zipf", 0.0) > 0.0:
            # skewed negatives (configs[4]): inverse-CDF generator at the reference's LabelGenerator plug point; the
            # host loop (pre-sampled ids of `value`) and the device sampler (`e2e`) draw from the same distribution
            m.set_negative_distribution(nv.zipf_cdf(w["D"], w["neg_zipf"]))
        return m, r

    model, rng = make_model(args.gemm_mode)
    srng = nv.RNG(rng.state + rank)

    def make_staged(m, rows, seed):
        """Synthetic batches of `rows` n-grams for this rank, staged in the model's device slots."""
        raw = make_batches(w, rows, seed + rank, NUM_BATCHES)
        bs, ids_keep = [], []
        for k, (f, labels) in enumerate(raw):
            b = nv.Batch(rows, w["n"]).fill(f, labels)
            ids = m.generate_labels(labels, srng)
            m.stage_batch(k, b, ids)
            bs.append(b); ids_keep.append(ids)
        return bs, raw, ids_keep

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)

    def timed(m, fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = m.kernel_launches()
        t0 = time.perf_counter()
        e0.record(stream)
        for it in range(steps):
            fn(it)
        e1.record(stream)
        e1.synchronize()
        sampler.window(t0, time.perf_counter())
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            ms = sharding.max_over_ranks(dist, ms, device="cuda")
        return ms, m.kernel_launches() - l0

    lr = w["lr"]
    breakdown = []

    def measure(m, batches, steps, warmup):
        """(device-resident ms, launches, host-fed ms) of `steps` steps on the batches staged in / fed to `m`."""
        staged = lambda it: m.train_step_staged(it % NUM_BATCHES, lr)

        def host_fed(it):
            # the reference-facing call on HOST buffers: features / weights / positive labels go up every step,
            # the z negatives per instance are drawn on the device (bit-exact with the reference's host sampler)
            m.step_sampled(batches[it % NUM_BATCHES], lr)
            if it > 0:
                m.last_cost(1)  # loss of the previous step: D2H read every step, one step lagged
        for it in range(warmup):
            staged(it)
        ms, launches = timed(m, staged, steps)
        cost = m.last_cost()
        m.sampler_seed(srng)
        for it in range(max(3, min(warmup, 5))):
            host_fed(it)
        ms_e2e, _ = timed(m, host_fed, steps)
        m.last_cost()
        if args.e2e_breakdown:
            # where the host-fed step differs from the device-resident one: host time inside the call (enqueue cost),
            # the per-step loss read, and the H2D + sampling work itself
            host_s = [0.0]

            def host_fed_timed(it):
                t0 = time.perf_counter()
                m.step_sampled(batches[it % NUM_BATCHES], lr)
                host_s[0] += time.perf_counter() - t0
                if it > 0:
                    m.last_cost(1)
            ms_a, _ = timed(m, host_fed_timed, steps)
            ms_b, _ = timed(m, lambda it: m.step_sampled(batches[it % NUM_BATCHES], lr), steps)   # no per-step loss read
            m.last_cost()
            stage_s = [0.0]

            def staged_timed(it):
                t0 = time.perf_counter()
                m.train_step_staged(it % NUM_BATCHES, lr)
                stage_s[0] += time.perf_counter() - t0
            ms_c, _ = timed(m, staged_timed, steps)
            breakdown.append({"rows": batches[0].num_instances_, "e2e_ms": ms_a / steps, "e2e_no_loss_read_ms": ms_b / steps,
                              "staged_ms": ms_c / steps, "host_us_in_step_sampled": 1e6 * host_s[0] / steps,
                              "host_us_in_train_step_staged": 1e6 * stage_s[0] / steps})
        return ms, launches, ms_e2e, cost

    batches, raw, ids_keep = make_staged(model, B, 1234)
    if rank == 0:
        sampler.lead_s = 0.0
        sampler.start()
        time.sleep(0.05)
    t_warm = time.perf_counter()
    ms, launches, ms_e2e, final_cost = measure(model, batches, args.steps, args.warmup)
    sampler.lead_s = (min(a for a, _ in sampler.windows) - t_warm) if sampler.windows else 0.0
    clocks = sampler.stop() if rank == 0 else None

    # per-phase device time (CUDA events around every phase on the model's stream)
    model.set_profiling(True)
    model.reset_phase_ms()
    prof_steps = min(args.steps, 20)
    for it in range(prof_steps):
        model.train_step_staged(it % NUM_BATCHES, lr)
    phases = {k: v / prof_steps for k, v in model.phase_ms().items()}
    model.set_profiling(False)

    if args.timeline and rank == 0:
        # overlapped timeline of fused steps: every phase's (start, end) on whatever stream it ran (nvsm_set_profiling 2)
        for it in range(3):
            model.train_step_staged(it % NUM_BATCHES, lr)
        model.synchronize()
        model.set_profiling(2)
        for it in range(3):
            model.train_step_staged(it % NUM_BATCHES, lr)
        tl = model.timeline()
        model.set_profiling(False)
        # the same for the host-fed step (h2d = this batch's copy-stream work: H2D copies, id validation, device sampling)
        for it in range(4):
            model.step_sampled(batches[it % NUM_BATCHES], lr)
        model.synchronize()
        model.set_profiling(2)
        for it in range(4):
            model.step_sampled(batches[it % NUM_BATCHES], lr)
        tl_e2e = model.timeline()
        model.set_profiling(False)
        with open(args.timeline, "w") as f:
            for title, rows in (("3 fused steps on device-resident batches", tl), ("4 host-fed steps (nvsm_step_sampled)", tl_e2e)):
                f.write("# timeline of %s (ms since the first launch; streams overlap)\n\n| phase | start | end | us |\n|---|---|---|---|\n" % title)
                for name, a, b in sorted(rows, key=lambda x: x[1]):
                    f.write("| %s | %.4f | %.4f | %.1f |\n" % (name, a, b, 1e3 * (b - a)))
                f.write("\n")

    # strong scaling (BASELINE configs[3] as written): the SAME global batch w["B"] sharded over the ranks
    strong = None
    if world > 1 and args.scaling in ("both", "strong") and w["B"] % world == 0:
        Bs = w["B"] // world
        sbatches, _, _ = make_staged(model, Bs, 5678)
        s_ms, _, s_ms_e2e, _ = measure(model, sbatches, args.steps, max(3, args.warmup))
        strong = {"global_batch": w["B"], "per_gpu_batch": Bs, "value": w["B"] * args.steps / (s_ms * 1e-3),
                  "ms_per_step": s_ms / args.steps,
                  "e2e": {"value": w["B"] * args.steps / (s_ms_e2e * 1e-3), "ms_per_step": s_ms_e2e / args.steps},
                  "unit": "n-grams/s",
                  "note": "speed-up = value / the N=1 line's value (the driver computes it); what does not shrink with "
                          "N on replicated tables: the dense optimiser pass over every table row and the bucket build "
                          "over all table rows (see DESIGN.md section 5)"}
        model.set_profiling(True)
        model.reset_phase_ms()
        for it in range(prof_steps):
            model.train_step_staged(it % NUM_BATCHES, lr)
        strong["phase_ms"] = {k: round(v / prof_steps, 4) for k, v in model.phase_ms().items()}
        model.set_profiling(False)

    # roofline denominators measured in place (rank 0): L2-resident gather and streaming copy
    probes = None
    if rank == 0 and not args.no_probes:
        table_bytes = min(4 * w["V"] * w["dw"], 64 << 20)
        rowf = (w["dw"] + 3) // 4 * 4
        gathers = {u: model.bench_memory(kind, table_bytes, row_floats=rowf, rows_per_item=w["n"], items=B, iters=20)
                   for kind, u in ((4, 1), (0, 2), (3, 4))}
        best_u = max(gathers, key=gathers.get)
        probes = {"l2_gather_gbs": gathers[best_u],
                  "l2_gather_probe": "warp-per-item gather of n=%d rows x %d floats from a %d MB (L2-resident) table, %d rows in "
                                     "flight per warp (best of 1 / 2 / 4: %s GB/s)" % (
                                         w["n"], rowf, table_bytes >> 20, best_u, " / ".join("%.0f" % gathers[u] for u in (1, 2, 4))),
                  "l2_read_gbs": model.bench_memory(2, 32 << 20, items=32, iters=5),
                  "stream_copy_gbs": model.bench_memory(1, 1 << 30, iters=10)}

    peer_ok = world > 1 and model.comm_peer_status()[0]
    if world > 1 and model.comm_peer_status()[1]:
        raise RuntimeError("NVLink peer exchange timed out waiting for a peer")

    alt = None
    if args.gemm_mode == 2 and not args.no_alt:
        # the same timed loop with single-pass TF32 GEMMs (looser parity, see tests/test_gpu_loss_curve.py)
        model.close()
        model, _ = make_model(1)
        make_staged(model, B, 1234)
        for it in range(args.warmup):
            model.train_step_staged(it % NUM_BATCHES, lr)
        ms_alt, _ = timed(model, lambda it: model.train_step_staged(it % NUM_BATCHES, lr), args.steps)
        alt = {"gemm": "tf32_tcgen05", "value": B * world * args.steps / (ms_alt * 1e-3), "ms_per_step": ms_alt / args.steps}
    model.close()

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = parity_check(args, w, nv, sharding, dist, rank, world, local_rank)

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        f0, l0 = raw[0]
        uniq_w, uniq_e = int(np.unique(f0).size), int(np.unique(ids_keep[0]).size)
        comp = compulsory_bytes_per_launch(w, B, args.gemm_mode, unique_words=uniq_w, unique_entities=uniq_e, world=world)
        alg = algorithmic_bytes_per_ngram(w)
        cand = {k: phases.get(k, 0.0) for k in comp}
        dom = max(cand, key=cand.get)
        dom_ms = cand[dom]
        achieved = comp[dom] / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        traffic, tr = None, {}
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(args.workload, {}).get(dom)
        except Exception:
            pass
        l2_peak = max(probes["l2_gather_gbs"], probes["l2_read_gbs"]) if probes else None
        l2_bytes = tr.get(args.workload + "_l2", {})   # SM <-> L2 bytes per launch (ncu lts__t_sectors_srcunit_tex x 32 B)
        per_kernel = {}
        for k in comp:
            t = phases.get(k, 0.0) * 1e-3
            if t <= 0:
                continue
            l2b = l2_bytes.get(k)
            per_kernel[k] = {"ms": round(phases[k], 4), "hbm_compulsory_gbs": round(comp[k] / t / 1e9, 1),
                             "hbm_frac": round(comp[k] / t / 1e9 / peak, 3),
                             "l2_gbs": round(l2b / t / 1e9, 1) if l2b else None,
                             "l2_frac": round(l2b / t / 1e9 / l2_peak, 3) if (l2b and l2_peak) else None,
                             "per_reference_gbs": round(alg[k] * B / t / 1e9, 1)}
        dom_l2 = l2_bytes.get(dom)
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    # achieved = COMPULSORY HBM bytes of the kernel as built (distinct table rows once, streamed per-step
                    # tensors, one read-modify-write of every optimiser-state row; compulsory_bytes_per_launch) / its
                    # CUDA-event time. traffic = ncu dram bytes of the same launch (profiles/traffic.json).
                    "compulsory_bytes_per_launch": comp[dom], "kernel_ms": dom_ms,
                    "dram_gbs": (traffic / (dom_ms * 1e-3) / 1e9) if (traffic and dom_ms > 0) else None,
                    "dram_frac": (traffic / (dom_ms * 1e-3) / 1e9 / peak) if (traffic and dom_ms > 0) else None,
                    "excess_traffic": (traffic / comp[dom]) if traffic else None,
                    # second ceiling: the gather-type kernels re-read every table row ~10x per batch out of L2. achieved =
                    # SM <-> L2 bytes of the launch (ncu lts__t_sectors_srcunit_tex, profiles/traffic.json) / CUDA-event time;
                    # peak = the best of an L2-resident row gather and an L2-resident streaming read measured on this GPU now
                    "l2": {"achieved": (dom_l2 / (dom_ms * 1e-3) / 1e9) if (dom_l2 and dom_ms > 0) else None, "peak": l2_peak,
                           "unit": "GB/s", "frac": (dom_l2 / (dom_ms * 1e-3) / 1e9 / l2_peak) if (dom_l2 and l2_peak and dom_ms > 0) else None,
                           "bytes_per_launch": dom_l2, "gather_probe_gbs": probes["l2_gather_gbs"] if probes else None,
                           "read_probe_gbs": probes["l2_read_gbs"] if probes else None,
                           "peak_source": (probes["l2_gather_probe"] + "; L2-resident streaming read of a 32 MB buffer") if probes else None},
                    "stream_copy_gbs_here": probes["stream_copy_gbs"] if probes else None,
                    "per_kernel": per_kernel,
                    "step_compulsory_hbm_gbs": sum(comp.values()) / (ms / args.steps * 1e-3) / 1e9,
                    "phase_ms": {k: round(v, 4) for k, v in phases.items()}}
        # word ids + positive labels (8-byte ids, the reference's `long`); the synthetic workload uses uniform word and
        # instance weights, which the C ABI takes as NULL and fills on the device: no bytes for them
        h2d = int(B * w["n"] * 8 + B * 8)
        ngrams = B * world * args.steps
        cpu = None
        if not args.no_cpu_baseline and world == 1:
            # bounded sample: whole batches of the same workload for ~args.cpu_seconds of host time (probe step first)
            sample_B = min(args.cpu_sample, w["B"]) if args.cpu_sample > 0 else w["B"]
            probe = cpu_port_run(w, 1, 1, sample_B)
            cpu_steps = int(min(40, max(2, round(args.cpu_seconds / max(probe["ms_per_step"] * 1e-3, 1e-6)))))
            r = cpu_port_run(w, cpu_steps, 0, sample_B)
            cpu = {"value": r["value"], "unit": "n-grams/s", "cores": r["cores"], "kind": "port",
                   "sample": "%d n-grams/step x %d steps (%.1f s) of the same workload (full-size tables), float32 oracle "
                             "port, host sampler + forward + backward + update, OpenMP over %d threads, %s build" % (
                                 sample_B, cpu_steps, r["ms_per_step"] * cpu_steps * 1e-3, r["cores"],
                                 "-march=native" if r["native"] else "portable")}
        line = {
            "metric": "n-grams/sec", "value": ngrams / (ms * 1e-3), "unit": "n-grams/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["name"], "update_method": w["update_method"], "per_gpu_batch": B,
                       "global_batch": B * world, "gemm": ["fp32_simt", "tf32_tcgen05", "3xtf32_tcgen05"][args.gemm_mode],
                       "l2": "working set (tables + moments + per-step tensors, > 1 GB) exceeds the 126 MB L2; "
                             "%d distinct batches cycled" % NUM_BATCHES,
                       "sparse_tables": ("single GPU" if world == 1 else "replicated, per-rank local updates" if args.sparse_sync == "local"
                                         else "replicated, rows all-gathered: every replica applies the global update"),
                       "negatives": ("value: pre-sampled (bit-exact host sampler) and staged with the batch; e2e: drawn "
                                     "inside the timed step by the bit-exact device sampler"
                                     + ("; Zipf(%g) over the entity ids (inverse-CDF generator)" % w["neg_zipf"]
                                        if w.get("neg_zipf", 0.0) > 0.0 else "; uniform (the reference's generator)")),
                       "word_ids": "Zipf(%g)" % w["word_zipf"] if w.get("word_zipf", 0.0) > 0.0 else "uniform",
                       "weights": "uniform word / instance weights (the reference's default weighting): passed as NULL over the "
                                  "C ABI, ones filled on the device, not part of the H2D bytes",
                       "host_cores_per_rank": cores_per_rank},
            "e2e": {"value": ngrams / (ms_e2e * 1e-3), "unit": "n-grams/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps,
                    "d2h": "the 8-byte loss sum of every step, written by the score kernel's last block into pinned host memory "
                           "(N > 1 without peer memory: cudaMemcpyAsync) and read by the host one step lagged"},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "final_cost": final_cost, "alt_single_pass_tf32": alt,
            "collectives": (None if world == 1 else
                            {"small_reductions": ("NVLink peer exchange fused into the compute kernels (batch-norm sums: "
                                                   "col_stats_reduce_finalize_kernel<true>; backward sums + loss: the score kernel's last "
                                                   "block; peer_allreduce.cuh)" + ("" if not os.environ.get("NVSM_NO_FUSED_XCHG") else
                                                                                    " -- stand-alone one-block launches (NVSM_NO_FUSED_XCHG)"))
                              if peer_ok else "ncclAllReduce",
                             "grad_transform": ("NVLink inboxes (gt_reduce_push_kernel -> transform_update_kernel)"
                                                if peer_ok and os.environ.get("NVSM_FUSED_GT") == "1" else
                                                "ncclAllReduce on a side stream under grad_phrase + word update")}),
        }
        if world > 1:
            line["strong"] = strong
            line["parity_check"] = parity
        if breakdown:
            line["e2e_breakdown"] = breakdown
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=sorted(WORKLOADS))
    ap.add_argument("--update_method", default=None)
    ap.add_argument("--gemm_mode", type=int, default=2, help="0 fp32 SIMT, 1 tf32 tcgen05, 2 3xtf32 tcgen05 (default: fp32-level parity)")
    ap.add_argument("--cpu_sample", type=int, default=0, help="n-grams per step of the CPU sample (0 = the workload's whole batch)")
    ap.add_argument("--cpu_seconds", type=float, default=10.0, help="host time budget of the cpu_baseline sample")
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--reference_kind", default="auto", choices=["auto", "cpu"],
                    help="--impl reference: auto = the reference's own CUDA step (oracle/_ref) when built, else the CPU port")
    ap.add_argument("--sparse_sync", default="local", choices=["local", "allgather"],
                    help="N>1: local = per-rank sparse updates (north-star prescription, default); allgather = exact "
                         "single-GPU trajectory (every replica applies all rows' updates)")
    ap.add_argument("--no_peer", action="store_true", help="N>1: small reductions through ncclAllReduce instead of the NVLink peer exchange")
    ap.add_argument("--no_alt", action="store_true", help="skip the extra single-pass TF32 measurement")
    ap.add_argument("--scaling", default="both", choices=["weak", "strong", "both"],
                    help="N>1: `value` is always the weak-scaling number (per-GPU batch fixed); both / strong add a "
                         "`strong` block with the workload's batch itself sharded over the ranks (BASELINE configs[3])")
    ap.add_argument("--no_parity_check", action="store_true", help="N>1: skip the sharded-vs-unsharded-vs-reference step check")
    ap.add_argument("--no_ref_check", action="store_true", help="N>1 parity check without the oracle/_ref leg")
    ap.add_argument("--e2e_breakdown", action="store_true", help="extra timed loops that split the host-fed step's overhead (rank 0's view)")
    ap.add_argument("--timeline", default=None, help="write the overlapped per-phase timeline of three fused steps to this markdown file")
    ap.add_argument("--no_probes", action="store_true", help="skip the L2-gather / stream-copy roofline probes")

    ap.add_argument("--zipf_words", type=float, default=0.0, help="word ids ~ Zipf(s) instead of uniform (C3 gather/scatter sweep)")
    ap.add_argument("--zipf_negatives", type=float, default=None, help="negatives ~ Zipf(s) over the entity ids (C5 default 1.0; 0 = uniform)")
    args = ap.parse_args()
    w = dict(WORKLOADS[args.workload])
    if args.update_method:
        w["update_method"] = args.update_method
    if args.zipf_words > 0.0:
        w["word_zipf"] = args.zipf_words
        w["name"] += ", Zipf(%g) word ids" % args.zipf_words
    if args.zipf_negatives is not None:
        w["neg_zipf"] = args.zipf_negatives
    rank, world, local_rank = env_int("RANK", 0), env_int("WORLD_SIZE", 1), env_int("LOCAL_RANK", 0)
    if args.impl == "reference":
        run_reference(args, w, rank)
        return
    if world == 1 and args.gpus > 1:
        # launched without torchrun: re-exec under torch.distributed.run
        cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(args.gpus),
               "--master-addr", "127.0.0.1", "--master-port", "29533", os.path.abspath(__file__)] + sys.argv[1:]
        sys.exit(subprocess.call(cmd))
    run_ours(args, w, rank, world, local_rank)


if __name__ == "__main__":
    main()
