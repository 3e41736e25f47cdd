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
    """gather W and E rows once, read-modify-write the same rows once, ids as
    delivered (8 B), weights (4 B). Dense optimiser passes are amortised separately."""
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


def env_int(name, default):
    try:
        return int(os.environ.get(name, default))
    except ValueError:
        return default


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.proc, self.lines = gpu_index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "20"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1])); mx.append(float(parts[2]))
            except ValueError:
                continue
            for nm, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons),
                "samples": len(sm)}


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
                            "reference_build": "unmodified /root/reference/cpp/*.cu, nvcc -O3 -use_fast_math float32 NDEBUG, "
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
    sample_B = min(args.cpu_sample, w["B"])
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
def run_ours(args, w, rank, world, local_rank):
    import torch
    import torch.distributed as dist

    import cunvsm_b200 as nv

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", local_rank))

    B = w["B"]  # per GPU (weak scaling)
    method, mode = nv.UPDATE_METHODS[w["update_method"]]
    desc = nv.ModelDesc(word_repr_size=w["dw"], entity_repr_size=w["dd"], batch_normalization=w["bn"],
                        nonlinearity=nv.NONLINEARITIES[w["nonlinearity"]], clip_sigmoid=True,
                        bias_negative_samples=w["bias_neg"])
    tc = nv.TrainConfig(batch_size=B, window_size=w["n"], num_random_entities=w["z"],
                        regularization_lambda=w["lam"], learning_rate=w["lr"], update_method=method, adam_mode=mode)
    model = nv.Model(w["V"], w["D"], desc, tc, device=local_rank, gemm_mode=args.gemm_mode,
                     num_batch_slots=NUM_BATCHES)
    stream = torch.cuda.Stream(device=local_rank)
    model.set_stream(stream.cuda_stream)
    rng = nv.RNG(1)
    model.initialize(rng)   # identical on every rank: same seed, same engine
    from cunvsm_b200 import sharding
    sharding.init_model_comm(model, dist, rank, world, sparse_mode=1 if args.sparse_sync == "allgather" else 0,
                             peer_exchange=not args.no_peer)

    if w.get("neg_zipf", 0.0) > 0.0:
        # skewed negatives (configs[4]): inverse-CDF generator at the reference's LabelGenerator plug point; the host
        # loop (pre-sampled ids of `value`) and the device sampler (`e2e`) draw from the same distribution
        model.set_negative_distribution(nv.zipf_cdf(w["D"], w["neg_zipf"]))
    # synthetic batches: every rank owns its own shard of n-gram rows
    raw = make_batches(w, B, 1234 + rank, NUM_BATCHES)
    batches, ids_list = [], []
    srng = nv.RNG(rng.state + rank)
    for f, labels in raw:
        b = nv.Batch(B, w["n"]).fill(f, labels)
        ids = model.generate_labels(labels, srng)
        pinned_ids = torch.from_numpy(ids).pin_memory()
        batches.append(b); ids_list.append(pinned_ids)
    ids_np = [t.numpy() for t in ids_list]
    for s in range(NUM_BATCHES):
        model.stage_batch(s, batches[s], ids_np[s])

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = model.kernel_launches()
        e0.record(stream)
        for it in range(steps):
            fn(it)
        e1.record(stream)
        e1.synchronize()
        barrier()
        ms = e0.elapsed_time(e1)
        if world > 1:
            ms = sharding.max_over_ranks(dist, ms, device="cuda")
        return ms, model.kernel_launches() - l0

    lr = w["lr"]
    staged = lambda it: model.train_step_staged(it % NUM_BATCHES, lr)

    def host_fed(it):
        # the reference-facing call on HOST buffers: features / weights / positive labels go up every step,
        # the z negatives per instance are drawn on the device (bit-exact with the reference's host sampler)
        k = it % NUM_BATCHES
        model.step_sampled(batches[k], lr)
        if it > 0:
            model.last_cost(1)  # loss of the previous step: D2H read every step, one step lagged

    for it in range(args.warmup):
        staged(it)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms, launches = timed(staged, args.steps)
    clocks = sampler.stop() if rank == 0 else None
    final_cost = model.last_cost()

    model.sampler_seed(srng)
    for it in range(max(3, min(args.warmup, 5))):
        host_fed(it)
    ms_e2e, _ = timed(host_fed, args.steps)
    model.last_cost()

    # per-phase device time (CUDA events around every phase on the model's stream)
    model.set_profiling(True)
    model.reset_phase_ms()
    prof_steps = min(args.steps, 20)
    for it in range(prof_steps):
        staged(it)
    phases = {k: v / prof_steps for k, v in model.phase_ms().items()}
    model.set_profiling(False)

    alt = None
    if args.gemm_mode == 2 and not args.no_alt:
        # the same timed loop with single-pass TF32 GEMMs (looser parity, see tests/test_gpu_loss_curve.py)
        model.close()
        model = nv.Model(w["V"], w["D"], desc, tc, device=local_rank, gemm_mode=1, num_batch_slots=NUM_BATCHES)
        model.set_stream(stream.cuda_stream)
        model.initialize(nv.RNG(1))
        sharding.init_model_comm(model, dist, rank, world, sparse_mode=1 if args.sparse_sync == "allgather" else 0,
                             peer_exchange=not args.no_peer)
        for s_ in range(NUM_BATCHES):
            model.stage_batch(s_, batches[s_], ids_np[s_])
        staged = lambda it: model.train_step_staged(it % NUM_BATCHES, lr)
        for it in range(args.warmup):
            staged(it)
        ms_alt, _ = timed(staged, args.steps)
        alt = {"gemm": "tf32_tcgen05", "value": B * world * args.steps / (ms_alt * 1e-3), "ms_per_step": ms_alt / args.steps}

    peer_ok = world > 1 and model.comm_peer_status()[0]
    if world > 1 and model.comm_peer_status()[1]:
        raise RuntimeError("NVLink peer exchange timed out waiting for a peer")
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        alg = algorithmic_bytes_per_ngram(w)
        cand = {k: phases.get(k, 0.0) for k in alg}
        dom = max(cand, key=cand.get)
        dom_ms = cand[dom]
        achieved = alg[dom] * B / (dom_ms * 1e-3) / 1e9 if dom_ms > 0 else 0.0
        total_alg = sum(alg.values()) * B
        traffic = None
        try:   # dram__bytes_read.sum + dram__bytes_write.sum per launch, from the committed ncu --set full capture
            tr = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
            traffic = tr.get(args.workload, {}).get(dom)
        except Exception:
            pass
        roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                    # achieved counts ALGORITHMIC bytes (every referenced row gathered once and read-modify-written
                    # once); rows are referenced ~10x per step and the tables fit the 126 MB L2, so most of that is
                    # served by L2 and frac can exceed 1. dram_gbs is the ncu-measured DRAM traffic over the same time.
                    "dram_gbs": (traffic / (dom_ms * 1e-3) / 1e9) if (traffic and dom_ms > 0) else None,
                    "dram_frac": (traffic / (dom_ms * 1e-3) / 1e9 / peak) if (traffic and dom_ms > 0) else None,
                    "algorithmic_bytes_per_launch": alg[dom] * B, "kernel_ms": dom_ms,
                    "step_algorithmic_gbs": total_alg / (ms / args.steps * 1e-3) / 1e9,
                    "phase_ms": {k: round(v, 4) for k, v in phases.items()}}
        h2d = int(B * w["n"] * 8 + B * w["n"] * 4 + B * 8 + B * 4)
        ngrams = B * world * args.steps
        cpu = None
        if not args.no_cpu_baseline:
            r = cpu_port_run(w, 2, 1, min(args.cpu_sample, w["B"]))
            cpu = {"value": r["value"], "unit": "n-grams/s", "cores": r["cores"], "kind": "port",
                   "sample": "%d n-grams/step x 2 steps of the same workload (full-size tables), float32 oracle, "
                             "sampler+forward+backward+update, %s build" % (args.cpu_sample, "-march=native" if r["native"] else "portable")}
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
                       "word_ids": "Zipf(%g)" % w["word_zipf"] if w.get("word_zipf", 0.0) > 0.0 else "uniform"},
            "e2e": {"value": ngrams / (ms_e2e * 1e-3), "unit": "n-grams/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": 8, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu,
            "final_cost": final_cost, "alt_single_pass_tf32": alt,
            "collectives": (None if world == 1 else
                            {"small_reductions": "nvlink peer exchange (peer_allreduce.cuh)" if peer_ok else "ncclAllReduce",
                             "grad_transform": "ncclAllReduce on a side stream under grad_phrase + word update"}),
        }
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
    ap.add_argument("--cpu_sample", type=int, default=5120)
    ap.add_argument("--no_cpu_baseline", action="store_true")
    ap.add_argument("--reference_kind", default="auto", choices=["auto", "cpu"],
                    help="--impl reference: auto = the reference's own CUDA step (oracle/_ref) when built, else the CPU port")
    ap.add_argument("--sparse_sync", default="local", choices=["local", "allgather"],
                    help="N>1: local = per-rank sparse updates (north-star prescription, default); allgather = exact "
                         "single-GPU trajectory (every replica applies all rows' updates)")
    ap.add_argument("--no_peer", action="store_true", help="N>1: small reductions through ncclAllReduce instead of the NVLink peer exchange")
    ap.add_argument("--no_alt", action="store_true", help="skip the extra single-pass TF32 measurement")
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
