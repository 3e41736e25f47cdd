"""world_size-2 gloo tests (CPU) of the host-side multi-GPU logic in cunvsm_b200/sharding.py:
row sharding, unique-id exchange, max-over-ranks, and the reduction recipe the library applies
across ranks (sum of per-shard loss sums / global B, summed grad_transform / grad_bias, global
batch-norm statistics), checked with the CPU oracle as the per-shard calculator."""
import os
import socket
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, out):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from cunvsm_b200 import sharding
    from oracle import binding as O
    from tests.util import make_batch

    # unique-id plumbing (a fake 128-byte id: ncclGetUniqueId needs no GPU but is not the point here)
    uid = sharding.broadcast_unique_id(dist, lambda: bytes(range(128)), rank)
    assert uid == bytes(range(128))
    assert sharding.max_over_ranks(dist, 1.0 + rank) == float(world)
    with pytest.raises(ValueError):
        sharding.shard_range(7, rank, world)

    V, D, dw, dd, n, z, B = 50, 40, 8, 6, 3, 2, 64
    f, fw, labels, w = make_batch(np.random.default_rng(5), B, n, V, D, z)
    ids, _ = O.generate_labels(labels, z, D, 17)
    lo, hi = sharding.shard_range(B, rank, world)
    sf, sfw, sl, sw, sids = sharding.shard_batch(f, fw, labels, w, ids, rank, world)
    assert sf.shape == (B // world, n) and sids.size == (B // world) * (z + 1)
    assert (sids.reshape(-1, z + 1)[:, 0] == labels[lo:hi]).all()

    def model():
        m = O.Model(V, D, dw, dd, nonlinearity=O.TANH, clip_sigmoid=True, num_random_entities=z)
        m.initialize(3)
        return m

    full = model()
    cost_full = full.compute_cost(f, fw, ids, w, n)
    full.compute_gradients()
    # per-shard forward/backward, then the cross-rank reduction the CUDA library performs:
    # loss sums and dense gradients add up; 1/B uses the GLOBAL batch (mult carries 1/B_local here).
    part = model()
    cost_part = part.compute_cost(sf, sfw, sids, sw, n)
    part.compute_gradients()
    scale = (hi - lo) / B
    red = torch.tensor(np.concatenate([[cost_part * scale], part.get("gT") * scale, part.get("gb") * scale]))
    dist.all_reduce(red)
    red = red.numpy()
    np.testing.assert_allclose(red[0], cost_full, rtol=1e-12)
    np.testing.assert_allclose(red[1:1 + dw * dd], full.get("gT"), rtol=1e-10, atol=1e-14)
    np.testing.assert_allclose(red[1 + dw * dd:], full.get("gb"), rtol=1e-10, atol=1e-14)
    # rows of grad_phrase stay local: shard rows equal the matching rows of the full batch
    np.testing.assert_allclose(part.get("gP") * scale, full.get("gP").reshape(B, dw)[lo:hi].ravel(), rtol=1e-10, atol=1e-15)
    # global batch-norm statistics = all-reduced sums over shards
    Z = np.random.default_rng(rank).standard_normal((hi - lo, dd))
    allZ = [None] * world
    dist.all_gather_object(allZ, Z)
    s = torch.tensor(np.concatenate([Z.sum(0), (Z * Z).sum(0)])); dist.all_reduce(s); s = s.numpy()
    mean = s[:dd] / B; var = s[dd:] / B - mean ** 2
    cat = np.concatenate(allZ)
    np.testing.assert_allclose(mean, cat.mean(0), rtol=1e-12, atol=1e-14)
    np.testing.assert_allclose(var, cat.var(0), rtol=1e-10)
    out.put((rank, "ok"))
    dist.destroy_process_group()


def test_two_rank_sharding_logic_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(world))
    assert got == [(0, "ok"), (1, "ok")]
