"""Host-side logic of the multi-GPU path: one process per GPU, the batch sharded by n-gram row.

The reference is single-GPU; this is the new data-parallel layer around the same step:
rank r owns rows [r*B/N, (r+1)*B/N) of every batch (and the matching slice of sampled entity
ids), parameters are replicated, and inside the step the library all-reduces the batch-norm
statistics, the backward column sums + loss, and grad_transform (NCCL, see nvsm_comm_init).
torch.distributed is only plumbing here: exchanging the NCCL unique id, barriers and the
max-over-ranks of timings.
"""
import numpy as np


def shard_range(num_instances, rank, world):
    """Rows of the global batch owned by `rank`. Shards must be equal (the library derives the
    global batch as local * world), so the batch size has to divide evenly."""
    if num_instances % world != 0:
        raise ValueError("batch of %d rows does not split evenly over %d ranks" % (num_instances, world))
    per = num_instances // world
    return rank * per, (rank + 1) * per


def shard_batch(features, feature_weights, labels, weights, entity_ids, rank, world):
    """Slice host arrays of a global batch ([B, n], [B, n], [B], [B], [B*(z+1)]) for one rank."""
    B = len(labels)
    lo, hi = shard_range(B, rank, world)
    R = len(entity_ids) // B
    ids = np.asarray(entity_ids).reshape(B, R)
    return (np.asarray(features)[lo:hi], np.asarray(feature_weights)[lo:hi], np.asarray(labels)[lo:hi],
            np.asarray(weights)[lo:hi], ids[lo:hi].reshape(-1))


def broadcast_unique_id(dist, make_id, rank, src=0):
    """Rank `src` creates the 128-byte NCCL unique id, everyone receives it."""
    box = [make_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    uid = box[0]
    if not isinstance(uid, (bytes, bytearray)) or len(uid) != 128:
        raise RuntimeError("bad NCCL unique id")
    return bytes(uid)


def init_model_comm(model, dist, rank, world, sparse_mode=0, peer_exchange=True):
    """Attach an NCCL communicator to `model` (no-op for a single rank). sparse_mode: 0 = per-rank local
    updates of the replicated tables, 1 = all-gather the rows so every replica applies the global update."""
    if world <= 1:
        return
    from .model import comm_unique_id
    uid = broadcast_unique_id(dist, comm_unique_id, rank)
    model.comm_init(uid, world, rank)
    if sparse_mode:
        model.comm_set_sparse_mode(sparse_mode)
    if peer_exchange:
        init_peer_exchange(model, dist, world)


def init_peer_exchange(model, dist, world):
    """NVLink peer exchange for the small per-step reductions; every rank must take the same decision, so the
    outcome of the IPC export / import is agreed on before anybody uses it (NCCL stays the fallback)."""
    try:
        blob = model.comm_peer_export()
    except Exception:
        blob = None
    blobs = [None] * world
    dist.all_gather_object(blobs, blob)
    ok = all(isinstance(b, (bytes, bytearray)) and len(b) == 128 for b in blobs)
    if ok:
        try:
            model.comm_peer_import([bytes(b) for b in blobs])
        except Exception:
            ok = False
    flags = [None] * world
    dist.all_gather_object(flags, ok)
    if not all(flags):
        model.peer_ready_override_off()
    dist.barrier()
    return all(flags)


def max_over_ranks(dist, value, device=None):
    """Device-timed milliseconds -> the slowest rank's figure."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
