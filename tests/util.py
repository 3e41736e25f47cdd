"""Shared helpers for the parity tests: seeded synthetic batches and oracle/CUDA twins."""
import numpy as np

import cunvsm_b200 as nv
from oracle import binding as O


def make_batch(rng, B, n, V, D, z, weighted=True):
    features = rng.integers(0, V, size=(B, n), dtype=np.int64)
    fw = rng.uniform(0.0, 2.0, size=(B, n)).astype(np.float32) if weighted else np.ones((B, n), np.float32)
    labels = rng.integers(0, D, size=B, dtype=np.int64)
    w = rng.uniform(0.0, 2.0, size=B).astype(np.float32) if weighted else np.ones(B, np.float32)
    return features, fw, labels, w


def twin_models(V, D, dw, dd, *, n, z, B, nonlinearity=nv.TANH, bn=False, clip=True, bias_neg=False,
                method=nv.SGD, adam_mode=nv.SPARSE, lam=0.01, seed=7, oracle_dtype=np.float32, gemm_mode=nv.GEMM_FP32,
                num_batch_slots=1):
    """Build the CUDA model and the oracle with identical (Glorot, seeded) parameters."""
    desc = nv.ModelDesc(word_repr_size=dw, entity_repr_size=dd, batch_normalization=bn, nonlinearity=nonlinearity,
                        clip_sigmoid=clip, bias_negative_samples=bias_neg)
    tc = nv.TrainConfig(batch_size=B, window_size=n, num_random_entities=z, regularization_lambda=lam,
                        update_method=method, adam_mode=adam_mode)
    gm = nv.Model(V, D, desc, tc, gemm_mode=gemm_mode, num_batch_slots=num_batch_slots)
    rng = nv.RNG(seed)
    gm.initialize(rng)
    om = O.Model(V, D, dw, dd, nonlinearity=nonlinearity, batch_normalization=bn, clip_sigmoid=clip,
                 bias_negative_samples=bias_neg, update_method=method, adam_mode=adam_mode,
                 num_random_entities=z, regularization_lambda=float(np.float32(lam)), dtype=oracle_dtype)
    om.set("W", gm.get_tensor(nv.WORD_REPRS))
    om.set("E", gm.get_tensor(nv.ENTITY_REPRS))
    om.set("T", gm.get_tensor(nv.TRANSFORM))
    om.set("b", gm.get_tensor(nv.BIAS))
    return gm, om, rng


def assert_close(actual, expected, rtol, atol_scale=1e-6, what=""):
    actual = np.asarray(actual, dtype=np.float64).ravel()
    expected = np.asarray(expected, dtype=np.float64).ravel()
    assert actual.shape == expected.shape, (what, actual.shape, expected.shape)
    scale = max(np.abs(expected).max(), 1e-30)
    np.testing.assert_allclose(actual, expected, rtol=rtol, atol=atol_scale * scale, err_msg=what)
