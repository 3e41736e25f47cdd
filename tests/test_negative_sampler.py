"""The skewed (inverse-CDF) negative sampler of BASELINE configs[4] — the reference's LabelGenerator plug point
(include/cuNVSM/labels.h:7-18) — host loop vs the oracle restatement (CPU), device kernel vs both (GPU)."""
import ctypes

import numpy as np
import pytest

import cunvsm_b200 as nv
from cunvsm_b200 import _lib
from oracle import binding as O


def host_cdf_labels(labels, z, cdf, state):
    L = _lib.load()
    labels = np.ascontiguousarray(labels, dtype=np.int64)
    out = np.zeros(labels.size * (z + 1), dtype=np.int64)
    st = ctypes.c_ulong(state)
    _lib.check(L.nvsm_generate_labels_cdf(labels.ctypes.data_as(ctypes.POINTER(ctypes.c_long)), labels.size, z,
                                          cdf.ctypes.data_as(ctypes.POINTER(ctypes.c_double)), cdf.size, ctypes.byref(st),
                                          out.ctypes.data_as(ctypes.POINTER(ctypes.c_long))))
    return out, st.value


@pytest.mark.parametrize("D,z,B,s,seed", [(7, 3, 50, 1.0, 1), (1000, 4, 2000, 1.0, 10), (100000, 32, 512, 0.75, 5),
                                           (3, 1, 1, 2.0, 2147483646), (50, 0, 9, 1.0, 3)])
def test_host_cdf_sampler_matches_oracle(D, z, B, s, seed):
    cdf = nv.zipf_cdf(D, s)
    labels = np.random.default_rng(seed % 1000).integers(0, D, size=B)
    got, st = host_cdf_labels(labels, z, cdf, seed)
    exp, est = O.generate_labels_cdf(labels, z, cdf, seed)
    assert (got == exp).all() and st == est
    assert (got.reshape(B, z + 1)[:, 0] == labels).all()


def test_host_cdf_sampler_follows_the_distribution():
    D, z, B = 20, 50, 4000
    cdf = nv.zipf_cdf(D, 1.0)
    got, _ = host_cdf_labels(np.zeros(B, dtype=np.int64), z, cdf, 42)
    neg = got.reshape(B, z + 1)[:, 1:].ravel()
    freq = np.bincount(neg, minlength=D) / neg.size
    np.testing.assert_allclose(freq, np.diff(np.concatenate([[0.0], cdf])), atol=4e-3)


def test_host_cdf_sampler_rejects_bad_distributions():
    labels = np.zeros(4, dtype=np.int64)
    for bad in (np.array([0.5, 0.4, 1.0]), np.array([0.2, 0.9]), np.array([-0.1, 1.0]), np.array([0.3, np.nan, 1.0])):
        with pytest.raises(nv.NvsmError):
            host_cdf_labels(labels, 2, bad, 1)


@pytest.mark.gpu
@pytest.mark.parametrize("D,z,B,s,seed", [(7, 3, 50, 1.0, 1), (1000000, 32, 4096, 1.0, 77), (50000, 10, 51200, 1.0, 1),
                                           (200, 1, 1024, 0.5, 2147483646)])
def test_device_cdf_sampler_bit_exact(D, z, B, s, seed):
    """nvsm_step_sampled with a distribution installed draws the ids of the host loop; the engine state chains."""
    desc = nv.ModelDesc(word_repr_size=8, entity_repr_size=8)
    tc = nv.TrainConfig(batch_size=B, window_size=2, num_random_entities=z)
    m = nv.Model(50, D, desc, tc, gemm_mode=nv.GEMM_FP32)
    m.initialize(nv.RNG(3))
    cdf = nv.zipf_cdf(D, s)
    m.set_negative_distribution(cdf)
    rng = nv.RNG(seed)
    m.sampler_seed(rng)
    state = seed
    nrng = np.random.default_rng(0)
    for step in range(3):
        labels = nrng.integers(0, D, size=B)
        batch = nv.Batch(B, 2).fill(nrng.integers(0, 50, size=(B, 2)), labels)
        m.step_sampled(batch, 0.0, train=False)
        got = m.entity_ids(B)
        exp, state = host_cdf_labels(labels, z, cdf, state)
        assert (got == exp).all()
        assert m.generate_labels(labels, nv.RNG(1)).shape == exp.shape      # host path of the mirror uses the same cdf
    assert m.sampler_state() == state
    # back to the reference's uniform generator
    m.set_negative_distribution(None)
    labels = nrng.integers(0, D, size=B)
    m.step_sampled(nv.Batch(B, 2).fill(nrng.integers(0, 50, size=(B, 2)), labels), 0.0, train=False)
    exp, state = O.generate_labels(labels, z, D, state)
    assert (m.entity_ids(B) == exp).all() and m.sampler_state() == state
