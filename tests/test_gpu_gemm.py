"""-m gpu: the tcgen05 / TMA projection GEMM in isolation against numpy (float64 reference).

kind::tf32 keeps 10 mantissa bits of each operand, so the tolerance is the TF32 one:
|err| <= 2e-3 * sqrt(K) * rms(a) * rms(b)-scale; the bit-level sanity check is the exact test with
operands that are exactly representable in TF32 (small integers), which must match to fp32
round-off and catches any descriptor / swizzle / layout mistake."""
import ctypes

import numpy as np
import pytest

import cunvsm_b200 as nv
from cunvsm_b200 import _lib

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def model():
    return nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))


def gemm(model, variant, A, B, alpha=1.0, bias=None, splits=1):
    pf = ctypes.POINTER(ctypes.c_float)
    A = np.ascontiguousarray(A, dtype=np.float32); B = np.ascontiguousarray(B, dtype=np.float32)
    if variant in (0, 2):
        M, K = A.shape; N = B.shape[0]
    else:
        K, M = A.shape; N = B.shape[1]
    C = np.zeros((M, N), dtype=np.float32)
    bp = None
    if bias is not None:
        bias = np.ascontiguousarray(bias, dtype=np.float32); bp = bias.ctypes.data_as(pf)
    _lib.check(model.L.nvsm_test_gemm_tc(model.h, variant, M, N, K, A.ctypes.data_as(pf), B.ctypes.data_as(pf),
                                         C.ctypes.data_as(pf), alpha, bp, splits))
    return C


SHAPES_K = [(128, 256, 32), (128, 256, 300), (256, 256, 300), (1000, 256, 300), (384, 300, 256), (130, 64, 64),
            (128, 16, 8), (4096, 128, 128), (257, 304, 40)]


@pytest.mark.parametrize("M,N,K", SHAPES_K)
def test_k_major_exact_small_integers(model, M, N, K):
    rng = np.random.default_rng(M * 7 + N * 3 + K)
    A = rng.integers(-4, 5, size=(M, K)).astype(np.float32)
    B = rng.integers(-4, 5, size=(N, K)).astype(np.float32)
    bias = rng.integers(-3, 4, size=N).astype(np.float32)
    C = gemm(model, 0, A, B, alpha=0.5, bias=bias)
    ref = 0.5 * (A.astype(np.float64) @ B.astype(np.float64).T) + bias
    np.testing.assert_array_equal(C, ref.astype(np.float32))


SHAPES_MN = [(128, 256, 32, 1), (300, 256, 1024, 4), (300, 256, 5000, 37), (64, 64, 4096, 8), (128, 32, 64, 2),
             (300, 256, 51200, 49)]


@pytest.mark.parametrize("M,N,K,splits", SHAPES_MN)
def test_mn_major_exact_small_integers(model, M, N, K, splits):
    rng = np.random.default_rng(M + N + K)
    A = rng.integers(-2, 3, size=(K, M)).astype(np.float32)
    B = rng.integers(-2, 3, size=(K, N)).astype(np.float32)
    C = gemm(model, 1, A, B, splits=splits)
    ref = A.astype(np.float64).T @ B.astype(np.float64)
    np.testing.assert_array_equal(C, ref.astype(np.float32))


def test_tf32_error_bound_random(model):
    rng = np.random.default_rng(0)
    A = rng.standard_normal((512, 300)).astype(np.float32)
    B = rng.standard_normal((256, 300)).astype(np.float32)
    C = gemm(model, 0, A, B)
    ref = A.astype(np.float64) @ B.astype(np.float64).T
    err = np.abs(C - ref).max()
    assert err <= 2e-3 * np.sqrt(300) * 3, err


@pytest.mark.parametrize("variant,M,N,K,splits", [(2, 512, 256, 300, 1), (2, 1000, 300, 256, 1), (3, 300, 256, 4096, 8), (3, 300, 256, 51200, 49)])
def test_3xtf32_reaches_fp32_accuracy(model, variant, M, N, K, splits):
    """3xTF32 (hi.hi + lo.hi + hi.lo): error must be at the fp32 level (1e-6 relative to the
    sum of |a||b|), ~1000x below single-pass TF32."""
    rng = np.random.default_rng(3)
    if variant == 2:
        A = rng.standard_normal((M, K)).astype(np.float32); B = rng.standard_normal((N, K)).astype(np.float32)
        ref = A.astype(np.float64) @ B.astype(np.float64).T
        mag = np.abs(A).astype(np.float64) @ np.abs(B).astype(np.float64).T
    else:
        A = rng.standard_normal((K, M)).astype(np.float32); B = rng.standard_normal((K, N)).astype(np.float32)
        ref = A.astype(np.float64).T @ B.astype(np.float64)
        mag = np.abs(A).astype(np.float64).T @ np.abs(B).astype(np.float64)
    C = gemm(model, variant, A, B, splits=splits)
    rel = (np.abs(C - ref) / mag).max()
    assert rel <= 2e-6, rel
