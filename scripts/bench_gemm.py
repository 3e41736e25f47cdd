"""Time the tcgen05 projection GEMMs alone at the C2 shapes (device-resident zero operands):
forward Z = P.T (K-major, CTA pairs), grad_phrase (K-major, two 160-wide N tiles), grad_transform (MN-major, split-K),
each in single-pass TF32 and 3xTF32, the forward one with and without the fused column statistics.
    NVSM_TC_2CTA=0 python scripts/bench_gemm.py     # forward on the single-SM kernel"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunvsm_b200 as nv
from cunvsm_b200 import _lib

m = nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))


def t(name, variant, M, N, K, splits=1, stats=False, split3=True, iters=50):
    ms = ctypes.c_float()
    _lib.check(m.L.nvsm_bench_gemm_tc(m.h, variant, M, N, K, splits, (1 if stats else 0) | (2 if split3 else 0), iters, ctypes.byref(ms)))
    fl = 2.0 * M * N * K * (3 if split3 else 1)
    print("%-28s M=%d N=%d K=%d splits=%d %s%s: %.1f us  %.0f TFLOP/s issued" % (
        name, M, N, K, splits, "3xTF32" if split3 else "TF32", " +stats" if stats else "", ms.value * 1e3, fl / ms.value / 1e9))


B = int(os.environ.get("B", 51200))
for split3 in (True, False):
    t("forward", 0, B, 256, 300, split3=split3)
    t("forward + column stats", 0, B, 256, 300, stats=True, split3=split3)
    t("grad_phrase", 0, B, 300, 256, split3=split3)
    t("grad_transform", 1, 300, 256, B, splits=49, split3=split3)
