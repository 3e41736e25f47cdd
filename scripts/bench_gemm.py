"""Time the tcgen05 projection GEMMs alone at the C2 shapes (device-resident operands)."""
import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunvsm_b200 as nv
from cunvsm_b200 import _lib
m = nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))
def t(variant, M, N, K, splits=1, stats=0, iters=50):
    ms = ctypes.c_float()
    _lib.check(m.L.nvsm_bench_gemm_tc(m.h, variant, M, N, K, splits, stats, iters, ctypes.byref(ms)))
    fl = 2.0 * M * N * K
    print("variant %d M=%d N=%d K=%d splits=%d stats=%d: %.1f us  %.1f TFLOP/s" % (variant, M, N, K, splits, stats, ms.value * 1e3, fl / ms.value / 1e9))
B = 51200
t(0, B, 256, 300); t(0, B, 256, 300, stats=1); t(0, B, 300, 256); t(1, 300, 256, B, splits=49)
t(0, B, 256, 320); t(0, B, 256, 256); t(0, B, 128, 300)
