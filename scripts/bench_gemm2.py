import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunvsm_b200 as nv
from cunvsm_b200 import _lib
m = nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))
def t(variant, M, N, K, splits=1, stats=0, iters=30):
    ms = ctypes.c_float()
    _lib.check(m.L.nvsm_bench_gemm_tc(m.h, variant, M, N, K, splits, stats, iters, ctypes.byref(ms)))
    fl = 2.0 * M * N * K
    byts = 4.0 * (M * K + (M / 128) * N * K + M * N)
    print("v%d M=%d N=%d K=%d: %.1f us  %.1f TFLOP/s  %.2f TB/s(smem-fill+store)" % (variant, M, N, K, ms.value * 1e3, fl / ms.value / 1e9, byts / ms.value / 1e9), flush=True)
for (M, K) in [(148 * 128, 320), (148 * 128, 3200), (148 * 128, 32000), (148 * 128 * 2, 320), (148 * 128 * 3, 320), (51200, 320), (51200, 3200), (512000, 320)]:
    t(0, M, 256, K)
