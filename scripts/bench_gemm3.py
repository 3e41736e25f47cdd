import ctypes, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cunvsm_b200 as nv
from cunvsm_b200 import _lib
m = nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))
def t(variant, M, N, K, splits=1, stats=0, iters=30):
    ms = ctypes.c_float()
    _lib.check(m.L.nvsm_bench_gemm_tc(m.h, variant, M, N, K, splits, stats, iters, ctypes.byref(ms)))
    print("v%d M=%d N=%d K=%d: %.1f us" % (variant, M, N, K, ms.value * 1e3), flush=True)
t(0, 51200, 256, 300)
