"""Does tcgen05 kind::tf32 truncate or round fp32 operands? (decides how the 3xTF32 split is built)"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cunvsm_b200 as nv
from tests.test_gpu_gemm import gemm
m = nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))
M, N, K = 128, 16, 32
for name, val in (("1+2^-11", 1 + 2.0**-11), ("1+2^-11+2^-12", 1 + 2.0**-11 + 2.0**-12), ("1+2^-10", 1 + 2.0**-10),
                  ("1+2^-10+2^-11", 1 + 2.0**-10 + 2.0**-11), ("1+2^-12", 1 + 2.0**-12), ("-(1+2^-11+2^-12)", -(1 + 2.0**-11 + 2.0**-12))):
    A = np.zeros((M, K), np.float32); A[:, 0] = val
    B = np.zeros((N, K), np.float32); B[:, 0] = 1.0
    C = gemm(m, 0, A, B)
    print("%-20s A-operand: C=%.10f  (exact %.10f; trunc->%.10f)" % (name, C[0, 0], np.float32(val), np.float32(np.frombuffer((np.float32(val).view(np.uint32) & np.uint32(0xFFFFE000)).tobytes(), np.float32)[0])))
    C = gemm(m, 0, B[:M//8*0+N].repeat(8, 0)[:M] if False else np.tile(B[:1], (M, 1)), np.tile(A[:1], (N, 1)))
    print("%-20s B-operand: C=%.10f" % (name, C[0, 0]))
