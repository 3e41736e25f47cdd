import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cunvsm_b200 as nv
from tests.test_gpu_gemm import gemm
m = nv.Model(8, 8, nv.ModelDesc(word_repr_size=8, entity_repr_size=8), nv.TrainConfig(batch_size=8, window_size=1))
np.set_printoptions(linewidth=250, threshold=100000)
M, N, K = 128, 64, 32
for (k0, m0) in [(0, 0), (0, 1), (0, 5), (1, 0), (3, 33), (9, 70), (31, 127)]:
    A = np.zeros((K, M), np.float32); A[k0, m0] = 1
    B = np.zeros((K, N), np.float32); B[k0, :] = np.arange(1, N + 1)
    C = gemm(m, 1, A, B)
    rows = np.nonzero(np.abs(C).sum(1))[0]
    print("A one-hot k=%d m=%d -> nonzero rows %s" % (k0, m0, rows[:10]), "row vals:", C[rows[0]][:40] if len(rows) else None)
for (k0, n0) in [(0, 0), (0, 1), (0, 7), (2, 40), (8, 3), (31, 63)]:
    A = np.zeros((K, M), np.float32); A[k0, :] = np.arange(1, M + 1)
    B = np.zeros((K, N), np.float32); B[k0, n0] = 1
    C = gemm(m, 1, A, B)
    cols = np.nonzero(np.abs(C).sum(0))[0]
    print("B one-hot k=%d n=%d -> nonzero cols %s" % (k0, n0, cols[:10]), "col vals:", C[:, cols[0]][:40] if len(cols) else None)
# cross-k check: A at k=ka, B at k=kb should give zero unless ka==kb
for ka, kb in [(0, 1), (0, 8), (1, 9), (3, 3)]:
    A = np.zeros((K, M), np.float32); A[ka, 2] = 1
    B = np.zeros((K, N), np.float32); B[kb, 3] = 1
    C = gemm(m, 1, A, B)
    print("ka=%d kb=%d sum=%g C[2,3]=%g nz=%s" % (ka, kb, C.sum(), C[2, 3], np.argwhere(C != 0)[:5].tolist()))
