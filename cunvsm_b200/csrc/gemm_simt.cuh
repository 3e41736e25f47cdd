// FP32 SIMT GEMM (exact-arithmetic path, any shape). This is the fp32 reference mode of
// the three projection GEMMs (NVSM_GEMM_FP32); the tensor-core path lives in
// gemm_tcgen05.cuh. Replaces the cuBLAS calls behind device_matrix's matrix_mult:
//   forward       Z  = P . T           cpp/params.cu:417-421
//   grad_transform gT = P^T . dX        cpp/params.cu:528-531   (long K: split-K)
//   grad_phrase   gP = dX . T^T / n     cpp/objective.cu:453-476
//
// C[M,N] (row-major, ldc) = alpha * opA(A) . opB(B) (+ bias[n])
//   TA == false: A is [M,K] row-major (lda);  TA == true: A is [K,M] row-major (lda)
//   TB == false: B is [K,N] row-major (ldb);  TB == true: B is [N,K] row-major (ldb)
// blockIdx.z selects a K-slice of k_per_split; slice z writes C + z*M*ldc (partials).
#pragma once

#include "common.cuh"

namespace nvsm {

constexpr int GEMM_BM = 128, GEMM_BN = 128, GEMM_BK = 8;

template <bool TA, bool TB>
__global__ void __launch_bounds__(256) sgemm_kernel(int M, int N, int K, const float* __restrict__ A, int lda,
                                                    const float* __restrict__ B, int ldb,
                                                    float* __restrict__ C, int ldc, int k_per_split,
                                                    float alpha, const float* __restrict__ bias) {
    __shared__ float As[2][GEMM_BK][GEMM_BM + 4];
    __shared__ float Bs[2][GEMM_BK][GEMM_BN + 4];
    const int tid = threadIdx.x;
    const int m0 = blockIdx.y * GEMM_BM, n0 = blockIdx.x * GEMM_BN;
    const int kbeg = blockIdx.z * k_per_split;
    const int kend = min(K, kbeg + k_per_split);
    C += (long)blockIdx.z * M * ldc;

    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, 8 x 8 outputs each
    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    float ra[4], rb[4];
    auto load_tiles = [&](int k0) {
        // A tile: BM x BK
        if constexpr (!TA) {
            const int m = m0 + (tid >> 1), kk = (tid & 1) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + kk + j;
                ra[j] = (m < M && k < kend) ? __ldg(A + (long)m * lda + k) : 0.f;
            }
        } else {
            const int kk = tid >> 5, mm = (tid & 31) * 4;
            const int k = k0 + kk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int m = m0 + mm + j;
                ra[j] = (m < M && k < kend) ? __ldg(A + (long)k * lda + m) : 0.f;
            }
        }
        if constexpr (!TB) {
            const int kk = tid >> 5, nn = (tid & 31) * 4;
            const int k = k0 + kk;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int n = n0 + nn + j;
                rb[j] = (n < N && k < kend) ? __ldg(B + (long)k * ldb + n) : 0.f;
            }
        } else {
            const int n = n0 + (tid >> 1), kk = (tid & 1) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const int k = k0 + kk + j;
                rb[j] = (n < N && k < kend) ? __ldg(B + (long)n * ldb + k) : 0.f;
            }
        }
    };
    auto store_tiles = [&](int buf) {
        if constexpr (!TA) {
            const int mm = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) As[buf][kk + j][mm] = ra[j];
        } else {
            const int kk = tid >> 5, mm = (tid & 31) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) As[buf][kk][mm + j] = ra[j];
        }
        if constexpr (!TB) {
            const int kk = tid >> 5, nn = (tid & 31) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[buf][kk][nn + j] = rb[j];
        } else {
            const int nn = tid >> 1, kk = (tid & 1) * 4;
#pragma unroll
            for (int j = 0; j < 4; ++j) Bs[buf][kk + j][nn] = rb[j];
        }
    };

    int buf = 0;
    if (kbeg < kend) {
        load_tiles(kbeg);
        store_tiles(0);
    }
    __syncthreads();
    for (int k0 = kbeg; k0 < kend; k0 += GEMM_BK) {
        const bool has_next = k0 + GEMM_BK < kend;
        if (has_next) load_tiles(k0 + GEMM_BK);  // global loads in flight during the FMAs
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            float a[8], b[8];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][ty * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 8]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][tx * 8 + 4]);
            a[0] = a0.x; a[1] = a0.y; a[2] = a0.z; a[3] = a0.w; a[4] = a1.x; a[5] = a1.y; a[6] = a1.z; a[7] = a1.w;
            b[0] = b0.x; b[1] = b0.y; b[2] = b0.z; b[3] = b0.w; b[4] = b1.x; b[5] = b1.y; b[6] = b1.z; b[7] = b1.w;
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] += a[i] * b[j];
        }
        if (has_next) {
            store_tiles(buf ^ 1);
            __syncthreads();
            buf ^= 1;
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const int m = m0 + ty * 8 + i;
        if (m >= M) continue;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int n = n0 + tx * 8 + j;
            if (n < N) {
                float v = alpha * acc[i][j];
                if (bias) v += __ldg(bias + n);
                C[(long)m * ldc + n] = v;
            }
        }
    }
}

}  // namespace nvsm
