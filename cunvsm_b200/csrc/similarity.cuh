// RepresentationSimilarity objective (EntityEntity / TermTerm, and the second constituent of the
// TextEntityEntityEntity / TextEntityTermTerm mixtures): cpp/objective.cu:485-700 of the reference.
//
//   forward  (:487-573): rows a = table[ids[2i]], b = table[ids[2i+1]]; p_i = clamp(sigmoid(a . b));
//                        mass_i = w_i log p_i; cost = -(1/N) sum_i mass_i
//   backward (:575-672): mult_i = w_i (1/N) (p_i in the clamp band ? 0 : 1 - p_i);
//                        grad[:, 2i] = mult_i b, grad[:, 2i+1] = mult_i a   (flip_adjacent_columns)
// One warp per pair does both while the two rows are in registers; the reference materialises the gathered
// rows, their product, a copy, and the flipped matrix (4 x dim x 2N floats). `scale` carries the mixture weight
// w_k / sum_k w_k that MergeGradientsFn applies to every constituent gradient (cpp/intermediate_results.cu:3-60).
#pragma once

#include "common.cuh"

namespace nvsm {

struct PairParams {
    const float* table;   // [objects, dim]
    int dim;
    const idx_t* ids;     // [2N]
    const float* w;       // [N]
    long N;
    float sig_lo_cmp, sig_lo_val, sig_hi_cmp, sig_hi_val;  // forward clamp (see nvsm.cu:forward)
    float der_lo_cmp, der_hi_cmp;                          // backward zero-gradient band
    float bsn;            // exp(-log(N))
    float scale;          // mixture weight of this constituent (1 when it is the only objective)
    float* probs;         // [N]
    float* mult;          // [N]
    float* G;             // [2N, dim] gradient columns, already multiplied by `scale`
    double* loss_acc;     // [1] sum_i w_i log p_i
};

template <int VEC, int NCH>
__global__ void __launch_bounds__(256) pair_forward_backward_kernel(const PairParams p) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = p.dim / VEC;
    float loss = 0.f;
    for (long i = warp0; i < p.N; i += nwarps) {
        const idx_t ia = __ldg(p.ids + 2 * i), ib = __ldg(p.ids + 2 * i + 1);
        float a[NCH][VEC], b[NCH][VEC];
        float dot = 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
#pragma unroll
            for (int v = 0; v < VEC; ++v) { a[j][v] = 0.f; b[j][v] = 0.f; }
            if (c < nvec) {
                load_vec_ro<VEC>(p.table + ia * p.dim + c * VEC, a[j]);
                load_vec_ro<VEC>(p.table + ib * p.dim + c * VEC, b[j]);
            }
#pragma unroll
            for (int v = 0; v < VEC; ++v) dot += a[j][v] * b[j][v];
        }
        dot = warp_sum(dot);
        // numerically stable sigmoid (include/cuNVSM/cuda_utils.h:192-214)
        float prob;
        if (dot >= 0.f) {
            prob = 1.0f / (1.0f + expf(-dot));
        } else {
            const float ex = expf(dot);
            prob = ex / (1.0f + ex);
        }
        prob = prob < p.sig_lo_cmp ? p.sig_lo_val : (prob > p.sig_hi_cmp ? p.sig_hi_val : prob);
        const float w = __ldg(p.w + i);
        const float der = (prob >= p.der_hi_cmp || prob <= p.der_lo_cmp) ? 0.0f : 1.0f - prob;
        const float m = w * (der * p.bsn);
        if (lane == 0) {
            loss += w * logf(prob);
            p.probs[i] = prob;
            p.mult[i] = m;
        }
        const float ms = m * p.scale;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            if (c < nvec) {
                float ga[VEC], gb[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) { ga[v] = ms * b[j][v]; gb[v] = ms * a[j][v]; }
                store_vec<VEC>(p.G + (2 * i) * p.dim + c * VEC, ga);
                store_vec<VEC>(p.G + (2 * i + 1) * p.dim + c * VEC, gb);
            }
        }
    }
    loss = warp_sum(loss);
    if (lane == 0 && loss != 0.f) atomicAdd(p.loss_acc, (double)loss);
}

// update_repr_kernel with window 1 and no weights (cpp/storage.cu:37-49):
//   target[ids[c], :] += scale * (acc ? 1 / sqrt(acc[ids[c]] + eps) : 1) * G[c, :]
template <int VEC>
__global__ void __launch_bounds__(256) rows_scatter_kernel(const float* __restrict__ G, const idx_t* __restrict__ ids,
                                                           long M, int dim, float* __restrict__ target, float scale,
                                                           const float* __restrict__ acc, float eps) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim / VEC;
    for (long c = warp0; c < M; c += nwarps) {
        const idx_t id = __ldg(ids + c);
        float coef = scale;
        if (acc) coef = coef / sqrtf(__ldg(acc + id) + eps);
        for (int k = lane; k < nvec; k += kWarp) {
            float g[VEC];
            load_vec_ro<VEC>(G + c * dim + k * VEC, g);
#pragma unroll
            for (int v = 0; v < VEC; ++v) g[v] *= coef;
            red_add_vec<VEC>(target + id * dim + k * VEC, g);
        }
    }
}

// acc[ids[c]] += scale * msq[c]   (scalar moments of Adagrad / sparse and dense-update Adam, window 1)
__global__ void rows_scalar_scatter_kernel(const idx_t* __restrict__ ids, const float* __restrict__ msq, long M,
                                           float scale, float* __restrict__ acc) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= M) return;
    atomicAdd(acc + ids[c], scale * msq[c]);
}

}  // namespace nvsm
