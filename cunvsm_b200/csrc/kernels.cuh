// Hand-written sm_100a kernels of the NVSM/LSE step: embedding gather-mean, batch-norm
// statistics, the fused score / NCE-loss / backward-through-dot kernel, batch-norm
// backward, and the sparse scatter-add + optimiser kernels. All are HBM/L2-bandwidth
// bound integer-indexed row traffic: one warp per n-gram, 128-bit accesses, warp-shuffle
// reductions, vector REDG for the scatter.
//
// Reference behaviour each kernel replaces is cited per kernel (paths relative to the
// reference tree). Tables are row-major [objects, dim] fp32, i.e. the memory image of the
// reference's column-major dim x objects matrices.
#pragma once

#include "common.cuh"
#include "peer_allreduce.cuh"

namespace nvsm {

// =====================================================================================
// gather_mean: P[o, :] = (1/window) * sum_w wt[o, w] * table[idx[o, w], :]
// Replaces average_repr_kernel (cpp/params.cu:75-95) for both get_average_representations
// (:138-172) and the plain gather get_representations (:97-121; window 1, no weights).
// The mean divides by the window even when weighted (quirk pinned by
// cpp/model_tests.cu:115-122).
// =====================================================================================
template <int VEC>
__global__ void __launch_bounds__(256) gather_mean_kernel(const float* __restrict__ table, int dim,
                                                          const idx_t* __restrict__ ids,
                                                          const float* __restrict__ wts,
                                                          long num_out, int window,
                                                          float* __restrict__ out, int ld_out, int tf32,
                                                          float* __restrict__ out_lo) {
    pdl_launch_dependents();   // the forward GEMM behind may set up (barriers, TMEM) while the last blocks drain
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim / VEC;
    const float fwin = (float)window;
    for (long o = warp0; o < num_out; o += nwarps) {
        const idx_t* oid = ids + o * window;
        const float* ow = wts ? wts + o * window : nullptr;
        for (int c = lane; c < nvec; c += kWarp) {
            float acc[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) acc[v] = 0.f;
#pragma unroll 5
            for (int w = 0; w < window; ++w) {
                const idx_t id = __ldg(oid + w);
                const float wt = ow ? __ldg(ow + w) : 1.0f;
                float x[VEC];
                load_vec_ro<VEC>(table + id * dim + c * VEC, x);
#pragma unroll
                for (int v = 0; v < VEC; ++v) acc[v] += wt * x[v];
            }
            float lo[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                acc[v] = acc[v] / fwin;
                lo[v] = 0.f;
                if (tf32) {                              // P only feeds the tensor-core GEMMs
                    const float hi = round_tf32(acc[v]);
                    lo[v] = acc[v] - hi;                 // exact; second operand of the 3xTF32 split
                    acc[v] = hi;
                }
            }
            store_vec<VEC>(out + o * ld_out + c * VEC, acc);
            if (out_lo) store_vec<VEC>(out_lo + o * ld_out + c * VEC, lo);
        }
    }
}

// Single-pass variant for dim % 4 == 0 and dim <= 512 floats: a group of LG lanes (power of two, <= 32) owns
// one n-gram; lane sl < L of the group keeps K float4 column groups (c = sl + j L) in registers, so every
// table row is visited ONCE (K independent 16-byte loads per lane per word) instead of once per 32-column
// pass, and the window's ids / weights are loaded once per n-gram (one lane each) and broadcast with
// shuffles. ncu on the pass-per-32-columns kernel above (C2, d_w = 300): 942 warp instructions per n-gram,
// issue slots 49 % busy, i.e. issue-bound, the third pass running 11 of 32 lanes; this layout needs ~270.
// U words are loaded back to back before their FMAs (NVSM_GATHER_U).
// Division by the window follows the reference's release build (-use_fast_math => __fdividef,
// CMakeLists.txt:71-73 of the reference).
template <int K, int LG>
__global__ void __launch_bounds__(256) gather_mean_lanes_kernel(const float* __restrict__ table, int dim,
                                                                const idx_t* __restrict__ ids,
                                                                const float* __restrict__ wts,
                                                                long num_out, int window, int L,
                                                                float* __restrict__ out, int ld_out, int tf32,
                                                                float* __restrict__ out_lo) {
    constexpr int kGroups = kWarp / LG;
#ifndef NVSM_GATHER_U
#define NVSM_GATHER_U 2   // measured on C2 (us): U=1 83.2, 2 81.5, 3 90.9, 5 87.9, 10 117 -- occupancy beats loads in flight
#endif
    constexpr int U = NVSM_GATHER_U;   // words in flight per lane: U * K independent 16-byte loads
    pdl_launch_dependents();   // the forward GEMM behind may set up (barriers, TMEM) while the last blocks drain
    const int lane = threadIdx.x & 31;
    const int sl = lane & (LG - 1);
    const int grp = lane / LG;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim >> 2;
    const float fwin = (float)window;
    const float4* __restrict__ tab4 = reinterpret_cast<const float4*>(table);
    // Lanes beyond the row (and groups beyond the batch) read clamped, valid addresses and skip the store:
    // the load / FMA loop stays branch-free, which lets all U * K loads issue back to back.
    int cj[K];
#pragma unroll
    for (int j = 0; j < K; ++j) cj[j] = min(sl + j * L, nvec - 1);
    for (long o0 = warp0 * kGroups; o0 < num_out; o0 += nwarps * kGroups) {
        const long o = min(o0 + grp, num_out - 1);
        float4 acc[K];
#pragma unroll
        for (int j = 0; j < K; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int w0 = 0; w0 < window; w0 += LG) {
            const int chunk = min(LG, window - w0);
            // one lane per word: vector offset of its row and its weight
            unsigned my_row = 0;
            float my_wt = 0.f;
            if (sl < chunk) {
                my_row = (unsigned)(__ldg(ids + o * window + w0 + sl) * nvec);
                my_wt = wts ? __ldg(wts + o * window + w0 + sl) : 1.0f;
            }
            int w = 0;
            for (; w + U <= chunk; w += U) {
                float4 x[U][K];
                float wt[U];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const unsigned row = __shfl_sync(kFull, my_row, w + u, LG);
                    wt[u] = __shfl_sync(kFull, my_wt, w + u, LG);
#pragma unroll
                    for (int j = 0; j < K; ++j) x[u][j] = __ldg(tab4 + row + cj[j]);
                }
#pragma unroll
                for (int u = 0; u < U; ++u) {
#pragma unroll
                    for (int j = 0; j < K; ++j) {
                        acc[j].x = fmaf(wt[u], x[u][j].x, acc[j].x); acc[j].y = fmaf(wt[u], x[u][j].y, acc[j].y);
                        acc[j].z = fmaf(wt[u], x[u][j].z, acc[j].z); acc[j].w = fmaf(wt[u], x[u][j].w, acc[j].w);
                    }
                }
            }
            for (; w < chunk; ++w) {
                const unsigned row = __shfl_sync(kFull, my_row, w, LG);
                const float wt1 = __shfl_sync(kFull, my_wt, w, LG);
                float4 x1[K];
#pragma unroll
                for (int j = 0; j < K; ++j) x1[j] = __ldg(tab4 + row + cj[j]);
#pragma unroll
                for (int j = 0; j < K; ++j) {
                    acc[j].x = fmaf(wt1, x1[j].x, acc[j].x); acc[j].y = fmaf(wt1, x1[j].y, acc[j].y);
                    acc[j].z = fmaf(wt1, x1[j].z, acc[j].z); acc[j].w = fmaf(wt1, x1[j].w, acc[j].w);
                }
            }
        }
        if (sl < L && o0 + grp < num_out) {
#pragma unroll
            for (int j = 0; j < K; ++j) {
                const int c = sl + j * L;
                if (c >= nvec) continue;
                float a[4] = {__fdividef(acc[j].x, fwin), __fdividef(acc[j].y, fwin), __fdividef(acc[j].z, fwin),
                              __fdividef(acc[j].w, fwin)};
                float lo[4] = {0.f, 0.f, 0.f, 0.f};
                if (tf32) {   // P only feeds the tensor-core GEMMs: hi = rn_tf32(x), lo = x - hi (exact)
#pragma unroll
                    for (int v = 0; v < 4; ++v) {
                        const float hi = round_tf32(a[v]);
                        lo[v] = a[v] - hi;
                        a[v] = hi;
                    }
                }
                store_vec<4>(out + o * ld_out + c * 4, a);
                if (out_lo) store_vec<4>(out_lo + o * ld_out + c * 4, lo);
            }
        }
    }
}

// =====================================================================================
// col_stats: sums[c] += sum_i Z[i, c], sums[dd + c] += sum_i Z[i, c]^2 (double).
// First half of cudnnBatchNormalizationForwardTraining, per-activation mode
// (cpp/cudnn_utils.cu:107-124). fp32 partials per thread, one double atomic per column
// per block.
// =====================================================================================
__global__ void __launch_bounds__(256) col_stats_kernel(const float* __restrict__ Z, long rows, int dd,
                                                        double* __restrict__ sums) {
    // thread t owns column (t % cols_per_pass) of every (t / cols_per_pass)-th row of its slab.
    const int tpr = min(dd, (int)blockDim.x);          // threads per row
    const int rpp = blockDim.x / tpr;                   // rows per pass
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    if (tr >= rpp) return;
    const long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = min(rows, r0 + rows_per_block);
    for (int c = tc; c < dd; c += tpr) {
        float s = 0.f, q = 0.f;
        for (long r = r0 + tr; r < r1; r += rpp) {
            const float x = __ldg(Z + r * dd + c);
            s += x;
            q += x * x;
        }
        if (r0 + tr < r1) {
            atomicAdd(sums + c, (double)s);
            atomicAdd(sums + dd + c, (double)q);
        }
    }
}

// Same statistics for dd % 4 == 0, without atomics: every thread owns one float4 column group
// and streams its slab of rows with four independent 128-bit loads in flight (Z is L2-resident:
// it was just written by the projection GEMM); the row groups of a block are combined in
// shared memory and each block writes one fp32 partial row [2*dd]. col_stats_reduce_kernel
// then sums the partials in double (deterministic).
__global__ void __launch_bounds__(256) col_stats4_kernel(const float* __restrict__ Z, long rows, int dd,
                                                         float* __restrict__ partials) {
    extern __shared__ float sm[];  // [rpp][2*dd]
    const int nvec = dd >> 2;
    const int tpr = min(nvec, (int)blockDim.x);
    const int rpp = blockDim.x / tpr;
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    const long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = min(rows, r0 + rows_per_block);
    if (tr < rpp) {
        for (int c = tc; c < nvec; c += tpr) {
            float4 s = make_float4(0.f, 0.f, 0.f, 0.f), q = s;
            const float4* base = reinterpret_cast<const float4*>(Z) + c;
            long r = r0 + tr;
            for (; r + 3 * rpp < r1; r += 4 * rpp) {
                const float4 x0 = __ldg(base + (r) * nvec), x1 = __ldg(base + (r + rpp) * nvec);
                const float4 x2 = __ldg(base + (r + 2 * rpp) * nvec), x3 = __ldg(base + (r + 3 * rpp) * nvec);
                s.x += (x0.x + x1.x) + (x2.x + x3.x); s.y += (x0.y + x1.y) + (x2.y + x3.y);
                s.z += (x0.z + x1.z) + (x2.z + x3.z); s.w += (x0.w + x1.w) + (x2.w + x3.w);
                q.x += (x0.x * x0.x + x1.x * x1.x) + (x2.x * x2.x + x3.x * x3.x);
                q.y += (x0.y * x0.y + x1.y * x1.y) + (x2.y * x2.y + x3.y * x3.y);
                q.z += (x0.z * x0.z + x1.z * x1.z) + (x2.z * x2.z + x3.z * x3.z);
                q.w += (x0.w * x0.w + x1.w * x1.w) + (x2.w * x2.w + x3.w * x3.w);
            }
            for (; r < r1; r += rpp) {
                const float4 x = __ldg(base + r * nvec);
                s.x += x.x; s.y += x.y; s.z += x.z; s.w += x.w;
                q.x += x.x * x.x; q.y += x.y * x.y; q.z += x.z * x.z; q.w += x.w * x.w;
            }
            float* o = sm + (long)tr * 2 * dd;
            *reinterpret_cast<float4*>(o + 4 * c) = s;
            *reinterpret_cast<float4*>(o + dd + 4 * c) = q;
        }
    }
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * dd; t += blockDim.x) {
        float a = 0.f;
        for (int g = 0; g < rpp; ++g) a += sm[(long)g * 2 * dd + t];
        partials[(long)blockIdx.x * 2 * dd + t] = a;
    }
}

// sums[t] = sum over blocks of partials[b][t], in double. One warp-row of 32 columns x 8 block
// slices per CTA so the ~300 partial rows are read with 8-way parallelism per column.
__global__ void __launch_bounds__(256) col_stats_reduce_kernel(const float* __restrict__ partials, int nblocks, int n,
                                                               double* __restrict__ sums) {
    __shared__ double sm[8][32];
    const int col = blockIdx.x * 32 + (threadIdx.x & 31);
    const int slice = threadIdx.x >> 5;
    double a = 0.0;
    if (col < n)
        for (int b = slice; b < nblocks; b += 8) a += (double)__ldg(partials + (long)b * n + col);
    sm[slice][threadIdx.x & 31] = a;
    __syncthreads();
    if (slice == 0 && col < n) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < 8; ++g) t += sm[g][threadIdx.x];
        sums[col] = t;
    }
}

// mean / invstd from the (all-reduced) sums; biased variance, eps inside the sqrt.
// Also writes the affine form the staged score kernel consumes: scale = invstd, shift = bias - mean * invstd.
__global__ void bn_finalize_kernel(const double* __restrict__ sums, int dd, double batch, double eps,
                                   float* __restrict__ mean, float* __restrict__ invstd,
                                   const float* __restrict__ bias, float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dd) return;
    const double mu = sums[c] / batch;
    double var = sums[dd + c] / batch - mu * mu;
    if (var < 0.0) var = 0.0;
    const float mf = (float)mu, isf = (float)(1.0 / sqrt(var + eps));
    mean[c] = mf;
    invstd[c] = isf;
    scale[c] = isf;
    shift[c] = bias[c] - mf * isf;
}

// Single-GPU: the partial reduction (col_stats_reduce_kernel) and the finalisation in one launch. One CTA owns 8
// columns x 128 slices of the partial rows (ncu r1i: the former 32-column x 32-slice shape ran 8 CTAs for 14.7 us, a
// serial chain of ~19 dependent L2 round trips per thread; this one runs dd / 8 CTAs with <= 5 per thread). Sums and
// sums of squares together, double accumulation, fixed summation order.
// XCHG (N > 1, NVLink peer exchange): between the reduction and the finalisation every block pushes the sums of ITS eight
// columns to every peer and waits for the peers' (per-block flags), i.e. partial rows -> local sums -> global sums ->
// mean / invstd in one launch instead of reduce, one-block all-reduce and finalize kernels.
template <bool XCHG>
__global__ void __launch_bounds__(1024) col_stats_reduce_finalize_kernel(const float* __restrict__ partials, int nblocks, int dd,
                                                                         double batch, double eps, double* __restrict__ sums,
                                                                         float* __restrict__ mean, float* __restrict__ invstd,
                                                                         const float* __restrict__ bias,
                                                                         float* __restrict__ scale, float* __restrict__ shift,
                                                                         const PeerXchg* __restrict__ xp, unsigned long long epoch,
                                                                         int* __restrict__ xchg_error,
                                                                         double* __restrict__ zero_buf, int zero_n) {
    constexpr int C = 8, S = 128;   // blockDim = C columns x S slices
    pdl_wait();                // (launched as a programmatic dependent of the forward GEMM)
    pdl_launch_dependents();
    // accumulators of the kernels behind this one (backward column sums + loss): zeroed here instead of by a memset
    // operation at the head of the step
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < zero_n; i += gridDim.x * blockDim.x) zero_buf[i] = 0.0;
    __shared__ double sm[2][32][C];
    const int cl = threadIdx.x & (C - 1), slice = threadIdx.x / C;
    const int col = blockIdx.x * C + cl;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double a = 0.0, q = 0.0;
    if (col < dd)
        for (int b = slice; b < nblocks; b += S) {
            a += (double)__ldg(partials + (long)b * 2 * dd + col);
            q += (double)__ldg(partials + (long)b * 2 * dd + dd + col);
        }
    // lanes l, l ^ 8, l ^ 16, l ^ 24 hold the same column
    a += __shfl_xor_sync(kFull, a, 8);  q += __shfl_xor_sync(kFull, q, 8);
    a += __shfl_xor_sync(kFull, a, 16); q += __shfl_xor_sync(kFull, q, 16);
    if (lane < C) { sm[0][warp][lane] = a; sm[1][warp][lane] = q; }
    __syncthreads();
    if (warp != 0) return;
    const bool own = threadIdx.x < C && col < dd;
    double s = 0.0, s2 = 0.0;
    if (own) {
#pragma unroll 8
        for (int g = 0; g < 32; ++g) { s += sm[0][g][threadIdx.x]; s2 += sm[1][g][threadIdx.x]; }
    }
    if (XCHG) {
        const PeerXchg& x = *xp;
        const int parity = (int)(epoch & 1ull);
        if (own)
            for (int p = 0; p < x.nranks; ++p) {
                double* dst = x.inbox[p] + peer_slot_index(x, 0, parity, x.rank);
                dst[col] = s;
                dst[dd + col] = s2;
            }
        __threadfence_system();
        __syncwarp();
        peer_publish(x, 0, parity, blockIdx.x, epoch, lane);
        peer_wait(x, 0, parity, blockIdx.x, epoch, lane, xchg_error);
        __syncwarp();
        if (own) {
            s = 0.0; s2 = 0.0;
            for (int p = 0; p < x.nranks; ++p) {   // rank order: bit-identical on every rank
                s += peer_inbox_value(x, 0, parity, p, col);
                s2 += peer_inbox_value(x, 0, parity, p, dd + col);
            }
        }
    }
    if (own) {
        sums[col] = s;
        sums[dd + col] = s2;
        const double mu = s / batch;
        double var = s2 / batch - mu * mu;
        if (var < 0.0) var = 0.0;
        const float mf = (float)mu, isf = (float)(1.0 / sqrt(var + eps));
        mean[col] = mf;
        invstd[col] = isf;
        scale[col] = isf;
        shift[col] = bias[col] - mf * isf;
    }
}

// Numerically careful variant: second pass accumulating sum (x - mean)^2.
__global__ void __launch_bounds__(256) col_var_kernel(const float* __restrict__ Z, long rows, int dd,
                                                      const double* __restrict__ sums, double batch,
                                                      double* __restrict__ var_sums) {
    const int tpr = min(dd, (int)blockDim.x);
    const int rpp = blockDim.x / tpr;
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    if (tr >= rpp) return;
    const long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * rows_per_block;
    const long r1 = min(rows, r0 + rows_per_block);
    for (int c = tc; c < dd; c += tpr) {
        const float mu = (float)(sums[c] / batch);
        float q = 0.f;
        for (long r = r0 + tr; r < r1; r += rpp) {
            const float d = __ldg(Z + r * dd + c) - mu;
            q += d * d;
        }
        if (r0 + tr < r1) atomicAdd(var_sums + c, (double)q);
    }
}

__global__ void bn_finalize2_kernel(const double* __restrict__ sums, const double* __restrict__ var_sums,
                                    int dd, double batch, double eps, float* __restrict__ mean,
                                    float* __restrict__ invstd) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dd) return;
    mean[c] = (float)(sums[c] / batch);
    invstd[c] = (float)(1.0 / sqrt(var_sums[c] / batch + eps));
}

// =====================================================================================
// score: everything between the projection and the backward GEMMs, without ever
// materialising a d_d x B*R matrix. Per n-gram i (one warp):
//   y      = f(BN?(Z[i]))                                  cpp/params.cu:425-446
//   s_r    = +-(y . E[id[i,r]])        negatives negated    cpp/objective.cu:159-239
//   p_r    = clamp(sigmoid(s_r))                            :242-246
//   loss  += wbc_r * log p_r                                :249-305, intermediate_results.cu:80-124
//   mult_r = wbc_r * (p_r in (eps, 1-eps) ? 1 - p_r : 0) / B     :354-371
//   Gp[i]  = (sum_r mult_r * (+-E[id])) .* f'(y)            :420-425, params.cu:474-491
//   column sums of Gp (-> grad_bias, BN backward)           params.cu:510-520
// =====================================================================================
struct ScoreParams {
    const float* Z;        // [B, dd] pre-activation
    const float* E;        // [D, dd]
    const idx_t* ids;      // [B*R], positive first
    const float* inst_w;   // [B]
    long B;
    int R, dd;
    float w_scale;         // (z+1)/(2z) when rebalancing, else 1   (objective.cu:268-274)
    float pos_scale;       // z when rebalancing, else 1            (objective.cu:282-290)
    // The reference clamps / tests the float probability against double constants
    // (`1.0 - epsilon_`). For a float p those comparisons are equivalent to comparisons with
    // the float thresholds below (host: smallest float >= / largest float <= the double).
    float sig_lo_cmp, sig_lo_val;   // p <  sig_lo_cmp -> sig_lo_val     forward clamp, eps 1e-7 | 0
    float sig_hi_cmp, sig_hi_val;   // p >  sig_hi_cmp -> sig_hi_val
    float der_lo_cmp, der_hi_cmp;   // p <= der_lo_cmp || p >= der_hi_cmp -> zero gradient (eps 1e-6 | 0)
    float bsn;             // exp(-log(B_global))
    ActParams act;
    float* probs;          // [B*R]
    float* mult;           // [B*R]
    float* Gp;             // [B, dd]
    float* Y;              // [B, dd] post-activation (nullable; kept for the pull-style update)
    int tf32_gp;           // round Gp to tf32 (tensor-core GEMMs, no batch-norm: Gp is dX)
    float* Gp_lo;          // nullable: Gp - rn_tf32(Gp) for the 3xTF32 GEMMs (no batch-norm)
    double* loss_acc;      // [1]  sum_c wbc_c log p_c
    double* col_sums;      // [2*dd]: sum_i dy, sum_i dy * xhat
    // l2_normalize_entity_reprs (Normalizer::forward on the gathered rows, cpp/objective.cu:170-176,
    // cpp/cuda_utils.cu:12-46): scores use e / |e|. Non-null => on; per reference the norm |e| and the
    // normalised dot product e.y / |e| are kept for the backward pass (cpp/cuda_utils.cu:69-127).
    float* enorm;          // [B*R] nullable
    float* escore;         // [B*R] nullable
    // Last-block tail (score_sums_tail; xchg_counter null = off). N > 1 with the NVLink peer exchange (xchg non-null):
    // col_sums + loss_acc (contiguous, [2*dd + 1]) are all-reduced across the ranks inside this kernel. loss_host
    // (mapped pinned, nullable): the final loss sum is written there, no D2H copy in the stream.
    const PeerXchg* xchg;
    unsigned long long xchg_epoch;
    unsigned int* xchg_counter;
    int* xchg_error;
    double* loss_host;
};

// Reduce four per-lane partial sums across the warp with 6 shuffles (instead of 20): after the
// call, `d0` of lane L holds the warp total of partial ((L >> 3) & 3).
__device__ __forceinline__ float warp_sum4_transposed(float d0, float d1, float d2, float d3, int lane) {
    const bool b4 = lane & 16, b3 = lane & 8;
    float a0 = b4 ? d2 : d0, a1 = b4 ? d3 : d1;
    const float s0 = b4 ? d0 : d2, s1 = b4 ? d1 : d3;
    a0 += __shfl_xor_sync(kFull, s0, 16);
    a1 += __shfl_xor_sync(kFull, s1, 16);
    float b = b3 ? a1 : a0;
    const float s = b3 ? a0 : a1;
    b += __shfl_xor_sync(kFull, s, 8);
    b += __shfl_xor_sync(kFull, b, 4);
    b += __shfl_xor_sync(kFull, b, 2);
    b += __shfl_xor_sync(kFull, b, 1);
    return b;
}

template <int VEC, int NCH>
__global__ void __launch_bounds__(256) score_kernel(const ScoreParams p) {
    constexpr int RB = 4;  // entity rows in flight per warp
    extern __shared__ float smem[];  // [2*dd] column sums + [1] loss
    const int lane = threadIdx.x & 31;
    const int dd = p.dd;
    const int nvec = dd / VEC;
    for (int t = threadIdx.x; t < 2 * dd + 1; t += blockDim.x) smem[t] = 0.f;
    __syncthreads();

    float cs[NCH][VEC], cx[NCH][VEC];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { cs[j][v] = 0.f; cx[j][v] = 0.f; }
    float loss = 0.f;
    const int my_slot = (lane >> 3) & 3;      // which of the RB rows this lane finishes
    const bool slot_leader = (lane & 7) == 0;

    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long i = warp0; i < p.B; i += nwarps) {
        float y[NCH][VEC], xh[NCH][VEC], gp[NCH][VEC];
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            float z[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) { z[v] = 0.f; gp[j][v] = 0.f; }
            if (c < nvec) load_vec_ro<VEC>(p.Z + i * dd + c * VEC, z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                y[j][v] = 0.f; xh[j][v] = 0.f;
                if (c < nvec) y[j][v] = act_forward(p.act, z[v], c * VEC + v, xh[j][v]);
            }
            if (p.Y && c < nvec) store_vec<VEC>(p.Y + i * dd + c * VEC, y[j]);
        }
        const float wneg = __ldg(p.inst_w + i) * p.w_scale;
        const float wpos = wneg * p.pos_scale;
        const idx_t* rid = p.ids + i * p.R;

        for (int r0 = 0; r0 < p.R; r0 += RB) {
            float e[RB][NCH][VEC];
            float dot[RB];
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                const bool valid = r0 + rr < p.R;
                const idx_t id = valid ? __ldg(rid + r0 + rr) : 0;
                dot[rr] = 0.f;
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = lane + j * kWarp;
#pragma unroll
                    for (int v = 0; v < VEC; ++v) e[rr][j][v] = 0.f;
                    if (valid && c < nvec) load_vec_ro<VEC>(p.E + id * dd + c * VEC, e[rr][j]);
                }
            }
#pragma unroll
            for (int rr = 0; rr < RB; ++rr)
#pragma unroll
                for (int j = 0; j < NCH; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) dot[rr] += y[j][v] * e[rr][j][v];
            // every lane finishes the scalar chain of ONE row (slot (lane >> 3) & 3)
            float d = warp_sum4_transposed(dot[0], dot[1], dot[2], dot[3], lane);
            const int r = r0 + my_slot;
            float inv_norm = 1.0f;
            if (p.enorm) {   // Normalizer::forward: the row enters the dot product as e / |e|
                float nsq[RB];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    nsq[rr] = 0.f;
#pragma unroll
                    for (int j = 0; j < NCH; ++j)
#pragma unroll
                        for (int v = 0; v < VEC; ++v) nsq[rr] += e[rr][j][v] * e[rr][j][v];
                }
                const float nrm = sqrtf(warp_sum4_transposed(nsq[0], nsq[1], nsq[2], nsq[3], lane));
                inv_norm = 1.0f / nrm;
                d = d / nrm;
                if (r < p.R && slot_leader) {
                    p.enorm[i * p.R + r] = nrm;
                    p.escore[i * p.R + r] = d;
                }
            }
            float coef = 0.f;
            if (r < p.R) {
                const float sign = r == 0 ? 1.0f : -1.0f;
                const float s = sign * d;
                // numerically stable sigmoid (include/cuNVSM/cuda_utils.h:192-214)
                float prob;
                if (s >= 0.f) {
                    prob = 1.0f / (1.0f + expf(-s));
                } else {
                    const float ex = expf(s);
                    prob = ex / (1.0f + ex);
                }
                prob = prob < p.sig_lo_cmp ? p.sig_lo_val : (prob > p.sig_hi_cmp ? p.sig_hi_val : prob);
                const float w = r == 0 ? wpos : wneg;
                const float der = (prob >= p.der_hi_cmp || prob <= p.der_lo_cmp) ? 0.0f : 1.0f - prob;
                const float m = w * (der * p.bsn);
                if (slot_leader) {
                    loss += w * logf(prob);
                    p.probs[i * p.R + r] = prob;
                    p.mult[i * p.R + r] = m;
                }
                coef = (sign * m) * inv_norm;   // d cost / d projection goes through the normalised row
            }
#pragma unroll
            for (int rr = 0; rr < RB; ++rr) {
                const float cf = __shfl_sync(kFull, coef, rr * 8);
#pragma unroll
                for (int j = 0; j < NCH; ++j)
#pragma unroll
                    for (int v = 0; v < VEC; ++v) gp[j][v] += cf * e[rr][j][v];
            }
        }
        // d cost / d (pre-activation); column sums for grad_bias and BN backward.
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            float dy[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                dy[v] = act_deriv(p.act, y[j][v]) * gp[j][v];
                cs[j][v] += dy[v];
                cx[j][v] += dy[v] * xh[j][v];
            }
            if (p.tf32_gp) {
                float lo[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) { const float hi = round_tf32(dy[v]); lo[v] = dy[v] - hi; dy[v] = hi; }
                if (p.Gp_lo && c < nvec) store_vec<VEC>(p.Gp_lo + i * dd + c * VEC, lo);
            }
            if (c < nvec) store_vec<VEC>(p.Gp + i * dd + c * VEC, dy);
        }
    }
    // block reduction of the per-warp partials, then one double atomic per column per block
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c = lane + j * kWarp;
        if (c < nvec) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                atomicAdd(&smem[c * VEC + v], cs[j][v]);
                atomicAdd(&smem[dd + c * VEC + v], cx[j][v]);
            }
        }
    }
    loss = warp_sum(loss);
    if (lane == 0) atomicAdd(&smem[2 * dd], loss);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * dd; t += blockDim.x) atomicAdd(p.col_sums + t, (double)smem[t]);
    if (threadIdx.x == 0) atomicAdd(p.loss_acc, (double)smem[2 * dd]);
    score_sums_tail(p.xchg, p.col_sums, 2 * dd + 1, p.xchg_epoch, 3, p.xchg_counter, p.xchg_error, p.loss_host);
}

// =====================================================================================
// bn_backward: dx = invstd * (dy - sum(dy)/B - xhat * sum(dy*xhat)/B), in place over Gp.
// cudnnBatchNormalizationBackward, per-activation, gamma == 1 (cpp/cudnn_utils.cu:158-177;
// formula pinned by cpp/cudnn_utils_tests.cu:143-176).
// =====================================================================================
// Per-column constants of the backward pass, once per step: grad_bias = sum(dy) (with and
// without batch-norm, cpp/params.cu:510-520) and the two batch means BN backward needs.
// `shifted_by_bias` (nullable): the ring score kernel accumulates sum dy * (xhat + bias); remove
// the bias term here: sum dy * xhat = S2 - bias * S1.
__global__ void bn_backward_prep_kernel(const double* __restrict__ col_sums, int dd, double batch,
                                        const float* __restrict__ shifted_by_bias, float* __restrict__ gb,
                                        float* __restrict__ mean_dy, float* __restrict__ mean_dyx) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dd) return;
    const double s1 = col_sums[c];
    double s2 = col_sums[dd + c];
    if (shifted_by_bias) s2 -= (double)shifted_by_bias[c] * s1;
    gb[c] = (float)s1;
    mean_dy[c] = (float)(s1 / batch);
    mean_dyx[c] = (float)(s2 / batch);
}

template <int VEC>
__global__ void __launch_bounds__(256) bn_backward_kernel(float* __restrict__ Gp, const float* __restrict__ Z,
                                                          const float* __restrict__ mean,
                                                          const float* __restrict__ invstd,
                                                          const float* __restrict__ mean_dy,
                                                          const float* __restrict__ mean_dyx, long rows, int dd,
                                                          int tf32, float* __restrict__ lo_out) {
    const int nvec = dd / VEC;
    const long total = rows * nvec;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t % nvec) * VEC;
        float g[VEC], z[VEC], lo[VEC];
        load_vec<VEC>(Gp + t * VEC, g);
        load_vec_ro<VEC>(Z + t * VEC, z);
#pragma unroll
        for (int v = 0; v < VEC; ++v) {
            const float is = __ldg(invstd + c + v);
            const float xh = (z[v] - __ldg(mean + c + v)) * is;
            g[v] = is * (g[v] - __ldg(mean_dy + c + v) - xh * __ldg(mean_dyx + c + v));
            lo[v] = 0.f;
            if (tf32) {                          // dX only feeds the tensor-core GEMMs
                const float hi = round_tf32(g[v]);
                lo[v] = g[v] - hi;
                g[v] = hi;
            }
        }
        store_vec<VEC>(Gp + t * VEC, g);
        if (lo_out) store_vec<VEC>(lo_out + t * VEC, lo);
    }
}

// Same arithmetic for dd % 4 == 0 with (dd / 4) dividing the block size: a thread keeps ONE float4 column group for
// all of its rows, so the four per-column parameters live in registers instead of being re-read per element
// (ncu on the kernel above: 18 loads per 2 stores, L1TEX 88 % busy, LSU queue throttling, DRAM at 58 %).
// The per-column means of dy and dy * xhat come straight from the (all-reduced) double column sums of the score kernel
// (what bn_backward_prep_kernel computes), and the first row of threads also publishes grad_bias / mean_dy / mean_dyx.
__global__ void __launch_bounds__(256) bn_backward_cols_kernel(float* __restrict__ Gp, const float* __restrict__ Z,
                                                               const float* __restrict__ mean,
                                                               const float* __restrict__ invstd,
                                                               const double* __restrict__ col_sums, double batch,
                                                               const float* __restrict__ shifted_by_bias,
                                                               float* __restrict__ gb, float* __restrict__ mean_dy,
                                                               float* __restrict__ mean_dyx, long rows, int dd,
                                                               int tf32, float* __restrict__ lo_out) {
    pdl_wait();
    pdl_launch_dependents();
    const int nvec = dd >> 2;
    const long tid = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nthreads = (long)gridDim.x * blockDim.x;
    const int c = (int)(tid % nvec);
    const long row_stride = nthreads / nvec;
    const float4 is = __ldg(reinterpret_cast<const float4*>(invstd) + c), mu = __ldg(reinterpret_cast<const float4*>(mean) + c);
    float mdv[4], mxv[4];
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        const int col = 4 * c + v;
        const double s1 = col_sums[col];
        double s2 = col_sums[dd + col];
        if (shifted_by_bias) s2 -= (double)__ldg(shifted_by_bias + col) * s1;
        mdv[v] = (float)(s1 / batch);
        mxv[v] = (float)(s2 / batch);
        if (tid < nvec) { gb[col] = (float)s1; mean_dy[col] = mdv[v]; mean_dyx[col] = mxv[v]; }
    }
    const float4 md = make_float4(mdv[0], mdv[1], mdv[2], mdv[3]), mx = make_float4(mxv[0], mxv[1], mxv[2], mxv[3]);
    float4* __restrict__ G4 = reinterpret_cast<float4*>(Gp);
    const float4* __restrict__ Z4 = reinterpret_cast<const float4*>(Z);
    float4* __restrict__ L4 = reinterpret_cast<float4*>(lo_out);
    auto one = [&](float g, float z, float isv, float muv, float mdv, float mxv, float& lo) {
        const float xh = (z - muv) * isv;
        float r = isv * (g - mdv - xh * mxv);
        lo = 0.f;
        if (tf32) { const float hi = round_tf32(r); lo = r - hi; r = hi; }   // dX only feeds the tensor-core GEMMs
        return r;
    };
    constexpr int U = 4;
    long r = tid / nvec;
    for (; r + (U - 1) * row_stride < rows; r += U * row_stride) {
        float4 g[U], z[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long o = (r + u * row_stride) * nvec + c;
            g[u] = G4[o];
            z[u] = __ldg(Z4 + o);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const long o = (r + u * row_stride) * nvec + c;
            float4 lo, out;
            out.x = one(g[u].x, z[u].x, is.x, mu.x, md.x, mx.x, lo.x); out.y = one(g[u].y, z[u].y, is.y, mu.y, md.y, mx.y, lo.y);
            out.z = one(g[u].z, z[u].z, is.z, mu.z, md.z, mx.z, lo.z); out.w = one(g[u].w, z[u].w, is.w, mu.w, md.w, mx.w, lo.w);
            G4[o] = out;
            if (lo_out) L4[o] = lo;
        }
    }
    for (; r < rows; r += row_stride) {
        const long o = r * nvec + c;
        const float4 g = G4[o], z = __ldg(Z4 + o);
        float4 lo, out;
        out.x = one(g.x, z.x, is.x, mu.x, md.x, mx.x, lo.x); out.y = one(g.y, z.y, is.y, mu.y, md.y, mx.y, lo.y);
        out.z = one(g.z, z.z, is.z, mu.z, md.z, mx.z, lo.z); out.w = one(g.w, z.w, is.w, mu.w, md.w, mx.w, lo.w);
        G4[o] = out;
        if (lo_out) L4[o] = lo;
    }
}

// =====================================================================================
// Split-K partial reduction for grad_transform: out[i] = sum_z part[z][i].
// =====================================================================================
__global__ void reduce_partials_kernel(const float* __restrict__ part, int splits, long n,
                                       float* __restrict__ out) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float s = 0.f;
    for (int z = 0; z < splits; ++z) s += part[(long)z * n + i];
    out[i] = s;
}

// =====================================================================================
// Sparse scatter-add (replaces update_repr_kernel, cpp/storage.cu:37-49). The gradient of
// the entity table is never materialised: grad_entity[:, c] = +-mult[c] * y[i(c)]
// (cpp/objective.cu:381-401) is recomputed from Z on the fly.
//   target[id[i,r], :] += scale * factor(c) * (+-mult[c]) * y[i, :]
// factor(c) = 1/sqrt(acc[id] + eps) for Adagrad (adagrad_update_kernel,
// cpp/updates_adagrad.cu:83-97, window 1), else 1.
// =====================================================================================
struct EntityScatterParams {
    const float* Z;
    ActParams act;
    const idx_t* ids;
    const float* mult;
    long B;
    int R, dd;
    float* target;
    float scale;
    const float* acc;  // nullable
    float eps;
};

template <int VEC>
__global__ void __launch_bounds__(256) entity_scatter_kernel(const EntityScatterParams p) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = p.dd / VEC;
    for (long i = warp0; i < p.B; i += nwarps) {
        for (int c = lane; c < nvec; c += kWarp) {
            float z[VEC], y[VEC];
            load_vec_ro<VEC>(p.Z + i * p.dd + c * VEC, z);
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                float xh;
                y[v] = act_forward(p.act, z[v], c * VEC + v, xh);
            }
            for (int r = 0; r < p.R; ++r) {
                const long col = i * p.R + r;
                const idx_t id = __ldg(p.ids + col);
                float coef = p.scale * __ldg(p.mult + col);
                if (r != 0) coef = -coef;
                if (p.acc) coef = coef / sqrtf(__ldg(p.acc + id) + p.eps);
                float g[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) g[v] = coef * y[v];
                red_add_vec<VEC>(p.target + id * p.dd + c * VEC, g);
            }
        }
    }
}

//   target[id[i,w], :] += scale * factor(i) * fw[i,w] * G[i, :]
// (update_repr_kernel with window n and per-word weights; factor(i) =
// 1/sqrt(mean_w acc[id[i,w]] + eps) for Adagrad.)
struct WordScatterParams {
    const float* G;  // [B, dw]
    const idx_t* ids;
    const float* fw;
    long B;
    int n, dw;
    float* target;
    float scale;
    const float* acc;  // nullable
    float eps;
};

template <int VEC>
__global__ void __launch_bounds__(256) word_scatter_kernel(const WordScatterParams p) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = p.dw / VEC;
    for (long i = warp0; i < p.B; i += nwarps) {
        const idx_t* wid = p.ids + i * p.n;
        const float* ww = p.fw + i * p.n;
        float factor = 1.0f;
        if (p.acc) {
            float a = 0.f;
            for (int w = 0; w < p.n; ++w) a += __ldg(p.acc + __ldg(wid + w));
            a /= (float)p.n;
            factor = 1.0f / sqrtf(a + p.eps);
        }
        for (int c = lane; c < nvec; c += kWarp) {
            float g[VEC];
            load_vec_ro<VEC>(p.G + i * p.dw + c * VEC, g);
#pragma unroll
            for (int v = 0; v < VEC; ++v) g[v] *= factor;
#pragma unroll 5
            for (int w = 0; w < p.n; ++w) {
                const idx_t id = __ldg(wid + w);
                const float coef = p.scale * __ldg(ww + w);
                float u[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) u[v] = coef * g[v];
                red_add_vec<VEC>(p.target + id * p.dw + c * VEC, u);
            }
        }
    }
}

// =====================================================================================
// Per-object scalar second moments (Adagrad, sparse / dense-update Adam):
//   acc[id] += scale * wt * mean_k grad[k, col]^2
// (reduce_axis<square> + scale by 1/rows, then update_repr_kernel with one-thread blocks:
// cpp/updates_adagrad.cu:130-158, cpp/updates_adam.cu:215-250.)
// =====================================================================================
// ysq[i] = mean_k y[i,k]^2 (entity side: mean_k grad_entity[k,c]^2 = mult[c]^2 * ysq[i]).
__global__ void __launch_bounds__(256) row_meansq_act_kernel(const float* __restrict__ Z, const ActParams act,
                                                             long rows, int dd, float inv_dim,
                                                             float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long i = warp0; i < rows; i += nwarps) {
        float s = 0.f;
        for (int c = lane; c < dd; c += kWarp) {
            float xh;
            const float y = act_forward(act, __ldg(Z + i * dd + c), c, xh);
            s += y * y;
        }
        s = warp_sum(s);
        if (lane == 0) out[i] = s * inv_dim;
    }
}

__global__ void __launch_bounds__(256) row_meansq_kernel(const float* __restrict__ G, long rows, int dim,
                                                         float inv_dim, float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long i = warp0; i < rows; i += nwarps) {
        float s = 0.f;
        for (int c = lane; c < dim; c += kWarp) {
            const float g = __ldg(G + i * dim + c);
            s += g * g;
        }
        s = warp_sum(s);
        if (lane == 0) out[i] = s * inv_dim;
    }
}

// With entity normalisation the column is (mult / |e|) * (y - s e / |e|), s = e.y / |e| (mult already holds
// mult / |e|): mean_k of its square = mult^2 * (mean_k y^2 - s^2 / dd).
// acc[id] += v for every thread of the block (id < 0: nothing), with the block's contributions to one id combined
// in shared memory first: a Zipfian id stream otherwise lands tens of thousands of atomics on one address.
// Call with all threads of a kAggThreads block.
__device__ __forceinline__ void block_aggregated_add(float* __restrict__ acc, int id, float v) {
    __shared__ int s_id[kAggSlots];
    __shared__ float s_sum[kAggSlots];
    for (int i = threadIdx.x; i < kAggSlots; i += blockDim.x) { s_id[i] = -1; s_sum[i] = 0.f; }
    __syncthreads();
    if (id >= 0) {
        const int slot = agg_slot(id);
        const int prev = atomicCAS(&s_id[slot], -1, id);
        if (prev == -1 || prev == id) atomicAdd(&s_sum[slot], v);
        else atomicAdd(acc + id, v);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAggSlots; i += blockDim.x)
        if (s_id[i] >= 0) atomicAdd(acc + s_id[i], s_sum[i]);
}

__global__ void entity_scalar_scatter_kernel(const idx_t* __restrict__ ids, const float* __restrict__ mult,
                                             const float* __restrict__ ysq, long total, int R, float scale,
                                             float* __restrict__ acc, const float* __restrict__ escore,
                                             float inv_dim) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float v = 0.f;
    if (c < total) {
        const float m = mult[c];
        float q = ysq[c / R];
        if (escore) { const float sc = escore[c]; q = fmaxf(q - sc * sc * inv_dim, 0.f); }
        v = scale * (m * m * q);
    }
    block_aggregated_add(acc, c < total ? (int)ids[c] : -1, v);
}

// =====================================================================================
// L2 Normalizer (cpp/cuda_utils.cu:3-141; golden vectors cpp/cuda_utils_tests.cu:51-92).
//   forward : out[:, c] = in[:, c] / |in[:, c]|                          (per instance = per row here)
//   backward: gin = gout / n - x (x . gout) / n^3 = (gout - xhat (xhat . gout)) / n
// =====================================================================================
// Phrase side, forward: rows of P (hi + lo when the tensor-core split is on) are normalised in place.
__global__ void __launch_bounds__(256) row_l2_normalize_kernel(float* __restrict__ P, float* __restrict__ P_lo,
                                                               long rows, int dim, int ld, int tf32,
                                                               float* __restrict__ norms) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long i = warp0; i < rows; i += nwarps) {
        float s = 0.f;
        for (int c = lane; c < dim; c += kWarp) {
            const float x = P[i * ld + c] + (P_lo ? P_lo[i * ld + c] : 0.f);
            s += x * x;
        }
        const float nrm = sqrtf(warp_sum(s));
        if (lane == 0) norms[i] = nrm;
        for (int c = lane; c < dim; c += kWarp) {
            float x = (P[i * ld + c] + (P_lo ? P_lo[i * ld + c] : 0.f)) / nrm;
            float lo = 0.f;
            if (tf32) { const float hi = round_tf32(x); lo = x - hi; x = hi; }
            P[i * ld + c] = x;
            if (P_lo) P_lo[i * ld + c] = lo;
        }
    }
}

// Phrase side, backward, in place over grad_phrase [rows, dim] (xhat = the normalised P kept by the forward).
__global__ void __launch_bounds__(256) row_l2_normalize_backward_kernel(float* __restrict__ G, int ldg,
                                                                        const float* __restrict__ P,
                                                                        const float* __restrict__ P_lo, int ld,
                                                                        const float* __restrict__ norms, long rows,
                                                                        int dim) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long i = warp0; i < rows; i += nwarps) {
        float s = 0.f;
        for (int c = lane; c < dim; c += kWarp) {
            const float xh = P[i * ld + c] + (P_lo ? P_lo[i * ld + c] : 0.f);
            s += xh * G[i * ldg + c];
        }
        const float dot = warp_sum(s);
        const float nrm = norms[i];
        for (int c = lane; c < dim; c += kWarp) {
            const float xh = P[i * ld + c] + (P_lo ? P_lo[i * ld + c] : 0.f);
            G[i * ldg + c] = (G[i * ldg + c] - xh * dot) / nrm;
        }
    }
}

// Entity side. Per reference c = (i, r) with row d = ids[c], n = |E_d|, s = E_d . y_i / n:
//   grad column = (+-mult / n) * y_i  -  (+-mult * s / n^2) * E_d
// The first term is the ordinary scatter with mult_eff = mult / n; the second is a per-row multiple of the row
// itself: kself[d] = sum_c +-mult_c s_c / n_c^2, applied by the *_self kernels below against the pre-update E.
__global__ void entity_norm_prep_kernel(const idx_t* __restrict__ ids, const float* __restrict__ mult,
                                        const float* __restrict__ enorm, const float* __restrict__ escore,
                                        long total, int R, float* __restrict__ mult_eff,
                                        float* __restrict__ kself) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    const float n = enorm[c], m = mult[c];
    mult_eff[c] = m / n;
    const float signed_m = (c % R) != 0 ? -m : m;
    atomicAdd(kself + ids[c], signed_m * escore[c] / (n * n));
}

// theta[d, :] *= decay - lr * kself[d] * (acc ? 1 / sqrt(acc[d] + eps) : 1)   (dense decay + self term of SGD / Adagrad)
__global__ void __launch_bounds__(256) scale_rows_self_kernel(float* __restrict__ theta, long num_rows, int dim,
                                                              float decay, float lr, const float* __restrict__ kself,
                                                              const float* __restrict__ acc, float eps) {
    const long total = num_rows * dim;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long d = t / dim;
        float k = kself[d];
        if (acc) k = k / sqrtf(acc[d] + eps);
        theta[t] = theta[t] * decay - (lr * k) * theta[t];
    }
}

// target[d, :] += coef * kself[d] * source[d, :]
__global__ void __launch_bounds__(256) row_self_axpy_kernel(float* __restrict__ target, const float* __restrict__ source,
                                                            long num_rows, int dim, float coef,
                                                            const float* __restrict__ kself) {
    const long total = num_rows * dim;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x)
        target[t] += (coef * kself[t / dim]) * source[t];
}

__global__ void word_scalar_scatter_kernel(const idx_t* __restrict__ ids, const float* __restrict__ fw,
                                           const float* __restrict__ msq, long total, int n, float scale,
                                           float* __restrict__ acc) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    block_aggregated_add(acc, c < total ? (int)ids[c] : -1, c < total ? scale * fw[c] * msq[c / n] : 0.f);
}

// =====================================================================================
// Dense helpers.
// =====================================================================================
__global__ void __launch_bounds__(256) scale_kernel(float* __restrict__ x, long n, float s) {
    const long stride = (long)gridDim.x * blockDim.x;
    const long n4 = n >> 2;
    float4* x4 = reinterpret_cast<float4*>(x);
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 v = x4[i];
        v.x *= s; v.y *= s; v.z *= s; v.w *= s;
        x4[i] = v;
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) x[i] *= s;
}

// Sparse Adam step for the entity table (window 1), cpp/updates_adam.cu:132-151,339-385:
//   E[id_c, :] += lr * bc * m[id_c, :] / (sqrt(v[id_c]) + eps)      for every column c.
template <int VEC>
__global__ void __launch_bounds__(256) adam_sparse_entity_kernel(const idx_t* __restrict__ ids, long total, int dd,
                                                                 const float* __restrict__ m,
                                                                 const float* __restrict__ v, float bc,
                                                                 float eps, float lr,
                                                                 float* __restrict__ E) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dd / VEC;
    for (long c = warp0; c < total; c += nwarps) {
        const idx_t id = __ldg(ids + c);
        const float den = sqrtf(__ldg(v + id)) + eps;
        for (int k = lane; k < nvec; k += kWarp) {
            float mm[VEC], u[VEC];
            load_vec<VEC>(m + id * dd + k * VEC, mm);
#pragma unroll
            for (int q = 0; q < VEC; ++q) u[q] = lr * (bc * mm[q] / den);
            red_add_vec<VEC>(E + id * dd + k * VEC, u);
        }
    }
}

// Sparse Adam "gradient" for the word table (window n): overwrites G[i, :] with
//   bc * mean_w m[id_w, :] / (sqrt(mean_w v[id_w]) + eps)      (adam_sparse_update_kernel)
template <int VEC>
__global__ void __launch_bounds__(256) adam_sparse_word_grad_kernel(const idx_t* __restrict__ ids, long B, int n,
                                                                    int dw, const float* __restrict__ m,
                                                                    const float* __restrict__ v, float bc,
                                                                    float eps, float* __restrict__ G) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dw / VEC;
    const float fn = (float)n;
    for (long i = warp0; i < B; i += nwarps) {
        const idx_t* wid = ids + i * n;
        float av = 0.f;
        for (int w = 0; w < n; ++w) av += __ldg(v + __ldg(wid + w));
        av /= fn;
        const float den = sqrtf(av) + eps;
        for (int k = lane; k < nvec; k += kWarp) {
            float am[VEC];
#pragma unroll
            for (int q = 0; q < VEC; ++q) am[q] = 0.f;
            for (int w = 0; w < n; ++w) {
                float mm[VEC];
                load_vec<VEC>(m + __ldg(wid + w) * dw + k * VEC, mm);
#pragma unroll
                for (int q = 0; q < VEC; ++q) am[q] += mm[q];
            }
#pragma unroll
            for (int q = 0; q < VEC; ++q) am[q] = bc * (am[q] / fn) / den;
            store_vec<VEC>(G + i * dw + k * VEC, am);
        }
    }
}

// DENSE_UPDATE Adam (cpp/updates_adam.cu:286-309): per-object scalar v.
//   theta = theta * (1 - lambda*lr) + lr * bc * m / (sqrt(v[obj]) + eps)
__global__ void __launch_bounds__(256) adam_dense_update_kernel(float* __restrict__ theta, const float* __restrict__ m,
                                                                const float* __restrict__ v, long num_objects,
                                                                int dim, float decay, float lr, float bc,
                                                                float eps) {
    const long total = num_objects * dim;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const float den = sqrtf(__ldg(v + t / dim)) + eps;
        theta[t] = theta[t] * decay + ((m[t] / den) * bc) * lr;
    }
}

// DENSE_UPDATE_DENSE_VARIANCE ("full_adam", cpp/updates_adam.cu:199-213,251-283,310-328),
// one fused pass over theta / m / v / agg (agg = scatter-added gradient of this step):
//   g  = agg - lambda * theta
//   m  = s1 * m + lr1 * agg - (lr1' * lambda) * theta      (== s1*m + (1-b1) * g)
//   v  = s2 * v + lr2 * g^2
//   theta += lr * bc * m / (sqrt(v) + eps);   agg = 0 (ready for the next step)
__global__ void __launch_bounds__(256) adam_full_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                        float* __restrict__ v, float* __restrict__ agg, long n,
                                                        float s1, float lr1, float reg1, float s2, float lr2,
                                                        float lambda, float lr, float bc, float eps) {
    const long stride = (long)gridDim.x * blockDim.x;
    const long n4 = n >> 2;
    float4* t4 = reinterpret_cast<float4*>(theta);
    float4* m4 = reinterpret_cast<float4*>(m);
    float4* v4 = reinterpret_cast<float4*>(v);
    float4* a4 = reinterpret_cast<float4*>(agg);
    auto one = [&](float& th, float& mm, float& vv, float& ag) {
        const float g = ag + (-lambda * th);
        mm = (mm * s1 + lr1 * ag) + (-reg1 * th);
        vv = vv * s2 + (g * g) * lr2;
        th = th + ((mm / (sqrtf(vv) + eps)) * bc) * lr;
        ag = 0.f;
    };
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
        float4 th = t4[i], mm = m4[i], vv = v4[i], ag = a4[i];
        one(th.x, mm.x, vv.x, ag.x);
        one(th.y, mm.y, vv.y, ag.y);
        one(th.z, mm.z, vv.z, ag.z);
        one(th.w, mm.w, vv.w, ag.w);
        t4[i] = th; m4[i] = mm; v4[i] = vv; a4[i] = ag;
    }
    for (long i = (n4 << 2) + (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        one(theta[i], m[i], v[i], agg[i]);
}

// Dense optimiser for the projection: T [dw*dd] followed by b [dd] in one launch.
// SGD: cpp/storage.cu:198-228; Adagrad: cpp/updates_adagrad.cu:33-70; Adam:
// cpp/updates_adam.cu:46-105 — the bias is never regularised and its Adam moments never
// decay (pinned by cpp/updates_tests.cu:352-366,409-423).
struct TransformUpdateParams {
    float* T; float* b;
    const float* gT; const float* gb;
    long nT; int nb;
    int method;      // 0 sgd, 1 adagrad, 2 adam
    float lr, lambda;
    float* aT; float* ab;   // adagrad acc / adam m
    float* vT; float* vb;   // adam v
    float s1, lr1, s2, lr2, bc, eps;
    // fused in (single GPU): the split-K partial reduction of grad_transform (gT_out = sum_z part[z]) ...
    const float* gT_part; int nparts; float* gT_out;
    // ... and the tensor-core operand copies of the NEW T for the next step's GEMMs: Tr = rn_tf32(T) [dw, dd],
    // Tt = its transpose [dd, ldT] (K-major B operand of the forward GEMM) and the 3xTF32 remainders
    float* Tr; float* Tt; float* Tr_lo; float* Tt_lo; int dd; int ldT;
    // N > 1, fused grad_transform exchange (gt_reduce_push_kernel): gT_part = this rank's gT inbox of the running parity
    // ([nranks][nT], written by the peers over NVLink), nparts = nranks; before reading element k wait until the flag
    // of chunk k / xchg_chunk from every rank has reached xchg_epoch. Null: off.
    const PeerXchg* xchg; unsigned long long xchg_epoch; int xchg_chunk; int* xchg_error;
};

__global__ void __launch_bounds__(256) transform_update_kernel(const TransformUpdateParams p) {
    const long total = p.nT + p.nb;
    if (p.xchg) {
        // Fused grad_transform exchange: the elements of this block (one grid-stride pass: the grid covers `total`) live in
        // at most `span` consecutive chunks; thread (c, r) polls the flag of chunk c from rank r, one system-scope fence
        // per polling thread, then the block goes on (per-thread acquire loads measured +11 us on the step).
        const PeerXchg& x = *p.xchg;
        const int parity = (int)(p.xchg_epoch & 1ull);
        const long e0 = (long)blockIdx.x * blockDim.x, e1 = min(p.nT, e0 + (long)blockDim.x) - 1;
        if (e0 < p.nT) {
            const int c0 = (int)(e0 / p.xchg_chunk), span = (int)(e1 / p.xchg_chunk) - c0 + 1;
            for (int t = threadIdx.x; t < span * x.nranks; t += blockDim.x) {
                const int c = c0 + t / x.nranks, r = t % x.nranks;
                const volatile unsigned long long* f = x.flags[x.rank] + peer_flag_index(x, kPeerKindGt, parity, r, c);
                long spins = 0;
                while (*f < p.xchg_epoch)
                    if (++spins > (1L << 26)) { atomicExch(p.xchg_error, 1); break; }
                __threadfence_system();
            }
        }
        __syncthreads();
    }
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const bool is_bias = t >= p.nT;
        const long k = is_bias ? t - p.nT : t;
        float* th = is_bias ? p.b + k : p.T + k;
        float g;
        if (is_bias) {
            g = p.gb[k];
        } else if (p.gT_part && p.xchg) {
            g = 0.f;
            for (int r = 0; r < p.nparts; ++r) g += __ldcg(p.gT_part + (long)r * p.nT + k);   // rank order, past L1
            p.gT_out[k] = g;
        } else if (p.gT_part) {
            // same summation order as reduce_partials_kernel; eight independent loads in flight
            g = 0.f;
            int z = 0;
            for (; z + 8 <= p.nparts; z += 8) {
                float x[8];
#pragma unroll
                for (int u = 0; u < 8; ++u) x[u] = __ldg(p.gT_part + (long)(z + u) * p.nT + k);
#pragma unroll
                for (int u = 0; u < 8; ++u) g += x[u];
            }
            for (; z < p.nparts; ++z) g += __ldg(p.gT_part + (long)z * p.nT + k);
            p.gT_out[k] = g;
        } else {
            g = p.gT[k];
        }
        const float lam = is_bias ? 0.f : p.lambda;
        if (p.method == 0) {
            *th = *th * (1.0f - lam * p.lr) + g * p.lr;
        } else if (p.method == 1) {
            float* a = is_bias ? p.ab + k : p.aT + k;
            const float acc = *a + g * g;
            *a = acc;
            g = g / sqrtf(acc + p.eps);
            *th = *th * (1.0f - lam * p.lr) + g * p.lr;
        } else {
            float* mm = is_bias ? p.ab + k : p.aT + k;
            float* vv = is_bias ? p.vb + k : p.vT + k;
            g = g + (-lam * *th);
            const float m = *mm * (is_bias ? 1.0f : p.s1) + g * p.lr1;
            const float v = *vv * (is_bias ? 1.0f : p.s2) + (g * g) * p.lr2;
            *mm = m;
            *vv = v;
            g = (m * p.bc) / (sqrtf(v) + p.eps);
            *th = *th + g * p.lr;
        }
        if (!is_bias && p.Tr) {
            const float raw = *th, hi = round_tf32(raw);
            const long r = k / p.dd, c = k - r * p.dd;
            p.Tr[k] = hi;
            p.Tt[c * p.ldT + r] = hi;
            if (p.Tr_lo) { p.Tr_lo[k] = raw - hi; p.Tt_lo[c * p.ldT + r] = raw - hi; }
        }
    }
}

// Dense-ify grad_entity for inspection: out[c, :] = +-mult[c] * y[i(c), :].
// With entity normalisation (enorm != null): +-(mult / n) * (y - s E_d / n).
__global__ void __launch_bounds__(256) materialize_grad_entity_kernel(const float* __restrict__ Z, const ActParams act,
                                                                      const float* __restrict__ mult, long total_cols,
                                                                      int R, int dd, float* __restrict__ out,
                                                                      const float* __restrict__ E,
                                                                      const idx_t* __restrict__ ids,
                                                                      const float* __restrict__ enorm,
                                                                      const float* __restrict__ escore) {
    const long total = total_cols * dd;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const long c = t / dd;
        const int k = (int)(t % dd);
        const long i = c / R;
        float xh;
        float y = act_forward(act, Z[i * dd + k], k, xh);
        float m = mult[c];
        if (enorm) {
            const float n = enorm[c];
            y = y - escore[c] * E[ids[c] * dd + k] / n;
            m = m / n;
        }
        const float g = y * m;
        out[t] = (c % R) != 0 ? -g : g;
    }
}

// Y = f(BN?(Z)) for inspection / inference.
__global__ void __launch_bounds__(256) materialize_activation_kernel(const float* __restrict__ Z, const ActParams act,
                                                                     long rows, int dd, float* __restrict__ out) {
    const long total = rows * dd;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        float xh;
        out[t] = act_forward(act, Z[t], (int)(t % dd), xh);
    }
}

__global__ void increment_kernel(float* p, float eps) { *p += eps; }

// Ids outside their table would gather from and update memory that is not theirs (the reference only DCHECKs them,
// include/cuNVSM/storage.h). One pass right behind the upload of a batch clamps such ids to row 0 and raises a flag
// in mapped host memory (written only when something is wrong) that the host reports at its next synchronising call.
__global__ void validate_ids_kernel(idx_t* __restrict__ words, long num_words, long word_limit,
                                    idx_t* __restrict__ entities, long num_entities, long entity_limit,
                                    volatile int* __restrict__ host_flags) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < num_words + num_entities; i += stride) {
        const bool is_word = i < num_words;
        idx_t* const p = is_word ? words + i : entities + (i - num_words);
        const idx_t v = *p;
        if (v < 0 || v >= (is_word ? word_limit : entity_limit)) {
            *p = 0;
            host_flags[is_word ? 0 : 1] = 1;
            __threadfence_system();
        }
    }
}

// x -> rn_tf32(x) in place, lo = x - rn_tf32(x) (test hook for the 3xTF32 GEMM).
__global__ void split_tf32_kernel(float* __restrict__ x, float* __restrict__ lo, long n) {
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long)gridDim.x * blockDim.x) {
        const float v = x[i], hi = round_tf32(v);
        x[i] = hi;
        lo[i] = v - hi;
    }
}

}  // namespace nvsm
