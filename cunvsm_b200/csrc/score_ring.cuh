// score_ring_kernel — the fused score / NCE loss / backward-through-dot kernel (see
// kernels.cuh:score_kernel for the arithmetic and the reference citations) with the row traffic
// moved off the register file: every warp owns `S` shared-memory stages (S = 1 by default, see try_score_ring), one stage = the
// pre-activation row Z[i] plus the R entity rows E[id[i,0..R)] of one n-gram, filled with 16-byte
// cp.async (LDGSTS) copies, one commit group per n-gram. While the warp works on n-gram j from
// stage j % S, the copies of n-grams j+1 .. j+S-1 are in flight, and the ids / instance weight of
// n-gram j+S are being fetched into registers. No register is spent on memory latency and the
// gather no longer depends on occupancy. (A cp.async.bulk-per-row variant was measured first: at
// 1 KB per request it is bound by the TMA unit's request rate, ~70 cycles per row per SM.)
#pragma once

#include "common.cuh"
#include "kernels.cuh"

namespace nvsm {

__device__ __forceinline__ uint32_t ring_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void ring_bulk_copy(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
                 : "memory");
}

__device__ __forceinline__ void ring_cp16(uint32_t dst, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}

__device__ __forceinline__ void ring_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 26); ++it) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(bar), "r"(parity)
            : "memory");
        if (ok) return;
    }
    __trap();
}

struct ScoreRingParams {
    ScoreParams s;
    int stages;          // S
    int warps;           // warps per block
    const float* bn_scale;   // [dd] invstd            (batch-norm only)
    const float* bn_shift;   // [dd] bias - mean*invstd
};

// Requires dd % 4 == 0 (16-byte rows) and R <= 32. FULL: dd == NCH * 128 (no column predicates).
//
// Per n-gram, all from shared memory:
//   pass 1  partial dot products of y with the R staged rows, four rows per transposing butterfly
//           (6 shuffles), result of row r parked in lane r;
//   chain   ONE sigmoid / clamp / log / multiplier evaluation for all R rows (lane r = row r).
//           exp / log / divide use the fast intrinsics: the reference's release build is compiled
//           with -use_fast_math (CMakeLists.txt:71-73);
//   pass 2  Gp += coef_r * E_r, re-reading the rows from shared memory.
template <int NCH, bool FULL>
__global__ void __launch_bounds__(256) score_ring_kernel(const ScoreRingParams q) {
    constexpr int VEC = 4;
    const ScoreParams& p = q.s;
    extern __shared__ __align__(128) uint8_t ring_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int dd = FULL ? NCH * 128 : p.dd;   // compile-time row length on the FULL path: immediate offsets, shifts
    const int R = p.R, S = q.stages, W = q.warps;
    const int nvec = dd / VEC;
    const uint32_t row_bytes = (uint32_t)dd * 4u;
    const uint32_t stage_bytes = (uint32_t)(R + 1) * row_bytes;   // row 0 = Z[i], rows 1..R = E rows
    const uint32_t stage_floats = stage_bytes / 4;
    // layout: [W][S] stages | [2*dd] BN scale/shift | [2*dd + 1] column sums + loss | [W][S] barriers
    float* stages = reinterpret_cast<float*>(ring_raw);
    float* bnp = stages + (size_t)W * S * stage_floats;
    float* sums = bnp + 2 * dd;
    uint64_t* bars = reinterpret_cast<uint64_t*>(sums + 2 * dd + 2);

    for (int t = threadIdx.x; t < 2 * dd + 1; t += blockDim.x) sums[t] = 0.f;
    pdl_wait();                // (programmatic dependent of the statistics kernel: nothing global is touched above)
    pdl_launch_dependents();
    if (p.act.use_bn)
        for (int t = threadIdx.x; t < dd; t += blockDim.x) { bnp[t] = q.bn_scale[t]; bnp[dd + t] = q.bn_shift[t]; }
    __syncthreads();

    float* my_stages = stages + (size_t)w * S * stage_floats;
    const int lane4 = lane * VEC;
    (void)bars;
    const long warp0 = (long)blockIdx.x * W + w;
    const long nwarps = (long)gridDim.x * W;
    const long my_count = warp0 < p.B ? (p.B - warp0 + nwarps - 1) / nwarps : 0;   // n-grams of this warp

    // id of row `lane` and instance weight of an upcoming n-gram
    // (ids are validated against the table size behind every upload: 32 bits hold a row index, one shuffle moves it)
    auto fetch_meta = [&](long j, uint32_t& id0, float& iw) {
        id0 = 0; iw = 0.f;
        if (j < my_count) {
            const long i = warp0 + j * nwarps;
            if (lane < R) id0 = (uint32_t)__ldg(p.ids + i * R + lane);
            iw = __ldg(p.inst_w + i);
        }
    };
    // One commit group per n-gram (an empty group past the end keeps the group arithmetic uniform). s = stage.
    // (ncu r2v: the issue sequence was 15 % of the kernel's instructions -- 64-bit id shuffles, a 64-bit j % S -- on a
    // kernel whose issue slots are 61 % busy.)
    const float* const e_lane = p.E + lane4;
    auto issue = [&](long j, uint32_t id0, int s) {
        if (j < my_count) {
            const long i = warp0 + j * nwarps;
            const uint32_t dst = ring_smem_u32(my_stages) + (uint32_t)s * stage_bytes + (uint32_t)lane4 * 4u;
            const float* zsrc = p.Z + i * dd + lane4;
#pragma unroll
            for (int k = 0; k < NCH; ++k)
                if (FULL || lane4 + k * 128 < dd) ring_cp16(dst + k * 512u, zsrc + k * 128);
            uint32_t d = dst + row_bytes;
            for (int r0 = 0; r0 < R; r0 += 4, d += 4 * row_bytes) {   // groups of four, predicated: no remainder loop
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const uint32_t id = __shfl_sync(kFull, id0, (r0 + rr) & 31);
                    if (r0 + rr < R) {
                        const float* src = e_lane + (size_t)id * dd;
#pragma unroll
                        for (int k = 0; k < NCH; ++k)
                            if (FULL || lane4 + k * 128 < dd) ring_cp16(d + (uint32_t)rr * row_bytes + k * 512u, src + k * 128);
                    }
                }
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };

    uint32_t nid0;
    float niw;
    float iw_q[4] = {0.f, 0.f, 0.f, 0.f};   // instance weights of n-grams j .. j+S-1 (S <= 4), rotated
#pragma unroll
    for (int jj = 0; jj < 3; ++jj) {
        if (jj < S - 1) {
            fetch_meta(jj, nid0, niw);
            iw_q[jj] = niw;
            issue(jj, nid0, jj);
        }
    }
    fetch_meta(S - 1, nid0, niw);
    int s = 0;   // stage of n-gram j (j % S, kept incrementally)

    float cs[NCH][VEC], cx[NCH][VEC];
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int v = 0; v < VEC; ++v) { cs[j][v] = 0.f; cx[j][v] = 0.f; }
    float loss = 0.f;
    const bool use_bn = p.act.use_bn != 0;
    const bool is_tanh = p.act.nonlinearity == 0;
    const float clip_min = p.act.clip_min, clip_max = p.act.clip_max;
    const int nbatch = (R + 3) >> 2;

    long i = warp0 - nwarps;
    for (long j = 0; j < my_count; ++j) {
        i += nwarps;   // n-gram of iteration j: warp0 + j * nwarps
        // keep the ring full: n-gram j+S-1 goes into the stage consumed at iteration j-1
        if (S == 1) iw_q[0] = niw; else if (S == 2) iw_q[1] = niw; else if (S == 3) iw_q[2] = niw; else iw_q[3] = niw;
        issue(j + S - 1, nid0, s == 0 ? S - 1 : s - 1);
        fetch_meta(j + S, nid0, niw);

        // groups j .. j+S-1 are outstanding: wait until at most S-1 remain, then make every lane's
        // copies visible to the whole warp
        if (S == 1) asm volatile("cp.async.wait_group 0;" ::: "memory");
        else if (S == 2) asm volatile("cp.async.wait_group 1;" ::: "memory");
        else if (S == 3) asm volatile("cp.async.wait_group 2;" ::: "memory");
        else asm volatile("cp.async.wait_group 3;" ::: "memory");
        __syncwarp();
        const float* st = my_stages + (size_t)s * stage_floats;

        float y[NCH][VEC], tt[NCH][VEC];
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c4 = lane4 + k * 128;
            const bool ok = FULL || c4 < dd;
            float4 z = make_float4(0.f, 0.f, 0.f, 0.f), a = make_float4(1.f, 1.f, 1.f, 1.f), b = make_float4(0.f, 0.f, 0.f, 0.f);
            if (ok) {
                z = *reinterpret_cast<const float4*>(st + c4);
                if (use_bn) {
                    a = *reinterpret_cast<const float4*>(bnp + c4);
                    b = *reinterpret_cast<const float4*>(bnp + dd + c4);
                }
            }
            tt[k][0] = fmaf(z.x, a.x, b.x); tt[k][1] = fmaf(z.y, a.y, b.y);   // BN: (z - mean) * invstd + bias
            tt[k][2] = fmaf(z.z, a.z, b.z); tt[k][3] = fmaf(z.w, a.w, b.w);
#pragma unroll
            for (int v = 0; v < VEC; ++v)
                y[k][v] = ok ? (is_tanh ? tanhf(tt[k][v]) : fminf(fmaxf(tt[k][v], clip_min), clip_max)) : 0.f;
            if (p.Y && ok) store_vec<VEC>(p.Y + i * dd + c4, y[k]);
        }

        // ---- one pass over the staged rows, four at a time: dot products (transposing butterfly, 6 shuffles), the
        // sigmoid / clamp / log / multiplier chain of THESE four rows (lane l evaluates row 4 bch + (l & 3); the chain of
        // a row only depends on its own dot product), and Gp += coef_r * E_r while the rows are still in registers.
        // The former layout (all dot products, ONE chain with lane r = row r, then a second sweep over the rows for Gp)
        // read every staged row twice: 35 KB of shared-memory traffic per n-gram against 24 KB now, on a kernel whose
        // shared-memory time (1.8 GB per launch / 36 TB/s) was half of its duration. Same arithmetic per row and same
        // accumulation order of Gp; the loss is still accumulated by lane r = row r.
        float gp[NCH][VEC];
#pragma unroll
        for (int k = 0; k < NCH; ++k)
#pragma unroll
            for (int v = 0; v < VEC; ++v) gp[k][v] = 0.f;
        const float* erow = st + dd;     // first entity row
        const float wneg = iw_q[0] * p.w_scale;
        float my_prob = 1.f, my_mult = 0.f, my_wgt = 0.f;
        // Branch-free: rows past R are clamped to row R-1 and get coefficient zero.
        for (int bch = 0; bch < nbatch; ++bch) {
            float4 x[4][NCH];
            const float* rp0 = erow + (size_t)(bch * 4) * dd + lane4;
            if (bch * 4 + 3 < R) {   // a full group: offsets are immediates on the FULL path
#pragma unroll
                for (int rr = 0; rr < 4; ++rr)
#pragma unroll
                    for (int k = 0; k < NCH; ++k)
                        x[rr][k] = (FULL || lane4 + k * 128 < dd) ? *reinterpret_cast<const float4*>(rp0 + rr * dd + k * 128)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) {
                    const float* rp = erow + (size_t)min(bch * 4 + rr, R - 1) * dd + lane4;
#pragma unroll
                    for (int k = 0; k < NCH; ++k)
                        x[rr][k] = (FULL || lane4 + k * 128 < dd) ? *reinterpret_cast<const float4*>(rp + k * 128)
                                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
            float dot[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                dot[rr] = 0.f;
#pragma unroll
                for (int k = 0; k < NCH; ++k) {   // (one FMA per element; the four rows are independent chains)
                    dot[rr] = fmaf(y[k][0], x[rr][k].x, dot[rr]); dot[rr] = fmaf(y[k][1], x[rr][k].y, dot[rr]);
                    dot[rr] = fmaf(y[k][2], x[rr][k].z, dot[rr]); dot[rr] = fmaf(y[k][3], x[rr][k].w, dot[rr]);
                }
            }
            const float d = warp_sum4_transposed(dot[0], dot[1], dot[2], dot[3], lane);   // slot (lane>>3)&3
            const float my_dot = __shfl_sync(kFull, d, (lane & 3) * 8);                  // row 4 bch + (lane & 3)
            float coef = 0.f;
            {
                const int r = bch * 4 + (lane & 3);
                const float sign = r == 0 ? 1.0f : -1.0f;
                const float sv = sign * my_dot;
                // numerically stable sigmoid (include/cuNVSM/cuda_utils.h:192-214), fast-math intrinsics
                const float ex = __expf(-fabsf(sv));
                const float inv = __fdividef(1.0f, 1.0f + ex);
                float prob = sv >= 0.f ? inv : ex * inv;
                prob = prob < p.sig_lo_cmp ? p.sig_lo_val : (prob > p.sig_hi_cmp ? p.sig_hi_val : prob);
                const float wgt = r == 0 ? wneg * p.pos_scale : wneg;
                const float der = (prob >= p.der_hi_cmp || prob <= p.der_lo_cmp) ? 0.0f : 1.0f - prob;
                const float m = wgt * (der * p.bsn);
                if (r < R) coef = sign * m;
                // every group of four lanes evaluated the same four rows: group bch keeps them, so that after the
                // sweep lane r holds row r (one coalesced store of probs / mult, one log per row)
                if ((lane >> 2) == bch) { my_prob = prob; my_mult = m; my_wgt = wgt; }
            }
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const float cf = __shfl_sync(kFull, coef, rr);
#pragma unroll
                for (int k = 0; k < NCH; ++k) {
                    gp[k][0] += cf * x[rr][k].x; gp[k][1] += cf * x[rr][k].y;
                    gp[k][2] += cf * x[rr][k].z; gp[k][3] += cf * x[rr][k].w;
                }
            }
        }
        if (lane < R) {   // lane r = row r
            loss += my_wgt * __logf(my_prob);
            p.probs[i * R + lane] = my_prob;
            p.mult[i * R + lane] = my_mult;
        }
#pragma unroll
        for (int k = 0; k < NCH; ++k) {
            const int c4 = lane4 + k * 128;
            float dy[VEC];
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                const float yy = y[k][v];
                // (clip_min == -clip_max, host: act_params)
                const float dv = is_tanh ? (1.0f - yy * yy) : (fabsf(yy) < clip_max ? 1.0f : 0.0f);
                dy[v] = dv * gp[k][v];
                cs[k][v] += dy[v];
                cx[k][v] += dy[v] * tt[k][v];   // sum dy * (xhat + bias); the bias term is removed by the caller
            }
            if (p.tf32_gp) {
                float lo[VEC];
#pragma unroll
                for (int v = 0; v < VEC; ++v) { const float hi = round_tf32(dy[v]); lo[v] = dy[v] - hi; dy[v] = hi; }
                if (p.Gp_lo && (FULL || c4 < dd)) store_vec<VEC>(p.Gp_lo + i * dd + c4, lo);
            }
            if (FULL || c4 < dd) store_vec<VEC>(p.Gp + i * dd + c4, dy);
        }
        // rotate the instance-weight queue; all lanes are done with stage s before it is refilled
#pragma unroll
        for (int t = 0; t < 3; ++t) iw_q[t] = iw_q[t + 1];
        s = (s + 1 == S) ? 0 : s + 1;
        __syncwarp();
    }
#pragma unroll
    for (int k = 0; k < NCH; ++k) {
        const int c4 = lane4 + k * 128;
        if (FULL || c4 < dd) {
#pragma unroll
            for (int v = 0; v < VEC; ++v) {
                atomicAdd(&sums[c4 + v], cs[k][v]);
                atomicAdd(&sums[dd + c4 + v], cx[k][v]);
            }
        }
    }
    loss = warp_sum(loss);
    if (lane == 0) atomicAdd(&sums[2 * dd], loss);
    __syncthreads();
    for (int t = threadIdx.x; t < 2 * dd; t += blockDim.x) atomicAdd(p.col_sums + t, (double)sums[t]);
    if (threadIdx.x == 0) atomicAdd(p.loss_acc, (double)sums[2 * dd]);
    score_sums_tail(p.xchg, p.col_sums, 2 * dd + 1, p.xchg_epoch, 3, p.xchg_counter, p.xchg_error, p.loss_host);
    (void)nvec;
}

// bn_scale = invstd, bn_shift = bias - mean * invstd
__global__ void bn_affine_kernel(const float* __restrict__ mean, const float* __restrict__ invstd,
                                 const float* __restrict__ bias, int dd, float* __restrict__ scale,
                                 float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= dd) return;
    scale[c] = invstd[c];
    shift[c] = bias[c] - mean[c] * invstd[c];
}

}  // namespace nvsm
