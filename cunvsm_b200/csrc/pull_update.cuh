// Pull-style ("gather, don't scatter") table update for full Adam
// (DENSE_UPDATE_DENSE_VARIANCE, cpp/updates_adam.cu:199-213,251-283,310-328).
//
// full Adam touches every row of the table every step (both moments decay and the L2 term
// acts on all rows), so the update is a dense pass anyway. Instead of scatter-adding the
// sparse gradient into a table-sized buffer with atomics (reference: update_repr_kernel,
// cpp/storage.cu:37-49) and then streaming that buffer back in, the batch's references are
// bucketed by row (counting sort: histogram -> exclusive scan -> fill) and the dense pass
// pulls each row's gradient from the L2-resident per-n-gram tensors while it already holds
// theta / m / v of that row in registers:
//   entities: agg[d] = sum_{c : id[c] = d} (+-mult[c]) * Y[c / R]        (cpp/objective.cu:381-401)
//   words   : agg[w] = sum_{c : id[c] = w} fw[c] * gP[c / n]             (intermediate_results.cu:283-317)
// No atomics on floats, no memset, one read and one write of theta / m / v per step.
#pragma once

#include "common.cuh"

namespace nvsm {

// Histogram of the batch's ids. Skewed id streams (Zipfian words, skewed negatives) would hammer one address with
// tens of thousands of atomics — the L2 slice that owns it serialises them and stalls whatever else runs (measured: the
// forward GEMM next to this kernel 54 -> 140 us under Zipf(1) word ids). Each block therefore first aggregates in a
// small direct-mapped shared-memory table (slot claimed with a CAS on the id); ids that lose their slot to another id
// go to global memory directly, as before. One global atomic per (block, hot id).

__global__ void __launch_bounds__(kAggThreads) ref_count_kernel(const idx_t* __restrict__ ids, long total,
                                                                int* __restrict__ counts) {
    __shared__ int s_id[kAggSlots], s_cnt[kAggSlots];
    for (int i = threadIdx.x; i < kAggSlots; i += kAggThreads) { s_id[i] = -1; s_cnt[i] = 0; }
    __syncthreads();
    const long c = (long)blockIdx.x * kAggThreads + threadIdx.x;
    if (c < total) {
        const int id = (int)__ldg(ids + c);
        const int slot = agg_slot(id);
        const int prev = atomicCAS(&s_id[slot], -1, id);
        if (prev == -1 || prev == id) atomicAdd(&s_cnt[slot], 1);
        else atomicAdd(counts + id, 1);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAggSlots; i += kAggThreads)
        if (s_cnt[i] > 0) atomicAdd(counts + s_id[i], s_cnt[i]);
}

// ---- skewed reference counts -------------------------------------------------------------------------------------
// One warp per row is the right shape while rows have ~10 references (uniform ids). Real id streams are Zipfian (the
// most frequent word of a 51200 x 10 batch over 50k words is referenced ~45 000 times; BASELINE configs[4] draws
// Zipf-skewed negatives), and one warp walking thousands of references is a millisecond-long tail. Rows with more
// than kHeavyRefs references are therefore skipped by the row kernel: the bucket build (scan_add_kernel, which runs on
// the auxiliary stream under the forward pass) appends one work item per segment of kHeavyRefs references of such a
// row to a list, and pull_heavy_kernel gives every segment its own warp, publishes the segment's
// partial sum and lets the LAST segment to arrive (per-row arrival counter) add the partials in segment order and
// apply the row's update. No float atomics; the summation order inside a row is fixed by the bucket order.
constexpr int kHeavyRefs = 64;

struct HeavyWork {
    int2* items;      // (row, segment), appended by scan_add_kernel
    int* count;       // number of items; reset by the first scan kernel of the bucket build
    float* part;      // [capacity][ld] partial sums (+ the squared-gradient partial at [ld - 4])
    int* arrivals;    // arrival counter of a row, at the index of its first item; reset by the last arrival
    int capacity;     // items allocated: >= 2 * references / kHeavyRefs + 1 cannot overflow
    int ld;           // floats per partial: dim rounded up to 4, + 4
};

// Exclusive scan, three small launches: per-block scan of 1024 counts + block totals,
// scan of the (<= 1024 * 1024 / 1024) block totals, add back.
__global__ void __launch_bounds__(1024) scan_blocks_kernel(const int* __restrict__ in, long n, int* __restrict__ out,
                                                           int* __restrict__ block_sums, int* __restrict__ reset = nullptr) {
    __shared__ int warp_tot[32];
    if (reset && blockIdx.x == 0 && threadIdx.x == 0) *reset = 0;   // HeavyWork::count, appended to by scan_add_kernel
    const long i = (long)blockIdx.x * 1024 + threadIdx.x;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int x = i < n ? in[i] : 0;
    int v = x;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int t = __shfl_up_sync(kFull, v, o);
        if (lane >= o) v += t;
    }
    if (lane == 31) warp_tot[w] = v;
    __syncthreads();
    if (w == 0) {
        int t = warp_tot[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int u = __shfl_up_sync(kFull, t, o);
            if (lane >= o) t += u;
        }
        warp_tot[lane] = t;
    }
    __syncthreads();
    const int prefix = (w > 0 ? warp_tot[w - 1] : 0) + v - x;  // exclusive within the block
    if (i < n) out[i] = prefix;
    if (threadIdx.x == 1023 && block_sums) block_sums[blockIdx.x] = prefix + x;
}

__global__ void __launch_bounds__(1024) scan_add_kernel(int* __restrict__ out, long n, const int* __restrict__ block_prefix,
                                                        long total_refs, const int* __restrict__ counts,
                                                        const HeavyWork* __restrict__ heavy) {
    const long i = (long)blockIdx.x * 1024 + threadIdx.x;
    if (i < n) out[i] += block_prefix[blockIdx.x];
    if (i == 0) out[n] = (int)total_refs;  // sentinel: offsets[n] = number of references
    if (heavy && i < n) {                  // rows the row kernel will skip: one work item per segment
        const int cnt = counts[i];
        if (cnt > kHeavyRefs) {
            const HeavyWork hw = *heavy;
            const int nseg = (cnt + kHeavyRefs - 1) / kHeavyRefs;
            const int pos = atomicAdd(hw.count, nseg);
            for (int sgm = 0; sgm < nseg && pos + sgm < hw.capacity; ++sgm) hw.items[pos + sgm] = make_int2((int)i, sgm);
        }
    }
}

// counts[] still holds the histogram; popping it hands out the slots of each bucket. Same block-level aggregation as
// ref_count_kernel: a block reserves a range of a hot bucket with ONE atomicSub and its threads take consecutive slots.
__global__ void __launch_bounds__(kAggThreads) ref_fill_kernel(const idx_t* __restrict__ ids, long total,
                                                               const int* __restrict__ offsets, int* __restrict__ counts,
                                                               int* __restrict__ refs) {
    __shared__ int s_id[kAggSlots], s_cnt[kAggSlots], s_base[kAggSlots];
    for (int i = threadIdx.x; i < kAggSlots; i += kAggThreads) { s_id[i] = -1; s_cnt[i] = 0; }
    __syncthreads();
    const long c = (long)blockIdx.x * kAggThreads + threadIdx.x;
    int id = -1, slot = -1, rank = 0;
    if (c < total) {
        id = (int)__ldg(ids + c);
        slot = agg_slot(id);
        const int prev = atomicCAS(&s_id[slot], -1, id);
        if (prev == -1 || prev == id) rank = atomicAdd(&s_cnt[slot], 1);
        else slot = -1;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kAggSlots; i += kAggThreads)
        if (s_cnt[i] > 0) s_base[i] = atomicSub(counts + s_id[i], s_cnt[i]);
    __syncthreads();
    if (c < total) {
        const int pos = slot >= 0 ? s_base[slot] - 1 - rank : atomicSub(counts + id, 1) - 1;
        refs[offsets[id] + pos] = (int)c;
    }
}

struct AdamFullConsts {
    float s1, lr1, reg1, s2, lr2, lambda, lr, bc, eps;
};

// One warp per table row. SRC rows (Y or gP) and the per-reference coefficient:
//   ENTITY:  src row = ref / group, coef = (ref % group == 0 ? +1 : -1) * coefs[ref]   (group = R)
//   WORD  :  src row = ref / group, coef = coefs[ref]                                  (group = n)
template <int VEC, int NCH, bool ENTITY>
__global__ void __launch_bounds__(256) adam_full_pull_kernel(float* __restrict__ theta, float* __restrict__ m,
                                                             float* __restrict__ v, long num_rows, int dim,
                                                             const int* __restrict__ offsets,
                                                             const int* __restrict__ refs,
                                                             const float* __restrict__ coefs,
                                                             const float* __restrict__ src, int group,
                                                             const AdamFullConsts k,
                                                             const float* __restrict__ self_k, const int heavy_above) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim / VEC;
    for (long row = warp0; row < num_rows; row += nwarps) {
        const int beg = __ldg(offsets + row), end = __ldg(offsets + row + 1);
        if (end - beg > heavy_above) continue;   // pull_heavy_kernel owns this row
        // entity normalisation: the gradient carries - self_k[row] * theta[row] (entity_norm_prep_kernel)
        const float ks = self_k ? __ldg(self_k + row) : 0.f;
        float agg[NCH][VEC];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int q = 0; q < VEC; ++q) agg[j][q] = 0.f;
        for (int base = beg; base < end; base += 32) {
            // lanes fetch up to 32 references of this row at once, then walk them together
            const int cnt = min(32, end - base);
            int my_src = 0;
            float my_coef = 0.f;
            if (lane < cnt) {
                const int ref = __ldg(refs + base + lane);
                my_src = ref / group;
                const float cf = __ldg(coefs + ref);
                my_coef = (ENTITY && (ref - my_src * group) != 0) ? -cf : cf;
            }
            for (int t = 0; t < cnt; ++t) {
                const int srow = __shfl_sync(kFull, my_src, t);
                const float cf = __shfl_sync(kFull, my_coef, t);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = lane + j * kWarp;
                    if (c < nvec) {
                        float x[VEC];
                        load_vec_ro<VEC>(src + (long)srow * dim + c * VEC, x);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) agg[j][q] += cf * x[q];
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            if (c < nvec) {
                const long o = row * dim + c * VEC;
                float th[VEC], mm[VEC], vv[VEC];
                // theta / m / v stream through once per step: evict-first, so the gathered source rows (Y, grad_phrase:
                // 52 / 61 MB, each row referenced ~10 times) stay L2-resident (ncu before: 2.7 - 3x DRAM re-reads of them)
                load_vec_cs<VEC>(theta + o, th);
                load_vec_cs<VEC>(m + o, mm);
                load_vec_cs<VEC>(v + o, vv);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const float ag = agg[j][q] - ks * th[q];
                    const float g = ag + (-k.lambda * th[q]);
                    mm[q] = (mm[q] * k.s1 + k.lr1 * ag) + (-k.reg1 * th[q]);
                    vv[q] = vv[q] * k.s2 + (g * g) * k.lr2;
                    th[q] = th[q] + (fast_div(mm[q], fast_sqrt(vv[q]) + k.eps) * k.bc) * k.lr;
                }
                store_vec_cs<VEC>(theta + o, th);
                store_vec_cs<VEC>(m + o, mm);
                store_vec_cs<VEC>(v + o, vv);
            }
        }
    }
}

// Pull-style SGD / Adagrad (cpp/storage.cu:51-102, cpp/updates_adagrad.cu:99-179) over the same reference buckets:
//   theta[row] = theta[row] * decay + lr * rs * sum_{refs of row} coef_ref * src[ref / group]
// decay = 1 - lambda_s * lr is the reference's dense whole-table decay (touch_all != 0), otherwise rows without
// references are not touched at all. Replaces one float atomic per element per reference (update_repr_kernel) by
// reads of the L2-resident per-n-gram rows; every row is written once.
//   entity Adagrad (acc != null): acc[row] += sum_refs coef^2 * ysq[ref / group]  (mean_k of the squared gradient
//   column, window 1), rs = 1 / sqrt(acc[row] + eps) with the UPDATED accumulator, like adagrad_update_kernel.
//   word Adagrad: the per-n-gram factor 1 / sqrt(mean_w acc[id_w] + eps) is folded into coefs by the caller.
template <int VEC, int NCH, bool ENTITY>
__global__ void __launch_bounds__(256) sgd_pull_kernel(float* __restrict__ theta, long num_rows, int dim,
                                                       const int* __restrict__ offsets, const int* __restrict__ refs,
                                                       const float* __restrict__ coefs, const float* __restrict__ src,
                                                       int group, float decay, float lr, int touch_all,
                                                       float* __restrict__ acc, const float* __restrict__ ysq, float eps,
                                                       const int heavy_above) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim / VEC;
    for (long row = warp0; row < num_rows; row += nwarps) {
        const int beg = __ldg(offsets + row), end = __ldg(offsets + row + 1);
        if (beg == end && !touch_all) continue;
        if (end - beg > heavy_above) continue;   // pull_heavy_kernel owns this row
        float agg[NCH][VEC];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int q = 0; q < VEC; ++q) agg[j][q] = 0.f;
        float sq = 0.f;
        for (int base = beg; base < end; base += 32) {
            const int cnt = min(32, end - base);
            int my_src = 0;
            float my_coef = 0.f;
            if (lane < cnt) {
                const int ref = __ldg(refs + base + lane);
                my_src = ref / group;
                const float cf = __ldg(coefs + ref);
                my_coef = (ENTITY && (ref - my_src * group) != 0) ? -cf : cf;
                if (acc) sq += cf * cf * __ldg(ysq + my_src);
            }
            for (int t = 0; t < cnt; ++t) {
                const int srow = __shfl_sync(kFull, my_src, t);
                const float cf = __shfl_sync(kFull, my_coef, t);
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = lane + j * kWarp;
                    if (c < nvec) {
                        float x[VEC];
                        load_vec_ro<VEC>(src + (long)srow * dim + c * VEC, x);
#pragma unroll
                        for (int q = 0; q < VEC; ++q) agg[j][q] += cf * x[q];
                    }
                }
            }
        }
        float rs = 1.0f;
        if (acc) {
            sq = warp_sum(sq);
            const float a = acc[row] + sq;
            if (lane == 0 && beg != end) acc[row] = a;
            rs = 1.0f / sqrtf(a + eps);
        }
        const float step = lr * rs;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            if (c < nvec) {
                const long o = row * dim + c * VEC;
                float th[VEC];
                load_vec_cs<VEC>(theta + o, th);
#pragma unroll
                for (int q = 0; q < VEC; ++q) th[q] = th[q] * decay + step * agg[j][q];
                store_vec_cs<VEC>(theta + o, th);
            }
        }
    }
}

// The same update when NO row is touched unless it is referenced (touch_all == 0: lambda == 0, or the float decay factor
// 1 - lambda_s lr rounds to exactly 1.0f -- a bit-exact no-op, C3 / C5). Then (i) the scan over the rows is the cost when
// most rows are empty (C5: 1 M entity rows, 135 k references: one offsets round trip per row and warp was 70 us of a
// 0.18 ms step), and (ii) a referenced row is a chain of dependent loads -- offsets -> refs -> coefs -> source rows ->
// theta -- with ~1.7 references per row (C3: 57 % of the HBM peak, ncu r2a). Here the 32 lanes of a warp fetch the
// bucket bounds AND the first two references (source row, signed coefficient, squared-gradient term) of the warp's next
// 32 rows at once; the warp then walks only the rows that have work, with everything but the source rows and theta
// already in registers, and theta is requested before the gather. Rows stay dealt round-robin (warp w owns rows w,
// w + nwarps, ...): skewed id streams put the referenced rows next to each other, and a warp owning 32 CONSECUTIVE
// rows of the hot region ran 4x longer than the rest. Same arithmetic and summation order as sgd_pull_kernel.
template <int VEC, int NCH, bool ENTITY>
__global__ void __launch_bounds__(256) sgd_pull_sparse_kernel(float* __restrict__ theta, long num_rows, int dim,
                                                              const int* __restrict__ offsets, const int* __restrict__ refs,
                                                              const float* __restrict__ coefs, const float* __restrict__ src,
                                                              int group, float decay, float lr, float* __restrict__ acc,
                                                              const float* __restrict__ ysq, float eps, const int heavy_above) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const int nvec = dim / VEC;
    for (long row0 = warp0; row0 < num_rows; row0 += nwarps * 32) {
        // ---- lane l: metadata of row row0 + l * nwarps ----
        const long lrow = row0 + (long)lane * nwarps;
        int lbeg = 0, lend = 0;
        if (lrow < num_rows) { lbeg = __ldg(offsets + lrow); lend = __ldg(offsets + lrow + 1); }
        const int lcnt = lend - lbeg;
        const bool lwork = lcnt > 0 && lcnt <= heavy_above;     // (rows above heavy_above: pull_heavy_kernel)
        int src0 = 0, src1 = 0;
        float cf0 = 0.f, cf1 = 0.f, sq0 = 0.f, sq1 = 0.f;
        if (lwork) {
            const int ref = __ldg(refs + lbeg);
            src0 = ref / group;
            const float c = __ldg(coefs + ref);
            cf0 = (ENTITY && (ref - src0 * group) != 0) ? -c : c;
            if (acc) sq0 = c * c * __ldg(ysq + src0);
            if (lcnt > 1) {
                const int ref1 = __ldg(refs + lbeg + 1);
                src1 = ref1 / group;
                const float c1 = __ldg(coefs + ref1);
                cf1 = (ENTITY && (ref1 - src1 * group) != 0) ? -c1 : c1;
                if (acc) sq1 = c1 * c1 * __ldg(ysq + src1);
            }
        }
        unsigned todo = __ballot_sync(kFull, lwork);
        while (todo) {
            const int sub = __ffs(todo) - 1;
            todo &= todo - 1;
            const long row = row0 + (long)sub * nwarps;
            const int beg = __shfl_sync(kFull, lbeg, sub), end = __shfl_sync(kFull, lend, sub);
            float th[NCH][VEC];
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const int c = lane + j * kWarp;
                if (c < nvec) load_vec_cs<VEC>(theta + row * dim + c * VEC, th[j]);
            }
            float acc_old = 0.f;
            if (acc) acc_old = acc[row];
            float agg[NCH][VEC];
#pragma unroll
            for (int j = 0; j < NCH; ++j)
#pragma unroll
                for (int q = 0; q < VEC; ++q) agg[j][q] = 0.f;
            float sq;
            if (end - beg <= 2) {
                // both references came with the metadata; sq: lanes 0 and 1 of sgd_pull_kernel hold one term each and
                // warp_sum adds them -- the same two-term sum
                const int s0 = __shfl_sync(kFull, src0, sub), s1 = __shfl_sync(kFull, src1, sub);
                const float c0 = __shfl_sync(kFull, cf0, sub), c1 = __shfl_sync(kFull, cf1, sub);
                sq = __shfl_sync(kFull, sq0, sub) + __shfl_sync(kFull, sq1, sub);
                float x0[NCH][VEC], x1[NCH][VEC];
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = lane + j * kWarp;
                    if (c < nvec) {
                        load_vec_ro<VEC>(src + (long)s0 * dim + c * VEC, x0[j]);
                        load_vec_ro<VEC>(src + (long)s1 * dim + c * VEC, x1[j]);     // (one reference: row s1 = 0, coefficient 0)
                    }
                }
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = lane + j * kWarp;
                    if (c < nvec) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) agg[j][q] += c0 * x0[j][q];
                        if (end - beg == 2) {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) agg[j][q] += c1 * x1[j][q];
                        }
                    }
                }
            } else {
                sq = 0.f;
                for (int base = beg; base < end; base += 32) {
                    const int cnt = min(32, end - base);
                    int my_src = 0;
                    float my_coef = 0.f;
                    if (lane < cnt) {
                        const int ref = __ldg(refs + base + lane);
                        my_src = ref / group;
                        const float cf = __ldg(coefs + ref);
                        my_coef = (ENTITY && (ref - my_src * group) != 0) ? -cf : cf;
                        if (acc) sq += cf * cf * __ldg(ysq + my_src);
                    }
                    for (int t = 0; t < cnt; ++t) {
                        const int srow = __shfl_sync(kFull, my_src, t);
                        const float cf = __shfl_sync(kFull, my_coef, t);
#pragma unroll
                        for (int j = 0; j < NCH; ++j) {
                            const int c = lane + j * kWarp;
                            if (c < nvec) {
                                float x[VEC];
                                load_vec_ro<VEC>(src + (long)srow * dim + c * VEC, x);
#pragma unroll
                                for (int q = 0; q < VEC; ++q) agg[j][q] += cf * x[q];
                            }
                        }
                    }
                }
                if (acc) sq = warp_sum(sq);
            }
            float rs = 1.0f;
            if (acc) {
                const float a = acc_old + sq;
                if (lane == 0) acc[row] = a;
                rs = 1.0f / sqrtf(a + eps);
            }
            const float step = lr * rs;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const int c = lane + j * kWarp;
                if (c < nvec) {
#pragma unroll
                    for (int q = 0; q < VEC; ++q) th[j][q] = th[j][q] * decay + step * agg[j][q];
                    store_vec_cs<VEC>(theta + row * dim + c * VEC, th[j]);
                }
            }
        }
    }
}

// Row updates of the two pull kernels as functors, for pull_heavy_kernel (same arithmetic, same order).
struct AdamFullApply {
    float *theta, *m, *v;
    AdamFullConsts k;
    const float* self_k;
    template <int VEC, int NCH>
    __device__ __forceinline__ void run(long row, int dim, int lane, int nvec, const float (&agg)[NCH][VEC], float) const {
        const float ks = self_k ? __ldg(self_k + row) : 0.f;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            if (c < nvec) {
                const long o = row * dim + c * VEC;
                float th[VEC], mm[VEC], vv[VEC];
                load_vec_cs<VEC>(theta + o, th);
                load_vec_cs<VEC>(m + o, mm);
                load_vec_cs<VEC>(v + o, vv);
#pragma unroll
                for (int q = 0; q < VEC; ++q) {
                    const float ag = agg[j][q] - ks * th[q];
                    const float g = ag + (-k.lambda * th[q]);
                    mm[q] = (mm[q] * k.s1 + k.lr1 * ag) + (-k.reg1 * th[q]);
                    vv[q] = vv[q] * k.s2 + (g * g) * k.lr2;
                    th[q] = th[q] + (fast_div(mm[q], fast_sqrt(vv[q]) + k.eps) * k.bc) * k.lr;
                }
                store_vec_cs<VEC>(theta + o, th);
                store_vec_cs<VEC>(m + o, mm);
                store_vec_cs<VEC>(v + o, vv);
            }
        }
    }
};

struct SgdApply {
    float* theta;
    float decay, lr;
    float* acc;   // entity Adagrad accumulator or null
    float eps;
    template <int VEC, int NCH>
    __device__ __forceinline__ void run(long row, int dim, int lane, int nvec, const float (&agg)[NCH][VEC], float sq) const {
        float rs = 1.0f;
        if (acc) {
            const float a = acc[row] + sq;
            __syncwarp();
            if (lane == 0) acc[row] = a;
            rs = 1.0f / sqrtf(a + eps);
        }
        const float step = lr * rs;
#pragma unroll
        for (int j = 0; j < NCH; ++j) {
            const int c = lane + j * kWarp;
            if (c < nvec) {
                const long o = row * dim + c * VEC;
                float th[VEC];
                load_vec_cs<VEC>(theta + o, th);
#pragma unroll
                for (int q = 0; q < VEC; ++q) th[q] = th[q] * decay + step * agg[j][q];
                store_vec_cs<VEC>(theta + o, th);
            }
        }
    }
};

constexpr int kHeavyGroup = 16;   // segments per first-level reduction group of pull_heavy_kernel

// Rows a lane keeps in flight in pull_heavy_kernel (register budget: 128 at two blocks per SM).
template <int VEC, int NCH>
__host__ __device__ constexpr int heavy_in_flight() { return VEC * NCH >= 32 ? 1 : (VEC * NCH >= 12 ? 2 : 4); }

// partial sum of a warp -> part[0 .. dim) (+ the scalar at [ld - 4]); made visible device-wide before returning
template <int VEC, int NCH>
__device__ __forceinline__ void store_partial(float* p, int ld, int lane, int nvec, const float (&agg)[NCH][VEC], float sq) {
#pragma unroll
    for (int j = 0; j < NCH; ++j) {
        const int c = lane + j * kWarp;
        if (c < nvec) store_vec<VEC>(p + c * VEC, agg[j]);
    }
    __syncwarp();   // (every lane has read the scalar of a slot that is being overwritten)
    if (lane == 0) p[ld - 4] = sq;
    __threadfence();
    __syncwarp();
}

// Count this warp in at `counter`; true for the warp that completes `expected` arrivals (it also re-arms the counter
// for the next step). The fences pair the partial stores of the others with the loads of the last one.
__device__ __forceinline__ bool arrive_last(int* counter, int expected, int lane) {
    int last = 0;
    if (lane == 0) {
        last = atomicAdd(counter, 1) == expected - 1;
        if (last) *counter = 0;
    }
    last = __shfl_sync(kFull, last, 0);
    if (last) __threadfence();
    return last != 0;
}

// agg = p[0] + p[stride] + ... (count partials of width ld, in that order; heavy_in_flight rows in flight per lane)
template <int VEC, int NCH>
__device__ __forceinline__ void sum_partials(const float* p, long stride, int ld, int count, int lane, int nvec,
                                             float (&agg)[NCH][VEC], float& sq) {
#pragma unroll
    for (int j = 0; j < NCH; ++j)
#pragma unroll
        for (int q = 0; q < VEC; ++q) agg[j][q] = 0.f;
    sq = 0.f;
    constexpr int U = heavy_in_flight<VEC, NCH>();
    for (int k0 = 0; k0 < count; k0 += U) {
        float x[U][NCH][VEC], s4[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const float* const pu = p + (long)min(k0 + u, count - 1) * stride;
#pragma unroll
            for (int j = 0; j < NCH; ++j) {
                const int c = lane + j * kWarp;
                if (c < nvec) load_vec_cg<VEC>(pu + c * VEC, x[u][j]);
            }
            s4[u] = __ldcg(pu + ld - 4);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            if (k0 + u < count) {
#pragma unroll
                for (int j = 0; j < NCH; ++j) {
                    const int c = lane + j * kWarp;
                    if (c < nvec) {
#pragma unroll
                        for (int q = 0; q < VEC; ++q) agg[j][q] += x[u][j][q];
                    }
                }
                sq += s4[u];
            }
        }
    }
}

// One warp per (row, segment) item of the heavy list; see HeavyWork. Several references are in flight per lane (the
// FMAs still run in reference order), since here the walk is long enough for load latency to be the limit.
template <int VEC, int NCH, bool ENTITY, typename Apply>
__global__ void __launch_bounds__(256, 2) pull_heavy_kernel(int dim, const int* __restrict__ offsets,
                                                         const int* __restrict__ refs, const float* __restrict__ coefs,
                                                         const float* __restrict__ src, int group,
                                                         const float* __restrict__ ysq, const HeavyWork hw,
                                                         const Apply apply) {
    const int lane = threadIdx.x & 31;
    const int warp0 = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const int nvec = dim / VEC;
    const int num_items = min(*hw.count, hw.capacity);
    for (int item = warp0; item < num_items; item += nwarps) {
        const int2 it = hw.items[item];
        const long row = it.x;
        const int base = item - it.y;   // the items of a row are contiguous, segment 0 first
        const int rbeg = __ldg(offsets + row), rend = __ldg(offsets + row + 1);
        const int nseg = (rend - rbeg + kHeavyRefs - 1) / kHeavyRefs;
        const int beg = rbeg + it.y * kHeavyRefs, end = min(beg + kHeavyRefs, rend);
        float agg[NCH][VEC];
#pragma unroll
        for (int j = 0; j < NCH; ++j)
#pragma unroll
            for (int q = 0; q < VEC; ++q) agg[j][q] = 0.f;
        float sq = 0.f;
        for (int b0 = beg; b0 < end; b0 += 32) {
            const int cnt = min(32, end - b0);
            int my_src = 0;
            float my_coef = 0.f;
            if (lane < cnt) {
                const int ref = __ldg(refs + b0 + lane);
                my_src = ref / group;
                const float cf = __ldg(coefs + ref);
                my_coef = (ENTITY && (ref - my_src * group) != 0) ? -cf : cf;
                if (ysq) sq += cf * cf * __ldg(ysq + my_src);
            }
            constexpr int U = heavy_in_flight<VEC, NCH>();
            for (int t = 0; t < cnt; t += U) {   // lanes >= cnt hold source row 0: loaded, never added
                int srow[U];
                float cf[U], x[U][NCH][VEC];
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    srow[u] = __shfl_sync(kFull, my_src, (t + u) & 31);
                    cf[u] = __shfl_sync(kFull, my_coef, (t + u) & 31);
                }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const int c = lane + j * kWarp;
                        if (c < nvec) load_vec_ro<VEC>(src + (long)srow[u] * dim + c * VEC, x[u][j]);
                    }
#pragma unroll
                for (int u = 0; u < U; ++u)
#pragma unroll
                    for (int j = 0; j < NCH; ++j) {
                        const int c = lane + j * kWarp;
                        if (c < nvec && t + u < cnt) {
#pragma unroll
                            for (int q = 0; q < VEC; ++q) agg[j][q] += cf[u] * x[u][j][q];
                        }
                    }
            }
        }
        // Publish this segment's partial and count it in. The partials of a row are added up in two levels so that
        // no warp walks hundreds of them: the last segment to arrive in a group of kHeavyGroup consecutive segments
        // adds the group's partials (in segment order) and publishes the group sum in the group's first slot; the
        // last GROUP to arrive adds the group sums (in group order) and applies the row update.
        const int grp = it.y / kHeavyGroup;
        const int gfirst = base + grp * kHeavyGroup;                    // item index of the group's first segment
        const int gsize = min(kHeavyGroup, nseg - grp * kHeavyGroup);
        const int ngrp = (nseg + kHeavyGroup - 1) / kHeavyGroup;
        sq = warp_sum(sq);
        // level 0: my partial -> my slot, group counter; level 1: the group sum -> the group's first slot, row counter
        long slot = item, first = gfirst, stride = hw.ld;
        int* counter = hw.arrivals + gfirst;
        int expected = gsize;
        bool mine_to_apply = false;
        for (int level = 0; level < 2; ++level) {
            store_partial<VEC, NCH>(hw.part + slot * hw.ld, hw.ld, lane, nvec, agg, sq);
            if (!arrive_last(counter, expected, lane)) break;
            sum_partials<VEC, NCH>(hw.part + first * hw.ld, stride, hw.ld, expected, lane, nvec, agg, sq);
            if (level == 1 || ngrp == 1) { mine_to_apply = true; break; }
            // (row counter: the slot after the row's first item — a heavy row has at least two segments, and group
            // counters sit at multiples of kHeavyGroup >= 2)
            slot = gfirst; first = base; stride = (long)hw.ld * kHeavyGroup;
            counter = hw.arrivals + base + 1;
            expected = ngrp;
        }
        if (!mine_to_apply) continue;
        apply.template run<VEC, NCH>(row, dim, lane, nvec, agg, sq);
    }
}

// Word-side Adagrad coefficients: wcoef[i * n + w] = fw[i, w] / sqrt(mean_w' acc[id[i, w']] + eps)
// (adagrad_update_kernel with window n, cpp/updates_adagrad.cu:83-97; acc already holds this batch's contribution).
__global__ void word_adagrad_coef_kernel(const idx_t* __restrict__ ids, const float* __restrict__ fw,
                                         const float* __restrict__ acc, long B, int n, float eps,
                                         float* __restrict__ wcoef) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= B) return;
    float a = 0.f;
    for (int w = 0; w < n; ++w) a += __ldg(acc + __ldg(ids + i * n + w));
    a /= (float)n;
    const float factor = 1.0f / sqrtf(a + eps);
    for (int w = 0; w < n; ++w) wcoef[i * n + w] = __ldg(fw + i * n + w) * factor;
}

}  // namespace nvsm
