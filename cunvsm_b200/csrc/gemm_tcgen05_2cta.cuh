// Two-SM variant of the K-major projection GEMMs (forward and grad_phrase): a cluster of two CTAs (one TPC)
// computes one 256 x bn output tile with tcgen05.mma.cta_group::2. CTA r of the pair stages ITS 128 rows of A and
// ITS bn/2 rows of B per k-block; the pair's tensor cores read each other's B half, so per SM the shared-memory
// fill and the operand reads per MMA drop from (128 + bn) to (128 + bn/2) k-rows. The single-SM kernel
// (gemm_tcgen05.cuh) is bound exactly there: tf32 operands are 4 bytes, 3xTF32 issues three MMAs per k-step, and
// TMA fill + MMA operand reads exceed the 128 B/cycle shared-memory port (ncu: tensor pipe ~52 % busy, L2 ~50 %).
//
// Protocol (per CTA: warp 0 TMA producer, warp 1 MMA issuer -- leader CTA only --, warps 2-5 epilogue):
//   full_bar[s]       leader's copy only: armed by the leader's producer with the bytes of BOTH CTAs; every TMA of the
//                     pair signals it (the peer addresses it through mapa).
//   empty_bar[s]      one per CTA, released by the leader's tcgen05.commit multicast to both CTAs.
//   tmem_full_bar[b]  one per CTA, same multicast commit after the last k-block of a tile.
//   tmem_empty_bar[b] leader's copy only, 8 arrivals (4 epilogue warps x 2 CTAs; the peer's arrive remotely).
// Each CTA drains its own 128 accumulator rows from its own TMEM.
#pragma once

#include "gemm_tcgen05.cuh"

namespace nvsm {
namespace tc {

__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of the same shared variable in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t map_to_cta(uint32_t local_addr, uint32_t rank) {
    uint32_t ra;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(ra) : "r"(local_addr), "r"(rank));
    return ra;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint32_t bar, uint16_t mask) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(bar), "h"(mask)
                 : "memory");
}
__device__ __forceinline__ void umma_tf32_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                              uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// tf32 x tf32 -> f32, M = 256 (pair), both operands K-major
__host__ __device__ inline uint32_t make_idesc_2sm(int n) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format: F32
    d |= 2u << 7;                    // A format: TF32
    d |= 2u << 10;                   // B format: TF32
    d |= (uint32_t)(n >> 3) << 17;   // N / 8
    d |= (uint32_t)(256 >> 4) << 24; // M / 16
    return d;
}

// Params: m_tiles counts 256-row tiles; stage_bytes is PER CTA: (128 + bn / 2) k-rows of 128 bytes, x2 when SPLIT.
template <bool SPLIT>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kThreads, 1)
gemm_tc2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                const __grid_constant__ CUtensorMap tmAlo, const __grid_constant__ CUtensorMap tmBlo, const Params p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[8];
    __shared__ __align__(8) uint64_t empty_bar[8];
    __shared__ __align__(8) uint64_t tmem_full_bar[2];
    __shared__ __align__(8) uint64_t tmem_empty_bar[2];
    __shared__ uint32_t tmem_base_smem;
    __shared__ float stat_strip[4][2 * kStatCols];   // fused column statistics (Params::stat_part)

    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = cluster_ctarank();
    const bool leader = rank == 0;
    const int cluster_id = blockIdx.x >> 1, num_clusters = gridDim.x >> 1;
    const int num_kb_total = (p.K + kBlockK - 1) / kBlockK;
    const int num_tiles = p.m_tiles * p.n_tiles;
    const int stages = p.stages;
    const int half_bn = p.bn >> 1;
    const uint32_t b_bytes = (uint32_t)half_bn * 128u;
    const uint32_t lo_off = kATileBytes + b_bytes;   // SPLIT: [A_hi][B_hi][A_lo][B_lo] per stage

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&tmem_full_bar[b]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[b]), 8);   // 4 epilogue warps of each CTA of the pair
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    cluster_sync_all();   // both CTAs' barriers exist before anything signals across the pair
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    asm volatile("griddepcontrol.wait;" ::: "memory");                 // (see gemm_tc_kernel)
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    auto decode = [&](int t, int& m0, int& n0) {
        n0 = (t % p.n_tiles) * p.bn;
        m0 = (t / p.n_tiles) * 256;
    };

    if (warp == 0) {
        // ===== TMA producer (both CTAs) =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            if constexpr (SPLIT) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
            }
            uint32_t it = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters) {
                int m0, n0;
                decode(t, m0, n0);
                const int my_m = m0 + 128 * (int)rank, my_n = n0 + half_bn * (int)rank;
                for (int kb = 0; kb < num_kb_total; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait(smem_u32(&empty_bar[s]), ((it / stages) & 1u) ^ 1u);
                    const uint32_t bar_local = smem_u32(&full_bar[s]);
                    if (leader) mbar_expect_tx(bar_local, 2u * p.stage_bytes);   // bytes of both CTAs
                    const uint32_t bar = map_to_cta(bar_local, 0);
                    const uint32_t a_dst = smem_base + (uint32_t)s * p.stage_bytes;
                    const uint32_t b_dst = a_dst + kATileBytes;
                    const int k0 = kb * kBlockK;
                    tma_load_2d_2sm(a_dst, &tmA, bar, k0, my_m);
                    tma_load_2d_2sm(b_dst, &tmB, bar, k0, my_n);
                    if constexpr (SPLIT) {
                        tma_load_2d_2sm(a_dst + lo_off, &tmAlo, bar, k0, my_m);
                        tma_load_2d_2sm(b_dst + lo_off, &tmBlo, bar, k0, my_n);
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (leader CTA only) =====
        if (lane == 0 && leader) {
            const uint32_t idesc = make_idesc_2sm(p.bn);
            uint32_t it = 0, lt = 0;
            for (int t = cluster_id; t < num_tiles; t += num_clusters, ++lt) {
                const uint32_t buf = lt & 1u;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), ((lt >> 1) & 1u) ^ 1u);
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn;
                for (int kb = 0; kb < num_kb_total; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait(smem_u32(&full_bar[s]), (it / stages) & 1u);
                    tc_fence_after();
                    const uint32_t a_base = smem_base + (uint32_t)s * p.stage_bytes;
                    const uint32_t b_base = a_base + kATileBytes;
#pragma unroll
                    for (int j = 0; j < kBlockK / kUmmaK; ++j) {
                        const uint64_t a_desc = make_desc(a_base + j * 32u, 16u, 1024u, 2u);
                        const uint64_t b_desc = make_desc(b_base + j * 32u, 16u, 1024u, 2u);
                        umma_tf32_2sm(tacc, a_desc, b_desc, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        if constexpr (SPLIT) {
                            const uint64_t a_lo = make_desc(a_base + lo_off + j * 32u, 16u, 1024u, 2u);
                            const uint64_t b_lo = make_desc(b_base + lo_off + j * 32u, 16u, 1024u, 2u);
                            umma_tf32_2sm(tacc, a_lo, b_desc, idesc, 1u);
                            umma_tf32_2sm(tacc, a_desc, b_lo, idesc, 1u);
                        }
                    }
                    umma_commit_2sm(smem_u32(&empty_bar[s]), 3);        // frees the stage in both CTAs
                }
                umma_commit_2sm(smem_u32(&tmem_full_bar[buf]), 3);      // accumulator ready in both CTAs
            }
        }
    } else {
        // ===== epilogue (both CTAs, own 128 rows) =====
        const int q = warp & 3;
        const float alpha = p.alpha;
        const float* __restrict__ bias = p.bias;
        const int Nv = p.N;
        const bool vec_ok = (p.ldc & 3) == 0;
        float* const strip = stat_strip[q];
        if (p.stat_part)
            for (int t = lane; t < 2 * kStatCols; t += 32) strip[t] = 0.f;
        uint32_t lt = 0;
        for (int t = cluster_id; t < num_tiles; t += num_clusters, ++lt) {
            int m0, n0;
            decode(t, m0, n0);
            const uint32_t buf = lt & 1u;
            const int m = m0 + 128 * (int)rank + q * 32 + lane;
            const bool row_ok = m < p.M;
            float* crow = p.C + (long)m * p.ldc;
            const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn + ((uint32_t)(q * 32) << 16);
            mbar_wait(smem_u32(&tmem_full_bar[buf]), (lt >> 1) & 1u);
            tc_fence_after();
            const int ncols = min(p.bn, Nv - n0);
            const int nchunks = (min(p.bn, max(ncols, 0)) + 15) / 16;
            uint32_t rcur[16], rnext[16];
            if (nchunks > 0) { tmem_ld16_nowait(tacc, rcur); tmem_ld_wait(); }
            for (int ch = 0; ch < nchunks; ++ch) {
                const int c0 = n0 + ch * 16;
                if (ch + 1 < nchunks) tmem_ld16_nowait(tacc + (uint32_t)(ch + 1) * 16u, rnext);
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(rcur[i]);
                const bool full = c0 + 16 <= Nv;
                float bv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) bv[i] = 0.f;
                if (bias) {
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if (c0 + i < Nv) bv[i] = __ldg(bias + c0 + i);
                }
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = fmaf(alpha, v[i], bv[i]);
                if (row_ok) {
                    if (full && vec_ok) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4)
                            *reinterpret_cast<float4*>(crow + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < Nv) crow[c0 + i] = v[i];
                    }
                }
                if (p.stat_part) stat_accumulate(strip, ch, lane, v);   // (rows past M are zero: zero-filled A, no bias)
                if (ch + 1 < nchunks) {
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) rcur[i] = rnext[i];
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                const uint32_t bar_local = smem_u32(&tmem_empty_bar[buf]);
                if (leader) mbar_arrive(bar_local);
                else mbar_arrive_cluster(map_to_cta(bar_local, 0));
            }
        }
        if (p.stat_part) {
            __syncwarp();
            stat_flush(strip, p.stat_part + ((size_t)blockIdx.x * 4 + q) * 2 * Nv, Nv, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // the peer may still be reading operands / signalling barriers of this CTA
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace tc
}  // namespace nvsm
