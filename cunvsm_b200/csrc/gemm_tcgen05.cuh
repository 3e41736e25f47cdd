// Projection GEMMs on the 5th-generation tensor cores (sm_100a): tcgen05.mma kind::tf32,
// operands staged in shared memory by TMA (128-byte swizzle), fp32 accumulators in TMEM,
// read back with tcgen05.ld for the epilogue. One warp-specialised CTA per output tile:
//   warp 0     TMA producer (one elected lane)
//   warp 1     MMA issuer   (one elected lane)
//   warps 2-5  epilogue     (each owns the 32 TMEM lanes of its warp_id % 4 quarter)
// connected by a ring of `stages` full/empty mbarriers and one accumulator-ready mbarrier.
//
// The three contractions of the step (replacing cuBLAS SGEMM behind device_matrix's
// matrix_mult; cpp/params.cu:417-421,528-531, cpp/objective.cu:453-456):
//   forward         Z [B, dd]  = P [B, dw] . Tt[dd, dw]^T          A K-major, B K-major
//   grad_phrase     gP[B, dw]  = dX[B, dd] . T [dw, dd]^T / n      A K-major, B K-major
//   grad_transform  gT[dw, dd] = P [B, dw]^T . dX[B, dd]           A MN-major, B MN-major,
//                                                                  split-K over the batch
// "K-major" = reduction index contiguous in memory; "MN-major" = output index contiguous.
//
// Shared-memory tile formats (all 128B-swizzled, 8-row atoms of 1024 bytes):
//   K-major  : rows of 32 fp32 (128 B) along k; 8-row atoms stacked along M/N (SBO = 1024).
//   MN-major : rows of 32 fp32 (128 B) along M/N; for 32-bit operands the hardware only takes
//              the SWIZZLE_128B_BASE32B format (TMA: SWIZZLE_128B_ATOM_32B): 32-byte chunks
//              XOR-ed with k-row % 4, 4 k-rows per 512-byte atom, atoms stacked along k
//              (SBO = 512); 32-wide M/N groups are separate TMA boxes, LBO bytes apart.
// One tcgen05.mma consumes K = 8 tf32: 32 bytes further along a K-major row, or one whole
// atom (1024 B) further in an MN-major tile.
#pragma once

#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace nvsm {
namespace tc {

constexpr int kBlockM = 128;        // UMMA M (cta_group::1)
constexpr int kBlockK = 32;         // fp32 elements per 128-byte swizzled row (KB = 32); KB = 16 uses 64-byte rows
constexpr int kUmmaK = 8;           // tf32 K per instruction
constexpr int kThreads = 192;
constexpr uint32_t kATileBytes = kBlockM * kBlockK * 4;  // 16 KB
constexpr uint32_t kGroupBytes = kBlockK * 32 * 4;       // one 32(MN) x 32(k) box: 4 KB

struct Params {
    int M, N, K;           // problem: C[M, N] = A . B over K
    int bn;                // tile N extent = UMMA N (multiple of 16, <= 256; MN-major B: multiple of 32)
    int m_tiles, n_tiles;  // output tiles
    int splits;            // split-K factor; tile index = (m_tile * n_tiles + n_tile) * splits + split
    int kb_per_split;      // k-blocks (of 32) per split
    int stages;
    int kb;                // k-block: fp32 elements along K per pipeline stage (16 or 32) = template KB
    uint32_t stage_bytes;
    uint32_t tmem_cols;    // power of two >= 2 * bn (two accumulator buffers)
    float* C;
    int ldc;
    long split_stride;     // elements between split-K partial outputs
    float alpha;
    const float* bias;     // nullable, per output column
    // Optional fused column statistics of C (batch-norm forward, cpp/cudnn_utils.cu:107-124): every epilogue warp
    // adds the column sums and sums of squares of the rows it drains and writes one fp32 partial row
    // stat_part[(blockIdx.x * 4 + warp) * 2 * N + {0, N} + column] when the kernel ends (n_tiles == 1, splits == 1,
    // N % 16 == 0, no bias: C = alpha * A.B). Replaces a separate pass over C (col_stats4_kernel).
    float* stat_part;
};

constexpr int kStatCols = 256;   // widest tile the fused statistics cover

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok = 0;
    for (uint32_t it = 0; it < (1u << 27); ++it) {
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// 64-bit shared-memory matrix descriptor (sm_100 "version 1").
// layout_type: 2 = SWIZZLE_128B (16-byte chunks, 8-row atoms; K-major operands),
//              1 = SWIZZLE_128B_BASE32B (32-byte chunks, 4-row atoms) — the only layout the
//                  hardware accepts for MN-major 32-bit (tf32) operands.
__device__ __forceinline__ uint64_t make_desc(uint32_t addr, uint32_t lbo_bytes, uint32_t sbo_bytes,
                                              uint32_t layout_type) {
    uint64_t d = 0;
    d |= (uint64_t)((addr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= (uint64_t)1 << 46;   // descriptor version (Blackwell)
    d |= (uint64_t)layout_type << 61;
    return d;
}

// 32-bit instruction descriptor: tf32 x tf32 -> f32, M = 128.
__host__ __device__ inline uint32_t make_idesc(int n, bool a_mn, bool b_mn) {
    uint32_t d = 0;
    d |= 1u << 4;                    // D format: F32
    d |= 2u << 7;                    // A format: TF32
    d |= 2u << 10;                   // B format: TF32
    d |= (a_mn ? 1u : 0u) << 15;     // A major: 0 = K, 1 = MN
    d |= (b_mn ? 1u : 0u) << 16;     // B major
    d |= (uint32_t)(n >> 3) << 17;   // N / 8
    d |= (uint32_t)(kBlockM >> 4) << 24;  // M / 16
    return d;
}

__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                          uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

__device__ __forceinline__ void tmem_ld16_nowait(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Transposing butterfly: every lane brings 32 partials x[0..31]; afterwards lane l holds the warp total of partial l.
// 31 shuffles instead of 160.
__device__ __forceinline__ float warp_reduce32_transposed(float (&x)[32], int lane) {
#pragma unroll
    for (int step = 16; step >= 1; step >>= 1) {
        const bool up = (lane & step) != 0;
#pragma unroll
        for (int i = 0; i < step; ++i) {
            const float keep = up ? x[i + step] : x[i];
            const float send = up ? x[i] : x[i + step];
            x[i] = keep + __shfl_xor_sync(0xffffffffu, send, step);
        }
    }
    return x[0];
}

// Column statistics of one drained chunk: v = this lane's row, 16 consecutive columns of the tile (chunk `ch`).
// Lane l < 16 accumulates the sum of column 16 ch + l, lane l >= 16 the sum of squares of column 16 ch + l - 16, in
// the warp's own 512-float strip of shared memory (slot 32 ch + l: no other thread touches it).
__device__ __forceinline__ void stat_accumulate(float* warp_strip, int ch, int lane, const float (&v)[16]) {
    float x[32];
#pragma unroll
    for (int i = 0; i < 16; ++i) { x[i] = v[i]; x[16 + i] = v[i] * v[i]; }
    const float r = warp_reduce32_transposed(x, lane);
    warp_strip[32 * ch + lane] += r;
}
__device__ __forceinline__ void stat_flush(const float* warp_strip, float* part_row, int n, int lane) {
    for (int ch = 0; ch * 16 < n; ++ch)
        part_row[(lane >> 4) * n + 16 * ch + (lane & 15)] = warp_strip[32 * ch + lane];
}

// Persistent, warp-specialised GEMM: grid = min(#tiles, #SMs); every CTA walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ... The TMA producer runs ahead across tile boundaries
// (the smem ring never drains), the MMA warp alternates between two TMEM accumulator buffers,
// and the four epilogue warps drain buffer b of tile t while the MMAs of tile t+1 fill buffer
// b ^ 1 (tmem_full / tmem_empty mbarriers).
// SPLIT = 3xTF32: every operand arrives as hi = rn_tf32(x) and lo = x - hi (two arrays written by the
// producers); per k-step the accumulator receives hi.hi + lo.hi + hi.lo, which restores fp32-level
// accuracy (the dropped lo.lo term is ~2^-22 relative) at three tensor-core instructions per step.
// KB = k-block per pipeline stage: 32 fp32 (128-byte rows, SWIZZLE_128B) or 16 fp32 (64-byte rows, SWIZZLE_64B for
// K-major operands; MN-major tiles simply hold 16 instead of 32 k-rows). Halving the k-block halves the stage and
// doubles the ring depth in the same shared memory: the 3xTF32 stages are (A + B) x (hi + lo) = 96 KB at KB = 32,
// i.e. a 2-deep ring that cannot cover the TMA latency (ncu: tensor pipe ~52 % busy, L2 ~50 %).
template <bool A_MN, bool B_MN, bool SPLIT, int KB>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
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
    constexpr uint32_t kATile = kBlockM * KB * 4;        // A tile bytes per stage
    constexpr uint32_t kGroup = KB * 32 * 4;             // one 32(MN) x KB(k) MN-major box
    constexpr uint32_t kSboK = KB == 32 ? 1024u : 512u;  // K-major: 8-row atom of KB*4-byte rows
    constexpr uint32_t kLayoutK = KB == 32 ? 2u : 4u;    // SWIZZLE_128B : SWIZZLE_64B
    const int num_kb_total = (p.K + KB - 1) / KB;
    const int num_tiles = p.m_tiles * p.n_tiles * p.splits;
    const int stages = p.stages;
    const uint32_t lo_off = kATile + (uint32_t)p.bn * (KB * 4u);   // SPLIT: [A_hi][B_hi][A_lo][B_lo] per stage

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(smem_u32(&full_bar[s]), 1);
            mbar_init(smem_u32(&empty_bar[s]), 1);
        }
        for (int b = 0; b < 2; ++b) {
            mbar_init(smem_u32(&tmem_full_bar[b]), 1);
            mbar_init(smem_u32(&tmem_empty_bar[b]), 4);   // one arrival per epilogue warp
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)),
                     "r"(p.tmem_cols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_base_smem;
    // programmatic dependent launch: the set-up above overlaps the tail of the kernel in front; operands are only read
    // (and C written) behind this point; the kernel behind may start its own set-up as this one's CTAs retire
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

    // tile -> (m0, n0, first k-block, number of k-blocks)
    auto decode = [&](int t, int& m0, int& n0, int& kb0, int& nkb, int& split) {
        split = t % p.splits;
        const int mn = t / p.splits;
        n0 = (mn % p.n_tiles) * p.bn;
        m0 = (mn / p.n_tiles) * kBlockM;
        kb0 = split * p.kb_per_split;
        nkb = max(0, min(p.kb_per_split, num_kb_total - kb0));
    };

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
            if constexpr (SPLIT) {
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmAlo) : "memory");
                asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
            }
            uint32_t it = 0;   // running k-block counter across tiles
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                int m0, n0, kb0, nkb, split;
                decode(t, m0, n0, kb0, nkb, split);
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait(smem_u32(&empty_bar[s]), ((it / stages) & 1u) ^ 1u);
                    const uint32_t bar = smem_u32(&full_bar[s]);
                    const uint32_t a_dst = smem_base + (uint32_t)s * p.stage_bytes;
                    const uint32_t b_dst = a_dst + kATile;
                    const int k0 = (kb0 + kb) * KB;
                    mbar_expect_tx(bar, p.stage_bytes);
                    if constexpr (!A_MN) {
                        tma_load_2d(a_dst, &tmA, bar, k0, m0);
                    } else {
#pragma unroll
                        for (int g = 0; g < kBlockM / 32; ++g) tma_load_2d(a_dst + g * kGroup, &tmA, bar, m0 + 32 * g, k0);
                    }
                    if constexpr (!B_MN) {
                        tma_load_2d(b_dst, &tmB, bar, k0, n0);
                    } else {
                        for (int g = 0; g < p.bn / 32; ++g) tma_load_2d(b_dst + g * kGroup, &tmB, bar, n0 + 32 * g, k0);
                    }
                    if constexpr (SPLIT) {
                        if constexpr (!A_MN) {
                            tma_load_2d(a_dst + lo_off, &tmAlo, bar, k0, m0);
                        } else {
#pragma unroll
                            for (int g = 0; g < kBlockM / 32; ++g)
                                tma_load_2d(a_dst + lo_off + g * kGroup, &tmAlo, bar, m0 + 32 * g, k0);
                        }
                        if constexpr (!B_MN) {
                            tma_load_2d(b_dst + lo_off, &tmBlo, bar, k0, n0);
                        } else {
                            for (int g = 0; g < p.bn / 32; ++g)
                                tma_load_2d(b_dst + lo_off + g * kGroup, &tmBlo, bar, n0 + 32 * g, k0);
                        }
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer =====
        if (lane == 0) {
            const uint32_t idesc = make_idesc(p.bn, A_MN, B_MN);
            uint32_t it = 0, lt = 0;   // k-block counter, local tile counter
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
                int m0, n0, kb0, nkb, split;
                decode(t, m0, n0, kb0, nkb, split);
                const uint32_t buf = lt & 1u;
                mbar_wait(smem_u32(&tmem_empty_bar[buf]), ((lt >> 1) & 1u) ^ 1u);   // epilogue drained this buffer
                tc_fence_after();
                const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn;
                for (int kb = 0; kb < nkb; ++kb, ++it) {
                    const int s = it % stages;
                    mbar_wait(smem_u32(&full_bar[s]), (it / stages) & 1u);
                    tc_fence_after();
                    const uint32_t a_base = smem_base + (uint32_t)s * p.stage_bytes;
                    const uint32_t b_base = a_base + kATile;
#pragma unroll
                    for (int j = 0; j < KB / kUmmaK; ++j) {
                        // K-major: 32 B further along the swizzled 128-byte row per K = 8.
                        // MN-major: 8 k-rows = two 4-row atoms (SBO = 512 B apart) per K = 8.
                        const uint64_t a_desc = A_MN ? make_desc(a_base + j * 1024u, kGroup, 512u, 1u)
                                                     : make_desc(a_base + j * 32u, 16u, kSboK, kLayoutK);
                        const uint64_t b_desc = B_MN ? make_desc(b_base + j * 1024u, kGroup, 512u, 1u)
                                                     : make_desc(b_base + j * 32u, 16u, kSboK, kLayoutK);
                        umma_tf32(tacc, a_desc, b_desc, idesc, (kb > 0 || j > 0) ? 1u : 0u);
                        if constexpr (SPLIT) {
                            const uint64_t a_lo = A_MN ? make_desc(a_base + lo_off + j * 1024u, kGroup, 512u, 1u)
                                                       : make_desc(a_base + lo_off + j * 32u, 16u, kSboK, kLayoutK);
                            const uint64_t b_lo = B_MN ? make_desc(b_base + lo_off + j * 1024u, kGroup, 512u, 1u)
                                                       : make_desc(b_base + lo_off + j * 32u, 16u, kSboK, kLayoutK);
                            umma_tf32(tacc, a_lo, b_desc, idesc, 1u);
                            umma_tf32(tacc, a_desc, b_lo, idesc, 1u);
                        }
                    }
                    umma_commit(smem_u32(&empty_bar[s]));       // frees the smem stage when these MMAs retire
                }
                umma_commit(smem_u32(&tmem_full_bar[buf]));     // accumulator of this tile complete
            }
        }
    } else {
        // ===== epilogue: TMEM -> registers -> global, overlapped with the next tile's MMAs =====
        const int q = warp & 3;                 // TMEM lane quarter this warp may access
        const float alpha = p.alpha;
        const float* __restrict__ bias = p.bias;
        const int Nv = p.N;
        const bool vec_ok = (p.ldc & 3) == 0;
        float* const strip = stat_strip[q];
        if (p.stat_part)
            for (int t = lane; t < 2 * kStatCols; t += 32) strip[t] = 0.f;
        uint32_t lt = 0;
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x, ++lt) {
            int m0, n0, kb0, nkb, split;
            decode(t, m0, n0, kb0, nkb, split);
            const uint32_t buf = lt & 1u;
            const int m = m0 + q * 32 + lane;
            const bool row_ok = m < p.M;
            float* crow = p.C + (long)split * p.split_stride + (long)m * p.ldc;
            const uint32_t tacc = tmem_base + buf * (uint32_t)p.bn + ((uint32_t)(q * 32) << 16);
            if (nkb > 0) {
                mbar_wait(smem_u32(&tmem_full_bar[buf]), (lt >> 1) & 1u);
                tc_fence_after();
            }
            const int ncols = min(p.bn, Nv - n0);                 // valid columns of this tile
            const int nchunks = (min(p.bn, max(ncols, 0)) + 15) / 16;
            uint32_t rcur[16], rnext[16];
            if (nkb > 0 && nchunks > 0) { tmem_ld16_nowait(tacc, rcur); tmem_ld_wait(); }
            for (int ch = 0; ch < nchunks; ++ch) {
                const int c0 = n0 + ch * 16;
                if (nkb > 0 && ch + 1 < nchunks) tmem_ld16_nowait(tacc + (uint32_t)(ch + 1) * 16u, rnext);   // in flight during the stores
                float v[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) v[i] = nkb > 0 ? __uint_as_float(rcur[i]) : 0.f;
                const bool full = c0 + 16 <= Nv;
                float bv[16];
#pragma unroll
                for (int i = 0; i < 16; ++i) bv[i] = 0.f;
                if (bias) {
                    if (full) {
#pragma unroll
                        for (int i = 0; i < 16; i += 4) {
                            const float4 tb = __ldg(reinterpret_cast<const float4*>(bias + c0 + i));
                            bv[i] = tb.x; bv[i + 1] = tb.y; bv[i + 2] = tb.z; bv[i + 3] = tb.w;
                        }
                    } else {
#pragma unroll
                        for (int i = 0; i < 16; ++i)
                            if (c0 + i < Nv) bv[i] = __ldg(bias + c0 + i);
                    }
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
                if (nkb > 0 && ch + 1 < nchunks) {
                    tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 16; ++i) rcur[i] = rnext[i];
                }
            }
            // this warp no longer reads the buffer: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&tmem_empty_bar[buf]));
        }
        if (p.stat_part) {
            __syncwarp();
            stat_flush(strip, p.stat_part + ((size_t)blockIdx.x * 4 + q) * 2 * Nv, Nv, lane);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(p.tmem_cols) : "memory");
    }
}

}  // namespace tc
}  // namespace nvsm
