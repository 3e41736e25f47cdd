// Roofline denominators measured on the box the bench runs on (bench.py `roofline.l2`): what a plain row gather
// out of an L2-resident table sustains. The gather-type kernels of the step (gather_mean, score, the two pull
// updates) read every table row ~10x per batch out of L2, so their ceiling is this number, not the HBM copy peak.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

namespace nvsm {

// One warp per item: sums `rows_per_item` pseudo-random rows of `row_vec4` float4 each (lane l owns float4 l, l + 32,
// ...; up to K per lane) and writes one float4 per lane and chunk (so the loads cannot be elided). U rows in flight.
template <int K, int U>
__global__ void __launch_bounds__(256) l2_gather_probe_kernel(const float4* __restrict__ table, long num_rows,
                                                              int row_vec4, int rows_per_item, long items,
                                                              float4* __restrict__ out, unsigned seed) {
    const int lane = threadIdx.x & 31;
    const long warp0 = ((long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long nwarps = ((long)gridDim.x * blockDim.x) >> 5;
    for (long item = warp0; item < items; item += nwarps) {
        float4 acc[K];
#pragma unroll
        for (int k = 0; k < K; ++k) acc[k] = make_float4(0.f, 0.f, 0.f, 0.f);
        for (int r0 = 0; r0 < rows_per_item; r0 += U) {
            float4 v[U][K];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                // row id: a per-(item, reference) hash, identical in all lanes of the warp
                unsigned h = (unsigned)(item * 0x9E3779B1u) ^ ((unsigned)(r0 + u) * 0x85EBCA77u) ^ seed;
                h ^= h >> 15; h *= 0x2C1B3C6Du; h ^= h >> 12; h *= 0x297A2D39u; h ^= h >> 15;
                const long row = (long)(h % (unsigned)num_rows);
                const float4* src = table + row * row_vec4;
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    const int c = lane + 32 * k;
                    v[u][k] = (r0 + u < rows_per_item && c < row_vec4) ? __ldg(src + c) : make_float4(0.f, 0.f, 0.f, 0.f);
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u)
#pragma unroll
                for (int k = 0; k < K; ++k) {
                    acc[k].x += v[u][k].x; acc[k].y += v[u][k].y; acc[k].z += v[u][k].z; acc[k].w += v[u][k].w;
                }
        }
#pragma unroll
        for (int k = 0; k < K; ++k) {
            const int c = lane + 32 * k;
            if (c < row_vec4) out[item * row_vec4 + c] = acc[k];
        }
    }
}

// (volatile: the sweeps re-read the same addresses and must not be merged by the compiler)
__device__ __forceinline__ float4 ld_cg_volatile(const float4* p) {
    float4 v;
    asm volatile("ld.global.cg.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}

// L2-resident streaming read: every thread walks the whole `n`-float4 buffer `reps` times with 8 independent 128-bit
// loads in flight (the buffer is sized to stay in L2, so after the first sweep every load is an L2 hit). GB/s of this
// is the SM <-> L2 ceiling every gather-type kernel of the step sits under.
__global__ void __launch_bounds__(256) l2_read_probe_kernel(const float4* __restrict__ in, long n, int reps,
                                                            float4* __restrict__ out) {
    const long stride = (long)gridDim.x * blockDim.x;
    const long t0 = (long)blockIdx.x * blockDim.x + threadIdx.x;
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int r = 0; r < reps; ++r) {
        long i = t0;
        for (; i + 7 * stride < n; i += 8 * stride) {
            float4 v[8];
#pragma unroll
            for (int u = 0; u < 8; ++u) v[u] = ld_cg_volatile(in + i + u * stride);
#pragma unroll
            for (int u = 0; u < 8; ++u) { acc.x += v[u].x; acc.y += v[u].y; acc.z += v[u].z; acc.w += v[u].w; }
        }
        for (; i < n; i += stride) { const float4 v = ld_cg_volatile(in + i); acc.x += v.x; acc.y += v.y; acc.z += v.z; acc.w += v.w; }
    }
    if (acc.x == 123.456f) out[t0] = acc;   // never true for a zero-filled buffer: keeps the loads alive
}

// Streaming read + write of `n` float4 (HBM copy peak cross-check against MEASURED_PEAKS.json).
__global__ void __launch_bounds__(256) stream_copy_probe_kernel(const float4* __restrict__ in, float4* __restrict__ out, long n) {
    const long stride = (long)gridDim.x * blockDim.x;
    for (long i = (long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) __stcs(out + i, __ldcs(in + i));
}

}  // namespace nvsm
