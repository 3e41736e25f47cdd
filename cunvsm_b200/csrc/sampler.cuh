// Device-side negative sampler, bit-exact with the reference's host loop
// (UniformLabelGenerator::generate, cpp/labels.cu:3-22 -> generate_random_indexes,
// include/cuNVSM/cuda_utils.h:24-33): for every instance, the positive label followed by z draws of
// std::uniform_int_distribution<long>(0, D-1) on the shared std::minstd_rand0.
//
// libstdc++'s distribution on this engine (range [1, 2^31-2], not a power of two) is the classic
// down-scaling loop:   scaling = 2147483645 / D;  past = D * scaling;
//                      do ret = rng() - 1; while (ret >= past);   return ret / scaling;
// i.e. a stream of candidates with (rare) rejections. minstd_rand0 is the Lehmer generator
// x_{k+1} = 16807 x_k mod (2^31-1), so candidate k is a^(k+1) x_0 mod m: every thread jumps to the
// start of its chunk with a modular power, walks the chunk, and a prefix sum over the per-chunk
// accept counts compacts the accepted candidates into draw order. The engine state after the call is
// the candidate that produced the last accepted draw — exactly where the host loop would stop.
#pragma once

#include "common.cuh"

namespace nvsm {

constexpr unsigned long long kLehmerM = 2147483647ull;   // 2^31 - 1
constexpr unsigned long long kLehmerA = 16807ull;
constexpr int kSamplerChunk = 16;                         // candidates per thread (64: 8000 threads for a C2 batch, ncu 22 us; 16: 32000)

__device__ __forceinline__ unsigned int lehmer_mulmod(unsigned long long a, unsigned long long b) {
    unsigned long long p = a * b;                         // < 2^62
    p = (p & kLehmerM) + (p >> 31);
    p = (p & kLehmerM) + (p >> 31);
    return (unsigned int)(p >= kLehmerM ? p - kLehmerM : p);
}

__device__ __forceinline__ unsigned int lehmer_pow(unsigned long long e) {   // a^e mod m
    unsigned int result = 1, base = (unsigned int)kLehmerA;
    while (e) {
        if (e & 1ull) result = lehmer_mulmod(result, base);
        base = lehmer_mulmod(base, base);
        e >>= 1;
    }
    return result;
}

struct SamplerParams {
    const unsigned int* state_in;   // x_0
    unsigned int* state_out;        // state after the last consumed candidate
    long num_draws;                 // N = B * z
    long num_candidates;            // T >= N + rejections (multiple of kSamplerChunk)
    unsigned int scaling, past;
    int z, R;
    const idx_t* labels;            // [B]
    idx_t* ids;                     // [B * R]
    int* counts;                    // [T / chunk] accepted per chunk
    const int* offsets;             // exclusive scan of counts (+ total at [nchunks])
    int* error_flag;                // set when T was too small (caller falls back to the host sampler)
};

__global__ void __launch_bounds__(256) sampler_count_kernel(const SamplerParams p) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nchunks = p.num_candidates / kSamplerChunk;
    if (t >= nchunks) return;
    unsigned int x = lehmer_mulmod(lehmer_pow((unsigned long long)t * kSamplerChunk), *p.state_in);
    int c = 0;
#pragma unroll 8
    for (int k = 0; k < kSamplerChunk; ++k) {
        x = lehmer_mulmod(x, kLehmerA);
        c += (x - 1u) < p.past;
    }
    p.counts[t] = c;
}

__global__ void __launch_bounds__(256) sampler_fill_kernel(const SamplerParams p) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nchunks = p.num_candidates / kSamplerChunk;
    // positives
    const long B = p.num_draws / max(p.z, 1);
    for (long i = t; i < (p.z > 0 ? B : 0); i += (long)gridDim.x * blockDim.x) p.ids[i * p.R] = p.labels[i];
    if (t >= nchunks) return;
    long g = p.offsets[t];
    if (t == nchunks - 1 && p.offsets[nchunks] < p.num_draws) *p.error_flag = 1;
    if (g >= p.num_draws) return;
    unsigned int x = lehmer_mulmod(lehmer_pow((unsigned long long)t * kSamplerChunk), *p.state_in);
    for (int k = 0; k < kSamplerChunk && g < p.num_draws; ++k) {
        x = lehmer_mulmod(x, kLehmerA);
        const unsigned int ret = x - 1u;
        if (ret < p.past) {
            p.ids[(g / p.z) * p.R + 1 + (g % p.z)] = (idx_t)(ret / p.scaling);
            if (g == p.num_draws - 1) *p.state_out = x;
            ++g;
        }
    }
}

// Rare-rejection path (what every table below ~2^27 rows takes: the rejection probability of libstdc++'s down-scaling loop is
// (2^31 - 2) mod D / (2^31 - 2), 1.6e-5 for D = 50k, i.e. ~8 rejected candidates in a 512 000-draw batch): instead of
// a per-chunk count array + three scan launches + a total launch, the count kernel appends the few chunks that saw a
// rejection to a short list and the fill kernel derives its chunk's offset from that list:
//   offset(t) = chunk * t - sum_{listed chunks c < t} rejected(c).
// Two launches instead of six on the copy stream, which runs under the main stream's gather / forward GEMM.
constexpr int kRejListCap = 2048;

struct RejList {
    int2* items;        // (chunk, rejected candidates in it), unordered
    int* count;         // entries appended by THIS call
    int* count_next;    // the other parity's counter: reset by the fill kernel for the next call
};

__global__ void __launch_bounds__(256) sampler_count_list_kernel(const SamplerParams p, const RejList rl) {
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nchunks = p.num_candidates / kSamplerChunk;
    if (t >= nchunks) return;
    unsigned int x = lehmer_mulmod(lehmer_pow((unsigned long long)t * kSamplerChunk), *p.state_in);
    int rej = 0;
#pragma unroll
    for (int k = 0; k < kSamplerChunk; ++k) {
        x = lehmer_mulmod(x, kLehmerA);
        rej += (x - 1u) >= p.past;
    }
    if (rej > 0) {
        const int slot = atomicAdd(rl.count, 1);
        if (slot < kRejListCap) rl.items[slot] = make_int2((int)t, rej);
    }
}

__global__ void __launch_bounds__(256) sampler_fill_list_kernel(const SamplerParams p, const RejList rl) {
    __shared__ int2 s_items[256];
    const long t = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long nchunks = p.num_candidates / kSamplerChunk;
    // positives
    const long B = p.num_draws / max(p.z, 1);
    for (long i = t; i < (p.z > 0 ? B : 0); i += (long)gridDim.x * blockDim.x) p.ids[i * p.R] = p.labels[i];
    const int listed = *rl.count;
    const int n = min(listed, kRejListCap);
    long before = 0, total_rej = 0;   // rejected candidates in chunks before mine / overall
    for (int base = 0; base < n; base += 256) {
        __syncthreads();
        if (base + (int)threadIdx.x < n) s_items[threadIdx.x] = rl.items[base + threadIdx.x];
        __syncthreads();
        const int m = min(256, n - base);
        for (int k = 0; k < m; ++k) {
            total_rej += s_items[k].y;
            if (s_items[k].x < t) before += s_items[k].y;
        }
    }
    if (t == 0) {
        if (listed > kRejListCap || p.num_candidates - total_rej < p.num_draws) *p.error_flag = 1;
        *rl.count_next = 0;
    }
    if (t >= nchunks) return;
    long g = t * kSamplerChunk - before;
    if (g >= p.num_draws) return;
    unsigned int x = lehmer_mulmod(lehmer_pow((unsigned long long)t * kSamplerChunk), *p.state_in);
    for (int k = 0; k < kSamplerChunk && g < p.num_draws; ++k) {
        x = lehmer_mulmod(x, kLehmerA);
        const unsigned int ret = x - 1u;
        if (ret < p.past) {
            p.ids[(g / p.z) * p.R + 1 + (g % p.z)] = (idx_t)(ret / p.scaling);
            if (g == p.num_draws - 1) *p.state_out = x;
            ++g;
        }
    }
}

// offsets[nchunks] = total accepted (scan_add_kernel wrote its own `total` argument there)
__global__ void sampler_total_kernel(const int* __restrict__ counts, int* __restrict__ offsets, long nchunks) {
    offsets[nchunks] = offsets[nchunks - 1] + counts[nchunks - 1];
}

// Skewed negatives (BASELINE configs[4]: Zipf-skewed sampling; the reference ships only the uniform generator and
// leaves LabelGenerator, include/cuNVSM/labels.h:7-18, as the plug point). Inverse-CDF sampling over a caller-supplied
// cumulative distribution cdf[0..D-1] (non-decreasing, cdf[D-1] = 1): draw j = i * z + r consumes exactly one engine
// output x_j = a^(j+1) x_0 mod m,  u = (x_j - 1) / (m - 1) in [0, 1),  id = min{k : cdf[k] > u}. No rejections, so the
// jump-ahead makes every draw independent; the host loop (nvsm_generate_labels_cdf) is the same arithmetic in order.
__global__ void __launch_bounds__(256) sampler_cdf_kernel(const unsigned int* __restrict__ state_in,
                                                          unsigned int* __restrict__ state_out,
                                                          const idx_t* __restrict__ labels, idx_t* __restrict__ ids,
                                                          long num_draws, int z, const double* __restrict__ cdf,
                                                          long num_objects) {
    const long j = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= num_draws) return;
    const long i = j / z;
    const int r = (int)(j - i * z);
    const unsigned int x = lehmer_mulmod(lehmer_pow((unsigned long long)j + 1ull), *state_in);
    const double u = (double)(x - 1u) / 2147483646.0;
    long lo = 0, hi = num_objects;
    while (lo < hi) {
        const long mid = (lo + hi) >> 1;
        if (__ldg(cdf + mid) <= u) lo = mid + 1; else hi = mid;
    }
    idx_t* const row = ids + i * (z + 1);
    row[1 + r] = (idx_t)min(lo, num_objects - 1);
    if (r == 0) row[0] = labels[i];
    if (j == num_draws - 1) *state_out = x;
}

__global__ void sampler_copy_labels_kernel(const idx_t* __restrict__ labels, long B, idx_t* __restrict__ ids) {
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < B) ids[i] = labels[i];   // z == 0: R == 1
}

}  // namespace nvsm
