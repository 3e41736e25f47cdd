// Shared device helpers for the sm_100a NVSM/LSE kernels.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

namespace nvsm {

typedef long idx_t;  // reference: `typedef long int32` (include/cuNVSM/base.h:28)

constexpr int kWarp = 32;
constexpr unsigned kFull = 0xffffffffu;

// Programmatic dependent launch (sm_90+). A kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization
// may become resident while the kernel in front of it in the stream is still draining: pdl_wait() blocks until that
// grid has completed and its writes are visible (call it before the first dependent global access; everything above
// it -- barrier init, TMEM allocation, shared-memory zeroing -- overlaps the predecessor's tail). pdl_launch_dependents()
// in the predecessor lets the dependent grid start being scheduled once every block of the predecessor has issued it
// (or exited). Both are no-ops for launches without the attribute / without a dependent.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

// ---- vector access: VEC == 4 (16-byte rows) or VEC == 1 (any dim) -----------------
template <int VEC>
__device__ __forceinline__ void load_vec(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = *reinterpret_cast<const float4*>(p);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = *p;
    }
}

// Read-only (non-coherent) path for data that is never written by the running kernel.
template <int VEC>
__device__ __forceinline__ void load_vec_ro(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldg(p);
    }
}

// Streaming (evict-first) access for data touched once per kernel, so that it does not push re-used lines
// (gathered rows) out of L2.
template <int VEC>
__device__ __forceinline__ void load_vec_cs(const float* __restrict__ p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldcs(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldcs(p);
    }
}

// L2-coherent load (bypasses L1): data another CTA of the running kernel has just published.
template <int VEC>
__device__ __forceinline__ void load_vec_cg(const float* p, float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        const float4 t = __ldcg(reinterpret_cast<const float4*>(p));
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        v[0] = __ldcg(p);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec_cs(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        __stcs(reinterpret_cast<float4*>(p), make_float4(v[0], v[1], v[2], v[3]));
    } else {
        __stcs(p, v[0]);
    }
}

template <int VEC>
__device__ __forceinline__ void store_vec(float* __restrict__ p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
        *p = v[0];
    }
}

// Fire-and-forget reduction into global memory: one 16-byte REDG.ADD.F32x4 per call when
// VEC == 4 (sm_90+), else a scalar RED.
template <int VEC>
__device__ __forceinline__ void red_add_vec(float* p, const float (&v)[VEC]) {
    if constexpr (VEC == 4) {
        asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v[0]), "f"(v[1]),
                     "f"(v[2]), "f"(v[3])
                     : "memory");
    } else {
        asm volatile("red.global.add.f32 [%0], %1;" ::"l"(p), "f"(v[0]) : "memory");
    }
}

// Round-to-nearest fp32 -> tf32 (10-bit mantissa), returned as an fp32 whose low 13 bits are zero.
// tcgen05 kind::tf32 TRUNCATES the operands it reads (measured: scripts/probe_tf32_rounding.py), a
// systematic toward-zero bias; operands that only feed the GEMMs are therefore pre-rounded by
// their producers, which makes the hardware conversion exact and the error unbiased.
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}

// sqrt / divide as the reference's release build compiles them (-use_fast_math => --prec-sqrt=false
// --prec-div=false, CMakeLists.txt:71-73 of the reference): sqrt.approx / div.approx, one MUFU each instead of
// the ~20-instruction IEEE sequences.
__device__ __forceinline__ float fast_sqrt(float x) {
    float r;
    asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x));
    return r;
}
__device__ __forceinline__ float fast_div(float a, float b) {
    float r;
    asm("div.approx.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(a), "f"(b));
    return r;
}

// Block-level aggregation of per-id atomics (bucket build, per-object scalar accumulators): ids hash into a small
// direct-mapped shared-memory table; see ref_count_kernel.
constexpr int kAggSlots = 512;
constexpr int kAggThreads = 1024;
__device__ __forceinline__ int agg_slot(int id) { return (int)(((unsigned)id * 2654435761u) >> 23); }   // 9 bits

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    return v;
}

// ---- activation (+ optional batch-norm) applied to the stored pre-activation Z ----
// Z holds T.P (+ b when batch-norm is off). With batch-norm: y = f((z - mean) * invstd + b)
// (cuDNN per-activation, gamma == 1; cpp/cudnn_utils.cu:82-129). f is tanh or the
// reference's clip whose bounds sit one ulp outside [-1, 1]
// (include/cuNVSM/cuda_utils.h:86-147).
struct ActParams {
    int nonlinearity;  // 0 tanh, 1 hard_tanh
    int use_bn;
    float clip_min, clip_max;
    const float* mean;    // [dd] (batch-norm only)
    const float* invstd;  // [dd]
    const float* bias;    // [dd]
};

__device__ __forceinline__ float act_forward(const ActParams& a, float z, int col, float& xhat) {
    float t = z;
    xhat = 0.f;
    if (a.use_bn) {
        xhat = (z - __ldg(a.mean + col)) * __ldg(a.invstd + col);
        t = xhat + __ldg(a.bias + col);
    }
    return a.nonlinearity == 0 ? tanhf(t) : fminf(fmaxf(t, a.clip_min), a.clip_max);
}

__device__ __forceinline__ float act_deriv(const ActParams& a, float y) {
    return a.nonlinearity == 0 ? (1.0f - y * y) : ((y > a.clip_min && y < a.clip_max) ? 1.0f : 0.0f);
}

}  // namespace nvsm
