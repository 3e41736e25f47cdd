// Device side of the stand-alone operator entry points (nvsm_op_* in include/nvsm_b200.h): the reference's
// Representations / Transform / *Storage / *GradientUpdater / BatchNormalization classes call them one operation at a
// time on caller-owned tensors (include/cuNVSM/{params,storage,updates,cudnn_utils}.h of this repo). The fused
// training step does not go through here; these kernels favour generality (any dim, any window, optional weights) over
// the last percent -- the gather / scatter / dense-optimiser kernels of the step (kernels.cuh) are reused where their
// signature already is the generic one.
#pragma once

#include "common.cuh"

namespace nvsm {

// y[i, c] = (x[i, c] - mean[c]) * invstd[c] + bias[c]      (cuDNN per-activation batch-norm, gamma == 1,
// cpp/cudnn_utils.cu:107-124). y may alias x.
__global__ void __launch_bounds__(256) op_bn_apply_kernel(const float* x, const float* __restrict__ mean,
                                                          const float* __restrict__ invstd, const float* __restrict__ bias,
                                                          long rows, int C, float* y) {
    const long total = rows * C;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        y[t] = (x[t] - mean[c]) * invstd[c] + bias[c];
    }
}

// sums[c] += sum_i dy[i, c];  sums[C + c] += sum_i dy[i, c] * xhat[i, c]      (cpp/cudnn_utils.cu:158-177)
__global__ void __launch_bounds__(256) op_bn_backward_sums_kernel(const float* __restrict__ dy, const float* __restrict__ x,
                                                                  const float* __restrict__ mean,
                                                                  const float* __restrict__ invstd, long rows, int C,
                                                                  double* __restrict__ sums) {
    const int tpr = min(C, (int)blockDim.x), rpp = blockDim.x / tpr;
    const int tr = threadIdx.x / tpr, tc = threadIdx.x % tpr;
    if (tr >= rpp) return;
    const long rows_per_block = (rows + gridDim.x - 1) / gridDim.x;
    const long r0 = (long)blockIdx.x * rows_per_block, r1 = min(rows, r0 + rows_per_block);
    for (int c = tc; c < C; c += tpr) {
        const float mu = mean[c], is = invstd[c];
        float s = 0.f, q = 0.f;
        for (long r = r0 + tr; r < r1; r += rpp) {
            const float g = dy[r * C + c];
            s += g;
            q += g * ((x[r * C + c] - mu) * is);
        }
        if (r0 + tr < r1) {
            atomicAdd(sums + c, (double)s);
            atomicAdd(sums + C + c, (double)q);
        }
    }
}

// dx = invstd * (dy - sum_dy / N - xhat * sum_dy_xhat / N); dbias = sum_dy. dx may alias dy.
__global__ void __launch_bounds__(256) op_bn_backward_apply_kernel(const float* dy, const float* __restrict__ x,
                                                                   const float* __restrict__ mean,
                                                                   const float* __restrict__ invstd,
                                                                   const double* __restrict__ sums, long rows, int C,
                                                                   float* dx, float* __restrict__ dbias) {
    const long total = rows * C;
    const float inv_n = 1.0f / (float)rows;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const int c = (int)(t % C);
        const float is = invstd[c];
        const float xh = (x[t] - mean[c]) * is;
        const float sb = (float)sums[c], sg = (float)sums[C + c];
        dx[t] = is * (dy[t] - sb * inv_n - xh * (sg * inv_n));
        if (t < C) dbias[t] = (float)sums[t];
    }
}

// update_dense (include/cuNVSM/storage_inl.h:4-32): param = param * (1 - lambda * lr) + op(grad) * lr, op in
// {identity, square}.
__global__ void __launch_bounds__(256) op_update_dense_kernel(float* __restrict__ param, const float* __restrict__ grad,
                                                              long n, float decay, float lr, int square) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const float g = grad[t];
        param[t] = param[t] * decay + (square ? g * g : g) * lr;
    }
}

// x[t] = value
__global__ void __launch_bounds__(256) op_fill_kernel(float* __restrict__ x, long n, float value) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) x[t] = value;
}

// One sparse gradient descriptor (RepresentationsStorage::SingleGradientType, include/cuNVSM/storage.h:64-69) in row-major
// device memory: grad [count, dim], ids [count * window], weights [count * window] or null.
//   target[ids[x, y], :] += scale * factor(x) * wt[x, y] * grad[x, :]          (update_repr_kernel, cpp/storage.cu:37-49)
// factor(x) = 1 / sqrt(mean_y acc[ids[x, y]] + eps) when acc != null (adagrad_update_kernel, cpp/updates_adagrad.cu:83-97).
__global__ void __launch_bounds__(256) op_scatter_rows_kernel(const float* __restrict__ grad, const idx_t* __restrict__ ids,
                                                              const float* __restrict__ wts, long count, int window, int dim,
                                                              float* __restrict__ target, float scale,
                                                              const float* __restrict__ acc, float eps) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long x = warp0; x < count; x += nwarps) {
        float factor = 1.0f;
        if (acc) {
            float a = 0.f;
            for (int y = 0; y < window; ++y) a += __ldg(acc + __ldg(ids + x * window + y));
            a /= (float)window;
            factor = 1.0f / sqrtf(a + eps);
        }
        for (int k = lane; k < dim; k += kWarp) {
            const float g = __ldg(grad + x * dim + k) * factor;
            for (int y = 0; y < window; ++y) {
                const float wt = wts ? __ldg(wts + x * window + y) : 1.0f;
                atomicAdd(target + __ldg(ids + x * window + y) * dim + k, (scale * wt) * g);
            }
        }
    }
}

// grad[x, :] /= sqrt(mean_y acc[ids[x, y]] + eps), in place (adagrad_update_kernel, cpp/updates_adagrad.cu:83-97): the
// reference rewrites the descriptor's gradient before the SGD scatter, and its tests read it back.
__global__ void __launch_bounds__(256) op_adagrad_scale_grad_kernel(float* __restrict__ grad, const idx_t* __restrict__ ids,
                                                                    long count, int window, int dim,
                                                                    const float* __restrict__ acc, float eps) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    for (long x = warp0; x < count; x += nwarps) {
        float a = 0.f;
        for (int y = 0; y < window; ++y) a += __ldg(acc + __ldg(ids + x * window + y));
        a /= (float)window;
        const float den = sqrtf(a + eps);
        for (int k = lane; k < dim; k += kWarp) grad[x * dim + k] = grad[x * dim + k] / den;
    }
}

// Dense optimiser of the projection with the reference's in-place gradient semantics ({SGD,Adagrad,Adam}Transform-
// GradientUpdater::update): Adagrad leaves g / sqrt(acc + eps) in the gradient tensors, Adam leaves the step direction
// bc m / (sqrt(v) + eps) (after g -= lambda T for the matrix); both are then applied by TransformStorage::update. The
// bias is never regularised and its Adam moments never decay (cpp/storage.cu:222-227, cpp/updates_tests.cu:352-366).
struct OpTransformUpdate {
    float* T; float* b; float* gT; float* gb;
    long nT; int nb;
    int method;
    float lr, lambda;
    float* aT; float* ab; float* vT; float* vb;
    float s1, lr1, s2, lr2, bc, eps;
};
__global__ void __launch_bounds__(256) op_transform_update_kernel(const OpTransformUpdate p) {
    const long total = p.nT + p.nb;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const bool is_bias = t >= p.nT;
        const long k = is_bias ? t - p.nT : t;
        float* th = is_bias ? p.b + k : p.T + k;
        float* gp = is_bias ? p.gb + k : p.gT + k;
        float g = *gp;
        const float lam = is_bias ? 0.f : p.lambda;
        if (p.method == 0) {
            *th = *th * (1.0f - lam * p.lr) + g * p.lr;
        } else if (p.method == 1) {
            float* a = is_bias ? p.ab + k : p.aT + k;
            const float acc = *a + g * g;
            *a = acc;
            g = g / sqrtf(acc + p.eps);
            *gp = g;
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
            *gp = g;
            *th = *th + g * p.lr;
        }
    }
}

// acc[ids[x, y]] += scale * wt[x, y] * msq[x]       (the scalar moments: update_repr_kernel on a 1 x objects table)
__global__ void op_scatter_scalar_kernel(const idx_t* __restrict__ ids, const float* __restrict__ wts,
                                         const float* __restrict__ msq, long total, int window, float scale,
                                         float* __restrict__ acc) {
    const long c = (long)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= total) return;
    atomicAdd(acc + ids[c], scale * (wts ? wts[c] : 1.0f) * msq[c / window]);
}

// Sparse Adam: grad[x, :] = bc * mean_y m[ids[x, y], :] / (sqrt(mean_y v[ids[x, y]]) + eps)
// (adam_sparse_update_kernel, cpp/updates_adam.cu:132-151) -- overwrites the descriptor's gradient like the reference.
__global__ void __launch_bounds__(256) op_adam_sparse_grad_kernel(const idx_t* __restrict__ ids, long count, int window,
                                                                  int dim, const float* __restrict__ m,
                                                                  const float* __restrict__ v, float bc, float eps,
                                                                  float* __restrict__ grad) {
    const int lane = threadIdx.x & 31;
    const long warp0 = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const long nwarps = (long)gridDim.x * (blockDim.x >> 5);
    const float fw = (float)window;
    for (long x = warp0; x < count; x += nwarps) {
        float av = 0.f;
        for (int y = 0; y < window; ++y) av += __ldg(v + __ldg(ids + x * window + y));
        av /= fw;
        const float den = sqrtf(av) + eps;
        for (int k = lane; k < dim; k += kWarp) {
            float am = 0.f;
            for (int y = 0; y < window; ++y) am += m[__ldg(ids + x * window + y) * dim + k];
            grad[x * dim + k] = bc * (am / fw) / den;
        }
    }
}

// agg = (agg - lambda * theta)^2 elementwise: the dense second-moment input of DENSE_UPDATE_DENSE_VARIANCE
// (cpp/updates_adam.cu:251-283).
__global__ void __launch_bounds__(256) op_full_adam_variance_input_kernel(float* __restrict__ agg, const float* __restrict__ theta,
                                                                          long n, float lambda) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) {
        const float g = agg[t] + (-lambda * theta[t]);
        agg[t] = g * g;
    }
}

// theta += lr * bc * m / (sqrt(v) + eps), v per element (full Adam) or per object (dense-update Adam); decay on theta.
__global__ void __launch_bounds__(256) op_adam_apply_kernel(float* __restrict__ theta, const float* __restrict__ m,
                                                            const float* __restrict__ v, long num_objects, int dim,
                                                            int v_per_object, float decay, float lr, float bc, float eps) {
    const long total = num_objects * dim;
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (long)gridDim.x * blockDim.x) {
        const float vv = v_per_object ? v[t / dim] : v[t];
        theta[t] = theta[t] * decay + ((m[t] / (sqrtf(vv) + eps)) * bc) * lr;
    }
}

// m += -l * theta (the L2 term of full Adam's first moment, cpp/updates_adam.cu:199-213)
__global__ void __launch_bounds__(256) op_axpy_kernel(float* __restrict__ y, const float* __restrict__ x, long n, float a) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x) y[t] += a * x[t];
}

// y = f(x): tanh or the reference's clip (bounds one ulp outside [-1, 1])
__global__ void __launch_bounds__(256) op_activation_kernel(const float* x, long n, int nonlinearity, float clip_min,
                                                            float clip_max, float* y) {
    for (long t = (long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long)gridDim.x * blockDim.x)
        y[t] = nonlinearity == 0 ? tanhf(x[t]) : fminf(fmaxf(x[t], clip_min), clip_max);
}

}  // namespace nvsm
