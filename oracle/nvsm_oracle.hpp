// ORACLE — TEST INFRASTRUCTURE ONLY.
//
// CPU restatement of the cuNVSM TextEntity (LSE / NVSM) training step. This
// file is the *checker* for the CUDA path in cunvsm_b200/csrc; it is never on
// the product path. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may build, load or call it.
//
// Parity status: PINNED. The restatement is checked (tests/test_oracle_golden.py)
// against the known-answer vectors of the reference's own test-suite:
//   cpp/model_tests.cu:52-123   (gather-mean), :153-243 (scatter + decay),
//   :245-275 (update_dense), :277-339 (Transform), :341-466 (Transform_backward,
//   the whole forward/backward chain, seed 10), :468-548 (BN + tanh forward),
//   cpp/cudnn_utils_tests.cu:115-177 (BN fwd/bwd), cpp/updates_tests.cu (every
//   optimiser variant), cpp/cuda_utils_tests.cu:8-21 (truncated sigmoid).
// The reference itself cannot be compiled here (its arithmetic library,
// cvangysel/device_matrix@master, is un-vendored) so there is
// no oracle/_ref.
//
// Conventions: every matrix of the reference is column-major rows=feature dim,
// cols=objects/instances, i.e. in memory one contiguous row of `dim` values per
// object. We store the same memory image: W[V][d_w], E[D][d_d],
// T[d_w][d_d] (reference: d_d x d_w column-major, cpp/storage.cu:185-191),
// P[B][d_w], Y[B][d_d], gE[B*R][d_d], gP[B][d_w], gT[d_w][d_d].
#pragma once

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <limits>
#include <memory>
#include <random>
#include <vector>

namespace oracle {

typedef long idx_t;                 // include/cuNVSM/base.h:28  (int32 is `long`)
typedef std::minstd_rand0 RNG;      // include/cuNVSM/base.h:36

enum Nonlinearity { TANH = 0, HARD_TANH = 1 };       // proto/nvsm.proto:11-14
enum UpdateMethod { SGD = 0, ADAGRAD = 1, ADAM = 2 };  // proto/nvsm.proto:41-45
enum AdamMode { ADAM_NONE = 0, SPARSE = 1, DENSE_UPDATE = 2,
                DENSE_UPDATE_DENSE_VARIANCE = 3 };    // proto/nvsm.proto:51-56

// ---------------------------------------------------------------------------
// RNG-driven host pieces.
// ---------------------------------------------------------------------------

// include/cuNVSM/cuda_utils.h:24-33 — a fresh distribution object per draw.
inline void generate_random_indexes(idx_t max, size_t num, RNG* rng, idx_t* out) {
    for (size_t i = 0; i < num; ++i)
        out[i] = std::uniform_int_distribution<idx_t>(0, max - 1)(*rng);
}

// cpp/labels.cu:3-22 — positive label first, then z uniform draws over all D
// (the true label is not excluded).
inline void generate_labels(const idx_t* labels, size_t num_labels, size_t z,
                            idx_t num_objects, RNG* rng, idx_t* out) {
    const size_t R = z + 1;
    for (size_t i = 0; i < num_labels; ++i) {
        out[i * R] = labels[i];
        generate_random_indexes(num_objects, z, rng, out + i * R + 1);
    }
}

// include/cuNVSM/cuda_utils.h:35-56 — Glorot uniform, linear memory order,
// rows = feature dim, cols = #objects.
template <typename F>
void init_matrix_glorot(F* data, size_t rows, size_t cols, RNG* rng) {
    const F max = std::sqrt(6.0 / (rows + cols));
    const size_t n = rows * cols;
    for (size_t i = 0; i < n; ++i)
        data[i] = 2 * max * (std::generate_canonical<F, 1>(*rng) - 0.5);
}

// ---------------------------------------------------------------------------
// Functors (include/cuNVSM/cuda_utils.h:58-237).
// ---------------------------------------------------------------------------

template <typename F>
struct Clip {  // :86-111 — bounds one ulp outside [min, max].
    F min_, max_;
    Clip(F lo, F hi, F eps = 1e-5)
        : min_(std::nextafter(lo, lo - eps)), max_(std::nextafter(hi, hi + eps)) {}
    F operator()(F x) const { return std::min(std::max(x, min_), max_); }
    F deriv(F y) const { return (y > min_ && y < max_) ? 1.0 : 0.0; }  // :119-147
};

template <typename F>
inline F truncated_sigmoid(F x, F epsilon) {  // :185-214
    const F prob = (x >= 0) ? 1.0 / (1.0 + std::exp(-x))
                            : std::exp(x) / (1.0 + std::exp(x));
    // min/max against `1.0 - epsilon_` promote to double in the reference.
    const double lo = epsilon, hi = 1.0 - epsilon;
    return static_cast<F>(std::min(std::max(static_cast<double>(prob), lo), hi));
}

template <typename F>
inline F sigmoid_to_log_sigmoid_deriv(F p, F epsilon) {  // :217-235
    return (p >= (1.0 - epsilon) || p <= epsilon) ? 0.0 : 1.0 - p;
}

// ---------------------------------------------------------------------------
// Representations (cpp/params.cu, cpp/storage.cu).
// ---------------------------------------------------------------------------

// average_repr_kernel, cpp/params.cu:75-95 — divides by the window even when
// weighted. weights may be null.
template <typename F>
void gather_mean(const F* repr, size_t dim, const idx_t* indices, const F* weights,
                 size_t num_out, size_t window, F* out) {
#pragma omp parallel for schedule(static)
    for (long o = 0; o < (long)num_out; ++o) {
        for (size_t k = 0; k < dim; ++k) {
            F agg = 0.0;
            for (size_t w = 0; w < window; ++w) {
                const idx_t id = indices[o * window + w];
                const F wt = weights ? weights[o * window + w] : F(1.0);
                agg += wt * repr[id * dim + k];
            }
            out[o * dim + k] = agg / window;
        }
    }
}

// One sparse gradient descriptor: (grad[dim x num_grads], indices[num_grads*window],
// window, weights-or-null) — include/cuNVSM/storage.h SingleGradientType.
template <typename F>
struct SparseGrad {
    F* grad;
    const idx_t* indices;
    size_t num_grads;
    size_t window;
    const F* weights;
};

// RepresentationsStorage::update + update_repr_kernel, cpp/storage.cu:37-102.
template <typename F>
void repr_storage_update(F* repr, size_t num_objects, size_t dim,
                         const std::vector<SparseGrad<F>>& descs, F lr, F lambda) {
    if (lambda > 0.0) {
        const F s = 1.0 - (lambda * lr);
        const long n = (long)(num_objects * dim);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < n; ++i) repr[i] *= s;
    }
    for (const SparseGrad<F>& d : descs) {
        for (size_t x = 0; x < d.num_grads; ++x)
            for (size_t y = 0; y < d.window; ++y) {
                const F wt = d.weights ? d.weights[x * d.window + y] : F(1.0);
                const idx_t id = d.indices[x * d.window + y];
                for (size_t k = 0; k < dim; ++k)
                    repr[id * dim + k] += lr * wt * d.grad[x * dim + k];
            }
    }
}

// update_dense, include/cuNVSM/storage_inl.h:4-32, with op in {identity, square}.
template <typename F, typename Op>
void update_dense(F* param, size_t n, const F* grad, F lr, F lambda, Op op) {
    const F s = 1.0 - lambda * lr;
#pragma omp parallel for schedule(static)
    for (long i = 0; i < (long)n; ++i) param[i] = param[i] * s + op(grad[i]) * lr;
}
template <typename F> struct Identity { F operator()(F x) const { return x; } };
template <typename F> struct Square { F operator()(F x) const { return x * x; } };

// TransformStorage::update, cpp/storage.cu:198-228 — bias never regularised.
template <typename F, typename Op>
void transform_storage_update(F* T, size_t nT, F* b, size_t nb, const F* gT,
                              const F* gb, F lr, F lambda, Op op) {
    update_dense(T, nT, gT, lr, lambda, op);
    update_dense(b, nb, gb, lr, F(0.0), op);
}

// ---------------------------------------------------------------------------
// Batch normalisation (cpp/cudnn_utils.cu:82-183; cuDNN per-activation,
// training mode, gamma == 1, beta = transform bias; formulas pinned by
// cpp/cudnn_utils_tests.cu:143-176).
// ---------------------------------------------------------------------------
// ---------------------------------------------------------------------------
// RepresentationSimilarity objective (EntityEntity / TermTerm; cpp/objective.cu:485-672) on one table[rows][dim]:
// pairs (ids[2i], ids[2i+1]) with weight w[i].
//   compute_cost (:487-573): s_i = <table[a_i], table[b_i]>, p_i = truncated_sigmoid(s_i, clip ? 1e-7 : 0),
//                            mass_i = w_i log p_i;  cost = -(1/N) sum_i mass_i (intermediate_results.cu:80-124)
//   compute_gradients (:575-672): mult_i = w_i * sigmoid_to_log_sigmoid_deriv(p_i, clip ? 1e-6 : 0) * exp(-log N);
//                            the gradient column of a pair member is mult_i times its PARTNER's row
//                            (flip_adjacent_columns), ascent direction; applied like any window-1 sparse gradient.
// ---------------------------------------------------------------------------
template <typename F>
F similarity_step(const F* table, size_t dim, const long* ids, const F* weights, size_t N, bool clip_sigmoid,
                  F* probs, F* grad /* [2N][dim] or null */) {
    const F eps_fwd = clip_sigmoid ? F(1e-7) : F(0), eps_bwd = clip_sigmoid ? F(1e-6) : F(0);
    const F bsn = std::exp(-std::log(F(N)));
    F total = 0;
    for (size_t i = 0; i < N; ++i) {
        const F* a = table + ids[2 * i] * dim;
        const F* b = table + ids[2 * i + 1] * dim;
        F sc = 0;
        for (size_t k = 0; k < dim; ++k) sc += a[k] * b[k];
        const F p = truncated_sigmoid<F>(sc, eps_fwd);
        if (probs) probs[i] = p;
        total += weights[i] * std::log(p);
        if (grad) {
            const F mult = weights[i] * (sigmoid_to_log_sigmoid_deriv<F>(p, eps_bwd) * bsn);
            for (size_t k = 0; k < dim; ++k) {
                grad[(2 * i) * dim + k] = mult * b[k];
                grad[(2 * i + 1) * dim + k] = mult * a[k];
            }
        }
    }
    return -total / F(N);
}

// ---------------------------------------------------------------------------
// L2 Normalizer (cpp/cuda_utils.cu:3-141; golden vectors cpp/cuda_utils_tests.cu:51-92).
// Instances are the columns of the reference's matrices = rows here: x[N][dim].
//   forward  (:12-45):  norms[i] = sqrt(sum_k x[i][k]^2);  y[i] = x[i] / norms[i]
//   backward (:69-128): dx[i] = (dy[i] * norms[i]^2 - x[i] * (x[i] . dy[i])) / norms[i]^3   (x = the cached INPUT)
// ---------------------------------------------------------------------------
template <typename F>
struct Normalizer {
    std::vector<F> norms, input_cache;
    size_t dim = 0;

    // y may alias x (the reference's own test normalises in place).
    void forward(const F* x, size_t N, size_t dim_, F* y) {
        dim = dim_;
        input_cache.assign(x, x + N * dim);
        norms.assign(N, F(0));
        for (size_t i = 0; i < N; ++i) {
            F s = 0;
            for (size_t k = 0; k < dim; ++k) s += input_cache[i * dim + k] * input_cache[i * dim + k];
            norms[i] = std::sqrt(s);
            for (size_t k = 0; k < dim; ++k) y[i * dim + k] = input_cache[i * dim + k] / norms[i];
        }
    }
    void backward(const F* dy, size_t N, F* dx) const {
        for (size_t i = 0; i < N; ++i) {
            F cross = 0;
            for (size_t k = 0; k < dim; ++k) cross += input_cache[i * dim + k] * dy[i * dim + k];
            const F n2 = norms[i] * norms[i], n3 = std::pow(norms[i], F(3.0));
            for (size_t k = 0; k < dim; ++k) dx[i * dim + k] = (dy[i * dim + k] * n2 + F(-1.0) * (input_cache[i * dim + k] * cross)) / n3;
        }
    }
};

template <typename F>
struct BatchNorm {
    size_t C = 0;
    double eps = 1e-4;
    std::vector<F> mean, invstd, input_cache;

    // x, y: [N][C]; y may alias x.
    void forward(const F* x, const F* bias, size_t N, size_t C_, F* y) {
        C = C_;
        mean.assign(C, 0);
        invstd.assign(C, 0);
        input_cache.assign(x, x + N * C);
        // Row-outer loops; every column is still summed in instance order.
        std::vector<F> s(C, 0), v(C, 0);
        for (size_t i = 0; i < N; ++i)
            for (size_t c = 0; c < C; ++c) s[c] += input_cache[i * C + c];
        for (size_t c = 0; c < C; ++c) mean[c] = s[c] / N;
        for (size_t i = 0; i < N; ++i)
            for (size_t c = 0; c < C; ++c) {
                const F d = input_cache[i * C + c] - mean[c];
                v[c] += d * d;
            }
        for (size_t c = 0; c < C; ++c) invstd[c] = 1.0 / std::sqrt(v[c] / N + eps);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i)
            for (size_t c = 0; c < C; ++c)
                y[i * C + c] = (input_cache[i * C + c] - mean[c]) * invstd[c] + bias[c];
    }

    // dy, dx: [N][C]; dx may alias dy. x = cached (or supplied) forward input.
    void backward(const F* dy, const F* x, size_t N, F* dx, F* dbias) const {
        std::vector<F> dgamma(C, 0), sb(C, 0);
        for (size_t i = 0; i < N; ++i)
            for (size_t c = 0; c < C; ++c) {
                const F xh = (x[i * C + c] - mean[c]) * invstd[c];
                sb[c] += dy[i * C + c];
                dgamma[c] += dy[i * C + c] * xh;
            }
        for (size_t c = 0; c < C; ++c) dbias[c] = sb[c];
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i)
            for (size_t c = 0; c < C; ++c) {
                const F xh = (x[i * C + c] - mean[c]) * invstd[c];
                dx[i * C + c] =
                    invstd[c] * (dy[i * C + c] - dbias[c] / N - xh * dgamma[c] / N);
            }
    }
};

// ---------------------------------------------------------------------------
// Optimisers (cpp/updates.cu, updates_adagrad.cu, updates_adam.cu).
// ---------------------------------------------------------------------------
template <typename F>
struct TransformUpdater {
    int method = SGD;
    size_t nT = 0, nb = 0;
    F beta1 = 0.9, beta2 = 0.999, epsilon = 1e-6;
    uint64_t t = 1;
    std::vector<F> aT, ab;          // Adagrad accumulators / Adam m
    std::vector<F> vT, vb;          // Adam v

    void init(int method_, size_t nT_, size_t nb_, F b1 = 0.9, F b2 = 0.999, F eps = 1e-6) {
        method = method_; nT = nT_; nb = nb_; beta1 = b1; beta2 = b2; epsilon = eps; t = 1;
        aT.assign(nT, 0); ab.assign(nb, 0); vT.assign(nT, 0); vb.assign(nb, 0);
    }

    // gT / gb are modified in place exactly like the reference's gradient tensors.
    void update(F* T, F* b, F* gT, F* gb, F lr, F lambda) {
        if (method == SGD) {  // cpp/updates.cu:24-35
            transform_storage_update(T, nT, b, nb, gT, gb, lr, lambda, Identity<F>());
        } else if (method == ADAGRAD) {  // cpp/updates_adagrad.cu:33-70
            transform_storage_update(aT.data(), nT, ab.data(), nb, gT, gb, F(1.0), F(0.0),
                                     Square<F>());
            for (size_t i = 0; i < nT; ++i) gT[i] = gT[i] / std::sqrt(aT[i] + epsilon);
            for (size_t i = 0; i < nb; ++i) gb[i] = gb[i] / std::sqrt(ab[i] + epsilon);
            transform_storage_update(T, nT, b, nb, gT, gb, lr, lambda, Identity<F>());
        } else {  // cpp/updates_adam.cu:46-105
            for (size_t i = 0; i < nT; ++i) gT[i] += -lambda * T[i];  // updates.h:23-62
            const F lr1 = 1.0 - beta1, lr2 = 1.0 - beta2;
            // lambda = 1.0 for the matrix => decay factor 1 - 1*(1-beta); 0 for the bias
            // => the bias moments never decay (pinned by updates_tests.cu:352-366,409-423).
            transform_storage_update(aT.data(), nT, ab.data(), nb, gT, gb, lr1, F(1.0),
                                     Identity<F>());
            transform_storage_update(vT.data(), nT, vb.data(), nb, gT, gb, lr2, F(1.0),
                                     Square<F>());
            const F bc = std::sqrt(1.0 - std::pow(beta2, t)) / (1.0 - std::pow(beta1, t));
            for (size_t i = 0; i < nT; ++i)
                gT[i] = (aT[i] * bc) / (std::sqrt(vT[i]) + epsilon);
            for (size_t i = 0; i < nb; ++i)
                gb[i] = (ab[i] * bc) / (std::sqrt(vb[i]) + epsilon);
            t += 1;
            transform_storage_update(T, nT, b, nb, gT, gb, lr, F(0.0), Identity<F>());
        }
    }
};

template <typename F>
struct RepresentationsUpdater {
    int method = SGD;
    int adam_mode = SPARSE;
    size_t num_objects = 0, dim = 0;
    F beta1 = 0.9, beta2 = 0.999, epsilon = 1e-6;
    uint64_t t = 1;
    std::vector<F> acc;  // Adagrad: [num_objects] scalar accumulators
    std::vector<F> m;    // Adam: [num_objects][dim]
    std::vector<F> v;    // Adam: [num_objects] (SPARSE, DENSE_UPDATE) or [num_objects][dim]

    void init(int method_, int adam_mode_, size_t num_objects_, size_t dim_, F b1 = 0.9,
              F b2 = 0.999, F eps = 1e-6) {
        method = method_; adam_mode = adam_mode_; num_objects = num_objects_; dim = dim_;
        beta1 = b1; beta2 = b2; epsilon = eps; t = 1;
        acc.clear(); m.clear(); v.clear();
        if (method == ADAGRAD) acc.assign(num_objects, 0);
        if (method == ADAM) {
            m.assign(num_objects * dim, 0);
            v.assign(adam_mode < DENSE_UPDATE_DENSE_VARIANCE ? num_objects : num_objects * dim, 0);
        }
    }

    // mean_k g[k, x]^2 per gradient column (reduce_axis<square> then scale by 1/rows).
    static std::vector<F> mean_square(const SparseGrad<F>& d, size_t dim) {
        std::vector<F> out(d.num_grads);
        const F inv = std::exp(-std::log((double)dim));
        for (size_t x = 0; x < d.num_grads; ++x) {
            F s = 0;
            for (size_t k = 0; k < dim; ++k) s += d.grad[x * dim + k] * d.grad[x * dim + k];
            out[x] = s * inv;
        }
        return out;
    }

    void update(F* repr, std::vector<SparseGrad<F>>& descs, F lr, F lambda) {
        if (method == SGD) {  // cpp/updates.cu:37-48
            repr_storage_update(repr, num_objects, dim, descs, lr, lambda);
            return;
        }
        if (method == ADAGRAD) {  // cpp/updates_adagrad.cu:99-179 (single descriptor only)
            SparseGrad<F>& d = descs.front();
            std::vector<F> avg = mean_square(d, dim);
            std::vector<SparseGrad<F>> one{{avg.data(), d.indices, d.num_grads, d.window, d.weights}};
            repr_storage_update(acc.data(), num_objects, 1, one, F(1.0), F(0.0));
            for (size_t x = 0; x < d.num_grads; ++x) {  // adagrad_update_kernel :83-97
                F a = 0.0;
                for (size_t w = 0; w < d.window; ++w) a += acc[d.indices[x * d.window + w]];
                a /= d.window;
                for (size_t k = 0; k < dim; ++k) d.grad[x * dim + k] /= std::sqrt(a + epsilon);
            }
            repr_storage_update(repr, num_objects, dim, descs, lr, lambda);
            return;
        }
        // Adam, cpp/updates_adam.cu:153-385.
        const bool sgd_reg = adam_mode < DENSE_UPDATE_DENSE_VARIANCE;
        repr_storage_update(m.data(), num_objects, dim, descs, F(1.0 - beta1), F(1.0));
        if (!sgd_reg) {
            const F l = (1.0 - beta1) * lambda;
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)(num_objects * dim); ++i) m[i] += -l * repr[i];
        }
        if (sgd_reg) {
            std::vector<std::vector<F>> keep;
            std::vector<SparseGrad<F>> sq;
            for (const SparseGrad<F>& d : descs) {
                keep.push_back(mean_square(d, dim));
                sq.push_back({keep.back().data(), d.indices, d.num_grads, d.window, d.weights});
            }
            repr_storage_update(v.data(), num_objects, 1, sq, F(1.0 - beta2), F(1.0));
        } else {
            std::vector<F> agg(num_objects * dim, 0);
            repr_storage_update(agg.data(), num_objects, dim, descs, F(1.0), F(0.0));
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)agg.size(); ++i) {
                agg[i] += -lambda * repr[i];
                agg[i] = agg[i] * agg[i];
            }
            update_dense(v.data(), v.size(), agg.data(), F(1.0 - beta2), F(1.0), Identity<F>());
        }
        const F bc = std::sqrt(1.0 - std::pow(beta2, t)) / (1.0 - std::pow(beta1, t));
        t += 1;
        if (adam_mode == DENSE_UPDATE) {
            const F s = 1.0 - lambda * lr;
#pragma omp parallel for schedule(static)
            for (long o = 0; o < (long)num_objects; ++o)
                for (size_t k = 0; k < dim; ++k) {
                    const F g = (m[o * dim + k] / (std::sqrt(v[o]) + epsilon)) * bc;
                    repr[o * dim + k] = repr[o * dim + k] * s + g * lr;
                }
        } else if (adam_mode == DENSE_UPDATE_DENSE_VARIANCE) {
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)(num_objects * dim); ++i) {
                const F g = (m[i] / (std::sqrt(v[i]) + epsilon)) * bc;
                repr[i] = repr[i] * F(1.0) + g * lr;
            }
        } else {  // SPARSE: adam_sparse_update_kernel :132-151 (single descriptor only)
            SparseGrad<F>& d = descs.front();
            for (size_t x = 0; x < d.num_grads; ++x) {
                F agg_v = 0.0;
                for (size_t w = 0; w < d.window; ++w) agg_v += v[d.indices[x * d.window + w]];
                agg_v /= d.window;
                for (size_t k = 0; k < dim; ++k) {
                    F agg_m = 0.0;
                    for (size_t w = 0; w < d.window; ++w)
                        agg_m += m[d.indices[x * d.window + w] * dim + k];
                    agg_m /= d.window;
                    d.grad[x * dim + k] = bc * agg_m / (std::sqrt(agg_v) + epsilon);
                }
            }
            repr_storage_update(repr, num_objects, dim, descs, lr, lambda);
        }
    }
};

// ---------------------------------------------------------------------------
// The TextEntity model + objective (cpp/model.cu, cpp/objective.cu:30-481,
// cpp/params.cu:377-535, cpp/intermediate_results.cu:80-129).
// ---------------------------------------------------------------------------
struct Config {
    long num_words = 0, num_entities = 0;
    int word_repr_size = 0, entity_repr_size = 0;
    int nonlinearity = TANH;
    int batch_normalization = 0;
    int clip_sigmoid = 0;
    int bias_negative_samples = 0;
    int update_method = SGD;
    int adam_mode = SPARSE;
    int num_random_entities = 1;
    double regularization_lambda = 0.0;
    double bn_epsilon = 1e-4;  // cpp/objective.cu:114
};

template <typename F>
struct Model {
    Config cfg;
    size_t V, D, dw, dd;
    std::vector<F> W, E, T, b;
    RepresentationsUpdater<F> word_updater, entity_updater;
    TransformUpdater<F> transform_updater;

    // Forward result (cpp/intermediate_results.h:269-290).
    size_t B = 0, n = 0, R = 0;
    std::vector<idx_t> word_ids, entity_ids;
    std::vector<F> word_weights;
    std::vector<F> P, Y, probs, mass, wbc;
    BatchNorm<F> bn;
    // Gradients.
    std::vector<F> mult, gE, Gp, gT, gb, gP;

    explicit Model(const Config& c)
        : cfg(c), V(c.num_words), D(c.num_entities), dw(c.word_repr_size), dd(c.entity_repr_size) {
        W.assign(V * dw, 0); E.assign(D * dd, 0); T.assign(dw * dd, 0); b.assign(dd, 0);
        word_updater.init(cfg.update_method, cfg.adam_mode, V, dw);
        entity_updater.init(cfg.update_method, cfg.adam_mode, D, dd);
        transform_updater.init(cfg.update_method, dw * dd, dd);
    }

    // cpp/model.cu:37-43 — order W, E, T; bias = 0 (cpp/params.cu:361-372).
    void initialize(RNG* rng) {
        init_matrix_glorot(W.data(), dw, V, rng);
        init_matrix_glorot(E.data(), dd, D, rng);
        init_matrix_glorot(T.data(), dd, dw, rng);
        std::fill(b.begin(), b.end(), F(0));
    }

    // Transform::transform, cpp/params.cu:377-451. in: [N][dw] -> out: [N][dd].
    void transform(const F* in, size_t N, bool use_bn, F* out) {
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)N; ++i) {
            F* o = out + i * dd;
            for (size_t r = 0; r < dd; ++r) o[r] = use_bn ? F(0) : b[r];
            for (size_t c = 0; c < dw; ++c) {
                const F x = in[i * dw + c];
                const F* t = T.data() + c * dd;
                for (size_t r = 0; r < dd; ++r) o[r] += t[r] * x;
            }
        }
        if (use_bn) {
            bn.eps = cfg.bn_epsilon;
            bn.forward(out, b.data(), N, dd, out);
        }
        const size_t tot = N * dd;
        if (cfg.nonlinearity == TANH) {
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)tot; ++i) out[i] = std::tanh(out[i]);
        } else {
            const Clip<F> clip(-1.0, 1.0);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)tot; ++i) out[i] = clip(out[i]);
        }
    }

    // Objective::compute_cost with the sampled ids supplied by the caller
    // (cpp/objective.cu:30-313). ids: [B*R], positive first.
    F compute_cost(const idx_t* features, const F* feature_weights, const idx_t* ids,
                   const F* weights, size_t B_, size_t n_) {
        B = B_; n = n_; R = (size_t)cfg.num_random_entities + 1;
        const size_t z = cfg.num_random_entities;
        word_ids.assign(features, features + B * n);
        word_weights.assign(feature_weights, feature_weights + B * n);
        entity_ids.assign(ids, ids + B * R);

        P.assign(B * dw, 0);
        gather_mean(W.data(), dw, word_ids.data(), word_weights.data(), B, n, P.data());
        Y.assign(B * dd, 0);
        transform(P.data(), B, cfg.batch_normalization != 0, Y.data());

        std::vector<F> iw(weights, weights + B);
        const bool rebalance = !cfg.bias_negative_samples && z > 1;
        if (rebalance) {  // :268-274
            const F s = (static_cast<F>(z) + 1.0) / (2.0 * static_cast<F>(z));
            for (size_t i = 0; i < B; ++i) iw[i] *= s;
        }
        wbc.assign(B * R, 0);
        for (size_t c = 0; c < B * R; ++c) wbc[c] = iw[c / R];  // broadcast_columns
        if (rebalance)
            for (size_t c = 0; c < B * R; c += R) wbc[c] *= static_cast<F>(z);  // :282-290

        probs.assign(B * R, 0);
        mass.assign(B * R, 0);
        const F eps = cfg.clip_sigmoid ? 1e-7 : 0.0;
#pragma omp parallel for schedule(static)
        for (long c = 0; c < (long)(B * R); ++c) {
            const size_t i = c / R;
            const bool neg = (c % R) != 0;
            const F* e = E.data() + entity_ids[c] * dd;
            const F* y = Y.data() + i * dd;
            F s = 0;
            for (size_t k = 0; k < dd; ++k) s += y[k] * (neg ? -e[k] : e[k]);
            probs[c] = truncated_sigmoid<F>(s, eps);
            mass[c] = wbc[c] * std::log(probs[c]);
        }
        return get_cost();
    }

    F get_cost() const {  // cpp/intermediate_results.cu:80-124
        F s = 0;
        for (size_t c = 0; c < B * R; ++c) s += mass[c];
        s /= B;
        return -s;
    }

    F scaled_regularization_lambda() const {  // :126-129
        return static_cast<F>(cfg.regularization_lambda) / B;
    }

    // Objective::compute_gradients (cpp/objective.cu:315-481) + Transform::backward
    // (cpp/params.cu:453-535).
    void compute_gradients() {
        const F bsn = std::exp(-std::log((double)B));
        const F eps = cfg.clip_sigmoid ? 1e-6 : 0.0;
        mult.assign(B * R, 0);
        for (size_t c = 0; c < B * R; ++c)
            mult[c] = wbc[c] * (sigmoid_to_log_sigmoid_deriv<F>(probs[c], eps) * bsn);

        gE.assign(B * R * dd, 0);
        Gp.assign(B * dd, 0);
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)B; ++i) {
            for (size_t r = 0; r < R; ++r) {
                const size_t c = i * R + r;
                const bool neg = r != 0;
                const F* e = E.data() + entity_ids[c] * dd;
                for (size_t k = 0; k < dd; ++k) {
                    const F g = Y[i * dd + k] * mult[c];
                    gE[c * dd + k] = neg ? -g : g;
                    Gp[i * dd + k] += mult[c] * (neg ? -e[k] : e[k]);  // fold_columns
                }
            }
        }
        // Transform::backward.
        if (cfg.nonlinearity == TANH) {
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)(B * dd); ++i) Gp[i] = (1.0 - Y[i] * Y[i]) * Gp[i];
        } else {
            const Clip<F> clip(-1.0, 1.0);
#pragma omp parallel for schedule(static)
            for (long i = 0; i < (long)(B * dd); ++i) Gp[i] = clip.deriv(Y[i]) * Gp[i];
        }
        gb.assign(dd, 0);
        if (!cfg.batch_normalization) {
            for (size_t i = 0; i < B; ++i)
                for (size_t k = 0; k < dd; ++k) gb[k] += Gp[i * dd + k];
        } else {
            bn.backward(Gp.data(), bn.input_cache.data(), B, Gp.data(), gb.data());
        }
        gT.assign(dw * dd, 0);
#pragma omp parallel for schedule(static)
        for (long c = 0; c < (long)dw; ++c)
            for (size_t i = 0; i < B; ++i) {
                const F p = P[i * dw + c];
                for (size_t r = 0; r < dd; ++r) gT[c * dd + r] += Gp[i * dd + r] * p;
            }
        gP.assign(B * dw, 0);
        const F inv_n = std::exp(-std::log((double)n));
#pragma omp parallel for schedule(static)
        for (long i = 0; i < (long)B; ++i)
            for (size_t c = 0; c < dw; ++c) {
                F s = 0;
                for (size_t r = 0; r < dd; ++r) s += T[c * dd + r] * Gp[i * dd + r];
                gP[i * dw + c] = s * inv_n;
            }
    }

    // Model::update, cpp/model.cu:187-220 — entities, words, transform.
    void update(F lr, F scaled_lambda) {
        std::vector<SparseGrad<F>> ge{{gE.data(), entity_ids.data(), B * R, 1, nullptr}};
        entity_updater.update(E.data(), ge, lr, scaled_lambda);
        std::vector<SparseGrad<F>> gw{{gP.data(), word_ids.data(), B, n, word_weights.data()}};
        word_updater.update(W.data(), gw, lr, scaled_lambda);
        transform_updater.update(T.data(), b.data(), gT.data(), gb.data(), lr, scaled_lambda);
    }

    // Model::infer, cpp/model.cu:105-133 — no BN at inference.
    void infer(const idx_t* words, size_t N, size_t window, F* out) {
        std::vector<F> p(N * dw);
        gather_mean(W.data(), dw, words, (const F*)nullptr, N, window, p.data());
        transform(p.data(), N, false, out);
    }
};

}  // namespace oracle
