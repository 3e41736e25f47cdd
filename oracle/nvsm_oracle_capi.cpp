// ORACLE — TEST INFRASTRUCTURE ONLY (see nvsm_oracle.hpp). C entry points for
// ctypes, instantiated for float64 (golden-vector checks; the reference's tests
// all link the float64 build, cpp/CMakeLists.txt:18) and float32 (the release
// arithmetic the CUDA path is compared with, cpp/CMakeLists.txt:17).
#include "nvsm_oracle.hpp"

#include <cstdio>
#include <sstream>
#ifdef _OPENMP
#include <omp.h>
#endif
#include <string>

using namespace oracle;

namespace {

unsigned long rng_state(const RNG& rng) {
    std::ostringstream ss;
    ss << rng;
    return std::stoul(ss.str());
}

template <typename F>
std::vector<F>* model_array(Model<F>* m, const char* name) {
    const std::string s(name);
    if (s == "W") return &m->W;
    if (s == "E") return &m->E;
    if (s == "T") return &m->T;
    if (s == "b") return &m->b;
    if (s == "P") return &m->P;
    if (s == "Y") return &m->Y;
    if (s == "probs") return &m->probs;
    if (s == "mass") return &m->mass;
    if (s == "wbc") return &m->wbc;
    if (s == "mult") return &m->mult;
    if (s == "gE") return &m->gE;
    if (s == "Gp") return &m->Gp;
    if (s == "gT") return &m->gT;
    if (s == "gb") return &m->gb;
    if (s == "gP") return &m->gP;
    if (s == "bn_mean") return &m->bn.mean;
    if (s == "bn_invstd") return &m->bn.invstd;
    if (s == "bn_input") return &m->bn.input_cache;
    if (s == "word_m") return &m->word_updater.m;
    if (s == "word_v") return &m->word_updater.v;
    if (s == "word_acc") return &m->word_updater.acc;
    if (s == "entity_m") return &m->entity_updater.m;
    if (s == "entity_v") return &m->entity_updater.v;
    if (s == "entity_acc") return &m->entity_updater.acc;
    if (s == "T_m") return &m->transform_updater.aT;
    if (s == "T_v") return &m->transform_updater.vT;
    if (s == "b_m") return &m->transform_updater.ab;
    if (s == "b_v") return &m->transform_updater.vb;
    return nullptr;
}

template <typename F>
std::vector<SparseGrad<F>> make_descs(int n, F** grads, const idx_t** idx, const long* num_grads,
                                      const long* window, const F** weights) {
    std::vector<SparseGrad<F>> d;
    for (int i = 0; i < n; ++i)
        d.push_back({grads[i], idx[i], (size_t)num_grads[i], (size_t)window[i],
                     weights ? weights[i] : nullptr});
    return d;
}

}  // namespace

extern "C" {

// --- RNG pieces (type independent). state in/out = minstd_rand0 state. ------
void oracle_generate_labels(const long* labels, long num_labels, long z, long num_objects,
                            unsigned long* state, long* out) {
    RNG rng;
    rng.seed(*state);
    generate_labels(labels, num_labels, z, num_objects, &rng, out);
    *state = rng_state(rng);
}

void oracle_generate_random_indexes(long max, long num, unsigned long* state, long* out) {
    RNG rng;
    rng.seed(*state);
    generate_random_indexes(max, num, &rng, out);
    *state = rng_state(rng);
}

struct oracle_config {  // mirrors oracle::Config field by field
    long num_words, num_entities;
    int word_repr_size, entity_repr_size;
    int nonlinearity, batch_normalization, clip_sigmoid, bias_negative_samples;
    int update_method, adam_mode, num_random_entities;
    double regularization_lambda, bn_epsilon;
};

}  // extern "C"

static Config to_config(const oracle_config* c) {
    Config k;
    k.num_words = c->num_words; k.num_entities = c->num_entities;
    k.word_repr_size = c->word_repr_size; k.entity_repr_size = c->entity_repr_size;
    k.nonlinearity = c->nonlinearity; k.batch_normalization = c->batch_normalization;
    k.clip_sigmoid = c->clip_sigmoid; k.bias_negative_samples = c->bias_negative_samples;
    k.update_method = c->update_method; k.adam_mode = c->adam_mode;
    k.num_random_entities = c->num_random_entities;
    k.regularization_lambda = c->regularization_lambda; k.bn_epsilon = c->bn_epsilon;
    return k;
}

#define ORACLE_API(SUF, F)                                                                      \
    extern "C" {                                                                                \
    void oracle_glorot_##SUF(F* data, long rows, long cols, unsigned long* state) {             \
        RNG rng; rng.seed(*state);                                                              \
        init_matrix_glorot<F>(data, rows, cols, &rng);                                          \
        *state = rng_state(rng);                                                                \
    }                                                                                           \
    F oracle_truncated_sigmoid_##SUF(F x, F eps) { return truncated_sigmoid<F>(x, eps); }       \
    F oracle_sigmoid_deriv_##SUF(F p, F eps) { return sigmoid_to_log_sigmoid_deriv<F>(p, eps); }\
    F oracle_clip_##SUF(F x) { return Clip<F>(-1.0, 1.0)(x); }                                  \
    F oracle_clip_deriv_##SUF(F y) { return Clip<F>(-1.0, 1.0).deriv(y); }                      \
    void oracle_gather_mean_##SUF(const F* repr, long dim, const long* idx, const F* weights,   \
                                  long num_out, long window, F* out) {                          \
        gather_mean<F>(repr, dim, idx, weights, num_out, window, out);                          \
    }                                                                                           \
    void oracle_update_dense_##SUF(F* param, long n, const F* grad, F lr, F lambda,             \
                                   int square) {                                                \
        if (square) update_dense<F>(param, n, grad, lr, lambda, Square<F>());                   \
        else update_dense<F>(param, n, grad, lr, lambda, Identity<F>());                        \
    }                                                                                           \
    /* BN: returns an opaque object holding mean / invstd / cached input. */                    \
    void* oracle_bn_create_##SUF(double eps) {                                                  \
        BatchNorm<F>* bn = new BatchNorm<F>(); bn->eps = eps; return bn;                        \
    }                                                                                           \
    void oracle_bn_destroy_##SUF(void* p) { delete (BatchNorm<F>*)p; }                          \
    void oracle_bn_forward_##SUF(void* p, const F* x, const F* bias, long N, long C, F* y) {    \
        ((BatchNorm<F>*)p)->forward(x, bias, N, C, y);                                          \
    }                                                                                           \
    void oracle_bn_backward_##SUF(void* p, const F* dy, const F* x, long N, F* dx, F* dbias) {  \
        ((BatchNorm<F>*)p)->backward(dy, x, N, dx, dbias);                                      \
    }                                                                                           \
    F oracle_similarity_step_##SUF(const F* table, long dim, const long* ids, const F* weights, \
                                   long N, int clip, F* probs, F* grad) {                       \
        return similarity_step<F>(table, dim, ids, weights, N, clip != 0, probs, grad);         \
    }                                                                                           \
    /* L2 Normalizer: forward into y (may alias x) + backward of dy against the cached input */ \
    void oracle_normalizer_##SUF(const F* x, const F* dy, long N, long dim, F* y, F* dx) {      \
        Normalizer<F> nz;                                                                       \
        nz.forward(x, N, dim, y);                                                               \
        if (dy && dx) nz.backward(dy, N, dx);                                                   \
    }                                                                                           \
    /* Optimisers as free-standing objects (updates_tests.cu style). */                         \
    void* oracle_repr_updater_create_##SUF(int method, int adam_mode, long num_objects,         \
                                           long dim, F b1, F b2, F eps) {                       \
        RepresentationsUpdater<F>* u = new RepresentationsUpdater<F>();                         \
        u->init(method, adam_mode, num_objects, dim, b1, b2, eps);                              \
        return u;                                                                               \
    }                                                                                           \
    void oracle_repr_updater_destroy_##SUF(void* p) { delete (RepresentationsUpdater<F>*)p; }   \
    void oracle_repr_updater_update_##SUF(void* p, F* repr, int n, F** grads, const long** idx, \
                                          const long* num_grads, const long* window,            \
                                          const F** weights, F lr, F lambda) {                  \
        std::vector<SparseGrad<F>> d = make_descs<F>(n, grads, idx, num_grads, window, weights);\
        ((RepresentationsUpdater<F>*)p)->update(repr, d, lr, lambda);                           \
    }                                                                                           \
    /* which: 0 = acc (Adagrad), 1 = m, 2 = v */                                                \
    long oracle_repr_updater_state_##SUF(void* p, int which, F* out, long cap) {                \
        RepresentationsUpdater<F>* u = (RepresentationsUpdater<F>*)p;                           \
        std::vector<F>& s = which == 0 ? u->acc : (which == 1 ? u->m : u->v);                   \
        if (out) for (long i = 0; i < cap && i < (long)s.size(); ++i) out[i] = s[i];            \
        return (long)s.size();                                                                  \
    }                                                                                           \
    void* oracle_transform_updater_create_##SUF(int method, long nT, long nb, F b1, F b2,       \
                                                F eps) {                                        \
        TransformUpdater<F>* u = new TransformUpdater<F>();                                     \
        u->init(method, nT, nb, b1, b2, eps);                                                   \
        return u;                                                                               \
    }                                                                                           \
    void oracle_transform_updater_destroy_##SUF(void* p) { delete (TransformUpdater<F>*)p; }    \
    void oracle_transform_updater_update_##SUF(void* p, F* T, F* b, F* gT, F* gb, F lr,         \
                                               F lambda) {                                      \
        ((TransformUpdater<F>*)p)->update(T, b, gT, gb, lr, lambda);                            \
    }                                                                                           \
    /* which: 0 = aT (acc or m), 1 = ab, 2 = vT, 3 = vb */                                      \
    long oracle_transform_updater_state_##SUF(void* p, int which, F* out, long cap) {           \
        TransformUpdater<F>* u = (TransformUpdater<F>*)p;                                       \
        std::vector<F>& s = which == 0 ? u->aT : which == 1 ? u->ab : which == 2 ? u->vT : u->vb;\
        if (out) for (long i = 0; i < cap && i < (long)s.size(); ++i) out[i] = s[i];            \
        return (long)s.size();                                                                  \
    }                                                                                           \
    /* The model. */                                                                            \
    void* oracle_model_create_##SUF(const oracle_config* c) {                                   \
        return new Model<F>(to_config(c));                                                      \
    }                                                                                           \
    void oracle_model_destroy_##SUF(void* m) { delete (Model<F>*)m; }                           \
    void oracle_model_initialize_##SUF(void* m, unsigned long* state) {                         \
        RNG rng; rng.seed(*state);                                                              \
        ((Model<F>*)m)->initialize(&rng);                                                       \
        *state = rng_state(rng);                                                                \
    }                                                                                           \
    long oracle_model_array_size_##SUF(void* m, const char* name) {                             \
        std::vector<F>* a = model_array<F>((Model<F>*)m, name);                                 \
        return a ? (long)a->size() : -1;                                                        \
    }                                                                                           \
    int oracle_model_get_##SUF(void* m, const char* name, F* out, long n) {                     \
        std::vector<F>* a = model_array<F>((Model<F>*)m, name);                                 \
        if (!a || (long)a->size() != n) return 1;                                               \
        std::copy(a->begin(), a->end(), out);                                                   \
        return 0;                                                                               \
    }                                                                                           \
    int oracle_model_set_##SUF(void* m, const char* name, const F* in, long n) {                \
        std::vector<F>* a = model_array<F>((Model<F>*)m, name);                                 \
        if (!a || (long)a->size() != n) return 1;                                               \
        std::copy(in, in + n, a->begin());                                                      \
        return 0;                                                                               \
    }                                                                                           \
    F oracle_model_compute_cost_##SUF(void* m, const long* features, const F* fw,               \
                                      const long* ids, const F* w, long B, long n) {            \
        return ((Model<F>*)m)->compute_cost(features, fw, ids, w, B, n);                        \
    }                                                                                           \
    void oracle_model_compute_gradients_##SUF(void* m) { ((Model<F>*)m)->compute_gradients(); } \
    void oracle_model_update_##SUF(void* m, F lr, F scaled_lambda) {                            \
        ((Model<F>*)m)->update(lr, scaled_lambda);                                              \
    }                                                                                           \
    F oracle_model_scaled_lambda_##SUF(void* m) {                                               \
        return ((Model<F>*)m)->scaled_regularization_lambda();                                  \
    }                                                                                           \
    void oracle_model_infer_##SUF(void* m, const long* words, long N, long window, F* out) {    \
        ((Model<F>*)m)->infer(words, N, window, out);                                           \
    }                                                                                           \
    }

ORACLE_API(f64, double)
ORACLE_API(f32, float)

extern "C" int oracle_num_threads() {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
