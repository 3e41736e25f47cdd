// C entry points around the UNMODIFIED reference (`Model<TextEntity::Objective>`, include/cuNVSM/model.h:76-130
// of /root/reference) compiled against oracle/ref_shim's stand-in headers -> oracle/_ref/libcunvsm_ref_{f32,f64}.so.
//
// TEST INFRASTRUCTURE: loaded only by tests/ (parity of the sm_100a path against the reference's own
// kernels and call structure) and by bench.py's `--impl reference` arm. One call = one reference call:
//   ref_forward            -> Model::compute_cost            (cpp/model.cu:136-143, cpp/objective.cu:30-305)
//   ref_get_cost           -> ForwardResult::get_cost        (cpp/intermediate_results.cu:80-124)
//   ref_compute_gradients  -> Model::compute_gradients       (cpp/objective.cu:315-481)
//   ref_update             -> Model::update                  (cpp/model.cu:187-220)
//   ref_step               -> the body of iterate_data       (cpp/main.cu:405-444)
// Internals (sampled ids, gradients, batch arrays) are read through the FRIEND_TEST doors the reference
// leaves open for its own tests; no reference source is modified or copied.
#include <cstring>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <vector>

#include "cuNVSM/model.h"

typedef FLOATING_POINT_TYPE FloatT;
typedef Model<TextEntity::Objective> RefModel;

// Friend of Gradients<FloatT> and TextEntity::ForwardResult (include/cuNVSM/intermediate_results.h:87-90,296-298).
class ParamsTest_Transform_BatchNormalization_Test {
 public:
  typedef TextEntity::ForwardResult<FloatT, int32, int32> ResultT;

  static const device_matrix<FloatT>* gradient(const Gradients<FloatT>& g, const std::string& name) {
      if (name == "grad_entity_repr") return g.grad_entity_repr_.get();
      if (name == "grad_phrase_reprs") return g.grad_phrase_reprs_.get();
      if (name == "grad_transform") return g.grad_transform_matrix_.get();
      if (name == "grad_bias") return g.grad_bias_.get();
      return nullptr;
  }

  static const device_matrix<FloatT>* result(const ResultT& r, const std::string& name) {
      if (name == "phrase_reprs") return r.phrase_reprs_.get();
      if (name == "word_projections") return r.word_projections_.get();
      if (name == "similarity_probs") return r.get_similarity_probs();
      if (name == "instance_weights") return r.broadcasted_instance_weights_.get();
      return nullptr;
  }

  static const device_matrix<int32>* entity_ids(const ResultT& r) { return r.entity_ids_.get(); }
};

namespace TextEntity {
// Friend of TextEntity::Batch (include/cuNVSM/data.h:170-171).
class IndriSourceTest_IndriSource_Test {
 public:
  static void fill(Batch* const batch, const int32* const features, const FloatT* const feature_weights,
                   const int32* const labels, const FloatT* const weights, const size_t num_instances) {
      CHECK_LE(num_instances, batch->maximum_size());
      const size_t n = batch->window_size();
      std::memcpy(batch->features_, features, num_instances * n * sizeof(int32));
      std::memcpy(batch->feature_weights_, feature_weights, num_instances * n * sizeof(FloatT));
      std::memcpy(batch->labels_, labels, num_instances * sizeof(int32));
      std::memcpy(batch->weights_, weights, num_instances * sizeof(FloatT));
      batch->num_instances_ = num_instances;
  }
};
}  // namespace TextEntity

namespace {

struct Handle {
  std::unique_ptr<RefModel> model;
  RNG rng;
  std::unique_ptr<RefModel::ForwardResult> result;
  std::unique_ptr<RefModel::Gradients> gradients;
  lse::TrainConfig train_config;
};

const char* full_name(const std::string& name) {
    if (name == "word_representations") return "word_representations-representations";
    if (name == "entity_representations") return "entity_representations-representations";
    if (name == "transform") return "word_entity_mapping-transform";
    if (name == "bias") return "word_entity_mapping-bias";
    return nullptr;
}

const device_matrix<FloatT>* find_tensor(Handle* const h, const char* const name) {
    const std::string key(name);
    const char* const param = full_name(key);
    if (param != nullptr) {
        const Storage<FloatT>::DataType data = h->model->get_data();
        const auto it = data.find(param);
        return it == data.end() ? nullptr : it->second;
    }
    if (h->gradients != nullptr) {
        const device_matrix<FloatT>* const m =
            ParamsTest_Transform_BatchNormalization_Test::gradient(*h->gradients, key);
        if (m != nullptr) return m;
    }
    if (h->result != nullptr) {
        return ParamsTest_Transform_BatchNormalization_Test::result(*h->result, key);
    }
    return nullptr;
}

}  // namespace

extern "C" {

int ref_float_bytes() { return static_cast<int>(sizeof(FloatT)); }

void* ref_create(const long num_words, const long num_entities,
                 const int word_repr_size, const int entity_repr_size,
                 const int nonlinearity, const int batch_normalization,
                 const int clip_sigmoid, const int bias_negative_samples,
                 const int l2_normalize_phrase_reprs, const int l2_normalize_entity_reprs,
                 const int update_method, const int adam_mode,
                 const int batch_size, const int window_size, const int num_random_entities,
                 const double regularization_lambda, const unsigned long seed) {
    lse::ModelDesc desc;
    desc.set_word_repr_size(word_repr_size);
    desc.set_entity_repr_size(entity_repr_size);
    desc.mutable_transform_desc()->set_nonlinearity(
        static_cast<lse::ModelDesc::TransformDesc::Nonlinearity>(nonlinearity));
    desc.mutable_transform_desc()->set_batch_normalization(batch_normalization != 0);
    desc.set_clip_sigmoid(clip_sigmoid != 0);
    desc.set_bias_negative_samples(bias_negative_samples != 0);
    desc.set_l2_normalize_phrase_reprs(l2_normalize_phrase_reprs != 0);
    desc.set_l2_normalize_entity_reprs(l2_normalize_entity_reprs != 0);

    Handle* const h = new Handle;
    h->train_config.set_batch_size(batch_size);
    h->train_config.set_window_size(window_size);
    h->train_config.set_num_random_entities(num_random_entities);
    h->train_config.set_regularization_lambda(regularization_lambda);
    h->train_config.mutable_update_method()->set_type(
        static_cast<lse::TrainConfig::UpdateMethod>(update_method));
    h->train_config.mutable_update_method()->mutable_adam_conf()->set_mode(
        static_cast<lse::TrainConfig::UpdateMethodConf::AdamConf::AdamMode>(adam_mode));

    h->rng.seed(seed);
    h->model.reset(new RefModel(num_words, num_entities, desc, h->train_config));
    h->model->initialize(&h->rng);  // W, E, T Glorot from the shared engine; b = 0 (cpp/model.cu:37-43).
    CCE(cudaDeviceSynchronize());
    return h;
}

void ref_destroy(void* const handle) {
    Handle* const h = static_cast<Handle*>(handle);
    h->gradients.reset();
    h->result.reset();
    delete h;
    cudaDeviceSynchronize();
}

unsigned long ref_get_rng_state(void* const handle) {
    std::stringstream ss;
    ss << static_cast<Handle*>(handle)->rng;
    unsigned long state = 0;
    ss >> state;
    return state;
}

void ref_set_rng_state(void* const handle, const unsigned long state) {
    std::stringstream ss;
    ss << state;
    ss >> static_cast<Handle*>(handle)->rng;
}

// Tensor access: parameters ("word_representations" [V x d_w], "entity_representations" [D x d_d],
// "transform" [d_w x d_d], "bias" [d_d] as row-major images of the column-major device_matrix),
// gradients ("grad_entity_repr" [B R x d_d], "grad_phrase_reprs" [B x d_w], "grad_transform", "grad_bias") and
// forward tensors ("phrase_reprs", "word_projections", "similarity_probs", "instance_weights").
long ref_tensor_size(void* const handle, const char* const name) {
    const device_matrix<FloatT>* const m = find_tensor(static_cast<Handle*>(handle), name);
    return m == nullptr ? -1 : static_cast<long>(m->size());
}

int ref_get_tensor(void* const handle, const char* const name, FloatT* const host) {
    const device_matrix<FloatT>* const m = find_tensor(static_cast<Handle*>(handle), name);
    if (m == nullptr) return -1;
    CCE(cudaDeviceSynchronize());
    CCE(cudaMemcpy(host, m->getData(), m->size() * sizeof(FloatT), cudaMemcpyDeviceToHost));
    return 0;
}

int ref_set_tensor(void* const handle, const char* const name, const FloatT* const host) {
    const device_matrix<FloatT>* const m = find_tensor(static_cast<Handle*>(handle), name);
    if (m == nullptr || full_name(name) == nullptr) return -1;
    CCE(cudaDeviceSynchronize());
    CCE(cudaMemcpy(m->getData(), host, m->size() * sizeof(FloatT), cudaMemcpyHostToDevice));
    return 0;
}

long ref_get_entity_ids(void* const handle, long* const host, const long capacity) {
    Handle* const h = static_cast<Handle*>(handle);
    if (h->result == nullptr) return -1;
    const device_matrix<int32>* const ids =
        ParamsTest_Transform_BatchNormalization_Test::entity_ids(*h->result);
    if (static_cast<long>(ids->size()) > capacity) return -1;
    CCE(cudaDeviceSynchronize());
    CCE(cudaMemcpy(host, ids->getData(), ids->size() * sizeof(int32), cudaMemcpyDeviceToHost));
    return static_cast<long>(ids->size());
}

// TextEntity::Batch (pinned host memory, include/cuNVSM/data.h:114-177).
void* ref_batch_create(const long batch_size, const long window_size) {
    return new TextEntity::Batch(batch_size, window_size);
}

void ref_batch_destroy(void* const batch) { delete static_cast<TextEntity::Batch*>(batch); }

void ref_batch_fill(void* const batch, const long* const features, const FloatT* const feature_weights,
                    const long* const labels, const FloatT* const weights, const long num_instances) {
    TextEntity::IndriSourceTest_IndriSource_Test::fill(
        static_cast<TextEntity::Batch*>(batch), features, feature_weights, labels, weights, num_instances);
}

void ref_forward(void* const handle, void* const batch) {
    Handle* const h = static_cast<Handle*>(handle);
    h->gradients.reset();  // gradients hold a raw pointer to the result they came from.
    h->result.reset(h->model->compute_cost(*static_cast<TextEntity::Batch*>(batch), &h->rng));
}

double ref_get_cost(void* const handle) {
    return static_cast<Handle*>(handle)->result->get_cost();
}

double ref_scaled_regularization_lambda(void* const handle) {
    return static_cast<Handle*>(handle)->result->scaled_regularization_lambda();
}

void ref_compute_gradients(void* const handle) {
    Handle* const h = static_cast<Handle*>(handle);
    h->gradients.reset(h->model->compute_gradients(*h->result));
}

void ref_update(void* const handle, const double learning_rate, const double scaled_regularization_lambda) {
    Handle* const h = static_cast<Handle*>(handle);
    h->model->update(*h->gradients, learning_rate, scaled_regularization_lambda);
}

// One iteration of the reference's training loop, in its order (cpp/main.cu:405-444): compute_cost,
// compute_gradients, update, then the blocking get_cost().
double ref_step(void* const handle, void* const batch, const double learning_rate) {
    Handle* const h = static_cast<Handle*>(handle);
    h->gradients.reset();
    h->result.reset(h->model->compute_cost(*static_cast<TextEntity::Batch*>(batch), &h->rng));
    h->gradients.reset(h->model->compute_gradients(*h->result));
    h->model->update(*h->gradients, learning_rate, h->result->scaled_regularization_lambda());
    return h->result->get_cost();
}

void ref_synchronize() { CCE(cudaDeviceSynchronize()); }

}  // extern "C"

// -------------------------------------------------------------------------------------------------
// The other Model<...> instantiations of the reference (cpp/model.cu:222-228): EntityEntity, TermTerm and the
// two mixtures. The handle owns its batch (a RepresentationSimilarity::Batch, or the std::tuple of batches the
// mixture objectives take).
// -------------------------------------------------------------------------------------------------
namespace RepresentationSimilarity {
// Friend of RepresentationSimilarity::Batch (include/cuNVSM/data.h:611).
class RepresentationSimilarity_DataSource_Test {
 public:
  static void fill(Batch* const batch, const int32* const pair_ids, const FloatT* const weights, const size_t num_pairs) {
      CHECK_LE(num_pairs, batch->maximum_size());
      std::memcpy(batch->features_, pair_ids, 2 * num_pairs * sizeof(int32));
      std::memcpy(batch->weights_, weights, num_pairs * sizeof(FloatT));
      batch->num_instances_ = num_pairs;
  }
};
}  // namespace RepresentationSimilarity

namespace {

typedef RepresentationSimilarity::RepresentationSimilarity_DataSource_Test PairFill;
typedef TextEntity::IndriSourceTest_IndriSource_Test TextFill;

struct AnyHandle {
  RNG rng;
  lse::TrainConfig train_config;
  virtual ~AnyHandle() {}
  virtual Storage<FloatT>::DataType data() const = 0;
  virtual void fill_text(const int32*, const FloatT*, const int32*, const FloatT*, size_t) = 0;
  virtual void fill_pairs(const int32*, const FloatT*, size_t) = 0;
  virtual void forward() = 0;
  virtual double cost() = 0;
  virtual double scaled_lambda() = 0;
  virtual void gradients() = 0;
  virtual void update(double lr, double lambda) = 0;
  virtual const device_matrix<FloatT>* gradient(const std::string& name) = 0;
};

template <typename ObjectiveT>
struct BatchOps;

template <>
struct BatchOps<RepresentationSimilarity::Batch> {
  typedef RepresentationSimilarity::Batch BatchT;
  static BatchT* make(size_t, size_t, size_t N) { return new BatchT(N); }
  static void fill_text(BatchT*, const int32*, const FloatT*, const int32*, const FloatT*, size_t) {
      LOG(FATAL) << "this objective has no TextEntity batch";
  }
  static void fill_pairs(BatchT* b, const int32* ids, const FloatT* w, size_t N) { PairFill::fill(b, ids, w, N); }
};

template <>
struct BatchOps<std::tuple<TextEntity::Batch, RepresentationSimilarity::Batch> > {
  typedef std::tuple<TextEntity::Batch, RepresentationSimilarity::Batch> BatchT;
  static BatchT* make(size_t B, size_t n, size_t N) {
      return new BatchT(std::make_tuple(B, n), std::make_tuple(N));
  }
  static void fill_text(BatchT* b, const int32* f, const FloatT* fw, const int32* l, const FloatT* w, size_t B) {
      TextFill::fill(&std::get<0>(*b), f, fw, l, w, B);
  }
  static void fill_pairs(BatchT* b, const int32* ids, const FloatT* w, size_t N) {
      PairFill::fill(&std::get<1>(*b), ids, w, N);
  }
};

template <typename ObjectiveT>
struct TypedHandle : public AnyHandle {
  typedef Model<ObjectiveT> ModelT;
  typedef BatchOps<typename ModelT::Batch> Ops;
  std::unique_ptr<ModelT> model;
  std::unique_ptr<typename ModelT::Batch> batch;
  std::unique_ptr<typename ModelT::ForwardResult> result;
  std::unique_ptr<typename ModelT::Gradients> grads;

  virtual ~TypedHandle() { grads.reset(); result.reset(); }
  virtual Storage<FloatT>::DataType data() const { return model->get_data(); }
  virtual void fill_text(const int32* f, const FloatT* fw, const int32* l, const FloatT* w, size_t B) {
      Ops::fill_text(batch.get(), f, fw, l, w, B);
  }
  virtual void fill_pairs(const int32* ids, const FloatT* w, size_t N) { Ops::fill_pairs(batch.get(), ids, w, N); }
  virtual void forward() {
      grads.reset();
      result.reset(model->compute_cost(*batch, &rng));
  }
  virtual double cost() { return result->get_cost(); }
  virtual double scaled_lambda() { return result->scaled_regularization_lambda(); }
  virtual void gradients() { grads.reset(model->compute_gradients(*result)); }
  virtual void update(double lr, double lambda) { model->update(*grads, lr, lambda); }
  virtual const device_matrix<FloatT>* gradient(const std::string& name) {
      return grads == nullptr ? nullptr : ParamsTest_Transform_BatchNormalization_Test::gradient(*grads, name);
  }
};

template <typename ObjectiveT>
AnyHandle* make_typed(const long V, const long D, const lse::ModelDesc& desc, const lse::TrainConfig& tc,
                      const unsigned long seed, const size_t B, const size_t n, const size_t N) {
    TypedHandle<ObjectiveT>* const h = new TypedHandle<ObjectiveT>;
    h->train_config = tc;
    h->rng.seed(seed);
    h->model.reset(new Model<ObjectiveT>(V, D, desc, tc));
    h->model->initialize(&h->rng);
    h->batch.reset(TypedHandle<ObjectiveT>::Ops::make(B, n, N));
    CCE(cudaDeviceSynchronize());
    return h;
}

}  // namespace

extern "C" {

// objective: 1 EntityEntity, 2 TermTerm, 3 TextEntityEntityEntity, 4 TextEntityTermTerm (= NVSM_OBJECTIVE_*).
void* ref2_create(const int objective, const long num_words, const long num_entities,
                  const int word_repr_size, const int entity_repr_size,
                  const int nonlinearity, const int batch_normalization,
                  const int clip_sigmoid, const int bias_negative_samples,
                  const int update_method, const int adam_mode,
                  const int batch_size, const int window_size, const int num_random_entities,
                  const int similarity_batch_size,
                  const double regularization_lambda,
                  const double text_entity_weight, const double similarity_weight,
                  const unsigned long seed) {
    lse::ModelDesc desc;
    desc.set_word_repr_size(word_repr_size);
    desc.set_entity_repr_size(entity_repr_size);
    desc.mutable_transform_desc()->set_nonlinearity(
        static_cast<lse::ModelDesc::TransformDesc::Nonlinearity>(nonlinearity));
    desc.mutable_transform_desc()->set_batch_normalization(batch_normalization != 0);
    desc.set_clip_sigmoid(clip_sigmoid != 0);
    desc.set_bias_negative_samples(bias_negative_samples != 0);
    lse::TrainConfig tc;
    tc.set_batch_size(batch_size);
    tc.set_window_size(window_size);
    tc.set_num_random_entities(num_random_entities);
    tc.set_regularization_lambda(regularization_lambda);
    tc.mutable_update_method()->set_type(static_cast<lse::TrainConfig::UpdateMethod>(update_method));
    tc.mutable_update_method()->mutable_adam_conf()->set_mode(
        static_cast<lse::TrainConfig::UpdateMethodConf::AdamConf::AdamMode>(adam_mode));
    tc.set_text_entity_weight(text_entity_weight);
    tc.set_entity_entity_weight(objective == 1 || objective == 3 ? similarity_weight : 0.0);
    tc.set_term_term_weight(objective == 2 || objective == 4 ? similarity_weight : 0.0);
    switch (objective) {
    case 1: return make_typed<EntityEntity::Objective>(num_words, num_entities, desc, tc, seed, batch_size, window_size, similarity_batch_size);
    case 2: return make_typed<TermTerm::Objective>(num_words, num_entities, desc, tc, seed, batch_size, window_size, similarity_batch_size);
    case 3: return make_typed<TextEntityEntityEntity::Objective>(num_words, num_entities, desc, tc, seed, batch_size, window_size, similarity_batch_size);
    case 4: return make_typed<TextEntityTermTerm::Objective>(num_words, num_entities, desc, tc, seed, batch_size, window_size, similarity_batch_size);
    default: return nullptr;
    }
}

void ref2_destroy(void* const handle) {
    delete static_cast<AnyHandle*>(handle);
    cudaDeviceSynchronize();
}

unsigned long ref2_get_rng_state(void* const handle) {
    std::stringstream ss;
    ss << static_cast<AnyHandle*>(handle)->rng;
    unsigned long state = 0;
    ss >> state;
    return state;
}

static const device_matrix<FloatT>* ref2_find(AnyHandle* const h, const char* const name) {
    const char* const param = full_name(name);
    if (param != nullptr) {
        const Storage<FloatT>::DataType data = h->data();
        const auto it = data.find(param);
        return it == data.end() ? nullptr : it->second;
    }
    return h->gradient(name);
}

long ref2_tensor_size(void* const handle, const char* const name) {
    const device_matrix<FloatT>* const m = ref2_find(static_cast<AnyHandle*>(handle), name);
    return m == nullptr ? -1 : static_cast<long>(m->size());
}

int ref2_get_tensor(void* const handle, const char* const name, FloatT* const host) {
    const device_matrix<FloatT>* const m = ref2_find(static_cast<AnyHandle*>(handle), name);
    if (m == nullptr) return -1;
    CCE(cudaDeviceSynchronize());
    CCE(cudaMemcpy(host, m->getData(), m->size() * sizeof(FloatT), cudaMemcpyDeviceToHost));
    return 0;
}

void ref2_fill_text(void* const handle, const long* const features, const FloatT* const feature_weights,
                    const long* const labels, const FloatT* const weights, const long num_instances) {
    static_cast<AnyHandle*>(handle)->fill_text(features, feature_weights, labels, weights, num_instances);
}

void ref2_fill_pairs(void* const handle, const long* const pair_ids, const FloatT* const weights, const long num_pairs) {
    static_cast<AnyHandle*>(handle)->fill_pairs(pair_ids, weights, num_pairs);
}

void ref2_forward(void* const handle) { static_cast<AnyHandle*>(handle)->forward(); }
double ref2_get_cost(void* const handle) { return static_cast<AnyHandle*>(handle)->cost(); }
double ref2_scaled_regularization_lambda(void* const handle) { return static_cast<AnyHandle*>(handle)->scaled_lambda(); }
void ref2_compute_gradients(void* const handle) { static_cast<AnyHandle*>(handle)->gradients(); }
void ref2_update(void* const handle, const double learning_rate, const double scaled_regularization_lambda) {
    static_cast<AnyHandle*>(handle)->update(learning_rate, scaled_regularization_lambda);
}

}  // extern "C"
