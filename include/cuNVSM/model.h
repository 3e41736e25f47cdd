// cuNVSM/model.h — the Model<TextEntity::Objective> class surface of the reference
// (reference: include/cuNVSM/model.h:20-133, objective.h, intermediate_results.h) as a thin C++
// façade over the C ABI of libnvsm_b200 (include/nvsm_b200.h). Same names, argument meaning and
// ownership rules: factory methods return raw `new` pointers the caller wraps in unique_ptr
// (reference: cpp/main.cu:405-411); the ForwardResult must outlive the Gradients built from it;
// every error aborts the process like the reference's glog CHECK / LOG(FATAL).
//
// Where the reference's arithmetic lives now:
//   Representations::get_average_representations, Transform::transform, BatchNormalization,
//   Objective::compute_cost / compute_gradients, RepresentationsStorage::update,
//   SGD/Adagrad/Adam*GradientUpdater  ->  hand-written sm_100a kernels behind nvsm_compute_cost /
//   nvsm_compute_gradients / nvsm_update (cunvsm_b200/csrc/).
#ifndef CUNVSM_B200_MODEL_H
#define CUNVSM_B200_MODEL_H

#include <map>
#include <memory>
#include <string>
#include <vector>

#include "../nvsm_b200.h"
#include "base.h"
#include "data.h"
#include "labels.h"
#include "nvsm.pb.h"

#define NVSM_ABORT_ON(rc) NVSM_CHECK((rc) == 0, nvsm_last_error())

// Host copy of a parameter tensor, row-major [objects, dim] (what get_array() returns in the
// reference for the column-major dim x objects device_matrix).
template <typename FloatT>
struct HostMatrix {
  size_t rows = 0, cols = 0;   // rows = feature dim, cols = #objects (reference orientation)
  std::vector<FloatT> data;    // data[object * rows + k]
};

class Typedefs {
 public:
  typedef int32 IdxType;
  typedef IdxType WordIdxType;
  typedef IdxType EntityIdxType;
  typedef FLOATING_POINT_TYPE FloatT;
};

template <typename ObjectiveT>
class Model;
class MultiForwardResult;

namespace TextEntity {

// Handle on the forward state held in the model's device workspace
// (reference: SimpleForwardResult / TextEntity::ForwardResult, intermediate_results.h:148-309).
class ForwardResult {
 public:
  typedef Typedefs::FloatT FloatT;
  FloatT get_cost() const {  // cpp/intermediate_results.cu:80-124 (synchronises)
    if (!have_cost_) { NVSM_ABORT_ON(nvsm_read_cost(model_, static_cast<int>(age()), &cost_)); have_cost_ = true; }
    return cost_;
  }
  FloatT scaled_regularization_lambda() const { return lambda_; }  // :126-129
  std::vector<FloatT> get_similarity_probs() const {
    NVSM_CHECK(age() == 0, "forward state has been overwritten by a later compute_cost");
    const long n = nvsm_tensor_size(model_, "similarity_probs");
    std::vector<FloatT> out(n);
    NVSM_ABORT_ON(nvsm_get_tensor(model_, "similarity_probs", out.data(), n));
    return out;
  }
  size_t batch_size() const { return batch_size_; }

 private:
  ForwardResult(nvsm_model* m, size_t batch_size, FloatT lambda, const long* counter)
      : model_(m), batch_size_(batch_size), lambda_(lambda), counter_(counter), serial_(*counter) {}
  long age() const { return *counter_ - serial_; }
  nvsm_model* model_;
  size_t batch_size_;
  FloatT lambda_;
  const long* counter_;
  long serial_;
  mutable FloatT cost_ = 0;
  mutable bool have_cost_ = false;
  template <typename O> friend class ::Model;
  friend class ::MultiForwardResult;
};

// reference: Gradients / SingleGradients (intermediate_results.h:45-110)
class Gradients {
 public:
  typedef Typedefs::FloatT FloatT;
  std::vector<FloatT> get(const std::string& name) const {  // grad_transform, grad_bias, grad_phrase_reprs, grad_entity_repr
    const long n = nvsm_tensor_size(model_, name.c_str());
    NVSM_CHECK(n >= 0, "unknown gradient tensor");
    std::vector<FloatT> out(n);
    NVSM_ABORT_ON(nvsm_get_tensor(model_, name.c_str(), out.data(), n));
    return out;
  }
 private:
  explicit Gradients(nvsm_model* m, const void* r) : model_(m), result_(r) {}
  nvsm_model* model_;
  const void* result_;   // the forward result these gradients belong to (must outlive them, like the reference)
  template <typename O> friend class ::Model;
};

// Tag type selecting the LSE / NVSM objective (reference: TextEntity::Objective, objective.h:66-96).
class Objective {
 public:
  typedef Typedefs::WordIdxType WordIdxType;
  typedef Typedefs::EntityIdxType EntityIdxType;
  typedef Typedefs::FloatT FloatT;
  typedef Batch BatchType;
  typedef ForwardResult ForwardResultType;
  typedef Gradients GradientsType;
  static const int kObjective = NVSM_OBJECTIVE_TEXT_ENTITY;
  template <typename ModelT>
  static ForwardResultType* compute_cost(const ModelT* model, const BatchType& batch, RNG* const rng) {
    return model->text_forward(batch, rng);
  }
};

}  // namespace TextEntity

namespace RepresentationSimilarity {

// reference: RepresentationSimilarity::ForwardResult, intermediate_results.h:311-361
class ForwardResult {
 public:
  typedef Typedefs::FloatT FloatT;
  // The library keeps ONE pair loss (the latest nvsm_similarity_compute_cost), so the cost is read when the result is
  // made (Model::pair_forward synchronises anyway): a result read after a later compute_cost -- the CLI reads the
  // loss of batch k-1 under batch k -- still reports its own batch.
  FloatT get_cost() const { return cost_; }
  FloatT scaled_regularization_lambda() const { return lambda_; }
  std::vector<FloatT> get_similarity_probs() const {
    NVSM_CHECK(age() == 0, "forward state has been overwritten by a later compute_cost");
    const long n = nvsm_tensor_size(model_, "similarity_pair_probs");
    std::vector<FloatT> out(n);
    NVSM_ABORT_ON(nvsm_get_tensor(model_, "similarity_pair_probs", out.data(), n));
    return out;
  }
  long age() const { return *counter_ - serial_; }

 private:
  ForwardResult(nvsm_model* m, FloatT lambda, FloatT cost, const long* counter)
      : model_(m), lambda_(lambda), cost_(cost), counter_(counter), serial_(*counter) {}
  nvsm_model* model_;
  FloatT lambda_;
  FloatT cost_;
  const long* counter_;
  long serial_;
  template <typename O> friend class ::Model;
};

template <int kObjectiveId>
class ObjectiveT {
 public:
  typedef Typedefs::WordIdxType WordIdxType;
  typedef Typedefs::EntityIdxType EntityIdxType;
  typedef Typedefs::FloatT FloatT;
  typedef Batch BatchType;
  typedef ForwardResult ForwardResultType;
  typedef TextEntity::Gradients GradientsType;
  static const int kObjective = kObjectiveId;
  template <typename ModelT>
  static ForwardResultType* compute_cost(const ModelT* model, const BatchType& batch, RNG* const) {
    return model->pair_forward(batch);
  }
};

}  // namespace RepresentationSimilarity

// reference: EntityEntity::Objective / TermTerm::Objective, objective.h:134-162
namespace EntityEntity { typedef RepresentationSimilarity::ObjectiveT<NVSM_OBJECTIVE_ENTITY_ENTITY> Objective; }
namespace TermTerm { typedef RepresentationSimilarity::ObjectiveT<NVSM_OBJECTIVE_TERM_TERM> Objective; }

// reference: MultiForwardResultBase, intermediate_results.h:167-195 / cpp/intermediate_results.cu:186-240 — cost and
// scaled lambda are the plain averages over the constituents.
class MultiForwardResult {
 public:
  typedef Typedefs::FloatT FloatT;
  FloatT get_cost() const { return (text_->get_cost() + pair_->get_cost()) / 2; }
  FloatT scaled_regularization_lambda() const {
    return (text_->scaled_regularization_lambda() + pair_->scaled_regularization_lambda()) / 2;
  }
  long age() const { return text_->age() > pair_->age() ? text_->age() : pair_->age(); }
  const TextEntity::ForwardResult& text() const { return *text_; }
  const RepresentationSimilarity::ForwardResult& similarity() const { return *pair_; }

 private:
  MultiForwardResult(TextEntity::ForwardResult* t, RepresentationSimilarity::ForwardResult* p) : text_(t), pair_(p) {}
  std::unique_ptr<TextEntity::ForwardResult> text_;
  std::unique_ptr<RepresentationSimilarity::ForwardResult> pair_;
  template <typename O> friend class ::Model;
};

// reference: TextEntityEntityEntity::Objective / TextEntityTermTerm::Objective, objective.h:164-238 — both
// constituents on their own batch, gradients merged with weights w_k / sum_k w_k.
template <int kObjectiveId>
class MixtureObjectiveT {
 public:
  typedef Typedefs::WordIdxType WordIdxType;
  typedef Typedefs::EntityIdxType EntityIdxType;
  typedef Typedefs::FloatT FloatT;
  typedef std::tuple<TextEntity::Batch, RepresentationSimilarity::Batch> BatchType;
  typedef MultiForwardResult ForwardResultType;
  typedef TextEntity::Gradients GradientsType;
  static const int kObjective = kObjectiveId;
  template <typename ModelT>
  static ForwardResultType* compute_cost(const ModelT* model, const BatchType& batch, RNG* const rng) {
    return model->mixture_forward(std::get<0>(batch), std::get<1>(batch), rng);
  }
};
namespace TextEntityEntityEntity { typedef MixtureObjectiveT<NVSM_OBJECTIVE_TEXT_ENTITY_ENTITY_ENTITY> Objective; }
namespace TextEntityTermTerm { typedef MixtureObjectiveT<NVSM_OBJECTIVE_TEXT_ENTITY_TERM_TERM> Objective; }

template <typename ObjectiveT>
class Model {
 public:
  typedef ObjectiveT Objective;
  typedef typename Objective::WordIdxType WordIdxType;
  typedef typename Objective::EntityIdxType EntityIdxType;
  typedef typename Objective::FloatT FloatT;
  typedef typename Objective::BatchType Batch;
  typedef typename Objective::ForwardResultType ForwardResult;
  typedef typename Objective::GradientsType Gradients;
  typedef std::map<std::string, HostMatrix<FloatT>> DataType;

  // reference: Model::Model, include/cuNVSM/model.h:82-85
  Model(const size_t num_words, const size_t num_entities, const lse::ModelDesc& desc,
        const lse::TrainConfig& train_config, const int device = 0, const int gemm_mode = NVSM_GEMM_TF32)
      : desc_(desc), train_config_(train_config), num_words_(num_words), num_entities_(num_entities) {
    nvsm_config c = nvsm_config();
    c.num_words = num_words; c.num_entities = num_entities;
    c.word_repr_size = desc.word_repr_size(); c.entity_repr_size = desc.entity_repr_size();
    c.nonlinearity = desc.transform_desc().nonlinearity();
    c.batch_normalization = desc.transform_desc().batch_normalization();
    c.clip_sigmoid = desc.clip_sigmoid();
    c.bias_negative_samples = desc.bias_negative_samples();
    c.l2_normalize_phrase_reprs = desc.l2_normalize_phrase_reprs();
    c.l2_normalize_entity_reprs = desc.l2_normalize_entity_reprs();
    c.update_method = train_config.update_method().type();
    c.adam_mode = train_config.update_method().adam_conf().mode();
    c.num_random_entities = train_config.num_random_entities();
    c.max_batch_size = train_config.batch_size();
    c.window_size = train_config.window_size();
    c.regularization_lambda = train_config.regularization_lambda();
    c.device = device; c.gemm_mode = gemm_mode; c.num_batch_slots = 1;
    c.objective = ObjectiveT::kObjective;
    c.text_entity_weight = train_config.text_entity_weight();
    c.similarity_weight = (c.objective == NVSM_OBJECTIVE_ENTITY_ENTITY || c.objective == NVSM_OBJECTIVE_TEXT_ENTITY_ENTITY_ENTITY)
                              ? train_config.entity_entity_weight() : train_config.term_term_weight();
    c.max_similarity_batch_size = train_config.batch_size();
    NVSM_ABORT_ON(nvsm_create(&c, &handle_));
  }
  virtual ~Model() { nvsm_destroy(handle_); }
  Model(const Model&) = delete;
  Model& operator=(const Model&) = delete;

  // reference: ModelBase::initialize, cpp/model.cu:37-43 — consumes the shared RNG (W, E, T)
  void initialize(RNG* const rng) {
    unsigned long st = nvsm_detail::rng_get_state(*rng);
    NVSM_ABORT_ON(nvsm_initialize(handle_, &st));
    nvsm_detail::rng_set_state(rng, st);
    initialized_ = true;
  }
  bool initialized() const { return initialized_; }
  size_t num_parameters() const {
    return num_words_ * desc_.word_repr_size() + num_entities_ * desc_.entity_repr_size() +
           static_cast<size_t>(desc_.word_repr_size()) * desc_.entity_repr_size() + desc_.entity_repr_size();
  }

  // Optional: keep the std::minstd_rand0 state on the device and draw the negatives there (bit-exact
  // with the host loop). While enabled, compute_cost ignores its RNG* argument; sync_rng() writes the
  // device engine state back into the caller's RNG (synchronises).
  void use_device_sampler(RNG* const rng) {
    NVSM_CHECK(label_generator_->on_device(), "the installed label generator only runs on the host");
    NVSM_ABORT_ON(nvsm_sampler_seed(handle_, nvsm_detail::rng_get_state(*rng)));
    device_sampler_ = true;
  }
  // The reference's Objective owns a LabelGenerator (UniformLabelGenerator, include/cuNVSM/objective.h / labels.h).
  // Takes ownership. Generators with on_device() keep working under use_device_sampler (same ids); any other
  // generator switches the model back to the host path.
  void set_label_generator(LabelGenerator<FloatT, EntityIdxType>* const generator, RNG* const rng = nullptr) {
    NVSM_CHECK(generator != nullptr, "null label generator");
    if (device_sampler_ && !generator->on_device()) {
      NVSM_CHECK(rng != nullptr, "switching to a host-only generator needs the RNG to hand the engine state back");
      sync_rng(rng);
      device_sampler_ = false;
    }
    label_generator_.reset(generator);
    const std::vector<double>* const cdf = generator->on_device() ? generator->distribution() : nullptr;
    NVSM_ABORT_ON(nvsm_sampler_set_cdf(handle_, cdf ? cdf->data() : nullptr, cdf ? static_cast<long>(cdf->size()) : 0));
  }
  bool uses_device_sampler() const { return device_sampler_; }
  void sync_rng(RNG* const rng) {
    if (!device_sampler_) return;
    unsigned long st = 0;
    NVSM_ABORT_ON(nvsm_sampler_state(handle_, &st));
    nvsm_detail::rng_set_state(rng, st);
  }

  // reference: Model::compute_cost (cpp/model.cu:136-143) -> ObjectiveT::compute_cost
  ForwardResult* compute_cost(const Batch& batch, RNG* const rng) const {
    return ObjectiveT::compute_cost(this, batch, rng);
  }

  // TextEntity::Objective::compute_cost, cpp/objective.cu:30-313. Negatives are drawn from `rng` on the
  // host exactly like UniformLabelGenerator (cpp/labels.cu:3-22).
  TextEntity::ForwardResult* text_forward(const TextEntity::Batch& batch, RNG* const rng) const {
    const size_t B = batch.num_instances(), R = train_config_.num_random_entities() + 1;
    NVSM_CHECK(batch.window_size() == static_cast<size_t>(train_config_.window_size()), "window size mismatch");
    if (device_sampler_) {
      NVSM_ABORT_ON(nvsm_step_sampled(handle_, batch.features(), fw_or_null(batch), batch.labels(), w_or_null(batch),
                                      B, 0.0f, /*train=*/0));
      NVSM_ABORT_ON(nvsm_wait_upload(handle_));   // the caller may recycle the batch (AsyncSource) once we return
      ++forward_counter_;
      return new TextEntity::ForwardResult(handle_, B, nvsm_scaled_regularization_lambda(handle_), &forward_counter_);
    }
    // Objective::generate_labels, cpp/objective.cu:5-28
    label_generator_->generate(batch.labels(), num_entities_, B, train_config_.num_random_entities(), &entity_ids_, rng);
    NVSM_CHECK(entity_ids_.size() == B * R, "the label generator returned a wrong number of ids");
    NVSM_ABORT_ON(nvsm_compute_cost(handle_, batch.features(), fw_or_null(batch), entity_ids_.data(), w_or_null(batch), B));
    NVSM_ABORT_ON(nvsm_wait_upload(handle_));     // (entity_ids_ is reused by the next call as well)
    ++forward_counter_;
    return new TextEntity::ForwardResult(handle_, B, nvsm_scaled_regularization_lambda(handle_), &forward_counter_);
  }

  // RepresentationSimilarity::Objective::compute_cost, cpp/objective.cu:487-573
  RepresentationSimilarity::ForwardResult* pair_forward(const RepresentationSimilarity::Batch& batch) const {
    NVSM_ABORT_ON(nvsm_similarity_compute_cost(handle_, batch.features(), batch.weights(), batch.num_instances()));
    NVSM_ABORT_ON(nvsm_synchronize(handle_));     // pair batches are copied on the compute stream: cheap, and recyclable after
    FloatT cost = 0;
    NVSM_ABORT_ON(nvsm_similarity_get_cost(handle_, &cost));   // (already on the host: the forward ends in its D2H copy)
    ++pair_counter_;
    return new RepresentationSimilarity::ForwardResult(handle_, nvsm_similarity_scaled_regularization_lambda(handle_), cost,
                                                       &pair_counter_);
  }

  // TextEntity{EntityEntity,TermTerm}::Objective::compute_cost, cpp/objective.cu:713-724,762-773
  MultiForwardResult* mixture_forward(const TextEntity::Batch& text, const RepresentationSimilarity::Batch& pairs,
                                      RNG* const rng) const {
    TextEntity::ForwardResult* const t = text_forward(text, rng);
    return new MultiForwardResult(t, pair_forward(pairs));
  }

  // reference: Model::compute_gradients, cpp/objective.cu:315-481
  Gradients* compute_gradients(const ForwardResult& result) {
    NVSM_CHECK(result.age() == 0, "compute_gradients needs the most recent ForwardResult");
    NVSM_ABORT_ON(nvsm_compute_gradients(handle_));
    return new Gradients(handle_, &result);
  }

  // reference: Model::update, cpp/model.cu:187-220 (entities, words, transform)
  void update(const Gradients& gradients, const FloatT learning_rate, const FloatT scaled_regularization_lambda) {
    (void)gradients;
    NVSM_ABORT_ON(nvsm_update(handle_, learning_rate, scaled_regularization_lambda));
  }

  // reference: Model::backprop, cpp/model.cu:176-185
  void backprop(const ForwardResult& result, const FloatT learning_rate) {
    std::unique_ptr<Gradients> gradients(compute_gradients(result));
    update(*gradients, learning_rate, result.scaled_regularization_lambda());
  }

  // reference: Model::get_cost, cpp/model.cu:154-174 — optionally restores the RNG first
  FloatT get_cost(const Batch& batch, const std::stringstream* const rng_state, RNG* const rng) const {
    if (rng_state != nullptr) {
      std::stringstream copy;
      copy << rng_state->str();
      copy >> *rng;
    }
    std::unique_ptr<ForwardResult> result(compute_cost(batch, rng));
    return result->get_cost();
  }

  // reference: Model::infer, cpp/model.cu:105-133 (no batch-norm at inference); returns [N, d_d]
  HostMatrix<FloatT> infer(const std::vector<std::vector<WordIdxType>>& words, const size_t window_size) const {
    std::vector<WordIdxType> flat;
    for (const auto& w : words) { NVSM_CHECK(w.size() == window_size, "ragged inference window"); flat.insert(flat.end(), w.begin(), w.end()); }
    HostMatrix<FloatT> out;
    out.rows = desc_.entity_repr_size(); out.cols = words.size();
    out.data.resize(out.rows * out.cols);
    NVSM_ABORT_ON(nvsm_infer(handle_, flat.data(), words.size(), window_size, out.data.data()));
    return out;
  }

  // reference: ModelBase::get_data, cpp/model.cu:64-93 — same four names
  DataType get_data() const {
    DataType data;
    fetch(&data, "word_representations-representations", desc_.word_repr_size(), num_words_);
    fetch(&data, "entity_representations-representations", desc_.entity_repr_size(), num_entities_);
    fetch(&data, "word_entity_mapping-transform", desc_.entity_repr_size(), desc_.word_repr_size());
    fetch(&data, "word_entity_mapping-bias", desc_.entity_repr_size(), 1);
    return data;
  }

  nvsm_model* handle() const { return handle_; }
  const lse::ModelDesc& desc() const { return desc_; }

  // uniform weighting declared by the data source (Batch::set_uniform_weights): NULL over the C ABI, no H2D copy
  static const float* fw_or_null(const TextEntity::Batch& b) { return b.uniform_feature_weights() ? nullptr : b.feature_weights(); }
  static const float* w_or_null(const TextEntity::Batch& b) { return b.uniform_weights() ? nullptr : b.weights(); }

 private:
  void fetch(DataType* data, const char* name, size_t rows, size_t cols) const {
    HostMatrix<FloatT> m;
    m.rows = rows; m.cols = cols; m.data.resize(rows * cols);
    NVSM_ABORT_ON(nvsm_get_tensor(handle_, name, m.data.data(), static_cast<long>(rows * cols)));
    (*data)[name] = std::move(m);
  }
  const lse::ModelDesc desc_;
  const lse::TrainConfig train_config_;
  const size_t num_words_, num_entities_;
  nvsm_model* handle_ = nullptr;
  bool initialized_ = false;
  bool device_sampler_ = false;
  std::unique_ptr<LabelGenerator<FloatT, EntityIdxType>> label_generator_{new UniformLabelGenerator<FloatT, EntityIdxType>()};
  mutable std::vector<long> entity_ids_;
  mutable long forward_counter_ = 0;
  mutable long pair_counter_ = 0;
};

typedef Model<TextEntity::Objective> DefaultModel;
typedef DefaultModel LSE;

#endif  // CUNVSM_B200_MODEL_H
