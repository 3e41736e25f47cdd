// cuNVSM/params.h — Parameters / Representations / Transform of the reference (include/cuNVSM/params.h:31-203,
// cpp/params.cu) as header-only classes over libnvsm_b200's stand-alone operators: a Representations IS a
// RepresentationsStorage plus its gradient updater, a Transform IS a TransformStorage plus its updater, exactly like the
// reference's inheritance. The fused training step (Model<...>, include/cuNVSM/model.h) does not go through these
// classes; they are the drop-in for callers of the class surface itself (the reference's own params / updates tests).
//
// Not carried over: the `Gradients` / `ForwardResult` container overloads of update() and Transform::backward
// (include/cuNVSM/intermediate_results.h) -- those containers only exist inside the fused step here; update() takes the
// gradient descriptor the containers would have produced (ConstructGradient, cpp/params.cu:10-28).
#ifndef CUNVSM_B200_PARAMS_H
#define CUNVSM_B200_PARAMS_H

#include <cmath>
#include <memory>
#include <random>
#include <vector>

#include "cudnn_utils.h"
#include "nvsm.pb.h"
#include "storage.h"
#include "updates.h"

typedef lse::TrainConfig::UpdateMethod UpdateMethod;
typedef lse::TrainConfig::UpdateMethodConf UpdateMethodConf;

static const char* const ParamName[] = {"word_representations", "word_entity_mapping", "entity_representations"};

// reference: init_matrix_glorot, include/cuNVSM/cuda_utils.h:35-56 — uniform in +-sqrt(6 / (rows + cols)), drawn on the
// host from the shared engine in linear memory order.
template <typename FloatT>
void init_matrix_glorot(cudaStream_like, device_matrix<FloatT>* const matrix, RNG* const rng) {
  const FloatT max = std::sqrt(6.0 / (matrix->getRows() + matrix->getCols()));
  std::vector<FloatT> h(matrix->size());
  for (size_t i = 0; i < h.size(); ++i) h[i] = 2 * max * (std::generate_canonical<FloatT, 1>(*rng) - 0.5);
  matrix->fillwith(nullptr, h);
}

template <typename FloatT>
class Parameters {
 public:
  explicit Parameters(const ParamIdentifier id) : id_(id), initialized_(false) {}
  virtual ~Parameters() {}
  virtual void initialize(RNG* const) { initialized_ = true; }
  bool initialized() const { return initialized_; }
  const char* name() const { return ParamName[id_]; }

 protected:
  const ParamIdentifier id_;

 private:
  bool initialized_;
};

template <typename FloatT, typename IdxType>
class Representations : public Parameters<FloatT>, public RepresentationsStorage<FloatT, IdxType> {
 public:
  using RepresentationsStorage<FloatT, IdxType>::reprs_;
  typedef typename RepresentationsStorage<FloatT, IdxType>::GradientType GradientType;

  // reference: cpp/params.cu:36-65 — the update method picks the updater
  Representations(const ParamIdentifier id, const size_t num_objects, const size_t size, const UpdateMethodConf& update_method,
                  Streams* const streams)
      : Parameters<FloatT>(id), RepresentationsStorage<FloatT, IdxType>(num_objects, size, streams), streams_(streams) {
    if (update_method.type() == lse::TrainConfig::SGD)
      updater_.reset(new SGDRepresentationsGradientUpdater<FloatT, IdxType>(streams));
    else if (update_method.type() == lse::TrainConfig::ADAGRAD)
      updater_.reset(new AdagradRepresentationsGradientUpdater<FloatT, IdxType>(num_objects, streams, DEFAULT_EPSILON, size));
    else if (update_method.type() == lse::TrainConfig::ADAM)
      updater_.reset(new AdamRepresentationsGradientUpdater<FloatT, IdxType>(num_objects, size, update_method.adam_conf(), streams));
    NVSM_CHECK(updater_ != nullptr, "unknown update method");
  }

  void initialize(RNG* const rng) override {   // cpp/params.cu:67-73
    init_matrix_glorot(nullptr, &reprs_, rng);
    Parameters<FloatT>::initialize(rng);
  }

  size_t num_objects() const { return reprs_.getCols(); }
  size_t size() const { return reprs_.getRows(); }

  // cpp/params.cu:97-117: one column per requested index
  device_matrix<FloatT>* get_representations(cudaStream_like, const device_matrix<IdxType>& indices) const {
    return gather(indices, 1, nullptr);
  }
  // cpp/params.cu:119-136 (debugging aid)
  device_matrix<FloatT>* get_representation(const IdxType idx) const {
    device_matrix<IdxType> one(1, 1, nullptr, streams_);
    one.fillwith(nullptr, std::vector<IdxType>(1, idx));
    return gather(one, 1, nullptr);
  }
  // cpp/params.cu:138-172: the mean divides by the window even when weighted (average_repr_kernel :75-95)
  device_matrix<FloatT>* get_average_representations(cudaStream_like, const device_matrix<IdxType>& indices, const size_t window_size,
                                                     const device_matrix<FloatT>* const indices_weights = nullptr) const {
    NVSM_CHECK(window_size > 0 && indices.size() % window_size == 0, "indices are not a multiple of the window");
    NVSM_CHECK(indices_weights == nullptr || indices_weights->size() == indices.size(), "weights / index dimensions disagree");
    return gather(indices, window_size, indices_weights);
  }

  // cpp/params.cu:295-312 with the descriptor ConstructGradient would hand over; nullptr = "No gradient": nothing happens
  void update(GradientType* const gradient_desc, const FloatT learning_rate, const FloatT scaled_regularization_lambda,
              Streams* const streams) {
    if (gradient_desc == nullptr) return;
    updater_->update(this, gradient_desc, learning_rate, scaled_regularization_lambda, streams);
  }

  RepresentationsGradientUpdater<FloatT, IdxType>* updater() { return updater_.get(); }

 private:
  device_matrix<FloatT>* gather(const device_matrix<IdxType>& indices, const size_t window, const device_matrix<FloatT>* const weights) const {
    const size_t num_out = indices.size() / window;
    NVSM_CHECK(num_out > 0, "no representation requested");
    device_matrix<FloatT>* const out = new device_matrix<FloatT>(size(), num_out, nullptr, streams_);
    NVSM_ABORT_ON(nvsm_op_average_representations(streams_->ops(), reprs_.getData(), static_cast<long>(num_objects()),
                                                  static_cast<int>(size()), indices.getData(), weights ? weights->getData() : nullptr,
                                                  static_cast<long>(num_out), static_cast<int>(window), out->getData()));
    return out;
  }
  Streams* const streams_;
  std::unique_ptr<RepresentationsGradientUpdater<FloatT, IdxType>> updater_;
};

template <typename FloatT>
class Transform : public Parameters<FloatT>, public TransformStorage<FloatT> {
 public:
  using TransformStorage<FloatT>::transform_;
  using TransformStorage<FloatT>::bias_;
  typedef typename TransformStorage<FloatT>::GradientType GradientType;

  // reference: cpp/params.cu:330-358
  Transform(const ParamIdentifier id, const lse::ModelDesc::TransformDesc& desc, const size_t word_repr_size,
            const size_t entity_repr_size, const UpdateMethodConf& update_method, Streams* const streams)
      : Parameters<FloatT>(id), TransformStorage<FloatT>(word_repr_size, entity_repr_size, streams), desc_(desc), streams_(streams) {
    if (update_method.type() == lse::TrainConfig::SGD)
      updater_.reset(new SGDTransformGradientUpdater<FloatT>(streams));
    else if (update_method.type() == lse::TrainConfig::ADAGRAD)
      updater_.reset(new AdagradTransformGradientUpdater<FloatT>(source_repr_size(), target_repr_size(), streams));
    else if (update_method.type() == lse::TrainConfig::ADAM)
      updater_.reset(new AdamTransformGradientUpdater<FloatT>(source_repr_size(), target_repr_size(), streams));
    NVSM_CHECK(updater_ != nullptr, "unknown update method");
  }

  size_t source_repr_size() const { return transform_.getCols(); }
  size_t target_repr_size() const { return transform_.getRows(); }

  void initialize(RNG* const rng) override {   // cpp/params.cu:360-372: Glorot projection, zero bias
    init_matrix_glorot(nullptr, &transform_, rng);
    bias_.fillwith(nullptr, FloatT(0.0));
    Parameters<FloatT>::initialize(rng);
  }

  // reference: cpp/params.cu:377-451 — f(T p + b), or f(BN(T p; beta = b)) when a BatchNormalization is handed in
  device_matrix<FloatT>* transform(cudaStream_like, const device_matrix<FloatT>& word_repr,
                                   BatchNormalization<FloatT>* const batch_normalization) const {
    NVSM_CHECK(word_repr.getRows() == source_repr_size() && word_repr.getCols() >= 1, "word representations have the wrong shape");
    const size_t num_instances = word_repr.getCols();
    std::unique_ptr<device_matrix<FloatT>> out(new device_matrix<FloatT>(target_repr_size(), num_instances, nullptr, streams_));
    NVSM_ABORT_ON(nvsm_op_project(streams_->ops(), transform_.getData(), static_cast<int>(source_repr_size()),
                                  static_cast<int>(target_repr_size()), word_repr.getData(), static_cast<long>(num_instances),
                                  batch_normalization ? nullptr : bias_.getData(), out->getData()));
    if (batch_normalization != nullptr) batch_normalization->forward(*out, bias_, out.get());
    const int nl = desc_.nonlinearity() == lse::ModelDesc::TransformDesc::TANH ? NVSM_TANH
                   : desc_.nonlinearity() == lse::ModelDesc::TransformDesc::HARD_TANH ? NVSM_HARD_TANH : -1;
    NVSM_CHECK(nl >= 0, "nonlinearity not implemented.");
    NVSM_ABORT_ON(nvsm_op_activation(streams_->ops(), out->getData(), static_cast<long>(out->size()), nl, out->getData()));
    return out.release();
  }

  // cpp/params.cu:537-554 with the descriptor ConstructGradient would hand over
  void update(GradientType* const gradient_desc, const FloatT learning_rate, const FloatT scaled_regularization_lambda,
              Streams* const streams) {
    if (gradient_desc == nullptr) return;
    updater_->update(this, gradient_desc, learning_rate, scaled_regularization_lambda, streams);
  }

  TransformGradientUpdater<FloatT>* updater() { return updater_.get(); }

 protected:
  const lse::ModelDesc::TransformDesc desc_;

 private:
  Streams* const streams_;
  std::unique_ptr<TransformGradientUpdater<FloatT>> updater_;
};

#endif  // CUNVSM_B200_PARAMS_H
