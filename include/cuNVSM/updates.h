// cuNVSM/updates.h — the nine GradientUpdater classes of the reference (include/cuNVSM/updates.h:66-258,
// cpp/updates.cu, cpp/updates_adagrad.cu, cpp/updates_adam.cu) over libnvsm_b200's nvsm_updater_* entry points. The
// optimiser state (the reference's `storages_`: Adagrad accumulators, Adam moments) lives in the library handle;
// state(name) downloads it for inspection the way the reference's tests read storages_[k]->get().
#ifndef CUNVSM_B200_UPDATES_H
#define CUNVSM_B200_UPDATES_H

#include <string>
#include <vector>

#include "nvsm.pb.h"
#include "storage.h"

#define DEFAULT_EPSILON 1e-6

typedef lse::TrainConfig::UpdateMethodConf::AdamConf AdamConf;

template <typename FloatT>
class GradientUpdater {
 public:
  virtual ~GradientUpdater() { nvsm_updater_destroy(handle_); }
  GradientUpdater(const GradientUpdater&) = delete;
  GradientUpdater& operator=(const GradientUpdater&) = delete;

  // "acc" | "m" | "v" (+ "_bias" for a transform): linear memory order of the matching storage
  std::vector<FloatT> state(const std::string& name) const {
    float* dev = nullptr;
    long count = 0;
    NVSM_ABORT_ON(nvsm_updater_state(handle_, name.c_str(), &dev, &count));
    std::vector<FloatT> out(count);
    NVSM_ABORT_ON(nvsm_dev_download(streams_->ops(), out.data(), dev, count * sizeof(FloatT)));
    return out;
  }

 protected:
  GradientUpdater(Streams* const streams, const int kind, const int method, const int adam_mode, const size_t num_objects,
                  const size_t dim, const size_t target_dim, const FloatT beta1, const FloatT beta2, const FloatT epsilon)
      : epsilon_(epsilon), streams_(streams) {
    NVSM_ABORT_ON(nvsm_updater_create(streams->ops(), kind, method, adam_mode, static_cast<long>(num_objects), static_cast<int>(dim),
                                      static_cast<int>(target_dim), beta1, beta2, epsilon, &handle_));
  }
  const FloatT epsilon_;
  Streams* const streams_;
  nvsm_updater* handle_ = nullptr;
};

template <typename FloatT>
class TransformGradientUpdater : public GradientUpdater<FloatT> {
 public:
  // reference signature: include/cuNVSM/updates.h:96-100
  virtual void update(TransformStorage<FloatT>* const storage, typename TransformStorage<FloatT>::GradientType* const gradient_desc,
                      const FloatT learning_rate, const FloatT scaled_regularization_lambda, Streams* const /*streams*/) {
    typename TransformStorage<FloatT>::ParamType p = storage->get();
    device_matrix<FloatT>& gT = std::get<0>(*gradient_desc);
    device_matrix<FloatT>& gb = std::get<1>(*gradient_desc);
    NVSM_CHECK(gT.size() == std::get<0>(p)->size() && gb.size() == std::get<1>(p)->size(), "transform gradient has the wrong shape");
    NVSM_ABORT_ON(nvsm_updater_update_transform(this->handle_, std::get<0>(p)->getData(), std::get<1>(p)->getData(), gT.getData(),
                                                gb.getData(), learning_rate, scaled_regularization_lambda));
  }

 protected:
  TransformGradientUpdater(Streams* const streams, const int method, const size_t source_dim, const size_t target_dim,
                           const FloatT beta1, const FloatT beta2, const FloatT epsilon)
      : GradientUpdater<FloatT>(streams, 1, method, 0, 0, source_dim, target_dim, beta1, beta2, epsilon) {}
};

template <typename FloatT, typename IdxType>
class RepresentationsGradientUpdater : public GradientUpdater<FloatT> {
 public:
  // reference signature: include/cuNVSM/updates.h:111-115
  virtual void update(RepresentationsStorage<FloatT, IdxType>* const storage,
                      typename RepresentationsStorage<FloatT, IdxType>::GradientType* const gradient_desc,
                      const FloatT learning_rate, const FloatT scaled_regularization_lambda, Streams* const /*streams*/) {
    std::vector<nvsm_grad_desc> d;
    for (const auto& g : *gradient_desc) d.push_back(nvsm_detail::to_desc(g, storage->repr_size()));
    NVSM_ABORT_ON(nvsm_updater_update_representations(this->handle_, storage->get()->getData(), d.data(), static_cast<int>(d.size()),
                                                      learning_rate, scaled_regularization_lambda));
  }

 protected:
  RepresentationsGradientUpdater(Streams* const streams, const int method, const int adam_mode, const size_t num_objects,
                                 const size_t repr_size, const FloatT beta1, const FloatT beta2, const FloatT epsilon)
      : GradientUpdater<FloatT>(streams, 0, method, adam_mode, num_objects, repr_size, 0, beta1, beta2, epsilon) {}
};

// ---- SGD (cpp/updates.cu:24-48): stateless, any shape -------------------------------------------------------------
template <typename FloatT>
class SGDTransformGradientUpdater : public TransformGradientUpdater<FloatT> {
 public:
  explicit SGDTransformGradientUpdater(Streams* const streams = DefaultStream::get())
      : TransformGradientUpdater<FloatT>(streams, NVSM_SGD, 1, 1, 0.9f, 0.999f, DEFAULT_EPSILON) {}
  void update(TransformStorage<FloatT>* const storage, typename TransformStorage<FloatT>::GradientType* const gradient_desc,
              const FloatT learning_rate, const FloatT scaled_regularization_lambda, Streams* const streams) override {
    storage->update(*gradient_desc, learning_rate, scaled_regularization_lambda, streams);   // TransformStorage::update
  }
};

template <typename FloatT, typename IdxType>
class SGDRepresentationsGradientUpdater : public RepresentationsGradientUpdater<FloatT, IdxType> {
 public:
  explicit SGDRepresentationsGradientUpdater(Streams* const streams = DefaultStream::get())
      : RepresentationsGradientUpdater<FloatT, IdxType>(streams, NVSM_SGD, 0, 1, 1, 0.9f, 0.999f, DEFAULT_EPSILON) {}
  void update(RepresentationsStorage<FloatT, IdxType>* const storage,
              typename RepresentationsStorage<FloatT, IdxType>::GradientType* const gradient_desc, const FloatT learning_rate,
              const FloatT scaled_regularization_lambda, Streams* const streams) override {
    storage->update(*gradient_desc, learning_rate, scaled_regularization_lambda, streams);
  }
};

// ---- Adagrad (cpp/updates_adagrad.cu) -------------------------------------------------------------------------------
template <typename FloatT>
class AdagradTransformGradientUpdater : public TransformGradientUpdater<FloatT> {
 public:
  AdagradTransformGradientUpdater(const size_t source_vector_dim, const size_t target_vector_dim, Streams* const streams,
                                  const FloatT epsilon = DEFAULT_EPSILON)
      : TransformGradientUpdater<FloatT>(streams, NVSM_ADAGRAD, source_vector_dim, target_vector_dim, 0.9f, 0.999f, epsilon) {}
};

template <typename FloatT, typename IdxType>
class AdagradRepresentationsGradientUpdater : public RepresentationsGradientUpdater<FloatT, IdxType> {
 public:
  // the accumulator is one scalar per object: the representation size only enters through the gradient's shape
  AdagradRepresentationsGradientUpdater(const size_t num_objects, Streams* const streams, const FloatT epsilon = DEFAULT_EPSILON,
                                        const size_t repr_size = 0)
      : RepresentationsGradientUpdater<FloatT, IdxType>(streams, NVSM_ADAGRAD, 0, num_objects, repr_size ? repr_size : 1, 0.9f, 0.999f,
                                                        epsilon), num_objects_(num_objects), repr_size_(repr_size), eps_(epsilon) {}
  void update(RepresentationsStorage<FloatT, IdxType>* const storage,
              typename RepresentationsStorage<FloatT, IdxType>::GradientType* const gradient_desc, const FloatT learning_rate,
              const FloatT scaled_regularization_lambda, Streams* const streams) override {
    if (repr_size_ != storage->repr_size()) {   // first use (or a differently shaped table): bind the handle to the table's width
      NVSM_CHECK(repr_size_ == 0, "the updater was used with a table of another representation size");
      nvsm_updater_destroy(this->handle_);
      this->handle_ = nullptr;
      repr_size_ = storage->repr_size();
      NVSM_ABORT_ON(nvsm_updater_create(this->streams_->ops(), 0, NVSM_ADAGRAD, 0, static_cast<long>(num_objects_),
                                        static_cast<int>(repr_size_), 0, 0.9f, 0.999f, eps_, &this->handle_));
    }
    RepresentationsGradientUpdater<FloatT, IdxType>::update(storage, gradient_desc, learning_rate, scaled_regularization_lambda, streams);
  }

 private:
  const size_t num_objects_;
  size_t repr_size_;
  const FloatT eps_;
};

// ---- Adam (cpp/updates_adam.cu) ---------------------------------------------------------------------------------------
template <typename FloatT>
class AdamTransformGradientUpdater : public TransformGradientUpdater<FloatT> {
 public:
  AdamTransformGradientUpdater(const size_t source_vector_dim, const size_t target_vector_dim, Streams* const streams,
                               const FloatT beta1 = 0.9, const FloatT beta2 = 0.999, const FloatT epsilon = DEFAULT_EPSILON)
      : TransformGradientUpdater<FloatT>(streams, NVSM_ADAM, source_vector_dim, target_vector_dim, beta1, beta2, epsilon) {}
};

template <typename FloatT, typename IdxType>
class AdamRepresentationsGradientUpdater : public RepresentationsGradientUpdater<FloatT, IdxType> {
 public:
  AdamRepresentationsGradientUpdater(const size_t num_objects, const size_t repr_size, const AdamConf& conf, Streams* const streams,
                                     const FloatT beta1 = 0.9, const FloatT beta2 = 0.999, const FloatT epsilon = DEFAULT_EPSILON)
      : RepresentationsGradientUpdater<FloatT, IdxType>(streams, NVSM_ADAM, static_cast<int>(conf.mode()), num_objects, repr_size, beta1,
                                                        beta2, epsilon) {}
};

#endif  // CUNVSM_B200_UPDATES_H
