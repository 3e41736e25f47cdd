// cuNVSM/storage.h — Storage / RepresentationsStorage / TransformStorage of the reference
// (include/cuNVSM/storage.h:24-206, cpp/storage.cu) as header-only classes over the stand-alone operators of
// libnvsm_b200 (nvsm_op_*): the parameter tensors live in device_matrix objects the storage owns, update() is
// RepresentationsStorage::update / TransformStorage::update (dense decay, then the sparse scatter / the dense step).
// Same names, argument meaning, ownership (get() / get_data() hand out borrowed pointers) and error behaviour (abort).
#ifndef CUNVSM_B200_STORAGE_H
#define CUNVSM_B200_STORAGE_H

#include <map>
#include <string>
#include <tuple>
#include <vector>

#include "device_matrix.h"

template <typename FloatT>
class Storage {
 public:
  typedef FloatT FloatType;
  typedef int32 WordIdxType;
  typedef int32 EntityIdxType;
  typedef std::map<std::string, const device_matrix<FloatT>*> DataType;

  Storage() {}
  virtual ~Storage() {}
  Storage(const Storage&) = delete;
  Storage& operator=(const Storage&) = delete;

  void initialize_with_null() { initialize_with_constant(0.0); }
  virtual DataType get_data() const = 0;
  virtual size_t num_parameters() const = 0;
  virtual void increment_parameter(const size_t idx, const FloatT epsilon) = 0;
  // protected in the reference, reached by its tests through `friend class UpdatesTest`
  virtual void initialize_with_constant(const FloatT value) = 0;

 protected:
  static void bump(device_matrix<FloatT>* const m, const size_t idx, const FloatT epsilon) {
    std::vector<FloatT> h = m->to_host();   // debugging aid (gradient checking), like the reference's increment_scalar
    h[idx] += epsilon;
    m->fillwith(nullptr, h);
  }
};

namespace nvsm_detail {
// one RepresentationsStorage::SingleGradientType -> the C ABI's descriptor (device pointers are borrowed)
template <typename Tuple>
inline nvsm_grad_desc to_desc(const Tuple& g, const size_t repr_size) {
  auto& grad = std::get<0>(g);
  const auto& idx = std::get<1>(g);
  const size_t window = std::get<2>(g);
  const auto* const weights = std::get<3>(g);
  NVSM_CHECK(window > 0 && idx.size() % window == 0, "indices are not a multiple of the window");
  NVSM_CHECK(grad.getRows() == repr_size && grad.getCols() == idx.size() / window, "gradient / index dimensions disagree");
  NVSM_CHECK(weights == nullptr || weights->size() == idx.size(), "weights / index dimensions disagree");
  nvsm_grad_desc d;
  d.grad = grad.getData();
  d.ids = idx.getData();
  d.count = static_cast<long>(idx.size() / window);
  d.window = static_cast<int>(window);
  d.weights = weights ? weights->getData() : nullptr;
  return d;
}
}  // namespace nvsm_detail

template <typename FloatT, typename IdxType>
class RepresentationsStorage : public Storage<FloatT> {
 public:
  typedef std::tuple<device_matrix<FloatT>&,         /* grad_repr [size x n] */
                     const device_matrix<IdxType>&,   /* repr_idx [window x n], linear order */
                     const size_t,                    /* window_size */
                     const device_matrix<FloatT>*     /* idx_weights or nullptr */> SingleGradientType;
  typedef std::vector<SingleGradientType> GradientType;

  RepresentationsStorage(const size_t num_objects, const size_t size, Streams* const streams)
      : reprs_(size, num_objects, streams->next(), streams) {}

  size_t repr_size() const { return reprs_.getRows(); }
  size_t num_objects() const { return reprs_.getCols(); }

  // reference: cpp/storage.cu:51-102 (identity / plus instantiation: what every caller of the class surface uses)
  void update(const GradientType& gradient_descs, const FloatT learning_rate, const FloatT scaled_regularization_lambda,
              Streams* const streams) {
    NVSM_CHECK(learning_rate >= 0.0 && scaled_regularization_lambda >= 0.0, "negative learning rate / lambda");
    std::vector<nvsm_grad_desc> d;
    for (const SingleGradientType& g : gradient_descs) d.push_back(nvsm_detail::to_desc(g, repr_size()));
    NVSM_ABORT_ON(nvsm_op_representations_update(streams->ops(), reprs_.getData(), static_cast<long>(num_objects()),
                                                 static_cast<int>(repr_size()), d.data(), static_cast<int>(d.size()),
                                                 learning_rate, scaled_regularization_lambda));
  }

  // reference: include/cuNVSM/storage.h:96-117 / storage_inl.h:4-32 with a dense gradient of the table's shape;
  // square = the func::square instantiation (second moments)
  void update_dense(cudaStream_like, const device_matrix<FloatT>& grad_reprs, const FloatT learning_rate,
                    const FloatT scaled_regularization_lambda, const bool square = false) {
    NVSM_CHECK(grad_reprs.size() == reprs_.size(), "dense gradient has the wrong shape");
    NVSM_ABORT_ON(nvsm_op_update_dense(reprs_.streams()->ops(), reprs_.getData(), grad_reprs.getData(),
                                       static_cast<long>(reprs_.size()), learning_rate, scaled_regularization_lambda, square));
  }

  device_matrix<FloatT>* get() { return &reprs_; }
  typename Storage<FloatT>::DataType get_data() const override { return {{"representations", &reprs_}}; }
  size_t num_parameters() const override { return reprs_.size(); }
  void increment_parameter(const size_t idx, const FloatT epsilon) override {
    NVSM_CHECK(idx < num_parameters(), "parameter index out of range");
    Storage<FloatT>::bump(&reprs_, idx, epsilon);
  }
  void initialize_with_constant(const FloatT value) override { reprs_.fillwith(nullptr, value); }

  // reference: cpp/storage.cu:139-183 — the (weighted) sum of the gradient entries that land on one parameter
  FloatT get_parameter_gradient(const GradientType& gradient_descs, const size_t param_idx) const {
    NVSM_CHECK(param_idx < num_parameters(), "parameter index out of range");
    const size_t object = param_idx / repr_size(), k = param_idx % repr_size();
    FloatT total = 0.0;
    for (const SingleGradientType& g : gradient_descs) {
      const std::vector<FloatT> grad = std::get<0>(g).to_host();
      const std::vector<IdxType> idx = std::get<1>(g).to_host();
      const size_t window = std::get<2>(g);
      std::vector<FloatT> wts;
      if (std::get<3>(g)) wts = std::get<3>(g)->to_host();
      for (size_t j = 0; j < idx.size(); ++j)
        if (static_cast<size_t>(idx[j]) == object) total += (wts.empty() ? FloatT(1.0) : wts[j]) * grad[(j / window) * repr_size() + k];
    }
    return total;
  }

 protected:
  device_matrix<FloatT> reprs_;
};

template <typename FloatT>
class TransformStorage : public Storage<FloatT> {
 public:
  typedef std::tuple<device_matrix<FloatT>&, /* grad_transform */ device_matrix<FloatT>& /* grad_bias */> GradientType;
  typedef std::tuple<device_matrix<FloatT>*, /* transform */ device_matrix<FloatT>* /* bias */> ParamType;

  // reference: cpp/storage.cu:185-196 — transform_ is entity_repr_size x word_repr_size (column-major)
  TransformStorage(const size_t word_repr_size, const size_t entity_repr_size, Streams* const streams)
      : transform_(entity_repr_size, word_repr_size, streams->next(), streams), bias_(entity_repr_size, 1, streams->next(), streams) {}

  // reference: cpp/storage.cu:198-228 — the bias is never regularised
  void update(const GradientType& gradient_desc, const FloatT learning_rate, const FloatT scaled_regularization_lambda,
              Streams* const streams, const bool square = false) {
    const device_matrix<FloatT>& gT = std::get<0>(gradient_desc);
    const device_matrix<FloatT>& gb = std::get<1>(gradient_desc);
    NVSM_CHECK(gT.size() == transform_.size() && gb.size() == bias_.size(), "transform gradient has the wrong shape");
    NVSM_ABORT_ON(nvsm_op_update_dense(streams->ops(), transform_.getData(), gT.getData(), static_cast<long>(transform_.size()),
                                       learning_rate, scaled_regularization_lambda, square));
    NVSM_ABORT_ON(nvsm_op_update_dense(streams->ops(), bias_.getData(), gb.getData(), static_cast<long>(bias_.size()),
                                       learning_rate, 0.0f, square));
  }

  ParamType get() { return ParamType(&transform_, &bias_); }
  typename Storage<FloatT>::DataType get_data() const override { return {{"transform", &transform_}, {"bias", &bias_}}; }
  size_t num_parameters() const override { return transform_.size() + bias_.size(); }
  void increment_parameter(const size_t idx, const FloatT epsilon) override {
    NVSM_CHECK(idx < num_parameters(), "parameter index out of range");
    if (idx < transform_.size()) Storage<FloatT>::bump(&transform_, idx, epsilon);
    else Storage<FloatT>::bump(&bias_, idx - transform_.size(), epsilon);
  }
  void initialize_with_constant(const FloatT value) override {
    transform_.fillwith(nullptr, value);
    bias_.fillwith(nullptr, value);
  }
  FloatT get_parameter_gradient(const GradientType& gradient_desc, const size_t idx) const {
    NVSM_CHECK(idx < num_parameters(), "parameter index out of range");
    if (idx < transform_.size()) return std::get<0>(gradient_desc).to_host()[idx];
    return std::get<1>(gradient_desc).to_host()[idx - transform_.size()];
  }

 protected:
  device_matrix<FloatT> transform_;
  device_matrix<FloatT> bias_;
};

#endif  // CUNVSM_B200_STORAGE_H
