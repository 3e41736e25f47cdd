// cuNVSM/data.h — host-side batch of the TextEntity objective and the data-source interface.
// TextEntity::Batch has the reference's layout (reference: include/cuNVSM/data.h:114-177,
// cpp/data.cu:8-30,94-124): four page-locked arrays, instance-major,
//   features_[B*n] (long), feature_weights_[B*n], labels_[B] (long), weights_[B].
// The Indri-backed sources of the reference are out of scope; SyntheticSource generates seeded
// uniform / Zipf n-grams with the same DataSource contract (has_next / next / reset / progress).
#ifndef CUNVSM_B200_DATA_H
#define CUNVSM_B200_DATA_H

#include <algorithm>
#include <cmath>
#include <cstring>
#include <tuple>
#include <vector>

#include "../nvsm_b200.h"
#include "base.h"
#include "nvsm.pb.h"

typedef int32 WordIdxType;
typedef int32 ObjectIdxType;
typedef FLOATING_POINT_TYPE WeightType;

class BatchInterface {
 public:
  virtual ~BatchInterface() {}
  virtual void clear() = 0;
  virtual bool full() const = 0;
  virtual bool empty() const = 0;
  virtual size_t num_instances() const = 0;
  virtual size_t maximum_size() const = 0;
};

template <typename BatchT>
class DataSource {
 public:
  virtual ~DataSource() {}
  virtual void reset() = 0;
  virtual void next(BatchT* const batch) = 0;
  virtual bool has_next() const = 0;
  virtual float32 progress() const = 0;
};

namespace TextEntity {

class Objective;

class Batch : public BatchInterface {
 public:
  Batch(const size_t batch_size, const size_t window_size)
      : batch_size_(batch_size), window_size_(window_size), num_instances_(0) {
    NVSM_CHECK(batch_size_ > 0 && window_size_ > 0, "batch and window size must be positive");
    alloc(&features_, batch_size_ * window_size_);
    alloc(&feature_weights_, batch_size_ * window_size_);
    alloc(&labels_, batch_size_);
    alloc(&weights_, batch_size_);
    clear();
  }
  explicit Batch(const lse::TrainConfig& train_config) : Batch(train_config.batch_size(), train_config.window_size()) {}
  // Forward constructor (reference: include/cuNVSM/data.h:120-121), for std::tuple<Batch...> of the mixtures.
  Batch(const std::tuple<size_t, size_t>& args) : Batch(std::get<0>(args), std::get<1>(args)) {}
  virtual ~Batch() {
    nvsm_host_free(features_); nvsm_host_free(feature_weights_); nvsm_host_free(labels_); nvsm_host_free(weights_);
  }
  Batch(const Batch&) = delete;
  Batch& operator=(const Batch&) = delete;

  virtual void clear() override { num_instances_ = 0; }
  virtual bool full() const override { return num_instances_ == batch_size_; }
  virtual bool empty() const override { return num_instances_ == 0; }
  virtual size_t num_instances() const override { return num_instances_; }
  virtual size_t maximum_size() const override { return batch_size_; }
  size_t window_size() const { return window_size_; }

  // DataSource::push_instance (reference: cpp/data.cu:94-124); empty weights => 1.0
  bool push_instance(const std::vector<WordIdxType>& features, const std::vector<WeightType>& feature_weights,
                     const ObjectIdxType object_id, const WeightType weight) {
    if (full()) return false;
    NVSM_CHECK(features.size() == window_size_, "instance window mismatch");
    std::copy(features.begin(), features.end(), &features_[num_instances_ * window_size_]);
    if (!feature_weights.empty()) {
      NVSM_CHECK(feature_weights.size() == features.size(), "feature weight count mismatch");
      std::copy(feature_weights.begin(), feature_weights.end(), &feature_weights_[num_instances_ * window_size_]);
    } else {
      std::fill(&feature_weights_[num_instances_ * window_size_], &feature_weights_[(num_instances_ + 1) * window_size_],
                static_cast<WeightType>(1.0));
    }
    labels_[num_instances_] = object_id;
    weights_[num_instances_] = weight;
    ++num_instances_;
    return true;
  }

  // Raw access (the reference grants it to friends: DataSource, Objective).
  WordIdxType* features() { return features_; }
  WeightType* feature_weights() { return feature_weights_; }
  ObjectIdxType* labels() { return labels_; }
  WeightType* weights() { return weights_; }
  const WordIdxType* features() const { return features_; }
  const WeightType* feature_weights() const { return feature_weights_; }
  const ObjectIdxType* labels() const { return labels_; }
  const WeightType* weights() const { return weights_; }
  void set_num_instances(size_t n) { NVSM_CHECK(n <= batch_size_, "too many instances"); num_instances_ = n; }

 private:
  template <typename T>
  static void alloc(T** p, size_t count) {
    void* raw = nullptr;
    NVSM_CHECK(nvsm_host_alloc(&raw, count * sizeof(T)) == 0, nvsm_last_error());
    *p = static_cast<T*>(raw);
  }
  const size_t batch_size_, window_size_;
  WordIdxType* features_ = nullptr;
  WeightType* feature_weights_ = nullptr;
  ObjectIdxType* labels_ = nullptr;
  WeightType* weights_ = nullptr;
  size_t num_instances_;
  friend class TextEntity::Objective;
};

typedef ::DataSource<Batch> DataSourceBase;

// Seeded synthetic n-gram source: word ids and positive document ids uniform (or Zipf(s) when
// zipf_exponent > 0), unit weights; `num_batches` full batches per epoch.
class SyntheticSource : public DataSourceBase {
 public:
  SyntheticSource(size_t num_words, size_t num_entities, size_t num_batches, uint64 seed, double zipf_exponent = 0.0)
      : num_words_(num_words), num_entities_(num_entities), num_batches_(num_batches), seed_(seed),
        zipf_(zipf_exponent), emitted_(0), rng_(seed) {
    if (zipf_ > 0.0) { build_cdf(num_words_, &word_cdf_); build_cdf(num_entities_, &entity_cdf_); }
  }
  virtual void reset() override { emitted_ = 0; rng_.seed(seed_); }
  virtual bool has_next() const override { return emitted_ < num_batches_; }
  virtual float32 progress() const override { return static_cast<float32>(emitted_) / num_batches_; }
  virtual void next(Batch* const batch) override {
    batch->clear();
    const size_t B = batch->maximum_size(), n = batch->window_size();
    for (size_t i = 0; i < B * n; ++i) { batch->features()[i] = draw(num_words_, word_cdf_); batch->feature_weights()[i] = 1.0f; }
    for (size_t i = 0; i < B; ++i) { batch->labels()[i] = draw(num_entities_, entity_cdf_); batch->weights()[i] = 1.0f; }
    batch->set_num_instances(B);
    ++emitted_;
  }

 private:
  void build_cdf(size_t n, std::vector<double>* cdf) const {
    cdf->resize(n);
    double acc = 0.0;
    for (size_t k = 0; k < n; ++k) { acc += 1.0 / std::pow(static_cast<double>(k + 1), zipf_); (*cdf)[k] = acc; }
    for (size_t k = 0; k < n; ++k) (*cdf)[k] /= acc;
  }
  long draw(size_t n, const std::vector<double>& cdf) {
    if (cdf.empty()) return static_cast<long>(rng_() % n);
    const double u = std::generate_canonical<double, 53>(rng_);
    return static_cast<long>(std::lower_bound(cdf.begin(), cdf.end(), u) - cdf.begin());
  }
  const size_t num_words_, num_entities_, num_batches_;
  const uint64 seed_;
  const double zipf_;
  size_t emitted_;
  std::mt19937_64 rng_;
  std::vector<double> word_cdf_, entity_cdf_;
};

}  // namespace TextEntity

// reference: RepresentationSimilarity::Batch, include/cuNVSM/data.h:551-614 / cpp/data.cu:157-222 — pairs of object
// ids (features_[2 i], features_[2 i + 1]) with one weight per pair, pinned host memory.
namespace RepresentationSimilarity {

typedef std::tuple<ObjectIdxType, ObjectIdxType, WeightType> InstanceT;

class Batch : public BatchInterface {
 public:
  explicit Batch(const size_t batch_size) : batch_size_(batch_size), num_instances_(0) {
    NVSM_CHECK(batch_size_ > 0, "batch size must be positive");
    void* raw = nullptr;
    NVSM_CHECK(nvsm_host_alloc(&raw, batch_size_ * 2 * sizeof(ObjectIdxType)) == 0, nvsm_last_error());
    features_ = static_cast<ObjectIdxType*>(raw);
    NVSM_CHECK(nvsm_host_alloc(&raw, batch_size_ * sizeof(WeightType)) == 0, nvsm_last_error());
    weights_ = static_cast<WeightType*>(raw);
  }
  explicit Batch(const lse::TrainConfig& train_config) : Batch(static_cast<size_t>(train_config.batch_size())) {}
  Batch(const size_t batch_size, const size_t /* window_size, ignored like the reference */) : Batch(batch_size) {}
  Batch(const std::tuple<size_t>& args) : Batch(std::get<0>(args)) {}
  virtual ~Batch() { nvsm_host_free(features_); nvsm_host_free(weights_); }
  Batch(const Batch&) = delete;
  Batch& operator=(const Batch&) = delete;

  virtual void clear() override { num_instances_ = 0; }
  virtual bool full() const override { return num_instances_ == batch_size_; }
  virtual bool empty() const override { return num_instances_ == 0; }
  virtual size_t num_instances() const override { return num_instances_; }
  virtual size_t maximum_size() const override { return batch_size_; }

  // RepresentationSimilarity::DataSource::next (cpp/data.cu:316-334)
  bool push_instance(const InstanceT& instance) {
    if (full()) return false;
    features_[2 * num_instances_] = std::get<0>(instance);
    features_[2 * num_instances_ + 1] = std::get<1>(instance);
    weights_[num_instances_] = std::get<2>(instance);
    ++num_instances_;
    return true;
  }

  const ObjectIdxType* features() const { return features_; }
  const WeightType* weights() const { return weights_; }

 private:
  const size_t batch_size_;
  ObjectIdxType* features_ = nullptr;
  WeightType* weights_ = nullptr;
  size_t num_instances_;
};

}  // namespace RepresentationSimilarity

namespace EntityEntity { using RepresentationSimilarity::Batch; using RepresentationSimilarity::InstanceT; }
namespace TermTerm { using RepresentationSimilarity::Batch; using RepresentationSimilarity::InstanceT; }

#endif  // CUNVSM_B200_DATA_H
