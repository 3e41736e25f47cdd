// cuNVSM/data.h — host-side batch of the TextEntity objective and the data-source interface.
// TextEntity::Batch has the reference's layout (reference: include/cuNVSM/data.h:114-177,
// cpp/data.cu:8-30,94-124): four page-locked arrays, instance-major,
//   features_[B*n] (long), feature_weights_[B*n], labels_[B] (long), weights_[B].
// The Indri-backed sources of the reference are out of scope; SyntheticSource generates seeded uniform / Zipf n-grams
// and NGramFileSource reads pre-tokenised n-grams, both with the DataSource contract (has_next / next / reset /
// progress); AsyncSource is the reference's prefetching wrapper (cpp/data_async.cpp).
#ifndef CUNVSM_B200_DATA_H
#define CUNVSM_B200_DATA_H

#include <algorithm>
#include <atomic>
#include <cmath>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <deque>
#include <fstream>
#include <limits>
#include <map>
#include <memory>
#include <mutex>
#include <numeric>
#include <sstream>
#include <string>
#include <thread>
#include <tuple>
#include <utility>
#include <vector>

#include "../nvsm_b200.h"
#include "base.h"
#include "nvsm.pb.h"

typedef int32 WordIdxType;
typedef int32 ObjectIdxType;
typedef FLOATING_POINT_TYPE WeightType;
// external identifier (document name / term string) -> model id; reference: IdentifiersMapT, include/cuNVSM/data.h
typedef std::map<std::string, ObjectIdxType> IdentifiersMapT;

class BatchInterface {
 public:
  virtual ~BatchInterface() {}
  virtual void clear() = 0;
  virtual bool full() const = 0;
  virtual bool empty() const = 0;
  virtual size_t num_instances() const = 0;
  virtual size_t maximum_size() const = 0;
  virtual void swap(BatchInterface* const other) = 0;   // reference: include/cuNVSM/data.h:57
};

template <typename BatchT>
class DataSource {
 public:
  virtual ~DataSource() {}
  virtual void reset() = 0;
  virtual void next(BatchT* const batch) = 0;
  virtual bool has_next() const = 0;
  virtual float32 progress() const = 0;
  // reference: DataSourceInterface::extract_metadata, include/cuNVSM/data.h:75 (sources without a mapping leave it empty)
  virtual void extract_metadata(lse::Metadata* const metadata) const { (void)metadata; }
};

namespace nvsm_detail {
// Metadata of a source whose ids are already model ids: term i <-> i (with its frequency), object j <-> j.
inline void identity_metadata(const size_t num_words, const size_t num_entities, const long* const term_frequency,
                              const long total_terms, lse::Metadata* const metadata) {
  for (size_t i = 0; i < num_words; ++i) {
    lse::Metadata::TermInfo* const term = metadata->add_term();
    term->set_index_term_id(static_cast<int>(i));
    term->set_model_term_id(static_cast<int>(i));
    term->set_term_frequency(term_frequency != nullptr ? static_cast<int>(term_frequency[i]) : 0);
  }
  metadata->set_total_terms(static_cast<int>(total_terms));
  for (size_t j = 0; j < num_entities; ++j) {
    lse::Metadata::ObjectInfo* const object = metadata->add_object();
    object->set_model_object_id(static_cast<int>(j));
    object->set_index_object_id(static_cast<int>(j));
  }
}
}  // namespace nvsm_detail

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

  virtual void clear() override { num_instances_ = 0; uniform_feature_weights_ = uniform_weights_ = false; }
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
  // A source whose weighting is uniform (the reference's default: every feature weight / instance weight is 1.0,
  // include/cuNVSM/data.h:465-467) says so after filling the batch; Model::compute_cost then passes NULL for that array
  // and the library fills ones on the device instead of copying them over PCIe. Reset by clear().
  void set_uniform_weights(const bool feature_weights, const bool instance_weights) {
    uniform_feature_weights_ = feature_weights; uniform_weights_ = instance_weights;
  }
  bool uniform_feature_weights() const { return uniform_feature_weights_; }
  bool uniform_weights() const { return uniform_weights_; }

  // reference: TextEntity::Batch::swap, cpp/data.cu:76-92 — exchanges the pinned arrays, no copy
  virtual void swap(BatchInterface* const other) override {
    Batch* const o = dynamic_cast<Batch*>(other);
    NVSM_CHECK(o != nullptr, "swap with a different batch type");
    NVSM_CHECK(batch_size_ == o->batch_size_ && window_size_ == o->window_size_, "swap needs equally sized batches");
    std::swap(features_, o->features_); std::swap(feature_weights_, o->feature_weights_);
    std::swap(labels_, o->labels_); std::swap(weights_, o->weights_);
    std::swap(num_instances_, o->num_instances_);
    std::swap(uniform_feature_weights_, o->uniform_feature_weights_); std::swap(uniform_weights_, o->uniform_weights_);
  }

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
  bool uniform_feature_weights_ = false, uniform_weights_ = false;
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
    batch->set_uniform_weights(true, true);
    ++emitted_;
  }
  // identity id mapping; synthetic ids carry no corpus statistics
  virtual void extract_metadata(lse::Metadata* const metadata) const override {
    nvsm_detail::identity_metadata(num_words_, num_entities_, nullptr, 0, metadata);
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

// Pre-tokenised n-gram file (stands in for IndriSource, cpp/data_indri.cpp, whose Indri index is out of scope):
// one instance per line, `<entity_id> <word_id_1> ... <word_id_n>` with an optional `| <instance_weight>` tail;
// lines starting with '#' are comments. Like IndriSource::reset (cpp/data_indri.cpp:386-397) every epoch walks the
// instances in an order shuffled with the SHARED RNG (unless no_shuffle), so the engine consumption order
// init -> shuffle -> negatives of the reference is kept. Only full batches are emitted (the CLI skips others).
// reference: include/cuNVSM/data.h:373-379
enum WeightingStrategy { AUTOMATIC_WEIGHTING, UNIFORM, INV_DOC_FREQUENCY };
enum TermWeightingStrategy { UNIFORM_TERM_WEIGHTING, SELF_INFORMATION_TERM_WEIGHTING };

class NGramFileSource : public DataSourceBase {
 public:
  // Instance weights (--weighting): INV_DOC_FREQUENCY = exp(log(avg_document_length) - log(document_length))
  // (cpp/data_indri.cpp:302-312; a document's length is the number of its n-grams in the file), AUTOMATIC =
  // UNIFORM when shuffling, INV_DOC_FREQUENCY otherwise (:640-646); an explicit `| weight` on a line multiplies it.
  // Word weights (--feature_weighting): SELF_INFORMATION = -log(tf / total_terms) (include/cuNVSM/data.h:464-487).
  NGramFileSource(const std::string& path, const size_t window_size, RNG* const rng, const bool no_shuffle = false,
                  WeightingStrategy weighting_strategy = UNIFORM,
                  const TermWeightingStrategy term_weighting_strategy = UNIFORM_TERM_WEIGHTING)
      : window_size_(window_size), rng_(rng), no_shuffle_(no_shuffle), position_(0), max_word_(-1), max_entity_(-1) {
    std::ifstream file(path);
    NVSM_CHECK(file.good(), ("cannot open n-gram file " + path).c_str());
    std::string line;
    while (std::getline(file, line)) {
      if (line.empty() || line[0] == '#') continue;
      float weight = 1.0f;
      const size_t bar = line.find('|');
      if (bar != std::string::npos) { weight = std::stof(line.substr(bar + 1)); line = line.substr(0, bar); }
      std::istringstream iss(line);
      long entity = -1;
      iss >> entity;
      NVSM_CHECK(!iss.fail() && entity >= 0, "malformed n-gram line: entity id");
      for (size_t w = 0; w < window_size_; ++w) {
        long id = -1;
        iss >> id;
        NVSM_CHECK(!iss.fail() && id >= 0, "malformed n-gram line: fewer word ids than the window size");
        words_.push_back(id);
        max_word_ = std::max(max_word_, id);
      }
      entities_.push_back(entity);
      weights_.push_back(weight);
      max_entity_ = std::max(max_entity_, entity);
    }
    NVSM_CHECK(!entities_.empty(), "empty n-gram file");
    order_.resize(entities_.size());
    std::iota(order_.begin(), order_.end(), 0);
    if (weighting_strategy == AUTOMATIC_WEIGHTING) weighting_strategy = no_shuffle_ ? INV_DOC_FREQUENCY : UNIFORM;
    if (weighting_strategy == INV_DOC_FREQUENCY) {
      std::vector<long> length(corpus_size(), 0);
      for (const long e : entities_) ++length[e];
      size_t documents = 0;
      for (const long l : length) documents += l > 0;
      const WeightType avg_document_length = static_cast<WeightType>(entities_.size()) / documents;
      for (size_t k = 0; k < entities_.size(); ++k)
        weights_[k] *= std::exp(std::log(avg_document_length) - std::log(static_cast<WeightType>(length[entities_[k]])));
    }
    instance_weights_uniform_ = true;
    for (const float w : weights_) instance_weights_uniform_ = instance_weights_uniform_ && w == 1.0f;
    if (term_weighting_strategy == SELF_INFORMATION_TERM_WEIGHTING) {
      std::vector<long> frequency(vocabulary_size(), 0);
      for (const long w : words_) ++frequency[w];
      word_weights_.resize(vocabulary_size());
      for (size_t w = 0; w < word_weights_.size(); ++w)
        word_weights_[w] = frequency[w] > 0 ? -std::log(static_cast<WeightType>(frequency[w]) / static_cast<WeightType>(words_.size())) : 0;
    }
  }
  // Instance k of the file as next() delivers it: words[n], word weights[n], entity, instance weight.
  void instance(const size_t k, long* const words, WeightType* const word_weights, long* const entity, WeightType* const weight) const {
    for (size_t w = 0; w < window_size_; ++w) {
      words[w] = words_[k * window_size_ + w];
      word_weights[w] = word_weights_.empty() ? static_cast<WeightType>(1.0) : word_weights_[words[w]];
    }
    *entity = entities_[k];
    *weight = weights_[k];
  }
  size_t num_instances() const { return entities_.size(); }
  size_t vocabulary_size() const { return static_cast<size_t>(max_word_ + 1); }
  size_t corpus_size() const { return static_cast<size_t>(max_entity_ + 1); }

  virtual void reset() override {
    position_ = 0;
    std::iota(order_.begin(), order_.end(), 0);
    if (!no_shuffle_ && rng_ != nullptr) std::shuffle(order_.begin(), order_.end(), *rng_);
  }
  // (batch size is only known at next(): ask "is there any instance left", like the reference's sources)
  virtual bool has_next() const override { return position_ < order_.size(); }
  virtual float32 progress() const override { return static_cast<float32>(position_) / order_.size(); }
  virtual void next(Batch* const batch) override {
    batch->clear();
    NVSM_CHECK(batch->window_size() == window_size_, "batch window does not match the n-gram file");
    const size_t B = std::min(batch->maximum_size(), order_.size() - position_);
    for (size_t i = 0; i < B; ++i)
      instance(order_[position_ + i], &batch->features()[i * window_size_], &batch->feature_weights()[i * window_size_],
               &batch->labels()[i], &batch->weights()[i]);
    batch->set_num_instances(B);
    batch->set_uniform_weights(word_weights_.empty(), instance_weights_uniform_);
    position_ += B;
  }
  // reference: IndriSource::extract_metadata, cpp/data_indri.cpp:534-555. The file is pre-tokenised, so index ids ==
  // model ids; term_frequency = occurrences of the word in the file, total_terms = all occurrences (what the
  // self-information weighting of py/nvsm/base.py:query_representation divides by).
  virtual void extract_metadata(lse::Metadata* const metadata) const override {
    std::vector<long> frequency(vocabulary_size(), 0);
    for (const long w : words_) ++frequency[w];
    nvsm_detail::identity_metadata(vocabulary_size(), corpus_size(), frequency.data(), static_cast<long>(words_.size()), metadata);
  }

 private:
  const size_t window_size_;
  RNG* const rng_;
  const bool no_shuffle_;
  size_t position_;
  long max_word_, max_entity_;
  std::vector<long> words_, entities_;
  std::vector<float> weights_;
  std::vector<WeightType> word_weights_;   // per word id; empty = uniform term weighting
  bool instance_weights_uniform_ = false;  // every instance weight is exactly 1.0 (uniform weighting, unweighted file)
  std::vector<size_t> order_;
};

}  // namespace TextEntity

// reference: AsyncSource, include/cuNVSM/data.h:663-711 / cpp/data_async.cpp — a worker thread keeps
// `num_concurrent_batches` pinned batches filled from the wrapped source; next() SWAPS the caller's empty batch with a
// full buffer (no copy). reset() restarts the worker after resetting the source on the calling thread (which is where
// the shared RNG is touched, as in the reference). Condition variables instead of the reference's yield-spinning
// on boost::lockfree queues.
template <typename BatchT>
class AsyncSource : public DataSource<BatchT> {
 public:
  typedef BatchT BatchType;

  // Takes ownership of `source`.
  AsyncSource(const size_t num_concurrent_batches, const size_t batch_size, const size_t window_size,
              DataSource<BatchT>* const source)
      : source_(source), stop_(false), running_(false) {
    NVSM_CHECK(num_concurrent_batches > 0, "need at least one buffer");
    for (size_t i = 0; i < num_concurrent_batches; ++i) {
      buffers_.emplace_back(new BatchT(batch_size, window_size));
      empty_.push_back(buffers_.back().get());
    }
    start_worker();
  }
  virtual ~AsyncSource() { stop_worker(); }

  virtual void reset() override {
    stop_worker();
    source_->reset();
    for (BatchT* b : full_) { b->clear(); empty_.push_back(b); }   // drop batches prefetched from the old epoch
    full_.clear();
    start_worker();
  }
  virtual bool has_next() const override {
    std::unique_lock<std::mutex> lock(mutex_);
    cv_.wait(lock, [&] { return !full_.empty() || !running_; });
    return !full_.empty();
  }
  virtual void next(BatchT* const batch) override {
    NVSM_CHECK(batch->empty(), "AsyncSource::next needs an empty batch");
    std::unique_lock<std::mutex> lock(mutex_);
    cv_.wait(lock, [&] { return !full_.empty() || !running_; });
    NVSM_CHECK(!full_.empty(), "AsyncSource::next called without has_next");
    BatchT* const buffer = full_.front();
    full_.pop_front();
    batch->swap(buffer);
    buffer->clear();
    empty_.push_back(buffer);
    cv_.notify_all();
  }
  virtual float32 progress() const override { return source_->progress(); }
  virtual void extract_metadata(lse::Metadata* const metadata) const override { source_->extract_metadata(metadata); }  // cpp/data_async.cpp

 private:
  void start_worker() {
    stop_ = false;
    running_ = true;
    thread_.reset(new std::thread([this] {
      for (;;) {
        BatchT* buffer = nullptr;
        {
          std::unique_lock<std::mutex> lock(mutex_);
          cv_.wait(lock, [&] { return stop_ || !empty_.empty(); });
          if (stop_ || !source_->has_next()) break;
          buffer = empty_.front();
          empty_.pop_front();
        }
        source_->next(buffer);   // outside the lock: this is the expensive part
        {
          std::lock_guard<std::mutex> lock(mutex_);
          full_.push_back(buffer);
        }
        cv_.notify_all();
      }
      {
        std::lock_guard<std::mutex> lock(mutex_);
        running_ = false;
      }
      cv_.notify_all();
    }));
  }
  void stop_worker() {
    {
      std::lock_guard<std::mutex> lock(mutex_);
      stop_ = true;
    }
    cv_.notify_all();
    if (thread_ && thread_->joinable()) thread_->join();
    thread_.reset();
  }

  std::unique_ptr<DataSource<BatchT>> source_;
  std::vector<std::unique_ptr<BatchT>> buffers_;
  std::deque<BatchT*> empty_, full_;
  mutable std::mutex mutex_;
  mutable std::condition_variable cv_;
  std::unique_ptr<std::thread> thread_;
  bool stop_, running_;
};

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

  virtual void swap(BatchInterface* const other) override {   // cpp/data.cu:196-210
    Batch* const o = dynamic_cast<Batch*>(other);
    NVSM_CHECK(o != nullptr && batch_size_ == o->batch_size_, "swap needs an equally sized batch of the same type");
    std::swap(features_, o->features_); std::swap(weights_, o->weights_);
    std::swap(num_instances_, o->num_instances_);
  }

 private:
  const size_t batch_size_;
  ObjectIdxType* features_ = nullptr;
  WeightType* weights_ = nullptr;
  size_t num_instances_;
};

// reference: LoadSimilarities, cpp/data.cu:233-285 — one `<first> <second> <weight>` triple per line, identifiers
// resolved through the map; pairs with an unknown identifier are skipped (with a warning).
inline std::vector<InstanceT>* LoadSimilarities(std::istream& file, const IdentifiersMapT& identifiers_map) {
  NVSM_CHECK(file.good(), "cannot read the similarity file");
  NVSM_CHECK(!identifiers_map.empty(), "empty identifiers map");
  std::vector<InstanceT>* const data = new std::vector<InstanceT>;
  std::string line;
  while (file.good() && std::getline(file, line)) {
    std::istringstream iss(line);
    std::string first, second;
    WeightType weight = 0;
    iss >> first >> second >> weight;
    if (first.empty() && second.empty()) continue;
    const auto a = identifiers_map.find(first), b = identifiers_map.find(second);
    if (a == identifiers_map.end() || b == identifiers_map.end()) {
      std::fprintf(stderr, "Entity '%s' not found; skipping pair.\n", (a == identifiers_map.end() ? first : second).c_str());
      continue;
    }
    data->push_back(std::make_tuple(a->second, b->second, weight));
  }
  return data;
}
inline std::vector<InstanceT>* LoadSimilarities(const std::string& path, const IdentifiersMapT& identifiers_map) {
  NVSM_CHECK(!path.empty(), "empty similarity path");
  std::ifstream file(path);
  return LoadSimilarities(file, identifiers_map);
}

// reference: RepresentationSimilarity::DataSource, include/cuNVSM/data.h:626-653 / cpp/data.cu:287-345 — the pairs in
// an order shuffled with the SHARED RNG at construction and at every reset(); the last batch of a pass may be partial.
class DataSource : public ::DataSource<Batch> {
 public:
  DataSource(const std::string& path, const IdentifiersMapT& identifiers_map, RNG* const rng)
      : DataSource(LoadSimilarities(path, identifiers_map), rng) {}
  // Takes ownership.
  DataSource(const std::vector<InstanceT>* const data, RNG* const rng) : data_(data), rng_(rng) {
    NVSM_CHECK(data_ != nullptr && rng_ != nullptr, "null similarity data or RNG");
    reset();
  }
  virtual void reset() override {
    instance_order_.resize(data_->size());
    std::iota(instance_order_.begin(), instance_order_.end(), 0);
    std::shuffle(instance_order_.begin(), instance_order_.end(), *rng_);
  }
  virtual void next(Batch* const batch) override {
    NVSM_CHECK(batch->empty(), "RepresentationSimilarity::DataSource::next needs an empty batch");
    while (!batch->full() && !instance_order_.empty()) {
      batch->push_instance(data_->at(instance_order_.front()));
      instance_order_.pop_front();
    }
  }
  virtual bool has_next() const override { return !instance_order_.empty(); }
  virtual float32 progress() const override {
    return 1.0f - static_cast<float32>(instance_order_.size()) / static_cast<float32>(data_->size());
  }
  size_t num_instances() const { return data_->size(); }

 private:
  std::unique_ptr<const std::vector<InstanceT>> data_;
  RNG* const rng_;
  std::deque<size_t> instance_order_;
};

}  // namespace RepresentationSimilarity

// reference: RepeatingSource, include/cuNVSM/data.h:737-761 / cpp/data_repeating.cpp — replays the wrapped source
// `num_repeats` times (size_t(-1): practically for ever, how the CLI pairs a short similarity file with a long text
// epoch, cpp/main.cu:254-256). Takes ownership.
template <typename BatchT>
class RepeatingSource : public DataSource<BatchT> {
 public:
  typedef BatchT BatchType;
  RepeatingSource(const size_t num_repeats, DataSource<BatchT>* const source)
      : source_(source), num_repeats_(num_repeats), current_iteration_(0) {}
  virtual void reset() override { current_iteration_ = 0; source_->reset(); }
  virtual void next(BatchT* const batch) override {
    if (!source_->has_next()) {
      source_->reset();
      ++current_iteration_;
      NVSM_CHECK(current_iteration_ < num_repeats_, "RepeatingSource::next called without has_next");
    }
    source_->next(batch);
  }
  virtual bool has_next() const override {
    if (current_iteration_ + 1 < num_repeats_) return true;
    NVSM_CHECK(current_iteration_ + 1 == num_repeats_, "RepeatingSource ran past its last repeat");
    return source_->has_next();
  }
  virtual float32 progress() const override {
    return source_->progress() + static_cast<float32>(current_iteration_) / num_repeats_;
  }
  virtual void extract_metadata(lse::Metadata* const metadata) const override { source_->extract_metadata(metadata); }

 private:
  std::unique_ptr<DataSource<BatchT>> source_;
  const size_t num_repeats_;
  size_t current_iteration_;
};

// reference: MultiSource, include/cuNVSM/data.h:713-735 / cpp/data_multi.cpp — one source per element of a batch tuple:
// next() advances all of them, has_next() needs all of them, progress() is the slowest one. Takes ownership.
template <typename... BatchT>
class MultiSource : public DataSource<std::tuple<BatchT...>> {
 public:
  typedef std::tuple<BatchT...> BatchType;
  explicit MultiSource(const std::tuple<DataSource<BatchT>*...>& sources) { adopt(sources, std::index_sequence_for<BatchT...>()); }
  virtual void reset() override { for_each([](auto& s) { s->reset(); }); }
  virtual void next(BatchType* const batch) override { next_impl(batch, std::index_sequence_for<BatchT...>()); }
  virtual bool has_next() const override {
    bool value = true;
    for_each([&](auto& s) { value = value && s->has_next(); });
    return value;
  }
  virtual float32 progress() const override {
    float32 value = 1.0f;
    for_each([&](auto& s) { value = std::min(value, s->progress()); });
    return value;
  }
  virtual void extract_metadata(lse::Metadata* const metadata) const override {
    for_each([&](auto& s) { s->extract_metadata(metadata); });
  }

 private:
  template <size_t... I>
  void adopt(const std::tuple<DataSource<BatchT>*...>& sources, std::index_sequence<I...>) {
    (void)std::initializer_list<int>{(std::get<I>(sources_).reset(std::get<I>(sources)), 0)...};
  }
  template <size_t... I>
  void next_impl(BatchType* const batch, std::index_sequence<I...>) {
    (void)std::initializer_list<int>{(std::get<I>(sources_)->next(&std::get<I>(*batch)), 0)...};
  }
  template <typename Fn>
  void for_each(Fn fn) const { std::apply([&](auto&... s) { (void)std::initializer_list<int>{(fn(s), 0)...}; }, sources_); }
  std::tuple<std::unique_ptr<DataSource<BatchT>>...> sources_;
};

namespace EntityEntity { using RepresentationSimilarity::Batch; using RepresentationSimilarity::InstanceT; using RepresentationSimilarity::DataSource; }
namespace TermTerm { using RepresentationSimilarity::Batch; using RepresentationSimilarity::InstanceT; using RepresentationSimilarity::DataSource; }

#endif  // CUNVSM_B200_DATA_H
