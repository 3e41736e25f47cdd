// Data-source tests in the manner of the reference's cpp/data_tests.cpp (MetaSourceTest.MultiSource :828-877,
// MetaSourceTest.RepeatingSource :879-908, the similarity source tests). The wrappers are templates over the batch
// type, so the CPU part runs them on a plain batch; `data_test --pinned` repeats the similarity / async part on the
// real page-locked batches (needs the CUDA runtime, i.e. the GPU box).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>

#include "cuNVSM/data.h"

#define EXPECT(cond) do { if (!(cond)) { std::fprintf(stderr, "FAILED %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

struct PlainBatch {           // what the wrappers need from a batch: nothing
  long value = -1;
  void clear() { value = -1; }
};

class CountingSource : public DataSource<PlainBatch> {   // cpp/data_tests.cpp: emits 0, 1, ..., n-1
 public:
  explicit CountingSource(size_t n) : n_(n), i_(0) {}
  virtual void reset() override { i_ = 0; ++resets; }
  virtual void next(PlainBatch* const batch) override { batch->value = static_cast<long>(i_++); }
  virtual bool has_next() const override { return i_ < n_; }
  virtual float32 progress() const override { return static_cast<float32>(i_) / n_; }
  virtual void extract_metadata(lse::Metadata* const metadata) const override { metadata->add_term()->set_model_term_id(static_cast<int>(n_)); }
  int resets = 0;
 private:
  const size_t n_;
  size_t i_;
};

static int test_multi_source() {
  MultiSource<PlainBatch, PlainBatch> source(
      std::make_tuple<DataSource<PlainBatch>*, DataSource<PlainBatch>*>(new CountingSource(8), new CountingSource(9)));
  std::tuple<PlainBatch, PlainBatch> batch;
  size_t idx = 0;
  while (source.has_next()) {
    source.next(&batch);
    EXPECT(std::get<0>(batch).value == static_cast<long>(idx) && std::get<1>(batch).value == static_cast<long>(idx));
    ++idx;
    EXPECT(source.progress() == std::min(idx / 8.0f, idx / 9.0f));
  }
  EXPECT(idx == 8);                       // the shorter source ends the pass
  source.reset();
  EXPECT(source.has_next());
  lse::Metadata meta;
  source.extract_metadata(&meta);         // every constituent contributes
  EXPECT(meta.term_size() == 2 && meta.term(0).model_term_id() == 8 && meta.term(1).model_term_id() == 9);
  return 0;
}

static int test_repeating_source() {
  RepeatingSource<PlainBatch> source(3, new CountingSource(2));
  PlainBatch batch;
  size_t idx = 0;
  while (source.has_next()) {
    source.next(&batch);
    EXPECT(batch.value == static_cast<long>(idx % 2));
    batch.clear();
    ++idx;
  }
  EXPECT(idx == 6);
  source.reset();
  EXPECT(source.has_next());
  RepeatingSource<PlainBatch> forever(static_cast<size_t>(-1), new CountingSource(3));   // cpp/main.cu:254-256
  for (int k = 0; k < 1000; ++k) { EXPECT(forever.has_next()); forever.next(&batch); EXPECT(batch.value == k % 3); }
  return 0;
}

static int test_load_similarities() {
  const IdentifiersMapT ids = {{"doc-a", 0}, {"doc-b", 1}, {"17", 2}};
  std::istringstream file("doc-a doc-b 0.5\ndoc-b unknown 1.0\n17 doc-a 2\n\nmissing doc-a 3\n");
  std::unique_ptr<std::vector<RepresentationSimilarity::InstanceT>> data(RepresentationSimilarity::LoadSimilarities(file, ids));
  EXPECT(data->size() == 2);              // pairs with an unknown identifier are skipped
  EXPECT(data->at(0) == std::make_tuple(0L, 1L, 0.5f));
  EXPECT(data->at(1) == std::make_tuple(2L, 0L, 2.0f));
  return 0;
}

// AsyncSource (cpp/data_async.cpp) on a heap batch: the prefetch / swap / reset protocol without page-locked memory,
// so that it also runs on the CPU (and under -fsanitize=thread).
struct HeapBatch {
  HeapBatch(size_t batch_size, size_t /*window_size*/) : capacity(batch_size) {}
  void clear() { values.clear(); }
  bool empty() const { return values.empty(); }
  void swap(HeapBatch* const other) { values.swap(other->values); std::swap(capacity, other->capacity); }
  size_t capacity;
  std::vector<long> values;
};

class HeapCountingSource : public DataSource<HeapBatch> {   // epoch e emits batches filled with e * 100 + i, i < n
 public:
  explicit HeapCountingSource(size_t n) : n_(n), i_(0), epoch_(0) {}
  virtual void reset() override { i_ = 0; ++epoch_; }
  virtual void next(HeapBatch* const batch) override { batch->values.assign(batch->capacity, static_cast<long>(epoch_ * 100 + i_)); ++i_; }
  virtual bool has_next() const override { return i_ < n_; }
  virtual float32 progress() const override { return static_cast<float32>(i_) / n_; }
 private:
  const size_t n_;
  size_t i_, epoch_;
};

static int test_async_source_protocol() {
  AsyncSource<HeapBatch> async(3, 16, 1, new HeapCountingSource(7));
  HeapBatch batch(16, 1);
  for (long epoch = 0; epoch < 4; ++epoch) {
    long seen = 0;
    while (async.has_next()) {
      batch.clear();
      async.next(&batch);
      EXPECT(batch.values.size() == 16 && batch.values[0] == epoch * 100 + seen && batch.values[15] == epoch * 100 + seen);
      ++seen;
      if (epoch == 2 && seen == 3) break;   // abandon an epoch half way: reset() must drop what was prefetched
    }
    EXPECT(seen == (epoch == 2 ? 3 : 7));
    async.reset();
  }
  return 0;
}

// NGramFileSource weighting strategies (cpp/data_indri.cpp:302-312,640-646; include/cuNVSM/data.h:464-487)
static int test_ngram_file_weighting(const char* dir) {
  const std::string path = std::string(dir) + "/weighting_ngrams.txt";
  {
    std::ofstream f(path);
    f << "# entity w1 w2\n0 1 2\n0 2 3\n0 3 3\n1 0 1 | 2.0\n";   // document 0: 3 n-grams, document 1: 1 n-gram
  }
  long words[2], entity; float ww[2], weight;
  {
    TextEntity::NGramFileSource plain(path, 2, nullptr, true, TextEntity::UNIFORM, TextEntity::UNIFORM_TERM_WEIGHTING);
    plain.instance(3, words, ww, &entity, &weight);
    EXPECT(words[0] == 0 && words[1] == 1 && entity == 1 && ww[0] == 1.0f && ww[1] == 1.0f && weight == 2.0f);
  }
  {
    // no_shuffle + AUTOMATIC -> INV_DOC_FREQUENCY: avg length = 4 n-grams / 2 documents = 2
    TextEntity::NGramFileSource src(path, 2, nullptr, true, TextEntity::AUTOMATIC_WEIGHTING, TextEntity::SELF_INFORMATION_TERM_WEIGHTING);
    src.instance(0, words, ww, &entity, &weight);
    EXPECT(std::fabs(weight - 2.0f / 3.0f) < 1e-6f);
    // 8 word occurrences: id 1 twice, id 2 twice, id 3 three times, id 0 once
    EXPECT(std::fabs(ww[0] - (-std::log(2.0f / 8.0f))) < 1e-6f && std::fabs(ww[1] - (-std::log(2.0f / 8.0f))) < 1e-6f);
    src.instance(2, words, ww, &entity, &weight);
    EXPECT(std::fabs(ww[0] - (-std::log(3.0f / 8.0f))) < 1e-6f);
    src.instance(3, words, ww, &entity, &weight);
    EXPECT(std::fabs(weight - 2.0f * 2.0f) < 1e-5f && std::fabs(ww[0] - (-std::log(1.0f / 8.0f))) < 1e-6f);
    RNG rng; rng.seed(3);
    TextEntity::NGramFileSource shuffled(path, 2, &rng, false, TextEntity::AUTOMATIC_WEIGHTING);   // shuffling -> UNIFORM
    shuffled.instance(0, words, ww, &entity, &weight);
    EXPECT(weight == 1.0f);
  }
  std::remove(path.c_str());
  return 0;
}

// page-locked batches: RepresentationSimilarity::DataSource fills pair batches in the order shuffled by the shared RNG,
// last batch partial, reshuffled at reset(); RepeatingSource keeps it going; AsyncSource delivers the same batches.
static int test_similarity_source_pinned() {
  std::vector<RepresentationSimilarity::InstanceT>* data = new std::vector<RepresentationSimilarity::InstanceT>;
  for (long i = 0; i < 10; ++i) data->push_back(std::make_tuple(i, 100 + i, 0.25f * i));
  RNG rng; rng.seed(7);
  RNG expect_rng; expect_rng.seed(7);
  std::vector<size_t> order(10);
  std::iota(order.begin(), order.end(), 0);
  std::shuffle(order.begin(), order.end(), expect_rng);
  RepresentationSimilarity::DataSource source(data, &rng);
  EXPECT(nvsm_detail::rng_get_state(rng) == nvsm_detail::rng_get_state(expect_rng));   // consumed the shared engine identically
  RepresentationSimilarity::Batch batch(4);
  size_t seen = 0, batches = 0;
  while (source.has_next()) {
    batch.clear();
    source.next(&batch);
    EXPECT(batch.num_instances() == (batches < 2 ? 4u : 2u));
    for (size_t k = 0; k < batch.num_instances(); ++k, ++seen) {
      const long i = static_cast<long>(order[seen]);
      EXPECT(batch.features()[2 * k] == i && batch.features()[2 * k + 1] == 100 + i && batch.weights()[k] == 0.25f * i);
    }
    ++batches;
  }
  EXPECT(seen == 10 && batches == 3 && source.progress() == 1.0f);
  source.reset();
  EXPECT(source.has_next() && nvsm_detail::rng_get_state(rng) != nvsm_detail::rng_get_state(expect_rng));

  // instance overflow (cpp/data_tests.cpp:130-190): a full batch refuses further instances and keeps its contents
  {
    RepresentationSimilarity::Batch small(2);
    EXPECT(small.push_instance(std::make_tuple(1L, 2L, 1.0f)) && small.push_instance(std::make_tuple(3L, 4L, 2.0f)));
    EXPECT(small.full() && !small.push_instance(std::make_tuple(5L, 6L, 3.0f)));
    EXPECT(small.num_instances() == 2 && small.features()[2] == 3 && small.features()[3] == 4 && small.weights()[1] == 2.0f);
    TextEntity::Batch text(2, 3);
    EXPECT(text.push_instance({1, 2, 3}, {}, 7, 1.0f) && text.push_instance({4, 5, 6}, {0.5f, 1.5f, 2.5f}, 8, 0.25f));
    EXPECT(text.full() && !text.push_instance({7, 8, 9}, {}, 9, 1.0f));
    EXPECT(text.features()[3] == 4 && text.feature_weights()[0] == 1.0f && text.feature_weights()[4] == 1.5f &&
           text.labels()[1] == 8 && text.weights()[1] == 0.25f);
    text.clear();
    EXPECT(text.empty() && text.push_instance({9, 9, 9}, {}, 1, 1.0f));
  }

  // TextEntity n-gram batches through AsyncSource == the same source read directly
  struct Seq : public DataSource<TextEntity::Batch> {
    size_t i = 0;
    virtual void reset() override { i = 0; }
    virtual bool has_next() const override { return i < 5; }
    virtual float32 progress() const override { return i / 5.0f; }
    virtual void next(TextEntity::Batch* const b) override {
      for (size_t k = 0; k < b->maximum_size(); ++k) b->push_instance({static_cast<long>(i), static_cast<long>(k)}, {}, static_cast<long>(i), 1.0f);
      ++i;
    }
  };
  AsyncSource<TextEntity::Batch> async(2, 8, 2, new Seq);
  TextEntity::Batch tb(8, 2);
  for (int epoch = 0; epoch < 2; ++epoch) {
    size_t n = 0;
    while (async.has_next()) {
      tb.clear();
      async.next(&tb);
      EXPECT(tb.num_instances() == 8 && tb.features()[0] == static_cast<long>(n) && tb.labels()[7] == static_cast<long>(n));
      ++n;
    }
    EXPECT(n == 5);
    async.reset();
  }
  return 0;
}

int main(int argc, char** argv) {
  const char* tmp = std::getenv("TMPDIR");
  int failed = test_multi_source() + test_repeating_source() + test_load_similarities() + test_async_source_protocol() +
               test_ngram_file_weighting(tmp ? tmp : "/tmp");
  if (argc > 1 && std::strcmp(argv[1], "--pinned") == 0) failed += test_similarity_source_pinned();
  if (failed) return 1;
  std::printf("data tests ok%s\n", argc > 1 ? " (pinned batches included)" : "");
  return 0;
}
