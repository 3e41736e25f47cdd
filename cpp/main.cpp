// cuNVSMTrainModel — training CLI with the reference's flag surface (reference: cpp/main.cu:15-76,
// 623-768) driving the B200-native step through the Model façade (include/cuNVSM/model.h).
//
// The reference reads an Indri index; Indri is out of scope here, so the positional argument is
// replaced by a seeded synthetic n-gram source (--synthetic_* flags). Everything else keeps the
// reference's meaning: per-batch compute_cost / compute_gradients / update / get_cost
// (iterate_data, cpp/main.cu:366-469), batches that are not a multiple of 1024 instances are
// skipped (:392-398), default learning rates 0.01 (SGD/Adagrad) / 0.001 (Adam) (:710-721),
// clip_sigmoid forced on (:645), --seed must be > 0 (:708), per-epoch batches/second logging
// (:604-612) plus n-grams/second.
#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "cuNVSM/gradient_check.h"
#include "cuNVSM/model.h"

// NVTX ranges with the reference's names (Epoch / Batch / FetchData / ComputeCost / ComputeGradients /
// UpdateParameters, cpp/main.cu:386-431,463,582,619). nvtx3 is header-only and a no-op without a profiler attached.
#if defined(__has_include)
#if __has_include(<nvtx3/nvToolsExt.h>)
#include <nvtx3/nvToolsExt.h>
#define NVSM_HAVE_NVTX 1
#endif
#endif

namespace {

struct NvtxRange {
#ifdef NVSM_HAVE_NVTX
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
#else
  explicit NvtxRange(const char*) {}
#endif
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

struct Flags {
  std::map<std::string, std::string> values;
  Flags() {
    values = {{"num_epochs", "100000"}, {"word_repr_size", "4"}, {"entity_repr_size", "4"}, {"batch_size", "1024"},
              {"window_size", "8"}, {"num_random_entities", "1"}, {"seed", "0"}, {"regularization_lambda", "0.01"},
              {"learning_rate", "0.0"}, {"update_method", ""}, {"weighting", "auto"}, {"feature_weighting", "uniform"},
              {"bias_negative_samples", "false"}, {"nonlinearity", ""}, {"l2_phrase_normalization", "false"},
              {"l2_entity_normalization", "false"}, {"batch_normalization", "false"}, {"compute_initial_cost", "false"},
              {"check_gradients", "false"}, {"gradient_check_epsilon", "1e-2"}, {"no_shuffle", "false"}, {"dump_initial_model", "false"}, {"dump_every", "0"},
              {"entity_similarity_weight", "0.0"}, {"term_similarity_weight", "0.0"}, {"output", ""},
              // replacements for the Indri positional argument
              {"synthetic_num_words", "50000"}, {"synthetic_num_entities", "50000"}, {"synthetic_num_batches", "100"},
              {"synthetic_zipf", "0.0"}, {"device", "0"}, {"gemm", "3xtf32"}, {"host_sampler", "false"}, {"v", "0"},
              // pre-tokenised n-gram file (`<entity> <w_1> ... <w_n> [| weight]` per line) instead of the synthetic source,
              // prefetched by an AsyncSource with this many pinned batches (the reference uses 10, cpp/main.cu:212-219)
              {"ngram_file", ""}, {"num_concurrent_batches", "10"},
              // negatives ~ Zipf(s) over the entity ids instead of the reference's uniform draws (0 = uniform)
              {"negative_sampling_zipf", "0.0"},
              // `<id_a> <id_b> <weight>` pairs for the similarity objectives (DataConfig.similarity_path of the reference);
              // ids are the entity (or word) ids of the n-gram file / synthetic source
              {"similarity_file", ""}};
  }
  void parse(int argc, char** argv) {
    for (int i = 1; i < argc; ++i) {
      std::string a = argv[i];
      NVSM_CHECK(a.rfind("--", 0) == 0, ("unexpected argument " + a).c_str());
      a = a.substr(2);
      std::string key = a, val;
      const size_t eq = a.find('=');
      if (eq != std::string::npos) { key = a.substr(0, eq); val = a.substr(eq + 1); }
      else if (key.rfind("no", 0) == 0 && values.count(key.substr(2)) && is_bool(key.substr(2))) { key = key.substr(2); val = "false"; }
      else if (values.count(key) && is_bool(key)) { val = "true"; }
      else { NVSM_CHECK(i + 1 < argc, ("missing value for --" + key).c_str()); val = argv[++i]; }
      NVSM_CHECK(values.count(key), ("unknown flag --" + key).c_str());
      values[key] = val;
    }
  }
  bool is_bool(const std::string& k) const { const std::string& v = values.at(k); return v == "true" || v == "false"; }
  std::string str(const std::string& k) const { return values.at(k); }
  long i(const std::string& k) const { return std::stol(values.at(k)); }
  double d(const std::string& k) const { return std::stod(values.at(k)); }
  bool b(const std::string& k) const { return values.at(k) == "true" || values.at(k) == "1"; }
};

// Write the four parameter tensors as <output>_<suffix>.<name>.npy, row-major [objects, dim] — the
// shapes the reference's HDF5 dump uses (cpp/hdf5.cu:26-53, lse_hdf5_inl.h:4-27).
template <typename ModelT>
void dump_model(const ModelT& model, const std::string& output, const std::string& suffix) {
  if (output.empty()) return;
  for (const auto& kv : model.get_data()) {
    const std::string path = output + "_" + suffix + "." + kv.first + ".npy";
    std::ofstream f(path, std::ios::binary);
    std::string hdr = "{'descr': '<f4', 'fortran_order': False, 'shape': (" + std::to_string(kv.second.cols) + ", " +
                      std::to_string(kv.second.rows) + "), }";
    while ((10 + hdr.size() + 1) % 64 != 0) hdr += ' ';
    hdr += '\n';
    const unsigned short len = static_cast<unsigned short>(hdr.size());
    f.write("\x93NUMPY\x01\x00", 8);
    f.write(reinterpret_cast<const char*>(&len), 2);
    f.write(hdr.data(), hdr.size());
    f.write(reinterpret_cast<const char*>(kv.second.data.data()), kv.second.data.size() * sizeof(float));
  }
  std::printf("Dumped model to %s_%s.*.npy\n", output.c_str(), suffix.c_str());
}


// Seeded synthetic similarity source (used when no --similarity_file is given): `num_batches` full batches of uniform
// random id pairs with unit weights per epoch.
class SyntheticSimilaritySource : public DataSource<RepresentationSimilarity::Batch> {
 public:
  SyntheticSimilaritySource(size_t num_objects, size_t num_batches, unsigned long seed)
      : num_objects_(num_objects), num_batches_(num_batches), seed_(seed), emitted_(0), rng_(seed) {}
  virtual void reset() override { emitted_ = 0; rng_.seed(seed_); }
  virtual bool has_next() const override { return emitted_ < num_batches_; }
  virtual float32 progress() const override { return static_cast<float32>(emitted_) / num_batches_; }
  virtual void next(RepresentationSimilarity::Batch* batch) override {
    std::uniform_int_distribution<long> pick(0, static_cast<long>(num_objects_) - 1);
    while (!batch->full()) batch->push_instance(std::make_tuple(pick(rng_), pick(rng_), 1.0f));
    ++emitted_;
  }
 private:
  const size_t num_objects_, num_batches_;
  const unsigned long seed_;
  size_t emitted_;
  std::mt19937_64 rng_;
};

struct Sources {
  DataSource<TextEntity::Batch>* text;
  DataSource<RepresentationSimilarity::Batch>* pairs;
};

// BatchHandler of the reference (cpp/main.cu:159-228): uniform access to the batch type of every objective.
template <typename ObjectiveT>
struct BatchOps;

template <>
struct BatchOps<TextEntity::Objective> {
  typedef TextEntity::Batch BatchT;
  static const bool has_text = true, pairs_over_entities = true;
  static BatchT* make(const lse::TrainConfig& tc) { return new BatchT(tc); }
  static void clear(BatchT* b) { b->clear(); }
  static void next(Sources& s, BatchT* b) { s.text->next(b); }
  static bool has_next(Sources& s) { return s.text->has_next(); }
  static void reset(Sources& s) { s.text->reset(); }
  static size_t num_instances(const BatchT& b) { return b.num_instances(); }
  // --check_gradients (cpp/main.cu:414-420): every parameter, central differences, abort on failure
  template <typename ModelT, typename ResultT>
  static void check_gradients(ModelT* model, const BatchT& batch, const ResultT& result, const TextEntity::Gradients& gradients,
                              const double epsilon, const std::stringstream& rng_state, RNG* rng, const bool verbose) {
    GradientCheckFn<ModelT> check;
    const bool ok = check(model, batch, result, gradients, static_cast<float>(epsilon), 1e-1f /* relative_error_threshold */,
                          rng_state, rng, 2e-6, verbose ? 1 : 0);
    const auto& r = check.report();
    std::printf("Gradient check: %zu parameters checked, %zu below the noise floor, worst relative error %.3g (%s)\n", r.checked,
                r.skipped, r.worst_relative_error, r.worst.c_str());
    NVSM_CHECK(ok, "Gradient check failed.");
  }
};

template <int K>
struct BatchOps<RepresentationSimilarity::ObjectiveT<K>> {
  typedef RepresentationSimilarity::Batch BatchT;
  static const bool has_text = false, pairs_over_entities = (K == NVSM_OBJECTIVE_ENTITY_ENTITY);
  static BatchT* make(const lse::TrainConfig& tc) { return new BatchT(tc); }
  static void clear(BatchT* b) { b->clear(); }
  static void next(Sources& s, BatchT* b) { s.pairs->next(b); }
  static bool has_next(Sources& s) { return s.pairs->has_next(); }
  static void reset(Sources& s) { s.pairs->reset(); }
  static size_t num_instances(const BatchT& b) { return b.num_instances(); }
  template <typename ModelT, typename ResultT>
  static void check_gradients(ModelT*, const BatchT&, const ResultT&, const TextEntity::Gradients&, double, const std::stringstream&, RNG*, bool) {
    NVSM_CHECK(false, "--check_gradients covers the TextEntity objective (no similarity weights)");
  }
};

template <int K>
struct BatchOps<MixtureObjectiveT<K>> {
  typedef std::tuple<TextEntity::Batch, RepresentationSimilarity::Batch> BatchT;
  static const bool has_text = true, pairs_over_entities = (K == NVSM_OBJECTIVE_TEXT_ENTITY_ENTITY_ENTITY);
  static BatchT* make(const lse::TrainConfig& tc) { return new BatchT(tc, tc); }
  static void clear(BatchT* b) { std::get<0>(*b).clear(); std::get<1>(*b).clear(); }
  static void next(Sources& s, BatchT* b) { s.text->next(&std::get<0>(*b)); s.pairs->next(&std::get<1>(*b)); }
  static bool has_next(Sources& s) { return s.text->has_next() && s.pairs->has_next(); }   // MultiSource semantics
  static void reset(Sources& s) { s.text->reset(); s.pairs->reset(); }
  static size_t num_instances(const BatchT& b) { return std::get<0>(b).num_instances(); }
  template <typename ModelT, typename ResultT>
  static void check_gradients(ModelT*, const BatchT&, const ResultT&, const TextEntity::Gradients&, double, const std::stringstream&, RNG*, bool) {
    NVSM_CHECK(false, "--check_gradients covers the TextEntity objective (no similarity weights)");
  }
};

// train<ObjectiveT> of the reference (cpp/main.cu:471-621)
template <typename ObjectiveT>
int train(const Flags& flags, const lse::ModelDesc& model_desc, const lse::TrainConfig& train_config_in) {
  typedef BatchOps<ObjectiveT> Ops;
  lse::TrainConfig train_config = train_config_in;
  RNG rng;
  rng.seed(flags.i("seed"));

  size_t V = flags.i("synthetic_num_words"), D = flags.i("synthetic_num_entities");
  std::unique_ptr<DataSource<TextEntity::Batch>> data_source_ptr;
  const bool file_source = !flags.str("ngram_file").empty();
  if (file_source) {
    // construct_data_source + wrap_source_async of the reference (cpp/main.cu:212-240), on an n-gram file
    const std::map<std::string, TextEntity::WeightingStrategy> WEIGHTING_STRATEGIES = {   // cpp/main.cu:136-146
        {"auto", TextEntity::AUTOMATIC_WEIGHTING}, {"uniform", TextEntity::UNIFORM}, {"inv_doc_frequency", TextEntity::INV_DOC_FREQUENCY}};
    const std::map<std::string, TextEntity::TermWeightingStrategy> FEATURE_WEIGHTING_STRATEGIES = {
        {"uniform", TextEntity::UNIFORM_TERM_WEIGHTING}, {"self_information", TextEntity::SELF_INFORMATION_TERM_WEIGHTING}};
    TextEntity::NGramFileSource* const file = new TextEntity::NGramFileSource(
        flags.str("ngram_file"), train_config.window_size(), &rng, train_config.no_shuffle(),
        WEIGHTING_STRATEGIES.at(flags.str("weighting")), FEATURE_WEIGHTING_STRATEGIES.at(flags.str("feature_weighting")));
    std::printf("n-gram file: %zu instances\n", file->num_instances());
    data_source_ptr.reset(new AsyncSource<TextEntity::Batch>(flags.i("num_concurrent_batches"), train_config.batch_size(),
                                                             train_config.window_size(), file));
  } else {
    data_source_ptr.reset(new TextEntity::SyntheticSource(V, D, flags.i("synthetic_num_batches"), flags.i("seed"),
                                                          flags.d("synthetic_zipf")));
  }
  DataSource<TextEntity::Batch>& data_source = *data_source_ptr;
  // Extract meta data through a generic interface: it sizes the model and is written next to the dumps
  // (reference: cpp/main.cu:501-537; read back by py/nvsm/base.py:load_meta).
  lse::Metadata meta;
  data_source.extract_metadata(&meta);
  V = meta.term_size(); D = meta.object_size();
  NVSM_CHECK(V > 0 && D > 0, "the data source reports an empty vocabulary or corpus");
  std::printf("Training statistics: vocabulary size=%zu, corpus size=%zu\n", V, D);
  if (!flags.str("output").empty()) {
    std::ofstream meta_file(flags.str("output") + "_meta", std::ios::binary);
    NVSM_CHECK(meta.SerializeToOstream(&meta_file), "cannot write the _meta file");
  }
  // construct_data_source<EntityEntity / TermTerm>, cpp/main.cu:240-277: the similarity file resolved through the
  // identifiers map of the text source (here: the decimal ids of the metadata), shuffled with the shared RNG and
  // repeated for as long as the text epoch lasts; has_next of the pair = MultiSource semantics (BatchOps::has_next).
  // Like the reference the sources are built (and first shuffled) BEFORE the model draws its initial parameters.
  // Re-shuffles in the middle of an epoch use the host copy of the engine: with --host_sampler that is the reference's
  // single stream, with the device sampler the negatives keep their own (device-resident) continuation of it.
  std::unique_ptr<DataSource<RepresentationSimilarity::Batch>> similarity_source;
  if (!flags.str("similarity_file").empty() && (!Ops::has_text || train_config.text_entity_weight() < 1.0f)) {
    IdentifiersMapT identifiers_map;
    if (Ops::pairs_over_entities) {
      for (int j = 0; j < meta.object_size(); ++j) identifiers_map[std::to_string(meta.object(j).index_object_id())] = meta.object(j).model_object_id();
    } else {
      for (int i = 0; i < meta.term_size(); ++i) identifiers_map[std::to_string(meta.term(i).index_term_id())] = meta.term(i).model_term_id();
    }
    RepresentationSimilarity::DataSource* const pairs =
        new RepresentationSimilarity::DataSource(flags.str("similarity_file"), identifiers_map, &rng);
    std::printf("similarity file: %zu pairs\n", pairs->num_instances());
    NVSM_CHECK(pairs->num_instances() >= static_cast<size_t>(train_config.batch_size()),
               "the similarity file holds fewer pairs than one batch");
    similarity_source.reset(new RepeatingSource<RepresentationSimilarity::Batch>(static_cast<size_t>(-1), pairs));
  } else {
    similarity_source.reset(new SyntheticSimilaritySource(Ops::pairs_over_entities ? D : V, flags.i("synthetic_num_batches"),
                                                          flags.i("seed") + 17));
  }
  const int gemm_mode = flags.str("gemm") == "fp32" ? NVSM_GEMM_FP32 : (flags.str("gemm") == "tf32" ? NVSM_GEMM_TF32 : NVSM_GEMM_3XTF32);

  std::printf("Model: word_repr_size=%d entity_repr_size=%d batch_normalization=%d nonlinearity=%s\n",
              model_desc.word_repr_size(), model_desc.entity_repr_size(), (int)model_desc.transform_desc().batch_normalization(),
              flags.str("nonlinearity").c_str());
  std::printf("Training: batch_size=%d window_size=%d num_random_entities=%d lambda=%g lr=%g update_method=%s |V|=%zu |D|=%zu\n",
              train_config.batch_size(), train_config.window_size(), train_config.num_random_entities(),
              train_config.regularization_lambda(), train_config.learning_rate(), flags.str("update_method").c_str(), V, D);

  Model<ObjectiveT> model(V, D, model_desc, train_config, flags.i("device"), gemm_mode);
  model.initialize(&rng);
  // negatives: the reference draws them on the host training thread (cpp/labels.cu:3-22, ~3.6 ms per
  // 51200-batch); by default the same stream is produced on the device, --host_sampler restores the loop.
  if (flags.d("negative_sampling_zipf") > 0.0 && Ops::has_text)
    model.set_label_generator(InverseCdfLabelGenerator<float, long>::zipf(D, flags.d("negative_sampling_zipf")));
  // (--check_gradients replays every batch's negatives from the saved host RNG state: host sampler)
  if (!flags.b("host_sampler") && !flags.b("check_gradients") && Ops::has_text) model.use_device_sampler(&rng);
  if (flags.b("dump_initial_model")) dump_model(model, flags.str("output"), "initial");

  Sources sources{&data_source, similarity_source.get()};
  std::unique_ptr<typename Ops::BatchT> batch_ptr(Ops::make(train_config));
  typename Ops::BatchT& batch = *batch_ptr;
  const long max_threads_per_block = 1024;  // Runtime::props().maxThreadsPerBlock in the reference
  const bool verbose = flags.i("v") > 0;
  const bool check_gradients = flags.b("check_gradients");

  auto iterate = [&](const bool backpropagate, size_t* num_batches, double* agg_cost, double* seconds) {
    *num_batches = 0; *agg_cost = 0.0;
    const auto t0 = std::chrono::steady_clock::now();
    std::unique_ptr<typename ObjectiveT::ForwardResultType> previous;
    while (Ops::has_next(sources)) {
      NvtxRange batch_range("Batch");
      {
        NvtxRange fetch_range("FetchData");
        Ops::clear(&batch);
        Ops::next(sources, &batch);
      }
      if (Ops::num_instances(batch) % max_threads_per_block != 0) {
        std::fprintf(stderr, "Skipping Batch #%zu as it is not a multiple of %ld (%zu instances).\n", *num_batches,
                     max_threads_per_block, Ops::num_instances(batch));
      } else {
        std::unique_ptr<typename ObjectiveT::ForwardResultType> result;
        std::unique_ptr<TextEntity::Gradients> gradients;
        std::stringstream rng_state;   // (reference: cpp/main.cu:400-402 -- saved for the gradient check)
        if (check_gradients) rng_state << rng;
        { NvtxRange r("ComputeCost"); result.reset(model.compute_cost(batch, &rng)); }
        { NvtxRange r("ComputeGradients"); gradients.reset(model.compute_gradients(*result)); }
        if (check_gradients) {
          Ops::check_gradients(&model, batch, *result, *gradients, flags.d("gradient_check_epsilon"), rng_state, &rng, verbose);
          // the probes ran in the model's workspace: replay the step's own forward / backward (same negatives)
          std::stringstream copy; copy << rng_state.str(); copy >> rng;
          result.reset(model.compute_cost(batch, &rng));
          gradients.reset(model.compute_gradients(*result));
          (void)result->get_cost();   // read now (cached): the next batch's probes would push it out of the loss ring
        }
        if (backpropagate) {
          NvtxRange r("UpdateParameters");
          model.update(*gradients, train_config.learning_rate(), result->scaled_regularization_lambda());
        }
        // read the previous batch's loss while this one runs (the reference synchronises every batch)
        if (previous) {
          const float c = previous->get_cost();
          *agg_cost += c;
          if (verbose) std::printf("Batch #%zu: cost=%g\n", *num_batches - 1, c);
        }
        previous = std::move(result);
      }
      if (flags.i("dump_every") > 0 && *num_batches > 0 && *num_batches % flags.i("dump_every") == 0)
        dump_model(model, flags.str("output"), std::to_string(*num_batches));
      ++*num_batches;
    }
    if (previous) *agg_cost += previous->get_cost();
    *seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  };

  size_t nb; double cost, secs;
  if (flags.b("compute_initial_cost")) {
    Ops::reset(sources);
    iterate(false, &nb, &cost, &secs);
    std::printf("Initial cost: %g\n", cost / nb);
  }
  size_t total_batches = 0; double total_secs = 0.0;
  for (long epoch = 1; epoch <= train_config.num_epochs(); ++epoch) {
    // the file source shuffles with the shared engine (cpp/data_indri.cpp:386-397): bring its state back from the
    // device sampler first and hand it over again afterwards, so the stream matches the reference's single RNG
    const bool device_rng = (file_source || !flags.str("similarity_file").empty()) && !flags.b("host_sampler") && Ops::has_text;
    if (device_rng) model.sync_rng(&rng);
    Ops::reset(sources);
    if (device_rng) model.use_device_sampler(&rng);
    NvtxRange epoch_range("Epoch");
    iterate(true, &nb, &cost, &secs);
    total_batches += nb; total_secs += secs;
    std::printf("Epoch #%ld: mean cost %g; %.2f batches/second, %.0f n-grams/second\n", epoch, cost / nb,
                total_batches / total_secs, total_batches / total_secs * train_config.batch_size());
    dump_model(model, flags.str("output"), std::to_string(epoch));
  }
  NVSM_ABORT_ON(nvsm_synchronize(model.handle()));
  model.sync_rng(&rng);
  return 0;
}

}  // namespace

int main(int argc, char** argv) {
  Flags flags;
  flags.parse(argc, argv);

  const std::map<std::string, std::pair<lse::TrainConfig::UpdateMethod, lse::TrainConfig::UpdateMethodConf::AdamConf::AdamMode>>
      UPDATE_METHODS = {  // cpp/main.cu:479-485
          {"sgd", {lse::TrainConfig::SGD, lse::TrainConfig::UpdateMethodConf::AdamConf::NONE}},
          {"adagrad", {lse::TrainConfig::ADAGRAD, lse::TrainConfig::UpdateMethodConf::AdamConf::NONE}},
          {"sparse_adam", {lse::TrainConfig::ADAM, lse::TrainConfig::UpdateMethodConf::AdamConf::SPARSE}},
          {"dense_adam", {lse::TrainConfig::ADAM, lse::TrainConfig::UpdateMethodConf::AdamConf::DENSE_UPDATE}},
          {"full_adam", {lse::TrainConfig::ADAM, lse::TrainConfig::UpdateMethodConf::AdamConf::DENSE_UPDATE_DENSE_VARIANCE}}};
  const std::map<std::string, lse::ModelDesc::TransformDesc::Nonlinearity> NONLINEARITIES = {
      {"tanh", lse::ModelDesc::TransformDesc::TANH}, {"hard_tanh", lse::ModelDesc::TransformDesc::HARD_TANH}};

  NVSM_CHECK(UPDATE_METHODS.count(flags.str("update_method")), "Please specify a valid --update_method.");
  NVSM_CHECK(NONLINEARITIES.count(flags.str("nonlinearity")), "Please specify a valid --nonlinearity.");
  // cpp/main.cu:698-706
  NVSM_CHECK(flags.d("entity_similarity_weight") >= 0.0 && flags.d("entity_similarity_weight") <= 1.0, "--entity_similarity_weight must be in [0, 1]");
  NVSM_CHECK(flags.d("term_similarity_weight") >= 0.0 && flags.d("term_similarity_weight") <= 1.0, "--term_similarity_weight must be in [0, 1]");
  // cpp/main.cu:634-638
  NVSM_CHECK(flags.str("weighting") == "auto" || flags.str("weighting") == "uniform" || flags.str("weighting") == "inv_doc_frequency",
             "Please specify a valid --weighting.");
  NVSM_CHECK(flags.str("feature_weighting") == "uniform" || flags.str("feature_weighting") == "self_information",
             "Please specify a valid --feature_weighting.");

  lse::ModelDesc model_desc;
  model_desc.set_word_repr_size(flags.i("word_repr_size"));
  model_desc.set_entity_repr_size(flags.i("entity_repr_size"));
  model_desc.mutable_transform_desc()->set_batch_normalization(flags.b("batch_normalization"));
  model_desc.mutable_transform_desc()->set_nonlinearity(NONLINEARITIES.at(flags.str("nonlinearity")));
  model_desc.set_clip_sigmoid(true);
  model_desc.set_bias_negative_samples(flags.b("bias_negative_samples"));
  model_desc.set_l2_normalize_phrase_reprs(flags.b("l2_phrase_normalization"));
  model_desc.set_l2_normalize_entity_reprs(flags.b("l2_entity_normalization"));

  lse::TrainConfig train_config;
  train_config.set_num_epochs(flags.i("num_epochs"));
  train_config.set_batch_size(flags.i("batch_size"));
  train_config.set_window_size(flags.i("window_size"));
  train_config.set_num_random_entities(flags.i("num_random_entities"));
  train_config.set_regularization_lambda(flags.d("regularization_lambda"));
  train_config.set_learning_rate(flags.d("learning_rate"));
  train_config.mutable_update_method()->set_type(UPDATE_METHODS.at(flags.str("update_method")).first);
  train_config.mutable_update_method()->mutable_adam_conf()->set_mode(UPDATE_METHODS.at(flags.str("update_method")).second);
  train_config.set_no_shuffle(flags.b("no_shuffle"));
  train_config.set_text_entity_weight(1.0 - flags.d("entity_similarity_weight") - flags.d("term_similarity_weight"));
  train_config.set_entity_entity_weight(flags.d("entity_similarity_weight"));
  train_config.set_term_term_weight(flags.d("term_similarity_weight"));

  NVSM_CHECK(flags.i("seed") > 0, "Please specify a --seed value.");
  if (train_config.learning_rate() == 0.0) {
    train_config.set_learning_rate(train_config.update_method().type() == lse::TrainConfig::ADAM ? 0.001 : 0.01);
  }

  // objective selection, cpp/main.cu:729-757
  if (train_config.entity_entity_weight() > 0.0) return train<TextEntityEntityEntity::Objective>(flags, model_desc, train_config);
  if (train_config.term_term_weight() > 0.0) return train<TextEntityTermTerm::Objective>(flags, model_desc, train_config);
  return train<TextEntity::Objective>(flags, model_desc, train_config);
}
