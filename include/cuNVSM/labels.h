// cuNVSM/labels.h — the negative-sampling plug point of the reference (reference: include/cuNVSM/labels.h:7-29,
// cpp/labels.cu:3-22, used by TextEntity::Objective::generate_labels, cpp/objective.cu:5-28).
//
// LabelGenerator::generate fills instance_entities[i * (z+1)] = labels[i] followed by z sampled negatives. The
// reference passes the entity Representations only to ask it for num_objects(); here the count is passed directly.
//   UniformLabelGenerator    — the reference's only generator (bit-exact: std::uniform_int_distribution<long> per draw
//                              on the shared std::minstd_rand0), host loop or device sampler.
//   InverseCdfLabelGenerator — skewed negatives (BASELINE configs[4], Zipf): one engine output per draw,
//                              u = (x - 1) / 2147483646, id = min{k : cdf[k] > u}; host loop or device sampler, same ids.
// Any other subclass runs on the host (Model falls back to the host path for it).
#ifndef CUNVSM_B200_LABELS_H
#define CUNVSM_B200_LABELS_H

#include <cmath>
#include <vector>

#include "../nvsm_b200.h"
#include "base.h"

template <typename FloatT, typename EntityIdxType>
class LabelGenerator {
 public:
  virtual ~LabelGenerator() {}

  // instance_entities[i * (z + 1)] = labels[i], then z negatives out of [0, num_objects); consumes *rng in draw order
  virtual void generate(const EntityIdxType* const labels, const size_t num_objects, const size_t num_labels,
                        const size_t num_negative_labels, std::vector<EntityIdxType>* const instance_entities,
                        RNG* const rng) const = 0;

  // Device sampler support: true when nvsm_step_sampled reproduces generate() bit for bit, with the cumulative
  // distribution to install (nullptr = uniform).
  virtual bool on_device() const { return false; }
  virtual const std::vector<double>* distribution() const { return nullptr; }
};

template <typename FloatT, typename EntityIdxType>
class UniformLabelGenerator : public LabelGenerator<FloatT, EntityIdxType> {
 public:
  virtual void generate(const EntityIdxType* const labels, const size_t num_objects, const size_t num_labels,
                        const size_t num_negative_labels, std::vector<EntityIdxType>* const instance_entities,
                        RNG* const rng) const override {
    instance_entities->resize(num_labels * (num_negative_labels + 1));
    unsigned long state = nvsm_detail::rng_get_state(*rng);
    NVSM_CHECK(nvsm_generate_labels(labels, num_labels, num_negative_labels, num_objects, &state,
                                    instance_entities->data()) == 0, nvsm_last_error());
    nvsm_detail::rng_set_state(rng, state);
  }
  virtual bool on_device() const override { return true; }
};

template <typename FloatT, typename EntityIdxType>
class InverseCdfLabelGenerator : public LabelGenerator<FloatT, EntityIdxType> {
 public:
  // cdf[k] = P(id <= k): non-decreasing, last entry exactly 1.0
  explicit InverseCdfLabelGenerator(const std::vector<double>& cdf) : cdf_(cdf) {}

  // Zipf(s) over ids 0..num_objects-1: id k has weight (k+1)^-s
  static InverseCdfLabelGenerator* zipf(const size_t num_objects, const double exponent) {
    std::vector<double> cdf(num_objects);
    double acc = 0.0;
    for (size_t k = 0; k < num_objects; ++k) { acc += 1.0 / std::pow(static_cast<double>(k + 1), exponent); cdf[k] = acc; }
    for (size_t k = 0; k < num_objects; ++k) cdf[k] /= acc;
    cdf[num_objects - 1] = 1.0;
    return new InverseCdfLabelGenerator(cdf);
  }

  virtual void generate(const EntityIdxType* const labels, const size_t num_objects, const size_t num_labels,
                        const size_t num_negative_labels, std::vector<EntityIdxType>* const instance_entities,
                        RNG* const rng) const override {
    NVSM_CHECK(num_objects == cdf_.size(), "the distribution does not cover the entity table");
    instance_entities->resize(num_labels * (num_negative_labels + 1));
    unsigned long state = nvsm_detail::rng_get_state(*rng);
    NVSM_CHECK(nvsm_generate_labels_cdf(labels, num_labels, num_negative_labels, cdf_.data(), cdf_.size(), &state,
                                        instance_entities->data()) == 0, nvsm_last_error());
    nvsm_detail::rng_set_state(rng, state);
  }
  virtual bool on_device() const override { return true; }
  virtual const std::vector<double>* distribution() const override { return &cdf_; }

 private:
  const std::vector<double> cdf_;
};

#endif  // CUNVSM_B200_LABELS_H
