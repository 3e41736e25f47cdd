// cuNVSM/gradient_check.h — GradientCheckFn of the reference (include/cuNVSM/gradient_check.h, cpp/gradient_check.cu:3-133;
// CLI flag --check_gradients, cpp/main.cu:63,414-420): for EVERY scalar parameter of the model, the analytical gradient of
// the running ForwardResult / Gradients against a central difference of the cost, with the negatives re-drawn from the
// saved RNG state so that every probe sees the same sampled batch. A debugging aid for tiny models: 3 forward passes per
// parameter. The probes run forward passes in the model's single workspace, so the ForwardResult / Gradients handed in
// are stale afterwards: the caller restores the RNG state and calls compute_cost / compute_gradients again (cpp/main.cpp).
//
// Same verdict rules as the reference: a gradient pointing the wrong way fails; a relative error above the threshold
// fails unless the numerical gradient is exactly zero; the cost of the unperturbed model must be reproduced (1e-5) after
// every probe. Differences from the reference, both because this library computes in float32 where the reference's
// check runs in its float64 test build: the cost is read before its final rounding to float (nvsm_read_cost_f64), and
// `floor` (default 2e-6) is the absolute noise of a central difference over a float32 forward pass (cost noise ~1e-8 over
// 2 epsilon): parameters whose two gradients are both below it are not judged, and a disagreement smaller than it is not
// counted as one.
#ifndef CUNVSM_B200_GRADIENT_CHECK_H
#define CUNVSM_B200_GRADIENT_CHECK_H

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <sstream>
#include <string>
#include <vector>

#include "model.h"

template <typename ModelT>
class GradientCheckFn {
 public:
  typedef typename ModelT::FloatT FloatT;

  struct Report {
    size_t checked = 0, skipped = 0, wrong_direction = 0, above_threshold = 0;
    double worst_relative_error = 0.0;
    std::string worst;
  };
  const Report& report() const { return report_; }

  bool operator()(ModelT* const model, const typename ModelT::Batch& batch, const typename ModelT::ForwardResult& result,
                  const typename ModelT::Gradients& gradients, const FloatT epsilon, const FloatT relative_error_threshold,
                  const std::stringstream& rng_state, RNG* const rng, const double floor = 2e-6, const int verbose = 0) {
    NVSM_CHECK(model->initialized(), "gradient check on an uninitialised model");
    NVSM_CHECK(epsilon >= 0.0 && relative_error_threshold >= 0.0, "negative epsilon / threshold");
    NVSM_CHECK(ModelT::Objective::kObjective == NVSM_OBJECTIVE_TEXT_ENTITY, "the gradient check covers the TextEntity objective");
    NVSM_CHECK(!model->uses_device_sampler(), "the gradient check replays the negatives from the host RNG state: disable the device sampler");
    nvsm_model* const h = model->handle();
    const size_t B = batch.num_instances(), n = batch.window_size();
    const lse::ModelDesc& desc = model->desc();
    const size_t dw = desc.word_repr_size(), dd = desc.entity_repr_size();
    report_ = Report();

    // analytical gradients of the running step, dense-ified on the host (RepresentationsStorage::get_parameter_gradient,
    // cpp/storage.cu:139-183: the weighted sum of the gradient columns that land on a parameter)
    const std::vector<FloatT> gT = gradients.get("grad_transform"), gb = gradients.get("grad_bias");
    const std::vector<FloatT> gP = gradients.get("grad_phrase_reprs"), gE = gradients.get("grad_entity_repr");
    const size_t R = gE.size() / (B * dd);
    std::vector<long> ids(B * R);
    NVSM_ABORT_ON(nvsm_get_entity_ids(h, ids.data(), static_cast<long>(ids.size())));
    const long V = nvsm_tensor_size(h, "word_representations-representations") / static_cast<long>(dw);
    const long D = nvsm_tensor_size(h, "entity_representations-representations") / static_cast<long>(dd);
    std::vector<double> dW(static_cast<size_t>(V) * dw, 0.0), dE(static_cast<size_t>(D) * dd, 0.0);
    for (size_t i = 0; i < B; ++i)
      for (size_t w = 0; w < n; ++w) {
        const long id = batch.features()[i * n + w];
        const double wt = batch.feature_weights()[i * n + w];
        for (size_t k = 0; k < dw; ++k) dW[id * dw + k] += wt * gP[i * dw + k];
      }
    for (size_t c = 0; c < B * R; ++c)
      for (size_t k = 0; k < dd; ++k) dE[ids[c] * dd + k] += gE[c * dd + k];

    const double base_cost = cost_of(model, batch, rng_state, rng);
    {
      double now = 0.0;
      NVSM_ABORT_ON(nvsm_read_cost_f64(h, 0, &now));
      (void)result;
      NVSM_CHECK(std::fabs(now - base_cost) <= 1e-5, "the saved RNG state does not reproduce the forward result");
    }

    bool checked = true;
    // the reference walks model->params_ in ParamIdentifier order: word representations, transform (+ bias), entities
    struct Group { const char* tensor; const char* name; size_t count; const double* dense; const std::vector<FloatT>* grad; };
    const Group groups[4] = {{"word_representations-representations", "word_representations", dW.size(), dW.data(), nullptr},
                             {"word_entity_mapping-transform", "word_entity_mapping", gT.size(), nullptr, &gT},
                             {"word_entity_mapping-bias", "word_entity_mapping(bias)", gb.size(), nullptr, &gb},
                             {"entity_representations-representations", "entity_representations", dE.size(), dE.data(), nullptr}};
    for (const Group& g : groups)
      for (size_t idx = 0; idx < g.count; ++idx) {
        // gradients are ascent directions of -cost (cpp/objective.cu:324-326): negate, like the reference
        const double predict = -(g.dense ? g.dense[idx] : static_cast<double>((*g.grad)[idx]));
        NVSM_ABORT_ON(nvsm_increment_parameter(h, g.tensor, static_cast<long>(idx), epsilon));
        const double plus = cost_of(model, batch, rng_state, rng);
        NVSM_ABORT_ON(nvsm_increment_parameter(h, g.tensor, static_cast<long>(idx), -2.0f * epsilon));
        const double minus = cost_of(model, batch, rng_state, rng);
        NVSM_ABORT_ON(nvsm_increment_parameter(h, g.tensor, static_cast<long>(idx), epsilon));
        const double approx = (plus - minus) / (2.0 * epsilon);
        const double scale = std::max(std::fabs(predict), std::fabs(approx));
        const double error = std::fabs(predict - approx);
        if (scale < floor) { ++report_.skipped; continue; }
        ++report_.checked;
        const double rel = error / scale;
        if (error > floor && rel > report_.worst_relative_error) {
          report_.worst_relative_error = rel;
          report_.worst = std::string(g.name) + "[" + std::to_string(idx) + "]";
        }
        // `floor` is the absolute noise of a central difference over a float32 forward pass: differences below it are
        // not evidence either way
        if (predict * approx < 0.0 && error > floor) {
          if (report_.wrong_direction + report_.above_threshold < 10 || verbose > 0)
            std::fprintf(stderr, "Parameter %zu of %s has gradient with incorrect direction (approx=%g, predict=%g, relative error=%g).\n",
                         idx, g.name, approx, predict, rel);
          ++report_.wrong_direction;
          checked = false;
        } else if (rel >= relative_error_threshold && error > floor) {
          if (report_.wrong_direction + report_.above_threshold < 10 || verbose > 0)
            std::fprintf(stderr, "Parameter %zu of %s most likely has incorrect gradient (approx=%g, predict=%g, relative error=%g).\n",
                         idx, g.name, approx, predict, rel);
          if (approx != 0.0) { ++report_.above_threshold; checked = false; }
        }
      }
    NVSM_CHECK(std::fabs(cost_of(model, batch, rng_state, rng) - base_cost) <= 1e-5, "the parameters were not restored");
    return checked;
  }

 private:
  // Model::get_cost (cpp/model.cu:154-174) with the cost read in double
  static double cost_of(ModelT* const model, const typename ModelT::Batch& batch, const std::stringstream& rng_state, RNG* const rng) {
    std::stringstream copy;
    copy << rng_state.str();
    copy >> *rng;
    std::unique_ptr<typename ModelT::ForwardResult> r(model->compute_cost(batch, rng));
    double c = 0.0;
    NVSM_ABORT_ON(nvsm_read_cost_f64(model->handle(), 0, &c));
    return c;
  }
  Report report_;
};

#endif  // CUNVSM_B200_GRADIENT_CHECK_H
