// cuNVSM/cudnn_utils.h — BatchNormalization of the reference (include/cuNVSM/cudnn_utils.h:84-127,
// cpp/cudnn_utils.cu:49-183) without cuDNN: per-activation batch statistics over the instances, biased variance, epsilon
// inside the square root, gamma fixed at 1 (never trained), beta = the bias argument, no running averages
// (nvsm_op_batchnorm_forward / _backward of libnvsm_b200). Matrices are num_features x num_instances, column-major.
#ifndef CUNVSM_B200_CUDNN_UTILS_H
#define CUNVSM_B200_CUDNN_UTILS_H

#include <memory>

#include "device_matrix.h"

template <typename FloatT>
class BatchNormalization {
 public:
  explicit BatchNormalization(const size_t num_features, const FloatT momentum = 0.1, const FloatT epsilon = 1e-4,
                              const bool cache_input = false, Streams* const streams = DefaultStream::get())
      : num_features_(num_features), momentum_(momentum), epsilon_(epsilon), cache_input_(cache_input), streams_(streams),
        mean_cache_(num_features, 1, nullptr, streams), inv_variance_cache_(num_features, 1, nullptr, streams) {}
  virtual ~BatchNormalization() {}
  BatchNormalization(const BatchNormalization&) = delete;
  BatchNormalization& operator=(const BatchNormalization&) = delete;

  // reference: cpp/cudnn_utils.cu:82-129. output may be the input (in place).
  void forward(const device_matrix<FloatT>& input, const device_matrix<FloatT>& bias, device_matrix<FloatT>* const output) {
    NVSM_CHECK(input.getRows() == num_features_ && bias.size() == num_features_, "batch-norm input has the wrong number of features");
    NVSM_CHECK(output->getRows() == input.getRows() && output->getCols() == input.getCols(), "batch-norm output has the wrong shape");
    if (cache_input_) input_cache_.reset(input.copy());   // the in-place call overwrites what backward needs
    NVSM_ABORT_ON(nvsm_op_batchnorm_forward(streams_->ops(), input.getData(), bias.getData(), static_cast<long>(input.getCols()),
                                            static_cast<int>(num_features_), epsilon_, output->getData(), mean_cache_.getData(),
                                            inv_variance_cache_.getData()));
  }

  // reference: cpp/cudnn_utils.cu:131-141 — needs cache_input
  void backward(const device_matrix<FloatT>& grad_output, const device_matrix<FloatT>& bias, device_matrix<FloatT>* const grad_input,
                device_matrix<FloatT>* const grad_bias) {
    NVSM_CHECK(input_cache_ != nullptr, "backward without an input needs cache_input");
    backward(grad_output, *input_cache_, bias, grad_input, grad_bias);
  }

  // reference: cpp/cudnn_utils.cu:143-183. grad_input may be grad_output (in place).
  void backward(const device_matrix<FloatT>& grad_output, const device_matrix<FloatT>& input, const device_matrix<FloatT>& /*bias*/,
                device_matrix<FloatT>* const grad_input, device_matrix<FloatT>* const grad_bias) {
    NVSM_CHECK(grad_output.getRows() == num_features_ && input.size() == grad_output.size(), "batch-norm gradient has the wrong shape");
    NVSM_CHECK(grad_input->size() == grad_output.size() && grad_bias->size() == num_features_, "batch-norm gradient has the wrong shape");
    NVSM_ABORT_ON(nvsm_op_batchnorm_backward(streams_->ops(), grad_output.getData(), input.getData(), mean_cache_.getData(),
                                             inv_variance_cache_.getData(), static_cast<long>(grad_output.getCols()),
                                             static_cast<int>(num_features_), grad_input->getData(), grad_bias->getData()));
  }

  const device_matrix<FloatT>& mean() const { return mean_cache_; }
  const device_matrix<FloatT>& inv_variance() const { return inv_variance_cache_; }

 private:
  const size_t num_features_;
  const FloatT momentum_;   // kept for the signature: the reference passes it to cuDNN but never reads running averages
  const FloatT epsilon_;
  const bool cache_input_;
  Streams* const streams_;
  std::unique_ptr<device_matrix<FloatT>> input_cache_;
  device_matrix<FloatT> mean_cache_;
  device_matrix<FloatT> inv_variance_cache_;
};

#endif  // CUNVSM_B200_CUDNN_UTILS_H
