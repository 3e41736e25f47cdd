// cuNVSM/device_matrix.h — the slice of the un-vendored `device_matrix` library (cvangysel/device_matrix, fetched by the
// reference's third_party/device_matrix-CMakeLists.txt) that the reference's Params / Storage / Updates /
// BatchNormalization class surface and its tests use: a column-major rows x cols matrix in device memory with
// construction, fillwith, copy, and host round trips (to_device / to_host, include/cuNVSM/tests_base_cuda.h).
// Host-only C++ over the stand-alone operator entry points of libnvsm_b200 (nvsm_dev_*): no CUDA toolchain is needed
// to compile against it. Also here: Streams / DefaultStream (include/cuNVSM/cuda_utils.h), which carry the operator
// context instead of a set of cudaStream_t.
//
// Layout: element (r, c) lives at data[c * rows + r]. A dim x objects matrix is therefore the row-major [objects, dim]
// image every nvsm_op_* entry point expects (cpp/storage.cu:6-10).
#ifndef CUNVSM_B200_DEVICE_MATRIX_H
#define CUNVSM_B200_DEVICE_MATRIX_H

#include <cstddef>
#include <cstdio>
#include <initializer_list>
#include <memory>
#include <vector>

#include "../nvsm_b200.h"
#include "base.h"

#ifndef NVSM_ABORT_ON
#define NVSM_ABORT_ON(rc) NVSM_CHECK((rc) == 0, nvsm_last_error())
#endif

typedef void* cudaStream_like;   // the reference passes cudaStream_t; the operator context owns the only stream here

// reference: Streams (include/cuNVSM/cuda_utils.h) — a bag of streams handed to every storage / updater; here the
// launch context of the stand-alone operators.
class Streams {
 public:
  explicit Streams(const int device = 0) { NVSM_ABORT_ON(nvsm_ops_create(device, &ops_)); }
  ~Streams() { nvsm_ops_destroy(ops_); }
  Streams(const Streams&) = delete;
  Streams& operator=(const Streams&) = delete;
  cudaStream_like next() { return nullptr; }
  nvsm_ops* ops() const { return ops_; }
  void synchronize() const { NVSM_ABORT_ON(nvsm_ops_synchronize(ops_)); }

 private:
  nvsm_ops* ops_ = nullptr;
};

// reference: DefaultStream::get() — process-wide default.
class DefaultStream {
 public:
  static Streams* get() {
    static Streams instance(0);
    return &instance;
  }
};

template <typename T>
class device_matrix {
 public:
  device_matrix(const size_t rows, const size_t cols, cudaStream_like /*stream*/ = nullptr, Streams* const streams = DefaultStream::get())
      : rows_(rows), cols_(cols), streams_(streams) {
    NVSM_CHECK(rows > 0 && cols > 0, "empty device_matrix");
    void* p = nullptr;
    NVSM_ABORT_ON(nvsm_dev_malloc(streams_->ops(), &p, rows * cols * sizeof(T)));   // zero-initialised
    data_ = static_cast<T*>(p);
  }
  ~device_matrix() { nvsm_dev_free(streams_->ops(), data_); }
  device_matrix(const device_matrix&) = delete;
  device_matrix& operator=(const device_matrix&) = delete;

  size_t getRows() const { return rows_; }
  size_t getCols() const { return cols_; }
  size_t size() const { return rows_ * cols_; }
  T* getData() { return data_; }
  const T* getData() const { return data_; }
  cudaStream_like getStream() const { return nullptr; }
  Streams* streams() const { return streams_; }

  void fillwith(cudaStream_like, const T value) { fill_impl(value); }
  void fillwith(cudaStream_like, const std::vector<T>& values) {
    NVSM_CHECK(values.size() == size(), "fillwith: size mismatch");
    NVSM_ABORT_ON(nvsm_dev_upload(streams_->ops(), data_, values.data(), size() * sizeof(T)));
  }
  device_matrix<T>* copy(cudaStream_like = nullptr) const {
    device_matrix<T>* const out = new device_matrix<T>(rows_, cols_, nullptr, streams_);
    NVSM_ABORT_ON(nvsm_dev_copy(streams_->ops(), out->data_, data_, size() * sizeof(T)));
    return out;
  }
  void copyFrom(cudaStream_like, const device_matrix<T>& other) {
    NVSM_CHECK(other.size() == size(), "copyFrom: size mismatch");
    NVSM_ABORT_ON(nvsm_dev_copy(streams_->ops(), data_, other.data_, size() * sizeof(T)));
  }
  std::vector<T> to_host() const {
    std::vector<T> out(size());
    NVSM_ABORT_ON(nvsm_dev_download(streams_->ops(), out.data(), data_, size() * sizeof(T)));
    return out;
  }

 private:
  void fill_impl(const float value) {
    NVSM_ABORT_ON(nvsm_dev_fill(streams_->ops(), reinterpret_cast<float*>(data_), static_cast<long>(size()), value));
  }
  template <typename U>
  void fill_impl(const U value) {   // index matrices: no device fill kernel for integers
    fillwith(nullptr, std::vector<T>(size(), static_cast<T>(value)));
  }
  const size_t rows_, cols_;
  Streams* const streams_;
  T* data_ = nullptr;
};

// reference: to_device / to_host / print_matrix (include/cuNVSM/tests_base_cuda.h) — linear memory order.
template <typename T>
void to_device(const std::vector<T>& values, device_matrix<T>* const dst) { dst->fillwith(nullptr, values); }
template <typename T>
void to_device(std::initializer_list<T> values, device_matrix<T>* const dst) { dst->fillwith(nullptr, std::vector<T>(values)); }
template <typename T>
std::vector<T> to_host(const device_matrix<T>& src) { return src.to_host(); }
template <typename T>
void print_matrix(const device_matrix<T>& m, std::FILE* const out = stderr) {
  const std::vector<T> h = m.to_host();
  for (size_t r = 0; r < m.getRows(); ++r) {
    for (size_t c = 0; c < m.getCols(); ++c) std::fprintf(out, "%s%g", c ? " " : "", static_cast<double>(h[c * m.getRows() + r]));
    std::fprintf(out, "\n");
  }
}

#endif  // CUNVSM_B200_DEVICE_MATRIX_H
