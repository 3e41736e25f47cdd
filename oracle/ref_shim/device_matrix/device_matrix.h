// Reconstruction of the part of cvangysel/device_matrix (un-vendored dependency of the reference,
// `third_party/device_matrix-CMakeLists.txt:6-8`, GIT_TAG master, source NOT under /root/reference)
// that the reference's training step calls. Written from the reference's call sites and tests
// (SURVEY.md Appendix A); one straightforward Thrust / CUDA kernel or cuBLAS call per operation,
// column-major storage, a size-bucketed caching pool standing in for cnmem.
//
// TEST INFRASTRUCTURE: exists only so the UNMODIFIED reference sources compile into
// oracle/_ref/libcunvsm_ref_{f32,f64}.so (oracle/ref_shim/Makefile). Nothing under cunvsm_b200/
// includes or links it. Numbers measured through it are "reference kernels + call structure over a
// reconstructed device_matrix", not the original library.
#ifndef REF_SHIM_DEVICE_MATRIX_H
#define REF_SHIM_DEVICE_MATRIX_H

#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cmath>
#include <initializer_list>
#include <iostream>
#include <limits>
#include <map>
#include <mutex>
#include <memory>
#include <type_traits>
#include <utility>
#include <vector>

#include <thrust/copy.h>
#include <thrust/count.h>
#include <thrust/device_ptr.h>
#include <thrust/execution_policy.h>
#include <thrust/fill.h>
#include <thrust/for_each.h>
#include <thrust/functional.h>
#include <thrust/host_vector.h>
#include <thrust/iterator/counting_iterator.h>
#include <thrust/iterator/permutation_iterator.h>
#include <thrust/iterator/transform_iterator.h>
#include <thrust/iterator/zip_iterator.h>
#include <thrust/reduce.h>
#include <thrust/transform.h>
#include <thrust/transform_reduce.h>
#include <thrust/tuple.h>

#include <glog/logging.h>

namespace cuda {

inline void check_cuda(const cudaError_t e, const char* const file, const int line) {
    if (e != cudaSuccess) {
        LOG(FATAL) << "CUDA error " << cudaGetErrorString(e) << " at " << file << ":" << line;
    }
}

inline void check_cublas(const cublasStatus_t s, const char* const file, const int line) {
    if (s != CUBLAS_STATUS_SUCCESS) {
        LOG(FATAL) << "cuBLAS error " << static_cast<int>(s) << " at " << file << ":" << line;
    }
}

}  // namespace cuda

#define CCE(val) ::cuda::check_cuda((val), __FILE__, __LINE__)
#define CCBE(val) ::cuda::check_cublas((val), __FILE__, __LINE__)

#define MAX_THREADS_PER_BLOCK 1024

#define PROFILE_FUNCTION() do {} while (0)
#define PROFILE_FUNCTION_WITH_STREAM(stream) do {} while (0)

#define LAUNCH_KERNEL(...) do { __VA_ARGS__; CCE(cudaGetLastError()); } while (0)

#define CHECK_DIMENSIONS(m, rows, cols) \
  do { CHECK_EQ((m).getRows(), (rows)); CHECK_EQ((m).getCols(), (cols)); } while (0)
#define CHECK_DIMENSIONS_EQUAL(a, b) \
  do { CHECK_EQ((a).getRows(), (b).getRows()); CHECK_EQ((a).getCols(), (b).getCols()); } while (0)

#ifdef NDEBUG
// Release: the finiteness / norm scans and the "null" fill are debugging aids, compiled out.
#define CHECK_MATRIX(m) do {} while (0)
#define CHECK_MATRIX_FINITE(m) do {} while (0)
#define CHECK_MATRIX_NORM(m) do {} while (0)
#define MAKE_MATRIX_NULL(m) do {} while (0)
#else
#define CHECK_MATRIX(m) CHECK(::cuda::is_finite(m)) << "non-finite matrix " #m
#define CHECK_MATRIX_FINITE(m) CHECK(::cuda::is_finite(m)) << "non-finite matrix " #m
#define CHECK_MATRIX_NORM(m) do {} while (0)
// Poison, so that an operation which wrongly depended on the previous contents shows up as NaN.
#define MAKE_MATRIX_NULL(m) ::cuda::poison(&(m))
#endif

namespace cuda {

enum Axis { FIRST_AXIS = 0, SECOND_AXIS = 1 };

//
// Functors.
//

namespace func {

template <typename FloatT>
struct identity {
  typedef FloatT argument_type;
  typedef FloatT result_type;
  __host__ __device__ FloatT operator()(const FloatT x) const { return x; }
};

template <typename FloatT>
struct square {
  typedef FloatT argument_type;
  typedef FloatT result_type;
  __host__ __device__ FloatT operator()(const FloatT x) const { return x * x; }
};

template <typename FloatT>
struct scale_by_constant {
  typedef FloatT argument_type;
  typedef FloatT result_type;
  explicit scale_by_constant(const FloatT c) : c_(c) {}
  __host__ __device__ FloatT operator()(const FloatT x) const { return x * c_; }
  FloatT c_;
};

template <typename FloatT>
struct add_constant {
  typedef FloatT argument_type;
  typedef FloatT result_type;
  explicit add_constant(const FloatT c) : c_(c) {}
  __host__ __device__ FloatT operator()(const FloatT x) const { return x + c_; }
  FloatT c_;
};

template <typename FloatT>
struct power {
  typedef FloatT argument_type;
  typedef FloatT result_type;
  explicit power(const FloatT p) : p_(p) {}
  __host__ __device__ FloatT operator()(const FloatT x) const { return ::pow(x, p_); }
  FloatT p_;
};

template <typename FloatT>
struct divides_tuple {
  typedef FloatT result_type;
  template <typename Tuple>
  __host__ __device__ FloatT operator()(const Tuple& t) const {
      return thrust::get<0>(t) / thrust::get<1>(t);
  }
};

template <typename FloatT>
struct is_not_finite {
  __host__ __device__ bool operator()(const FloatT x) const { return !isfinite(static_cast<double>(x)); }
};

}  // namespace func

//
// Streams. The reference runs everything through `DefaultStream` ("Multiple streams do not seem to
// improve training speed", cpp/model.cu:13-14): one stream, merge_streams is the identity.
//

class Streams {
 public:
  virtual ~Streams() {}
  virtual cudaStream_t next() = 0;
  virtual void synchronize() { CCE(cudaStreamSynchronize(0)); }
};

class DefaultStream : public Streams {
 public:
  DefaultStream() {}
  virtual cudaStream_t next() { return 0; }

  static DefaultStream* get() {
      static DefaultStream instance;
      return &instance;
  }
};

inline cudaStream_t merge_streams(const cudaStream_t first, const cudaStream_t second) {
    if (first == second) {
        return first;
    }
    // Different streams: make `first` wait for everything queued on `second`.
    cudaEvent_t event;
    CCE(cudaEventCreateWithFlags(&event, cudaEventDisableTiming));
    CCE(cudaEventRecord(event, second));
    CCE(cudaStreamWaitEvent(first, event, 0));
    CCE(cudaEventDestroy(event));
    return first;
}

class ScopedProfiler {
 public:
  explicit ScopedProfiler(const char* const) {}
};

// Device memory pool standing in for cnmem (the allocator device_matrix links, CMakeLists.txt:49,78 of the
// reference): blocks are cached by size and handed back without touching the driver, so the steady-state step
// performs no cudaMalloc / cudaFree. All work is ordered on one stream, which makes immediate reuse safe.
class MemoryPool {
 public:
  static MemoryPool* getInstance() {
      static MemoryPool* const instance = new MemoryPool;  // leaked on purpose: outlives every static matrix.
      return instance;
  }

  void* allocate(const size_t bytes) {
      const size_t rounded = ((bytes > 0 ? bytes : 1) + 511) / 512 * 512;
      {
          ::std::lock_guard< ::std::mutex> guard(mutex_);
          ::std::vector<void*>& bucket = free_[rounded];
          if (!bucket.empty()) {
              void* const ptr = bucket.back();
              bucket.pop_back();
              return ptr;
          }
      }
      void* ptr = nullptr;
      cudaError_t status = cudaMalloc(&ptr, rounded);
      if (status != cudaSuccess) {
          cudaGetLastError();
          release_cached();
          status = cudaMalloc(&ptr, rounded);
      }
      CCE(status);
      return ptr;
  }

  void deallocate(void* const ptr, const size_t bytes) {
      const size_t rounded = ((bytes > 0 ? bytes : 1) + 511) / 512 * 512;
      ::std::lock_guard< ::std::mutex> guard(mutex_);
      free_[rounded].push_back(ptr);
  }

 private:
  MemoryPool() {}

  void release_cached() {
      ::std::lock_guard< ::std::mutex> guard(mutex_);
      cudaDeviceSynchronize();
      for (::std::map<size_t, ::std::vector<void*> >::iterator it = free_.begin(); it != free_.end(); ++it) {
          for (size_t i = 0; i < it->second.size(); ++i) cudaFree(it->second[i]);
          it->second.clear();
      }
  }

  ::std::mutex mutex_;
  ::std::map<size_t, ::std::vector<void*> > free_;
};

template <typename FloatT>
class Runtime {
 public:
  static Runtime* getInstance() {
      static Runtime instance;
      return &instance;
  }

  const cudaDeviceProp& props() const { return props_; }
  cublasHandle_t& handle() { return handle_; }

 private:
  Runtime() {
      int device = 0;
      CCE(cudaGetDevice(&device));
      CCE(cudaGetDeviceProperties(&props_, device));
      CCBE(cublasCreate(&handle_));
  }

  cudaDeviceProp props_;
  cublasHandle_t handle_;
};

template <typename T>
inline T get_scalar(const T* const device_ptr) {
    T value;
    CCE(cudaMemcpy(&value, device_ptr, sizeof(T), cudaMemcpyDeviceToHost));
    return value;
}

//
// device_matrix: column-major rows x cols.
//

template <typename FloatT>
class device_matrix {
 public:
  typedef FloatT value_type;

  device_matrix(const size_t rows, const size_t cols, const cudaStream_t stream)
      : rows_(rows), cols_(cols), stream_(stream), data_(nullptr) {
      data_ = static_cast<FloatT*>(MemoryPool::getInstance()->allocate(size() * sizeof(FloatT)));
  }

  ~device_matrix() {
      if (data_ != nullptr) {
          MemoryPool::getInstance()->deallocate(data_, size() * sizeof(FloatT));
      }
  }

  // From host memory [begin, end).
  static device_matrix* create(const cudaStream_t stream,
                               const FloatT* const begin, const FloatT* const end,
                               const size_t rows, const size_t cols) {
      CHECK_EQ(static_cast<size_t>(end - begin), rows * cols);
      device_matrix* const m = new device_matrix(rows, cols, stream);
      CCE(cudaMemcpyAsync(m->data_, begin, m->size() * sizeof(FloatT),
                          cudaMemcpyHostToDevice, stream));
      return m;
  }

  static device_matrix* create(const cudaStream_t stream,
                               const ::std::vector<FloatT>& values,
                               const size_t rows, const size_t cols) {
      CHECK_EQ(values.size(), rows * cols);
      device_matrix* const m = new device_matrix(rows, cols, stream);
      m->fillwith(stream, values);
      return m;
  }

  static device_matrix* create(const cudaStream_t stream,
                               ::std::initializer_list<FloatT> values,
                               const size_t rows, const size_t cols) {
      return create(stream, ::std::vector<FloatT>(values), rows, cols);
  }

  static device_matrix* create_column(const cudaStream_t stream,
                                      const ::std::vector<FloatT>& values) {
      return create(stream, values, values.size(), 1);
  }

  static device_matrix* create_shape_as(const cudaStream_t stream, const device_matrix& other) {
      return new device_matrix(other.getRows(), other.getCols(), stream);
  }

  inline size_t getRows() const { return rows_; }
  inline size_t getCols() const { return cols_; }
  inline size_t size() const { return rows_ * cols_; }
  inline FloatT* getData() const { return data_; }
  inline cudaStream_t getStream() const { return stream_; }

  inline thrust::device_ptr<FloatT> begin() const { return thrust::device_pointer_cast(data_); }
  inline thrust::device_ptr<FloatT> end() const { return thrust::device_pointer_cast(data_ + size()); }
  inline thrust::device_ptr<FloatT> begin(const size_t col) const {
      return thrust::device_pointer_cast(data_ + col * rows_);
  }

  inline bool hasSameShape(const device_matrix& other) const {
      return rows_ == other.rows_ && cols_ == other.cols_;
  }

  void fillwith(const cudaStream_t stream, const FloatT value) {
      thrust::fill(thrust::cuda::par.on(stream), begin(), end(), value);
  }

  void fillwith(const cudaStream_t stream, const ::std::vector<FloatT>& values) {
      CHECK_EQ(values.size(), size());
      // Pageable source: cudaMemcpyAsync stages it before returning, the vector may die afterwards.
      CCE(cudaMemcpyAsync(data_, values.data(), size() * sizeof(FloatT),
                          cudaMemcpyHostToDevice, stream));
  }

  device_matrix* copy(const cudaStream_t stream) const {
      device_matrix* const m = new device_matrix(rows_, cols_, stream);
      m->copyFrom(stream, *this);
      return m;
  }

  void copyFrom(const cudaStream_t stream, const device_matrix& other) {
      CHECK(hasSameShape(other));
      CCE(cudaMemcpyAsync(data_, other.data_, size() * sizeof(FloatT),
                          cudaMemcpyDeviceToDevice, stream));
  }

  void scale(const cudaStream_t stream, const FloatT alpha) {
      thrust::transform(thrust::cuda::par.on(stream), begin(), end(), begin(),
                        func::scale_by_constant<FloatT>(alpha));
  }

  void square(const cudaStream_t stream) {
      thrust::transform(thrust::cuda::par.on(stream), begin(), end(), begin(),
                        func::square<FloatT>());
  }

  void transfer(const cudaStream_t stream, FloatT* const host_dst, const size_t num) const {
      CHECK_LE(num, size());
      CCE(cudaMemcpyAsync(host_dst, data_, num * sizeof(FloatT), cudaMemcpyDeviceToHost, stream));
  }

 private:
  const size_t rows_;
  const size_t cols_;
  const cudaStream_t stream_;
  FloatT* data_;

  device_matrix(const device_matrix&);
  void operator=(const device_matrix&);
};

template <typename FloatT>
inline thrust::device_ptr<FloatT> begin(const device_matrix<FloatT>& m) { return m.begin(); }
template <typename FloatT>
inline thrust::device_ptr<FloatT> end(const device_matrix<FloatT>& m) { return m.end(); }
template <typename FloatT>
inline FloatT* raw_begin(const device_matrix<FloatT>& m) { return m.getData(); }

template <typename FloatT>
FloatT* get_array(const cudaStream_t stream, const device_matrix<FloatT>& m) {
    FloatT* const host = new FloatT[m.size()];
    CCE(cudaMemcpyAsync(host, m.getData(), m.size() * sizeof(FloatT), cudaMemcpyDeviceToHost, stream));
    CCE(cudaStreamSynchronize(stream));
    return host;
}

template <typename FloatT>
void print_matrix(const device_matrix<FloatT>& m, ::std::ostream& os = ::std::cerr) {
#ifndef NDEBUG
    (void) m; (void) os;  // VLOG-level output in the original; silent here.
#endif
}

template <typename FloatT>
::std::ostream& operator<<(::std::ostream& os, const device_matrix<FloatT>& m) {
    os << "device_matrix(" << m.getRows() << " x " << m.getCols() << ")";
    return os;
}

template <typename FloatT>
bool is_finite(const device_matrix<FloatT>& m) {
    return thrust::count_if(thrust::cuda::par.on(m.getStream()), m.begin(), m.end(),
                            func::is_not_finite<FloatT>()) == 0;
}

template <typename FloatT>
void poison(device_matrix<FloatT>* const m) {
    m->fillwith(m->getStream(), ::std::numeric_limits<FloatT>::quiet_NaN());
}
template <typename FloatT>
void poison(device_matrix<FloatT>* const* const) {}

template <typename T>
void flatten(const cudaStream_t stream,
             const ::std::vector< ::std::vector<T> >& nested,
             device_matrix<T>* const dst) {
    ::std::vector<T> flat;
    for (size_t i = 0; i < nested.size(); ++i) {
        flat.insert(flat.end(), nested[i].begin(), nested[i].end());
    }
    dst->fillwith(stream, flat);
}

//
// Iterators.
//

template <typename Iterator, typename FloatT>
inline thrust::transform_iterator<func::scale_by_constant<FloatT>, Iterator>
make_scalar_multiplication_iterator(Iterator it, const FloatT c) {
    return thrust::make_transform_iterator(it, func::scale_by_constant<FloatT>(c));
}

// Thrust deduces FloatT from the scalar; the reference passes doubles to float iterators
// (`1.0 - lambda * lr`), so narrow to the iterator's value type.
template <typename FloatT, typename ScalarT>
inline thrust::transform_iterator<func::scale_by_constant<FloatT>, thrust::device_ptr<FloatT> >
make_scalar_multiplication_iterator(thrust::device_ptr<FloatT> it, const ScalarT c) {
    return thrust::make_transform_iterator(it, func::scale_by_constant<FloatT>(static_cast<FloatT>(c)));
}

struct index_to_column {
  typedef size_t argument_type;
  typedef size_t result_type;
  explicit index_to_column(const size_t rows) : rows_(rows) {}
  __host__ __device__ size_t operator()(const size_t idx) const { return idx / rows_; }
  size_t rows_;
};

template <typename FloatT>
inline thrust::transform_iterator<index_to_column, thrust::counting_iterator<size_t> >
make_matrix_column_iterator(const device_matrix<FloatT>& m) {
    return thrust::make_transform_iterator(thrust::counting_iterator<size_t>(0),
                                           index_to_column(m.getRows()));
}

//
// Element-wise operations.
//

namespace detail {

inline unsigned int num_blocks(const size_t n, const unsigned int threads) {
    return static_cast<unsigned int>((n + threads - 1) / threads);
}

template <typename Policy>
inline cudaStream_t stream_of(const Policy& policy) {
    return thrust::cuda_cub::stream(
        const_cast<typename ::std::remove_const<Policy>::type&>(policy));
}

template <typename FloatT, typename Op>
__global__ void columns_unary_kernel(const size_t rows, const size_t n, const size_t N,
                                     const bool every_nth, FloatT* const data, Op op) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const size_t col = i / rows;
    const bool is_nth = (col % N) == 0;
    if (is_nth == every_nth) {
        data[i] = op(data[i]);
    }
}

template <typename FloatT, typename Op, typename FirstOp, typename SecondOp>
__global__ void columnwise_kernel(const size_t rows, const size_t n,
                                  const FloatT* const first, const FloatT* const vec,
                                  FloatT* const dst, Op op, FirstOp first_op, SecondOp second_op) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    dst[i] = op(first_op(first[i]), second_op(vec[i / rows]));
}

template <typename FloatT>
__global__ void broadcast_columns_kernel(const size_t rows, const size_t n, const size_t reps,
                                         const FloatT* const src, FloatT* const dst) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const size_t col = i / rows, row = i % rows;
    dst[i] = src[(col / reps) * rows + row];
}

template <typename FloatT>
__global__ void repmat_kernel(const size_t src_size, const size_t n,
                              const FloatT* const src, FloatT* const dst) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    dst[i] = src[i % src_size];
}

template <typename FloatT, typename Op>
__global__ void fold_columns_kernel(const size_t rows, const size_t n, const size_t k,
                                    const FloatT* const src, const FloatT* const weights,
                                    FloatT* const dst, Op op) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    const size_t col = i / rows, row = i % rows;
    const size_t first = col * k;
    FloatT agg = src[first * rows + row] * (weights != nullptr ? weights[first] : static_cast<FloatT>(1));
    for (size_t r = 1; r < k; ++r) {
        const FloatT value = src[(first + r) * rows + row] *
            (weights != nullptr ? weights[first + r] : static_cast<FloatT>(1));
        agg = op(agg, value);
    }
    dst[i] = agg;
}

template <typename FloatT>
__global__ void flip_adjacent_columns_kernel(const size_t rows, const size_t half_n, FloatT* const data) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= half_n) return;
    const size_t pair = i / rows, row = i % rows;
    FloatT* const a = data + (2 * pair) * rows + row;
    FloatT* const b = a + rows;
    const FloatT tmp = *a;
    *a = *b;
    *b = tmp;
}

// Per-column sum of op(x): one warp per column, lanes stride the (contiguous) column.
template <typename FloatT, typename Op>
__global__ void reduce_first_axis_kernel(const size_t rows, const size_t cols,
                                         const FloatT* const src, FloatT* const dst, Op op) {
    const size_t warp = (blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x) / 32;
    const unsigned int lane = threadIdx.x % 32;
    if (warp >= cols) return;
    const FloatT* const column = src + warp * rows;
    FloatT agg = 0;
    for (size_t r = lane; r < rows; r += 32) {
        agg += op(column[r]);
    }
    for (int offset = 16; offset > 0; offset /= 2) {
        agg += __shfl_down_sync(0xffffffffu, agg, offset);
    }
    if (lane == 0) dst[warp] = agg;
}

// Per-row sum of op(x) over a slab of columns; thread = row (coalesced across rows).
template <typename FloatT, typename Op>
__global__ void reduce_second_axis_kernel(const size_t rows, const size_t cols,
                                          const size_t cols_per_block,
                                          const FloatT* const src, FloatT* const partial, Op op) {
    const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (row >= rows) return;
    const size_t first = blockIdx.y * cols_per_block;
    const size_t last = first + cols_per_block < cols ? first + cols_per_block : cols;
    FloatT agg = 0;
    for (size_t c = first; c < last; ++c) {
        agg += op(src[c * rows + row]);
    }
    partial[blockIdx.y * rows + row] = agg;
}

template <typename FloatT>
__global__ void sum_partials_kernel(const size_t rows, const size_t num_partials,
                                    const FloatT* const partial, FloatT* const dst) {
    const size_t row = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (row >= rows) return;
    FloatT agg = 0;
    for (size_t p = 0; p < num_partials; ++p) {
        agg += partial[p * rows + row];
    }
    dst[row] = agg;
}

template <typename FloatT>
__global__ void hstack_kernel(const size_t n, const FloatT weight, const FloatT* const src, FloatT* const dst) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    dst[i] = weight * src[i];
}

}  // namespace detail

template <typename Op, typename Policy, typename FloatT>
void apply_elemwise(const Policy& policy, device_matrix<FloatT>* const m, Op op = Op()) {
    thrust::transform(policy, m->begin(), m->end(), m->begin(), op);
}

// M[:, c] = Op(M[:, c], v[c]).
template <typename Op, typename Policy, typename FloatT>
void apply_columnwise(const Policy& policy, const device_matrix<FloatT>& vec,
                      device_matrix<FloatT>* const m, Op op = Op()) {
    CHECK_EQ(vec.size(), m->getCols());
    const size_t n = m->size();
    if (n == 0) return;
    detail::columnwise_kernel<<<detail::num_blocks(n, 256), 256, 0, detail::stream_of(policy)>>>(
        m->getRows(), n, m->getData(), vec.getData(), m->getData(), op,
        func::identity<FloatT>(), func::identity<FloatT>());
    CCE(cudaGetLastError());
}

// dst[:, c] = Op(FirstOp(first[:, c]), SecondOp(v[c])).
template <typename Op, typename FirstOp, typename SecondOp, typename Policy, typename FloatT>
void apply_columnwise(const Policy& policy,
                      const device_matrix<FloatT>& first, const device_matrix<FloatT>& vec,
                      device_matrix<FloatT>* const dst,
                      FirstOp first_op = FirstOp(), SecondOp second_op = SecondOp(), Op op = Op()) {
    CHECK_EQ(vec.size(), first.getCols());
    CHECK(first.hasSameShape(*dst));
    const size_t n = dst->size();
    if (n == 0) return;
    detail::columnwise_kernel<<<detail::num_blocks(n, 256), 256, 0, detail::stream_of(policy)>>>(
        dst->getRows(), n, first.getData(), vec.getData(), dst->getData(), op, first_op, second_op);
    CCE(cudaGetLastError());
}

template <typename Op, typename Policy, typename FloatT>
void apply_except_every_Nth_column(const Policy& policy, const size_t N,
                                   device_matrix<FloatT>* const m, Op op = Op()) {
    const size_t n = m->size();
    if (n == 0) return;
    detail::columns_unary_kernel<<<detail::num_blocks(n, 256), 256, 0, detail::stream_of(policy)>>>(
        m->getRows(), n, N, false, m->getData(), op);
    CCE(cudaGetLastError());
}

template <typename Policy, typename FloatT, typename Op>
void apply_every_Nth_column(const Policy& policy, const size_t N,
                            device_matrix<FloatT>* const m, Op op) {
    const size_t n = m->size();
    if (n == 0) return;
    detail::columns_unary_kernel<<<detail::num_blocks(n, 256), 256, 0, detail::stream_of(policy)>>>(
        m->getRows(), n, N, true, m->getData(), op);
    CCE(cudaGetLastError());
}

// dst = first_op(first) * dst.
template <typename Policy, typename FloatT, typename FirstOp>
void hadamard_product(const Policy& policy, const device_matrix<FloatT>& first,
                      device_matrix<FloatT>* const second_and_dst, FirstOp first_op) {
    CHECK(first.hasSameShape(*second_and_dst));
    thrust::transform(policy,
                      thrust::make_transform_iterator(first.begin(), first_op),
                      thrust::make_transform_iterator(first.end(), first_op),
                      second_and_dst->begin(), second_and_dst->begin(),
                      thrust::multiplies<FloatT>());
}

template <typename Policy, typename FloatT>
void hadamard_product(const Policy& policy, const device_matrix<FloatT>& first,
                      device_matrix<FloatT>* const second_and_dst) {
    hadamard_product(policy, first, second_and_dst, func::identity<FloatT>());
}

template <typename FloatT>
device_matrix<FloatT>* hadamard_product(const cudaStream_t stream,
                                        const device_matrix<FloatT>& first,
                                        const device_matrix<FloatT>& second) {
    CHECK(first.hasSameShape(second));
    device_matrix<FloatT>* const dst = new device_matrix<FloatT>(first.getRows(), first.getCols(), stream);
    thrust::transform(thrust::cuda::par.on(stream), first.begin(), first.end(), second.begin(),
                      dst->begin(), thrust::multiplies<FloatT>());
    return dst;
}

// dst += src_op(src).
template <typename Policy, typename FloatT, typename SrcOp>
void elemwise_plus(const Policy& policy, const device_matrix<FloatT>& src,
                   device_matrix<FloatT>* const dst, SrcOp src_op) {
    CHECK(src.hasSameShape(*dst));
    thrust::transform(policy,
                      thrust::make_transform_iterator(src.begin(), src_op),
                      thrust::make_transform_iterator(src.end(), src_op),
                      dst->begin(), dst->begin(), thrust::plus<FloatT>());
}

template <typename Policy, typename FloatT>
void elemwise_plus(const Policy& policy, const device_matrix<FloatT>& src,
                   device_matrix<FloatT>* const dst) {
    elemwise_plus(policy, src, dst, func::identity<FloatT>());
}

// dst = op(first, dst).
template <typename Policy, typename FloatT, typename Op>
void elemwise_binary(const Policy& policy, const device_matrix<FloatT>& first,
                     device_matrix<FloatT>* const second_and_dst, Op op) {
    CHECK(first.hasSameShape(*second_and_dst));
    thrust::transform(policy, first.begin(), first.end(), second_and_dst->begin(),
                      second_and_dst->begin(), op);
}

// [a b] -> [a a b b].
template <typename FloatT>
device_matrix<FloatT>* broadcast_columns(const cudaStream_t stream, const device_matrix<FloatT>& src,
                                         const size_t reps) {
    device_matrix<FloatT>* const dst =
        new device_matrix<FloatT>(src.getRows(), src.getCols() * reps, stream);
    const size_t n = dst->size();
    if (n > 0) {
        detail::broadcast_columns_kernel<<<detail::num_blocks(n, 256), 256, 0, stream>>>(
            src.getRows(), n, reps, src.getData(), dst->getData());
        CCE(cudaGetLastError());
    }
    return dst;
}

// [a b] -> [a b a b].
template <typename FloatT>
device_matrix<FloatT>* repmat(const cudaStream_t stream, const device_matrix<FloatT>& src,
                              const size_t reps) {
    device_matrix<FloatT>* const dst =
        new device_matrix<FloatT>(src.getRows(), src.getCols() * reps, stream);
    const size_t n = dst->size();
    if (n > 0) {
        detail::repmat_kernel<<<detail::num_blocks(n, 256), 256, 0, stream>>>(
            src.size(), n, src.getData(), dst->getData());
        CCE(cudaGetLastError());
    }
    return dst;
}

// dst[:, i] = Op-fold over r < k of weights[i k + r] * src[:, i k + r].
template <typename FloatT, typename Op = thrust::plus<FloatT> >
device_matrix<FloatT>* fold_columns(const cudaStream_t stream, const device_matrix<FloatT>& src,
                                    const size_t k,
                                    const device_matrix<FloatT>* const weights = nullptr,
                                    Op op = Op()) {
    CHECK_EQ(src.getCols() % k, 0);
    if (weights != nullptr) {
        CHECK_EQ(weights->size(), src.getCols());
    }
    device_matrix<FloatT>* const dst =
        new device_matrix<FloatT>(src.getRows(), src.getCols() / k, stream);
    const size_t n = dst->size();
    if (n > 0) {
        detail::fold_columns_kernel<<<detail::num_blocks(n, 256), 256, 0, stream>>>(
            src.getRows(), n, k, src.getData(),
            weights != nullptr ? weights->getData() : static_cast<FloatT*>(nullptr),
            dst->getData(), op);
        CCE(cudaGetLastError());
    }
    return dst;
}

template <typename FloatT>
void flip_adjacent_columns(const cudaStream_t stream, device_matrix<FloatT>* const m) {
    CHECK_EQ(m->getCols() % 2, 0);
    const size_t half_n = m->size() / 2;
    if (half_n > 0) {
        detail::flip_adjacent_columns_kernel<<<detail::num_blocks(half_n, 256), 256, 0, stream>>>(
            m->getRows(), half_n, m->getData());
        CCE(cudaGetLastError());
    }
}

// FIRST_AXIS: dst[1 x cols] = per-column sums; SECOND_AXIS: dst[rows x 1] = per-row sums. Overwrites dst.
template <typename FloatT, typename Op = func::identity<FloatT> >
void reduce_axis(const cudaStream_t stream, const Axis axis, const device_matrix<FloatT>& src,
                 device_matrix<FloatT>* const dst, Op op = Op()) {
    const size_t rows = src.getRows(), cols = src.getCols();
    if (axis == FIRST_AXIS) {
        CHECK_EQ(dst->size(), cols);
        if (cols == 0) return;
        const unsigned int threads = 256;
        detail::reduce_first_axis_kernel<<<detail::num_blocks(cols * 32, threads), threads, 0, stream>>>(
            rows, cols, src.getData(), dst->getData(), op);
        CCE(cudaGetLastError());
    } else {
        CHECK_EQ(dst->size(), rows);
        if (rows == 0) return;
        const unsigned int threads = rows < 128 ? 32 : 128;
        const unsigned int row_blocks = detail::num_blocks(rows, threads);
        size_t slabs = (1184 + row_blocks - 1) / row_blocks;
        if (slabs > cols) slabs = cols > 0 ? cols : 1;
        const size_t cols_per_block = (cols + slabs - 1) / slabs;
        slabs = cols_per_block > 0 ? (cols + cols_per_block - 1) / cols_per_block : 1;
        device_matrix<FloatT> partial(rows, slabs, stream);
        detail::reduce_second_axis_kernel<<<dim3(row_blocks, slabs), threads, 0, stream>>>(
            rows, cols, cols_per_block, src.getData(), partial.getData(), op);
        CCE(cudaGetLastError());
        detail::sum_partials_kernel<<<row_blocks, threads, 0, stream>>>(
            rows, slabs, partial.getData(), dst->getData());
        CCE(cudaGetLastError());
    }
}

template <typename FloatT>
device_matrix<FloatT>* hstack(const cudaStream_t stream,
                              const ::std::vector< ::std::pair<device_matrix<FloatT>*, FloatT> >& parts) {
    CHECK(!parts.empty());
    const size_t rows = parts.front().first->getRows();
    size_t cols = 0;
    for (size_t i = 0; i < parts.size(); ++i) {
        CHECK_EQ(parts[i].first->getRows(), rows);
        cols += parts[i].first->getCols();
    }
    device_matrix<FloatT>* const dst = new device_matrix<FloatT>(rows, cols, stream);
    size_t offset = 0;
    for (size_t i = 0; i < parts.size(); ++i) {
        const size_t n = parts[i].first->size();
        if (n > 0) {
            detail::hstack_kernel<<<detail::num_blocks(n, 256), 256, 0, stream>>>(
                n, parts[i].second, parts[i].first->getData(), dst->getData() + offset);
            CCE(cudaGetLastError());
        }
        offset += n;
    }
    return dst;
}

//
// GEMM (cuBLAS, column-major): dst = op(A) op(B) (+ dst when dst_contains_bias).
//

namespace detail {

inline cublasStatus_t gemm(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb,
                           int m, int n, int k, const float* alpha, const float* A, int lda,
                           const float* B, int ldb, const float* beta, float* C, int ldc) {
    return cublasSgemm(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

inline cublasStatus_t gemm(cublasHandle_t h, cublasOperation_t ta, cublasOperation_t tb,
                           int m, int n, int k, const double* alpha, const double* A, int lda,
                           const double* B, int ldb, const double* beta, double* C, int ldc) {
    return cublasDgemm(h, ta, tb, m, n, k, alpha, A, lda, B, ldb, beta, C, ldc);
}

}  // namespace detail

template <typename FloatT>
void matrix_mult(const cudaStream_t stream,
                 const device_matrix<FloatT>& first, const cublasOperation_t first_op,
                 const device_matrix<FloatT>& second, const cublasOperation_t second_op,
                 device_matrix<FloatT>* const dst,
                 const bool dst_contains_bias = false) {
    const size_t m = first_op == CUBLAS_OP_N ? first.getRows() : first.getCols();
    const size_t k = first_op == CUBLAS_OP_N ? first.getCols() : first.getRows();
    const size_t k2 = second_op == CUBLAS_OP_N ? second.getRows() : second.getCols();
    const size_t n = second_op == CUBLAS_OP_N ? second.getCols() : second.getRows();
    CHECK_EQ(k, k2);
    CHECK_DIMENSIONS(*dst, m, n);

    cublasHandle_t& handle = Runtime<FloatT>::getInstance()->handle();
    CCBE(cublasSetStream(handle, stream));

    const FloatT alpha = 1.0;
    const FloatT beta = dst_contains_bias ? 1.0 : 0.0;
    CCBE(detail::gemm(handle, first_op, second_op,
                      static_cast<int>(m), static_cast<int>(n), static_cast<int>(k),
                      &alpha, first.getData(), static_cast<int>(first.getRows()),
                      second.getData(), static_cast<int>(second.getRows()),
                      &beta, dst->getData(), static_cast<int>(dst->getRows())));
}

}  // namespace cuda

#endif  // REF_SHIM_DEVICE_MATRIX_H
