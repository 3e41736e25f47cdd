// Stand-in for boost::lockfree::queue (Boost is not installed). Only AsyncSource (data prefetcher,
// outside the training step) holds one; a mutex-guarded deque keeps the declaration compilable.
#ifndef REF_SHIM_BOOST_LOCKFREE_QUEUE_HPP
#define REF_SHIM_BOOST_LOCKFREE_QUEUE_HPP
#include <cstddef>
#include <deque>
#include <mutex>
namespace boost {
namespace lockfree {
template <typename T>
class queue {
 public:
  explicit queue(std::size_t) {}
  bool push(const T& v) { std::lock_guard<std::mutex> g(m_); q_.push_back(v); return true; }
  bool pop(T& v) {
      std::lock_guard<std::mutex> g(m_);
      if (q_.empty()) return false;
      v = q_.front(); q_.pop_front(); return true;
  }
  bool empty() const { std::lock_guard<std::mutex> g(m_); return q_.empty(); }
 private:
  mutable std::mutex m_;
  std::deque<T> q_;
};
}  // namespace lockfree
}  // namespace boost
#endif
