/* oracle/shims/tbb/concurrent_queue.h -- mutex-based stand-in for tbb::concurrent_queue
 * (TEST INFRASTRUCTURE ONLY: lets the reference's core/ntsDataloador.hpp and toolkits/ be
 * compiled against the adaptor header without oneTBB installed). */
#ifndef NTS_ORACLE_SHIM_TBB_CONCURRENT_QUEUE_H
#define NTS_ORACLE_SHIM_TBB_CONCURRENT_QUEUE_H
#include <deque>
#include <mutex>
namespace tbb {
template <typename T> class concurrent_queue {
  std::deque<T> q_;
  mutable std::mutex m_;
public:
  concurrent_queue() {}
  concurrent_queue(const concurrent_queue &o) { std::lock_guard<std::mutex> g(o.m_); q_ = o.q_; }
  void push(const T &v) { std::lock_guard<std::mutex> g(m_); q_.push_back(v); }
  bool try_pop(T &v) { std::lock_guard<std::mutex> g(m_); if (q_.empty()) return false; v = q_.front(); q_.pop_front(); return true; }
  bool empty() const { std::lock_guard<std::mutex> g(m_); return q_.empty(); }
  size_t unsafe_size() const { std::lock_guard<std::mutex> g(m_); return q_.size(); }
  void clear() { std::lock_guard<std::mutex> g(m_); q_.clear(); }
};
}  // namespace tbb
#endif
