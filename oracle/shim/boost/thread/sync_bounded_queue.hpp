// Stand-in for <boost/thread/sync_bounded_queue.hpp> (test infrastructure).
// Blocking bounded FIFO with the three calls the reference's swap pipeline
// uses (reference: src/simulator-mpi/SwapArrays.hpp:66-82, SwapperMT.cpp:34-121).
#pragma once
#include <condition_variable>
#include <deque>
#include <mutex>
namespace boost {
template <class T>
class sync_bounded_queue {
public:
     explicit sync_bounded_queue(size_t cap) : cap_(cap) {}
     void push_back(const T& v)
     {
          std::unique_lock<std::mutex> l(m_);
          cv_.wait(l, [&] { return q_.size() < cap_; });
          q_.push_back(v);
          cv_.notify_all();
     }
     T pull_front()
     {
          std::unique_lock<std::mutex> l(m_);
          cv_.wait(l, [&] { return !q_.empty(); });
          T v = q_.front();
          q_.pop_front();
          cv_.notify_all();
          return v;
     }
     void pull_front(T& v) { v = pull_front(); }
private:
     std::deque<T> q_;
     std::mutex m_;
     std::condition_variable cv_;
     size_t cap_;
};
}  // namespace boost
