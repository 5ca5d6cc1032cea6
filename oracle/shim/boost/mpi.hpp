// Stand-in for <boost/mpi.hpp> (test infrastructure; neither Boost.MPI nor an
// MPI runtime is installed). It provides exactly the surface the reference
// engine uses (reference: src/simulator-mpi/SimulatorMPI.cpp:73,87,153,289,
// 335,648,692,863,908,964,1111 and SwapperMT.cpp:115).
//
// Two modes, chosen from the environment when the first communicator is built:
//   * single rank (default): collectives are local copies;
//   * multi rank: HIQ_REF_SIZE=R, HIQ_REF_RANK=r, HIQ_REF_SHM=/dev/shm/<name>.
//     R OS processes (one per rank) attach to one shared-memory arena that the
//     launcher created with ftruncate (all-zero). Collectives are
//     copy-in / barrier / copy-out / barrier on per-rank slots. Every
//     collective, including the ones on split communicators, synchronises on
//     the WORLD barrier: the reference calls them in lock-step on all ranks.
#pragma once
#include <fcntl.h>
#include <mpi.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>

#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace hiq_shim {

constexpr size_t kHeaderBytes = 4096;
constexpr size_t kSlotBytes = 1ul << 20;

struct ArenaHeader {
     std::atomic<uint32_t> count;
     std::atomic<uint32_t> sense;
};

class World {
public:
     static World& get()
     {
          static World w;
          return w;
     }
     int rank() const { return rank_; }
     int size() const { return size_; }
     char* slot(int r) const { return base_ + kHeaderBytes + kSlotBytes * static_cast<size_t>(r); }
     void barrier()
     {
          if (size_ == 1) return;
          local_sense_ ^= 1u;
          auto* h = reinterpret_cast<ArenaHeader*>(base_);
          if (h->count.fetch_add(1, std::memory_order_acq_rel) + 1 == static_cast<uint32_t>(size_)) {
               h->count.store(0, std::memory_order_relaxed);
               h->sense.store(local_sense_, std::memory_order_release);
          }
          else {
               while (h->sense.load(std::memory_order_acquire) != local_sense_) sched_yield();
          }
     }
     // publish `bytes` from `src` in my slot and wait until every rank has
     void publish(const void* src, size_t bytes)
     {
          if (bytes > kSlotBytes) {
               std::fprintf(stderr, "hiq_shim: message of %zu bytes exceeds slot\n", bytes);
               std::abort();
          }
          std::memcpy(slot(rank_), src, bytes);
          barrier();
     }

private:
     World()
     {
          const char* s = std::getenv("HIQ_REF_SIZE");
          const char* r = std::getenv("HIQ_REF_RANK");
          const char* p = std::getenv("HIQ_REF_SHM");
          size_ = s ? std::atoi(s) : 1;
          rank_ = r ? std::atoi(r) : 0;
          if (size_ > 1) {
               if (!p) { std::fprintf(stderr, "hiq_shim: HIQ_REF_SHM unset\n"); std::abort(); }
               int fd = ::open(p, O_RDWR);
               if (fd < 0) { std::perror("hiq_shim: open arena"); std::abort(); }
               size_t bytes = kHeaderBytes + kSlotBytes * static_cast<size_t>(size_);
               void* m = ::mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
               if (m == MAP_FAILED) { std::perror("hiq_shim: mmap arena"); std::abort(); }
               ::close(fd);
               base_ = static_cast<char*>(m);
          }
     }
     int rank_ = 0, size_ = 1;
     char* base_ = nullptr;
     uint32_t local_sense_ = 0;
};

}  // namespace hiq_shim

namespace boost {
namespace mpl {
struct true_ { static const bool value = true; };
struct false_ { static const bool value = false; };
}  // namespace mpl

namespace mpi {
namespace threading { enum level { single, funneled, serialized, multiple }; }

struct environment {
     environment() {}
     explicit environment(threading::level) {}
};

class communicator {
public:
     communicator()
     {
          auto& w = hiq_shim::World::get();
          members_.resize(w.size());
          for (int i = 0; i < w.size(); ++i) members_[i] = i;
          me_ = w.rank();
     }
     int size() const { return static_cast<int>(members_.size()); }
     int rank() const { return me_; }
     int world_rank_of(int comm_rank) const { return members_[comm_rank]; }
     void barrier() const { hiq_shim::World::get().barrier(); }
     // members of the new communicator are ordered by world rank
     communicator split(int color) const
     {
          auto& w = hiq_shim::World::get();
          communicator c(*this);
          if (w.size() == 1) return c;
          w.publish(&color, sizeof(color));
          c.members_.clear();
          for (int r: members_) {
               int other;
               std::memcpy(&other, w.slot(r), sizeof(other));
               if (other == color) {
                    if (r == w.rank()) c.me_ = static_cast<int>(c.members_.size());
                    c.members_.push_back(r);
               }
          }
          w.barrier();
          return c;
     }

private:
     std::vector<int> members_;
     int me_ = 0;
};

template <class T>
MPI_Datatype get_mpi_datatype(const T&) { return 0; }
template <class T>
struct is_mpi_complex_datatype : mpl::false_ {};
template <class Op, class T>
struct is_commutative : mpl::false_ {};

template <class T>
void broadcast(const communicator& c, T& v, int root)
{
     auto& w = hiq_shim::World::get();
     if (w.size() == 1) return;
     w.publish(&v, sizeof(T));
     std::memcpy(&v, w.slot(c.world_rank_of(root)), sizeof(T));
     w.barrier();
}

template <class T, class Op>
void all_reduce(const communicator& c, const T* in, int n, T* out, Op op)
{
     auto& w = hiq_shim::World::get();
     if (w.size() == 1) {
          for (int i = 0; i < n; ++i) out[i] = in[i];
          return;
     }
     w.publish(in, sizeof(T) * n);
     for (int i = 0; i < n; ++i) {
          T acc;
          std::memcpy(&acc, w.slot(c.world_rank_of(0)) + sizeof(T) * i, sizeof(T));
          for (int r = 1; r < c.size(); ++r) {
               T v;
               std::memcpy(&v, w.slot(c.world_rank_of(r)) + sizeof(T) * i, sizeof(T));
               acc = op(acc, v);
          }
          out[i] = acc;
     }
     w.barrier();
}
template <class T, class Op>
void all_reduce(const communicator& c, const T& in, T& out, Op op) { all_reduce(c, &in, 1, &out, op); }
template <class T, class Op>
T all_reduce(const communicator& c, const T& in, Op op)
{
     T out;
     all_reduce(c, &in, 1, &out, op);
     return out;
}

template <class T>
void all_gather(const communicator& c, const T* in, int n, T* out)
{
     auto& w = hiq_shim::World::get();
     if (w.size() == 1) {
          std::memcpy((void*) out, (const void*) in, sizeof(T) * n);
          return;
     }
     w.publish(in, sizeof(T) * n);
     for (int r = 0; r < c.size(); ++r)
          std::memcpy((void*) (out + static_cast<size_t>(n) * r), w.slot(c.world_rank_of(r)), sizeof(T) * n);
     w.barrier();
}

// rank j receives, at out[n*i ..], the block in[n*j ..] of rank i
template <class T>
void all_to_all(const communicator& c, const T* in, int n, T* out)
{
     auto& w = hiq_shim::World::get();
     if (w.size() == 1) {
          std::memcpy((void*) out, (const void*) in, sizeof(T) * n);
          return;
     }
     w.publish(in, sizeof(T) * n * c.size());
     for (int r = 0; r < c.size(); ++r)
          std::memcpy((void*) (out + static_cast<size_t>(n) * r),
                      w.slot(c.world_rank_of(r)) + sizeof(T) * n * static_cast<size_t>(c.rank()), sizeof(T) * n);
     w.barrier();
}

}  // namespace mpi
}  // namespace boost
