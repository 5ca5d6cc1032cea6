// Rank-to-rank communication of the engine: one process per GPU, NCCL over NVLink/NVSwitch.
//
// Replaces the reference's Boost.MPI call sites (reference: SimulatorMPI.cpp:87,289,335,648,692,
// 863,908,964 and the all_to_all in SwapperMT.cpp:115).  World size 1 needs no communicator and
// every collective degenerates to a local copy.
#pragma once
#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdint>
#include <vector>

#include "peer_ipc.hpp"
#include "slab.hpp"

namespace hiq {

class Comm {
public:
     Comm() = default;
     ~Comm();
     Comm(const Comm&) = delete;
     Comm& operator=(const Comm&) = delete;

     // world_size == 1: `unique_id` may be null.  Otherwise all ranks pass the 128-byte id that
     // rank 0 obtained from hiq_comm_unique_id() (the launcher distributes it).
     int init(int rank, int world_size, const void* unique_id, int device);

     // One communicator per process and unique id: an NCCL id can bootstrap exactly one communicator,
     // while a process may create many engines over its lifetime (every SimulatorMPI(...) of the
     // caller).  Returns the cached communicator of (id, rank) or creates it; never destroyed before
     // process exit (the reference's MPI world has the same lifetime).  Null + error on failure.
     static Comm* shared(int rank, int world_size, const void* unique_id, int device);

     int rank() const { return rank_; }
     int size() const { return size_; }
     ncclComm_t handle() const { return comm_; }
     // descriptor channel to the peer processes (opened with the communicator; world size > 1 only)
     FdChannel& fds() { return fds_; }
     // engines of this process created so far on this communicator: the slab generation ("epoch") that the
     // peer-mapping handshake tags its messages with (all ranks create their engines in the same order)
     uint64_t next_epoch() { return ++epoch_; }

     // Staging of the packed exchange (engine.cpp, Engine::exchange_packed): one buffer per process, a virtual-memory
     // allocation whose chunks the peers map through file descriptors (like the slabs).  Process-wide like the
     // communicator: an engine per circuit must not pay the allocation and the descriptor handshake again.
     // A peer process's view of one of my buffers (or mine of its): chunks received and mapped so far, chunks sent so far.
     struct PeerView {
          PeerSlab slab;
          size_t sent = 0;
     };
     struct PackedStaging {
          Slab mine;                    // virtual-memory allocation, exported to the peers chunk by chunk (file descriptors)
          size_t wanted = 0;            // the request that produced this buffer (it may have been halved to fit)
          std::vector<PeerView> peers;  // by world rank: the peers' staging buffers mapped into this process
          bool failed = false;
          size_t bytes() const { return mine.mapped_amps() * sizeof(double2); }
     };
     PackedStaging& packed() { return packed_; }
     // the peer-mapping handshake failed once in this process group (peers on another node, no descriptor channel ...):
     // every later engine goes straight to the staged NCCL exchange instead of waiting for the handshake to time out again
     bool p2p_broken = false;

     // host-value collectives (values staged through a small device buffer on `stream`)
     int allreduce_sum(double* vals, int n, cudaStream_t stream);
     int broadcast_bytes(void* host, size_t bytes, int root, cudaStream_t stream);
     // in-place sum of n doubles that already live in device memory (no host bounce, no synchronisation)
     int allreduce_sum_device(double* dev, int n, cudaStream_t stream);
     // device collective: every rank contributes n doubles, recv holds size()*n (rank-major)
     int allgather(const double* dev_send, double* dev_recv, size_t n, cudaStream_t stream);

private:
     int rank_ = 0;
     int size_ = 1;
     ncclComm_t comm_ = nullptr;
     double* stage_ = nullptr;  // 4 KiB device staging
     FdChannel fds_;
     uint64_t epoch_ = 0;
     PackedStaging packed_;
};

}  // namespace hiq
