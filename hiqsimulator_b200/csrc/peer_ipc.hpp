// Passing CUDA virtual-memory allocation handles between the per-GPU processes of one box.
//
// The local slab is a cuMemCreate/cuMemMap allocation (slab.hpp).  Exported as POSIX file
// descriptors, its physical chunks can be mapped into the address space of every peer process, after
// which a GPU reads and writes its peers' slabs directly over NVLink/NVSwitch — the transport of the
// in-place qubit swap (swap_kernels.cu).  File descriptors travel over abstract-namespace Unix
// datagram sockets (SCM_RIGHTS), one socket per rank, named after the NCCL unique id of the world.
#pragma once
#include <cstdint>
#include <string>

namespace hiq {

struct FdMessage {
     int src_rank = -1;
     uint32_t index = 0;   // chunk index in the sender's slab
     uint32_t total = 0;   // chunks the sender's slab has right now
     uint64_t size = 0;    // bytes of this chunk
     uint64_t epoch = 0;   // sender's slab generation (a new engine = a new slab)
     int fd = -1;          // received descriptor (owned by the receiver)
};

class FdChannel {
public:
     FdChannel() = default;
     ~FdChannel();
     FdChannel(const FdChannel&) = delete;
     FdChannel& operator=(const FdChannel&) = delete;

     // bind this rank's socket; `world_key` must be identical on all ranks (the NCCL unique id)
     int open(const std::string& world_key, int rank);
     bool is_open() const { return sock_ >= 0; }
     // send one descriptor to `dst_rank` (retries while the peer's socket is not bound yet)
     int send_fd(int dst_rank, const FdMessage& m, int timeout_ms = 20000);
     // receive one descriptor (blocks up to timeout_ms); HIQ_ERR_RUNTIME on timeout
     int recv_fd(FdMessage& m, int timeout_ms = 20000);

private:
     std::string name_of(int rank) const;
     int sock_ = -1;
     int rank_ = 0;
     std::string prefix_;
};

}  // namespace hiq
