// Growable device buffer for the local amplitude slab.
//
// The reference reserves address space for 2^max_local amplitudes up front and doubles the
// vector on every local allocation (reference: SimulatorMPI.cpp:95-96, :168-174).  The B200
// equivalent is CUDA virtual memory management: reserve the VA range once, map physical HBM in
// as the slab grows, so a 5-qubit test does not pin 137 GB and growth never copies.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <utility>
#include <vector>

namespace hiq {

class Slab {
public:
     Slab() = default;
     ~Slab() { release(); }
     Slab(const Slab&) = delete;
     Slab& operator=(const Slab&) = delete;

     // reserve address space for `max_amps` amplitudes on `device`; `shareable` allocations can be
     // exported to the peer processes of the box (POSIX file descriptors)
     int init(int device, uint64_t max_amps, bool shareable = false);
     // make at least `amps` amplitudes addressable (newly mapped memory is NOT cleared)
     int ensure(uint64_t amps);
     void release();

     double2* data() const { return reinterpret_cast<double2*>(base_); }
     uint64_t mapped_amps() const { return mapped_ / sizeof(double2); }
     uint64_t reserved_amps() const { return reserved_ / sizeof(double2); }
     size_t reserved_bytes() const { return reserved_; }
     size_t n_chunks() const { return chunks_.size(); }
     size_t chunk_bytes(size_t i) const { return chunks_[i].second; }
     // a new file descriptor for physical chunk i (the caller closes it after sending it away)
     int export_chunk(size_t i, int* fd) const;
     // exchange the two buffers (address ranges and physical memory): an out-of-place pass ends with a swap, not a copy
     void swap(Slab& o)
     {
          std::swap(device_, o.device_);
          std::swap(base_, o.base_);
          std::swap(reserved_, o.reserved_);
          std::swap(mapped_, o.mapped_);
          std::swap(gran_, o.gran_);
          std::swap(shareable_, o.shareable_);
          chunks_.swap(o.chunks_);
     }

private:
     int device_ = 0;
     CUdeviceptr base_ = 0;
     size_t reserved_ = 0;
     size_t mapped_ = 0;
     size_t gran_ = 0;
     bool shareable_ = false;
     std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks_;
};

// Read/write view of a peer process's slab: its physical chunks imported from file descriptors and
// mapped, in order, into a local address range with access granted to the local GPU (NVLink P2P).
class PeerSlab {
public:
     PeerSlab() = default;
     ~PeerSlab() { release(); }
     PeerSlab(const PeerSlab&) = delete;
     PeerSlab& operator=(const PeerSlab&) = delete;
     PeerSlab(PeerSlab&& o) noexcept { *this = std::move(o); }
     PeerSlab& operator=(PeerSlab&& o) noexcept;

     int init(int local_device, size_t reserve_bytes);
     // takes ownership of fd (closed after the import)
     int map_next_chunk(int fd, size_t bytes);
     void release();

     double2* data() const { return reinterpret_cast<double2*>(base_); }
     size_t n_chunks() const { return chunks_.size(); }
     size_t mapped_bytes() const { return mapped_; }
     uint64_t epoch = 0;  // generation of the peer's slab these chunks belong to

private:
     int device_ = 0;
     CUdeviceptr base_ = 0;
     size_t reserved_ = 0;
     size_t mapped_ = 0;
     std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks_;
};

}  // namespace hiq
