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
#include <vector>

namespace hiq {

class Slab {
public:
     Slab() = default;
     ~Slab() { release(); }
     Slab(const Slab&) = delete;
     Slab& operator=(const Slab&) = delete;

     // reserve address space for `max_amps` amplitudes on `device`
     int init(int device, uint64_t max_amps);
     // make at least `amps` amplitudes addressable (newly mapped memory is NOT cleared)
     int ensure(uint64_t amps);
     void release();

     double2* data() const { return reinterpret_cast<double2*>(base_); }
     uint64_t mapped_amps() const { return mapped_ / sizeof(double2); }
     uint64_t reserved_amps() const { return reserved_ / sizeof(double2); }

private:
     int device_ = 0;
     CUdeviceptr base_ = 0;
     size_t reserved_ = 0;
     size_t mapped_ = 0;
     size_t gran_ = 0;
     std::vector<std::pair<CUmemGenericAllocationHandle, size_t>> chunks_;
};

}  // namespace hiq
