// Device-side helpers shared by the sm_100a kernels of the state-vector engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hiq {

constexpr int kMaxTargets = 5;   // the reference's Run() accepts up to 5 fused qubits
                                 // (reference: src/simulator-mpi/SimulatorMPI.cpp:470-524)
constexpr int kNumSMs = 148;     // B200

// Positions (ascending) at which zero bits are inserted into a counter to
// enumerate indices whose target/control bits are clear ("bit deposit").
struct InsertBits {
     int n;
     uint8_t pos[64];
};

__host__ __device__ __forceinline__ uint64_t insert_zero_bits(uint64_t f, const InsertBits& ib)
{
#pragma unroll 1
     for (int i = 0; i < ib.n; ++i) {
          const uint64_t low = f & ((1ull << ib.pos[i]) - 1ull);
          f = ((f >> ib.pos[i]) << (ib.pos[i] + 1)) | low;
     }
     return f;
}

__device__ __forceinline__ void cmac(double2& acc, const double2 m, const double2 v)
{
     acc.x = fma(m.x, v.x, acc.x);
     acc.x = fma(-m.y, v.y, acc.x);
     acc.y = fma(m.x, v.y, acc.y);
     acc.y = fma(m.y, v.x, acc.y);
}

__device__ __forceinline__ double2 cmul(const double2 a, const double2 b)
{
     return make_double2(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x));
}

__device__ __forceinline__ double norm2(const double2 a) { return fma(a.x, a.x, a.y * a.y); }

// 128-bit global accesses. Slabs are streamed once per pass: the loads skip L1
// allocation, the stores are plain (L2 merges the sectors before eviction).
__device__ __forceinline__ double2 ldg_stream(const double2* p)
{
     double2 v;
     asm volatile("ld.global.L1::no_allocate.v2.f64 {%0,%1}, [%2];" : "=d"(v.x), "=d"(v.y) : "l"(p));
     return v;
}

__device__ __forceinline__ double ldg_stream_f64(const double* p)
{
     double v;
     asm volatile("ld.global.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
     return v;
}

__device__ __forceinline__ void cp_async16(void* smem, const void* gmem)
{
     const uint32_t s = static_cast<uint32_t>(__cvta_generic_to_shared(smem));
     asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(s), "l"(gmem) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

__device__ __forceinline__ double warp_sum(double v)
{
#pragma unroll
     for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
     return v;
}

// Deterministic CTA-wide sum (fixed tree); result valid in thread 0.
template <int THREADS>
__device__ __forceinline__ double block_sum(double v, double* scratch /* THREADS/32 doubles */)
{
     v = warp_sum(v);
     const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
     if (lane == 0) scratch[w] = v;
     __syncthreads();
     double r = 0.0;
     if (w == 0) {
          r = lane < THREADS / 32 ? scratch[lane] : 0.0;
          r = warp_sum(r);
     }
     __syncthreads();
     return r;
}

}  // namespace hiq
