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

// ---------------------------------------------------------------------------------------------
// Batched diagonal factors.  Consecutive diagonal fused gates of the reference's plan (each one
// would be a full pass of kernel_core_diag, reference: kernels/intrin/kernels_diag.hpp:35-144)
// commute with every index permutation and compose by multiplication, so a launch can apply up
// to kMaxDiagOps of them in ONE pass over HBM: psi[i] *= prod_j lut_j[bits of i at slots_j].
// The same structure rides along a dense launch as "pre-diagonals" applied to the loaded tuple.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxDiagOps = 16;

// Work is cut into *chunks* of consecutive indices (one CTA iteration each).  A slot whose index
// bit cannot change inside a chunk is "chunk-constant": its selector bits are computed once per
// chunk by one warp; an op made only of such slots collapses into a single per-chunk factor.
struct DiagBatch {
     int n;                               // ops in use
     int n_lo;                            // ops [0, n_lo) need per-element work, ops [n_lo, n) are chunk-constant
     uint8_t n_lo_slots[kMaxDiagOps];     // leading entries of slots[j] that vary inside a chunk
     uint8_t slots[kMaxDiagOps][8];       // reordered: chunk-varying slots first; unused entries = 63 (always-0 bit)
     double2 lut[kMaxDiagOps][1 << kMaxTargets];  // permuted to the reordered slots
};

// selector of an op for index idx (bit l of the selector = bit slots[l] of idx)
__device__ __forceinline__ uint32_t diag_select(const uint8_t (&slots)[8], uint64_t idx)
{
     uint32_t sel = 0;
#pragma unroll
     for (int l = 0; l < kMaxTargets; ++l) sel |= static_cast<uint32_t>((idx >> slots[l]) & 1ull) << l;
     return sel;
}

// selector bits contributed by the n_lo leading (chunk-varying) slots
__device__ __forceinline__ uint32_t diag_select_lo(const uint8_t (&slots)[8], int n_lo, uint64_t idx)
{
     uint32_t sel = 0;
     for (int l = 0; l < n_lo; ++l) sel |= static_cast<uint32_t>((idx >> slots[l]) & 1ull) << l;
     return sel;
}

struct DiagHoist {
     double2 s_hi;                  // product of the chunk-constant ops
     uint32_t selh[kMaxDiagOps];    // chunk-constant selector bits of every op
};

// Once per chunk (all threads of the CTA call it): selectors at the chunk's base index, then the
// product of the chunk-constant factors by a shuffle tree in warp 0.
__device__ __forceinline__ void diag_hoist(const DiagBatch& b, const double2 (*lut)[1 << kMaxTargets], uint64_t chunk_base_idx,
                                           DiagHoist& h)
{
     __syncthreads();  // readers of the previous chunk's values are done
     if (threadIdx.x < 32) {
          const int j = threadIdx.x;
          double2 f = make_double2(1.0, 0.0);
          if (j < b.n) {
               const uint32_t sel = diag_select(b.slots[j], chunk_base_idx);
               h.selh[j] = sel;
               if (j >= b.n_lo) f = lut[j][sel];
          }
#pragma unroll
          for (int o = kMaxDiagOps / 2; o > 0; o >>= 1) {
               double2 g;
               g.x = __shfl_xor_sync(0xffffffffu, f.x, o);
               g.y = __shfl_xor_sync(0xffffffffu, f.y, o);
               f = cmul(f, g);
          }
          if (j == 0) h.s_hi = f;
     }
     __syncthreads();
}

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
