// Device-side helpers shared by the sm_100a kernels of the state-vector engine.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace hiq {

constexpr int kMaxTargets = 5;   // the reference's Run() accepts up to 5 fused qubits
                                 // (reference: src/simulator-mpi/SimulatorMPI.cpp:470-524)
constexpr int kNumSMsB200 = 148;  // B200; the launchers size their grids from the device's own count (num_sms())

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
// compose by multiplication, so a launch can apply up to kMaxDiagOps of them in ONE pass over HBM:
// psi[i] *= prod_j lut_j[bits of i at slots_j].  The same program rides along a dense launch,
// applied to the tuple it loads.
//
// Cost model: an element index is split as  (chunk bits | u bits | tid bits).  The selector of op j
// is the OR of three partial selectors: selh_j(chunk) computed once per CTA iteration by warp 0,
// selt_j(tid) computed once per thread per kernel, usel_j[u] tabulated by the host.  Ops that do not
// depend on u (class S0) collapse into ONE factor per thread per chunk; ops that do (class S1) cost a
// lookup + complex multiply per element.  The host picks the u bit positions among the index bits the
// ops touch least, so S1 is usually empty.
// ---------------------------------------------------------------------------------------------
constexpr int kMaxDiagOps = 16;
constexpr int kMaxUBits = 4;

struct DiagProg {
     int n;                                        // ops in use, ordered S0 | S1 | E
     int n_s0, n_s1, n_e;                          // class sizes (n_e > 0 only along a dense launch)
     int n_s0a;                                    // leading S0 ops that do not depend on tid either: one factor per CTA per chunk
     uint8_t slots[kMaxDiagOps][8];                // selector bit l <-> index bit slots[j][l]; unused = 63 (always 0)
     uint8_t usel[kMaxDiagOps][1 << kMaxUBits];    // selector bits contributed by the u part of the index
     double2 lut[kMaxDiagOps][1 << kMaxTargets];
};

// selector of an op for index idx (bit l of the selector = bit slots[l] of idx)
__device__ __forceinline__ uint32_t diag_select(const uint8_t (&slots)[8], uint64_t idx)
{
     uint32_t sel = 0;
#pragma unroll
     for (int l = 0; l < kMaxTargets; ++l) sel |= static_cast<uint32_t>((idx >> slots[l]) & 1ull) << l;
     return sel;
}

struct DiagShared {
     double2 lut[kMaxDiagOps][1 << kMaxTargets];
     uint32_t selh[kMaxDiagOps];                   // per-chunk partial selectors
     double2 s_hi;                                 // per-chunk product of the CTA-uniform ops [0, n_s0a)
};

// kernel start: LUTs to shared memory, this thread's partial selectors to `selt` (packed 8 bits per op)
template <int THREADS>
__device__ __forceinline__ void diag_prog_init(const DiagProg& p, DiagShared& sh, uint64_t tid_idx, uint32_t (&selt)[kMaxDiagOps / 4])
{
     for (int i = threadIdx.x; i < p.n * (1 << kMaxTargets); i += THREADS)
          sh.lut[i >> kMaxTargets][i & ((1 << kMaxTargets) - 1)] = p.lut[i >> kMaxTargets][i & ((1 << kMaxTargets) - 1)];
#pragma unroll
     for (int w = 0; w < kMaxDiagOps / 4; ++w) {
          uint32_t v = 0;
#pragma unroll
          for (int b = 0; b < 4; ++b)
               if (4 * w + b < p.n) v |= diag_select(p.slots[4 * w + b], tid_idx) << (8 * b);
          selt[w] = v;
     }
}

__device__ __forceinline__ uint32_t diag_selt(const uint32_t (&selt)[kMaxDiagOps / 4], int j)
{
     // j is uniform across the CTA: a 4-way select on registers, no local memory
     const uint32_t w = (j < 4) ? selt[0] : (j < 8) ? selt[1] : (j < 12) ? selt[2] : selt[3];
     return (w >> (8 * (j & 3))) & 0xffu;
}

// once per chunk (all threads call it): partial selectors at the chunk's base index and the product
// of the CTA-uniform factors (shuffle tree in warp 0)
__device__ __forceinline__ void diag_prog_chunk(const DiagProg& p, DiagShared& sh, uint64_t chunk_idx)
{
     __syncthreads();  // readers of the previous chunk's values are done (and the LUT fill is visible)
     if (threadIdx.x < 32) {
          const int j = threadIdx.x;
          double2 f = make_double2(1.0, 0.0);
          if (j < p.n) {
               const uint32_t sel = diag_select(p.slots[j], chunk_idx);
               sh.selh[j] = sel;
               if (j < p.n_s0a) f = sh.lut[j][sel];
          }
#pragma unroll
          for (int o = kMaxDiagOps / 2; o > 0; o >>= 1) {
               double2 g;
               g.x = __shfl_xor_sync(0xffffffffu, f.x, o);
               g.y = __shfl_xor_sync(0xffffffffu, f.y, o);
               f = cmul(f, g);
          }
          if (j == 0) sh.s_hi = f;
     }
     __syncthreads();
}

// product of the class-S0 factors for this thread in this chunk (two independent chains)
__device__ __forceinline__ double2 diag_prog_s0(const DiagProg& p, const DiagShared& sh, const uint32_t (&selt)[kMaxDiagOps / 4])
{
     double2 s = sh.s_hi;
     double2 r = make_double2(1.0, 0.0);
     int j = p.n_s0a;
     for (; j + 1 < p.n_s0; j += 2) {
          s = cmul(s, sh.lut[j][sh.selh[j] | diag_selt(selt, j)]);
          r = cmul(r, sh.lut[j + 1][sh.selh[j + 1] | diag_selt(selt, j + 1)]);
     }
     if (j < p.n_s0) s = cmul(s, sh.lut[j][sh.selh[j] | diag_selt(selt, j)]);
     return cmul(s, r);
}

// multiply `s` by the class-S1 factors for element part u
__device__ __forceinline__ double2 diag_prog_s1(const DiagProg& p, const DiagShared& sh, const uint32_t (&selt)[kMaxDiagOps / 4], int u,
                                               double2 s)
{
     for (int j = p.n_s0; j < p.n_s0 + p.n_s1; ++j) s = cmul(s, sh.lut[j][sh.selh[j] | diag_selt(selt, j) | p.usel[j][u]]);
     return s;
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
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait_group()
{
     asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

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
