// Dense k-qubit complex128 gate application over the local amplitude slab (k = 1..5).
//
// Replaces the reference CPU kernels kernelK<V,M,kernel_core>
// (reference: src/simulator-mpi/kernels/intrin/kernel{1..5}.hpp, scalar statement in
// kernels/nointrin/kernel{1..5}.hpp; dispatched from SimulatorMPI::Run, SimulatorMPI.cpp:470-515):
//   for every base index I with target bits 0 and (I & ctrl_mask) == ctrl_mask
//     out[b] = sum_c m[b][c] * in[c],  element c at I + sum_l c_l << slots[l]
// One pass reads and writes the slab once: 32 B of HBM traffic per amplitude.
//
// Three kernels, picked by the launcher from the target slots:
//   DIRECT  one 2^k tuple per thread, consecutive lanes take consecutive free indices so every
//           warp-wide 128-bit access covers whole 32 B sectors (needs the lowest target slot >= 2).
//           The matrix lives in the kernel-parameter constant bank and is consumed as the
//           constant operand of DFMA (fully unrolled, k <= 4).
//   TILED   targets in slots 0/1: a tile = {lo contiguous low slots} U {targets above them} is
//           staged in shared memory with coalesced cp.async, tuples are gathered from the tile
//           through a per-launch XOR swizzle (bank-conflict free for any target set), written
//           back in place and stored coalesced.
//   DMMA    FP64 tensor-core path: the complex 2^k x 2^k product is the real 2^(k+1)-dim
//           product out^T = in^T * Mreal^T computed with mma.sync.m8n8k4.f64, 8 tuples per
//           warp-level MMA group; the A fragment is loaded straight from the slab (two lanes per
//           amplitude), the accumulator fragment is exactly one complex128 per lane and is
//           stored with one 128-bit store.
#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <type_traits>
#include <vector>

#include "dense_rows.cuh"
#include "hiq_device.cuh"
#include "hiq_host.hpp"

namespace hiq {

// ---------------------------------------------------------------------------------------------
// DIRECT
// ---------------------------------------------------------------------------------------------
template <int K>
struct DirectParams {
     double2* psi;
     uint64_t n_free;     // number of tuples to process
     uint64_t ctrl_mask;  // OR-ed into every base index
     InsertBits ins;      // target and control slots, ascending
     uint64_t off[1 << K];
     double2 m[1 << (2 * K)];
     double msum[K == 4 ? (1 << (2 * K)) : 1];  // Re + Im of every entry (three-multiplication product, K = 4 only)
};

template <int K>
__device__ __forceinline__ void load_tuple(double2 (&in)[1 << K], const double2* base,
                                           const uint64_t (&off)[1 << K])
{
#pragma unroll
     for (int c = 0; c < (1 << K); ++c) in[c] = ldg_stream(base + off[c]);
}

template <int K, int THREADS, int MINB, int KS = K, bool M3 = false>
__global__ void __launch_bounds__(THREADS, MINB) dense_direct_kernel(const __grid_constant__ DirectParams<K> p)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * THREADS;
     for (uint64_t f = static_cast<uint64_t>(blockIdx.x) * THREADS + threadIdx.x; f < p.n_free; f += stride) {
          double2* base = p.psi + (insert_zero_bits(f, p.ins) | p.ctrl_mask);
          double2 in[1 << K];
          load_tuple<K>(in, base, p.off);
          if constexpr (M3) apply_rows_3m(in, p.m, p.msum, [&](int b, double2 v) { base[p.off[b]] = v; });
          else apply_rows<K, KS>(in, p.m, [&](int b, double2 v) { base[p.off[b]] = v; });
     }
}

// DIRECT, staged: same arithmetic as dense_direct_kernel, but every thread's NEXT tuple is copied into
// its shared-memory column by cp.async while the current tuple is multiplied, so the HBM latency of a
// tuple hides behind 2^(2K) complex MACs instead of stalling the (register-limited, 4 warps/scheduler) CTA.
template <int K, int THREADS, int MINB, int KS = K, bool M3 = false>
__global__ void __launch_bounds__(THREADS, MINB) dense_direct_staged_kernel(const __grid_constant__ DirectParams<K> p)
{
     extern __shared__ double2 dyn_smem[];
     double2 (*stage)[THREADS] = reinterpret_cast<double2 (*)[THREADS]>(dyn_smem);
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * THREADS;
     auto tuple_base = [&](uint64_t f) { return p.psi + (insert_zero_bits(f, p.ins) | p.ctrl_mask); };
     auto prefetch = [&](const double2* base) {
#pragma unroll
          for (int c = 0; c < (1 << K); ++c) cp_async16(&stage[c][threadIdx.x], base + p.off[c]);
     };
     uint64_t f = static_cast<uint64_t>(blockIdx.x) * THREADS + threadIdx.x;
     if (f < p.n_free) prefetch(tuple_base(f));
     for (; f < p.n_free; f += stride) {
          double2* base = tuple_base(f);
          cp_async_wait_all();  // this thread's own copies; nobody else reads its column
          double2 in[1 << K];
#pragma unroll
          for (int c = 0; c < (1 << K); ++c) in[c] = stage[c][threadIdx.x];
          if (f + stride < p.n_free) prefetch(tuple_base(f + stride));
          if constexpr (M3) apply_rows_3m(in, p.m, p.msum, [&](int b, double2 v) { base[p.off[b]] = v; });
          else apply_rows<K, KS>(in, p.m, [&](int b, double2 v) { base[p.off[b]] = v; });
     }
}

// DIRECT with diagonals folded in: psi <- M * (prod_j D_j) * psi in one pass (DiagProg, hiq_device.cuh).
// Free index = (chunk bits | t bits | tid bits): a CTA iteration covers THREADS consecutive free indices
// (tid) times 2^n_t tuples per thread (t) whose index positions are the free bits the ops touch least.
// Ops that avoid the dense targets are per-tuple scalars — S0: one factor per thread per chunk, S1: one
// lookup per tuple — and are multiplied into ONE scalar s.  An op that overlaps the targets (class E)
// needs a factor per tuple element: the first one is combined with s into the 2^m distinct values
// s * lut[...] (m = overlapping target bits) staged in this thread's shared-memory column, so the whole
// diagonal program costs 2^K complex multiplies per tuple on top of the matrix product.
constexpr int kMaxTBits = 3;

template <int K>
struct DirectPreParams {
     DirectParams<K> d;                      // d.ins = targets U t positions; d.n_free = 2^(L - K - n_t)
     int n_t;
     int fast;                               // no op depends on t: every factor is fixed per thread per chunk
     uint64_t toff[1 << kMaxTBits];          // index offset of tuple t of a thread
     int e_npat;                             // distinct joint patterns of the class-E ops over the tuple elements
     uint8_t e_pat[kMaxDiagOps][1 << K];     // selector bits of class-E op j for joint pattern e
     uint8_t e_cmap[1 << K];                 // joint pattern of tuple element c
     DiagProg prog;
};

template <int K, int THREADS, int MINB, int KS = K, bool M3 = false>
__global__ void __launch_bounds__(THREADS, MINB) dense_direct_pre_kernel(const __grid_constant__ DirectPreParams<K> p)
{
     // dynamic shared memory: stage[2^K][THREADS] (this thread's NEXT tuple, filled by cp.async while the
     // current one is being multiplied) followed by sT[e_npat][THREADS] (this thread's class-E factors)
     extern __shared__ double2 dyn_smem[];
     double2 (*stage)[THREADS] = reinterpret_cast<double2 (*)[THREADS]>(dyn_smem);
     double2 (*sT)[THREADS] = reinterpret_cast<double2 (*)[THREADS]>(dyn_smem + (1 << K) * THREADS);
     __shared__ DiagShared sh;
     uint32_t selt[kMaxDiagOps / 4];
     diag_prog_init<THREADS>(p.prog, sh, insert_zero_bits(threadIdx.x, p.d.ins), selt);
     const uint64_t n_chunks = (p.d.n_free + THREADS - 1) / THREADS;
     const int nt = 1 << p.n_t;
     const int je = p.prog.n_s0 + p.prog.n_s1;  // first class-E op
     const bool valid = threadIdx.x < p.d.n_free;  // only a slab smaller than one chunk leaves threads idle
     // factors of the joint patterns for tuple t of this thread, scalar part s folded in
     auto build_patterns = [&](double2 s, int t) {
          for (int e = 0; e < p.e_npat; ++e) {
               double2 f = s;
               for (int j = je; j < p.prog.n; ++j)
                    f = cmul(f, sh.lut[j][sh.selh[j] | diag_selt(selt, j) | p.prog.usel[j][t] | p.e_pat[j][e]]);
               sT[e][threadIdx.x] = f;
          }
     };
     // index of this thread's tuple 0 in a chunk (bit deposit, once per chunk); tuple t adds toff[t]
     auto chunk_index = [&](uint64_t chunk) { return insert_zero_bits(chunk * THREADS + threadIdx.x, p.d.ins); };
     auto prefetch = [&](const double2* base) {
#pragma unroll
          for (int c = 0; c < (1 << K); ++c) cp_async16(&stage[c][threadIdx.x], base + p.d.off[c]);
     };
     uint64_t chunk = blockIdx.x;
     uint64_t bidx = chunk < n_chunks ? chunk_index(chunk) : 0;
     if (chunk < n_chunks && valid) prefetch(p.d.psi + bidx);
     for (; chunk < n_chunks; chunk += gridDim.x) {
          diag_prog_chunk(p.prog, sh, insert_zero_bits(chunk * THREADS, p.d.ins));
          if (!valid) continue;
          const double2 s0 = diag_prog_s0(p.prog, sh, selt);
          if (p.fast && p.prog.n_e) build_patterns(s0, 0);
          const uint64_t bidx_next = chunk + gridDim.x < n_chunks ? chunk_index(chunk + gridDim.x) : 0;
#pragma unroll 1
          for (int t = 0; t < nt; ++t) {
               double2* base = p.d.psi + (bidx | p.toff[t]);
               double2 s = s0;
               if (!p.fast) {
                    s = diag_prog_s1(p.prog, sh, selt, t, s0);
                    if (p.prog.n_e) build_patterns(s, t);
               }
               cp_async_wait_all();  // this thread's own copies: no CTA barrier needed, nobody else reads them
               double2 in[1 << K];
               if (p.prog.n_e == 0) {
#pragma unroll
                    for (int c = 0; c < (1 << K); ++c) in[c] = cmul(stage[c][threadIdx.x], s);
               }
               else {
#pragma unroll
                    for (int c = 0; c < (1 << K); ++c) in[c] = cmul(stage[c][threadIdx.x], sT[p.e_cmap[c]][threadIdx.x]);
               }
               // the staged tuple now lives in registers (the multiplies above consumed it): refill the stage
               // with this thread's next tuple so its HBM latency hides behind the matrix product
               if (t + 1 < nt) prefetch(p.d.psi + (bidx | p.toff[t + 1]));
               else if (chunk + gridDim.x < n_chunks) prefetch(p.d.psi + bidx_next);
               if constexpr (M3) apply_rows_3m(in, p.d.m, p.d.msum, [&](int b, double2 v) { base[p.d.off[b]] = v; });
               else apply_rows<K, KS>(in, p.d.m, [&](int b, double2 v) { base[p.d.off[b]] = v; });
          }
          bidx = bidx_next;
     }
}

// DIRECT with diagonals folded in, block form (KS < K), register-pipelined.
// The reduced product never mixes the 2^(K-KS) blocks of a tuple, so a thread only needs one block (2^KS elements)
// live at a time: the next block is loaded into registers while the current one is multiplied and stored, nothing is
// staged through shared memory (the staged kernel spends a quarter of its stalls on the shared-memory queue,
// profiles/r01n_ncu_full_blocks_L30.md) and the register budget allows more resident warps.  Same diagonal program,
// chunking and factor tables as dense_direct_pre_kernel; results are identical.
template <int K, int KS, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dense_direct_pre_blocks_kernel(const __grid_constant__ DirectPreParams<K> p)
{
     static_assert(KS < K, "block form only");
     constexpr int D = 1 << K;
     constexpr int DS = 1 << KS;
     constexpr int NB = 1 << (K - KS);
     extern __shared__ double2 dyn_smem[];
     double2 (*sT)[THREADS] = reinterpret_cast<double2 (*)[THREADS]>(dyn_smem);  // this thread's class-E factors
     __shared__ DiagShared sh;
     uint32_t selt[kMaxDiagOps / 4];
     diag_prog_init<THREADS>(p.prog, sh, insert_zero_bits(threadIdx.x, p.d.ins), selt);
     const uint64_t n_chunks = (p.d.n_free + THREADS - 1) / THREADS;
     const int nt = 1 << p.n_t;
     const int je = p.prog.n_s0 + p.prog.n_s1;
     const bool valid = threadIdx.x < p.d.n_free;
     auto build_patterns = [&](double2 s, int t) {
          for (int e = 0; e < p.e_npat; ++e) {
               double2 f = s;
               for (int j = je; j < p.prog.n; ++j)
                    f = cmul(f, sh.lut[j][sh.selh[j] | diag_selt(selt, j) | p.prog.usel[j][t] | p.e_pat[j][e]]);
               sT[e][threadIdx.x] = f;
          }
     };
     auto chunk_index = [&](uint64_t chunk) { return insert_zero_bits(chunk * THREADS + threadIdx.x, p.d.ins); };
     auto load_block = [&](double2 (&dst)[DS], const double2* base, int blk) {
#pragma unroll
          for (int j = 0; j < DS; ++j) dst[j] = ldg_stream(base + p.d.off[blk * DS + j]);
     };
     uint64_t chunk = blockIdx.x;
     uint64_t bidx = chunk < n_chunks ? chunk_index(chunk) : 0;
     double2 cur[DS], nxt[DS];
     if (chunk < n_chunks && valid) load_block(cur, p.d.psi + bidx, 0);
     for (; chunk < n_chunks; chunk += gridDim.x) {
          diag_prog_chunk(p.prog, sh, insert_zero_bits(chunk * THREADS, p.d.ins));
          if (!valid) continue;
          const double2 s0 = diag_prog_s0(p.prog, sh, selt);
          if (p.fast && p.prog.n_e) build_patterns(s0, 0);
          const bool more_chunks = chunk + gridDim.x < n_chunks;
          const uint64_t bidx_next = more_chunks ? chunk_index(chunk + gridDim.x) : 0;
#pragma unroll 1
          for (int t = 0; t < nt; ++t) {
               double2* base = p.d.psi + (bidx | p.toff[t]);
               double2 s = s0;
               if (!p.fast) {
                    s = diag_prog_s1(p.prog, sh, selt, t, s0);
                    if (p.prog.n_e) build_patterns(s, t);
               }
#pragma unroll
               for (int blk = 0; blk < NB; ++blk) {
                    // the block after this one (next block of the tuple, else block 0 of the thread's next tuple)
                    if (blk + 1 < NB) load_block(nxt, base, blk + 1);
                    else if (t + 1 < nt) load_block(nxt, p.d.psi + (bidx | p.toff[t + 1]), 0);
                    else if (more_chunks) load_block(nxt, p.d.psi + bidx_next, 0);
                    double2 in[DS];
                    if (p.prog.n_e == 0) {
#pragma unroll
                         for (int j = 0; j < DS; ++j) in[j] = cmul(cur[j], s);
                    }
                    else {
#pragma unroll
                         for (int j = 0; j < DS; ++j) in[j] = cmul(cur[j], sT[p.e_cmap[blk * DS + j]][threadIdx.x]);
                    }
#pragma unroll
                    for (int r = 0; r < DS; ++r) {
                         const int b = blk * DS + r;
                         double2 acc = make_double2(0.0, 0.0);
#pragma unroll
                         for (int j = 0; j < DS; ++j) cmac(acc, p.d.m[b * D + blk * DS + j], in[j]);
                         base[p.d.off[b]] = acc;
                    }
#pragma unroll
                    for (int j = 0; j < DS; ++j) cur[j] = nxt[j];
               }
          }
          bidx = bidx_next;
     }
}

// ---------------------------------------------------------------------------------------------
// TILED
// ---------------------------------------------------------------------------------------------
template <int K>
struct TiledParams {
     double2* psi;
     uint64_t n_tiles;
     uint64_t hi_ctrl_mask;  // control slots >= lo (global positions): fixed to 1
     uint32_t lo_ctrl_mask;  // control slots <  lo: predicate on the tile-local index
     int lo;                 // contiguous low slots in the tile
     int tile_bits;          // lo + number of targets >= lo
     int nswz;
     uint32_t swz_src[3];    // tile-local bit that is XOR-ed ...
     uint32_t swz_dst[3];    // ... into this bit (< 3) of the shared-memory position
     InsertBits outer;       // (slot - lo) of high targets and high controls, ascending
     InsertBits inner;       // tile-local positions of the targets, ascending
     uint64_t hoff[32];      // global offset of high-target combination h
     uint32_t loff[1 << K];  // tile-local offset of matrix index c
     double2 m[1 << (2 * K)];
};

template <int K>
__device__ __forceinline__ uint32_t swizzle(uint32_t j, const TiledParams<K>& p)
{
     uint32_t x = 0;
#pragma unroll
     for (int i = 0; i < 3; ++i)
          if (i < p.nswz) x |= ((j >> p.swz_src[i]) & 1u) << p.swz_dst[i];
     return j ^ x;
}

template <int K, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dense_tiled_kernel(const __grid_constant__ TiledParams<K> p)
{
     extern __shared__ double2 tile[];
     const uint32_t tile_amps = 1u << p.tile_bits;
     const uint32_t lo_mask = (1u << p.lo) - 1u;
     for (uint64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
          double2* base = p.psi + ((insert_zero_bits(t, p.outer) << p.lo) | p.hi_ctrl_mask);
          for (uint32_t j = threadIdx.x; j < tile_amps; j += THREADS)
               cp_async16(&tile[swizzle<K>(j, p)], base + (j & lo_mask) + p.hoff[j >> p.lo]);
          cp_async_wait_all();
          __syncthreads();
          for (uint32_t u = threadIdx.x; u < (tile_amps >> K); u += THREADS) {
               const uint32_t lb = static_cast<uint32_t>(insert_zero_bits(u, p.inner));
               if ((lb & p.lo_ctrl_mask) != p.lo_ctrl_mask) continue;
               const uint32_t pb = swizzle<K>(lb, p);  // swizzle sources are non-target bits
               double2 in[1 << K];
#pragma unroll
               for (int c = 0; c < (1 << K); ++c) in[c] = tile[pb ^ p.loff[c]];
               apply_rows<K>(in, p.m, [&](int b, double2 v) { tile[pb ^ p.loff[b]] = v; });
          }
          __syncthreads();
          for (uint32_t j = threadIdx.x; j < tile_amps; j += THREADS)
               base[(j & lo_mask) + p.hoff[j >> p.lo]] = tile[swizzle<K>(j, p)];
          __syncthreads();
     }
}

// ---------------------------------------------------------------------------------------------
// DMMA (FP64 tensor cores), K = 2..5
// ---------------------------------------------------------------------------------------------
template <int K>
struct DmmaParams {
     double2* psi;
     uint64_t n_groups;  // groups of 8 consecutive free indices
     uint64_t ctrl_mask;
     InsertBits ins;
     uint64_t off[1 << K];
     double2 m[1 << (2 * K)];
};

__device__ __forceinline__ void dmma884(double& d0, double& d1, double a, double b)
{
     asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                  : "+d"(d0), "+d"(d1)
                  : "d"(a), "d"(b));
}

template <int K, int THREADS, int MINB, int G>
__global__ void __launch_bounds__(THREADS, MINB) dense_dmma_kernel(const __grid_constant__ DmmaParams<K> p)
{
     constexpr int D = 1 << K;       // complex dimension
     constexpr int KT = 2 * D / 4;   // k-tiles of the real 2D x 2D product
     constexpr int NT = 2 * D / 8;   // n-tiles
     // B fragments of Mreal^T: Bs[(kt * NT + nt) * 32 + lane] = Mreal[8 nt + lane / 4][4 kt + lane % 4]
     extern __shared__ double bs[];
     __shared__ uint64_t s_off[D];
     for (int i = threadIdx.x; i < KT * NT * 32; i += THREADS) {
          const int lane = i & 31, tilei = i >> 5;
          const int kt = tilei / NT, nt = tilei % NT;
          const int r = 8 * nt + (lane >> 2), q = 4 * kt + (lane & 3);
          const double2 e = p.m[(r >> 1) * D + (q >> 1)];
          double v;
          if ((r & 1) == (q & 1)) v = e.x;
          else v = (r & 1) ? e.y : -e.y;
          bs[i] = v;
     }
     if (threadIdx.x < D) s_off[threadIdx.x] = p.off[threadIdx.x];
     __syncthreads();

     const int lane = threadIdx.x & 31;
     const int t = lane >> 2, j = lane & 3;
     uint64_t ld_off[KT];  // byte offsets of this lane's A-fragment element per k-tile
     uint64_t st_off[NT];  // byte offsets of this lane's accumulator amplitude per n-tile
#pragma unroll
     for (int kt = 0; kt < KT; ++kt) ld_off[kt] = s_off[2 * kt + (j >> 1)] * 16 + (j & 1) * 8;
#pragma unroll
     for (int nt = 0; nt < NT; ++nt) st_off[nt] = s_off[4 * nt + j] * 16;

     const uint64_t warps = (static_cast<uint64_t>(gridDim.x) * THREADS) >> 5;
     const uint64_t warp0 = (static_cast<uint64_t>(blockIdx.x) * THREADS + threadIdx.x) >> 5;
     for (uint64_t g0 = warp0 * G; g0 < p.n_groups; g0 += warps * G) {
          char* base[G];
          double a[G][KT];
#pragma unroll
          for (int g = 0; g < G; ++g) {
               const uint64_t grp = (g0 + g < p.n_groups) ? g0 + g : g0;  // tail: recompute group g0 (idempotent reads)
               base[g] = reinterpret_cast<char*>(p.psi + (insert_zero_bits(grp * 8 + t, p.ins) | p.ctrl_mask));
#pragma unroll
               for (int kt = 0; kt < KT; ++kt)
                    a[g][kt] = ldg_stream_f64(reinterpret_cast<const double*>(base[g] + ld_off[kt]));
          }
          double d[G][NT][2];
#pragma unroll
          for (int g = 0; g < G; ++g)
#pragma unroll
               for (int nt = 0; nt < NT; ++nt) d[g][nt][0] = d[g][nt][1] = 0.0;
#pragma unroll
          for (int kt = 0; kt < KT; ++kt)
#pragma unroll
               for (int nt = 0; nt < NT; ++nt) {
                    const double b = bs[(kt * NT + nt) * 32 + lane];
#pragma unroll
                    for (int g = 0; g < G; ++g) dmma884(d[g][nt][0], d[g][nt][1], a[g][kt], b);
               }
          // every lane of the warp has finished loading before any lane stores (mma.sync is warp-wide)
#pragma unroll
          for (int g = 0; g < G; ++g) {
               if (g0 + g < p.n_groups) {
#pragma unroll
                    for (int nt = 0; nt < NT; ++nt)
                         *reinterpret_cast<double2*>(base[g] + st_off[nt]) = make_double2(d[g][nt][0], d[g][nt][1]);
               }
          }
     }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int tile_bits_for(int k, int L)
{
     int tb = std::max(k + 7, 10);
     return std::min(tb, L);
}

template <int K>
static void fill_common(uint64_t (&off)[1 << K], double2 (&m)[1 << (2 * K)], const int* slots, const double* matrix)
{
     for (int c = 0; c < (1 << K); ++c) {
          uint64_t o = 0;
          for (int l = 0; l < K; ++l)
               if ((c >> l) & 1) o |= 1ull << slots[l];
          off[c] = o;
     }
     std::memcpy(m, matrix, sizeof(double2) << (2 * K));
}

template <int K>
static void fill_msum(DirectParams<K>& p)
{
     if constexpr (K == 4)
          for (int i = 0; i < (1 << (2 * K)); ++i) p.msum[i] = p.m[i].x + p.m[i].y;
     else p.msum[0] = 0.0;
}

// HIQ_DENSE_3M=0 in the environment keeps the full k = 4 product on four multiplications (A/B measurements)
static bool dense_3m_enabled()
{
     static const bool on = [] {
          const char* e = std::getenv("HIQ_DENSE_3M");
          return !(e && e[0] == '0');
     }();
     return on;
}

static InsertBits make_insert_bits(const int* slots, int k, uint64_t ctrl_mask, int shift = 0, int min_pos = 0)
{
     std::vector<int> pos;
     for (int l = 0; l < k; ++l)
          if (slots[l] >= min_pos) pos.push_back(slots[l] - shift);
     for (int s = min_pos; s < 64; ++s)
          if ((ctrl_mask >> s) & 1) pos.push_back(s - shift);
     std::sort(pos.begin(), pos.end());
     InsertBits ib;
     std::memset(&ib, 0, sizeof(ib));
     ib.n = static_cast<int>(pos.size());
     for (int i = 0; i < ib.n; ++i) ib.pos[i] = static_cast<uint8_t>(pos[i]);
     return ib;
}

// Block structure of a gate matrix.  Index bit l is a SELECT bit when m[b][c] == 0 for every pair b, c that
// differ in bit l: the qubit only chooses which block acts on the other ("mixing") qubits.  shape.order lists the
// mixing bits first (ascending), then the select bits; shape.ks = number of mixing bits (>= 1).
struct DenseShape {
     int ks;
     int order[kMaxTargets];
};

static DenseShape dense_shape(int k, const double* matrix)
{
     const int d = 1 << k;
     uint32_t mixes = 0;  // bit l set: some nonzero entry couples indices that differ in bit l
     for (int b = 0; b < d; ++b)
          for (int c = 0; c < d; ++c)
               if (b != c && (matrix[2 * (b * d + c)] != 0.0 || matrix[2 * (b * d + c) + 1] != 0.0)) mixes |= static_cast<uint32_t>(b ^ c);
     DenseShape sh;
     sh.ks = 0;
     for (int l = 0; l < k; ++l)
          if ((mixes >> l) & 1u) sh.order[sh.ks++] = l;
     int n = sh.ks;
     for (int l = 0; l < k; ++l)
          if (!((mixes >> l) & 1u)) sh.order[n++] = l;
     if (sh.ks == 0) sh.ks = 1;  // a diagonal matrix: any bit may play the mixing one
     return sh;
}

// slots and matrix re-expressed with index bit i' = old bit order[i']
static void permute_gate(int k, const int* order, const int* slots, const double* matrix, int* slots_out, double* matrix_out)
{
     const int d = 1 << k;
     int map[1 << kMaxTargets];  // new index -> old index
     for (int x = 0; x < d; ++x) {
          int o = 0;
          for (int i = 0; i < k; ++i)
               if ((x >> i) & 1) o |= 1 << order[i];
          map[x] = o;
     }
     for (int i = 0; i < k; ++i) slots_out[i] = slots[order[i]];
     for (int b = 0; b < d; ++b)
          for (int c = 0; c < d; ++c) {
               matrix_out[2 * (b * d + c)] = matrix[2 * (map[b] * d + map[c])];
               matrix_out[2 * (b * d + c) + 1] = matrix[2 * (map[b] * d + map[c]) + 1];
          }
}

// calls f(std::integral_constant<int, KS>) for the run-time ks in 1..K (K <= 4; K = 5 has the full form only)
template <int K, class F>
static int with_ks(int ks, F&& f)
{
     if constexpr (K >= 5) return f(std::integral_constant<int, K>{});
     else {
          if constexpr (K >= 2) if (ks == 1) return f(std::integral_constant<int, 1>{});
          if constexpr (K >= 3) if (ks == 2) return f(std::integral_constant<int, 2>{});
          if constexpr (K >= 4) if (ks == 3) return f(std::integral_constant<int, 3>{});
          return f(std::integral_constant<int, K>{});
     }
}

template <int K>
static void fill_direct(DirectParams<K>& p, double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask)
{
     p.psi = psi;
     const int nc = __builtin_popcountll(ctrl_mask);
     p.n_free = 1ull << (L - K - nc);
     p.ctrl_mask = ctrl_mask;
     p.ins = make_insert_bits(slots, K, ctrl_mask);
     fill_common<K>(p.off, p.m, slots, matrix);
     fill_msum<K>(p);
}

template <int K>
static int launch_direct(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask,
                         cudaStream_t stream, int ks = K)
{
     DirectParams<K> p;
     fill_direct<K>(p, psi, L, slots, matrix, ctrl_mask);
     constexpr int THREADS = (K >= 4) ? 128 : 256;
     constexpr int MINB = (K >= 5) ? 2 : (K == 4 ? 4 : 4);
     const uint64_t need = (p.n_free + THREADS - 1) / THREADS;
     const uint64_t cap = grid_cap(static_cast<uint64_t>(num_sms()) * MINB * 8);
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(need, cap));
     auto go = [&](auto ks_c, auto m3_c) {
          constexpr int KS = decltype(ks_c)::value;
          constexpr bool M3 = decltype(m3_c)::value;
          if constexpr (K >= 3 && K <= 4) {
               // register-limited shapes: stage the next tuple through shared memory (cp.async)
               if (need > cap) {
                    static bool attr_set = false;
                    if (!attr_set) {
                         cudaFuncSetAttribute(dense_direct_staged_kernel<K, THREADS, MINB, KS, M3>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                              64 * 1024);
                         cudaFuncSetAttribute(dense_direct_staged_kernel<K, THREADS, MINB, KS, M3>,
                                              cudaFuncAttributePreferredSharedMemoryCarveout, 100);
                         attr_set = true;
                    }
                    dense_direct_staged_kernel<K, THREADS, MINB, KS, M3><<<grid, THREADS, sizeof(double2) * THREADS << K, stream>>>(p);
                    count_launch();
                    return check_launch("dense_direct_staged_kernel");
               }
          }
          dense_direct_kernel<K, THREADS, MINB, KS, M3><<<grid, THREADS, 0, stream>>>(p);
          count_launch();
          return check_launch("dense_direct_kernel");
     };
     if constexpr (K == 4)
          if (ks == K && dense_3m_enabled()) return go(std::integral_constant<int, K>{}, std::true_type{});
     return with_ks<K>(ks, [&](auto ks_c) { return go(ks_c, std::false_type{}); });
}

// Host part of the folded-diagonal launch: index split (tid | t | chunk), diagonal program, class-E pattern tables.
template <int K>
struct DirectPreShape {
     static constexpr int THREADS = (K >= 3) ? 128 : 256;
     static constexpr int TID_BITS = (K >= 3) ? 7 : 8;
};

template <int K>
static int fill_direct_pre(DirectPreParams<K>& p, double2* psi, int L, const int* slots, const double* matrix, const hiqk_diag_op* pre,
                           int n_pre)
{
     constexpr int TID_BITS = DirectPreShape<K>::TID_BITS;
     uint64_t tmask = 0;
     for (int t = 0; t < K; ++t) tmask |= 1ull << slots[t];
     // tid occupies the TID_BITS lowest free index positions; t positions are picked among the rest
     uint64_t tid_mask = 0;
     for (int pos = 0, got = 0; pos < L && got < TID_BITS; ++pos)
          if (!((tmask >> pos) & 1ull)) {
               tid_mask |= 1ull << pos;
               ++got;
          }
     int tpos[kMaxTBits];
     const int n_t = choose_u_positions(L, pre, n_pre, tmask | tid_mask, kMaxTBits, tpos);
     int order[kMaxDiagOps];
     const int rc = build_diag_prog(p.prog, L, pre, n_pre, tpos, n_t, tmask, tid_mask, order, "hiqk_apply_dense_prediag");
     if (rc != HIQ_OK) return rc;
     p.n_t = n_t;
     for (int t = 0; t < (1 << kMaxTBits); ++t) {
          uint64_t o = 0;
          for (int b = 0; b < n_t; ++b)
               if ((t >> b) & 1) o |= 1ull << tpos[b];
          p.toff[t] = o;
     }
     // free-index deposit skips the targets and the t positions
     {
          std::vector<int> skip(slots, slots + K);
          skip.insert(skip.end(), tpos, tpos + n_t);
          p.d.ins = make_insert_bits(skip.data(), static_cast<int>(skip.size()), 0);
     }
     p.d.psi = psi;
     p.d.n_free = 1ull << (L - K - n_t);
     p.d.ctrl_mask = 0;
     fill_common<K>(p.d.off, p.d.m, slots, matrix);
     fill_msum<K>(p.d);
     // class-E tables: the selector bits every tuple element contributes to each op, grouped into the
     // distinct joint patterns (elements with the same bits for all class-E ops share one factor)
     std::memset(p.e_pat, 0, sizeof(p.e_pat));
     std::memset(p.e_cmap, 0, sizeof(p.e_cmap));
     const int je = p.prog.n_s0 + p.prog.n_s1;
     std::vector<std::vector<uint32_t>> pats;  // pats[e][j - je]
     for (int c = 0; c < (1 << K); ++c) {
          std::vector<uint32_t> sel(p.prog.n - je, 0);
          for (int j = je; j < p.prog.n; ++j)
               for (int l = 0; l < kMaxTargets; ++l)
                    for (int t = 0; t < K; ++t)
                         if (p.prog.slots[j][l] == slots[t] && ((c >> t) & 1)) sel[j - je] |= 1u << l;
          auto it = std::find(pats.begin(), pats.end(), sel);
          if (it == pats.end()) {
               pats.push_back(sel);
               it = pats.end() - 1;
          }
          p.e_cmap[c] = static_cast<uint8_t>(it - pats.begin());
     }
     p.e_npat = p.prog.n_e ? static_cast<int>(pats.size()) : 0;
     for (size_t e = 0; e < pats.size(); ++e)
          for (int j = je; j < p.prog.n; ++j) p.e_pat[j][e] = static_cast<uint8_t>(pats[e][j - je]);
     p.fast = p.prog.n_s1 == 0 ? 1 : 0;
     for (int j = je; j < p.prog.n; ++j)
          for (int t = 0; t < (1 << n_t); ++t)
               if (p.prog.usel[j][t]) p.fast = 0;
     return HIQ_OK;
}

template <int K>
static int launch_direct_pre(double2* psi, int L, const int* slots, const double* matrix, const hiqk_diag_op* pre, int n_pre,
                             cudaStream_t stream, int ks = K)
{
     static DirectPreParams<K> p;  // 10+ KB: keep it off the stack; launches are issued from one host thread per engine
     static std::mutex mu;
     std::lock_guard<std::mutex> lock(mu);
     constexpr int THREADS = DirectPreShape<K>::THREADS;
     constexpr int MINB = 4;
     const int rc = fill_direct_pre<K>(p, psi, L, slots, matrix, pre, n_pre);
     if (rc != HIQ_OK) return rc;
     const size_t smem = sizeof(double2) * THREADS * ((1u << K) + std::max(p.e_npat, 1));
     const uint64_t n_chunks = (p.d.n_free + THREADS - 1) / THREADS;
     const uint64_t cap = grid_cap(static_cast<uint64_t>(num_sms()) * MINB * 8);
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n_chunks, cap));
     auto go = [&](auto ks_c, auto m3_c) {
          constexpr int KS = decltype(ks_c)::value;
          constexpr bool M3 = decltype(m3_c)::value;
          if constexpr (KS < K) {
               // block-structured gate: register-pipelined block form — no tuple staging, more resident warps (measured on
               // B200, profiles/r02a_prediag_{staged,blockloop}_L30.jsonl: 6.46 -> 5.84 ms for the QFT-like launch at L = 30)
               constexpr int MINB2 = (K >= 3) ? (KS >= 3 ? 4 : 6) : 4;
               const size_t smem2 = sizeof(double2) * THREADS * std::max(p.e_npat, 1);
               const uint64_t cap2 = grid_cap(static_cast<uint64_t>(num_sms()) * MINB2 * 8);
               const unsigned grid2 = static_cast<unsigned>(std::min<uint64_t>(n_chunks, cap2));
               dense_direct_pre_blocks_kernel<K, KS, THREADS, MINB2><<<grid2, THREADS, smem2, stream>>>(p);
               count_launch();
               return check_launch("dense_direct_pre_blocks_kernel");
          }
          else {
          static bool attr_set = false;
          if (!attr_set) {
               cudaFuncSetAttribute(dense_direct_pre_kernel<K, THREADS, MINB, KS, M3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
               cudaFuncSetAttribute(dense_direct_pre_kernel<K, THREADS, MINB, KS, M3>, cudaFuncAttributePreferredSharedMemoryCarveout, 100);
               attr_set = true;
          }
          dense_direct_pre_kernel<K, THREADS, MINB, KS, M3><<<grid, THREADS, smem, stream>>>(p);
          count_launch();
          return check_launch("dense_direct_pre_kernel");
          }
     };
     if constexpr (K == 4)
          if (ks == K && dense_3m_enabled()) return go(std::integral_constant<int, K>{}, std::true_type{});
     return with_ks<K>(ks, [&](auto ks_c) { return go(ks_c, std::false_type{}); });
}

template <int K>
static int fill_tiled(TiledParams<K>& p, double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask)
{
     std::memset(&p, 0, sizeof(p));
     const int tb = tile_bits_for(K, L);
     if (tb < K + 3) return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: slab too small for the tiled kernel");
     // lo + #(targets >= lo) == tb
     int lo = tb;
     auto n_hi = [&](int l) { int n = 0; for (int i = 0; i < K; ++i) n += slots[i] >= l; return n; };
     while (lo + n_hi(lo) > tb) --lo;
     p.psi = psi;
     p.lo = lo;
     p.tile_bits = tb;
     std::vector<int> hi;  // high targets ascending
     for (int i = 0; i < K; ++i) if (slots[i] >= lo) hi.push_back(slots[i]);
     std::sort(hi.begin(), hi.end());
     const int nh = static_cast<int>(hi.size());
     for (int h = 0; h < (1 << nh); ++h) {
          uint64_t o = 0;
          for (int b = 0; b < nh; ++b) if ((h >> b) & 1) o |= 1ull << hi[b];
          p.hoff[h] = o;
     }
     // tile-local position of every target
     int lpos[kMaxTargets];
     for (int l = 0; l < K; ++l) {
          if (slots[l] < lo) lpos[l] = slots[l];
          else lpos[l] = lo + static_cast<int>(std::find(hi.begin(), hi.end(), slots[l]) - hi.begin());
     }
     for (int c = 0; c < (1 << K); ++c) {
          uint32_t o = 0;
          for (int l = 0; l < K; ++l) if ((c >> l) & 1) o |= 1u << lpos[l];
          p.loff[c] = o;
     }
     p.inner = make_insert_bits(lpos, K, 0);
     const uint64_t lo_mask = (1ull << lo) - 1ull;
     p.lo_ctrl_mask = static_cast<uint32_t>(ctrl_mask & lo_mask);
     p.hi_ctrl_mask = ctrl_mask & ~lo_mask;
     p.outer = make_insert_bits(hi.data(), nh, p.hi_ctrl_mask, lo, lo);
     const int n_hi_ctrl = __builtin_popcountll(p.hi_ctrl_mask);
     p.n_tiles = 1ull << (L - tb - n_hi_ctrl);
     // swizzle: the three lowest non-target tile bits must land on distinct 16-byte bank groups
     uint32_t tmask = 0;
     for (int l = 0; l < K; ++l) tmask |= 1u << lpos[l];
     std::vector<int> dst;  // low-3 positions occupied by targets
     for (int b = 0; b < 3; ++b) if ((tmask >> b) & 1) dst.push_back(b);
     int found = 0;
     p.nswz = 0;
     for (int b = 0; b < tb && found < 3; ++b) {
          if ((tmask >> b) & 1) continue;
          ++found;
          if (b >= 3) {
               p.swz_src[p.nswz] = b;
               p.swz_dst[p.nswz] = dst[p.nswz];
               ++p.nswz;
          }
     }
     std::memcpy(p.m, matrix, sizeof(double2) << (2 * K));
     return HIQ_OK;
}

template <int K>
static int launch_tiled(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask,
                        cudaStream_t stream)
{
     TiledParams<K> p;
     const int rc = fill_tiled<K>(p, psi, L, slots, matrix, ctrl_mask);
     if (rc != HIQ_OK) return rc;
     const int tb = p.tile_bits;
     constexpr int THREADS = 128;
     constexpr int MINB = (K >= 5) ? 2 : 4;
     const size_t smem = sizeof(double2) << tb;
     static bool attr_set = false;
     if (!attr_set) {
          cudaFuncSetAttribute(dense_tiled_kernel<K, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
          attr_set = true;
     }
     const uint64_t cap = grid_cap(static_cast<uint64_t>(num_sms()) * MINB * 4);
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n_tiles, cap));
     dense_tiled_kernel<K, THREADS, MINB><<<grid, THREADS, smem, stream>>>(p);
     count_launch();
     return check_launch("dense_tiled_kernel");
}

// the tensor-core kernel works on groups of 8 consecutive free indices
static bool dmma_fits(int L, int k, uint64_t ctrl_mask) { return (1ull << (L - k - __builtin_popcountll(ctrl_mask))) >= 8; }

template <int K>
static void fill_dmma(DmmaParams<K>& p, double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask)
{
     const int nc = __builtin_popcountll(ctrl_mask);
     const uint64_t n_free = 1ull << (L - K - nc);
     p.psi = psi;
     p.n_groups = n_free >> 3;
     p.ctrl_mask = ctrl_mask;
     p.ins = make_insert_bits(slots, K, ctrl_mask);
     fill_common<K>(p.off, p.m, slots, matrix);
}

template <int K>
static int launch_dmma(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask,
                       cudaStream_t stream)
{
     DmmaParams<K> p;
     if (!dmma_fits(L, K, ctrl_mask)) return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
     fill_dmma<K>(p, psi, L, slots, matrix, ctrl_mask);
     constexpr int THREADS = 128;
     constexpr int G = (K >= 5) ? 1 : 2;
     constexpr int MINB = (K >= 5) ? 3 : 4;
     constexpr int D = 1 << K;
     const size_t smem = sizeof(double) * (2 * D / 4) * (2 * D / 8) * 32;
     static bool attr_set = false;
     if (!attr_set) {
          cudaFuncSetAttribute(dense_dmma_kernel<K, THREADS, MINB, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
          attr_set = true;
     }
     const uint64_t groups_per_block = static_cast<uint64_t>(THREADS / 32) * G;
     const uint64_t need = (p.n_groups + groups_per_block - 1) / groups_per_block;
     const uint64_t cap = grid_cap(static_cast<uint64_t>(num_sms()) * MINB * 2);
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(need, cap));
     dense_dmma_kernel<K, THREADS, MINB, G><<<grid, THREADS, smem, stream>>>(p);
     count_launch();
     return check_launch("dense_dmma_kernel");
}

// HIQ_DENSE_BLOCKS=0 in the environment keeps every DIRECT launch on the full product (A/B measurements)
static bool dense_blocks_enabled()
{
     static const bool on = [] {
          const char* e = std::getenv("HIQ_DENSE_BLOCKS");
          return !(e && e[0] == '0');
     }();
     return on;
}

// Kernel choice measured on B200 under sustained load (profiles/r01k_sustained_L30.jsonl; a long circuit runs
// power-capped, the burst figures of profiles/r01_sweep_a_L30.jsonl rank some shapes differently): the 32x32
// product is FP64-bound and fastest on the tensor cores; k <= 4 is HBM-bound and fastest with one tuple per
// thread as long as every warp access covers whole 32 B sectors (lowest target slot >= 1, >= 2 for k = 2);
// a target in slot 0 takes the shared-memory tile (k <= 3) or, for k = 4, the tensor-core kernel, whose
// A fragments are loaded sector-complete whatever the slots are.
static int pick_variant(int L, int k, const int* slots)
{
     int min_slot = 64;
     for (int l = 0; l < k; ++l) min_slot = std::min(min_slot, slots[l]);
     const bool can_tile = tile_bits_for(k, L) >= k + 3 && L >= 10;
     const bool can_dmma = L - k >= 3;
     if (k == 5) return can_dmma ? HIQK_DENSE_DMMA : HIQK_DENSE_DIRECT;
     if (k == 4) {
          if (min_slot >= 1) return HIQK_DENSE_DIRECT;
          return can_dmma ? HIQK_DENSE_DMMA : (can_tile ? HIQK_DENSE_TILED : HIQK_DENSE_DIRECT);
     }
     if (k == 2) return (min_slot < 2 && can_tile) ? HIQK_DENSE_TILED : HIQK_DENSE_DIRECT;
     return (min_slot < 1 && can_tile) ? HIQK_DENSE_TILED : HIQK_DENSE_DIRECT;
}

template <int K>
static int dispatch_k(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask, int variant,
                      cudaStream_t stream)
{
     if (variant == HIQK_DENSE_AUTO) variant = pick_variant(L, K, slots);
     switch (variant) {
          case HIQK_DENSE_DIRECT:
               if constexpr (K >= 2 && K <= 4) {
                    if (dense_blocks_enabled()) {
                         const DenseShape sh = dense_shape(K, matrix);
                         if (sh.ks < K) {
                              int pslots[K];
                              double pm[2 << (2 * K)];
                              permute_gate(K, sh.order, slots, matrix, pslots, pm);
                              return launch_direct<K>(psi, L, pslots, pm, ctrl_mask, stream, sh.ks);
                         }
                    }
               }
               return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
          case HIQK_DENSE_DIRECT_FULL: return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
          case HIQK_DENSE_TILED: return launch_tiled<K>(psi, L, slots, matrix, ctrl_mask, stream);
          case HIQK_DENSE_DMMA:
               if constexpr (K >= 2) return launch_dmma<K>(psi, L, slots, matrix, ctrl_mask, stream);
               else return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
          default: return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: unknown variant");
     }
}

// Host-only parameter image of hiqk_apply_dense (see include/hiq_b200.h): the same variant resolution and block-shape
// permutation as dispatch_k, the same fill functions as the launchers.  Header words: magic, variant, K, mixing bits used
// (DIRECT), three-multiplication flag, sizeof(parameters), sizeof(InsertBits), then the field offsets of the variant:
//   DIRECT: n_free, ctrl_mask, ins, off, m, msum        DMMA: n_groups, ctrl_mask, ins, off, m
//   TILED:  n_tiles, hi_ctrl_mask, lo_ctrl_mask, lo, tile_bits, nswz, swz_src, swz_dst, outer, inner, hoff, loff, m
constexpr int kDenseImageHeaderWords = 32;

template <int K>
static int dense_image_k(int L, const int* slots, const double* matrix, uint64_t ctrl_mask, int variant, void* image)
{
     uint32_t* head = static_cast<uint32_t*>(image);
     void* body = head + kDenseImageHeaderWords;
     if (variant == HIQK_DENSE_AUTO) variant = pick_variant(L, K, slots);
     if (variant == HIQK_DENSE_DMMA && (K < 2 || !dmma_fits(L, K, ctrl_mask))) variant = HIQK_DENSE_DIRECT_FULL;  // launch_dmma's fallback
     int w = 0;
     head[w++] = 0x4e445148u;  // 'HQDN'
     int ks = K, m3 = 0;
     switch (variant) {
          case HIQK_DENSE_DIRECT:
          case HIQK_DENSE_DIRECT_FULL: {
               using P = DirectParams<K>;
               P* p = static_cast<P*>(body);
               int pslots[K];
               double pm[2 << (2 * K)];
               const int* use_slots = slots;
               const double* use_m = matrix;
               if constexpr (K >= 2 && K <= 4) {
                    if (variant == HIQK_DENSE_DIRECT && dense_blocks_enabled()) {
                         const DenseShape sh = dense_shape(K, matrix);
                         if (sh.ks < K) {
                              permute_gate(K, sh.order, slots, matrix, pslots, pm);
                              use_slots = pslots;
                              use_m = pm;
                              ks = sh.ks;
                         }
                    }
               }
               fill_direct<K>(*p, nullptr, L, use_slots, use_m, ctrl_mask);
               if (K == 4 && ks == K && dense_3m_enabled()) m3 = 1;
               head[w++] = HIQK_DENSE_DIRECT;
               head[w++] = K;
               head[w++] = static_cast<uint32_t>(ks);
               head[w++] = static_cast<uint32_t>(m3);
               head[w++] = sizeof(P);
               head[w++] = sizeof(InsertBits);
               head[w++] = static_cast<uint32_t>(offsetof(P, n_free));
               head[w++] = static_cast<uint32_t>(offsetof(P, ctrl_mask));
               head[w++] = static_cast<uint32_t>(offsetof(P, ins));
               head[w++] = static_cast<uint32_t>(offsetof(P, off));
               head[w++] = static_cast<uint32_t>(offsetof(P, m));
               head[w++] = static_cast<uint32_t>(offsetof(P, msum));
               return HIQ_OK;
          }
          case HIQK_DENSE_TILED: {
               using P = TiledParams<K>;
               P* p = static_cast<P*>(body);
               const int rc = fill_tiled<K>(*p, nullptr, L, slots, matrix, ctrl_mask);
               if (rc != HIQ_OK) return rc;
               head[w++] = HIQK_DENSE_TILED;
               head[w++] = K;
               head[w++] = K;
               head[w++] = 0;
               head[w++] = sizeof(P);
               head[w++] = sizeof(InsertBits);
               head[w++] = static_cast<uint32_t>(offsetof(P, n_tiles));
               head[w++] = static_cast<uint32_t>(offsetof(P, hi_ctrl_mask));
               head[w++] = static_cast<uint32_t>(offsetof(P, lo_ctrl_mask));
               head[w++] = static_cast<uint32_t>(offsetof(P, lo));
               head[w++] = static_cast<uint32_t>(offsetof(P, tile_bits));
               head[w++] = static_cast<uint32_t>(offsetof(P, nswz));
               head[w++] = static_cast<uint32_t>(offsetof(P, swz_src));
               head[w++] = static_cast<uint32_t>(offsetof(P, swz_dst));
               head[w++] = static_cast<uint32_t>(offsetof(P, outer));
               head[w++] = static_cast<uint32_t>(offsetof(P, inner));
               head[w++] = static_cast<uint32_t>(offsetof(P, hoff));
               head[w++] = static_cast<uint32_t>(offsetof(P, loff));
               head[w++] = static_cast<uint32_t>(offsetof(P, m));
               return HIQ_OK;
          }
          case HIQK_DENSE_DMMA: {
               if constexpr (K >= 2) {
                    using P = DmmaParams<K>;
                    P* p = static_cast<P*>(body);
                    fill_dmma<K>(*p, nullptr, L, slots, matrix, ctrl_mask);
                    head[w++] = HIQK_DENSE_DMMA;
                    head[w++] = K;
                    head[w++] = K;
                    head[w++] = 0;
                    head[w++] = sizeof(P);
                    head[w++] = sizeof(InsertBits);
                    head[w++] = static_cast<uint32_t>(offsetof(P, n_groups));
                    head[w++] = static_cast<uint32_t>(offsetof(P, ctrl_mask));
                    head[w++] = static_cast<uint32_t>(offsetof(P, ins));
                    head[w++] = static_cast<uint32_t>(offsetof(P, off));
                    head[w++] = static_cast<uint32_t>(offsetof(P, m));
                    return HIQ_OK;
               }
               return set_error(HIQ_ERR_ARG, "hiqk_dense_image: the tensor-core kernel needs k >= 2");
          }
          default: return set_error(HIQ_ERR_ARG, "hiqk_dense_image: unknown variant");
     }
}

constexpr size_t kDenseImageBodyBytes =
     sizeof(DirectParams<5>) > sizeof(TiledParams<5>) ? (sizeof(DirectParams<5>) > sizeof(DmmaParams<5>) ? sizeof(DirectParams<5>) : sizeof(DmmaParams<5>))
                                                      : (sizeof(TiledParams<5>) > sizeof(DmmaParams<5>) ? sizeof(TiledParams<5>) : sizeof(DmmaParams<5>));

}  // namespace hiq

extern "C" size_t hiqk_dense_image_bytes(void) { return hiq::kDenseImageHeaderWords * sizeof(uint32_t) + hiq::kDenseImageBodyBytes; }

extern "C" int hiqk_dense_image(int L, int k, const int* slots, const double* matrix, uint64_t ctrl_mask, int variant, void* image,
                                size_t image_bytes)
{
     using namespace hiq;
     if (!slots || !matrix || !image) return set_error(HIQ_ERR_ARG, "hiqk_dense_image: null argument");
     if (image_bytes < hiqk_dense_image_bytes()) return set_error(HIQ_ERR_ARG, "hiqk_dense_image: buffer too small");
     if (k < 1 || k > kMaxTargets) return set_error(HIQ_ERR_ARG, "hiqk_dense_image: k must be 1..5");
     if (L < k || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_dense_image: bad slab size");
     uint64_t tmask = 0;
     for (int l = 0; l < k; ++l) {
          if (slots[l] < 0 || slots[l] >= L || ((tmask >> slots[l]) & 1))
               return set_error(HIQ_ERR_ARG, "hiqk_dense_image: target slots must be distinct and < L");
          tmask |= 1ull << slots[l];
     }
     if ((ctrl_mask & tmask) || (L < 64 && (ctrl_mask >> L)))
          return set_error(HIQ_ERR_ARG, "hiqk_dense_image: control mask overlaps targets or exceeds the slab");
     std::memset(image, 0, hiqk_dense_image_bytes());
     switch (k) {
          case 1: return dense_image_k<1>(L, slots, matrix, ctrl_mask, variant, image);
          case 2: return dense_image_k<2>(L, slots, matrix, ctrl_mask, variant, image);
          case 3: return dense_image_k<3>(L, slots, matrix, ctrl_mask, variant, image);
          case 4: return dense_image_k<4>(L, slots, matrix, ctrl_mask, variant, image);
          default: return dense_image_k<5>(L, slots, matrix, ctrl_mask, variant, image);
     }
}

extern "C" int hiqk_dense_pick_variant(int L, int k, const int* slots)
{
     if (!slots || k < 1 || k > hiq::kMaxTargets) return HIQK_DENSE_DIRECT;
     return hiq::pick_variant(L, k, slots);
}

extern "C" int hiqk_dense_block_shape(int k, const double* matrix, int* order)
{
     if (!matrix || k < 1 || k > hiq::kMaxTargets) return hiq::set_error(HIQ_ERR_ARG, "hiqk_dense_block_shape: bad argument"), -1;
     const hiq::DenseShape sh = hiq::dense_shape(k, matrix);
     if (order)
          for (int i = 0; i < k; ++i) order[i] = sh.order[i];
     return sh.ks;
}

extern "C" int hiqk_dense_direct_mixing_bits(int k, const double* matrix)
{
     if (!matrix || k < 1 || k > hiq::kMaxTargets) return hiq::set_error(HIQ_ERR_ARG, "hiqk_dense_direct_mixing_bits: bad argument"), -1;
     if (k < 2 || k > 4 || !hiq::dense_blocks_enabled()) return k;
     return hiq::dense_shape(k, matrix).ks;
}

extern "C" int hiqk_dense_prediag_supported(int L, int k, const int* slots)
{
     if (!slots || k < 1 || k > 4 || L < k || L > 40) return 0;
     return hiq::pick_variant(L, k, slots) == HIQK_DENSE_DIRECT ? 1 : 0;
}

namespace hiq {
// argument checks and block-shape permutation shared by the launch and by its parameter image
static int prediag_prepare(const char* who, int L, int k, const int*& slots, const double*& matrix, int (&pslots)[kMaxTargets],
                           double (&pm)[2 << (2 * 4)], int& ks)
{
     if (!hiqk_dense_prediag_supported(L, k, slots))
          return set_error(HIQ_ERR_ARG, std::string(who) + ": needs k <= 4 and targets the DIRECT kernel takes");
     uint64_t tmask = 0;
     for (int l = 0; l < k; ++l) {
          if (slots[l] < 0 || slots[l] >= L || ((tmask >> slots[l]) & 1))
               return set_error(HIQ_ERR_ARG, std::string(who) + ": target slots must be distinct and < L");
          tmask |= 1ull << slots[l];
     }
     ks = k;
     if (k >= 2 && dense_blocks_enabled()) {
          const DenseShape sh = dense_shape(k, matrix);
          if (sh.ks < k) {
               permute_gate(k, sh.order, slots, matrix, pslots, pm);
               slots = pslots;
               matrix = pm;
               ks = sh.ks;
          }
     }
     return HIQ_OK;
}

constexpr int kPreImageHeaderWords = 64;

template <int K>
static int direct_pre_image(int L, const int* slots, const double* matrix, const hiqk_diag_op* pre, int n_pre, int ks, void* image)
{
     uint32_t* head = static_cast<uint32_t*>(image);
     DirectPreParams<K>* p = reinterpret_cast<DirectPreParams<K>*>(head + kPreImageHeaderWords);
     const int rc = fill_direct_pre<K>(*p, nullptr, L, slots, matrix, pre, n_pre);
     if (rc != HIQ_OK) return rc;
     using P = DirectPreParams<K>;
     int w = 0;
     head[w++] = 0x50445148u;  // 'HQDP'
     head[w++] = K;
     head[w++] = DirectPreShape<K>::THREADS;
     head[w++] = static_cast<uint32_t>(ks);  // mixing bits the kernel uses (block form when < K)
     head[w++] = (K == 4 && ks == K && dense_3m_enabled()) ? 1 : 0;  // three-multiplication product
     head[w++] = sizeof(P);
     head[w++] = kMaxDiagOps;
     head[w++] = 1 << kMaxUBits;
     head[w++] = 1 << kMaxTargets;
     head[w++] = 1 << kMaxTBits;
     head[w++] = sizeof(InsertBits);
     head[w++] = static_cast<uint32_t>(offsetof(P, d) + offsetof(DirectParams<K>, n_free));
     head[w++] = static_cast<uint32_t>(offsetof(P, d) + offsetof(DirectParams<K>, ins));
     head[w++] = static_cast<uint32_t>(offsetof(P, d) + offsetof(DirectParams<K>, off));
     head[w++] = static_cast<uint32_t>(offsetof(P, d) + offsetof(DirectParams<K>, m));
     head[w++] = static_cast<uint32_t>(offsetof(P, d) + offsetof(DirectParams<K>, msum));
     head[w++] = static_cast<uint32_t>(offsetof(P, n_t));
     head[w++] = static_cast<uint32_t>(offsetof(P, fast));
     head[w++] = static_cast<uint32_t>(offsetof(P, toff));
     head[w++] = static_cast<uint32_t>(offsetof(P, e_npat));
     head[w++] = static_cast<uint32_t>(offsetof(P, e_pat));
     head[w++] = static_cast<uint32_t>(offsetof(P, e_cmap));
     head[w++] = static_cast<uint32_t>(offsetof(P, prog));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, n));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, n_s0));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, n_s1));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, n_e));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, n_s0a));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, slots));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, usel));
     head[w++] = static_cast<uint32_t>(offsetof(DiagProg, lut));
     return HIQ_OK;
}
}  // namespace hiq

extern "C" size_t hiqk_dense_prediag_image_bytes(void)
{
     return hiq::kPreImageHeaderWords * sizeof(uint32_t) + sizeof(hiq::DirectPreParams<4>);
}

// Host-only: the kernel parameters hiqk_apply_dense_prediag would launch with (see include/hiq_b200.h).
extern "C" int hiqk_dense_prediag_image(int L, int k, const int* slots, const double* matrix, const hiqk_diag_op* pre, int n_pre,
                                        void* image, size_t image_bytes)
{
     using namespace hiq;
     if (!slots || !matrix || !pre || !image) return set_error(HIQ_ERR_ARG, "hiqk_dense_prediag_image: null argument");
     if (image_bytes < hiqk_dense_prediag_image_bytes()) return set_error(HIQ_ERR_ARG, "hiqk_dense_prediag_image: buffer too small");
     if (n_pre < 1 || n_pre > HIQK_MAX_DIAG_OPS) return set_error(HIQ_ERR_ARG, "hiqk_dense_prediag_image: needs 1..16 diagonal ops");
     int pslots[kMaxTargets];
     double pm[2 << (2 * 4)];
     int ks = k;
     const int rc = prediag_prepare("hiqk_dense_prediag_image", L, k, slots, matrix, pslots, pm, ks);
     if (rc != HIQ_OK) return rc;
     std::memset(image, 0, hiqk_dense_prediag_image_bytes());
     switch (k) {
          case 1: return direct_pre_image<1>(L, slots, matrix, pre, n_pre, ks, image);
          case 2: return direct_pre_image<2>(L, slots, matrix, pre, n_pre, ks, image);
          case 3: return direct_pre_image<3>(L, slots, matrix, pre, n_pre, ks, image);
          default: return direct_pre_image<4>(L, slots, matrix, pre, n_pre, ks, image);
     }
}

extern "C" int hiqk_apply_dense_prediag(void* slab, int L, int k, const int* slots, const double* matrix,
                                        const hiqk_diag_op* pre, int n_pre, void* stream)
{
     using namespace hiq;
     if (!slab || !slots || !matrix) return set_error(HIQ_ERR_ARG, "hiqk_apply_dense_prediag: null argument");
     if (n_pre == 0) return hiqk_apply_dense(slab, L, k, slots, matrix, 0, HIQK_DENSE_DIRECT, stream);
     if (n_pre < 0 || n_pre > HIQK_MAX_DIAG_OPS || !pre)
          return set_error(HIQ_ERR_ARG, "hiqk_apply_dense_prediag: needs 0..16 diagonal ops and a non-null op array");
     double2* psi = static_cast<double2*>(slab);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     int pslots[kMaxTargets];
     double pm[2 << (2 * 4)];
     int ks = k;
     const int rc = prediag_prepare("hiqk_apply_dense_prediag", L, k, slots, matrix, pslots, pm, ks);
     if (rc != HIQ_OK) return rc;
     switch (k) {
          case 1: return launch_direct_pre<1>(psi, L, slots, matrix, pre, n_pre, st);
          case 2: return launch_direct_pre<2>(psi, L, slots, matrix, pre, n_pre, st, ks);
          case 3: return launch_direct_pre<3>(psi, L, slots, matrix, pre, n_pre, st, ks);
          default: return launch_direct_pre<4>(psi, L, slots, matrix, pre, n_pre, st, ks);
     }
}

extern "C" int hiqk_apply_dense(void* slab, int L, int k, const int* slots, const double* matrix,
                                uint64_t ctrl_mask, int variant, void* stream)
{
     using namespace hiq;
     if (!slab || !slots || !matrix) return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: null argument");
     if (k < 1 || k > kMaxTargets) return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: k must be 1..5");
     if (L < k || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: bad slab size");
     uint64_t tmask = 0;
     for (int l = 0; l < k; ++l) {
          if (slots[l] < 0 || slots[l] >= L || ((tmask >> slots[l]) & 1))
               return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: target slots must be distinct and < L");
          tmask |= 1ull << slots[l];
     }
     if ((ctrl_mask & tmask) || (L < 64 && (ctrl_mask >> L)))
          return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: control mask overlaps targets or exceeds the slab");
     double2* psi = static_cast<double2*>(slab);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     switch (k) {
          case 1: return dispatch_k<1>(psi, L, slots, matrix, ctrl_mask, variant, st);
          case 2: return dispatch_k<2>(psi, L, slots, matrix, ctrl_mask, variant, st);
          case 3: return dispatch_k<3>(psi, L, slots, matrix, ctrl_mask, variant, st);
          case 4: return dispatch_k<4>(psi, L, slots, matrix, ctrl_mask, variant, st);
          default: return dispatch_k<5>(psi, L, slots, matrix, ctrl_mask, variant, st);
     }
}
