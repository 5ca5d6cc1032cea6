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
#include <cstring>
#include <mutex>
#include <vector>

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
};

template <int K>
__device__ __forceinline__ void load_tuple(double2 (&in)[1 << K], const double2* base,
                                           const uint64_t (&off)[1 << K])
{
#pragma unroll
     for (int c = 0; c < (1 << K); ++c) in[c] = ldg_stream(base + off[c]);
}

// out[b] = sum_c m[b][c] in[c]; every row is handed to `store(b, value)` as soon as it is done.
template <int K, class Store>
__device__ __forceinline__ void apply_rows(const double2 (&in)[1 << K], const double2* __restrict__ m,
                                           Store store)
{
     constexpr int D = 1 << K;
     if constexpr (K <= 4) {
#pragma unroll
          for (int b = 0; b < D; ++b) {
               double2 acc = make_double2(0.0, 0.0);
#pragma unroll
               for (int c = 0; c < D; ++c) cmac(acc, m[b * D + c], in[c]);
               store(b, acc);
          }
     }
     else {
          // 32x32: keep the row loop rolled (a full unroll is 64 KB of SASS)
#pragma unroll 2
          for (int b = 0; b < D; ++b) {
               double2 acc = make_double2(0.0, 0.0);
#pragma unroll
               for (int c = 0; c < D; ++c) cmac(acc, m[b * D + c], in[c]);
               store(b, acc);
          }
     }
}

template <int K, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dense_direct_kernel(const __grid_constant__ DirectParams<K> p)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * THREADS;
     for (uint64_t f = static_cast<uint64_t>(blockIdx.x) * THREADS + threadIdx.x; f < p.n_free; f += stride) {
          double2* base = p.psi + (insert_zero_bits(f, p.ins) | p.ctrl_mask);
          double2 in[1 << K];
          load_tuple<K>(in, base, p.off);
          apply_rows<K>(in, p.m, [&](int b, double2 v) { base[p.off[b]] = v; });
     }
}

// DIRECT with pre-diagonals: psi <- M * (prod_j D_j) * psi in one pass.  Tuples are processed in chunks
// of THREADS * T consecutive free indices (one CTA iteration): diagonal ops whose slots cannot change
// inside a chunk collapse into one factor per chunk (diag_hoist); an op with chunk-varying slots costs
// one lookup + multiply per tuple; an op that overlaps the dense targets costs one per tuple element,
// selected by (chunk bits | tuple bits | dsel[j][c]).
constexpr int kPreTuplesPerThread = 8;

template <int K>
struct DirectPreParams {
     DirectParams<K> d;
     uint32_t overlap_mask;                  // bit j: op j touches a dense target
     uint8_t dsel[kMaxDiagOps][1 << K];      // selector bits of op j contributed by tuple element c
     DiagBatch pre;
};

template <int K, int THREADS, int MINB>
__global__ void __launch_bounds__(THREADS, MINB) dense_direct_pre_kernel(const __grid_constant__ DirectPreParams<K> p)
{
     __shared__ double2 lut[kMaxDiagOps][1 << kMaxTargets];
     __shared__ DiagHoist h;
     for (int i = threadIdx.x; i < p.pre.n * (1 << kMaxTargets); i += THREADS)
          lut[i >> kMaxTargets][i & ((1 << kMaxTargets) - 1)] = p.pre.lut[i >> kMaxTargets][i & ((1 << kMaxTargets) - 1)];
     constexpr uint64_t CH = static_cast<uint64_t>(THREADS) * kPreTuplesPerThread;
     const uint64_t n_chunks = (p.d.n_free + CH - 1) / CH;
     for (uint64_t chunk = blockIdx.x; chunk < n_chunks; chunk += gridDim.x) {
          diag_hoist(p.pre, lut, insert_zero_bits(chunk * CH, p.d.ins), h);
          const double2 s_hi = h.s_hi;
#pragma unroll 1
          for (int t = 0; t < kPreTuplesPerThread; ++t) {
               const uint64_t f = chunk * CH + static_cast<uint64_t>(t) * THREADS + threadIdx.x;
               if (f >= p.d.n_free) break;
               const uint64_t bidx = insert_zero_bits(f, p.d.ins);
               double2* base = p.d.psi + bidx;
               double2 in[1 << K];
               load_tuple<K>(in, base, p.d.off);
               double2 s = s_hi;
               for (int j = 0; j < p.pre.n_lo; ++j) {
                    const uint32_t sb = h.selh[j] | diag_select_lo(p.pre.slots[j], p.pre.n_lo_slots[j], bidx);
                    if ((p.overlap_mask >> j) & 1u) {
#pragma unroll
                         for (int c = 0; c < (1 << K); ++c) in[c] = cmul(in[c], lut[j][sb | p.dsel[j][c]]);
                    }
                    else {
                         s = cmul(s, lut[j][sb]);
                    }
               }
#pragma unroll
               for (int c = 0; c < (1 << K); ++c) in[c] = cmul(in[c], s);
               apply_rows<K>(in, p.d.m, [&](int b, double2 v) { base[p.d.off[b]] = v; });
          }
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

template <int K>
static int launch_direct(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask,
                         cudaStream_t stream)
{
     DirectParams<K> p;
     p.psi = psi;
     const int nc = __builtin_popcountll(ctrl_mask);
     p.n_free = 1ull << (L - K - nc);
     p.ctrl_mask = ctrl_mask;
     p.ins = make_insert_bits(slots, K, ctrl_mask);
     fill_common<K>(p.off, p.m, slots, matrix);
     constexpr int THREADS = (K >= 4) ? 128 : 256;
     constexpr int MINB = (K >= 5) ? 2 : (K == 4 ? 4 : 4);
     const uint64_t need = (p.n_free + THREADS - 1) / THREADS;
     const uint64_t cap = static_cast<uint64_t>(kNumSMs) * MINB * 8;
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(need, cap));
     dense_direct_kernel<K, THREADS, MINB><<<grid, THREADS, 0, stream>>>(p);
     count_launch();
     return check_launch("dense_direct_kernel");
}

template <int K>
static int launch_direct_pre(double2* psi, int L, const int* slots, const double* matrix, const hiqk_diag_op* pre, int n_pre,
                             cudaStream_t stream)
{
     static DirectPreParams<K> p;  // 10+ KB: keep it off the stack; launches are issued from one host thread per engine
     static std::mutex mu;
     std::lock_guard<std::mutex> lock(mu);
     constexpr int THREADS = (K >= 3) ? 128 : 256;
     constexpr int MINB = 4;
     constexpr uint64_t CH = static_cast<uint64_t>(THREADS) * kPreTuplesPerThread;
     p.d.psi = psi;
     p.d.n_free = 1ull << (L - K);
     p.d.ctrl_mask = 0;
     p.d.ins = make_insert_bits(slots, K, 0);
     fill_common<K>(p.d.off, p.d.m, slots, matrix);
     uint64_t tmask = 0;
     for (int t = 0; t < K; ++t) tmask |= 1ull << slots[t];
     // index bits that change inside one chunk of consecutive free indices
     const uint64_t varying = insert_zero_bits(CH - 1, p.d.ins);
     const int rc = make_diag_batch(p.pre, L, pre, n_pre, varying, tmask, "hiqk_apply_dense_prediag");
     if (rc != HIQ_OK) return rc;
     p.overlap_mask = 0;
     std::memset(p.dsel, 0, sizeof(p.dsel));
     for (int j = 0; j < n_pre; ++j) {
          bool overlap = false;
          for (int c = 0; c < (1 << K); ++c) {
               uint32_t sel = 0;
               for (int l = 0; l < kMaxTargets; ++l)
                    for (int t = 0; t < K; ++t)
                         if (p.pre.slots[j][l] == slots[t] && ((c >> t) & 1)) sel |= 1u << l;
               p.dsel[j][c] = static_cast<uint8_t>(sel);
               overlap |= sel != 0;
          }
          if (overlap) p.overlap_mask |= 1u << j;
     }
     const uint64_t n_chunks = (p.d.n_free + CH - 1) / CH;
     const uint64_t cap = static_cast<uint64_t>(kNumSMs) * MINB * 4;
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(n_chunks, cap));
     dense_direct_pre_kernel<K, THREADS, MINB><<<grid, THREADS, 0, stream>>>(p);
     count_launch();
     return check_launch("dense_direct_pre_kernel");
}

template <int K>
static int launch_tiled(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask,
                        cudaStream_t stream)
{
     TiledParams<K> p;
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
     constexpr int THREADS = 128;
     constexpr int MINB = (K >= 5) ? 2 : 4;
     const size_t smem = sizeof(double2) << tb;
     static bool attr_set = false;
     if (!attr_set) {
          cudaFuncSetAttribute(dense_tiled_kernel<K, THREADS, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
          attr_set = true;
     }
     const uint64_t cap = static_cast<uint64_t>(kNumSMs) * MINB * 4;
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n_tiles, cap));
     dense_tiled_kernel<K, THREADS, MINB><<<grid, THREADS, smem, stream>>>(p);
     count_launch();
     return check_launch("dense_tiled_kernel");
}

template <int K>
static int launch_dmma(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask,
                       cudaStream_t stream)
{
     DmmaParams<K> p;
     const int nc = __builtin_popcountll(ctrl_mask);
     const uint64_t n_free = 1ull << (L - K - nc);
     if (n_free < 8) return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
     p.psi = psi;
     p.n_groups = n_free >> 3;
     p.ctrl_mask = ctrl_mask;
     p.ins = make_insert_bits(slots, K, ctrl_mask);
     fill_common<K>(p.off, p.m, slots, matrix);
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
     const uint64_t cap = static_cast<uint64_t>(kNumSMs) * MINB * 2;
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(need, cap));
     dense_dmma_kernel<K, THREADS, MINB, G><<<grid, THREADS, smem, stream>>>(p);
     count_launch();
     return check_launch("dense_dmma_kernel");
}

// Kernel choice measured on B200 (profiles/r01_sweep_a_L30.jsonl): the 32x32 product is FP64-bound and
// fastest on the tensor cores; k <= 4 is HBM-bound and fastest with one tuple per thread unless a
// target sits in the lowest slots, where the shared-memory tile keeps the accesses coalesced.
static int pick_variant(int L, int k, const int* slots)
{
     int min_slot = 64;
     for (int l = 0; l < k; ++l) min_slot = std::min(min_slot, slots[l]);
     const bool can_tile = tile_bits_for(k, L) >= k + 3 && L >= 10;
     if (k == 5) return (L - k >= 3) ? HIQK_DENSE_DMMA : HIQK_DENSE_DIRECT;
     if (k == 4) return (min_slot < 1 && can_tile) ? HIQK_DENSE_TILED : HIQK_DENSE_DIRECT;
     return (min_slot < 2 && can_tile) ? HIQK_DENSE_TILED : HIQK_DENSE_DIRECT;
}

template <int K>
static int dispatch_k(double2* psi, int L, const int* slots, const double* matrix, uint64_t ctrl_mask, int variant,
                      cudaStream_t stream)
{
     if (variant == HIQK_DENSE_AUTO) variant = pick_variant(L, K, slots);
     switch (variant) {
          case HIQK_DENSE_DIRECT: return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
          case HIQK_DENSE_TILED: return launch_tiled<K>(psi, L, slots, matrix, ctrl_mask, stream);
          case HIQK_DENSE_DMMA:
               if constexpr (K >= 2) return launch_dmma<K>(psi, L, slots, matrix, ctrl_mask, stream);
               else return launch_direct<K>(psi, L, slots, matrix, ctrl_mask, stream);
          default: return set_error(HIQ_ERR_ARG, "hiqk_apply_dense: unknown variant");
     }
}

}  // namespace hiq

extern "C" int hiqk_dense_pick_variant(int L, int k, const int* slots)
{
     if (!slots || k < 1 || k > hiq::kMaxTargets) return HIQK_DENSE_DIRECT;
     return hiq::pick_variant(L, k, slots);
}

extern "C" int hiqk_dense_prediag_supported(int L, int k, const int* slots)
{
     if (!slots || k < 1 || k > 4 || L < k) return 0;
     return hiq::pick_variant(L, k, slots) == HIQK_DENSE_DIRECT ? 1 : 0;
}

extern "C" int hiqk_apply_dense_prediag(void* slab, int L, int k, const int* slots, const double* matrix,
                                        const hiqk_diag_op* pre, int n_pre, void* stream)
{
     using namespace hiq;
     if (!slab || !slots || !matrix) return set_error(HIQ_ERR_ARG, "hiqk_apply_dense_prediag: null argument");
     if (n_pre == 0) return hiqk_apply_dense(slab, L, k, slots, matrix, 0, HIQK_DENSE_DIRECT, stream);
     if (!hiqk_dense_prediag_supported(L, k, slots))
          return set_error(HIQ_ERR_ARG, "hiqk_apply_dense_prediag: needs k <= 4 and targets the DIRECT kernel takes");
     uint64_t tmask = 0;
     for (int l = 0; l < k; ++l) {
          if (slots[l] < 0 || slots[l] >= L || ((tmask >> slots[l]) & 1))
               return set_error(HIQ_ERR_ARG, "hiqk_apply_dense_prediag: target slots must be distinct and < L");
          tmask |= 1ull << slots[l];
     }
     double2* psi = static_cast<double2*>(slab);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     switch (k) {
          case 1: return launch_direct_pre<1>(psi, L, slots, matrix, pre, n_pre, st);
          case 2: return launch_direct_pre<2>(psi, L, slots, matrix, pre, n_pre, st);
          case 3: return launch_direct_pre<3>(psi, L, slots, matrix, pre, n_pre, st);
          default: return launch_direct_pre<4>(psi, L, slots, matrix, pre, n_pre, st);
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
