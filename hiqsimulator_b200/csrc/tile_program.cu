// Tile-resident gate programs: several consecutive fused gates of the plan in ONE pass over HBM.
//
// The reference applies every fused cluster with its own sweep over the slab (reference: SimulatorMPI::Run,
// src/simulator-mpi/SimulatorMPI.cpp:441-541 -> kernels/intrin/kernel{1..4}.hpp; diagonal clusters through
// kernels_diag.hpp:35-144).  Consecutive clusters of a stage mostly act on neighbouring qubits (the stage/cluster plan of
// src/scheduler/cluster_scheduler.cpp:24-99 walks the circuit in program order), so the union of the targets of a few of
// them is small.  Here the engine hands such a run of gates to one launch: a CTA loads a tile of 2^T amplitudes —
// the `lo` lowest slots (contiguous >= 256 B runs in HBM) x every combination of the higher target slots of the run — into
// shared memory, applies the gates one after the other to the tile (a __syncthreads between two gates), and writes it back.
// HBM sees 32 B per amplitude for the whole run instead of 32 B per gate.
//
//  * Every gate is brought to a 4-target form (k < 4: extra "select" bits that the matrix does not mix) and to its block
//    structure (ks mixing bits first, dense_rows.cuh): a thread owns ONE 16-element tuple per gate, gathered from the tile
//    through an XOR swizzle of the three low address bits that the launcher solves per launch so that the gathers of every
//    gate of the run are bank-conflict free.  The matrices sit in the kernel-parameter constant bank.
//  * The diagonal fused gates of the plan that precede a gate ride along: factors that avoid the gate's targets collapse
//    into one scalar per tuple (a table lookup per op), the others ("class E") are applied per tuple element.
//    Per-thread selector bits are computed once per launch (a thread's tuples sit at the same tile positions in every
//    tile) and kept in shared memory; the part of a selector that depends on the tile is computed once per tile.
#include <algorithm>
#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <vector>

#include "dense_rows.cuh"
#include "hiq_device.cuh"
#include "hiq_host.hpp"

namespace hiq {

constexpr int kTileMaxSteps = HIQK_TILE_MAX_STEPS;
constexpr int kTileMaxOps = HIQK_TILE_MAX_OPS;  // diagonal ops per gate
constexpr int kTileLutEntries = 1024;           // table pool shared by all ops of a launch (shared memory holds what is used)
// per-thread context in shared memory, computed once per launch: the physical tile position of the thread's tuple of
// every gate (u32) and, per diagonal op, the selector bits its tuple base contributes (u8)
constexpr int kTileCtxBytes = kTileMaxSteps * (4 + kTileMaxOps);

struct TileStepDesc {
     int ks;                          // mixing bits, 1..4 (4 = full product, three-multiplication form)
     int n_ops, n_e;                  // diagonal ops applied to a tuple before the gate; the last n_e touch its targets
     int sel_off;                     // first selector-byte row of this gate's ops in shared memory
     int n_t;                         // the first n_t ops have every slot outside the tile: one factor per TILE
     int n_em;                        // ... and the last n_em of those touch MIXING bits (a factor per element; the other
                                      // class-E ops only see select bits: one factor per block of the reduced product)
     int in_mask;                     // tuple-block bits that are REAL select bits of the gate (the others pad the tuple to 16)
     int n_in, n_out;                 // select bits inside the tile / outside it (the latter choose the block per TILE)
     uint8_t out_slot[4];             // slab slots of the outside select bits, in matrix-bit order
     uint8_t mono_row[16];            // monomial form (ks == 0): the one row that column c feeds; its coefficient is m[c]
     uint8_t tpos[4];                 // tile-local target positions, ascending (bit deposit of the thread index)
     uint16_t ploff[16];              // physical (swizzled) tile offset of tuple element c
     uint16_t lut_off[kTileMaxOps];   // first table entry of op j in the pool
     uint8_t lpos[kTileMaxOps][5];    // selector bit l <- tile-local position (a non-target bit of the tile), 0xFF otherwise
     uint8_t outer[kTileMaxOps][5];   // selector bit l <- global slot outside the tile, 63 (always 0) otherwise
     uint8_t esel[kTileMaxOps][16];   // class-E ops: selector bits contributed by tuple element c
};

struct TileParams {
     double2* psi;
     uint64_t n_tiles;
     int n_steps;
     int lo;                // the tile's lowest `lo` bits are the slab's lowest slots
     InsertBits outer;      // (high tile slots - lo), ascending: tile number -> base index
     uint32_t swz_mask[3];  // physical bit r of a tile position = bit r XOR parity(position & swz_mask[r])
     uint64_t ioff[16];     // global offset spelled by the 4 highest tile bits (load / store iteration i)
     uint16_t pi[16];       // physical tile position of (i << (T - 4))
     uint8_t tslot[16];     // tile-local bit -> slab slot
     TileStepDesc step[kTileMaxSteps];
     const double2* lut;    // table pool in device memory (staged by the launcher), copied to shared memory at kernel start
     int n_lut;
     int lut_pad;           // table pool entries reserved in shared memory (n_lut rounded up)
     double2 m[kTileMaxSteps][256];
     double msum[kTileMaxSteps][256];
};
static_assert(sizeof(TileParams) <= 32764, "kernel parameters are limited to 32764 bytes");

__device__ __forceinline__ uint32_t tile_phys(uint32_t j, const uint32_t (&mask)[3])
{
     return j ^ ((__popc(j & mask[0]) & 1u) | ((__popc(j & mask[1]) & 1u) << 1) | ((__popc(j & mask[2]) & 1u) << 2));
}

// One gate applied to this thread's tuple, reduced product (KS < 4 mixing bits: 2^(4-KS) independent blocks).  The 16
// elements are gathered first (sixteen independent shared-memory loads in flight), then the diagonal factors: ops that
// avoid the targets give one scalar per tuple; class-E ops that only see select bits give one factor per BLOCK, folded
// into that scalar (2^(4-KS) products instead of 16); class-E ops on mixing bits are applied element by element.
template <int S, int THREADS, int KS>
__device__ __forceinline__ void tile_step_blocks(const TileParams& p, double2* __restrict__ tile, const double2* __restrict__ lut,
                                                 const uint32_t* __restrict__ ctx_pb, const uint8_t* __restrict__ ctx_sel,
                                                 const uint32_t (&selh)[kTileMaxSteps][kTileMaxOps], const double2 (&stile)[kTileMaxSteps], uint64_t tbase)
{
     constexpr int DS = 1 << KS;
     constexpr int NB = 16 / DS;
     const TileStepDesc& d = p.step[S];
     const int tid = threadIdx.x;
     const uint32_t pb = ctx_pb[S * THREADS + tid];
     const uint8_t* my_sel = ctx_sel + d.sel_off * THREADS + tid;  // op j: my_sel[j * THREADS]
     double2 in[16];
#pragma unroll
     for (int c = 0; c < 16; ++c) in[c] = tile[pb ^ d.ploff[c]];
     if (d.n_ops) {
          const int n_s = d.n_ops - d.n_e;
          const int n_es = d.n_e - d.n_em;  // class-E ops that only see select bits
          double2 sc = stile[S];  // product of the ops that only see slots outside the tile
          for (int j = d.n_t; j < n_s; ++j) sc = cmul(sc, lut[d.lut_off[j] + (selh[S][j] | my_sel[j * THREADS])]);
          double2 f[NB];
#pragma unroll
          for (int blk = 0; blk < NB; ++blk) f[blk] = sc;
          for (int j = n_s; j < n_s + n_es; ++j) {
               const uint32_t sel0 = d.lut_off[j] + (selh[S][j] | my_sel[j * THREADS]);
#pragma unroll
               for (int blk = 0; blk < NB; ++blk) f[blk] = cmul(f[blk], lut[sel0 + d.esel[j][blk * DS]]);
          }
          for (int j = n_s + n_es; j < d.n_ops; ++j) {
               const uint32_t sel0 = d.lut_off[j] + (selh[S][j] | my_sel[j * THREADS]);
#pragma unroll
               for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], lut[sel0 + d.esel[j][c]]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], f[c >> KS]);
     }
     // The select bits of the gate pick the 2^KS x 2^KS block that acts on a tuple block: those inside the tile are part
     // of the tuple (block blk of the tuple <-> block blk & in_mask of the matrix), those outside it are the same for the
     // whole tile (tbase) — so only MIXING bits have to be tile bits, and a run of gates needs a much smaller tile.
     uint32_t osel = 0;
#pragma unroll
     for (int i = 0; i < 3; ++i)
          if (i < d.n_out) osel |= static_cast<uint32_t>((tbase >> d.out_slot[i]) & 1ull) << i;
     const uint32_t obase = osel << d.n_in;
#pragma unroll
     for (int blk = 0; blk < NB; ++blk) {
          const double2* __restrict__ mb = p.m[S] + ((blk & d.in_mask) | obase) * (DS * 17);
#pragma unroll
          for (int r = 0; r < DS; ++r) {
               double2 acc = make_double2(0.0, 0.0);
#pragma unroll
               for (int j = 0; j < DS; ++j) cmac(acc, mb[r * 16 + j], in[blk * DS + j]);
               tile[pb ^ d.ploff[blk * DS + r]] = acc;
          }
     }
}

// Full 16 x 16 product (three-multiplication form): the whole tuple is live.
template <int S, int THREADS>
__device__ __forceinline__ void tile_step_full(const TileParams& p, double2* __restrict__ tile, const double2* __restrict__ lut,
                                               const uint32_t* __restrict__ ctx_pb, const uint8_t* __restrict__ ctx_sel,
                                               const uint32_t (&selh)[kTileMaxSteps][kTileMaxOps], const double2 (&stile)[kTileMaxSteps])
{
     const TileStepDesc& d = p.step[S];
     const int tid = threadIdx.x;
     const uint32_t pb = ctx_pb[S * THREADS + tid];
     const uint8_t* my_sel = ctx_sel + d.sel_off * THREADS + tid;
     double2 in[16];
#pragma unroll
     for (int c = 0; c < 16; ++c) in[c] = tile[pb ^ d.ploff[c]];
     if (d.n_ops) {
          const int n_s = d.n_ops - d.n_e;
          double2 sc = stile[S];
          for (int j = d.n_t; j < n_s; ++j) sc = cmul(sc, lut[d.lut_off[j] + (selh[S][j] | my_sel[j * THREADS])]);
          for (int j = n_s; j < d.n_ops; ++j) {
               const uint32_t sel0 = selh[S][j] | my_sel[j * THREADS];
#pragma unroll
               for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], lut[d.lut_off[j] + (sel0 | d.esel[j][c])]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], sc);
     }
     apply_rows_3m(in, p.m[S], p.msum[S], [&](int b, double2 v) { tile[pb ^ d.ploff[b]] = v; });
}

// Monomial matrix (one nonzero per row and column: products of X / Y / Z / phase gates, fused permutations): no products to
// sum — element c goes to row mono_row[c] times its coefficient.
template <int S, int THREADS>
__device__ __forceinline__ void tile_step_mono(const TileParams& p, double2* __restrict__ tile, const double2* __restrict__ lut,
                                               const uint32_t* __restrict__ ctx_pb, const uint8_t* __restrict__ ctx_sel,
                                               const uint32_t (&selh)[kTileMaxSteps][kTileMaxOps], const double2 (&stile)[kTileMaxSteps])
{
     const TileStepDesc& d = p.step[S];
     const int tid = threadIdx.x;
     const uint32_t pb = ctx_pb[S * THREADS + tid];
     const uint8_t* my_sel = ctx_sel + d.sel_off * THREADS + tid;
     double2 in[16];
#pragma unroll
     for (int c = 0; c < 16; ++c) in[c] = tile[pb ^ d.ploff[c]];
     if (d.n_ops) {
          const int n_s = d.n_ops - d.n_e;
          double2 sc = stile[S];
          for (int j = d.n_t; j < n_s; ++j) sc = cmul(sc, lut[d.lut_off[j] + (selh[S][j] | my_sel[j * THREADS])]);
          for (int j = n_s; j < d.n_ops; ++j) {
               const uint32_t sel0 = selh[S][j] | my_sel[j * THREADS];
#pragma unroll
               for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], lut[d.lut_off[j] + (sel0 | d.esel[j][c])]);
          }
#pragma unroll
          for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], sc);
     }
#pragma unroll
     for (int c = 0; c < 16; ++c) tile[pb ^ d.ploff[d.mono_row[c]]] = cmul(p.m[S][c], in[c]);
}

template <int S, int THREADS, bool FULL>
__device__ __forceinline__ void tile_step(const TileParams& p, double2* __restrict__ tile, const double2* __restrict__ lut,
                                          const uint32_t* __restrict__ ctx_pb, const uint8_t* __restrict__ ctx_sel,
                                          const uint32_t (&selh)[kTileMaxSteps][kTileMaxOps], const double2 (&stile)[kTileMaxSteps], uint64_t tbase)
{
     // two forms: 4 x 4 blocks (gates with at most two mixing bits; a lone mixing bit takes a select or padding bit as
     // its partner) and the full 16 x 16 product (everything else)
     if (p.step[S].ks == 0) tile_step_mono<S, THREADS>(p, tile, lut, ctx_pb, ctx_sel, selh, stile);
     else if (p.step[S].ks <= 2) tile_step_blocks<S, THREADS, 2>(p, tile, lut, ctx_pb, ctx_sel, selh, stile, tbase);
     else if constexpr (FULL) tile_step_full<S, THREADS>(p, tile, lut, ctx_pb, ctx_sel, selh, stile);
}

// FULL = false: no gate of the run is a full 16 x 16 product (the code of that form is left out).
// NBUF = 2: the tile of the NEXT iteration is copied into a second buffer (cp.async, its own commit group) while the gates
// run on the current one, so the HBM latency and most of the transfer hide behind the arithmetic; with one buffer the CTAs
// of an SM fall into lock-step (all loading, then all computing) and the pass costs memory time PLUS compute time.
// Shared memory (dynamic): tiles | table pool (lut_pad entries) | tuple positions [gate][thread] | selector bytes [op][thread].
template <int T, bool FULL, int NBUF>
__global__ void __launch_bounds__(1 << (T - 4), T >= 12 ? 2 : (NBUF == 2 ? 3 : 4)) tile_program_kernel(const __grid_constant__ TileParams p)
{
     constexpr int THREADS = 1 << (T - 4);
     static_assert(THREADS >= kTileMaxSteps * kTileMaxOps, "one thread per (gate, op) computes the per-tile selectors");
     extern __shared__ double2 dyn_smem[];
     double2* lut = dyn_smem + NBUF * (1 << T);                                // table pool
     uint32_t* ctx_pb = reinterpret_cast<uint32_t*>(lut + p.lut_pad);          // [gate][thread]
     uint8_t* ctx_sel = reinterpret_cast<uint8_t*>(ctx_pb + p.n_steps * THREADS);  // [op of the run][thread]
     __shared__ uint32_t selh[kTileMaxSteps][kTileMaxOps];
     __shared__ double2 stile[kTileMaxSteps];
     const int tid = threadIdx.x;
     for (int i = tid; i < p.n_lut; i += THREADS) lut[i] = p.lut[i];
     // this thread's slice of the tile in the load / store phases: positions tid + i * THREADS
     const uint32_t pt = tile_phys(tid, p.swz_mask);
     uint64_t goff_t = 0;
#pragma unroll
     for (int b = 0; b < T - 4; ++b)
          if ((tid >> b) & 1) goff_t |= 1ull << p.tslot[b];
     // this thread's tuple of every gate: the same tile positions in every tile
     for (int s = 0; s < p.n_steps; ++s) {
          const TileStepDesc& d = p.step[s];
          uint32_t base = tid;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
               const uint32_t low = base & ((1u << d.tpos[i]) - 1u);
               base = ((base >> d.tpos[i]) << (d.tpos[i] + 1)) | low;
          }
          ctx_pb[s * THREADS + tid] = tile_phys(base, p.swz_mask);
          for (int j = 0; j < d.n_ops; ++j) {
               uint32_t sel = 0;
#pragma unroll
               for (int l = 0; l < 5; ++l)
                    if (d.lpos[j][l] != 0xFF) sel |= ((base >> d.lpos[j][l]) & 1u) << l;
               ctx_sel[(d.sel_off + j) * THREADS + tid] = static_cast<uint8_t>(sel);
          }
     }
     auto tile_base = [&](uint64_t t) { return insert_zero_bits(t, p.outer) << p.lo; };
     auto issue_load = [&](uint64_t t, double2* buf) {
          const double2* g = p.psi + tile_base(t) + goff_t;
#pragma unroll
          for (int i = 0; i < 16; ++i) cp_async16(&buf[pt ^ p.pi[i]], g + p.ioff[i]);
     };
     uint64_t t = blockIdx.x;
     int cur = 0;
     if (NBUF == 2 && t < p.n_tiles) {
          issue_load(t, dyn_smem);
          cp_async_commit();
     }
     __syncthreads();  // table pool and per-thread context are in place
     for (; t < p.n_tiles; t += gridDim.x) {
          double2* tile = dyn_smem + cur * (1 << T);
          const uint64_t tbase = tile_base(t);
          if (NBUF == 2) {
               // the other buffer was stored (and its reads fenced by the barrier that ended the previous iteration)
               if (t + gridDim.x < p.n_tiles) issue_load(t + gridDim.x, dyn_smem + (cur ^ 1) * (1 << T));
               cp_async_commit();
          }
          else {
               issue_load(t, tile);
          }
          if (tid < kTileMaxSteps * kTileMaxOps) {
               const int s = tid / kTileMaxOps, j = tid % kTileMaxOps;
               if (s < p.n_steps && j < p.step[s].n_ops) {
                    uint32_t sel = 0;
#pragma unroll
                    for (int l = 0; l < 5; ++l) sel |= static_cast<uint32_t>((tbase >> p.step[s].outer[j][l]) & 1ull) << l;
                    selh[s][j] = sel;
               }
          }
          if (tid >= THREADS - kTileMaxSteps) {
               // one thread per gate: the factors of the ops that only see slots outside the tile, once per tile
               const int s = tid - (THREADS - kTileMaxSteps);
               if (s < p.n_steps) {
                    double2 f = make_double2(1.0, 0.0);
                    for (int j = 0; j < p.step[s].n_t; ++j) {
                         uint32_t sel = 0;
#pragma unroll
                         for (int l = 0; l < 5; ++l) sel |= static_cast<uint32_t>((tbase >> p.step[s].outer[j][l]) & 1ull) << l;
                         f = cmul(f, lut[p.step[s].lut_off[j] + sel]);
                    }
                    stile[s] = f;
               }
          }
          if (NBUF == 2) cp_async_wait_group<1>();  // everything but the copies of the next tile has landed
          else cp_async_wait_all();
          __syncthreads();
          tile_step<0, THREADS, FULL>(p, tile, lut, ctx_pb, ctx_sel, selh, stile, tbase);
          __syncthreads();
          if (p.n_steps > 1) {
               tile_step<1, THREADS, FULL>(p, tile, lut, ctx_pb, ctx_sel, selh, stile, tbase);
               __syncthreads();
          }
          if (p.n_steps > 2) {
               tile_step<2, THREADS, FULL>(p, tile, lut, ctx_pb, ctx_sel, selh, stile, tbase);
               __syncthreads();
          }
          if (p.n_steps > 3) {
               tile_step<3, THREADS, FULL>(p, tile, lut, ctx_pb, ctx_sel, selh, stile, tbase);
               __syncthreads();
          }
          double2* g = p.psi + tbase + goff_t;
#pragma unroll
          for (int i = 0; i < 16; ++i) g[p.ioff[i]] = tile[pt ^ p.pi[i]];
          __syncthreads();  // this buffer is refilled next (one buffer: by the next tile; two: by the tile after the next)
          if (NBUF == 2) cur ^= 1;
     }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

struct PlannedStep {
     int ks = 0;           // 2 = block form (4 x 4 blocks), 4 = full product
     int n_in = 0;         // select bits inside the tile
     int n_out = 0;        // select bits outside the tile
     int lp[4];            // tile positions of the thread's tuple bits: mixing bits, inside select bits, padding
     int out_slot[4];      // slab slots of the outside select bits
     int k = 0;
     int order[5];         // matrix bits of the gate in the planned order: mixing, inside selects, outside selects
     double m[2 * 256];    // the gate's matrix in that bit order, leading dimension 16
};

struct TilePlan {
     int T = 0, lo = 0, n_hi = 0;
     int hi[16];
     std::vector<PlannedStep> steps;
     uint32_t swz_mask[3] = {0, 0, 0};
};

int local_pos(const TilePlan& pl, int slot)
{
     if (slot < pl.lo) return slot;
     for (int i = 0; i < pl.n_hi; ++i)
          if (pl.hi[i] == slot) return pl.lo + i;
     return -1;
}

// fewest contiguous low slots a tile may have: 4 = 256-byte runs in HBM (HIQ_TILE_MIN_LO=3 allows 128-byte runs, which lets
// an 8-slot run take the 2^11 tile with four CTAs per SM instead of the 2^12 one with two)
int tile_min_lo()
{
     static const int v = [] {
          const char* e = std::getenv("HIQ_TILE_MIN_LO");
          const int x = e ? std::atoi(e) : 4;
          return x < 3 ? 3 : (x > 6 ? 6 : x);
     }();
     return v;
}

bool is_monomial(int k, const double* m)
{
     const int d = 1 << k;
     for (int b = 0; b < d; ++b) {
          int in_row = 0, in_col = 0;
          for (int c = 0; c < d; ++c) {
               in_row += (m[2 * (b * d + c)] != 0.0 || m[2 * (b * d + c) + 1] != 0.0) ? 1 : 0;
               in_col += (m[2 * (c * d + b)] != 0.0 || m[2 * (c * d + b) + 1] != 0.0) ? 1 : 0;
          }
          if (in_row != 1 || in_col != 1) return false;
     }
     return true;
}

// smallest supported tile that holds at least tile_min_lo() low slots and every MIXING bit of the run: a select bit (an
// index bit of the matrix that no nonzero entry mixes) only chooses which block acts on the mixing bits, so it may stay
// outside the tile — there it is the same for the whole tile
bool choose_tile(int L, int n_steps, const hiqk_tile_step* steps, TilePlan& pl, std::string& why)
{
     uint64_t u = 0;
     pl.steps.assign(n_steps, PlannedStep());
     for (int s = 0; s < n_steps; ++s) {
          const hiqk_tile_step& st = steps[s];
          if (st.k < 1 || st.k > 4 || !st.matrix) {
               why = "gates of a tile program have 1..4 targets";
               return false;
          }
          uint64_t seen = 0;
          for (int l = 0; l < st.k; ++l) {
               if (st.slots[l] < 0 || st.slots[l] >= L || ((seen >> st.slots[l]) & 1)) {
                    why = "target slots must be distinct and < L";
                    return false;
               }
               seen |= 1ull << st.slots[l];
          }
          PlannedStep& ps = pl.steps[s];
          ps.k = st.k;
          const int raw = hiqk_dense_block_shape(st.k, st.matrix, ps.order);  // mixing bits first in ps.order
          // block form when at most two bits mix (the first two bits of the order span the blocks: a lone mixing bit is
          // joined by a select bit, or by a padding bit of the tile when the gate has one target); full product otherwise
          ps.ks = raw <= 2 ? 2 : 4;
          if (raw > 2 && is_monomial(st.k, st.matrix)) ps.ks = 0;  // monomial form: as wide as the full one, no arithmetic to speak of
          const int n_block = ps.ks == 2 ? std::min(2, st.k) : st.k;
          for (int i = 0; i < n_block; ++i) u |= 1ull << st.slots[ps.order[i]];
     }
     static const int min_t = [] {
          const char* e = std::getenv("HIQ_TILE_MIN_T");  // measurements: force the 2^12 tile
          return e ? std::atoi(e) : 11;
     }();
     for (int T = std::max(11, std::min(12, min_t)); T <= 12; ++T) {
          if (T > L) break;
          int lo = T;
          auto n_hi = [&](int l) { return __builtin_popcountll(u >> l); };
          while (lo > 0 && lo + n_hi(lo) > T) --lo;
          if (lo + n_hi(lo) != T || lo < tile_min_lo()) continue;
          pl.T = T;
          pl.lo = lo;
          pl.n_hi = 0;
          for (int s = lo; s < L; ++s)
               if ((u >> s) & 1) pl.hi[pl.n_hi++] = s;
          return true;
     }
     why = "the mixing bits of the run do not fit a 2^12 tile with 4 low slots";
     return false;
}

void plan_steps(int n_steps, const hiqk_tile_step* steps, TilePlan& pl)
{
     for (int s = 0; s < n_steps; ++s) {
          const hiqk_tile_step& st = steps[s];
          PlannedStep& ps = pl.steps[s];
          const int n_block = ps.ks == 2 ? std::min(2, st.k) : st.k;  // gate bits that span a block
          // gate bits in the planned order: block bits, select bits inside the tile (low slots, block bits of other gates
          // of the run), select bits outside it
          int order[5], n = 0;
          for (int i = 0; i < n_block; ++i) order[n++] = ps.order[i];
          ps.n_in = ps.n_out = 0;
          for (int i = n_block; i < st.k; ++i)
               if (local_pos(pl, st.slots[ps.order[i]]) >= 0) {
                    order[n++] = ps.order[i];
                    ++ps.n_in;
               }
          for (int i = n_block; i < st.k; ++i)
               if (local_pos(pl, st.slots[ps.order[i]]) < 0) {
                    ps.out_slot[ps.n_out++] = st.slots[ps.order[i]];
                    order[n++] = ps.order[i];
               }
          for (int i = 0; i < st.k; ++i) ps.order[i] = order[i];
          // the thread's tuple = four tile bits: block bits (padded to the block width with tile bits the gate does not
          // touch: identity factors), inside select bits, padding.  Padding prefers tile positions >= 3, so that the
          // gathers keep the three low address bits for the lanes of a quarter warp.
          uint32_t used = 0;
          for (int i = 0; i < st.k; ++i) {
               const int lp = local_pos(pl, st.slots[i]);
               if (lp >= 0) used |= 1u << lp;
          }
          auto pad = [&] {
               for (int pos = 3; pos < pl.T; ++pos)
                    if (!((used >> pos) & 1)) {
                         used |= 1u << pos;
                         return pos;
                    }
               for (int pos = 0; pos < 3; ++pos)
                    if (!((used >> pos) & 1)) {
                         used |= 1u << pos;
                         return pos;
                    }
               return -1;
          };
          const int block_width = ps.ks == 2 ? 2 : 4;
          int nt = 0;
          for (int i = 0; i < n_block; ++i) ps.lp[nt++] = local_pos(pl, st.slots[order[i]]);
          const int n_block_pad = block_width - n_block;  // identity factors inside the block
          for (int i = 0; i < n_block_pad; ++i) ps.lp[nt++] = pad();
          for (int i = 0; i < ps.n_in; ++i) ps.lp[nt++] = local_pos(pl, st.slots[order[n_block + i]]);
          while (nt < 4) ps.lp[nt++] = pad();
          // the matrix over [block bits, block padding, inside selects, outside selects], leading dimension 16:
          // entry (b, c) = M(gate bits of b, gate bits of c) when the padding bits of b and c agree, else 0
          const int d0 = 1 << st.k;
          const int kk = st.k + n_block_pad;  // <= 4
          std::memset(ps.m, 0, sizeof(ps.m));
          auto gate_index = [&](int x) {  // planned index -> (index into the caller's matrix, padding bits)
               int o = 0;
               for (int i = 0; i < n_block; ++i)
                    if ((x >> i) & 1) o |= 1 << order[i];
               for (int i = n_block; i < st.k; ++i)
                    if ((x >> (i + n_block_pad)) & 1) o |= 1 << order[i];
               return o;
          };
          const int pad_mask = ((1 << n_block_pad) - 1) << n_block;
          for (int b = 0; b < (1 << kk); ++b)
               for (int c = 0; c < (1 << kk); ++c) {
                    if ((b & pad_mask) != (c & pad_mask)) continue;
                    const int gb = gate_index(b), gc = gate_index(c);
                    ps.m[2 * (b * 16 + c)] = st.matrix[2 * (gb * d0 + gc)];
                    ps.m[2 * (b * 16 + c) + 1] = st.matrix[2 * (gb * d0 + gc) + 1];
               }
          if (ps.ks != 2 && kk < 4) {
               // full form of a 3-target gate: the fourth tuple bit is padding — replicate the 8 x 8 matrix on its two values
               for (int b = 0; b < 8; ++b)
                    for (int c = 0; c < 8; ++c) {
                         ps.m[2 * ((b + 8) * 16 + c + 8)] = ps.m[2 * (b * 16 + c)];
                         ps.m[2 * ((b + 8) * 16 + c + 8) + 1] = ps.m[2 * (b * 16 + c) + 1];
                    }
          }
     }
}

// XOR swizzle: physical low bits = position low bits ^ (sum of col[b] over the set bits b >= 3).  The 8 lanes of a quarter
// warp differ in the three lowest tile positions that are NOT targets of the gate; their 16-byte bank groups are distinct
// iff the three columns (unit vectors for positions < 3) are linearly independent.  Small backtracking search over the
// columns of the positions that matter; falls back to the assignment with the fewest conflicting gates.
void solve_swizzle(TilePlan& pl)
{
     const int T = pl.T;
     std::vector<std::array<int, 3>> need;
     for (const PlannedStep& ps: pl.steps) {
          uint32_t tm = 0;
          for (int i = 0; i < 4; ++i) tm |= 1u << ps.lp[i];
          std::array<int, 3> f{};
          int n = 0;
          for (int pos = 0; pos < T && n < 3; ++pos)
               if (!((tm >> pos) & 1)) f[n++] = pos;
          need.push_back(f);
     }
     std::vector<int> col(T, 0);
     for (int b = 0; b < 3; ++b) col[b] = 1 << b;
     std::vector<int> vars;
     for (auto& f: need)
          for (int pos: f)
               if (pos >= 3 && std::find(vars.begin(), vars.end(), pos) == vars.end()) vars.push_back(pos);
     std::sort(vars.begin(), vars.end());
     auto independent = [&](const std::array<int, 3>& f) {
          const int a = col[f[0]], b = col[f[1]], c = col[f[2]];
          return a && b && c && a != b && a != c && b != c && (a ^ b) != c;
     };
     auto bad = [&] {
          int n = 0;
          for (auto& f: need) n += independent(f) ? 0 : 1;
          return n;
     };
     int best_bad = 1 << 30;
     std::vector<int> best = col;
     std::function<bool(size_t)> rec = [&](size_t i) {
          if (i == vars.size()) {
               const int n = bad();
               if (n < best_bad) {
                    best_bad = n;
                    best = col;
               }
               return n == 0;
          }
          for (int v = 1; v < 8; ++v) {
               col[vars[i]] = v;
               if (rec(i + 1)) return true;
          }
          return false;
     };
     if (vars.size() <= 6) rec(0);
     else {
          // many positions matter (wide runs): greedy, gate by gate
          for (int pos: vars) {
               int best_v = 1, best_n = 1 << 30;
               for (int v = 1; v < 8; ++v) {
                    col[pos] = v;
                    const int n = bad();
                    if (n < best_n) {
                         best_n = n;
                         best_v = v;
                    }
               }
               col[pos] = best_v;
          }
          best = col;
     }
     for (int r = 0; r < 3; ++r) {
          pl.swz_mask[r] = 0;
          for (int b = 3; b < T; ++b)
               if ((best[b] >> r) & 1) pl.swz_mask[r] |= 1u << b;
     }
}

uint32_t host_phys(uint32_t j, const uint32_t (&mask)[3])
{
     uint32_t x = 0;
     for (int r = 0; r < 3; ++r) x |= (static_cast<uint32_t>(__builtin_popcount(j & mask[r])) & 1u) << r;
     return j ^ x;
}

int fill_params(TileParams& p, double2* lut_host, void* slab, int L, int n_steps, const hiqk_tile_step* steps, const TilePlan& pl,
                std::string& why)
{
     std::memset(&p, 0, sizeof(p));
     const int T = pl.T;
     p.psi = static_cast<double2*>(slab);
     p.n_tiles = 1ull << (L - T);
     p.n_steps = n_steps;
     p.lo = pl.lo;
     p.outer.n = pl.n_hi;
     for (int i = 0; i < pl.n_hi; ++i) p.outer.pos[i] = static_cast<uint8_t>(pl.hi[i] - pl.lo);
     for (int r = 0; r < 3; ++r) p.swz_mask[r] = pl.swz_mask[r];
     for (int b = 0; b < T; ++b) p.tslot[b] = static_cast<uint8_t>(b < pl.lo ? b : pl.hi[b - pl.lo]);
     for (int i = 0; i < 16; ++i) {
          uint64_t o = 0;
          for (int b = 0; b < 4; ++b)
               if ((i >> b) & 1) o |= 1ull << p.tslot[T - 4 + b];
          p.ioff[i] = o;
          p.pi[i] = static_cast<uint16_t>(host_phys(static_cast<uint32_t>(i) << (T - 4), pl.swz_mask));
     }
     int lut_used = 0, ops_total = 0;
     for (int s = 0; s < n_steps; ++s) {
          const PlannedStep& ps = pl.steps[s];
          TileStepDesc& d = p.step[s];
          d.ks = ps.ks;
          d.n_in = ps.n_in;
          d.n_out = ps.n_out;
          d.in_mask = (1 << ps.n_in) - 1;
          for (int i = 0; i < 4; ++i) d.out_slot[i] = static_cast<uint8_t>(i < ps.n_out ? ps.out_slot[i] : 63);
          int sorted[4] = {ps.lp[0], ps.lp[1], ps.lp[2], ps.lp[3]};
          std::sort(sorted, sorted + 4);
          uint32_t tm = 0;
          for (int i = 0; i < 4; ++i) {
               d.tpos[i] = static_cast<uint8_t>(sorted[i]);
               tm |= 1u << sorted[i];
          }
          for (int c = 0; c < 16; ++c) {
               uint32_t o = 0;
               for (int i = 0; i < 4; ++i)
                    if ((c >> i) & 1) o |= 1u << ps.lp[i];
               d.ploff[c] = static_cast<uint16_t>(host_phys(o, pl.swz_mask));
          }
          std::memcpy(p.m[s], ps.m, sizeof(double) * 2 * 256);
          for (int i = 0; i < 256; ++i) p.msum[s][i] = ps.m[2 * i] + ps.m[2 * i + 1];
          if (ps.ks == 0) {
               for (int c = 0; c < 16; ++c)
                    for (int b = 0; b < 16; ++b)
                         if (ps.m[2 * (b * 16 + c)] != 0.0 || ps.m[2 * (b * 16 + c) + 1] != 0.0) {
                              d.mono_row[c] = static_cast<uint8_t>(b);
                              p.m[s][c] = make_double2(ps.m[2 * (b * 16 + c)], ps.m[2 * (b * 16 + c) + 1]);
                         }
          }
          // diagonal ops: the ones that avoid the gate's (padded) targets first, class E last
          const hiqk_tile_step& st = steps[s];
          if (st.n_pre < 0 || st.n_pre > kTileMaxOps || (st.n_pre && !st.pre)) {
               why = "a gate of a tile program carries at most " + std::to_string(kTileMaxOps) + " diagonal ops";
               return HIQ_ERR_ARG;
          }
          std::vector<int> order;
          std::vector<bool> is_e(st.n_pre, false), is_em(st.n_pre, false);
          uint32_t mix_mask = 0;  // tile positions of the block bits (tuple bits 0 .. ks-1; every tuple bit in the monomial form)
          for (int i = 0; i < (ps.ks == 0 ? 4 : ps.ks) && i < 4; ++i) mix_mask |= 1u << ps.lp[i];
          for (int j = 0; j < st.n_pre; ++j) {
               const hiqk_diag_op& o = st.pre[j];
               if (o.k < 0 || o.k > kMaxTargets) {
                    why = "diagonal op with k outside 0..5";
                    return HIQ_ERR_ARG;
               }
               uint64_t seen = 0;
               for (int l = 0; l < o.k; ++l) {
                    if (o.slots[l] < 0 || o.slots[l] >= L || ((seen >> o.slots[l]) & 1)) {
                         why = "diagonal op slots must be distinct and < L";
                         return HIQ_ERR_ARG;
                    }
                    seen |= 1ull << o.slots[l];
                    const int lp = local_pos(pl, o.slots[l]);
                    if (lp >= 0 && ((tm >> lp) & 1)) is_e[j] = true;
                    if (lp >= 0 && ((mix_mask >> lp) & 1)) is_em[j] = true;
               }
          }
          std::vector<bool> is_t(st.n_pre, true);  // every slot outside the tile
          for (int j = 0; j < st.n_pre; ++j)
               for (int l = 0; l < st.pre[j].k; ++l)
                    if (local_pos(pl, st.pre[j].slots[l]) >= 0) is_t[j] = false;
          for (int j = 0; j < st.n_pre; ++j)
               if (is_t[j]) order.push_back(j);
          for (int j = 0; j < st.n_pre; ++j)
               if (!is_e[j] && !is_t[j]) order.push_back(j);
          for (int j = 0; j < st.n_pre; ++j)
               if (is_e[j] && !is_em[j]) order.push_back(j);
          for (int j = 0; j < st.n_pre; ++j)
               if (is_em[j]) order.push_back(j);
          d.n_em = static_cast<int>(std::count(is_em.begin(), is_em.end(), true));
          d.n_ops = st.n_pre;
          d.n_e = static_cast<int>(std::count(is_e.begin(), is_e.end(), true));
          d.n_t = static_cast<int>(std::count(is_t.begin(), is_t.end(), true));
          d.sel_off = ops_total;
          ops_total += st.n_pre;
          for (int jj = 0; jj < st.n_pre; ++jj) {
               const hiqk_diag_op& o = st.pre[order[jj]];
               if (lut_used + (1 << o.k) > kTileLutEntries) {
                    why = "the diagonal tables of the run exceed the table pool";
                    return HIQ_ERR_ARG;
               }
               d.lut_off[jj] = static_cast<uint16_t>(lut_used);
               std::memcpy(lut_host + lut_used, o.lut, sizeof(double2) << o.k);
               lut_used += 1 << o.k;
               for (int l = 0; l < 5; ++l) {
                    d.lpos[jj][l] = 0xFF;
                    d.outer[jj][l] = 63;
               }
               for (int l = 0; l < o.k; ++l) {
                    const int lp = local_pos(pl, o.slots[l]);
                    if (lp < 0) d.outer[jj][l] = static_cast<uint8_t>(o.slots[l]);
                    else if (!((tm >> lp) & 1)) d.lpos[jj][l] = static_cast<uint8_t>(lp);
                    else {
                         for (int c = 0; c < 16; ++c)
                              for (int i = 0; i < 4; ++i)
                                   if (ps.lp[i] == lp && ((c >> i) & 1)) d.esel[jj][c] |= static_cast<uint8_t>(1u << l);
                    }
               }
          }
     }
     p.n_lut = lut_used;
     p.lut_pad = (lut_used + 15) & ~15;
     return HIQ_OK;
}

// The table pool travels through a small ring of pinned host / device buffer pairs (it does not fit the 32 KB of kernel
// parameters next to the matrices): slot i is reused only after the launch that read it has completed (an event per slot),
// whatever stream that launch went to.
struct LutRing {
     static constexpr int kDepth = 8;
     double2* host[kDepth] = {};
     double2* dev[kDepth] = {};
     cudaEvent_t ev[kDepth] = {};
     bool used[kDepth] = {};
     int next = 0;
     bool ready = false;
};
LutRing g_ring[16];

int ring_acquire(LutRing*& ring, int& slot)
{
     int device = 0;
     HIQ_CUDA(cudaGetDevice(&device));
     if (device < 0 || device >= 16) return set_error(HIQ_ERR_ARG, "hiqk_apply_tile_program: device index out of range");
     LutRing& r = g_ring[device];
     if (!r.ready) {
          for (int i = 0; i < LutRing::kDepth; ++i) {
               HIQ_CUDA(cudaMallocHost(&r.host[i], sizeof(double2) * kTileLutEntries));
               HIQ_CUDA(cudaMalloc(&r.dev[i], sizeof(double2) * kTileLutEntries));
               HIQ_CUDA(cudaEventCreateWithFlags(&r.ev[i], cudaEventDisableTiming));
          }
          r.ready = true;
     }
     slot = r.next;
     r.next = (r.next + 1) % LutRing::kDepth;
     if (r.used[slot]) HIQ_CUDA(cudaEventSynchronize(r.ev[slot]));
     ring = &r;
     return HIQ_OK;
}

size_t tile_smem_bytes(int T, int nbuf, const TileParams& p)
{
     int ops_total = 0;
     for (int s = 0; s < p.n_steps; ++s) ops_total += p.step[s].n_ops;
     const size_t threads = 1u << (T - 4);
     return sizeof(double2) * (static_cast<size_t>(nbuf) * (1u << T) + p.lut_pad) + threads * (4 * p.n_steps + ops_total) + 16;
}

}  // namespace
}  // namespace hiq

using namespace hiq;

extern "C" int hiqk_dense_is_monomial(int k, const double* matrix)
{
     if (!matrix || k < 1 || k > kMaxTargets) return 0;
     return is_monomial(k, matrix) ? 1 : 0;
}

extern "C" int hiqk_tile_program_fits(int L, int n_steps, const hiqk_tile_step* steps)
{
     if (!steps || n_steps < 1 || n_steps > kTileMaxSteps || L > 40) return 0;
     TilePlan pl;
     std::string why;
     if (!choose_tile(L, n_steps, steps, pl, why)) return 0;
     int lut = 0;
     for (int s = 0; s < n_steps; ++s) {
          if (steps[s].n_pre < 0 || steps[s].n_pre > kTileMaxOps || (steps[s].n_pre && !steps[s].pre)) return 0;
          for (int j = 0; j < steps[s].n_pre; ++j) {
               if (steps[s].pre[j].k < 0 || steps[s].pre[j].k > kMaxTargets) return 0;
               lut += 1 << steps[s].pre[j].k;
          }
     }
     if (lut > kTileLutEntries) return 0;
     if (const char* e = std::getenv("HIQ_TILE_DEBUG"); e && e[0] == '2') {
          // where the diagonal ops of the run would sit (host-only analysis)
          plan_steps(n_steps, steps, pl);
          std::fprintf(stderr, "  plan T=%d lo=%d high slots:", pl.T, pl.lo);
          for (int i = 0; i < pl.n_hi; ++i) std::fprintf(stderr, " %d", pl.hi[i]);
          std::fprintf(stderr, "\n");
          for (int s = 0; s < n_steps; ++s) {
               uint32_t tm = 0;
               for (int i = 0; i < 4; ++i) tm |= 1u << pl.steps[s].lp[i];
               int outer_only = 0, on_tuple = 0;
               for (int j = 0; j < steps[s].n_pre; ++j) {
                    bool any_in = false, tup = false;
                    for (int l = 0; l < steps[s].pre[j].k; ++l) {
                         const int lp = local_pos(pl, steps[s].pre[j].slots[l]);
                         any_in = any_in || lp >= 0;
                         tup = tup || (lp >= 0 && ((tm >> lp) & 1));
                    }
                    outer_only += any_in ? 0 : 1;
                    on_tuple += tup ? 1 : 0;
               }
               std::fprintf(stderr, "    gate %d (form %d, selects in/out %d/%d): %d ops, %d with every slot outside the tile, %d on tuple bits\n", s,
                            pl.steps[s].ks, pl.steps[s].n_in, pl.steps[s].n_out, steps[s].n_pre, outer_only, on_tuple);
          }
     }
     return pl.T;
}

// Host-only: the parameter image of a launch (see include/hiq_b200.h).  Header words, in order: magic, tile bits,
// sizeof(TileParams), sizeof(TileStepDesc), max steps, max ops, table-pool entries, sizeof(InsertBits), then the offsets
// of the TileParams fields {n_tiles, n_steps, lo, outer, swz_mask, ioff, pi, tslot, step, n_lut, lut_pad, m, msum} and of
// the TileStepDesc fields {ks, n_ops, n_e, sel_off, n_t, n_em, in_mask, n_in, n_out, out_slot, mono_row, tpos, ploff,
// lut_off, lpos, outer, esel}.
namespace {
constexpr int kImageHeaderWords = 64;
}

extern "C" size_t hiqk_tile_program_image_bytes(void)
{
     return kImageHeaderWords * sizeof(uint32_t) + sizeof(TileParams) + sizeof(double2) * kTileLutEntries;
}

extern "C" int hiqk_tile_program_image(int L, int n_steps, const hiqk_tile_step* steps, void* image, size_t image_bytes)
{
     if (!steps || !image) return set_error(HIQ_ERR_ARG, "hiqk_tile_program_image: null argument");
     if (image_bytes < hiqk_tile_program_image_bytes()) return set_error(HIQ_ERR_ARG, "hiqk_tile_program_image: buffer too small");
     if (n_steps < 1 || n_steps > kTileMaxSteps || L > 40)
          return set_error(HIQ_ERR_ARG, "hiqk_tile_program_image: a tile program holds 1.." + std::to_string(kTileMaxSteps) + " gates");
     TilePlan pl;
     std::string why;
     if (!choose_tile(L, n_steps, steps, pl, why)) return set_error(HIQ_ERR_ARG, "hiqk_tile_program_image: " + why);
     plan_steps(n_steps, steps, pl);
     solve_swizzle(pl);
     uint32_t* head = static_cast<uint32_t*>(image);
     TileParams* p = reinterpret_cast<TileParams*>(head + kImageHeaderWords);
     double2* lut = reinterpret_cast<double2*>(reinterpret_cast<char*>(p) + sizeof(TileParams));
     std::memset(image, 0, hiqk_tile_program_image_bytes());
     const int rc = fill_params(*p, lut, nullptr, L, n_steps, steps, pl, why);
     if (rc != HIQ_OK) return set_error(rc, "hiqk_tile_program_image: " + why);
     int w = 0;
     head[w++] = 0x50545148u;  // 'HQTP'
     head[w++] = static_cast<uint32_t>(pl.T);
     head[w++] = sizeof(TileParams);
     head[w++] = sizeof(TileStepDesc);
     head[w++] = kTileMaxSteps;
     head[w++] = kTileMaxOps;
     head[w++] = kTileLutEntries;
     head[w++] = sizeof(InsertBits);
#define HIQ_P_OFF(field) head[w++] = static_cast<uint32_t>(offsetof(TileParams, field))
     HIQ_P_OFF(n_tiles);
     HIQ_P_OFF(n_steps);
     HIQ_P_OFF(lo);
     HIQ_P_OFF(outer);
     HIQ_P_OFF(swz_mask);
     HIQ_P_OFF(ioff);
     HIQ_P_OFF(pi);
     HIQ_P_OFF(tslot);
     HIQ_P_OFF(step);
     HIQ_P_OFF(n_lut);
     HIQ_P_OFF(lut_pad);
     HIQ_P_OFF(m);
     HIQ_P_OFF(msum);
#undef HIQ_P_OFF
#define HIQ_S_OFF(field) head[w++] = static_cast<uint32_t>(offsetof(TileStepDesc, field))
     HIQ_S_OFF(ks);
     HIQ_S_OFF(n_ops);
     HIQ_S_OFF(n_e);
     HIQ_S_OFF(sel_off);
     HIQ_S_OFF(n_t);
     HIQ_S_OFF(n_em);
     HIQ_S_OFF(in_mask);
     HIQ_S_OFF(n_in);
     HIQ_S_OFF(n_out);
     HIQ_S_OFF(out_slot);
     HIQ_S_OFF(mono_row);
     HIQ_S_OFF(tpos);
     HIQ_S_OFF(ploff);
     HIQ_S_OFF(lut_off);
     HIQ_S_OFF(lpos);
     HIQ_S_OFF(outer);
     HIQ_S_OFF(esel);
#undef HIQ_S_OFF
     static_assert(8 + 13 + 17 <= kImageHeaderWords, "header words");
     return HIQ_OK;
}

extern "C" int hiqk_apply_tile_program(void* slab, int L, int n_steps, const hiqk_tile_step* steps, void* stream)
{
     if (!slab || !steps) return set_error(HIQ_ERR_ARG, "hiqk_apply_tile_program: null argument");
     if (n_steps < 1 || n_steps > kTileMaxSteps || L > 40)
          return set_error(HIQ_ERR_ARG, "hiqk_apply_tile_program: a tile program holds 1.." + std::to_string(kTileMaxSteps) + " gates");
     TilePlan pl;
     std::string why;
     if (!choose_tile(L, n_steps, steps, pl, why)) return set_error(HIQ_ERR_ARG, "hiqk_apply_tile_program: " + why);
     plan_steps(n_steps, steps, pl);
     solve_swizzle(pl);
     static TileParams p;  // ~27 KB: off the stack; launches are issued from one host thread per engine
     static std::mutex mu;
     std::lock_guard<std::mutex> lock(mu);
     LutRing* ring = nullptr;
     int slot = 0;
     int rc = ring_acquire(ring, slot);
     if (rc != HIQ_OK) return rc;
     rc = fill_params(p, ring->host[slot], slab, L, n_steps, steps, pl, why);
     if (rc != HIQ_OK) return set_error(rc, "hiqk_apply_tile_program: " + why);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     p.lut = ring->dev[slot];
     if (p.n_lut)
          HIQ_CUDA(cudaMemcpyAsync(ring->dev[slot], ring->host[slot], sizeof(double2) * p.n_lut, cudaMemcpyHostToDevice, st));
     bool full = false;
     for (const PlannedStep& ps: pl.steps) full = full || ps.ks >= 4;
     static const bool double_buffer = [] {
          const char* e = std::getenv("HIQ_TILE_DOUBLE_BUFFER");  // A/B measurements: 1 = two buffers
          return e && e[0] == '1';
     }();
     auto go = [&](void (*kernel)(TileParams), int threads, int nbuf) {
          const size_t smem = tile_smem_bytes(pl.T, nbuf, p);
          static std::vector<void (*)(TileParams)> configured;  // under `mu`: the kernels whose shared-memory limit is raised
          if (std::find(configured.begin(), configured.end(), kernel) == configured.end()) {
               cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
               configured.push_back(kernel);
          }
          int per_sm = 0;
          if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem) != cudaSuccess || per_sm < 1) {
               cudaGetLastError();
               per_sm = 1;
          }
          const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n_tiles, grid_cap(static_cast<uint64_t>(num_sms()) * per_sm)));
          kernel<<<grid, threads, smem, st>>>(p);
     };
     if (pl.T == 11 && double_buffer) {
          if (full) go(tile_program_kernel<11, true, 2>, 1 << 7, 2);
          else go(tile_program_kernel<11, false, 2>, 1 << 7, 2);
     }
     else if (pl.T == 11) {
          if (full) go(tile_program_kernel<11, true, 1>, 1 << 7, 1);
          else go(tile_program_kernel<11, false, 1>, 1 << 7, 1);
     }
     else if (full) go(tile_program_kernel<12, true, 1>, 1 << 8, 1);
     else go(tile_program_kernel<12, false, 1>, 1 << 8, 1);
     count_launch();
     ring->used[slot] = true;
     cudaEventRecord(ring->ev[slot], st);
     return check_launch("tile_program_kernel");
}
