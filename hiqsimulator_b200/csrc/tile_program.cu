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
constexpr int kTileLutEntries = 512;            // table pool shared by all ops of a launch (8 KB of shared memory)
// per-thread context in shared memory, computed once per launch: the physical tile position of the thread's tuple of
// every gate (u32) and, per diagonal op, the selector bits its tuple base contributes (u8)
constexpr int kTileCtxBytes = kTileMaxSteps * (4 + kTileMaxOps);

struct TileStepDesc {
     int ks;                          // mixing bits, 1..4 (4 = full product, three-multiplication form)
     int n_ops, n_e;                  // diagonal ops applied to a tuple before the gate; the last n_e touch its targets
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
     double2 m[kTileMaxSteps][256];
     double msum[kTileMaxSteps][256];
};
static_assert(sizeof(TileParams) <= 32764, "kernel parameters are limited to 32764 bytes");

__device__ __forceinline__ uint32_t tile_phys(uint32_t j, const uint32_t (&mask)[3])
{
     return j ^ ((__popc(j & mask[0]) & 1u) | ((__popc(j & mask[1]) & 1u) << 1) | ((__popc(j & mask[2]) & 1u) << 2));
}

template <int S, int THREADS>
__device__ __forceinline__ void tile_step(const TileParams& p, double2* __restrict__ tile, const double2* __restrict__ lut,
                                          const uint32_t* __restrict__ ctx_pb, const uint8_t* __restrict__ ctx_sel,
                                          const uint32_t (&selh)[kTileMaxSteps][kTileMaxOps])
{
     const TileStepDesc& d = p.step[S];
     const int tid = threadIdx.x;
     const uint32_t pb = ctx_pb[S * THREADS + tid];
     const uint8_t* my_sel = ctx_sel + (S * kTileMaxOps) * THREADS + tid;  // op j: my_sel[j * THREADS]
     double2 in[16];
#pragma unroll
     for (int c = 0; c < 16; ++c) in[c] = tile[pb ^ d.ploff[c]];
     if (d.n_ops) {
          const int n_s = d.n_ops - d.n_e;
          double2 sc = make_double2(1.0, 0.0);
          for (int j = 0; j < n_s; ++j) sc = cmul(sc, lut[d.lut_off[j] + (selh[S][j] | my_sel[j * THREADS])]);
          if (d.n_e == 0) {
#pragma unroll
               for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], sc);
          }
          else {
               for (int j = n_s; j < d.n_ops; ++j) {
                    const uint32_t sel0 = selh[S][j] | my_sel[j * THREADS];
#pragma unroll
                    for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], lut[d.lut_off[j] + (sel0 | d.esel[j][c])]);
               }
               if (n_s) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) in[c] = cmul(in[c], sc);
               }
          }
     }
     auto store = [&](int b, double2 v) { tile[pb ^ d.ploff[b]] = v; };
     switch (d.ks) {
          case 1: apply_rows<4, 1>(in, p.m[S], store); break;
          case 2: apply_rows<4, 2>(in, p.m[S], store); break;
          case 3: apply_rows<4, 3>(in, p.m[S], store); break;
          default: apply_rows_3m(in, p.m[S], p.msum[S], store); break;
     }
}

template <int T>
__global__ void __launch_bounds__(1 << (T - 4), T >= 12 ? 2 : 4) tile_program_kernel(const __grid_constant__ TileParams p)
{
     constexpr int THREADS = 1 << (T - 4);
     static_assert(THREADS >= kTileMaxSteps * kTileMaxOps, "one thread per (gate, op) computes the per-tile selectors");
     extern __shared__ double2 dyn_smem[];
     double2* tile = dyn_smem;                                         // 2^T amplitudes
     double2* lut = dyn_smem + (1 << T);                               // table pool
     uint32_t* ctx_pb = reinterpret_cast<uint32_t*>(lut + kTileLutEntries);   // [gate][thread]
     uint8_t* ctx_sel = reinterpret_cast<uint8_t*>(ctx_pb + kTileMaxSteps * THREADS);  // [gate][op][thread]
     __shared__ uint32_t selh[kTileMaxSteps][kTileMaxOps];
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
               ctx_sel[(s * kTileMaxOps + j) * THREADS + tid] = static_cast<uint8_t>(sel);
          }
     }
     __syncthreads();
     for (uint64_t t = blockIdx.x; t < p.n_tiles; t += gridDim.x) {
          const uint64_t tbase = insert_zero_bits(t, p.outer) << p.lo;
          double2* g = p.psi + tbase + goff_t;
#pragma unroll
          for (int i = 0; i < 16; ++i) cp_async16(&tile[pt ^ p.pi[i]], g + p.ioff[i]);
          if (tid < kTileMaxSteps * kTileMaxOps) {
               const int s = tid / kTileMaxOps, j = tid % kTileMaxOps;
               if (s < p.n_steps && j < p.step[s].n_ops) {
                    uint32_t sel = 0;
#pragma unroll
                    for (int l = 0; l < 5; ++l) sel |= static_cast<uint32_t>((tbase >> p.step[s].outer[j][l]) & 1ull) << l;
                    selh[s][j] = sel;
               }
          }
          cp_async_wait_all();
          __syncthreads();
          tile_step<0, THREADS>(p, tile, lut, ctx_pb, ctx_sel, selh);
          __syncthreads();
          if (p.n_steps > 1) {
               tile_step<1, THREADS>(p, tile, lut, ctx_pb, ctx_sel, selh);
               __syncthreads();
          }
          if (p.n_steps > 2) {
               tile_step<2, THREADS>(p, tile, lut, ctx_pb, ctx_sel, selh);
               __syncthreads();
          }
          if (p.n_steps > 3) {
               tile_step<3, THREADS>(p, tile, lut, ctx_pb, ctx_sel, selh);
               __syncthreads();
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) g[p.ioff[i]] = tile[pt ^ p.pi[i]];
          __syncthreads();  // the next tile's copies overwrite the tile
     }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
namespace {

struct PlannedStep {
     int lp[4];            // tile-local positions of the matrix bits (mixing bits first)
     int ks;
     double m[2 * 256];    // 16 x 16 complex, index bits in the order of lp
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

// smallest supported tile that holds at least tile_min_lo() low slots and every target of the run
bool choose_tile(int L, int n_steps, const hiqk_tile_step* steps, TilePlan& pl, std::string& why)
{
     uint64_t u = 0;
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
          u |= seen;
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
     why = "the targets of the run do not fit a 2^12 tile with 4 low slots";
     return false;
}

void plan_steps(int n_steps, const hiqk_tile_step* steps, TilePlan& pl)
{
     pl.steps.resize(n_steps);
     for (int s = 0; s < n_steps; ++s) {
          const hiqk_tile_step& st = steps[s];
          int lp[4];
          uint32_t used = 0;
          for (int l = 0; l < st.k; ++l) {
               lp[l] = local_pos(pl, st.slots[l]);
               used |= 1u << lp[l];
          }
          // k < 4: add select bits the matrix does not mix — tile positions >= 3 first, so the gathers keep the three low
          // address bits for the lanes of a quarter warp
          int k = st.k;
          for (int pos = 3; pos < pl.T && k < 4; ++pos)
               if (!((used >> pos) & 1)) {
                    lp[k++] = pos;
                    used |= 1u << pos;
               }
          for (int pos = 0; pos < 3 && k < 4; ++pos)
               if (!((used >> pos) & 1)) {
                    lp[k++] = pos;
                    used |= 1u << pos;
               }
          // I (x) M on the added high index bits
          const int d0 = 1 << st.k;
          std::vector<double> full(2 * 256, 0.0);
          for (int b = 0; b < 16; ++b)
               for (int c = 0; c < 16; ++c)
                    if ((b >> st.k) == (c >> st.k)) {
                         full[2 * (b * 16 + c)] = st.matrix[2 * ((b & (d0 - 1)) * d0 + (c & (d0 - 1)))];
                         full[2 * (b * 16 + c) + 1] = st.matrix[2 * ((b & (d0 - 1)) * d0 + (c & (d0 - 1))) + 1];
                    }
          int order[5];
          PlannedStep& ps = pl.steps[s];
          ps.ks = hiqk_dense_block_shape(4, full.data(), order);
          int map[16];
          for (int x = 0; x < 16; ++x) {
               int o = 0;
               for (int i = 0; i < 4; ++i)
                    if ((x >> i) & 1) o |= 1 << order[i];
               map[x] = o;
          }
          for (int i = 0; i < 4; ++i) ps.lp[i] = lp[order[i]];
          for (int b = 0; b < 16; ++b)
               for (int c = 0; c < 16; ++c) {
                    ps.m[2 * (b * 16 + c)] = full[2 * (map[b] * 16 + map[c])];
                    ps.m[2 * (b * 16 + c) + 1] = full[2 * (map[b] * 16 + map[c]) + 1];
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
     int lut_used = 0;
     for (int s = 0; s < n_steps; ++s) {
          const PlannedStep& ps = pl.steps[s];
          TileStepDesc& d = p.step[s];
          d.ks = ps.ks;
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
          // diagonal ops: the ones that avoid the gate's (padded) targets first, class E last
          const hiqk_tile_step& st = steps[s];
          if (st.n_pre < 0 || st.n_pre > kTileMaxOps || (st.n_pre && !st.pre)) {
               why = "a gate of a tile program carries at most " + std::to_string(kTileMaxOps) + " diagonal ops";
               return HIQ_ERR_ARG;
          }
          std::vector<int> order;
          std::vector<bool> is_e(st.n_pre, false);
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
               }
          }
          for (int j = 0; j < st.n_pre; ++j)
               if (!is_e[j]) order.push_back(j);
          for (int j = 0; j < st.n_pre; ++j)
               if (is_e[j]) order.push_back(j);
          d.n_ops = st.n_pre;
          d.n_e = static_cast<int>(std::count(is_e.begin(), is_e.end(), true));
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

size_t tile_smem_bytes(int T)
{
     return sizeof(double2) * ((1u << T) + kTileLutEntries) + static_cast<size_t>(kTileCtxBytes) * (1u << (T - 4));
}

}  // namespace
}  // namespace hiq

using namespace hiq;

extern "C" int hiqk_tile_program_fits(int L, int n_steps, const hiqk_tile_step* steps)
{
     if (!steps || n_steps < 1 || n_steps > kTileMaxSteps || L > 40) return 0;
     TilePlan pl;
     std::string why;
     if (!choose_tile(L, n_steps, steps, pl, why)) return 0;
     int lut = 0;
     for (int s = 0; s < n_steps; ++s) {
          if (steps[s].n_pre < 0 || steps[s].n_pre > kTileMaxOps) return 0;
          for (int j = 0; j < steps[s].n_pre; ++j) {
               if (steps[s].pre[j].k < 0 || steps[s].pre[j].k > kMaxTargets) return 0;
               lut += 1 << steps[s].pre[j].k;
          }
     }
     if (lut > kTileLutEntries) return 0;
     return pl.T;
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
     const size_t smem = tile_smem_bytes(pl.T);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     p.lut = ring->dev[slot];
     if (p.n_lut)
          HIQ_CUDA(cudaMemcpyAsync(ring->dev[slot], ring->host[slot], sizeof(double2) * p.n_lut, cudaMemcpyHostToDevice, st));
     if (pl.T == 11) {
          static bool attr = false;
          if (!attr) {
               cudaFuncSetAttribute(tile_program_kernel<11>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
               attr = true;
          }
          const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n_tiles, grid_cap(static_cast<uint64_t>(num_sms()) * 4)));
          tile_program_kernel<11><<<grid, 1 << 7, smem, st>>>(p);
     }
     else {
          static bool attr = false;
          if (!attr) {
               cudaFuncSetAttribute(tile_program_kernel<12>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
               attr = true;
          }
          const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n_tiles, grid_cap(static_cast<uint64_t>(num_sms()) * 2)));
          tile_program_kernel<12><<<grid, 1 << 8, smem, st>>>(p);
     }
     count_launch();
     ring->used[slot] = true;
     cudaEventRecord(ring->ev[slot], st);
     return check_launch("tile_program_kernel");
}
