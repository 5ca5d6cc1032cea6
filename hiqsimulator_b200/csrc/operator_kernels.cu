// Operator-level passes over the local slab: Pauli-string sums (get_expectation_value /
// apply_qubit_operator) and register permutations (emulate_math).
//
// The reference wrapper calls these on its C++ simulator (reference:
// hiq/projectq/backends/_sim/_simulator_mpi.py:180-183, 220-223, 467) but the reference class does not
// implement them (SimulatorMPI.hpp:217-225 throws; the other two are not exported,
// _cppsim_mpi.cpp:63-82).  Semantics: ProjectQ's C++ simulator (apply_term / get_expectation_value /
// apply_qubit_operator / emulate_math in projectq/backends/_sim/_cppkernels/simulator.hpp), which
// copies the state and applies X/Y/Z gate by gate — three slab copies and one pass per Pauli factor.
// Here a Pauli string P acts as  (P psi)[i ^ x] = i^{#Y} (-1)^{popcount(i & z)} psi[i]  and all the
// terms of an operator that share the flip mask x are ONE pass: F(i) = sum_t c_t (-1)^{popcount(i & z_t)}.
//   expectation   one read-only pass (16 B/amplitude): pairs (i, i^x) are visited once
//   apply         in place when every term shares x (32 B/amplitude), else accumulated into a second buffer
//   permutation   gather through the inverse map, optionally from peer-mapped slabs (NVLink loads)
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "hiq_device.cuh"
#include "hiq_host.hpp"

namespace hiq {

constexpr int kOpThreads = 256;
constexpr int kOpMaxPartials = 4096;  // per component; the workspace holds 2 x 8192 doubles

static unsigned op_grid(uint64_t items, int per_thread)
{
     const uint64_t need = (items + static_cast<uint64_t>(kOpThreads) * per_thread - 1) / (static_cast<uint64_t>(kOpThreads) * per_thread);
     const uint64_t cap = grid_cap(static_cast<uint64_t>(num_sms()) * 8 * 2);
     return static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(need, cap)));
}

struct PauliParams {
     const double2* psi;  // local slab: the partner side psi[i ^ xmask]
     const double2* src;  // S[i] = src[i - begin]  (psi + begin for a local pass)
     double2* out;        // apply: destination (slab itself in place, or the accumulator)
     uint64_t begin, count;
     uint64_t xmask;
     int top;             // highest set bit of xmask
     int n_terms;
     int accumulate;
     uint64_t zmask[HIQK_MAX_PAULI_TERMS];
     double2 coef[HIQK_MAX_PAULI_TERMS];
};

__device__ __forceinline__ double2 pauli_factor(const PauliParams& p, uint64_t i)
{
     double2 f = make_double2(0.0, 0.0);
#pragma unroll 1
     for (int t = 0; t < p.n_terms; ++t) {
          const double s = (__popcll(i & p.zmask[t]) & 1) ? -1.0 : 1.0;
          f.x = fma(s, p.coef[t].x, f.x);
          f.y = fma(s, p.coef[t].y, f.y);
     }
     return f;
}

// conj(a) * b
__device__ __forceinline__ double2 cmul_conj(const double2 a, const double2 b)
{
     return make_double2(fma(a.x, b.x, a.y * b.y), fma(a.x, b.y, -a.y * b.x));
}

__device__ __forceinline__ uint64_t insert_zero(uint64_t f, int pos)
{
     const uint64_t low = f & ((1ull << pos) - 1ull);
     return ((f >> pos) << (pos + 1)) | low;
}

enum { PAULI_DIAG = 0, PAULI_PAIRED = 1, PAULI_RANGE = 2 };

// ----------------------------------------------------------------------------- expectation
// partials[b] / partials[grid + b] = real / imaginary part of CTA b's sum (fixed tree: deterministic)
template <int MODE>
__global__ void __launch_bounds__(kOpThreads) pauli_expect_kernel(const __grid_constant__ PauliParams p, double* partials)
{
     __shared__ double scratch[kOpThreads / 32];
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kOpThreads;
     double2 acc = make_double2(0.0, 0.0);
     uint64_t g = static_cast<uint64_t>(blockIdx.x) * kOpThreads + threadIdx.x;
     if (MODE == PAULI_DIAG) {
          // x = 0: sum F(i) |psi[i]|^2
          for (; g + stride < p.count; g += 2 * stride) {
               const double2 v0 = ldg_stream(p.psi + g), v1 = ldg_stream(p.psi + g + stride);
               const double2 f0 = pauli_factor(p, g), f1 = pauli_factor(p, g + stride);
               const double n0 = norm2(v0), n1 = norm2(v1);
               acc.x += f0.x * n0 + f1.x * n1;
               acc.y += f0.y * n0 + f1.y * n1;
          }
          for (; g < p.count; g += stride) {
               const double2 v = ldg_stream(p.psi + g);
               const double2 f = pauli_factor(p, g);
               const double n = norm2(v);
               acc.x += f.x * n;
               acc.y += f.y * n;
          }
     }
     else if (MODE == PAULI_PAIRED) {
          // every pair (i, j = i ^ x) once: conj(psi[j]) F(i) psi[i] + conj(psi[i]) F(j) psi[j]
          for (; g < p.count; g += stride) {
               const uint64_t i = insert_zero(g, p.top);
               const uint64_t j = i ^ p.xmask;
               const double2 a = ldg_stream(p.psi + i), b = ldg_stream(p.psi + j);
               const double2 t0 = cmul(pauli_factor(p, i), cmul_conj(b, a));
               const double2 t1 = cmul(pauli_factor(p, j), cmul_conj(a, b));
               acc.x += t0.x + t1.x;
               acc.y += t0.y + t1.y;
          }
     }
     else {
          // i in [begin, begin + count): conj(psi[i ^ x]) F(i) S[i]
          for (; g < p.count; g += stride) {
               const uint64_t i = p.begin + g;
               const double2 s = ldg_stream(p.src + g), b = ldg_stream(p.psi + (i ^ p.xmask));
               const double2 t = cmul(pauli_factor(p, i), cmul_conj(b, s));
               acc.x += t.x;
               acc.y += t.y;
          }
     }
     const double sx = block_sum<kOpThreads>(acc.x, scratch);
     const double sy = block_sum<kOpThreads>(acc.y, scratch);
     if (threadIdx.x == 0) {
          partials[blockIdx.x] = sx;
          partials[gridDim.x + blockIdx.x] = sy;
     }
}

__global__ void __launch_bounds__(1024) fold_pairs_kernel(const double* partials, int n, double* out)
{
     // out[o] = sum_j partials[o * n + j], o = 0, 1 — fixed order, one CTA
     __shared__ double scratch[32];
     for (int o = 0; o < 2; ++o) {
          double acc = 0.0;
          for (int j = threadIdx.x; j < n; j += 1024) acc += partials[o * n + j];
          const double s = block_sum<1024>(acc, scratch);
          if (threadIdx.x == 0) out[o] = s;
     }
}

// ----------------------------------------------------------------------------- apply
template <int MODE>
__global__ void __launch_bounds__(kOpThreads) pauli_apply_kernel(const __grid_constant__ PauliParams p)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kOpThreads;
     uint64_t g = static_cast<uint64_t>(blockIdx.x) * kOpThreads + threadIdx.x;
     if (MODE == PAULI_DIAG) {
          // in place, x = 0: psi[i] *= F(i)
          for (; g + stride < p.count; g += 2 * stride) {
               const double2 v0 = ldg_stream(p.psi + g), v1 = ldg_stream(p.psi + g + stride);
               p.out[g] = cmul(pauli_factor(p, g), v0);
               p.out[g + stride] = cmul(pauli_factor(p, g + stride), v1);
          }
          for (; g < p.count; g += stride) p.out[g] = cmul(pauli_factor(p, g), ldg_stream(p.psi + g));
     }
     else if (MODE == PAULI_PAIRED) {
          // in place: the pair (i, j = i ^ x) trades places, (P psi)[j] = F(i) psi[i], (P psi)[i] = F(j) psi[j]
          for (; g < p.count; g += stride) {
               const uint64_t i = insert_zero(g, p.top);
               const uint64_t j = i ^ p.xmask;
               const double2 a = ldg_stream(p.psi + i), b = ldg_stream(p.psi + j);
               p.out[j] = cmul(pauli_factor(p, i), a);
               p.out[i] = cmul(pauli_factor(p, j), b);
          }
     }
     else {
          // out[i ^ x] (+)= F(i) S[i], i in [begin, begin + count)
          for (; g < p.count; g += stride) {
               const uint64_t i = p.begin + g;
               double2 v = cmul(pauli_factor(p, i), ldg_stream(p.src + g));
               double2* d = p.out + (i ^ p.xmask);
               if (p.accumulate) {
                    const double2 o = *d;
                    v.x += o.x;
                    v.y += o.y;
               }
               *d = v;
          }
     }
}

// ----------------------------------------------------------------------------- register permutation
constexpr int kMaxPermSlabs = 16;
constexpr int kMaxPermBits = 40;

struct PermParams {
     double2* dst;
     const double2* slabs[kMaxPermSlabs];  // slab of every rank that can hold a source amplitude (peer-mapped)
     const uint32_t* table;                // TABLE: source register value per destination value
     uint64_t n;                           // 2^L
     uint64_t rank_bits;                   // rank << L
     uint64_t ctrl_mask, reg_mask;         // over the global amplitude index
     uint64_t a, N;                        // ADD: subtrahend, modulus 2^n_bits in N;  *_MOD: a = inverse constant, N
     int L, kind, n_bits;
     int contig;                           // >= 0: the register is index bits [contig, contig + n_bits)
     uint8_t pos[kMaxPermBits];
};

__device__ __forceinline__ uint64_t perm_source_value(const PermParams& p, uint64_t v)
{
     switch (p.kind) {
          case HIQK_PERM_TABLE: return p.table[v];
          case HIQK_PERM_ADD: return (v - p.a) & (p.N - 1ull);  // N = 2^n_bits, a already reduced
          case HIQK_PERM_ADD_MOD: return v >= p.N ? v : (v >= p.a ? v - p.a : v + p.N - p.a);
          default: return v >= p.N ? v : (p.a * v) % p.N;       // a = multiplier^-1 mod N, a * v < 2^64
     }
}

__global__ void __launch_bounds__(kOpThreads) permute_gather_kernel(const __grid_constant__ PermParams p)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kOpThreads;
     const uint64_t lmask = p.n - 1ull;
     for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * kOpThreads + threadIdx.x; j < p.n; j += stride) {
          const uint64_t G = p.rank_bits | j;
          uint64_t s = G;
          if ((G & p.ctrl_mask) == p.ctrl_mask) {
               uint64_t v = 0;
               if (p.contig >= 0) v = (G >> p.contig) & ((1ull << p.n_bits) - 1ull);
               else
                    for (int b = 0; b < p.n_bits; ++b) v |= ((G >> p.pos[b]) & 1ull) << b;
               const uint64_t u = perm_source_value(p, v);
               uint64_t dep = 0;
               if (p.contig >= 0) dep = u << p.contig;
               else
                    for (int b = 0; b < p.n_bits; ++b) dep |= ((u >> b) & 1ull) << p.pos[b];
               s = (G & ~p.reg_mask) | dep;
          }
          p.dst[j] = ldg_stream(p.slabs[s >> p.L] + (s & lmask));
     }
}

static int fill_pauli(PauliParams& p, const char* who, const void* slab, int L, uint64_t xmask, const hiqk_pauli_term* terms,
                      int n_terms)
{
     if (!slab || !terms) return set_error(HIQ_ERR_ARG, std::string(who) + ": null argument");
     if (L < 0 || L > 40) return set_error(HIQ_ERR_ARG, std::string(who) + ": bad L");
     if (n_terms < 1 || n_terms > HIQK_MAX_PAULI_TERMS)
          return set_error(HIQ_ERR_ARG, std::string(who) + ": need 1.." + std::to_string(HIQK_MAX_PAULI_TERMS) + " terms");
     if (xmask >> L) return set_error(HIQ_ERR_ARG, std::string(who) + ": flip mask outside the slab");
     std::memset(&p, 0, sizeof(p));
     p.psi = static_cast<const double2*>(slab);
     p.xmask = xmask;
     p.top = 0;
     for (int b = 0; b < L; ++b)
          if ((xmask >> b) & 1ull) p.top = b;
     p.n_terms = n_terms;
     for (int t = 0; t < n_terms; ++t) {
          if (terms[t].zmask >> L) return set_error(HIQ_ERR_ARG, std::string(who) + ": sign mask outside the slab");
          p.zmask[t] = terms[t].zmask;
          p.coef[t] = make_double2(terms[t].re, terms[t].im);
     }
     return HIQ_OK;
}

}  // namespace hiq

using namespace hiq;

extern "C" int hiqk_pauli_expect(const void* slab, int L, uint64_t xmask, const hiqk_pauli_term* terms, int n_terms, const void* src,
                                 uint64_t begin, uint64_t count, double* d_out, void* workspace, void* stream)
{
     PauliParams p;
     int rc = fill_pauli(p, "hiqk_pauli_expect", slab, L, xmask, terms, n_terms);
     if (rc != HIQ_OK) return rc;
     if (!d_out || !workspace) return set_error(HIQ_ERR_ARG, "hiqk_pauli_expect: null output");
     const uint64_t n = 1ull << L;
     if (begin > n || count > n - begin) return set_error(HIQ_ERR_ARG, "hiqk_pauli_expect: range outside the slab");
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     double* partials = static_cast<double*>(workspace);
     unsigned grid;
     const bool whole_local = !src && begin == 0 && count == n;
     if (whole_local && xmask == 0) {
          p.count = n;
          grid = std::min<unsigned>(op_grid(p.count, 4), kOpMaxPartials);
          pauli_expect_kernel<PAULI_DIAG><<<grid, kOpThreads, 0, st>>>(p, partials);
     }
     else if (whole_local) {
          p.count = n >> 1;
          grid = std::min<unsigned>(op_grid(p.count, 2), kOpMaxPartials);
          pauli_expect_kernel<PAULI_PAIRED><<<grid, kOpThreads, 0, st>>>(p, partials);
     }
     else {
          p.begin = begin;
          p.count = count;
          p.src = src ? static_cast<const double2*>(src) : p.psi + begin;
          grid = std::min<unsigned>(op_grid(std::max<uint64_t>(count, 1), 4), kOpMaxPartials);
          pauli_expect_kernel<PAULI_RANGE><<<grid, kOpThreads, 0, st>>>(p, partials);
     }
     fold_pairs_kernel<<<1, 1024, 0, st>>>(partials, static_cast<int>(grid), d_out);
     count_launch(2);
     return check_launch("pauli_expect_kernel");
}

extern "C" int hiqk_pauli_apply(void* slab, int L, uint64_t xmask, const hiqk_pauli_term* terms, int n_terms, void* acc, int accumulate,
                                const void* src, uint64_t begin, uint64_t count, void* stream)
{
     PauliParams p;
     int rc = fill_pauli(p, "hiqk_pauli_apply", slab, L, xmask, terms, n_terms);
     if (rc != HIQ_OK) return rc;
     const uint64_t n = 1ull << L;
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     if (!acc) {
          // in place over the whole slab
          if (src || begin != 0 || count != n) return set_error(HIQ_ERR_ARG, "hiqk_pauli_apply: the in-place form covers the whole slab");
          p.out = static_cast<double2*>(slab);
          if (xmask == 0) {
               p.count = n;
               pauli_apply_kernel<PAULI_DIAG><<<op_grid(p.count, 4), kOpThreads, 0, st>>>(p);
          }
          else {
               p.count = n >> 1;
               pauli_apply_kernel<PAULI_PAIRED><<<op_grid(p.count, 2), kOpThreads, 0, st>>>(p);
          }
     }
     else {
          if (acc == slab) return set_error(HIQ_ERR_ARG, "hiqk_pauli_apply: the accumulator must not alias the slab");
          if (begin > n || count > n - begin) return set_error(HIQ_ERR_ARG, "hiqk_pauli_apply: range outside the slab");
          if (count == 0) return HIQ_OK;
          p.out = static_cast<double2*>(acc);
          p.accumulate = accumulate ? 1 : 0;
          p.begin = begin;
          p.count = count;
          p.src = src ? static_cast<const double2*>(src) : p.psi + begin;
          pauli_apply_kernel<PAULI_RANGE><<<op_grid(count, 4), kOpThreads, 0, st>>>(p);
     }
     count_launch();
     return check_launch("pauli_apply_kernel");
}

// y[j] += a * x[j] for every j with (j & mask) == val  (x may alias y: y *= 1 + a).  The Taylor accumulation and the
// control handling of emulate_time_evolution (ProjectQ simulator.hpp: output_state[j] += update[j] when the control
// bits of j are set; reference call site: _simulator_mpi.py:469-475).
namespace hiq {
__global__ void __launch_bounds__(kOpThreads) axpy_masked_kernel(double2* y, const double2* x, uint64_t n, uint64_t mask, uint64_t val,
                                                                 double2 a)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kOpThreads;
     for (uint64_t j = static_cast<uint64_t>(blockIdx.x) * kOpThreads + threadIdx.x; j < n; j += stride) {
          if ((j & mask) != val) continue;
          const double2 xv = ldg_stream(x + j);
          double2 yv = y[j];
          yv.x = fma(a.x, xv.x, fma(-a.y, xv.y, yv.x));
          yv.y = fma(a.x, xv.y, fma(a.y, xv.x, yv.y));
          y[j] = yv;
     }
}
}  // namespace hiq

extern "C" int hiqk_axpy_masked(void* y, const void* x, int L, uint64_t mask, uint64_t val, double a_re, double a_im, void* stream)
{
     if (!y || !x || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_axpy_masked: bad argument");
     if ((L < 64 && (mask >> L)) || (val & ~mask)) return set_error(HIQ_ERR_ARG, "hiqk_axpy_masked: mask / value outside the slab");
     const uint64_t n = 1ull << L;
     axpy_masked_kernel<<<op_grid(n, 4), kOpThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<double2*>(y), static_cast<const double2*>(x),
                                                                                               n, mask, val, make_double2(a_re, a_im));
     count_launch();
     return check_launch("axpy_masked_kernel");
}

extern "C" int hiq_modinv(uint64_t a, uint64_t N, uint64_t* out)
{
     // a^-1 mod N by the extended Euclidean algorithm; fails when gcd(a, N) != 1
     if (!out || N < 2) return set_error(HIQ_ERR_ARG, "hiq_modinv: bad argument");
     __int128 t = 0, nt = 1, r = N, nr = a % N;
     while (nr != 0) {
          const __int128 q = r / nr;
          const __int128 tt = t - q * nt;
          t = nt;
          nt = tt;
          const __int128 rr = r - q * nr;
          r = nr;
          nr = rr;
     }
     if (r != 1) return set_error(HIQ_ERR_ARG, "hiq_modinv: the constant is not invertible modulo N");
     if (t < 0) t += N;
     *out = static_cast<uint64_t>(t);
     return HIQ_OK;
}

extern "C" int hiqk_permute_gather(void* dst, const void* const* slabs, int n_slabs, int rank, int L, const hiqk_perm* perm, void* stream)
{
     if (!dst || !slabs || !perm) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: null argument");
     if (L < 0 || L > 40 || n_slabs < 1 || n_slabs > kMaxPermSlabs || (n_slabs & (n_slabs - 1)) || rank < 0 || rank >= n_slabs)
          return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: bad L / rank / slab count (a power of two <= 16)");
     if (perm->n_bits < 1 || perm->n_bits > kMaxPermBits) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: register of 1..40 bits");
     int g = 0;
     while ((1 << g) < n_slabs) ++g;
     PermParams p;
     std::memset(&p, 0, sizeof(p));
     p.dst = static_cast<double2*>(dst);
     p.n = 1ull << L;
     p.L = L;
     p.rank_bits = static_cast<uint64_t>(rank) << L;
     p.kind = perm->kind;
     p.n_bits = perm->n_bits;
     p.ctrl_mask = perm->ctrl_mask;
     p.contig = perm->pos[0];
     for (int b = 0; b < perm->n_bits; ++b) {
          const int pos = perm->pos[b];
          if (pos < 0 || pos >= L + g || ((p.reg_mask >> pos) & 1ull))
               return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: register bits must be distinct index bits below L + log2(slabs)");
          p.reg_mask |= 1ull << pos;
          p.pos[b] = static_cast<uint8_t>(pos);
          if (pos != perm->pos[0] + b) p.contig = -1;
     }
     if ((p.ctrl_mask & p.reg_mask) || (p.ctrl_mask >> (L + g))) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: bad control mask");
     // which slabs can be read: ranks that differ from `rank` only in register bits
     const uint64_t reach = p.reg_mask >> L;
     for (int r = 0; r < n_slabs; ++r) {
          const bool needed = ((static_cast<uint64_t>(r ^ rank)) & ~reach) == 0;
          if (needed && !slabs[r]) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: missing the slab of rank " + std::to_string(r));
          p.slabs[r] = static_cast<const double2*>(slabs[r]);
          if (needed && slabs[r] == dst) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: the destination must not alias a source slab");
     }
     const uint64_t space = 1ull << perm->n_bits;
     switch (perm->kind) {
          case HIQK_PERM_TABLE:
               if (!perm->table || perm->n_bits > 32) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: table form needs a table and <= 32 bits");
               p.table = perm->table;
               break;
          case HIQK_PERM_ADD:
               p.N = space;
               p.a = perm->a & (space - 1ull);
               break;
          case HIQK_PERM_ADD_MOD:
               if (perm->N < 1 || perm->N > space) return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: modulus outside the register range");
               p.N = perm->N;
               p.a = perm->a % perm->N;
               break;
          case HIQK_PERM_MUL_MOD: {
               if (perm->N < 2 || perm->N > space || perm->N > (1ull << 32))
                    return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: modulus must be in [2, min(2^n_bits, 2^32)]");
               uint64_t inv = 0;
               if (hiq_modinv(perm->a, perm->N, &inv) != HIQ_OK) return HIQ_ERR_ARG;
               p.N = perm->N;
               p.a = inv;
               break;
          }
          default: return set_error(HIQ_ERR_ARG, "hiqk_permute_gather: unknown permutation kind");
     }
     permute_gather_kernel<<<op_grid(p.n, 4), kOpThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
     count_launch();
     return check_launch("permute_gather_kernel");
}
