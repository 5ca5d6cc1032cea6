// Streaming (HBM-bound) kernels of the state-vector engine: diagonal gates, global phase,
// probability / block-norm / entropy reductions, collapse, fill, qubit compaction and the
// pack/unpack halves of the global<->local qubit swap.  Each launcher cites the reference code it
// replaces; all of them are one pass over (part of) the local slab with 128-bit accesses.
#include <algorithm>
#include <complex>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "hiq_device.cuh"
#include "hiq_host.hpp"

namespace hiq {

using cplx = std::complex<double>;
constexpr int kStreamThreads = 256;
constexpr int kMaxPartials = 8192;  // per-CTA partial sums (x2 for bit_norms)

static unsigned stream_grid(uint64_t items, int per_thread = 4)
{
     const uint64_t need = (items + static_cast<uint64_t>(kStreamThreads) * per_thread - 1) /
                           (static_cast<uint64_t>(kStreamThreads) * per_thread);
     const uint64_t cap = grid_cap(static_cast<uint64_t>(num_sms()) * 8 * 4);  // 8 CTAs/SM resident, 4 waves
     return static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(need, cap)));
}

// ----------------------------------------------------------------------------- diagonal gate
struct DiagParams {
     double2* psi;
     uint64_t n_free;     // indices enumerated (control bits are inserted as fixed ones)
     uint64_t ctrl_mask;
     InsertBits ins;      // control slots, ascending
     int k;
     int slots[kMaxTargets];
     double2 d[1 << kMaxTargets];
};

__global__ void __launch_bounds__(kStreamThreads) diag_kernel(const __grid_constant__ DiagParams p)
{
     __shared__ double2 lut[1 << kMaxTargets];
     if (threadIdx.x < (1 << p.k)) lut[threadIdx.x] = p.d[threadIdx.x];
     __syncthreads();
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     uint64_t f = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x;
     // four independent amplitudes in flight per thread
     for (; f + 3 * stride < p.n_free; f += 4 * stride) {
          uint64_t idx[4];
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
               idx[u] = insert_zero_bits(f + u * stride, p.ins) | p.ctrl_mask;
               v[u] = ldg_stream(p.psi + idx[u]);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
               int sel = 0;
#pragma unroll
               for (int l = 0; l < kMaxTargets; ++l)
                    if (l < p.k) sel |= static_cast<int>((idx[u] >> p.slots[l]) & 1ull) << l;
               p.psi[idx[u]] = cmul(v[u], lut[sel]);
          }
     }
     for (; f < p.n_free; f += stride) {
          const uint64_t idx = insert_zero_bits(f, p.ins) | p.ctrl_mask;
          int sel = 0;
          for (int l = 0; l < p.k; ++l) sel |= static_cast<int>((idx >> p.slots[l]) & 1ull) << l;
          p.psi[idx] = cmul(p.psi[idx], lut[sel]);
     }
}

// ----------------------------------------------------------------------------- batched diagonal gates
// One HBM pass for up to kMaxDiagOps diagonal fused gates (DiagProg, hiq_device.cuh).  Index =
// (chunk bits | u bits | tid bits): tid = index bits 0..7 (one coalesced 4 KiB run per (chunk, u)),
// the u positions are the index bits >= 8 that the ops touch least.
struct DiagBatchParams {
     double2* psi;
     uint64_t n;         // amplitudes
     uint64_t n_chunks;
     int n_u;            // u bits in use
     InsertBits ins;     // the u positions, ascending (zero bits inserted into chunk << 8)
     uint64_t uoff[1 << kMaxUBits];
     DiagProg prog;
};

__global__ void __launch_bounds__(kStreamThreads, 4) diag_batch_kernel(const __grid_constant__ DiagBatchParams p)
{
     __shared__ DiagShared sh;
     uint32_t selt[kMaxDiagOps / 4];
     diag_prog_init<kStreamThreads>(p.prog, sh, threadIdx.x, selt);
     const int nu = 1 << p.n_u;
     for (uint64_t chunk = blockIdx.x; chunk < p.n_chunks; chunk += gridDim.x) {
          const uint64_t cidx = insert_zero_bits(chunk << 8, p.ins);
          diag_prog_chunk(p.prog, sh, cidx);
          const double2 s0 = diag_prog_s0(p.prog, sh, selt);
          const uint64_t base = cidx | threadIdx.x;
#pragma unroll 1
          for (int u0 = 0; u0 < nu; u0 += 4) {
               uint64_t idx[4];
               double2 v[4], f[4];
               bool ok[4];
#pragma unroll
               for (int u = 0; u < 4; ++u) {
                    idx[u] = base | p.uoff[(u0 + u) & ((1 << kMaxUBits) - 1)];
                    ok[u] = (u0 + u < nu) && idx[u] < p.n;
                    if (ok[u]) v[u] = ldg_stream(p.psi + idx[u]);
               }
#pragma unroll
               for (int u = 0; u < 4; ++u) f[u] = p.prog.n_s1 ? diag_prog_s1(p.prog, sh, selt, (u0 + u) & ((1 << kMaxUBits) - 1), s0) : s0;
#pragma unroll
               for (int u = 0; u < 4; ++u)
                    if (ok[u]) p.psi[idx[u]] = cmul(v[u], f[u]);
          }
     }
}

// ----------------------------------------------------------------------------- scale / fill / collapse
__global__ void __launch_bounds__(kStreamThreads) scale_kernel(double2* psi, uint64_t n, double2 s)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x;
     for (; i + 3 * stride < n; i += 4 * stride) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ldg_stream(psi + i + u * stride);
#pragma unroll
          for (int u = 0; u < 4; ++u) psi[i + u * stride] = cmul(v[u], s);
     }
     for (; i < n; i += stride) psi[i] = cmul(psi[i], s);
}

__global__ void __launch_bounds__(kStreamThreads) fill_kernel(double2* psi, uint64_t n, double2 v)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x; i < n; i += stride) psi[i] = v;
}

// psi[i] = match ? psi[i] * scale : 0 — amplitudes that do not match are overwritten without being read
__global__ void __launch_bounds__(kStreamThreads) collapse_kernel(double2* psi, uint64_t n, uint64_t mask, uint64_t val,
                                                                   double scale)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x;
     for (; i + 3 * stride < n; i += 4 * stride) {
          double2 v[4];
          bool ok[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
               ok[u] = ((i + u * stride) & mask) == val;
               v[u] = ok[u] ? ldg_stream(psi + i + u * stride) : make_double2(0.0, 0.0);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) psi[i + u * stride] = make_double2(v[u].x * scale, v[u].y * scale);
     }
     for (; i < n; i += stride) {
          const bool ok = (i & mask) == val;
          const double2 v = ok ? psi[i] : make_double2(0.0, 0.0);
          psi[i] = make_double2(v.x * scale, v.y * scale);
     }
}

// ----------------------------------------------------------------------------- reductions
// Stage 1: one partial per CTA (fixed tree); stage 2: one CTA folds the partials in index order.
struct ReduceParams {
     const double2* psi;
     uint64_t n_free;
     uint64_t val;
     InsertBits ins;  // mask bits, ascending (fixed to `val`)
};

enum { RED_NORM = 0, RED_ENTROPY = 1 };

template <int MODE>
__global__ void __launch_bounds__(kStreamThreads) reduce_kernel(const __grid_constant__ ReduceParams p, double* partials)
{
     __shared__ double scratch[kStreamThreads / 32];
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     double acc[4] = {0.0, 0.0, 0.0, 0.0};
     uint64_t f = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x;
     auto term = [](double2 v) {
          const double pr = norm2(v);
          if (MODE == RED_NORM) return pr;
          return pr > 0.0 ? pr * log2(pr) : 0.0;
     };
     for (; f + 3 * stride < p.n_free; f += 4 * stride) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ldg_stream(p.psi + (insert_zero_bits(f + u * stride, p.ins) | p.val));
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += term(v[u]);
     }
     for (; f < p.n_free; f += stride) acc[0] += term(p.psi[insert_zero_bits(f, p.ins) | p.val]);
     const double s = block_sum<kStreamThreads>((acc[0] + acc[1]) + (acc[2] + acc[3]), scratch);
     if (threadIdx.x == 0) partials[blockIdx.x] = s;
}

__global__ void __launch_bounds__(1024) fold_kernel(const double* partials, int n, int n_out, double* out)
{
     // out[o] = sum_j partials[o * n + j]: fixed order -> run-to-run deterministic
     __shared__ double scratch[32];
     for (int o = 0; o < n_out; ++o) {
          double acc = 0.0;
          for (int j = threadIdx.x; j < n; j += 1024) acc += partials[o * n + j];
          const double s = block_sum<1024>(acc, scratch);
          if (threadIdx.x == 0) out[o] = s;
     }
}

__global__ void __launch_bounds__(kStreamThreads) bit_norms_kernel(const double2* psi, uint64_t n, int slot, double* partials)
{
     __shared__ double scratch[kStreamThreads / 32];
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     double a0 = 0.0, a1 = 0.0;
     for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x; i < n; i += stride) {
          const double pr = norm2(ldg_stream(psi + i));
          if ((i >> slot) & 1ull) a1 += pr;
          else a0 += pr;
     }
     const double s0 = block_sum<kStreamThreads>(a0, scratch);
     const double s1 = block_sum<kStreamThreads>(a1, scratch);
     if (threadIdx.x == 0) {
          partials[blockIdx.x] = s0;
          partials[gridDim.x + blockIdx.x] = s1;
     }
}

// one CTA per block of `len` consecutive amplitudes (len >= kStreamThreads)
__global__ void __launch_bounds__(kStreamThreads) block_norms_cta_kernel(const double2* psi, uint64_t len, double* out)
{
     __shared__ double scratch[kStreamThreads / 32];
     const double2* p = psi + static_cast<uint64_t>(blockIdx.x) * len;
     double acc[4] = {0.0, 0.0, 0.0, 0.0};
     uint64_t i = threadIdx.x;
     for (; i + 3 * kStreamThreads < len; i += 4 * kStreamThreads) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ldg_stream(p + i + u * kStreamThreads);
#pragma unroll
          for (int u = 0; u < 4; ++u) acc[u] += norm2(v[u]);
     }
     for (; i < len; i += kStreamThreads) acc[0] += norm2(p[i]);
     const double s = block_sum<kStreamThreads>((acc[0] + acc[1]) + (acc[2] + acc[3]), scratch);
     if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// small blocks: one thread per block, sequential like the reference loop
__global__ void block_norms_thread_kernel(const double2* psi, uint64_t len, uint64_t n_blocks, double* out)
{
     const uint64_t b = static_cast<uint64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
     if (b >= n_blocks) return;
     double acc = 0.0;
     for (uint64_t j = 0; j < len; ++j) acc += norm2(psi[b * len + j]);
     out[b] = acc;
}

// ----------------------------------------------------------------------------- compaction
__global__ void __launch_bounds__(kStreamThreads) compact_gather_kernel(const double2* src, double2* scratch, uint64_t j0,
                                                                         uint64_t count, int slot, uint64_t keep_bit)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     const uint64_t low_mask = (1ull << slot) - 1ull;
     for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x; i < count; i += stride) {
          const uint64_t j = j0 + i;
          const uint64_t s = ((j >> slot) << (slot + 1)) | keep_bit | (j & low_mask);
          scratch[i] = ldg_stream(src + s);
     }
}

__global__ void __launch_bounds__(kStreamThreads) copy_kernel(const double2* src, double2* dst, uint64_t count)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x; i < count; i += stride)
          dst[i] = src[i];
}

// ----------------------------------------------------------------------------- swap pack / unpack
struct SwapParams {
     double2* psi;
     double2* buf;
     uint64_t begin, count;
     uint64_t pat_bits;  // pattern spread onto the swapped slots
     InsertBits ins;     // swapped slots, ascending
};

template <bool PACK>
__global__ void __launch_bounds__(kStreamThreads) swap_pack_kernel(const __grid_constant__ SwapParams p)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kStreamThreads;
     for (uint64_t i = static_cast<uint64_t>(blockIdx.x) * kStreamThreads + threadIdx.x; i < p.count; i += stride) {
          const uint64_t idx = insert_zero_bits(p.begin + i, p.ins) | p.pat_bits;
          if (PACK) p.buf[i] = ldg_stream(p.psi + idx);
          else p.psi[idx] = ldg_stream(p.buf + i);
     }
}

static InsertBits bits_of_mask(uint64_t mask)
{
     InsertBits ib;
     std::memset(&ib, 0, sizeof(ib));
     for (int s = 0; s < 64; ++s)
          if ((mask >> s) & 1ull) ib.pos[ib.n++] = static_cast<uint8_t>(s);
     return ib;
}

}  // namespace hiq

using namespace hiq;

extern "C" int hiqk_apply_diag(void* slab, int L, int k, const int* slots, const double* diag, uint64_t ctrl_mask,
                               void* stream)
{
     if (!slab || !slots || !diag) return set_error(HIQ_ERR_ARG, "hiqk_apply_diag: null argument");
     if (k < 1 || k > kMaxTargets || L < k || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_apply_diag: bad k or L");
     DiagParams p;
     std::memset(&p, 0, sizeof(p));
     uint64_t tmask = 0;
     for (int l = 0; l < k; ++l) {
          if (slots[l] < 0 || slots[l] >= L || ((tmask >> slots[l]) & 1))
               return set_error(HIQ_ERR_ARG, "hiqk_apply_diag: target slots must be distinct and < L");
          tmask |= 1ull << slots[l];
          p.slots[l] = slots[l];
     }
     if ((ctrl_mask & tmask) || (ctrl_mask >> L)) return set_error(HIQ_ERR_ARG, "hiqk_apply_diag: bad control mask");
     p.psi = static_cast<double2*>(slab);
     p.k = k;
     p.ctrl_mask = ctrl_mask;
     p.ins = bits_of_mask(ctrl_mask);
     p.n_free = 1ull << (L - p.ins.n);
     std::memcpy(p.d, diag, sizeof(double2) << k);
     diag_kernel<<<stream_grid(p.n_free), kStreamThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
     count_launch();
     return check_launch("diag_kernel");
}

// Picks up to `want` index positions for the u part: positions < L outside `exclude`, the ones the ops
// touch least first.  Returns how many were found (ascending order in `out`).
int hiq::choose_u_positions(int L, const hiqk_diag_op* ops, int n_ops, uint64_t exclude, int want, int* out)
{
     int touches[64] = {0};
     for (int j = 0; j < n_ops; ++j)
          for (int l = 0; l < ops[j].k && l < kMaxTargets; ++l)
               if (ops[j].slots[l] >= 0 && ops[j].slots[l] < 64) ++touches[ops[j].slots[l]];
     std::vector<int> cand;
     for (int pos = 0; pos < L; ++pos)
          if (!((exclude >> pos) & 1ull)) cand.push_back(pos);
     std::stable_sort(cand.begin(), cand.end(), [&](int a, int c) { return touches[a] < touches[c]; });
     const int n = std::min<int>(want, static_cast<int>(cand.size()));
     std::sort(cand.begin(), cand.begin() + n);
     for (int i = 0; i < n; ++i) out[i] = cand[i];
     return n;
}

// Validates `ops` and builds the kernel-side program.  upos[b] = index position of u bit b;
// target_mask = dense targets (ops touching them become class E); tid_mask = index bits spelled by threadIdx.  order_out[j] = original index of
// the op placed at position j.
int hiq::build_diag_prog(DiagProg& p, int L, const hiqk_diag_op* ops, int n_ops, const int* upos, int n_u, uint64_t target_mask,
                         uint64_t tid_mask, int* order_out, const char* who)
{
     if (!ops || n_ops < 1 || n_ops > kMaxDiagOps)
          return set_error(HIQ_ERR_ARG, std::string(who) + ": need 1.." + std::to_string(kMaxDiagOps) + " diagonal ops");
     std::memset(&p, 0, sizeof(p));
     uint64_t umask = 0;
     for (int b = 0; b < n_u; ++b) umask |= 1ull << upos[b];
     std::vector<int> cls(n_ops);
     for (int j = 0; j < n_ops; ++j) {
          const hiqk_diag_op& o = ops[j];
          if (o.k < 0 || o.k > kMaxTargets) return set_error(HIQ_ERR_ARG, std::string(who) + ": diagonal op with k outside 0..5");
          uint64_t seen = 0;
          for (int l = 0; l < o.k; ++l) {
               if (o.slots[l] < 0 || o.slots[l] >= L || ((seen >> o.slots[l]) & 1))
                    return set_error(HIQ_ERR_ARG, std::string(who) + ": diagonal op slots must be distinct and < L");
               seen |= 1ull << o.slots[l];
          }
          // 0: CTA-uniform per chunk, 1: per thread per chunk, 2: per element (u-dependent), 3: touches dense targets
          cls[j] = (seen & target_mask) ? 3 : ((seen & umask) ? 2 : ((seen & tid_mask) ? 1 : 0));
     }
     std::vector<int> order(n_ops);
     for (int j = 0; j < n_ops; ++j) order[j] = j;
     std::stable_sort(order.begin(), order.end(), [&](int a, int c) { return cls[a] < cls[c]; });
     p.n = n_ops;
     for (int j = 0; j < n_ops; ++j) {
          const hiqk_diag_op& o = ops[order[j]];
          if (order_out) order_out[j] = order[j];
          if (cls[order[j]] == 0) ++p.n_s0a;
          if (cls[order[j]] <= 1) ++p.n_s0;
          else if (cls[order[j]] == 2) ++p.n_s1;
          else ++p.n_e;
          for (int l = 0; l < 8; ++l) p.slots[j][l] = l < o.k ? static_cast<uint8_t>(o.slots[l]) : 63;  // bit 63 is always 0
          for (int u = 0; u < (1 << n_u); ++u) {
               uint32_t sel = 0;
               for (int l = 0; l < o.k; ++l)
                    for (int b = 0; b < n_u; ++b)
                         if (o.slots[l] == upos[b] && ((u >> b) & 1)) sel |= 1u << l;
               p.usel[j][u] = static_cast<uint8_t>(sel);
          }
          // entries beyond 2^k are never selected (their selector bits read as 0)
          std::memcpy(p.lut[j], o.lut, sizeof(double2) << o.k);
     }
     return HIQ_OK;
}

namespace hiq {
// host part of the batched diagonal launch: index split (tid = bits 0..7 | u | chunk) and the diagonal program
static int fill_diag_batch(DiagBatchParams& p, double2* psi, int L, const hiqk_diag_op* ops, int n_ops, const char* who)
{
     int upos[kMaxUBits];
     const int n_u = choose_u_positions(L, ops, n_ops, 0xffull, kMaxUBits, upos);
     const int rc = build_diag_prog(p.prog, L, ops, n_ops, upos, n_u, 0, 0xffull, nullptr, who);
     if (rc != HIQ_OK) return rc;
     p.psi = psi;
     p.n = 1ull << L;
     p.n_u = n_u;
     p.n_chunks = L > 8 + n_u ? 1ull << (L - 8 - n_u) : 1;
     std::memset(&p.ins, 0, sizeof(p.ins));
     p.ins.n = n_u;
     for (int b = 0; b < n_u; ++b) p.ins.pos[b] = static_cast<uint8_t>(upos[b]);
     for (int u = 0; u < (1 << kMaxUBits); ++u) {
          uint64_t o = 0;
          for (int b = 0; b < n_u; ++b)
               if ((u >> b) & 1) o |= 1ull << upos[b];
          p.uoff[u] = o;
     }
     return HIQ_OK;
}
constexpr int kBatchImageHeaderWords = 32;
}  // namespace hiq

extern "C" size_t hiqk_diag_batch_image_bytes(void) { return kBatchImageHeaderWords * sizeof(uint32_t) + sizeof(DiagBatchParams); }

// Host-only: the kernel parameters hiqk_apply_diag_batch would launch with (see include/hiq_b200.h).
extern "C" int hiqk_diag_batch_image(int L, const hiqk_diag_op* ops, int n_ops, void* image, size_t image_bytes)
{
     if (!image || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_diag_batch_image: bad argument");
     if (!ops || n_ops < 1 || n_ops > kMaxDiagOps) return set_error(HIQ_ERR_ARG, "hiqk_diag_batch_image: need 1..16 diagonal ops");
     if (image_bytes < hiqk_diag_batch_image_bytes()) return set_error(HIQ_ERR_ARG, "hiqk_diag_batch_image: buffer too small");
     std::memset(image, 0, hiqk_diag_batch_image_bytes());
     uint32_t* head = static_cast<uint32_t*>(image);
     DiagBatchParams* p = reinterpret_cast<DiagBatchParams*>(head + kBatchImageHeaderWords);
     const int rc = fill_diag_batch(*p, nullptr, L, ops, n_ops, "hiqk_diag_batch_image");
     if (rc != HIQ_OK) return rc;
     int w = 0;
     head[w++] = 0x42445148u;  // 'HQDB'
     head[w++] = kStreamThreads;
     head[w++] = sizeof(DiagBatchParams);
     head[w++] = kMaxDiagOps;
     head[w++] = 1 << kMaxUBits;
     head[w++] = 1 << kMaxTargets;
     head[w++] = sizeof(InsertBits);
     head[w++] = static_cast<uint32_t>(offsetof(DiagBatchParams, n));
     head[w++] = static_cast<uint32_t>(offsetof(DiagBatchParams, n_chunks));
     head[w++] = static_cast<uint32_t>(offsetof(DiagBatchParams, n_u));
     head[w++] = static_cast<uint32_t>(offsetof(DiagBatchParams, ins));
     head[w++] = static_cast<uint32_t>(offsetof(DiagBatchParams, uoff));
     head[w++] = static_cast<uint32_t>(offsetof(DiagBatchParams, prog));
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

extern "C" int hiqk_apply_diag_batch(void* slab, int L, const hiqk_diag_op* ops, int n_ops, void* stream)
{
     if (!slab || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_apply_diag_batch: bad argument");
     if (!ops || n_ops < 1 || n_ops > kMaxDiagOps) return set_error(HIQ_ERR_ARG, "hiqk_apply_diag_batch: need 1..16 diagonal ops");
     static DiagBatchParams p;  // ~9 KB, launches are issued from one host thread per engine
     static std::mutex mu;
     std::lock_guard<std::mutex> lock(mu);
     const int rc = fill_diag_batch(p, static_cast<double2*>(slab), L, ops, n_ops, "hiqk_apply_diag_batch");
     if (rc != HIQ_OK) return rc;
     const unsigned grid = static_cast<unsigned>(std::min<uint64_t>(p.n_chunks, grid_cap(static_cast<uint64_t>(num_sms()) * 4 * 8)));
     diag_batch_kernel<<<grid, kStreamThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
     count_launch();
     return check_launch("diag_batch_kernel");
}

extern "C" int hiqk_scale(void* slab, int L, double re, double im, void* stream)
{
     if (!slab || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_scale: bad argument");
     const uint64_t n = 1ull << L;
     scale_kernel<<<stream_grid(n), kStreamThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<double2*>(slab), n,
                                                                                            make_double2(re, im));
     count_launch();
     return check_launch("scale_kernel");
}

extern "C" size_t hiqk_workspace_bytes(void) { return sizeof(double) * 2 * kMaxPartials; }

static int launch_reduce(int mode, const void* slab, int L, uint64_t mask, uint64_t val, double* d_out, void* workspace,
                         void* stream)
{
     if (!slab || !d_out || !workspace || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "reduce: bad argument");
     if ((mask >> L) || (val & ~mask)) return set_error(HIQ_ERR_ARG, "reduce: mask/val outside the slab");
     ReduceParams p;
     p.psi = static_cast<const double2*>(slab);
     p.ins = bits_of_mask(mask);
     p.n_free = 1ull << (L - p.ins.n);
     p.val = val;
     const unsigned grid = std::min<unsigned>(stream_grid(p.n_free), kMaxPartials);
     double* partials = static_cast<double*>(workspace);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     if (mode == RED_NORM) reduce_kernel<RED_NORM><<<grid, kStreamThreads, 0, st>>>(p, partials);
     else reduce_kernel<RED_ENTROPY><<<grid, kStreamThreads, 0, st>>>(p, partials);
     fold_kernel<<<1, 1024, 0, st>>>(partials, static_cast<int>(grid), 1, d_out);
     count_launch(2);
     return check_launch("reduce_kernel");
}

extern "C" int hiqk_prob_masked(const void* slab, int L, uint64_t mask, uint64_t val, double* d_out, void* workspace,
                                void* stream)
{
     return launch_reduce(RED_NORM, slab, L, mask, val, d_out, workspace, stream);
}

extern "C" int hiqk_entropy(const void* slab, int L, double* d_out, void* workspace, void* stream)
{
     return launch_reduce(RED_ENTROPY, slab, L, 0, 0, d_out, workspace, stream);
}

extern "C" int hiqk_bit_norms(const void* slab, int L, int slot, double* d_out, void* workspace, void* stream)
{
     if (!slab || !d_out || !workspace || L < 1 || L > 40 || slot < 0 || slot >= L)
          return set_error(HIQ_ERR_ARG, "hiqk_bit_norms: bad argument");
     const uint64_t n = 1ull << L;
     const unsigned grid = std::min<unsigned>(stream_grid(n), kMaxPartials);
     double* partials = static_cast<double*>(workspace);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     bit_norms_kernel<<<grid, kStreamThreads, 0, st>>>(static_cast<const double2*>(slab), n, slot, partials);
     fold_kernel<<<1, 1024, 0, st>>>(partials, static_cast<int>(grid), 2, d_out);
     count_launch(2);
     return check_launch("bit_norms_kernel");
}

extern "C" int hiqk_block_norms(const void* slab, int L, uint64_t n_blocks, double* d_out, void* stream)
{
     if (!slab || !d_out || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_block_norms: bad argument");
     const uint64_t n = 1ull << L;
     if (n_blocks == 0 || (n_blocks & (n_blocks - 1)) || n_blocks > n || n_blocks > (1ull << 20))
          return set_error(HIQ_ERR_ARG, "hiqk_block_norms: n_blocks must be a power of two <= 2^L");
     const uint64_t len = n / n_blocks;
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     const double2* psi = static_cast<const double2*>(slab);
     if (len >= kStreamThreads) block_norms_cta_kernel<<<static_cast<unsigned>(n_blocks), kStreamThreads, 0, st>>>(psi, len, d_out);
     else block_norms_thread_kernel<<<static_cast<unsigned>((n_blocks + 127) / 128), 128, 0, st>>>(psi, len, n_blocks, d_out);
     count_launch();
     return check_launch("block_norms_kernel");
}

extern "C" int hiqk_collapse(void* slab, int L, uint64_t mask, uint64_t val, double scale, void* stream)
{
     if (!slab || L < 0 || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_collapse: bad argument");
     const uint64_t n = 1ull << L;
     collapse_kernel<<<stream_grid(n), kStreamThreads, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<double2*>(slab), n,
                                                                                               mask, val, scale);
     count_launch();
     return check_launch("collapse_kernel");
}

extern "C" int hiqk_fill(void* slab, uint64_t begin, uint64_t count, double re, double im, void* stream)
{
     if (!slab) return set_error(HIQ_ERR_ARG, "hiqk_fill: null slab");
     if (count == 0) return HIQ_OK;
     fill_kernel<<<stream_grid(count), kStreamThreads, 0, static_cast<cudaStream_t>(stream)>>>(
         static_cast<double2*>(slab) + begin, count, make_double2(re, im));
     count_launch();
     return check_launch("fill_kernel");
}

extern "C" int hiqk_compact_bit(void* slab, int L, int slot, int keep, void* scratch, uint64_t scratch_amps, void* stream)
{
     if (!slab || !scratch || L < 1 || L > 40 || slot < 0 || slot >= L || scratch_amps == 0)
          return set_error(HIQ_ERR_ARG, "hiqk_compact_bit: bad argument");
     const uint64_t half = 1ull << (L - 1);
     double2* psi = static_cast<double2*>(slab);
     double2* tmp = static_cast<double2*>(scratch);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     const uint64_t keep_bit = keep ? (1ull << slot) : 0ull;
     // Chunks are processed in index order: chunk c only reads source indices >= its first
     // destination index, which no earlier chunk has written.
     for (uint64_t j0 = 0; j0 < half; j0 += scratch_amps) {
          const uint64_t cnt = std::min(scratch_amps, half - j0);
          compact_gather_kernel<<<stream_grid(cnt), kStreamThreads, 0, st>>>(psi, tmp, j0, cnt, slot, keep_bit);
          copy_kernel<<<stream_grid(cnt), kStreamThreads, 0, st>>>(tmp, psi + j0, cnt);
          count_launch(2);
     }
     return check_launch("compact_kernel");
}

static int launch_swap(bool pack, void* slab, int L, int q, const int* slots, uint64_t pat, uint64_t begin, uint64_t count,
                       void* buf, void* stream)
{
     if (!slab || !buf || !slots || q < 1 || q > L || L > 40) return set_error(HIQ_ERR_ARG, "hiqk_swap_pack: bad argument");
     SwapParams p;
     std::memset(&p, 0, sizeof(p));
     uint64_t mask = 0;
     std::vector<int> sorted(slots, slots + q);
     std::sort(sorted.begin(), sorted.end());
     for (int i = 0; i < q; ++i) {
          if (sorted[i] < 0 || sorted[i] >= L || ((mask >> sorted[i]) & 1)) return set_error(HIQ_ERR_ARG, "hiqk_swap_pack: bad slots");
          mask |= 1ull << sorted[i];
          if ((pat >> i) & 1ull) p.pat_bits |= 1ull << sorted[i];
     }
     if (begin + count > (1ull << (L - q))) return set_error(HIQ_ERR_ARG, "hiqk_swap_pack: range outside the slab");
     if (count == 0) return HIQ_OK;
     p.psi = static_cast<double2*>(slab);
     p.buf = static_cast<double2*>(buf);
     p.begin = begin;
     p.count = count;
     p.ins = bits_of_mask(mask);
     cudaStream_t st = static_cast<cudaStream_t>(stream);
     if (pack) swap_pack_kernel<true><<<stream_grid(count), kStreamThreads, 0, st>>>(p);
     else swap_pack_kernel<false><<<stream_grid(count), kStreamThreads, 0, st>>>(p);
     count_launch();
     return check_launch("swap_pack_kernel");
}

extern "C" int hiqk_swap_pack(const void* slab, int L, int q, const int* slots, uint64_t pat, uint64_t begin, uint64_t count,
                              void* dst, void* stream)
{
     return launch_swap(true, const_cast<void*>(slab), L, q, slots, pat, begin, count, dst, stream);
}

extern "C" int hiqk_swap_unpack(void* slab, int L, int q, const int* slots, uint64_t pat, uint64_t begin, uint64_t count,
                                const void* src, void* stream)
{
     return launch_swap(false, slab, L, q, slots, pat, begin, count, const_cast<void*>(src), stream);
}
