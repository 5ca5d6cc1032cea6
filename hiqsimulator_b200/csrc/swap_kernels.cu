// In-place global<->local qubit swap over peer-mapped slabs (NVLink 5 / NVSwitch load-store).
//
// Replaces the reference's pack -> mpi::all_to_all -> unpack pipeline (reference:
// src/simulator-mpi/SwapperMT.cpp:30-126, swapping.hpp:33-68) for the data that has to cross GPUs.
// The swap transposes global-index bit (L + gpos_i) with local slot_i: rank r's amplitudes
// whose swapped slots spell peer p's bits trade places with p's amplitudes whose swapped slots spell
// r's bits, same free index on both sides.  The transpositions are pairwise disjoint, so ONE kernel per
// GPU does the exchange in place with no staging buffer: for every pair it owns it loads its own value
// and the peer's (P2P load), then stores them crosswise (P2P store).  The two GPUs of a pair split the
// free-index range, so each direction of every link carries half loads' responses and half stores.
#include <algorithm>
#include <cstring>
#include <vector>

#include "hiq_device.cuh"
#include "hiq_host.hpp"

namespace hiq {

constexpr int kSwapThreads = 256;
constexpr int kMaxSwapPeers = 7;  // up to 3 swapped global qubits per kernel

struct SwapP2PParams {
     double2* local;
     double2* peer[kMaxSwapPeers];
     int n_peers;
     uint64_t begin[kMaxSwapPeers];      // free-index range this GPU handles for peer k
     uint64_t count[kMaxSwapPeers];
     uint64_t mine_bits[kMaxSwapPeers];  // peer k's pattern spread onto the swapped slots (my elements that leave)
     uint64_t theirs_bits;               // my pattern spread onto the swapped slots (the peer's elements that arrive)
     InsertBits ins;                     // swapped slots, ascending
};

__global__ void __launch_bounds__(kSwapThreads) swap_p2p_kernel(const __grid_constant__ SwapP2PParams p)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kSwapThreads;
     for (int k = 0; k < p.n_peers; ++k) {
          double2* __restrict__ remote = p.peer[k];
          const uint64_t mine = p.mine_bits[k];
          const uint64_t n = p.count[k];
          uint64_t g = static_cast<uint64_t>(blockIdx.x) * kSwapThreads + threadIdx.x;
          for (; g + 3 * stride < n; g += 4 * stride) {
               uint64_t base[4];
               double2 a[4], b[4];
#pragma unroll
               for (int u = 0; u < 4; ++u) {
                    base[u] = insert_zero_bits(p.begin[k] + g + u * stride, p.ins);
                    a[u] = ldg_stream(p.local + (base[u] | mine));
                    b[u] = ldg_stream(remote + (base[u] | p.theirs_bits));
               }
#pragma unroll
               for (int u = 0; u < 4; ++u) {
                    p.local[base[u] | mine] = b[u];
                    remote[base[u] | p.theirs_bits] = a[u];
               }
          }
          for (; g < n; g += stride) {
               const uint64_t base = insert_zero_bits(p.begin[k] + g, p.ins);
               const double2 a = ldg_stream(p.local + (base | mine));
               const double2 b = ldg_stream(remote + (base | p.theirs_bits));
               p.local[base | mine] = b;
               remote[base | p.theirs_bits] = a;
          }
     }
}

// Packed exchange (low swapped slots): the amplitudes bound for peer k are gathered from the local slab — strided by
// 2^(lowest swapped slot), at HBM speed through L2 — into a contiguous run of a staging buffer, which may live in the
// PEER's memory (push: the NVLink traffic is full-line posted writes whatever the slots are) or in local memory (pull:
// the peer reads it with contiguous loads).  The inverse scatter fills the local slab from a staging buffer.  One launch
// moves the pieces of all peers (blockIdx.y = peer); four independent 128-bit accesses per thread are in flight.
struct SwapMoveParams {
     double2* psi;
     int n_peers;
     double2* buf[kMaxSwapPeers];
     uint64_t pat_bits[kMaxSwapPeers];  // peer k's pattern spread onto the swapped slots
     uint64_t begin, count;             // free-index range of this piece
     InsertBits ins;                    // swapped slots, ascending
};

template <bool PACK>
__global__ void __launch_bounds__(kSwapThreads) swap_move_kernel(const __grid_constant__ SwapMoveParams p)
{
     const int k = blockIdx.y;
     double2* __restrict__ buf = p.buf[k];
     const uint64_t pat = p.pat_bits[k];
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * kSwapThreads;
     uint64_t i = static_cast<uint64_t>(blockIdx.x) * kSwapThreads + threadIdx.x;
     for (; i + 3 * stride < p.count; i += 4 * stride) {
          uint64_t idx[4];
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
               idx[u] = insert_zero_bits(p.begin + i + u * stride, p.ins) | pat;
               v[u] = PACK ? ldg_stream(p.psi + idx[u]) : ldg_stream(buf + i + u * stride);
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
               if (PACK) buf[i + u * stride] = v[u];
               else p.psi[idx[u]] = v[u];
          }
     }
     for (; i < p.count; i += stride) {
          const uint64_t idx = insert_zero_bits(p.begin + i, p.ins) | pat;
          if (PACK) buf[i] = ldg_stream(p.psi + idx);
          else p.psi[idx] = ldg_stream(buf + i);
     }
}

}  // namespace hiq

using namespace hiq;

extern "C" int hiqk_swap_move(void* slab, int L, int q, const int* slots, int n_peers, const uint64_t* peer_pats, uint64_t begin,
                              uint64_t count, void* const* bufs, int pack, void* stream)
{
     if (!slab || !slots || !peer_pats || !bufs) return set_error(HIQ_ERR_ARG, "hiqk_swap_move: null argument");
     if (q < 1 || q > L || L > 40 || n_peers < 1 || n_peers > kMaxSwapPeers || n_peers > (1 << q) - 1)
          return set_error(HIQ_ERR_ARG, "hiqk_swap_move: bad q / n_peers");
     if (begin + count > (1ull << (L - q))) return set_error(HIQ_ERR_ARG, "hiqk_swap_move: range outside the slab");
     if (count == 0) return HIQ_OK;
     SwapMoveParams p;
     std::memset(&p, 0, sizeof(p));
     std::vector<int> sorted(slots, slots + q);
     std::sort(sorted.begin(), sorted.end());
     uint64_t mask = 0;
     for (int i = 0; i < q; ++i) {
          if (sorted[i] < 0 || sorted[i] >= L || ((mask >> sorted[i]) & 1)) return set_error(HIQ_ERR_ARG, "hiqk_swap_move: bad slots");
          mask |= 1ull << sorted[i];
          p.ins.pos[p.ins.n++] = static_cast<uint8_t>(sorted[i]);
     }
     p.psi = static_cast<double2*>(slab);
     p.n_peers = n_peers;
     p.begin = begin;
     p.count = count;
     for (int k = 0; k < n_peers; ++k) {
          if (!bufs[k]) return set_error(HIQ_ERR_ARG, "hiqk_swap_move: null staging buffer");
          p.buf[k] = static_cast<double2*>(bufs[k]);
          for (int i = 0; i < q; ++i)
               if ((peer_pats[k] >> i) & 1ull) p.pat_bits[k] |= 1ull << sorted[i];
     }
     const uint64_t need = (count + static_cast<uint64_t>(kSwapThreads) * 4 - 1) / (static_cast<uint64_t>(kSwapThreads) * 4);
     const uint64_t per_peer = std::max<uint64_t>(1, grid_cap(static_cast<uint64_t>(num_sms()) * 8) / n_peers);
     dim3 grid(static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(need, per_peer))), static_cast<unsigned>(n_peers));
     if (pack) swap_move_kernel<true><<<grid, kSwapThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
     else swap_move_kernel<false><<<grid, kSwapThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
     count_launch();
     return check_launch("swap_move_kernel");
}

extern "C" int hiqk_swap_p2p(void* local, void* const* peer_slabs, int n_peers, int L, int q, const int* slots,
                             const uint64_t* peer_pats, uint64_t my_pat, const uint64_t* begin, const uint64_t* count, void* stream)
{
     if (!local || !peer_slabs || !slots || !peer_pats || !begin || !count) return set_error(HIQ_ERR_ARG, "hiqk_swap_p2p: null argument");
     if (q < 1 || q > L || L > 40 || n_peers < 1 || n_peers > kMaxSwapPeers || n_peers > (1 << q) - 1)
          return set_error(HIQ_ERR_ARG, "hiqk_swap_p2p: bad q / n_peers");
     SwapP2PParams p;
     std::memset(&p, 0, sizeof(p));
     std::vector<int> sorted(slots, slots + q);
     std::sort(sorted.begin(), sorted.end());
     uint64_t mask = 0;
     for (int i = 0; i < q; ++i) {
          if (sorted[i] < 0 || sorted[i] >= L || ((mask >> sorted[i]) & 1)) return set_error(HIQ_ERR_ARG, "hiqk_swap_p2p: bad slots");
          mask |= 1ull << sorted[i];
          p.ins.pos[p.ins.n++] = static_cast<uint8_t>(sorted[i]);
     }
     auto spread = [&](uint64_t pat) {
          uint64_t o = 0;
          for (int i = 0; i < q; ++i)
               if ((pat >> i) & 1ull) o |= 1ull << sorted[i];
          return o;
     };
     p.local = static_cast<double2*>(local);
     p.n_peers = n_peers;
     p.theirs_bits = spread(my_pat);
     uint64_t total = 0;
     for (int k = 0; k < n_peers; ++k) {
          if (!peer_slabs[k]) return set_error(HIQ_ERR_ARG, "hiqk_swap_p2p: null peer slab");
          if (peer_pats[k] == my_pat) return set_error(HIQ_ERR_ARG, "hiqk_swap_p2p: a peer with this GPU's own pattern");
          if (begin[k] + count[k] > (1ull << (L - q))) return set_error(HIQ_ERR_ARG, "hiqk_swap_p2p: range outside the slab");
          p.peer[k] = static_cast<double2*>(peer_slabs[k]);
          p.begin[k] = begin[k];
          p.count[k] = count[k];
          p.mine_bits[k] = spread(peer_pats[k]);
          total = std::max(total, count[k]);
     }
     if (total == 0) return HIQ_OK;
     const uint64_t need = (total + static_cast<uint64_t>(kSwapThreads) * 4 - 1) / (static_cast<uint64_t>(kSwapThreads) * 4);
     const unsigned grid = static_cast<unsigned>(std::max<uint64_t>(1, std::min<uint64_t>(need, grid_cap(static_cast<uint64_t>(num_sms()) * 8))));
     swap_p2p_kernel<<<grid, kSwapThreads, 0, static_cast<cudaStream_t>(stream)>>>(p);
     count_launch();
     return check_launch("swap_p2p_kernel");
}
