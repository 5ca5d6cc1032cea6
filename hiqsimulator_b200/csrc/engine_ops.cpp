// Operator-level calls of the reference wrapper that the reference C++ class does not provide:
// get_expectation_value / apply_qubit_operator (reference: _simulator_mpi.py:148-223 call
// self._simulator.get_expectation_value / apply_qubit_operator; _cppsim_mpi.cpp:63-82 exports neither),
// set_wavefunction (:279-305), emulate_math (:459-468; SimulatorMPI.hpp:217-225 throws) and the
// Allgather behind cheat() (:377-380).  Semantics follow the ProjectQ C++ simulator the wrapper was
// written against (projectq/backends/_sim/_cppkernels/simulator.hpp, third party, not vendored):
//   expectation  sum_t c_t Re<psi|P_t|psi>          apply   psi <- sum_t c_t P_t psi (no renormalisation)
//   emulate_math basis states matching the control mask get their register values replaced by f(values)
// Every method starts with run(), like its ProjectQ counterpart.
//
// Distribution (one process per GPU): a Pauli string with X/Y on global qubits couples rank r to the
// single partner r ^ gx; the partner's slab arrives piece by piece through a pairwise NCCL send/recv into a
// bounded staging buffer and is consumed by the same kernels.  A register permutation that involves
// global qubits gathers straight from the peers' slabs mapped into this process (NVLink loads).
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <sstream>

#include "engine.hpp"
#include "hiq_host.hpp"
#include "nccl_api.hpp"

namespace hiq {

namespace {
constexpr size_t kNpos = static_cast<size_t>(-1);
constexpr uint64_t kPieceAmps = 1ull << 24;   // 256 MiB staged per exchange
constexpr uint64_t kGatherAmps = 1ull << 21;  // 32 MiB pieces for cheat()

struct DeviceBuffer {
     void* p = nullptr;
     ~DeviceBuffer()
     {
          if (p) cudaFree(p);
     }
};
}  // namespace

double2* Engine::scratch_buffer(uint64_t amps, const char* who)
{
     if (!scratch_.data()) cu(scratch_.init(device_, 1ull << max_local_, false));
     if (scratch_.ensure(amps) != HIQ_OK)
          fail(std::string(who) + ": needs a second buffer of the slab's size, which does not fit in device memory");
     return scratch_.data();
}

// ------------------------------------------------------------------------------------ Pauli groups
std::vector<Engine::PauliGroup> Engine::pauli_groups(const std::vector<PauliTerm>& terms, const std::vector<Index>& ids,
                                                     const char* what) const
{
     // Composition rule (factors applied left to right, like ProjectQ's apply_term): with
     // P|i> = c (-1)^{popcount(i & z)} |i ^ x>,  X_b P: x ^= b;  Z_b P: z ^= b, c *= (-1)^{x_b};
     // Y_b P: z ^= b, c *= i (-1)^{x_b}, x ^= b.  Masks are split into local slots and rank bits.
     std::vector<PauliGroup> groups;
     std::map<std::pair<int, uint64_t>, size_t> index;
     for (const PauliTerm& t: terms) {
          uint64_t lx = 0, lz = 0;
          int gx = 0, gz = 0;
          cplx c = t.coef;
          for (const auto& f: t.factors) {
               if (f.first < 0 || static_cast<size_t>(f.first) >= ids.size())
                    fail(std::string(what) + ": qubit_operator acts on more qubits than contained in the qureg.");
               const Index q = ids[f.first];
               size_t pos = find(locals_, q);
               const bool local = pos != kNpos;
               if (!local) pos = find_sure(globals_, q);
               const bool flipped = local ? ((lx >> pos) & 1ull) : ((gx >> pos) & 1);
               auto flip_x = [&] {
                    if (local) lx ^= 1ull << pos;
                    else gx ^= 1 << pos;
               };
               auto flip_z = [&] {
                    if (local) lz ^= 1ull << pos;
                    else gz ^= 1 << pos;
               };
               switch (f.second) {
                    case 'X': flip_x(); break;
                    case 'Z':
                         flip_z();
                         if (flipped) c = -c;
                         break;
                    case 'Y':
                         flip_z();
                         c *= flipped ? cplx(0.0, -1.0) : cplx(0.0, 1.0);
                         flip_x();
                         break;
                    default: fail(std::string(what) + ": unknown Pauli operator '" + std::string(1, f.second) + "'");
               }
          }
          // sign contributed by the global qubits: the source amplitudes live on rank ^ gx
          if (__builtin_popcount((rank_ ^ gx) & gz) & 1) c = -c;
          const auto key = std::make_pair(gx, lx);
          auto it = index.find(key);
          if (it == index.end()) {
               it = index.emplace(key, groups.size()).first;
               groups.emplace_back();
               groups.back().lx = lx;
               groups.back().gx = gx;
          }
          groups[it->second].terms.push_back({lz, c.real(), c.imag()});
     }
     return groups;
}

double2* Engine::ensure_staging(uint64_t amps)
{
     const size_t need = amps * sizeof(double2);
     if (!swap_buf_ || swap_buf_bytes_ < need) {
          if (swap_buf_) {
               cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
               cudaFree(swap_buf_);
          }
          swap_buf_ = nullptr;
          swap_buf_bytes_ = 0;
          cu(check_cuda(cudaMalloc(&swap_buf_, need), "cudaMalloc staging"));
          swap_buf_bytes_ = need;
     }
     return static_cast<double2*>(swap_buf_);
}

void Engine::exchange_piece(int partner, uint64_t begin, uint64_t count, double2* staging)
{
     // pairwise and symmetric: the partner issues the mirror image of this group on its own stream
     nccl().GroupStart();
     nccl().Send(slab_.data() + begin, count * 2, ncclDouble, partner, comm_p_->handle(), stream_);
     nccl().Recv(staging, count * 2, ncclDouble, partner, comm_p_->handle(), stream_);
     const ncclResult_t r = nccl().GroupEnd();
     if (r != ncclSuccess) throw EngineError(HIQ_ERR_CUDA, std::string("ncclGroupEnd: ") + nccl().GetErrorString(r));
}

namespace {
Descriptor pauli_descriptor(int kind, int mode, uint64_t lx, int partner, const hiqk_pauli_term* terms, int n)
{
     Descriptor d;
     d.kind = kind;
     d.k = n;
     d.aux = {mode, static_cast<int64_t>(lx), partner};
     for (int t = 0; t < n; ++t) {
          d.aux.push_back(static_cast<int64_t>(terms[t].zmask));
          d.payload.emplace_back(terms[t].re, terms[t].im);
     }
     return d;
}
}  // namespace

double Engine::get_expectation_value(const std::vector<PauliTerm>& terms, const std::vector<Index>& ids)
{
     run();
     const std::vector<PauliGroup> groups = pauli_groups(terms, ids, "get_expectation_value()");
     const int L = static_cast<int>(locals_.size());
     const uint64_t n = 1ull << L;
     if (!dry_run_) flush_pending();
     double total = 0.0;
     for (const PauliGroup& g: groups) {
          const int partner = rank_ ^ g.gx;
          const uint64_t piece = g.gx ? std::min(n, kPieceAmps) : n;
          for (uint64_t begin = 0; begin < n; begin += piece) {
               double2* staging = nullptr;
               if (g.gx && !dry_run_) {
                    staging = ensure_staging(piece);
                    exchange_piece(partner, begin, piece, staging);
               }
               for (size_t t0 = 0; t0 < g.terms.size(); t0 += HIQK_MAX_PAULI_TERMS) {
                    const int nt = static_cast<int>(std::min<size_t>(HIQK_MAX_PAULI_TERMS, g.terms.size() - t0));
                    if (tracing_ && begin == 0) trace_op(pauli_descriptor(HIQ_DESC_PAULI_EXPECT, 0, g.lx, partner, &g.terms[t0], nt));
                    if (dry_run_) continue;
                    double v[2];
                    cu(hiqk_pauli_expect(slab_.data(), L, g.lx, &g.terms[t0], nt, staging, begin, piece, d_vals_, workspace_, stream_));
                    d2h(v, d_vals_, sizeof(v));
                    total += v[0];
               }
          }
     }
     if (dry_run_) return 0.0;  // the value needs the amplitudes; the trace carries the passes
     cu(comm_p_->allreduce_sum(&total, 1, stream_));
     return total;
}

void Engine::apply_qubit_operator(const std::vector<PauliTerm>& terms, const std::vector<Index>& ids)
{
     run();
     const std::vector<PauliGroup> groups = pauli_groups(terms, ids, "apply_qubit_operator()");
     const int L = static_cast<int>(locals_.size());
     const uint64_t n = 1ull << L;
     if (!dry_run_) flush_pending();
     if (groups.empty()) {
          // the empty operator is the zero operator
          if (tracing_) {
               Descriptor d;
               d.kind = HIQ_DESC_SCALE;
               d.payload = {cplx(0.0)};
               trace_op(d);
          }
          if (!dry_run_) cu(check_cuda(cudaMemsetAsync(slab_.data(), 0, n * sizeof(double2), stream_), "cudaMemsetAsync"));
          return;
     }
     if (groups.size() == 1 && groups[0].gx == 0 && groups[0].terms.size() <= HIQK_MAX_PAULI_TERMS) {
          // every term moves amplitude i to the same place: in place, one pass
          const PauliGroup& g = groups[0];
          const int nt = static_cast<int>(g.terms.size());
          if (tracing_) trace_op(pauli_descriptor(HIQ_DESC_PAULI_APPLY, 0, g.lx, rank_, g.terms.data(), nt));
          if (!dry_run_) cu(hiqk_pauli_apply(slab_.data(), L, g.lx, g.terms.data(), nt, nullptr, 0, nullptr, 0, n, stream_));
          return;
     }
     // general operator: new = sum over groups, accumulated in a second buffer (ProjectQ keeps three copies)
     struct {
          void* p = nullptr;
     } acc;
     if (!dry_run_) acc.p = scratch_buffer(n, "apply_qubit_operator(): an operator whose terms flip different qubit sets");
     bool first = true;
     for (const PauliGroup& g: groups) {
          const int partner = rank_ ^ g.gx;
          const uint64_t piece = g.gx ? std::min(n, kPieceAmps) : n;
          for (size_t t0 = 0; t0 < g.terms.size(); t0 += HIQK_MAX_PAULI_TERMS) {
               const int nt = static_cast<int>(std::min<size_t>(HIQK_MAX_PAULI_TERMS, g.terms.size() - t0));
               if (tracing_) trace_op(pauli_descriptor(HIQ_DESC_PAULI_APPLY, first ? 1 : 2, g.lx, partner, &g.terms[t0], nt));
               if (!dry_run_) {
                    for (uint64_t begin = 0; begin < n; begin += piece) {
                         double2* staging = nullptr;
                         if (g.gx) {
                              staging = ensure_staging(piece);
                              exchange_piece(partner, begin, piece, staging);
                         }
                         cu(hiqk_pauli_apply(slab_.data(), L, g.lx, &g.terms[t0], nt, acc.p, first ? 0 : 1, staging, begin, piece, stream_));
                    }
               }
               first = false;
          }
     }
     if (tracing_) {
          Descriptor d;
          d.kind = HIQ_DESC_PAULI_COMMIT;
          trace_op(d);
     }
     if (!dry_run_) {
          // every send of this rank's slab was issued on stream_ before this copy, so the partners have their data
          cu(check_cuda(cudaMemcpyAsync(slab_.data(), acc.p, n * sizeof(double2), cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync"));
          cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
     }
}

// ------------------------------------------------------------------------------------ emulate_time_evolution
void Engine::emulate_time_evolution(const std::vector<PauliTerm>& terms, double time, const std::vector<Index>& ids,
                                    const std::vector<Index>& ctrls)
{
     // psi <- exp(-i t H) psi where the control qubits are 1 — ProjectQ's algorithm (simulator.hpp, emulate_time_evolution;
     // reference call site _simulator_mpi.py:469-475): the identity terms are a phase, the rest is cut into
     // s = |t| * sum|c| + 1 slices, each summed as a Taylor series term by term (term k+1 = -i t / (s (k+1)) * H * term k,
     // applied to the WHOLE vector; only the accumulation looks at the controls) until a term's norm drops to 1e-12.
     run();
     need_device("emulate_time_evolution()");
     cplx tr(0.0);
     double op_nrm = 0.0;
     std::vector<PauliTerm> td;
     for (const PauliTerm& t: terms) {
          if (t.factors.empty()) tr += t.coef;
          else {
               td.push_back(t);
               op_nrm += std::abs(t.coef);
          }
     }
     const unsigned s = static_cast<unsigned>(std::abs(time) * op_nrm + 1.0);
     const cplx I(0.0, 1.0);
     const cplx correction = std::exp(-time * I * tr / static_cast<double>(s));
     uint64_t lm = 0, gm = 0;
     for (Index c: ctrls) {
          size_t pos = find(locals_, c);
          if (pos != kNpos) lm |= 1ull << pos;
          else gm |= 1ull << find_sure(globals_, c);
     }
     const bool takes_part = (static_cast<uint64_t>(rank_) & gm) == gm;
     const int L = static_cast<int>(locals_.size());
     const uint64_t n = 1ull << L;
     flush_pending();
     DeviceBuffer out;
     if (cudaMalloc(&out.p, n * sizeof(double2)) != cudaSuccess) {
          cudaGetLastError();
          fail("emulate_time_evolution(): needs two more buffers of the slab's size, which do not fit in device memory");
     }
     cu(check_cuda(cudaMemcpyAsync(out.p, slab_.data(), n * sizeof(double2), cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync"));
     for (unsigned i = 0; i < s; ++i) {
          for (unsigned k = 0;; ++k) {
               if (k > 10000) fail("emulate_time_evolution(): the Taylor series does not converge");
               const cplx coeff = (-time * I) / static_cast<double>(s * (k + 1));
               std::vector<PauliTerm> scaled = td;
               for (PauliTerm& t: scaled) t.coef *= coeff;
               apply_qubit_operator(scaled, ids);  // slab <- coeff * H' * slab
               double nrm = 0.0;
               if (takes_part) {
                    cu(hiqk_prob_masked(slab_.data(), L, lm, lm, d_vals_, workspace_, stream_));
                    d2h(&nrm, d_vals_, sizeof(double));
                    cu(hiqk_axpy_masked(out.p, slab_.data(), L, lm, lm, 1.0, 0.0, stream_));
               }
               cu(comm_p_->allreduce_sum(&nrm, 1, stream_));
               if (!(std::sqrt(nrm) > 1.e-12)) break;
          }
          if (takes_part) {
               const cplx a = correction - 1.0;
               cu(hiqk_axpy_masked(out.p, out.p, L, lm, lm, a.real(), a.imag(), stream_));
          }
          cu(check_cuda(cudaMemcpyAsync(slab_.data(), out.p, n * sizeof(double2), cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync"));
     }
     cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
}

// ------------------------------------------------------------------------------------ set_wavefunction
void Engine::set_wavefunction(const cplx* amps, uint64_t n_amps, const std::vector<Index>& ordering)
{
     // ProjectQ: the simulator adopts `ordering` (index bit i <-> ordering[i]) and copies the amplitudes.
     // Here the first L ids become the local slots and the others fill the occupied global positions in
     // ascending order; rank r copies the slice its global bits select (zeros when a bit is set at an empty
     // global position).
     run();
     std::vector<Index> allocated = locals_;
     std::vector<int> occupied;
     for (size_t p = 0; p < globals_.size(); ++p)
          if (globals_[p] != kNone) {
               allocated.push_back(globals_[p]);
               occupied.push_back(static_cast<int>(p));
          }
     std::vector<Index> a = allocated, b = ordering;
     std::sort(a.begin(), a.end());
     std::sort(b.begin(), b.end());
     if (a != b || ordering.size() >= 63 || n_amps != (1ull << ordering.size()) || !amps)
          fail("set_wavefunction(): Invalid mapping provided. Please make sure all qubits have been allocated previously "
               "(call eng.flush()).");
     const size_t L = locals_.size();
     locals_.assign(ordering.begin(), ordering.begin() + L);
     for (size_t k = 0; k < occupied.size(); ++k) globals_[occupied[k]] = ordering[L + k];
     int64_t slice = 0;
     for (size_t p = 0; p < globals_.size(); ++p) {
          if (!((rank_ >> p) & 1)) continue;
          const auto it = std::find(occupied.begin(), occupied.end(), static_cast<int>(p));
          if (it == occupied.end()) {
               slice = -1;
               break;
          }
          slice |= 1ll << (it - occupied.begin());
     }
     if (tracing_) {
          Descriptor d;
          d.kind = HIQ_DESC_LOAD;
          d.aux = {slice};
          trace_op(d);
     }
     if (dry_run_) return;
     flush_pending();
     const uint64_t n = 1ull << L;
     if (slice < 0) {
          cu(check_cuda(cudaMemsetAsync(slab_.data(), 0, n * sizeof(double2), stream_), "cudaMemsetAsync"));
     }
     else {
          cu(check_cuda(cudaMemcpyAsync(slab_.data(), amps + (static_cast<uint64_t>(slice) << L), n * sizeof(double2),
                                        cudaMemcpyHostToDevice, stream_),
                        "cudaMemcpyAsync"));
          stats_.h2d_bytes += static_cast<double>(n * sizeof(double2));
     }
     cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
}

// ------------------------------------------------------------------------------------ emulate_math
void Engine::emulate_math(int kind, uint64_t a, uint64_t N, const std::vector<uint64_t>& fwd_table, const std::vector<Index>& reg_ids,
                          const std::vector<Index>& ctrls)
{
     run();
     const int L = static_cast<int>(locals_.size());
     const uint64_t n = 1ull << L;
     hiqk_perm perm;
     std::memset(&perm, 0, sizeof(perm));
     perm.kind = kind;
     perm.a = a;
     perm.N = N;
     if (reg_ids.empty() || reg_ids.size() > 40) fail("emulate_math(): registers must hold 1..40 qubits in total");
     perm.n_bits = static_cast<int>(reg_ids.size());
     auto index_bit = [&](Index q) {
          size_t pos = find(locals_, q);
          if (pos != kNpos) return static_cast<int>(pos);
          return L + static_cast<int>(find_sure(globals_, q));
     };
     uint64_t reg_mask = 0;
     for (size_t b = 0; b < reg_ids.size(); ++b) {
          perm.pos[b] = index_bit(reg_ids[b]);
          if ((reg_mask >> perm.pos[b]) & 1ull) fail("emulate_math(): a qubit appears twice in the registers");
          reg_mask |= 1ull << perm.pos[b];
     }
     for (Index c: ctrls) perm.ctrl_mask |= 1ull << index_bit(c);
     if (perm.ctrl_mask & reg_mask) fail("emulate_math(): a control qubit is part of a register");
     const uint64_t space = 1ull << perm.n_bits;
     // validate the map on the host (the same checks the launcher repeats) and build the inverse table
     std::vector<uint32_t> inverse;
     switch (kind) {
          case HIQK_PERM_TABLE: {
               if (perm.n_bits > 30) fail("emulate_math(): a tabulated function is limited to 30 register qubits");
               if (fwd_table.size() != space) fail("emulate_math(): the function table must have 2^(register qubits) entries");
               inverse.assign(space, 0xffffffffu);
               for (uint64_t v = 0; v < space; ++v) {
                    const uint64_t w = fwd_table[v] & (space - 1ull);  // results are truncated to the register width
                    if (inverse[w] != 0xffffffffu)
                         fail("emulate_math(): the function is not reversible on the registers (two values map to " + std::to_string(w) + ")");
                    inverse[w] = static_cast<uint32_t>(v);
               }
               break;
          }
          case HIQK_PERM_ADD: break;
          case HIQK_PERM_ADD_MOD:
               if (N < 1 || N > space) fail("emulate_math(): the modulus must be in [1, 2^(register qubits)]");
               break;
          case HIQK_PERM_MUL_MOD: {
               if (N < 2 || N > space || N > (1ull << 32)) fail("emulate_math(): the modulus must be in [2, min(2^(register qubits), 2^32)]");
               uint64_t inv = 0;
               if (hiq_modinv(a, N, &inv) != HIQ_OK) fail("emulate_math(): the multiplier is not invertible modulo N (the map is not reversible)");
               break;
          }
          default: fail("emulate_math(): unknown function kind");
     }
     // ranks that can hold a source amplitude of this rank: they differ in global register bits only
     const int reach = static_cast<int>(reg_mask >> L);
     std::vector<int> peer_ranks;
     for (int r = 0; r < world_; ++r)
          if (r != rank_ && ((r ^ rank_) & ~reach) == 0) peer_ranks.push_back(r);
     const uint64_t global_ctrl = perm.ctrl_mask >> L;
     const bool participate = (static_cast<uint64_t>(rank_) & global_ctrl) == global_ctrl;
     if (tracing_) {
          Descriptor d;
          d.kind = HIQ_DESC_PERMUTE;
          d.k = perm.n_bits;
          d.aux = {kind, static_cast<int64_t>(a), static_cast<int64_t>(N), static_cast<int64_t>(perm.ctrl_mask), participate ? 1 : 0};
          for (int b = 0; b < perm.n_bits; ++b) d.aux.push_back(perm.pos[b]);
          if (inverse.size() <= (1u << 20))  // larger tables are not worth carrying in a trace
               for (uint32_t v: inverse) d.aux.push_back(v);
          trace_op(d);
     }
     if (dry_run_) return;
     flush_pending();
     if (world_ > 16) fail("emulate_math(): at most 16 ranks");
     if (!peer_ranks.empty() && !ensure_peer_views(peer_ranks))
          fail(std::string("emulate_math(): registers on global qubits need the peers' slabs mapped into this process: ") + hiq_last_error());
     if (!participate) return;  // a global control is 0 here — and on every rank this one could exchange with
     DeviceBuffer table;
     double2* tmp = scratch_buffer(n, "emulate_math()");
     if (!inverse.empty()) {
          cu(check_cuda(cudaMalloc(&table.p, inverse.size() * sizeof(uint32_t)), "cudaMalloc table"));
          cu(check_cuda(cudaMemcpyAsync(table.p, inverse.data(), inverse.size() * sizeof(uint32_t), cudaMemcpyHostToDevice, stream_),
                        "cudaMemcpyAsync"));
          stats_.h2d_bytes += static_cast<double>(inverse.size() * sizeof(uint32_t));
          perm.table = static_cast<const uint32_t*>(table.p);
     }
     std::vector<const void*> slabs(world_, nullptr);
     slabs[rank_] = slab_.data();
     for (int pr: peer_ranks) slabs[pr] = peer_views_[pr].slab.data();
     if (!peer_ranks.empty()) group_barrier(peer_ranks);  // the peers' gates are complete before their slabs are read
     cu(hiqk_permute_gather(tmp, slabs.data(), world_, rank_, L, &perm, stream_));
     if (!peer_ranks.empty()) group_barrier(peer_ranks);  // nobody overwrites a slab a peer still reads
     if (world_ == 1) {
          // one process: the permuted copy becomes the slab (the peers of a multi-GPU world have the slab's memory mapped,
          // there the result is copied back)
          slab_.swap(scratch_);
     }
     else {
          cu(check_cuda(cudaMemcpyAsync(slab_.data(), tmp, n * sizeof(double2), cudaMemcpyDeviceToDevice, stream_), "cudaMemcpyAsync"));
     }
     if (table.p) cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));  // the table is freed on return
}

// ------------------------------------------------------------------------------------ cheat()
void Engine::gather_state_to_host(void* dst, uint64_t cap_amps)
{
     // reference: MPI.COMM_WORLD.Allgather of the rank slabs, _simulator_mpi.py:377-380
     need_device("cheat()");
     const uint64_t n = 1ull << locals_.size();
     if (!dst || cap_amps < n * static_cast<uint64_t>(world_)) fail("cheat(): destination buffer too small");
     if (world_ == 1) {
          copy_slab_to_host(dst, cap_amps);
          return;
     }
     flush_pending();
     const uint64_t piece = std::min(n, kGatherAmps);
     double2* staging = ensure_staging(piece);
     cplx* out = static_cast<cplx*>(dst);
     for (int root = 0; root < world_; ++root) {
          for (uint64_t begin = 0; begin < n; begin += piece) {
               const ncclResult_t r = nccl().Broadcast(slab_.data() + begin, staging, piece * 2, ncclDouble, root, comm_p_->handle(), stream_);
               if (r != ncclSuccess) throw EngineError(HIQ_ERR_CUDA, std::string("ncclBroadcast: ") + nccl().GetErrorString(r));
               cu(check_cuda(cudaMemcpyAsync(out + static_cast<uint64_t>(root) * n + begin, staging, piece * sizeof(double2),
                                             cudaMemcpyDeviceToHost, stream_),
                             "cudaMemcpyAsync"));
               cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
               stats_.d2h_bytes += static_cast<double>(piece * sizeof(double2));
          }
     }
}

}  // namespace hiq
