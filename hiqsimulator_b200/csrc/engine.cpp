#include "engine.hpp"

#include <algorithm>
#include <cmath>
#include <unistd.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <sstream>

#include "hiq_host.hpp"
#include "nccl_api.hpp"

namespace hiq {

namespace {
constexpr size_t kNpos = static_cast<size_t>(-1);
constexpr uint64_t kMaxBlocks = 1ull << 15;        // measurement blocks per rank (reference: SimulatorMPI.cpp:903)
constexpr uint64_t kCompactScratchAmps = 1ull << 21;  // 32 MiB: stays in L2 between gather and copy
constexpr uint64_t kSwapPieceAmps = 1ull << 24;       // 256 MiB per peer per piece

std::string list_str(const std::vector<Index>& v)
{
     std::ostringstream o;
     o << "[";
     for (size_t i = 0; i < v.size(); ++i) o << (i ? ", " : "") << v[i];
     o << "]";
     return o.str();
}
double seconds_since(std::chrono::steady_clock::time_point t0)
{
     return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
}  // namespace

void Engine::fail(const std::string& msg) const { throw EngineError(HIQ_ERR_RUNTIME, msg); }

void Engine::cu(int rc) const
{
     if (rc != HIQ_OK) throw EngineError(rc, hiq_last_error());
}

void Engine::need_device(const char* what) const
{
     if (dry_run_) fail(std::string(what) + ": not available on a dry-run (descriptor-trace) engine");
}

Engine::Engine(uint64_t seed, int max_local, int max_cluster, int rank, int world_size, const void* nccl_id, int device,
               int flags)
    : min_local_(max_cluster),
      max_local_(max_local),
      max_cluster_(max_cluster),
      rank_(rank),
      world_(world_size),
      device_(device),
      dry_run_(flags & HIQ_FLAG_DRY_RUN),
      tracing_(flags & (HIQ_FLAG_TRACE | HIQ_FLAG_DRY_RUN)),
      timing_((flags & HIQ_FLAG_TIMING) && !(flags & HIQ_FLAG_DRY_RUN)),
      batching_(!(flags & HIQ_FLAG_NO_BATCH))
{
     const auto t_ctor = Clock::now();
     if (world_size < 1 || (world_size & (world_size - 1)) || rank < 0 || rank >= world_size)
          fail("ctor(): world size must be a power of two and 0 <= rank < world size");
     if (max_local < 1 || max_local > 40 || max_cluster < 1) fail("ctor(): bad max_local / max_cluster_size");
     max_global_ = 0;
     while ((1 << max_global_) < world_size) ++max_global_;
     globals_.assign(max_global_, kNone);
     // every rank is handed the same seed (the reference broadcasts rank 0's, SimulatorMPI.cpp:87)
     rnd_eng_ = std::mt19937(seed);
     std::uniform_real_distribution<double> dist(0., 1.);
     rng_ = std::bind(dist, std::ref(rnd_eng_));
     if (!dry_run_) {
          if (const char* m = std::getenv("HIQ_SWAP_MODE")) {
               if (!std::strcmp(m, "staged") || !std::strcmp(m, "nccl")) swap_mode_ = 1;
               else if (!std::strcmp(m, "p2p")) swap_mode_ = 2;
               else if (!std::strcmp(m, "packed")) swap_mode_ = 3;
          }
          if (const char* m = std::getenv("HIQ_SWAP_PACKED")) packed_enabled_ = m[0] == '1';
          if (const char* m = std::getenv("HIQ_SWAP_PACKED_BELOW")) packed_below_slot_ = std::atoi(m);
          if (const char* m = std::getenv("HIQ_SWAP_PACKED_PIECE")) packed_piece_cap_ = std::strtoull(m, nullptr, 10);
          if (const char* m = std::getenv("HIQ_SWAP_P2P_MIN_SLOT")) min_p2p_slot_ = std::atoi(m);
     }
     if (const char* m = std::getenv("HIQ_TILE")) tile_enabled_ = m[0] != '0';
     if (const char* m = std::getenv("HIQ_TILE_MAX_FULL")) tile_max_full_ = std::atoi(m);
     if (const char* m = std::getenv("HIQ_TILE_MAX_STEPS")) tile_max_steps_ = std::max(1, std::min(HIQK_TILE_MAX_STEPS, std::atoi(m)));
     if (const char* m = std::getenv("HIQ_TILE_SINGLE")) tile_single_ = m[0] == '1';
     if (!dry_run_) try {
          cu(check_cuda(cudaSetDevice(device_), "cudaSetDevice"));
          cu(slab_.init(device_, 1ull << max_local_, world_size > 1));
          cu(check_cuda(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "cudaStreamCreate"));
          cu(check_cuda(cudaStreamCreateWithFlags(&comm_stream_, cudaStreamNonBlocking), "cudaStreamCreate"));
          comm_p_ = Comm::shared(rank, world_size, nccl_id, device_);
          if (!comm_p_) throw EngineError(HIQ_ERR_CUDA, hiq_last_error());
          epoch_ = comm_p_->next_epoch();
          cu(check_cuda(cudaMalloc(&workspace_, hiqk_workspace_bytes()), "cudaMalloc workspace"));
          cu(check_cuda(cudaMalloc(&d_vals_, 64 * sizeof(double)), "cudaMalloc"));
          cu(slab_.ensure(1));
          cu(hiqk_fill(slab_.data(), 0, 1, rank_ == 0 ? 1.0 : 0.0, 0.0, stream_));
     }
     catch (...) {
          // a constructor that throws runs no destructor: give back what was created so far (hiq_create may be retried)
          release_device_resources();
          throw;
     }
     ++stats_.total_stages;
     stats_.ctor_s = seconds_since(t_ctor);
}

void Engine::release_device_resources()
{
     if (dry_run_) return;
     cudaSetDevice(device_);
     cudaDeviceSynchronize();
     if (workspace_) cudaFree(workspace_);
     if (d_vals_) cudaFree(d_vals_);
     if (d_blocks_) cudaFree(d_blocks_);
     if (swap_buf_) cudaFree(swap_buf_);
     workspace_ = nullptr;
     d_vals_ = d_blocks_ = nullptr;
     swap_buf_ = nullptr;
     slab_.release();
     scratch_.release();
     for (auto& t: timed_) {
          cudaEventDestroy(t.start);
          cudaEventDestroy(t.stop);
     }
     timed_.clear();
     for (auto ev: event_pool_) cudaEventDestroy(ev);
     event_pool_.clear();
     for (auto& ev: swap_events_)
          if (ev) {
               cudaEventDestroy(ev);
               ev = nullptr;
          }
     if (stream_) cudaStreamDestroy(stream_);
     if (comm_stream_) cudaStreamDestroy(comm_stream_);
     stream_ = comm_stream_ = nullptr;
}

Engine::~Engine() { release_device_resources(); }

void Engine::synchronize()
{
     flush_pending();  // on a dry-run engine: closes the launch accounting (stats), nothing is launched
     if (dry_run_) return;
     cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
}

size_t Engine::find(const std::vector<Index>& v, Index val) const
{
     auto it = std::find(v.begin(), v.end(), val);
     return it == v.end() ? kNpos : static_cast<size_t>(it - v.begin());
}

size_t Engine::find_sure(const std::vector<Index>& v, Index val) const
{
     const size_t pos = find(v, val);
     if (pos == kNpos) fail("ArrayFindSure(): Can't find " + std::to_string(val) + " in " + list_str(v));
     return pos;
}

uint64_t Engine::ids_to_bits(const std::vector<Index>& ids, const std::vector<Index>& perm) const
{
     uint64_t mask = 0;
     for (Index q: ids) {
          const size_t pos = find(perm, q);
          if (pos != kNpos) mask |= 1ull << pos;
     }
     return mask;
}

// ------------------------------------------------------------------------------------ allocation
void Engine::allocate_local(Index id)
{
     const uint64_t old = 1ull << locals_.size();
     // queued launches (held dense gate, batched diagonals) were planned for the slab as it is NOW: they go out
     // before the register grows — afterwards locals_.size() would describe a slab whose upper half is not mapped yet
     flush_pending();
     locals_.push_back(id);
     if (tracing_) {
          Descriptor d;
          d.kind = HIQ_DESC_GROW;
          d.k = static_cast<int>(locals_.size());
          trace_op(d);
     }
     if (dry_run_) return;
     grow_slab(2 * old);
     cu(check_cuda(cudaMemsetAsync(slab_.data() + old, 0, old * sizeof(double2), stream_), "cudaMemsetAsync"));
}

void Engine::grow_slab(uint64_t amps)
{
     const auto t_grow = Clock::now();
     if (slab_.ensure(amps) != HIQ_OK && scratch_.data()) {
          scratch_.release();  // the second buffer of the out-of-place passes gives its memory back to the register
          cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));
     }
     cu(slab_.ensure(amps));
     stats_.slab_grow_s += seconds_since(t_grow);
}

void Engine::allocate_global(Index id) { globals_[find_sure(globals_, kNone)] = id; }

void Engine::allocate_qubit(Index id)
{
     // policy of the reference (SimulatorMPI.cpp:184-216)
     const auto t0 = Clock::now();
     // the reference does not look (a second allocation of a live id leaves it in two places of the maps, -1 is its mark
     // of an empty global position); refused here, on every rank alike
     if (id < 0) fail("AllocateQubit(): qubit ids must be non-negative");
     if (find(locals_, id) != kNpos || find(globals_, id) != kNpos)
          fail("AllocateQubit(): qubit " + std::to_string(id) + " is already allocated");
     const size_t nloc = locals_.size();
     if (nloc < min_local_) allocate_local(id);
     else if (find(globals_, kNone) != kNpos) allocate_global(id);
     else if (nloc < max_local_) allocate_local(id);
     else fail("AllocateQubit(): can't allocate more than " + std::to_string(globals_.size() + nloc) + " qubits");
     stats_.allocs_s += seconds_since(t0);
}

void Engine::allocate_qureg(const std::vector<Index>& ids, cplx init)
{
     if (init != cplx(0.0) && !locals_.empty())
          fail("AllocateQureg(): initialization of only first qureg is supported");
     if (!dry_run_) {
          // map the register's final extent in ONE piece instead of one piece per doubling: a slab made of few
          // virtual-memory chunks is cheaper to hand to the peers (one descriptor + one mapping per chunk and peer)
          size_t nloc = locals_.size();
          size_t free_globals = static_cast<size_t>(std::count(globals_.begin(), globals_.end(), kNone));
          for (size_t i = 0; i < ids.size(); ++i) {  // the allocation policy of allocate_qubit()
               if (nloc < min_local_) ++nloc;
               else if (free_globals > 0) --free_globals;
               else if (nloc < max_local_) ++nloc;
               else break;  // allocate_qubit() raises when it gets there
          }
          if (nloc > locals_.size()) {
               flush_pending();
               grow_slab(1ull << nloc);
          }
     }
     for (Index q: ids) allocate_qubit(q);
     if (init != cplx(0.0)) {
          uint64_t gmsk = 0;
          for (size_t pos = 0; pos < globals_.size(); ++pos)
               if (globals_[pos] != kNone) gmsk |= 1ull << pos;
          if ((static_cast<uint64_t>(rank_) & ~gmsk) == 0) {
               if (tracing_) {
                    Descriptor d;
                    d.kind = HIQ_DESC_FILL;
                    d.payload = {init};
                    trace_op(d);
               }
               if (!dry_run_) cu(hiqk_fill(slab_.data(), 0, 1ull << locals_.size(), init.real(), init.imag(), stream_));
          }
     }
}

void Engine::ensure_scratch()
{
     if (!swap_buf_ || swap_buf_bytes_ < kCompactScratchAmps * sizeof(double2)) {
          if (swap_buf_) cudaFree(swap_buf_);
          swap_buf_ = nullptr;
          swap_buf_bytes_ = kCompactScratchAmps * sizeof(double2);
          cu(check_cuda(cudaMalloc(&swap_buf_, swap_buf_bytes_), "cudaMalloc scratch"));
     }
}

void Engine::deallocate_local(Index id)
{
     need_device("DeallocateLocalQubit()");
     flush_pending();
     const size_t pos = find_sure(locals_, id);
     const int L = static_cast<int>(locals_.size());
     double sums[2];
     cu(hiqk_bit_norms(slab_.data(), L, static_cast<int>(pos), d_vals_, workspace_, stream_));
     d2h(sums, d_vals_, sizeof(sums));
     cu(comm_p_->allreduce_sum(sums, 2, stream_));
     if (!((sums[0] > max_float_error_) ^ (sums[1] > max_float_error_)))
          fail("DeallocateLocalQubit(): qubit " + std::to_string(id) + " is entangled");
     const int keep = sums[0] > max_float_error_ ? 0 : 1;
     ensure_scratch();
     cu(hiqk_compact_bit(slab_.data(), L, static_cast<int>(pos), keep, swap_buf_, kCompactScratchAmps, stream_));
     locals_.erase(locals_.begin() + pos);
}

void Engine::deallocate_global(Index id)
{
     need_device("DeallocateGlobalQubit()");
     flush_pending();
     const size_t pos = find_sure(globals_, id);
     const int L = static_cast<int>(locals_.size());
     double local_norm = 0.0;
     cu(hiqk_prob_masked(slab_.data(), L, 0, 0, d_vals_, workspace_, stream_));
     d2h(&local_norm, d_vals_, sizeof(double));
     double sums[2] = {0.0, 0.0};
     sums[(rank_ >> pos) & 1] = local_norm;
     cu(comm_p_->allreduce_sum(sums, 2, stream_));
     if (!((sums[0] > max_float_error_) ^ (sums[1] > max_float_error_)))
          fail("DeallocateGlobalQubit(): qubit " + std::to_string(id) + " is entangled");
     if (sums[1] > max_float_error_) {
          // |1>: bring it to the last local slot, flip, send it back (reference: SimulatorMPI.cpp:350-361)
          const Index local_qubit = locals_.back();
          swap_qubits({id, local_qubit});
          GateMatrix x(2);
          x.at(0, 1) = 1.0;
          x.at(1, 0) = 1.0;
          apply_gate(x, {id}, {});
          run();
          swap_qubits({local_qubit, id});
     }
     globals_[pos] = kNone;
}

void Engine::deallocate_qubit(Index id)
{
     // routing of the reference (SimulatorMPI.cpp:369-426)
     const auto t0 = Clock::now();
     const size_t nloc = locals_.size();
     const size_t nglob = globals_.size() - std::count(globals_.begin(), globals_.end(), kNone);
     if (find(locals_, id) != kNpos) {
          if (nloc > min_local_ || nglob == 0) {
               deallocate_local(id);
          }
          else {
               const Index g = *std::find_if(globals_.begin(), globals_.end(), [](Index q) { return q != kNone; });
               swap_qubits({g, id});
               deallocate_global(id);
          }
     }
     else {
          if (nloc > min_local_ || nglob == 0) {
               if (locals_.empty()) fail("ArrayFindSure(): Can't find " + std::to_string(id) + " in " + list_str(globals_));
               swap_qubits({id, locals_.back()});
               deallocate_local(id);
          }
          else {
               deallocate_global(id);
          }
     }
     stats_.deallocs_s += seconds_since(t0);
}

// ------------------------------------------------------------------------------------ gates
void Engine::apply_gate(GateMatrix m, std::vector<Index> ids, std::vector<Index> ctrls)
{
     // per-rank preprocessing of the reference (SimulatorMPI.cpp:710-815)
     if (ids.size() > 30 || m.dim != (1 << ids.size())) fail("ApplyGate(): matrix size does not match the number of target qubits");
     // Checked here, on every rank alike (the id lists are the same everywhere), before the rank-dependent control filter:
     // the reference meets an unknown qubit in Run() on the ranks that kept the gate only, which then leave the collective
     // sequence the other ranks stay in.
     for (const std::vector<Index>* list: {&ids, &ctrls})
          for (Index q: *list)
               if (q < 0 || (find(locals_, q) == kNpos && find(globals_, q) == kNpos))
                    fail("ArrayFindSure(): Can't find " + std::to_string(q) + " in " + list_str(locals_));
     {
          // a control named twice (nested Control blocks on one qubit) is one control
          std::vector<Index> once;
          for (Index c: ctrls)
               if (std::find(once.begin(), once.end(), c) == once.end()) once.push_back(c);
          ctrls.swap(once);
          std::vector<Index> all(ids);
          all.insert(all.end(), ctrls.begin(), ctrls.end());
          std::sort(all.begin(), all.end());
          if (std::adjacent_find(all.begin(), all.end()) != all.end()) fail("ApplyGate(): target and control qubits must be distinct");
     }
     const bool diag = is_diagonal(m);
     const uint64_t global_id_mask = ids_to_bits(ids, globals_);
     const uint64_t global_ctrl_mask = ids_to_bits(ctrls, globals_);
     std::vector<Index> local_ctrls;
     for (Index c: ctrls)
          if (find(locals_, c) != kNpos) local_ctrls.push_back(c);
     const bool huge = ids.size() + local_ctrls.size() > max_cluster_;
     if (huge) run();  // flush whatever is pending before a gate wider than the cluster

     // Deviation: the reference tests this only on ranks that pass the control filter below and
     // then blocks in a barrier the other ranks never reach; here every rank raises together.
     if (global_id_mask != 0 && !diag) fail("ApplyGate(): can't apply non-diagonal gate to global qubits");
     // a gate wider than the cluster goes to the fusion as it is (below, as in the reference, which then cannot find the
     // global target in Run() on the ranks that kept the gate): refused here, on every rank
     if (huge && global_id_mask != 0) fail("ApplyGate(): a gate wider than the cluster size cannot act on global qubits");
     if ((static_cast<uint64_t>(rank_) & global_ctrl_mask) != global_ctrl_mask) return;  // not this rank

     ++stats_.total_gates;
     if (huge) {
          fused_.insert(std::move(m), diag, std::move(ids), ctrls);
          return;
     }
     // keep the rows/columns whose global-target bits equal this rank's bits
     auto ok_bit = [&](int msk) {
          for (size_t i = 0; i < ids.size(); ++i) {
               const size_t pos = find(globals_, ids[i]);
               if (pos != kNpos && ((msk >> i) & 1) != ((rank_ >> pos) & 1)) return false;
          }
          return true;
     };
     std::vector<int> keep;
     for (int i = 0; i < m.dim; ++i)
          if (ok_bit(i)) keep.push_back(i);
     GateMatrix sub(static_cast<int>(keep.size()));
     for (int i = 0; i < sub.dim; ++i)
          for (int j = 0; j < sub.dim; ++j) sub.at(i, j) = m.at(keep[i], keep[j]);
     FusionAccumulator::add_controls(sub, ids, local_ctrls);
     ids.erase(std::remove_if(ids.begin(), ids.end(), [&](Index q) { return find(globals_, q) != kNpos; }), ids.end());
     fused_.insert(std::move(sub), diag, std::move(ids), {});
}

cudaEvent_t Engine::take_event()
{
     if (!event_pool_.empty()) {
          cudaEvent_t ev = event_pool_.back();
          event_pool_.pop_back();
          return ev;
     }
     cudaEvent_t ev;
     cu(check_cuda(cudaEventCreate(&ev), "cudaEventCreate"));
     return ev;
}

std::vector<Engine::PassTime> Engine::collect_timings()
{
     std::vector<PassTime> out;
     if (dry_run_) return out;
     synchronize();
     for (auto& t: timed_) {
          float ms = 0.f;
          cudaEventElapsedTime(&ms, t.start, t.stop);
          out.push_back({t.kind, t.k, t.variant, t.n_ref, static_cast<double>(ms)});
          event_pool_.push_back(t.start);
          event_pool_.push_back(t.stop);
     }
     timed_.clear();
     return out;
}

void Engine::d2h(void* dst, const void* src, size_t bytes)
{
     flush_pending();
     cu(check_cuda(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, stream_), "cudaMemcpyAsync"));
     synchronize();
     stats_.d2h_bytes += static_cast<double>(bytes);
}

void Engine::trace_op(const Descriptor& d)
{
     if (!tracing_) return;
     // a dry-run engine closes its launch accounting where the device path flushes its queues: before anything that is
     // not a gate pass (the launch trace then lists the launches in device order)
     if (dry_run_) flush_pending();
     trace_.push_back(d);
     if (dry_run_) launches_.push_back(d);
}

void Engine::record_launch(int form, const std::vector<LaunchStepRef>& steps)
{
     Descriptor r;
     r.kind = HIQ_DESC_LAUNCH;
     r.k = static_cast<int>(steps.size());
     r.slots[0] = form;
     for (const LaunchStepRef& st: steps) {
          const int k = st.gate ? st.gate->k : -1;
          r.aux.push_back(k);
          for (int l = 0; l < k; ++l) r.aux.push_back(st.gate->slots[l]);
          r.aux.push_back(st.ops ? static_cast<int64_t>(st.ops->size()) : 0);
          if (st.gate) r.payload.insert(r.payload.end(), st.gate->payload.begin(), st.gate->payload.end());
          if (!st.ops) continue;
          for (const hiqk_diag_op& o: *st.ops) {
               r.aux.push_back(o.k);
               for (int l = 0; l < o.k; ++l) r.aux.push_back(o.slots[l]);
               const cplx* t = reinterpret_cast<const cplx*>(o.lut);
               r.payload.insert(r.payload.end(), t, t + (1 << o.k));
          }
     }
     launches_.push_back(std::move(r));
}

namespace {
uint64_t slot_mask(const hiqk_diag_op& o)
{
     uint64_t m = 0;
     for (int l = 0; l < o.k; ++l) m |= 1ull << o.slots[l];
     return m;
}

// Folds `op` into a queued op over a superset of its slots (one table instead of two lookups), or
// appends it when the queue has room.  Returns false when neither is possible.
bool merge_or_append(std::vector<hiqk_diag_op>& q, std::vector<int>& refs, const hiqk_diag_op& op, size_t cap)
{
     const uint64_t mo = slot_mask(op);
     for (size_t i = 0; i < q.size(); ++i) {
          hiqk_diag_op& t = q[i];
          if ((mo & ~slot_mask(t)) != 0) continue;
          cplx* tl = reinterpret_cast<cplx*>(t.lut);
          const cplx* ol = reinterpret_cast<const cplx*>(op.lut);
          for (int e = 0; e < (1 << t.k); ++e) {
               int sel = 0;
               for (int l = 0; l < op.k; ++l)
                    for (int m = 0; m < t.k; ++m)
                         if (t.slots[m] == op.slots[l] && ((e >> m) & 1)) sel |= 1 << l;
               tl[e] *= ol[sel];
          }
          ++refs[i];
          return true;
     }
     if (q.size() >= cap) return false;
     q.push_back(op);
     refs.push_back(1);
     return true;
}
}  // namespace

void Engine::execute(const Descriptor& d)
{
     if (tracing_) trace_.push_back(d);
     const int L = static_cast<int>(locals_.size());
     if (!dry_run_) stats_.h2d_bytes += static_cast<double>(d.payload.size() * sizeof(cplx));
     int variant = 0;
     if (d.kind == HIQ_DESC_DENSE) variant = dense_variant_ ? dense_variant_ : hiqk_dense_pick_variant(L, d.k, d.slots);
     if (batching_) {
          // diagonal passes of the plan are deferred: they ride along a neighbouring dense launch or go
          // out as one batched pass at the next observation of the slab
          if ((d.kind == HIQ_DESC_DIAG || d.kind == HIQ_DESC_SCALE) && d.ctrl_mask == 0) {
               queue_diagonal(d);
               return;
          }
          if (d.kind == HIQ_DESC_DENSE && d.ctrl_mask == 0 && d.k <= 4 && dense_variant_ == 0) {
               const bool direct = variant == HIQK_DENSE_DIRECT && hiqk_dense_prediag_supported(L, d.k, d.slots);
               HeldGate cand;
               cand.d = d;
               cand.variant = variant;
               cand.direct = direct;
               cand.full = hiqk_dense_direct_mixing_bits(d.k, reinterpret_cast<const double*>(d.payload.data())) >= 4 &&
                           !hiqk_dense_is_monomial(d.k, reinterpret_cast<const double*>(d.payload.data()));
               // the diagonals queued so far precede this gate; more than one launch carries go out as batched passes first
               const size_t cap = tile_enabled_ ? HIQK_TILE_MAX_OPS : HIQK_MAX_DIAG_OPS;
               if (pending_.size() > cap) flush_pending(cap);
               if (tile_enabled_ && !group_.empty() && group_accepts(cand, pending_)) {
                    cand.ops.swap(pending_);
                    cand.refs.swap(pending_ref_);
                    group_.push_back(std::move(cand));
                    return;
               }
               launch_group();  // the gates held so far go out now
               if (pending_.size() > cap) flush_pending(cap);
               if (direct || (tile_enabled_ && group_accepts(cand, pending_))) {
                    cand.ops.swap(pending_);  // applied to the tuples this gate loads
                    cand.refs.swap(pending_ref_);
                    group_.push_back(std::move(cand));
                    return;
               }
          }
     }
     flush_pending();
     launch(d, variant, {}, {});
}

bool Engine::group_accepts(const HeldGate& cand, const std::vector<hiqk_diag_op>& cand_ops) const
{
     // may `cand` (with the diagonals that precede it) join the gates held so far in one tile-resident pass?
     const int L = static_cast<int>(locals_.size());
     if (group_.size() + 1 > static_cast<size_t>(tile_max_steps_)) return false;
     int n_full = cand.full ? 1 : 0;
     for (const HeldGate& g: group_) n_full += g.full ? 1 : 0;
     if (n_full > tile_max_full_ && group_.size() + 1 > 1) return false;
     std::vector<hiqk_tile_step> steps(group_.size() + 1);
     auto fill = [](hiqk_tile_step& st, const HeldGate& g, const std::vector<hiqk_diag_op>& ops) {
          st.k = g.d.k;
          for (int l = 0; l < 5; ++l) st.slots[l] = g.d.slots[l];
          st.matrix = reinterpret_cast<const double*>(g.d.payload.data());
          st.pre = ops.data();
          st.n_pre = static_cast<int>(ops.size());
     };
     for (size_t i = 0; i < group_.size(); ++i) fill(steps[i], group_[i], group_[i].ops);
     fill(steps.back(), cand, cand_ops);
     return hiqk_tile_program_fits(L, static_cast<int>(steps.size()), steps.data()) != 0;
}

void Engine::queue_diagonal(const Descriptor& d)
{
     hiqk_diag_op op;
     std::memset(&op, 0, sizeof(op));
     op.k = d.kind == HIQ_DESC_SCALE ? 0 : d.k;
     for (int l = 0; l < op.k; ++l) op.slots[l] = d.slots[l];
     std::memcpy(op.lut, d.payload.data(), sizeof(cplx) << op.k);
     if (!group_.empty()) {
          HeldGate& h = group_.back();
          uint64_t tm = 0;
          for (int l = 0; l < h.d.k; ++l) tm |= 1ull << h.d.slots[l];
          // commutes with the most recent held gate: joins its launch as a per-tuple scalar
          if ((slot_mask(op) & tm) == 0) {
               const size_t cap = group_.size() > 1 || !h.direct ? HIQK_TILE_MAX_OPS : HIQK_MAX_DIAG_OPS;
               std::vector<hiqk_diag_op> ops = h.ops;
               std::vector<int> refs = h.refs;
               if (merge_or_append(ops, refs, op, cap)) {
                    bool ok = true;
                    if (group_.size() > 1 || !h.direct) {  // the run must still fit its tile program (table pool)
                         HeldGate last = std::move(group_.back());
                         group_.pop_back();
                         ok = group_accepts(last, ops);
                         group_.push_back(std::move(last));
                    }
                    if (ok) {
                         group_.back().ops.swap(ops);
                         group_.back().refs.swap(refs);
                         return;
                    }
               }
          }
     }
     merge_or_append(pending_, pending_ref_, op, static_cast<size_t>(-1));
}

void Engine::launch_group()
{
     if (group_.empty()) return;
     std::vector<HeldGate> g;
     g.swap(group_);
     if (g.size() == 1 && g[0].direct && !tile_single_) {
          launch(g[0].d, g[0].variant, g[0].ops, g[0].refs);
          return;
     }
     if (g.size() == 1 && g[0].ops.empty() && !tile_single_) {
          launch(g[0].d, g[0].variant, {}, {});
          return;
     }
     const int L = static_cast<int>(locals_.size());
     std::vector<hiqk_tile_step> steps(g.size());
     int n_ref = 0;
     for (size_t i = 0; i < g.size(); ++i) {
          steps[i].k = g[i].d.k;
          for (int l = 0; l < 5; ++l) steps[i].slots[l] = g[i].d.slots[l];
          steps[i].matrix = reinterpret_cast<const double*>(g[i].d.payload.data());
          steps[i].pre = g[i].ops.data();
          steps[i].n_pre = static_cast<int>(g[i].ops.size());
          n_ref += 1;
          for (int r: g[i].refs) n_ref += r;
     }
     if (std::getenv("HIQ_TILE_DEBUG")) {
          std::fprintf(stderr, "tile run (T=%d):", hiqk_tile_program_fits(L, static_cast<int>(steps.size()), steps.data()));
          for (size_t i = 0; i < g.size(); ++i) {
               uint64_t tm = 0;
               for (int l = 0; l < g[i].d.k; ++l) tm |= 1ull << g[i].d.slots[l];
               int n_e = 0;
               for (const hiqk_diag_op& o: g[i].ops) n_e += (slot_mask(o) & tm) ? 1 : 0;
               std::fprintf(stderr, "  [k=%d ks=%d slots", g[i].d.k, hiqk_dense_direct_mixing_bits(g[i].d.k, steps[i].matrix));
               for (int l = 0; l < g[i].d.k; ++l) std::fprintf(stderr, " %d", g[i].d.slots[l]);
               std::fprintf(stderr, " | ops %zu, on targets %d]", g[i].ops.size(), n_e);
          }
          std::fprintf(stderr, "\n");
     }
     TimedPass tp{HIQ_DESC_TILE, static_cast<int>(g.size()), 0, n_ref, nullptr, nullptr};
     if (timing_) {
          tp.start = take_event();
          tp.stop = take_event();
          cudaEventRecord(tp.start, stream_);
     }
     if (dry_run_) {
          std::vector<LaunchStepRef> rec;
          for (const HeldGate& h: g) rec.push_back({&h.d, &h.ops});
          record_launch(HIQ_LAUNCH_TILE, rec);
     }
     else
          cu(hiqk_apply_tile_program(slab_.data(), L, static_cast<int>(steps.size()), steps.data(), stream_));
     ++stats_.gate_launches;
     ++stats_.tile_launches;
     stats_.tile_steps += g.size();
     if (timing_) {
          cudaEventRecord(tp.stop, stream_);
          timed_.push_back(tp);
     }
}

void Engine::flush_pending(size_t keep)
{
     // the held gates first (queued diagonals that touch the last one's targets come after it), then
     // batched diagonal launches (full batches first) until at most `keep` ops remain queued
     launch_group();
     while (pending_.size() > keep) {
          const size_t take = std::min<size_t>(pending_.size(), HIQK_MAX_DIAG_OPS);
          int n_ref = 0;
          for (size_t i = 0; i < take; ++i) n_ref += pending_ref_[i];
          TimedPass tp{HIQ_DESC_DIAG, static_cast<int>(take), 0, n_ref, nullptr, nullptr};
          if (timing_) {
               tp.start = take_event();
               tp.stop = take_event();
               cudaEventRecord(tp.start, stream_);
          }
          if (dry_run_) {
               const std::vector<hiqk_diag_op> batch(pending_.begin(), pending_.begin() + take);
               record_launch(HIQ_LAUNCH_DIAG_BATCH, {{nullptr, &batch}});
          }
          else
               cu(hiqk_apply_diag_batch(slab_.data(), static_cast<int>(locals_.size()), pending_.data(), static_cast<int>(take), stream_));
          ++stats_.gate_launches;
          if (timing_) {
               cudaEventRecord(tp.stop, stream_);
               timed_.push_back(tp);
          }
          pending_.erase(pending_.begin(), pending_.begin() + take);
          pending_ref_.erase(pending_ref_.begin(), pending_ref_.begin() + take);
     }
}

void Engine::launch(const Descriptor& d, int variant, const std::vector<hiqk_diag_op>& ops, const std::vector<int>& refs)
{
     const int L = static_cast<int>(locals_.size());
     int n_ref = 1;
     for (int r: refs) n_ref += r;
     TimedPass tp{d.kind, d.k, variant, n_ref, nullptr, nullptr};
     if (timing_) {
          if (d.kind == HIQ_DESC_DENSE && variant == HIQK_DENSE_DIRECT) {
               // label the reduced product: bits 8.. of the variant = mixing bits when fewer than k
               const int ks = hiqk_dense_direct_mixing_bits(d.k, reinterpret_cast<const double*>(d.payload.data()));
               if (ks > 0 && ks < d.k) tp.variant |= ks << 8;
          }
          tp.start = take_event();
          tp.stop = take_event();
          cudaEventRecord(tp.start, stream_);
     }
     ++stats_.gate_launches;
     if (dry_run_) {
          if (ops.empty()) launches_.push_back(d);
          else record_launch(HIQ_LAUNCH_DENSE_PREDIAG, {{&d, &ops}});
          return;
     }
     switch (d.kind) {
          case HIQ_DESC_DENSE:
               if (!ops.empty())
                    cu(hiqk_apply_dense_prediag(slab_.data(), L, d.k, d.slots, reinterpret_cast<const double*>(d.payload.data()),
                                                ops.data(), static_cast<int>(ops.size()), stream_));
               else
                    cu(hiqk_apply_dense(slab_.data(), L, d.k, d.slots, reinterpret_cast<const double*>(d.payload.data()),
                                        d.ctrl_mask, variant, stream_));
               break;
          case HIQ_DESC_DIAG:
               cu(hiqk_apply_diag(slab_.data(), L, d.k, d.slots, reinterpret_cast<const double*>(d.payload.data()), d.ctrl_mask,
                                  stream_));
               break;
          case HIQ_DESC_SCALE: cu(hiqk_scale(slab_.data(), L, d.payload[0].real(), d.payload[0].imag(), stream_)); break;
          default: break;
     }
     if (timing_) {
          cudaEventRecord(tp.stop, stream_);
          timed_.push_back(tp);
     }
}

void Engine::run()
{
     // fuse and dispatch (reference: SimulatorMPI.cpp:441-541)
     const auto t0 = Clock::now();
     GateMatrix m;
     std::vector<Index> ids, ctrls;
     bool diag = true;
     // refused BEFORE the product is formed: a cluster of n qubits costs 4^n memory and 8^n time to fuse, and a caller that
     // keeps queueing gates after this error (the cluster stays queued, as in the reference) must not pay that every time
     if (fused_.num_qubits() > 5) fail("Run(): cannot apply " + std::to_string(fused_.num_qubits()) + " qubits gate");
     fused_.fuse(m, ids, ctrls, diag);
     Descriptor d;
     d.ctrl_mask = ids_to_bits(ctrls, locals_);
     d.k = static_cast<int>(ids.size());
     for (int l = 0; l < d.k; ++l) d.slots[l] = static_cast<int>(find_sure(locals_, ids[l]));
     if (d.k == 0) {
          if (m.at(0, 0) != cplx(1.0)) {
               d.kind = HIQ_DESC_SCALE;
               d.payload = {m.at(0, 0)};
               ++stats_.scale_passes;
          }
          else {
               d.kind = HIQ_DESC_NONE;
               ++stats_.skipped_passes;
          }
     }
     else if (diag) {
          d.kind = HIQ_DESC_DIAG;
          d.payload.resize(m.dim);
          for (int i = 0; i < m.dim; ++i) d.payload[i] = m.at(i, i);
          ++stats_.diag_passes;
     }
     else {
          d.kind = HIQ_DESC_DENSE;
          d.payload = std::move(m.a);
          ++stats_.dense_passes;
     }
     if (d.kind != HIQ_DESC_NONE) execute(d);
     fused_.reset();
     ++stats_.total_runs;
     stats_.runs_s += seconds_since(t0);
}

// ------------------------------------------------------------------------------------ reductions
void Engine::masks(const std::vector<Index>& ids, const std::vector<bool>& bits, const char* what, uint64_t& lm, uint64_t& lv,
                   uint64_t& gm, uint64_t& gv) const
{
     if (ids.size() != bits.size()) fail(std::string(what) + ": ids.size() != bit_string.size()");
     lm = lv = gm = gv = 0;
     for (size_t i = 0; i < ids.size(); ++i) {
          size_t pos = find(locals_, ids[i]);
          if (pos != kNpos) {
               lm |= 1ull << pos;
               if (bits[i]) lv |= 1ull << pos;
          }
          else {
               pos = find_sure(globals_, ids[i]);
               gm |= 1ull << pos;
               if (bits[i]) gv |= 1ull << pos;
          }
     }
}

double Engine::probability_internal(uint64_t lm, uint64_t lv, uint64_t gm, uint64_t gv)
{
     // reference: SimulatorMPI.cpp:841-870
     // the local sum stays on the device: kernel -> ncclAllReduce in place -> ONE copy to the host and one synchronisation
     double p = 0.0;
     flush_pending();
     if ((static_cast<uint64_t>(rank_) & gm) == gv)
          cu(hiqk_prob_masked(slab_.data(), static_cast<int>(locals_.size()), lm, lv, d_vals_, workspace_, stream_));
     else
          cu(check_cuda(cudaMemsetAsync(d_vals_, 0, sizeof(double), stream_), "cudaMemsetAsync"));
     cu(comm_p_->allreduce_sum_device(d_vals_, 1, stream_));
     d2h(&p, d_vals_, sizeof(double));
     return p;
}

void Engine::normalize(double norm, uint64_t lm, uint64_t lv, uint64_t gm, uint64_t gv)
{
     // reference: SimulatorMPI.cpp:872-895
     const int L = static_cast<int>(locals_.size());
     flush_pending();
     if ((static_cast<uint64_t>(rank_) & gm) != gv)
          cu(check_cuda(cudaMemsetAsync(slab_.data(), 0, sizeof(double2) << L, stream_), "cudaMemsetAsync"));
     else
          cu(hiqk_collapse(slab_.data(), L, lm, lv, 1.0 / std::sqrt(norm), stream_));
}

double Engine::get_probability(const std::vector<bool>& bits, const std::vector<Index>& ids)
{
     uint64_t lm, lv, gm, gv;
     masks(ids, bits, "GetProbability()", lm, lv, gm, gv);
     need_device("GetProbability()");
     return probability_internal(lm, lv, gm, gv);
}

cplx Engine::get_amplitude(const std::vector<bool>& bits, const std::vector<Index>& ids)
{
     // reference: SimulatorMPI.cpp:602-650
     const size_t nq = locals_.size() + globals_.size() - std::count(globals_.begin(), globals_.end(), kNone);
     if (bits.size() != nq || bits.size() != ids.size()) fail("GetAmplitude(): ids.size() != number of qubits");
     uint64_t owner = 0, num = 0, check = 0;
     for (size_t i = 0; i < ids.size(); ++i) {
          size_t pos = find(locals_, ids[i]);
          if (pos != kNpos) {
               if (bits[i]) num |= 1ull << pos;
               check |= 1ull << pos;
          }
          else {
               pos = find_sure(globals_, ids[i]);
               if (bits[i]) owner |= 1ull << pos;
               check |= 1ull << (pos + locals_.size());
          }
     }
     if ((1ull << nq) - 1 != check)
          fail("GetAmplitude(): the second argument must be a permutation of all allocated qubits.");
     need_device("GetAmplitude()");
     cplx value(0.0);
     if (static_cast<uint64_t>(rank_) == owner) {
          d2h(&value, slab_.data() + num, sizeof(cplx));
     }
     cu(comm_p_->broadcast_bytes(&value, sizeof(value), static_cast<int>(owner), stream_));
     return value;
}

void Engine::collapse_wavefunction(const std::vector<Index>& ids, const std::vector<bool>& values)
{
     // reference: SimulatorMPI.cpp:1010-1058
     if (ids.size() != values.size()) fail("collapseWaveFunction(): ids.size() != values.size()");
     uint64_t lm, lv, gm, gv;
     masks(ids, values, "collapseWaveFunction()", lm, lv, gm, gv);
     need_device("collapseWaveFunction()");
     const double norm = probability_internal(lm, lv, gm, gv);
     if (norm < 1.e-12) fail("collapseWaveFunction(): Invalid collapse! Probability is ~0.");
     normalize(norm, lm, lv, gm, gv);
}

std::vector<bool> Engine::measure_qubits(const std::vector<Index>& ids)
{
     // two-level inverse-CDF sampling of the reference (SimulatorMPI.cpp:897-1008)
     need_device("MeasureQubits()");
     flush_pending();
     cu(check_cuda(cudaStreamSynchronize(stream_), "cudaStreamSynchronize"));  // queued gates are not the measurement's time
     const auto t0 = Clock::now();
     const int L = static_cast<int>(locals_.size());
     const uint64_t size = 1ull << L;
     const uint64_t n = std::min(size, kMaxBlocks);
     if (!d_blocks_) cu(check_cuda(cudaMalloc(&d_blocks_, sizeof(double) * kMaxBlocks * world_), "cudaMalloc blocks"));
     cu(hiqk_block_norms(slab_.data(), L, n, d_blocks_ + n * rank_, stream_));
     cu(comm_p_->allgather(d_blocks_ + n * rank_, d_blocks_, n, stream_));
     std::vector<double> tot(n * world_);
     d2h(tot.data(), d_blocks_, sizeof(double) * tot.size());
     const std::vector<double> raw = tot;  // the block sums themselves (the outcome's probability may follow from them)
     // per-rank inclusive prefix, then the running shift over ranks — same order as the reference
     double shift = 0.0;
     for (int r = 0; r < world_; ++r) {
          double* blk = tot.data() + n * r;
          for (uint64_t j = 1; j < n; ++j) blk[j] += blk[j - 1];
          const double new_shift = blk[n - 1];
          for (uint64_t j = 0; j < n; ++j) blk[j] += shift;
          shift += new_shift;
     }
     const double rnd = rng_();
     uint64_t i = 0;
     for (; i < tot.size(); ++i)
          if (rnd <= tot[i]) break;
     const uint64_t src_rank = i / n, src_index = i % n;
     if (static_cast<int>(src_rank) == world_) {
          std::ostringstream o;
          o << "MeasureQubits(): Random number " << rnd << " > norm partial sum " << tot.back();
          fail(o.str());
     }
     const uint64_t block_size = size / n;
     uint64_t k = 0;
     double amp2 = 0.0;  // |amplitude|^2 of the sampled basis state (owner rank)
     if (rank_ == static_cast<int>(src_rank)) {
          double acc = i > 0 ? tot[i - 1] : 0.0;
          std::vector<cplx> blk(block_size);
          k = src_index * block_size;
          d2h(blk.data(), slab_.data() + k, sizeof(cplx) * block_size);
          for (uint64_t j = 0; j < block_size; ++j, ++k) {
               acc += std::norm(blk[j]);
               if (acc >= rnd) break;
          }
          // the scan of the reference may run to the end of the block (k one past it) when rounding leaves acc < rnd
          amp2 = k < (src_index + 1) * block_size ? std::norm(blk[k - src_index * block_size]) : 0.0;
     }
     struct {
          uint64_t index;
          double amp2;
     } msg = {(src_rank << L) + k, amp2};
     cu(comm_p_->broadcast_bytes(&msg, sizeof(msg), static_cast<int>(src_rank), stream_));
     const uint64_t res_index = msg.index;

     std::vector<bool> res(ids.size());
     uint64_t lm = 0, lv = 0, gm = 0, gv = 0;
     for (size_t q = 0; q < ids.size(); ++q) {
          size_t pos = find(locals_, ids[q]);
          if (pos != kNpos) {
               lm |= 1ull << pos;
               if (res_index & (1ull << pos)) {
                    lv |= 1ull << pos;
                    res[q] = true;
               }
          }
          else {
               pos = find_sure(globals_, ids[q]);
               gm |= 1ull << pos;
               if (src_rank & (1ull << pos)) {
                    gv |= 1ull << pos;
                    res[q] = true;
               }
          }
     }
     // Probability of the outcome (reference: a second sweep, getProbability_internal, SimulatorMPI.cpp:996-997).  Two
     // cases need no sweep: every measured local slot lies above the block granularity, so the outcome is a union of
     // whole blocks and its probability the sum of their block sums; or every qubit was measured, and it is the squared
     // modulus of the sampled amplitude.  Otherwise the masked reduction runs.
     double norm;
     const int block_bits = L - static_cast<int>(__builtin_ctzll(n));  // log2(block size)
     uint64_t all_globals = 0;
     for (size_t pos = 0; pos < globals_.size(); ++pos)
          if (globals_[pos] != kNone) all_globals |= 1ull << pos;
     if ((lm & ((1ull << block_bits) - 1ull)) == 0) {
          norm = 0.0;
          const uint64_t bm = lm >> block_bits, bv = lv >> block_bits;
          for (int r = 0; r < world_; ++r) {
               if ((static_cast<uint64_t>(r) & gm) != gv) continue;
               const double* blk = raw.data() + n * r;
               for (uint64_t b = 0; b < n; ++b)
                    if ((b & bm) == bv) norm += blk[b];
          }
     }
     else if (lm == size - 1 && gm == all_globals && msg.amp2 > 0.0) norm = msg.amp2;
     else norm = probability_internal(lm, lv, gm, gv);
     normalize(norm, lm, lv, gm, gv);
     stats_.measures_s += seconds_since(t0);
     return res;
}

double Engine::entropy()
{
     need_device("Entropy()");
     flush_pending();
     double e = 0.0;
     cu(hiqk_entropy(slab_.data(), static_cast<int>(locals_.size()), d_vals_, workspace_, stream_));
     d2h(&e, d_vals_, sizeof(double));
     cu(comm_p_->allreduce_sum(&e, 1, stream_));
     return -e;
}

// ------------------------------------------------------------------------------------ slot maps
std::vector<Index> Engine::qubits_permutation() const
{
     std::vector<Index> res = locals_;
     res.insert(res.end(), globals_.begin(), globals_.end());
     return res;
}

void Engine::set_qubits_permutation(const std::vector<Index>& p)
{
     // relabel only, no data motion (reference: SimulatorMPI.cpp:661-667)
     if (p.size() < locals_.size() || p.size() < globals_.size()) fail("SetQubitsPermutation(): permutation too short");
     {
          // a relabelling: the same labels (empty global positions included) in another order — anything else would put a
          // qubit in two places of the maps or drop one (the reference takes the list as it comes)
          std::vector<Index> a = qubits_permutation(), b = p;
          std::sort(a.begin(), a.end());
          std::sort(b.begin(), b.end());
          if (a != b) fail("SetQubitsPermutation(): not a permutation of the current qubit ids " + list_str(qubits_permutation()));
     }
     locals_.assign(p.begin(), p.begin() + locals_.size());
     globals_.assign(p.end() - globals_.size(), p.end());
}

std::map<Index, int> Engine::id2pos() const
{
     std::map<Index, int> m;
     for (size_t pos = 0; pos < locals_.size(); ++pos) m[locals_[pos]] = static_cast<int>(pos);
     for (size_t pos = 0; pos < globals_.size(); ++pos)
          if (globals_[pos] != kNone) m[globals_[pos]] = static_cast<int>(pos + locals_.size());
     return m;
}

void Engine::copy_slab_to_host(void* dst, uint64_t cap_amps)
{
     need_device("cheat_local()");
     const uint64_t n = 1ull << locals_.size();
     if (cap_amps < n) fail("cheat_local(): destination buffer too small");
     d2h(dst, slab_.data(), n * sizeof(double2));
}

void Engine::copy_slab_from_host(const void* src, uint64_t n_amps)
{
     need_device("set_local_slab()");
     if (n_amps != (1ull << locals_.size())) fail("set_local_slab(): size must equal 2^(local qubits)");
     group_.clear();  // the whole slab is overwritten
     pending_.clear();
     pending_ref_.clear();
     cu(check_cuda(cudaMemcpyAsync(slab_.data(), src, n_amps * sizeof(double2), cudaMemcpyHostToDevice, stream_), "cudaMemcpyAsync"));
     synchronize();
     stats_.h2d_bytes += static_cast<double>(n_amps * sizeof(double2));
}

// ------------------------------------------------------------------------------------ swaps
void Engine::swap_qubits_stage(const std::vector<Index>& pairs)
{
     const auto t0 = Clock::now();
     swap_qubits(pairs);
     if (!dry_run_) synchronize();
     stats_.swaps_s += seconds_since(t0);
     ++stats_.total_swaps;
     ++stats_.total_stages;
}

void Engine::swap_qubits(const std::vector<Index>& pairs)
{
     // pairs = [global id, local id, ...]; afterwards the ids trade places (reference: SimulatorMPI.cpp:1085-1138)
     if (pairs.size() % 2) fail("SwapQubits(): odd number of ids");
     // gates still waiting in the fusion were given for the layout as it is: they go out before qubits trade places (the
     // reference's wrapper runs before every swap / allocation / measurement, _simulator_mpi.py:505-507; a caller of the
     // class that does not would find its gate's qubit global in Run(), on the ranks that kept the gate)
     if (!fused_.empty()) run();
     std::map<Index, size_t> pos;
     for (size_t i = 0; i < pairs.size(); i += 2) {
          pos[pairs[i]] = find_sure(globals_, pairs[i]);
          pos[pairs[i + 1]] = find_sure(locals_, pairs[i + 1]);
     }
     if (pos.size() != pairs.size()) fail("SwapQubits(): each qubit should be unique");
     std::vector<int> gpos, slots;
     for (size_t i = 0; i < pairs.size(); i += 2) {
          gpos.push_back(static_cast<int>(pos[pairs[i]]));
          slots.push_back(static_cast<int>(pos[pairs[i + 1]]));
     }
     if (tracing_) {
          Descriptor d;
          d.kind = HIQ_DESC_SWAP;
          d.k = static_cast<int>(gpos.size());
          for (size_t i = 0; i < gpos.size(); ++i) {
               d.aux.push_back(gpos[i]);
               d.aux.push_back(slots[i]);
          }
          trace_op(d);
     }
     if (!dry_run_ && !gpos.empty()) {
          flush_pending();
          TimedPass tp{HIQ_DESC_SWAP, static_cast<int>(gpos.size()), 0, 0, nullptr, nullptr};
          if (timing_) {
               tp.start = take_event();
               tp.stop = take_event();
               swap_mark_[0] = take_event();
               swap_mark_[1] = take_event();
               swap_marked_ = false;
               cudaEventRecord(tp.start, stream_);
          }
          exchange(gpos, slots);
          if (timing_) {
               cudaEventRecord(tp.stop, stream_);
               if (swap_marked_) {
                    // variant 1 = waiting for the slowest peer of the group to arrive (rank skew: ranks whose global
                    // control bits are 0 skip gates, SimulatorMPI.cpp:735-738), variant 0 = the exchange itself
                    timed_.push_back(TimedPass{HIQ_DESC_SWAP, tp.k, 1, 0, tp.start, swap_mark_[0]});
                    tp.start = swap_mark_[1];
               }
               else {
                    event_pool_.push_back(swap_mark_[0]);
                    event_pool_.push_back(swap_mark_[1]);
               }
               timed_.push_back(tp);
               swap_mark_[0] = swap_mark_[1] = nullptr;
          }
     }
     for (size_t i = 0; i < gpos.size(); ++i) std::swap(locals_[slots[i]], globals_[gpos[i]]);
}

void Engine::exchange(const std::vector<int>& gpos, const std::vector<int>& slots)
{
     // Two transports for the same transposition:
     //   peer-mapped  one kernel per GPU swaps the pairs in place through NVLink loads/stores, no staging and
     //                no extra HBM passes; its accesses are runs of 2^(lowest swapped slot) amplitudes
     //   staged       pack -> NCCL send/recv -> unpack pipeline; contiguous messages whatever the slots
     const int lowest = *std::min_element(slots.begin(), slots.end());
     const bool want_packed = swap_mode_ == 3 || (swap_mode_ == 0 && packed_enabled_ && lowest < packed_below_slot_);
     if (want_packed && !comm_p_->packed().failed && exchange_packed(gpos, slots)) return;
     if (swap_mode_ == 3) fail(std::string("SwapQubits(): packed exchange unavailable: ") + hiq_last_error());
     const bool want_p2p = swap_mode_ == 2 || (swap_mode_ == 0 && lowest >= min_p2p_slot_);
     if (want_p2p && !comm_p_->p2p_broken && exchange_p2p(gpos, slots)) return;
     if (swap_mode_ == 2) fail(std::string("SwapQubits(): peer-mapped exchange unavailable: ") + hiq_last_error());
     exchange_staged(gpos, slots);
}

void Engine::group_barrier(const std::vector<int>& peer_ranks)
{
     // stream-ordered barrier among the swap group: a 1-element send/recv with every peer completes
     // only after the peer's stream has reached the same point
     double* buf = d_vals_ + 16;
     nccl().GroupStart();
     for (size_t k = 0; k < peer_ranks.size(); ++k) {
          nccl().Send(buf, 1, ncclDouble, peer_ranks[k], comm_p_->handle(), stream_);
          nccl().Recv(buf + 1 + k, 1, ncclDouble, peer_ranks[k], comm_p_->handle(), stream_);
     }
     ncclResult_t r = nccl().GroupEnd();
     if (r != ncclSuccess) throw EngineError(HIQ_ERR_CUDA, std::string("ncclGroupEnd: ") + nccl().GetErrorString(r));
}

bool Engine::map_peer_chunks(const Slab& mine, std::vector<PeerView>& views, uint64_t tag, const std::vector<int>& peer_ranks)
{
     // Handshake: send every listed peer the chunks of `mine` it has not seen yet, and map the chunks the peers send
     // me.  Buffers grow in lock-step on all ranks (allocation is collective), so a peer's view is complete when it
     // has as many chunks as my own buffer.  `tag` names the buffer generation (engine epoch for slabs, 0 for the
     // process-wide staging buffer); messages with another tag belong to a buffer that no longer exists and are dropped.
     FdChannel& ch = comm_p_->fds();
     if (!ch.is_open()) {
          set_error(HIQ_ERR_RUNTIME, "descriptor channel is not open");
          return false;
     }
     if (views.empty()) views.resize(world_);
     struct Out {
          int rank;
          size_t index;
     };
     std::vector<Out> outbox;
     for (int pr: peer_ranks) {
          PeerView& v = views[pr];
          for (size_t i = v.sent; i < mine.n_chunks(); ++i) outbox.push_back({pr, i});
     }
     auto complete = [&] {
          for (int pr: peer_ranks)
               if (views[pr].slab.n_chunks() < mine.n_chunks()) return false;
          return true;
     };
     auto receive_one = [&](int timeout_ms) -> bool {
          FdMessage m;
          if (ch.recv_fd(m, timeout_ms) != HIQ_OK) return false;
          if (m.src_rank < 0 || m.src_rank >= world_) {
               ::close(m.fd);
               set_error(HIQ_ERR_RUNTIME, "peer ipc: message from an unknown rank");
               return false;
          }
          PeerView& v = views[m.src_rank];
          if (m.epoch != tag) {  // a message of another buffer generation: not for this handshake
               ::close(m.fd);
               return true;
          }
          if (!v.slab.data()) {
               if (v.slab.init(device_, mine.reserved_bytes()) != HIQ_OK) {
                    ::close(m.fd);
                    return false;
               }
               v.slab.epoch = m.epoch;
          }
          if (m.index != v.slab.n_chunks()) {
               ::close(m.fd);
               set_error(HIQ_ERR_RUNTIME, "peer ipc: chunk arrived out of order");
               return false;
          }
          return v.slab.map_next_chunk(m.fd, m.size) == HIQ_OK;
     };
     size_t next = 0;
     const auto t0 = Clock::now();
     while (next < outbox.size() || !complete()) {
          if (seconds_since(t0) > 20.0) {
               set_error(HIQ_ERR_RUNTIME, "peer ipc: handshake timed out");
               return false;
          }
          if (next < outbox.size()) {
               const Out& o = outbox[next];
               FdMessage m;
               m.index = static_cast<uint32_t>(o.index);
               m.total = static_cast<uint32_t>(mine.n_chunks());
               m.size = mine.chunk_bytes(o.index);
               m.epoch = tag;
               if (mine.export_chunk(o.index, &m.fd) != HIQ_OK) return false;
               const int rc = ch.send_fd(o.rank, m, 0 /* do not wait: drain my own queue instead */);
               ::close(m.fd);
               if (rc == HIQ_OK) {
                    views[o.rank].sent = o.index + 1;
                    ++next;
                    continue;
               }
               // the peer's queue is full (or it is not there yet): make progress on my side, then retry
               if (!complete()) receive_one(5);
               continue;
          }
          if (!receive_one(20000)) return false;
     }
     return true;
}

bool Engine::ensure_peer_views(const std::vector<int>& peer_ranks)
{
     // The handshake runs only when some view is incomplete — a condition that is the same on every rank,
     // because slabs grow in lock-step — and its outcome is agreed on by the whole world, so that either
     // all ranks take the peer-mapped path or all fall back (swap: staged exchange; emulate_math: error).
     if (comm_p_->p2p_broken) {
          set_error(HIQ_ERR_RUNTIME, "peer-mapped slabs are unavailable in this process group");
          return false;
     }
     bool stale = peer_views_.empty();
     for (int pr: peer_ranks)
          if (!stale && (peer_views_[pr].sent < slab_.n_chunks() || peer_views_[pr].slab.n_chunks() < slab_.n_chunks())) stale = true;
     if (stale) {
          const auto t_map = Clock::now();
          double failed = map_peer_chunks(slab_, peer_views_, epoch_, peer_ranks) ? 0.0 : 1.0;
          stats_.peer_map_s += seconds_since(t_map);
          const std::string why = failed != 0.0 ? hiq_last_error() : "";
          cu(comm_p_->allreduce_sum(&failed, 1, stream_));
          if (failed != 0.0) {
               comm_p_->p2p_broken = true;
               set_error(HIQ_ERR_RUNTIME, why.empty() ? "a peer rank could not map the slabs" : why);
               return false;
          }
     }
     return true;
}

bool Engine::exchange_p2p(const std::vector<int>& gpos, const std::vector<int>& slots)
{
     const int L = static_cast<int>(locals_.size());
     const int q = static_cast<int>(gpos.size());
     if (q > 3) {
          set_error(HIQ_ERR_RUNTIME, "more than 3 swapped pairs");
          return false;
     }
     std::vector<int> order(q);
     for (int i = 0; i < q; ++i) order[i] = i;
     std::sort(order.begin(), order.end(), [&](int a, int b) { return slots[a] < slots[b]; });
     auto pattern_of = [&](int r) {
          uint64_t pat = 0;
          for (int j = 0; j < q; ++j)
               if ((r >> gpos[order[j]]) & 1) pat |= 1ull << j;
          return pat;
     };
     std::vector<int> peer_ranks;
     for (int x = 1; x < (1 << q); ++x) {
          int pr = rank_;
          for (int i = 0; i < q; ++i)
               if ((x >> i) & 1) pr ^= 1 << gpos[i];
          peer_ranks.push_back(pr);
     }
     if (!ensure_peer_views(peer_ranks)) return false;
     const uint64_t n = 1ull << (L - q);
     const uint64_t half = (n + 1) / 2;
     std::vector<void*> ptrs;
     std::vector<uint64_t> pats, begins, counts;
     for (int pr: peer_ranks) {
          ptrs.push_back(peer_views_[pr].slab.data());
          pats.push_back(pattern_of(pr));
          begins.push_back(rank_ < pr ? 0 : half);  // the lower rank of a pair takes the lower half of the free indices
          counts.push_back(rank_ < pr ? half : n - half);
     }
     group_barrier(peer_ranks);  // every peer has finished the gates before its slab is touched
     if (swap_mark_[0]) {
          cudaEventRecord(swap_mark_[0], stream_);
          cudaEventRecord(swap_mark_[1], stream_);
          swap_marked_ = true;
     }
     cu(hiqk_swap_p2p(slab_.data(), ptrs.data(), static_cast<int>(ptrs.size()), L, q, slots.data(), pats.data(), pattern_of(rank_),
                      begins.data(), counts.data(), stream_));
     group_barrier(peer_ranks);  // ... and nobody moves on while a peer still writes into its slab
     stats_.swap_bytes_sent += static_cast<double>(peer_ranks.size()) * n * sizeof(double2);
     ++stats_.swaps_p2p;
     return true;
}

bool Engine::ensure_packed_staging(size_t want_bytes, size_t min_bytes)
{
     // Collective over the world (every rank takes part in every swap, with the same arguments).  The staging buffer is a
     // virtual-memory allocation like the slab, shared the same way: its chunks travel to the peers as file descriptors
     // and are mapped there with access for the peer's GPU only.  (CUDA runtime IPC handles were measured to slow every
     // later cuMemMap / cuMemSetAccess of the process by 15-30x once peer access had been enabled through them.)
     // The buffer belongs to the process (Comm::packed()): the engines a caller creates one after the other share it.
     Comm::PackedStaging& S = comm_p_->packed();
     if (S.failed) {
          set_error(HIQ_ERR_RUNTIME, "the packed exchange could not be set up in this process group");
          return false;
     }
     std::vector<int> others;
     for (int r = 0; r < world_; ++r)
          if (r != rank_) others.push_back(r);
     auto views_complete = [&] {
          if (S.peers.empty()) return false;
          for (int r: others)
               if (S.peers[r].sent < S.mine.n_chunks() || S.peers[r].slab.n_chunks() < S.mine.n_chunks()) return false;
          return true;
     };
     if (S.mine.data() && want_bytes <= S.wanted && S.bytes() >= min_bytes && views_complete()) return true;
     if (!S.mine.data()) {
          if (S.mine.init(device_, (16ull << 30) / sizeof(double2), true) != HIQ_OK) {
               S.failed = true;
               return false;
          }
     }
     // grow to the wanted size, halving the request until every rank could map it
     size_t bytes = std::max(want_bytes, S.bytes());
     for (;;) {
          double failed = S.mine.ensure(bytes / sizeof(double2)) == HIQ_OK ? 0.0 : 1.0;
          cu(comm_p_->allreduce_sum(&failed, 1, stream_));
          if (failed == 0.0) break;
          bytes /= 2;
          if (bytes < min_bytes || bytes <= S.bytes()) {
               if (S.bytes() >= min_bytes) break;  // what is mapped already will do (more pieces)
               S.failed = true;
               set_error(HIQ_ERR_RUNTIME, "no device memory for the staging buffer of the packed exchange");
               return false;
          }
     }
     double failed = map_peer_chunks(S.mine, S.peers, 0 /* tag of the staging buffer */, others) ? 0.0 : 1.0;
     const std::string why = failed != 0.0 ? hiq_last_error() : "";
     cu(comm_p_->allreduce_sum(&failed, 1, stream_));
     if (failed != 0.0) {
          S.failed = true;
          set_error(HIQ_ERR_RUNTIME, why.empty() ? "a peer rank could not map the staging buffers" : why);
          return false;
     }
     S.wanted = std::max(S.wanted, want_bytes);
     return true;
}

bool Engine::exchange_packed(const std::vector<int>& gpos, const std::vector<int>& slots)
{
     // Same transposition as exchange_staged(), but the wire carries contiguous full-line traffic whatever the swapped
     // slots are (the in-place kernel makes 16-64 B runs when a slot is below 3, and NVLink moves 32 B sectors): piece i
     // is gathered from my slab straight into the PEERS' staging buffers (posted NVLink writes), a stream-ordered barrier
     // tells the group the pieces have landed, and every rank scatters its own staging buffer into its slab (local, HBM
     // speed).  Measured on 2 B200 at L = 32 (profiles/r02c_swap_n2_*.jsonl): 601 GB/s per direction for a swapped slot 1
     // or 2 and 500 for slot 0, against 387-504 and 316 for the in-place kernel; pulling the pieces with NVLink reads
     // instead was slower (495-545) and is not kept.
     // Pipeline over two streams: the gathers and the barriers run on the engine stream, the scatters on the second one,
     // so the scatter of piece i overlaps the gather of piece i + 1.  Two staging buffers; barrier i also certifies
     // that everybody has scattered piece i - 1, which is what frees the buffer piece i + 1 goes into.
     const int L = static_cast<int>(locals_.size());
     const int q = static_cast<int>(gpos.size());
     if (q > 3) {
          set_error(HIQ_ERR_RUNTIME, "more than 3 swapped pairs");
          return false;
     }
     std::vector<int> order(q);
     for (int i = 0; i < q; ++i) order[i] = i;
     std::sort(order.begin(), order.end(), [&](int a, int b) { return slots[a] < slots[b]; });
     const uint64_t chunk = 1ull << (L - q);
     const int n_peers = (1 << q) - 1;
     constexpr size_t kStagingMax = 8ull << 30;
     const size_t full = 2ull * n_peers * chunk * sizeof(double2);
     const size_t floor_bytes = 2ull * n_peers * std::min<uint64_t>(chunk, 1ull << 16) * sizeof(double2);
     if (!ensure_packed_staging(std::min(full, kStagingMax), floor_bytes)) return false;
     Comm::PackedStaging& S = comm_p_->packed();
     uint64_t piece = chunk;
     while (2ull * n_peers * piece * sizeof(double2) > S.bytes()) piece >>= 1;
     while (packed_piece_cap_ && piece > packed_piece_cap_ && piece > 1) piece >>= 1;
     std::vector<int> peer_ranks;
     std::vector<uint64_t> pats;
     for (int x = 1; x < (1 << q); ++x) {  // peer k of rank r is r ^ bits(x): the same k names me in the peer's list
          int pr = rank_;
          for (int i = 0; i < q; ++i)
               if ((x >> i) & 1) pr ^= 1 << gpos[i];
          uint64_t pat = 0;
          for (int j = 0; j < q; ++j)
               if ((pr >> gpos[order[j]]) & 1) pat |= 1ull << j;
          peer_ranks.push_back(pr);
          pats.push_back(pat);
     }
     if (!swap_events_[0]) {
          for (auto& ev: swap_events_) cu(check_cuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate"));
     }
     cudaEvent_t* gathered = &swap_events_[0];   // [2] piece gathered and certified by the barrier (engine stream)
     cudaEvent_t* scattered = &swap_events_[2];  // [2] piece scattered into my slab (second stream)
     // Entry barrier with the peers of THIS exchange.  The staging buffers belong to the process, and consecutive swaps may
     // pair a rank with different peers (another subset of the global bits): a peer that has not reached this exchange may
     // still be scattering the previous one out of the very buffer the first gather writes into.  Its stream reaches this
     // barrier only after that scatter (and the closing barrier of its previous exchange) has completed.
     group_barrier(peer_ranks);
     if (swap_mark_[0]) {
          cudaEventRecord(swap_mark_[0], stream_);
          cudaEventRecord(swap_mark_[1], stream_);
          swap_marked_ = true;
     }
     const uint64_t n_pieces = chunk / piece;
     std::vector<void*> mine(n_peers), theirs(n_peers);
     for (uint64_t i = 0; i < n_pieces; ++i) {
          const int b = static_cast<int>(i % 2);
          for (int k = 0; k < n_peers; ++k) {
               const size_t off = (static_cast<size_t>(b) * n_peers + k) * piece;
               mine[k] = S.mine.data() + off;
               theirs[k] = S.peers[peer_ranks[k]].slab.data() + off;
          }
          cu(hiqk_swap_move(slab_.data(), L, q, slots.data(), n_peers, pats.data(), i * piece, piece, theirs.data(), 1, stream_));
          if (i >= 1) cu(check_cuda(cudaStreamWaitEvent(stream_, scattered[(i - 1) % 2], 0), "cudaStreamWaitEvent"));
          group_barrier(peer_ranks);
          cu(check_cuda(cudaEventRecord(gathered[b], stream_), "cudaEventRecord"));
          cu(check_cuda(cudaStreamWaitEvent(comm_stream_, gathered[b], 0), "cudaStreamWaitEvent"));
          cu(hiqk_swap_move(slab_.data(), L, q, slots.data(), n_peers, pats.data(), i * piece, piece, mine.data(), 0, comm_stream_));
          cu(check_cuda(cudaEventRecord(scattered[b], comm_stream_), "cudaEventRecord"));
     }
     // the slab is complete, and nobody starts the next exchange (or frees anything) while a peer still reads a buffer
     cu(check_cuda(cudaStreamWaitEvent(stream_, scattered[(n_pieces - 1) % 2], 0), "cudaStreamWaitEvent"));
     group_barrier(peer_ranks);
     stats_.swap_bytes_sent += static_cast<double>(n_peers) * chunk * sizeof(double2);
     ++stats_.swaps_packed;
     return true;
}

void Engine::exchange_staged(const std::vector<int>& gpos, const std::vector<int>& slots)
{
     // Net effect: transpose global-index bit (L + gpos_i) with bit slot_i.
     // Peer p differs from this rank in a non-empty subset of the swapped global bits; the
     // amplitudes whose swapped slots spell p's bits go to p and are replaced, in place, by p's
     // amplitudes whose swapped slots spell this rank's bits.
     //
     // Pipeline: the free-index range is cut into pieces; piece i is packed on the engine stream,
     // exchanged on the communication stream (one NCCL group of send/recv to the 2^q - 1 peers) and
     // unpacked on the engine stream, with two staging buffers so that pack(i+1) and unpack(i-1)
     // overlap the NVLink transfer of piece i.  Pieces touch disjoint amplitudes, so in place is safe.
     const int L = static_cast<int>(locals_.size());
     const int q = static_cast<int>(gpos.size());
     std::vector<int> order(q);  // pair indices by ascending slot
     for (int i = 0; i < q; ++i) order[i] = i;
     std::sort(order.begin(), order.end(), [&](int a, int b) { return slots[a] < slots[b]; });
     const uint64_t chunk = 1ull << (L - q);
     const int n_peers = (1 << q) - 1;
     const uint64_t piece = std::min(chunk, kSwapPieceAmps);
     constexpr int NBUF = 2;
     const size_t need = 2ull * NBUF * n_peers * piece * sizeof(double2);
     if (!swap_buf_ || swap_buf_bytes_ < need) {
          if (swap_buf_) cudaFree(swap_buf_);
          swap_buf_ = nullptr;
          cu(check_cuda(cudaMalloc(&swap_buf_, need), "cudaMalloc swap staging"));
          swap_buf_bytes_ = need;
     }
     const uint64_t buf_amps = static_cast<uint64_t>(n_peers) * piece;
     double2* send[NBUF];
     double2* recv[NBUF];
     for (int b = 0; b < NBUF; ++b) {
          send[b] = static_cast<double2*>(swap_buf_) + (2 * b) * buf_amps;
          recv[b] = static_cast<double2*>(swap_buf_) + (2 * b + 1) * buf_amps;
     }
     if (!swap_events_[0]) {
          for (auto& ev: swap_events_) cu(check_cuda(cudaEventCreateWithFlags(&ev, cudaEventDisableTiming), "cudaEventCreate"));
     }
     cudaEvent_t* packed = &swap_events_[0];      // [NBUF] piece packed on stream_
     cudaEvent_t* exchanged = &swap_events_[NBUF];  // [NBUF] piece exchanged on comm_stream_
     struct Peer {
          int rank;
          uint64_t pat;  // pattern over the sorted swapped slots
     };
     std::vector<Peer> peers;
     for (int x = 1; x < (1 << q); ++x) {
          int pr = rank_;
          for (int i = 0; i < q; ++i)
               if ((x >> i) & 1) pr ^= 1 << gpos[i];
          uint64_t pat = 0;
          for (int j = 0; j < q; ++j)
               if ((pr >> gpos[order[j]]) & 1) pat |= 1ull << j;
          peers.push_back({pr, pat});
     }
     const uint64_t n_pieces = chunk / piece;
     auto unpack_piece = [&](uint64_t i) {
          const int b = static_cast<int>(i % NBUF);
          cu(check_cuda(cudaStreamWaitEvent(stream_, exchanged[b], 0), "cudaStreamWaitEvent"));
          for (int k = 0; k < n_peers; ++k)
               cu(hiqk_swap_unpack(slab_.data(), L, q, slots.data(), peers[k].pat, i * piece, piece, recv[b] + k * piece, stream_));
     };
     for (uint64_t i = 0; i < n_pieces; ++i) {
          const int b = static_cast<int>(i % NBUF);
          // send[b]/recv[b] were last used by piece i - NBUF, whose unpack is already queued on stream_
          for (int k = 0; k < n_peers; ++k)
               cu(hiqk_swap_pack(slab_.data(), L, q, slots.data(), peers[k].pat, i * piece, piece, send[b] + k * piece, stream_));
          cu(check_cuda(cudaEventRecord(packed[b], stream_), "cudaEventRecord"));
          cu(check_cuda(cudaStreamWaitEvent(comm_stream_, packed[b], 0), "cudaStreamWaitEvent"));
          nccl().GroupStart();
          for (int k = 0; k < n_peers; ++k) {
               nccl().Send(send[b] + k * piece, piece * 2, ncclDouble, peers[k].rank, comm_p_->handle(), comm_stream_);
               nccl().Recv(recv[b] + k * piece, piece * 2, ncclDouble, peers[k].rank, comm_p_->handle(), comm_stream_);
          }
          ncclResult_t r = nccl().GroupEnd();
          if (r != ncclSuccess) throw EngineError(HIQ_ERR_CUDA, std::string("ncclGroupEnd: ") + nccl().GetErrorString(r));
          cu(check_cuda(cudaEventRecord(exchanged[b], comm_stream_), "cudaEventRecord"));
          if (i >= 1) unpack_piece(i - 1);
     }
     unpack_piece(n_pieces - 1);
     stats_.swap_bytes_sent += static_cast<double>(n_peers) * chunk * sizeof(double2);
     ++stats_.swaps_staged;
}

}  // namespace hiq
