// Host engine: one instance per rank/GPU.  Owns the amplitude slab (HBM), the qubit->slot maps,
// the gate-fusion accumulator and the RNG, and turns the reference's operator API into fused-gate
// descriptors, swap plans and reduction requests executed by the sm_100a kernels.
//
// Behavioural spec: the reference class SimulatorMPI
// (reference: src/simulator-mpi/SimulatorMPI.hpp:43-308, SimulatorMPI.cpp).  Every rank-dependent
// rule (global-control filter, diagonal slicing, allocation policy, measurement ordering, swap
// colours) is evaluated exactly as the reference does for an MPI rank.
#pragma once
#include <chrono>
#include <cstdint>
#include <functional>
#include <map>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hiq_b200.h"
#include "comm.hpp"
#include "fusion.hpp"
#include "slab.hpp"

namespace hiq {

struct EngineError : std::runtime_error {
     int code;
     EngineError(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// What the host hands to the device layer; also the unit of the dry-run trace.
struct Descriptor {
     int kind = 0;  // HIQ_DESC_*
     int k = 0;
     int slots[5] = {0, 0, 0, 0, 0};
     uint64_t ctrl_mask = 0;
     std::vector<cplx> payload;   // dense: 4^k, diag: 2^k, scale: 1
     std::vector<int64_t> aux;    // swap: gpos0, slot0, gpos1, slot1, ...
};

struct EngineStats {
     uint64_t total_gates = 0, total_runs = 0, total_stages = 0, total_swaps = 0;
     uint64_t dense_passes = 0, diag_passes = 0, scale_passes = 0, skipped_passes = 0;
     double runs_s = 0, swaps_s = 0, measures_s = 0, allocs_s = 0, deallocs_s = 0;
     double swap_bytes_sent = 0;
     uint64_t swaps_p2p = 0, swaps_staged = 0, swaps_packed = 0;
     double h2d_bytes = 0, d2h_bytes = 0;
     uint64_t gate_launches = 0;  // device launches that carried the dense/diag/scale passes above
     double ctor_s = 0, slab_grow_s = 0, peer_map_s = 0;  // host seconds: constructor, mapping physical memory, peer-slab handshakes
     uint64_t tile_launches = 0, tile_steps = 0;           // multi-gate tile-resident launches and the dense gates they carried
};

class Engine {
public:
     Engine(uint64_t seed, int max_local, int max_cluster, int rank, int world_size, const void* nccl_id, int device,
            int flags);
     ~Engine();
     Engine(const Engine&) = delete;
     Engine& operator=(const Engine&) = delete;

     void allocate_qubit(Index id);
     void allocate_qureg(const std::vector<Index>& ids, cplx init);
     void deallocate_qubit(Index id);
     void apply_gate(GateMatrix m, std::vector<Index> ids, std::vector<Index> ctrls);
     void run();
     void swap_qubits_stage(const std::vector<Index>& pairs);  // SwapQubitsWrapper
     std::vector<bool> measure_qubits(const std::vector<Index>& ids);
     double get_probability(const std::vector<bool>& bits, const std::vector<Index>& ids);
     cplx get_amplitude(const std::vector<bool>& bits, const std::vector<Index>& ids);
     void collapse_wavefunction(const std::vector<Index>& ids, const std::vector<bool>& values);
     double entropy();
     std::vector<Index> qubits_permutation() const;
     const std::vector<Index>& locals() const { return locals_; }
     const std::vector<Index>& globals() const { return globals_; }
     void set_qubits_permutation(const std::vector<Index>& p);
     std::map<Index, int> id2pos() const;
     // copy the local slab to host memory (complex128, 2^L amplitudes)
     void copy_slab_to_host(void* dst, uint64_t cap_amps);
     void copy_slab_from_host(const void* src, uint64_t n_amps);
     double2* slab_ptr()
     {
          flush_pending();  // the caller is about to look at the amplitudes
          return slab_.data();
     }
     int local_qubits() const { return static_cast<int>(locals_.size()); }
     void synchronize();

     int rank() const { return rank_; }
     int world_size() const { return world_; }
     bool dry_run() const { return dry_run_; }
     cudaStream_t stream() const { return stream_; }
     const EngineStats& stats() const { return stats_; }
     struct PassTime {
          int kind, k, variant;
          int n_ref;  // fused-gate passes of the reference's plan carried by this launch
          double ms;
     };
     std::vector<PassTime> collect_timings();
     const std::vector<Descriptor>& trace() const { return trace_; }
     // dry-run engines only: what would really be LAUNCHED, in device order — the plan's non-gate descriptors as they are,
     // single gates as they are, and HIQ_DESC_LAUNCH records for launches that carry several passes of the plan
     const std::vector<Descriptor>& launch_trace() const { return launches_; }
     void clear_trace()
     {
          trace_.clear();
          launches_.clear();
     }
     void set_dense_variant(int v) { dense_variant_ = v; }

     // ---- operator-level calls of the reference wrapper that the reference class lacks (engine_ops.cpp;
     // reference call sites: _simulator_mpi.py:180-183, 220-223, 305, 459-468, 377-380; semantics: ProjectQ simulator.hpp)
     struct PauliTerm {
          std::vector<std::pair<int, char>> factors;  // (index into ids, 'X' | 'Y' | 'Z'), applied in this order
          cplx coef;
     };
     double get_expectation_value(const std::vector<PauliTerm>& terms, const std::vector<Index>& ids);
     void apply_qubit_operator(const std::vector<PauliTerm>& terms, const std::vector<Index>& ids);
     void emulate_time_evolution(const std::vector<PauliTerm>& terms, double time, const std::vector<Index>& ids,
                                 const std::vector<Index>& ctrls);
     void set_wavefunction(const cplx* amps, uint64_t n_amps, const std::vector<Index>& ordering);
     // kind = HIQK_PERM_*; fwd_table (TABLE only) = f(v) for every value of the concatenated registers
     void emulate_math(int kind, uint64_t a, uint64_t N, const std::vector<uint64_t>& fwd_table, const std::vector<Index>& reg_ids,
                       const std::vector<Index>& ctrls);
     // cheat(): every rank receives the concatenation of all rank slabs (world * 2^L amplitudes, rank-major)
     void gather_state_to_host(void* dst, uint64_t cap_amps);

private:
     using PeerView = Comm::PeerView;
     static constexpr Index kNone = -1;  // empty global slot (reference: kNotFound_ stored in an int64)
     using Clock = std::chrono::steady_clock;

     [[noreturn]] void fail(const std::string& msg) const;
     void cu(int rc) const;
     size_t find(const std::vector<Index>& v, Index val) const;       // npos when absent
     size_t find_sure(const std::vector<Index>& v, Index val) const;  // throws when absent
     uint64_t ids_to_bits(const std::vector<Index>& ids, const std::vector<Index>& perm) const;
     void need_device(const char* what) const;

     void allocate_local(Index id);
     void grow_slab(uint64_t amps);  // map the slab up to `amps` amplitudes (the scratch slab gives way when memory is short)
     void allocate_global(Index id);
     void deallocate_local(Index id);
     void deallocate_global(Index id);
     void swap_qubits(const std::vector<Index>& pairs);
     void exchange(const std::vector<int>& gpos, const std::vector<int>& slots);
     void exchange_staged(const std::vector<int>& gpos, const std::vector<int>& slots);  // pack -> NCCL send/recv -> unpack
     bool exchange_p2p(const std::vector<int>& gpos, const std::vector<int>& slots);     // in place over peer-mapped slabs
     // handshake over the descriptor channel: send the peers the chunks of `mine` they have not seen, map theirs
     bool map_peer_chunks(const Slab& mine, std::vector<PeerView>& views, uint64_t tag, const std::vector<int>& peer_ranks);
     bool ensure_peer_views(const std::vector<int>& peer_ranks);  // world-agreed: all ranks succeed or all fail
     struct PauliGroup {
          uint64_t lx = 0;  // flip mask over local slots
          int gx = 0;       // flip mask over rank bits: the source amplitudes live on rank ^ gx
          std::vector<hiqk_pauli_term> terms;
     };
     std::vector<PauliGroup> pauli_groups(const std::vector<PauliTerm>& terms, const std::vector<Index>& ids, const char* what) const;
     void exchange_piece(int partner, uint64_t begin, uint64_t count, double2* staging);
     double2* ensure_staging(uint64_t amps);
     void group_barrier(const std::vector<int>& peer_ranks);
     void masks(const std::vector<Index>& ids, const std::vector<bool>& bits, const char* what, uint64_t& lm, uint64_t& lv,
                uint64_t& gm, uint64_t& gv) const;
     double probability_internal(uint64_t lm, uint64_t lv, uint64_t gm, uint64_t gv);
     void normalize(double norm, uint64_t lm, uint64_t lv, uint64_t gm, uint64_t gv);
     void execute(const Descriptor& d);
     void launch(const Descriptor& d, int variant, const std::vector<hiqk_diag_op>& ops, const std::vector<int>& refs);
     void queue_diagonal(const Descriptor& d);
     void flush_pending(size_t keep = 0);
     void ensure_scratch();
     void release_device_resources();

     const double max_float_error_ = 1e-12;
     size_t min_local_, max_local_, max_global_, max_cluster_;
     int rank_, world_, device_;
     bool dry_run_, tracing_, timing_, batching_;
     std::vector<Index> locals_, globals_;
     FusionAccumulator fused_;
     std::mt19937 rnd_eng_;
     std::function<double()> rng_;

     Slab slab_;
     // Second buffer of the out-of-place passes (register permutations, operator accumulation): same reservation as the
     // slab, mapped on first use and kept for the engine's lifetime — Shor's algorithm calls emulate_math 2n times.
     Slab scratch_;
     double2* scratch_buffer(uint64_t amps, const char* who);
     Comm* comm_p_ = nullptr;  // process-wide communicator (Comm::shared), not owned
     cudaStream_t stream_ = nullptr;
     cudaStream_t comm_stream_ = nullptr;
     void* workspace_ = nullptr;  // reduction partials
     double* d_vals_ = nullptr;   // small device results
     double* d_blocks_ = nullptr; // world * 2^15 block sums
     void* swap_buf_ = nullptr;   // staging for the exchange (send | recv)
     size_t swap_buf_bytes_ = 0;
     cudaEvent_t swap_events_[4] = {nullptr, nullptr, nullptr, nullptr};  // packed[2], exchanged[2]
     // peer-mapped slabs (NVLink P2P): views of the other ranks' slabs and how many of my chunks each has been sent

     std::vector<PeerView> peer_views_;
     uint64_t epoch_ = 0;
     int swap_mode_ = 0;        // 0 auto, 1 staged NCCL only, 2 peer-mapped only, 3 packed only
     // packed exchange (low swapped slots): gather piece i straight into the PEERS' staging buffers (posted NVLink
     // writes of whole lines), stream-ordered barrier, scatter my own staging buffer into my slab
     bool packed_enabled_ = true;       // HIQ_SWAP_PACKED=0: low swapped slots take the in-place kernel too (A/B measurements)
     int packed_below_slot_ = 3;        // auto mode: used when the lowest swapped slot is below this
     uint64_t packed_piece_cap_ = 0;    // HIQ_SWAP_PACKED_PIECE: largest piece in amplitudes (tests drive the multi-piece pipeline with it)
     bool ensure_packed_staging(size_t want_bytes, size_t min_bytes);  // process-wide buffers live in Comm::packed()
     bool exchange_packed(const std::vector<int>& gpos, const std::vector<int>& slots);
     int min_p2p_slot_ = 0;     // lowest swapped slot for which the in-place kernel is used in auto mode
     int dense_variant_ = 0;

     EngineStats stats_;
     std::vector<Descriptor> trace_;
     std::vector<Descriptor> launches_;
     void trace_op(const Descriptor& d);  // a non-gate operation of the plan (swap, grow, fill, operator passes ...)
     struct LaunchStepRef {
          const Descriptor* gate;  // nullptr: diagonal factors only
          const std::vector<hiqk_diag_op>* ops;
     };
     void record_launch(int form, const std::vector<LaunchStepRef>& steps);
     struct TimedPass {
          int kind, k, variant, n_ref;
          cudaEvent_t start, stop;
     };
     cudaEvent_t swap_mark_[2] = {nullptr, nullptr};  // recorded right after the entry barrier of a timed exchange
     bool swap_marked_ = false;
     // Diagonal fused gates wait here until the next dense launch (which applies them to the tuples
     // it loads) or the next observation of the slab (one batched pass).  ref_passes[i] = how many
     // passes of the reference's plan op i stands for (ops over the same slots are multiplied on the host).
     std::vector<hiqk_diag_op> pending_;
     std::vector<int> pending_ref_;
     // Dense launches are held back until the next dense gate that cannot share their pass (or the next observation):
     //  * diagonal gates that arrive meanwhile and avoid the most recent held gate's targets commute with it and join its
     //    launch as per-tuple scalars; the ones that touch its targets wait in pending_ for the next dense gate;
     //  * consecutive dense gates whose targets (plus four low slots) fit one shared-memory tile form a RUN that goes out as
     //    one tile-resident launch (csrc/tile_program.cu): one HBM pass for the whole run.  The plan itself — clusters,
     //    swaps, their order — is the reference's; only the number of sweeps over the slab changes.
     struct HeldGate {
          Descriptor d;
          int variant = 0;
          bool direct = false;  // the one-gate DIRECT launch can carry its diagonals itself
          bool full = false;    // full 16 x 16 product (no block structure): FP64-bound
          std::vector<hiqk_diag_op> ops;  // diagonal factors applied before this gate
          std::vector<int> refs;          // passes of the reference's plan each of them stands for
     };
     std::vector<HeldGate> group_;
     bool tile_enabled_ = true;   // HIQ_TILE=0: one launch per dense gate (round-1 behaviour)
     int tile_max_steps_ = HIQK_TILE_MAX_STEPS;
     int tile_max_full_ = 1;      // full products per run (HIQ_TILE_MAX_FULL): two of them in one pass are FP64-bound
     bool tile_single_ = false;   // HIQ_TILE_SINGLE=1: single gates take the tile kernel too (A/B measurements)
     bool group_accepts(const HeldGate& cand, const std::vector<hiqk_diag_op>& cand_ops) const;
     void launch_group();
     std::vector<TimedPass> timed_;
     std::vector<cudaEvent_t> event_pool_;
     cudaEvent_t take_event();
     void d2h(void* dst, const void* src, size_t bytes);
};

}  // namespace hiq
