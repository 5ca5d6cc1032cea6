// Host-side stage / cluster scheduling (the `_sched_cpp` surface of the reference).
//
// Behavioural spec (must be reproduced bit-exactly):
//   SwapScheduler    reference: src/scheduler/swap_scheduler.{h,cpp}
//   ClusterScheduler reference: src/scheduler/cluster_scheduler.{h,cpp}
//   id <-> bit-position conversion: src/scheduler/convertors.cpp:45-73
// The algorithms are re-implemented, not transcribed: the cluster search replays the reference's
// deterministic first-visit order of candidate qubit sets without scoring, then scores the
// candidates with an early-exit walk (optionally on several host threads) and reduces
// lexicographically (more gates, fewer qubits, earlier visit) — the same winner, much sooner.
#pragma once
#include <cstdint>
#include <map>
#include <vector>

namespace hiq {
namespace sched {

using Id = int64_t;
using Mask = uint64_t;

struct Universe {
     std::vector<Id> pos_to_id;  // ascending ids
     std::map<Id, int> id_to_pos;
     void build(const std::vector<std::vector<Id>>& gate, const std::vector<std::vector<Id>>& gate_ctrl,
                const std::vector<Id>& extra_a, const std::vector<Id>& extra_b);
     Mask mask_of(const std::vector<Id>& ids) const;
     std::vector<Id> ids_of(Mask m) const;
};

class SwapScheduler {
public:
     SwapScheduler(const std::vector<std::vector<Id>>& gate, const std::vector<std::vector<Id>>& gate_ctrl,
                   std::vector<bool> gate_diag, int num_splits, int num_locals, bool fuse);
     // ids of the qubits that should be local during the next stage (ascending)
     std::vector<Id> ScheduleSwap();

private:
     bool can_take(int pos, Mask locals, Mask bad) const;
     void merge_into(int from, int to);
     bool merge_prev(int i);
     bool merge_next(int i);
     void fuse_single_qubit_gates();
     int search(int pos, Mask locals, Mask bad, int score, int splits);
     void tail(int pos, Mask locals, Mask bad, int score);

     int num_splits_, num_locals_;
     Universe u_;
     std::vector<Mask> gate_, ctrl_;
     std::vector<bool> diag_;
     std::vector<int> weight_;
     int best_score_ = 0;
     Mask best_locals_ = 0;
     // zero-budget tails
     std::vector<Mask> all_;
     std::vector<uint8_t> isdiag_;
     std::vector<int> suffix_w_;         // weight of gates pos..end
     std::vector<int> qubit_suffix_w_;   // [qubit][pos]: weight of gates pos..end touching the qubit
     std::vector<Mask> future_;          // qubits touched by gates pos..end
     int nq_ = 0, kmax_ = 1, floor_ = 0, min_qubits_ = 0;
};

class ClusterScheduler {
public:
     ClusterScheduler(const std::vector<std::vector<Id>>& gate, const std::vector<std::vector<Id>>& gate_ctrl,
                      std::vector<bool> gate_diag, const std::vector<Id>& locals, const std::vector<Id>& globals,
                      int cluster_size);
     // indices (program order) of the gates of the best next cluster; {} if nothing can run
     std::vector<int> ScheduleCluster();
     // number of candidate clusters scored by the last call (diagnostics)
     size_t candidates() const { return n_candidates_; }
     static void set_threads(int n);
     // 0 = bounded search (default), 1 = replay-and-score-everything (kept as the cross-check of the former)
     static void set_mode(int mode);
     // exact walks done by the last call (diagnostics)
     size_t evaluated() const { return n_evaluated_; }

private:
     std::vector<int> schedule_replay();
     std::vector<int> huge_gate() const;
     bool schedule_bounded(std::vector<int>& out);
     bool can_take(int i, Mask cluster, Mask bad) const;
     int score(Mask cluster) const;
     std::vector<int> gates_of(Mask cluster) const;
     void visit(Mask cluster, int b0);

     int cluster_size_;
     Universe u_;
     Mask locals_ = 0, globals_ = 0;
     std::vector<Mask> gate_, ctrl_, all_;
     std::vector<bool> diag_;
     bool early_exit_ok_ = true;
     struct Entry {
          Mask key = 0;
          uint8_t used = 0;
          uint8_t lo = 0;  // lowest bit position already expanded for this cluster
     };
     std::vector<Entry> table_;          // open-addressing memo: cluster -> expansion range
     size_t table_mask_ = 0;
     int table_shift_ = 0;
     int top_ = -1;                      // position of the highest local qubit
     std::vector<Mask> order_;           // candidates in first-visit order
     size_t n_candidates_ = 0;
     size_t n_evaluated_ = 0;
};

// The stage / cluster loop of the reference's GreedyScheduler (reference:
// hiq/projectq/cengines/_greedyscheduler.py:95-242) as ONE host object that emits the plan step by step:
// the initial relabelling, the controlled-Z role swaps, the gate clusters (= one fused-gate descriptor each once
// the engine has fused them) and the swap plans.  next() does the work of one step only, so the caller can hand
// every cluster to the device while the following one is being searched.
class GreedyPlanner {
public:
     enum Kind { DONE = 0, PERM = 1, ZSWAP = 2, CLUSTER = 3, SWAP = 4 };
     struct Step {
          int kind = DONE;
          // PERM: the whole permutation (locals then globals) for set_qubits_perm
          // ZSWAP: {gate index, control position}: that control and the target of the controlled-Z trade roles
          // CLUSTER: indices (into the constructor's gate list, program order) of the gates of the next cluster
          // SWAP: [global id, local id, ...] pairs for swap_qubits
          std::vector<Id> data;
     };
     GreedyPlanner(std::vector<std::vector<Id>> gate, std::vector<std::vector<Id>> gate_ctrl, std::vector<bool> is_z,
                   std::vector<Id> locals, std::vector<Id> globals, int cluster_size, int num_splits, bool first_stage);
     Step next();
     double cluster_seconds() const { return cluster_s_; }
     double swap_seconds() const { return swap_s_; }

private:
     void prepare_ctrlz();
     bool schedule_swap(std::vector<Id>& g_to_l, std::vector<Id>& l_to_g);
     void remaining(std::vector<std::vector<Id>>& gate, std::vector<std::vector<Id>>& ctrl) const;

     std::vector<std::vector<Id>> gate_, ctrl_;
     std::vector<bool> is_z_;
     std::vector<int> left_;  // indices of the gates not yet emitted, program order
     std::vector<Id> locals_, globals_;
     int cluster_size_, num_splits_;
     enum State { FIRST, STAGE_BEGIN, CLUSTERS, FINISHED } state_;
     std::vector<Step> queue_;  // steps decided but not yet handed out (front first)
     size_t queue_pos_ = 0;
     double cluster_s_ = 0.0, swap_s_ = 0.0;
};

}  // namespace sched
}  // namespace hiq
