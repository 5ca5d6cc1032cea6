#include "sched.hpp"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <stdexcept>
#include <thread>
#include <unordered_set>

namespace hiq {
namespace sched {

static inline int popcount(Mask m) { return __builtin_popcountll(m); }
static inline bool subset(Mask a, Mask b) { return (a & ~b) == 0; }

// ------------------------------------------------------------------------------------ Universe
void Universe::build(const std::vector<std::vector<Id>>& gate, const std::vector<std::vector<Id>>& gate_ctrl,
                     const std::vector<Id>& extra_a, const std::vector<Id>& extra_b)
{
     // position = rank of the id among all distinct ids seen (ascending id == ascending position)
     pos_to_id.clear();
     for (auto& g: gate) pos_to_id.insert(pos_to_id.end(), g.begin(), g.end());
     for (auto& g: gate_ctrl) pos_to_id.insert(pos_to_id.end(), g.begin(), g.end());
     pos_to_id.insert(pos_to_id.end(), extra_a.begin(), extra_a.end());
     pos_to_id.insert(pos_to_id.end(), extra_b.begin(), extra_b.end());
     std::sort(pos_to_id.begin(), pos_to_id.end());
     pos_to_id.erase(std::unique(pos_to_id.begin(), pos_to_id.end()), pos_to_id.end());
     if (pos_to_id.size() >= 64) throw std::runtime_error("CalcPos(): scheduling with 64 or more qubits is not yet supported");
     id_to_pos.clear();
     for (size_t i = 0; i < pos_to_id.size(); ++i) id_to_pos[pos_to_id[i]] = static_cast<int>(i);
}

Mask Universe::mask_of(const std::vector<Id>& ids) const
{
     Mask m = 0;
     for (Id q: ids) m |= Mask(1) << id_to_pos.at(q);
     return m;
}

std::vector<Id> Universe::ids_of(Mask m) const
{
     std::vector<Id> out;
     for (int p = 0; p < 64 && (Mask(1) << p) <= m; ++p)
          if ((m >> p) & 1) out.push_back(pos_to_id[p]);
     return out;
}

// ------------------------------------------------------------------------------------ SwapScheduler
SwapScheduler::SwapScheduler(const std::vector<std::vector<Id>>& gate, const std::vector<std::vector<Id>>& gate_ctrl,
                             std::vector<bool> gate_diag, int num_splits, int num_locals, bool fuse)
    : num_splits_(num_splits), num_locals_(num_locals), diag_(std::move(gate_diag)), weight_(gate.size(), 1)
{
     if (gate.size() != gate_ctrl.size() || gate.size() != diag_.size()) throw std::runtime_error("SwapScheduler: ctor(): size mismatch");
     u_.build(gate, gate_ctrl, {}, {});
     gate_.resize(gate.size());
     ctrl_.resize(gate.size());
     for (size_t i = 0; i < gate.size(); ++i) {
          gate_[i] = u_.mask_of(gate[i]);
          ctrl_[i] = u_.mask_of(gate_ctrl[i]);
     }
     if (fuse) fuse_single_qubit_gates();
}

bool SwapScheduler::can_take(int pos, Mask locals, Mask bad) const
{
     if ((ctrl_[pos] | gate_[pos]) & bad) return false;
     if (!diag_[pos]) return popcount(gate_[pos] | locals) <= num_locals_;
     return true;
}

void SwapScheduler::merge_into(int from, int to)
{
     weight_[to] += weight_[from];
     weight_[from] = 0;
     if (!diag_[from]) diag_[to] = false;
     if (ctrl_[to] & gate_[from]) {
          ctrl_[to] ^= gate_[from];  // the qubit stops being a control of the neighbour ...
          gate_[to] |= gate_[from];  // ... and becomes one of its targets
     }
}

bool SwapScheduler::merge_prev(int i)
{
     for (int j = i - 1; j >= 0; --j)
          if (gate_[i] & (gate_[j] | ctrl_[j])) {
               merge_into(i, j);
               return true;
          }
     return false;
}

bool SwapScheduler::merge_next(int i)
{
     for (int j = i + 1; j < static_cast<int>(gate_.size()); ++j)
          if (gate_[i] & (gate_[j] | ctrl_[j])) {
               merge_into(i, j);
               return true;
          }
     return false;
}

void SwapScheduler::fuse_single_qubit_gates()
{
     // last gate to first: a gate touching exactly one qubit is absorbed by a neighbour on that qubit
     for (int i = static_cast<int>(gate_.size()) - 1; i >= 0; --i) {
          bool absorbed = false;
          if (popcount(gate_[i] | ctrl_[i]) == 1) {
               if (!diag_[i]) absorbed = merge_next(i) || merge_prev(i);
               else absorbed = merge_prev(i) || merge_next(i);
          }
          if (absorbed) {
               gate_.erase(gate_.begin() + i);
               ctrl_.erase(ctrl_.begin() + i);
               diag_.erase(diag_.begin() + i);
               weight_.erase(weight_.begin() + i);
          }
     }
}

std::vector<Id> SwapScheduler::ScheduleSwap()
{
     if (gate_.empty()) return {};
     best_score_ = 0;
     best_locals_ = 0;
     // tables for the zero-budget tails (see tail())
     const int n = static_cast<int>(gate_.size());
     all_.resize(n);
     isdiag_.resize(n);
     suffix_w_.assign(n + 1, 0);
     kmax_ = 1;
     min_qubits_ = 64;
     future_.assign(n + 1, 0);
     Mask used = 0;
     for (int i = n - 1; i >= 0; --i) {
          all_[i] = gate_[i] | ctrl_[i];
          future_[i] = future_[i + 1] | all_[i];
          min_qubits_ = std::min(min_qubits_, popcount(all_[i]));
          isdiag_[i] = diag_[i] ? 1 : 0;
          suffix_w_[i] = suffix_w_[i + 1] + weight_[i];
          kmax_ = std::max(kmax_, popcount(all_[i]));
          used |= all_[i];
     }
     nq_ = used ? 64 - __builtin_clzll(used) : 0;
     qubit_suffix_w_.assign(static_cast<size_t>(nq_) * (n + 1), 0);
     for (int q = 0; q < nq_; ++q) {
          int* row = &qubit_suffix_w_[static_cast<size_t>(q) * (n + 1)];
          for (int i = n - 1; i >= 0; --i) row[i] = row[i + 1] + (((all_[i] >> q) & 1) ? weight_[i] : 0);
     }
     // the path that takes every gate it can is part of the walk for any budget: its score is a floor of the result
     {
          Mask locals = 0, bad = 0;
          int score = 0;
          for (int i = 0; i < n; ++i) {
               if (can_take(i, locals, bad)) {
                    score += weight_[i];
                    if (!isdiag_[i]) locals |= gate_[i];
               }
               else bad |= all_[i];
          }
          floor_ = score;
     }
     search(0, 0, 0, 0, num_splits_);
     return u_.ids_of(best_locals_);
}

// A node entered without budget never branches again (a takeable gate is taken, anything else is skipped), hands
// back no budget, and only its end matters: scores grow along the path and the local set stops changing after
// the last take.  The result of ScheduleSwap() is the FIRST node of the walk with the highest score, so a tail
// that cannot reach floor_ (<= that highest score) is irrelevant and is not walked.  Bound: a remaining gate on
// a blocked qubit is lost; every lost gate is counted at most kmax_ times in the per-qubit suffix weights.
void SwapScheduler::tail(int pos, Mask locals, Mask bad, int score)
{
     const int n = static_cast<int>(gate_.size());
     {
          long lost = 0;
          const size_t stride = static_cast<size_t>(n) + 1;
          for (Mask m = bad; m; m &= m - 1) {
               const int q = __builtin_ctzll(m);
               if (q < nq_) lost += qubit_suffix_w_[q * stride + pos];
          }
          const long reach = score + suffix_w_[pos] - (lost + kmax_ - 1) / kmax_;
          if (reach < floor_) return;
     }
     for (; pos < n; ++pos) {
          const Mask a = all_[pos];
          bool take = (a & bad) == 0;
          if (take && !isdiag_[pos]) take = popcount(gate_[pos] | locals) <= num_locals_;
          if (take) {
               score += weight_[pos];
               if (!isdiag_[pos]) locals |= gate_[pos];
          }
          else bad |= a;
     }
     if (score > best_score_) {
          best_score_ = score;
          best_locals_ = locals;
     }
}

// Budgeted backtracking; returns the unused split budget.  Branch order: skip first, then take.
// Only a gate that can be taken AND needs new local qubits branches (one unit of budget; the skip side gets half
// of what is left and hands back what it did not use).  Everything else — a gate that cannot be taken, a diagonal
// gate, a gate whose targets are already local — has a single continuation that receives and returns the whole
// budget, so those nodes are walked in a loop and only branches recurse; the take side of a branch continues in
// the same frame.  Visit order, budget flow and best-so-far updates are those of the node-per-gate recursion
// (reference: swap_scheduler.cpp:86-170); the reference walk is ~99 % such single-continuation nodes.
int SwapScheduler::search(int pos, Mask locals, Mask bad, int score, int splits)
{
     const int n = static_cast<int>(gate_.size());
     for (;;) {
          if (splits == 0) {
               tail(pos, locals, bad, score);
               return 0;
          }
          if (score > best_score_) {
               best_score_ = score;
               best_locals_ = locals;
          }
          if (pos == n) return splits;
          const Mask a = all_[pos];
          const bool is_diag = isdiag_[pos] != 0;
          const bool takeable = (a & bad) == 0 && (is_diag || popcount(gate_[pos] | locals) <= num_locals_);
          if (!takeable) {
               bad |= a;
               ++pos;
               // no remaining gate has all its qubits unblocked: the rest of the path neither scores nor branches
               if (popcount(future_[pos] & ~bad) < min_qubits_) return splits;
               continue;
          }
          if (is_diag || subset(gate_[pos], locals)) {
               score += weight_[pos];
               ++pos;
               continue;
          }
          splits -= 1;
          const int give = splits / 2;
          splits += search(pos + 1, locals, bad | a, score, give) - give;
          locals |= gate_[pos];
          score += weight_[pos];
          ++pos;
     }
}

// ------------------------------------------------------------------------------------ ClusterScheduler
static std::atomic<int> g_threads{0};
void ClusterScheduler::set_threads(int n) { g_threads.store(n); }

ClusterScheduler::ClusterScheduler(const std::vector<std::vector<Id>>& gate, const std::vector<std::vector<Id>>& gate_ctrl,
                                   std::vector<bool> gate_diag, const std::vector<Id>& locals, const std::vector<Id>& globals,
                                   int cluster_size)
    : cluster_size_(cluster_size), diag_(std::move(gate_diag))
{
     if (gate.size() != gate_ctrl.size() || gate.size() != diag_.size()) throw std::runtime_error("ClusterScheduler: ctor(): size mismatch");
     u_.build(gate, gate_ctrl, locals, globals);
     gate_.resize(gate.size());
     ctrl_.resize(gate.size());
     all_.resize(gate.size());
     locals_ = u_.mask_of(locals);
     globals_ = u_.mask_of(globals);
     for (size_t i = 0; i < gate.size(); ++i) {
          gate_[i] = u_.mask_of(gate[i]);
          ctrl_[i] = u_.mask_of(gate_ctrl[i]);
          all_[i] = gate_[i] | ctrl_[i];
          // A gate without any local qubit can only be taken if it is diagonal (or has no targets);
          // if none exists, a walk may stop as soon as every cluster qubit is blocked.
          if ((all_[i] & locals_) == 0 && (diag_[i] || gate_[i] == 0)) early_exit_ok_ = false;
     }
}

bool ClusterScheduler::can_take(int i, Mask cluster, Mask bad) const
{
     if ((all_[i] & bad) || !subset(all_[i] & locals_, cluster)) return false;
     if (!diag_[i]) return subset(gate_[i], locals_);
     return true;
}

// gates admitted by `cluster` in program order; a gate that is not admitted blocks its qubits
int ClusterScheduler::score(Mask cluster) const
{
     Mask bad = 0;
     int taken = 0;
     const int n = static_cast<int>(gate_.size());
     for (int i = 0; i < n; ++i) {
          if (can_take(i, cluster, bad)) ++taken;
          else {
               bad |= all_[i];
               if (early_exit_ok_ && subset(cluster, bad)) break;
          }
     }
     return taken;
}

std::vector<int> ClusterScheduler::gates_of(Mask cluster) const
{
     Mask bad = 0;
     std::vector<int> out;
     for (int i = 0; i < static_cast<int>(gate_.size()); ++i) {
          if (can_take(i, cluster, bad)) out.push_back(i);
          else bad |= all_[i];
     }
     return out;
}

// Replays the reference's enumeration and records each distinct cluster the first time it is met.
//
// Reference walk (cluster_scheduler.cpp:80-99): Rec(c, bit) scores c, then walks bit upwards; at
// every position it stops if (c, bit) was expanded before, marks it, recurses into (c | bit,
// bit << 1) when the bit is an unused local and the cluster is not full, and goes on with bit << 1
// until bit exceeds the highest local.  Because every walk runs to the top or into an already
// marked position, the marked positions of a cluster always form a range [lo, top]; remembering
// `lo` per cluster is therefore equivalent to the reference's per-(cluster, bit) memo, and the
// walk only needs to touch the addable bits in [b0, lo).
void ClusterScheduler::visit(Mask cluster, int b0)
{
     size_t slot = (cluster * 0x9E3779B97F4A7C15ull) >> table_shift_;
     while (table_[slot].used && table_[slot].key != cluster) slot = (slot + 1) & table_mask_;
     Entry& e = table_[slot];
     if (!e.used) {
          e.used = 1;
          e.key = cluster;
          e.lo = static_cast<uint8_t>(top_ + 1);
          order_.push_back(cluster);
     }
     if (b0 >= e.lo) return;
     const int hi = e.lo;
     e.lo = static_cast<uint8_t>(b0);
     if (popcount(cluster) >= cluster_size_) return;
     Mask range = (hi >= 64 ? ~Mask(0) : ((Mask(1) << hi) - 1)) & ~((Mask(1) << b0) - 1);
     Mask addable = locals_ & ~cluster & range;
     while (addable) {
          const int b = __builtin_ctzll(addable);
          addable &= addable - 1;
          visit(cluster | (Mask(1) << b), b + 1);
     }
}

static std::atomic<int> g_mode{0};
void ClusterScheduler::set_mode(int mode) { g_mode.store(mode); }

std::vector<int> ClusterScheduler::ScheduleCluster()
{
     if (gate_.empty()) return {};
     n_evaluated_ = 0;
     if (g_mode.load() == 0) {
          std::vector<int> out;
          if (schedule_bounded(out)) return out;
     }
     return schedule_replay();
}

// Bounded search for the same winner as schedule_replay() without replaying the enumeration.
//
// Facts used (C, D clusters; a gate's "seed" = its local qubits):
//  * score is monotone (C subset of D => score(C) <= score(D)) and score(C) = score(used(C)), used(C) = union of the
//    seeds of the gates C admits; so the winner (more gates, then fewer qubits) satisfies used(W) = W: every
//    qubit of W has its FIRST gate admitted.  Qubits whose first gate can never run are left out.
//  * chain bound: an admitted gate is preceded, on each of its local qubits, only by admitted gates, so
//    score(C) <= free + sum over q in C of sum over the leading gates of q's chain with seed inside C of 1/|seed|.
//    Only clusters whose bound reaches the best score so far are walked exactly.
//  * first-visit order of the reference's enumeration, needed for ties: a cluster X is first met in the
//    expansion of the first gate (program order) whose seed lies inside X, and inside one expansion clusters
//    are met in lexicographic order of their added qubits (ascending, prefix first).
bool ClusterScheduler::schedule_bounded(std::vector<int>& out)
{
     constexpr int W = 60;  // divisible by every seed size up to 6
     if (cluster_size_ > 6 || cluster_size_ < 0) return false;
     const int n = static_cast<int>(gate_.size());
     struct Link {
          Mask seed;
          int w;
     };
     std::vector<Link> chain[64];
     std::vector<Mask> seed(n);
     int free_takes = 0;
     bool empty_seed = false;
     for (int i = 0; i < n; ++i) {
          seed[i] = all_[i] & locals_;
          const bool takeable = diag_[i] || subset(gate_[i], locals_);
          const int pc = popcount(seed[i]);
          if (pc == 0) {
               empty_seed = true;
               if (takeable) ++free_takes;
               continue;
          }
          const bool ok = takeable && pc <= cluster_size_;
          for (Mask m = seed[i]; m; m &= m - 1) chain[__builtin_ctzll(m)].push_back(Link{ok ? seed[i] : ~Mask(0), ok ? W / pc : 0});
     }
     int act[64];
     int na = 0;
     for (int p = 0; p < 64; ++p)
          if (!chain[p].empty() && chain[p][0].w) act[na++] = p;

     auto bound = [&](Mask c) {
          int u = free_takes * W;
          for (Mask m = c; m; m &= m - 1) {
               const std::vector<Link>& ch = chain[__builtin_ctzll(m)];
               for (const Link& l: ch) {
                    if (!subset(l.seed, c)) break;
                    u += l.w;
               }
          }
          return u;
     };

     std::vector<std::pair<int, Mask>> cand;
     if (empty_seed) cand.emplace_back(free_takes * W, Mask(0));
     {
          // subsets of the active qubits with 1..cluster_size members
          int idx[8];
          Mask acc[8];
          int depth = 0;
          idx[0] = 0;
          acc[0] = 0;
          while (depth >= 0) {
               if (depth >= cluster_size_ || idx[depth] >= na) {
                    --depth;
                    continue;
               }
               const Mask c = acc[depth] | (Mask(1) << act[idx[depth]]);
               ++idx[depth];
               cand.emplace_back(bound(c), c);
               if (depth + 1 < cluster_size_) {
                    idx[depth + 1] = idx[depth];
                    acc[depth + 1] = c;
                    ++depth;
               }
          }
     }
     n_candidates_ = cand.size();

     int best_score = 0, best_bits = 0;
     std::vector<Mask> ties;
     auto consider = [&](Mask c) {
          ++n_evaluated_;
          const int s = score(c);
          if (s == 0) return;
          const int b = popcount(c);
          if (s > best_score || (s == best_score && b < best_bits)) {
               best_score = s;
               best_bits = b;
               ties.clear();
               ties.push_back(c);
          }
          else if (s == best_score && b == best_bits) ties.push_back(c);
     };
     // start from the most promising candidate so that the bound bites at once
     size_t top = 0;
     for (size_t k = 1; k < cand.size(); ++k)
          if (cand[k].first > cand[top].first) top = k;
     if (!cand.empty()) consider(cand[top].second);
     for (size_t k = 0; k < cand.size(); ++k)
          if (k != top && cand[k].first >= best_score * W && cand[k].first > 0) consider(cand[k].second);

     if (ties.empty()) {
          out = huge_gate();
          return true;
     }
     auto first_seed_gate = [&](Mask c) {
          for (int i = 0; i < n; ++i)
               if (subset(seed[i], c) && popcount(seed[i]) <= cluster_size_) return i;
          return n;
     };
     Mask win = ties[0];
     int win_gate = ties.size() > 1 ? first_seed_gate(win) : 0;
     for (size_t k = 1; k < ties.size(); ++k) {
          const Mask c = ties[k];
          const int g = first_seed_gate(c);
          bool earlier;
          if (g != win_gate) earlier = g < win_gate;
          else {
               const Mask diff = c ^ win;  // same seed, same size: the one holding the lowest differing qubit comes first
               earlier = (c & (diff & (~diff + 1))) != 0;
          }
          if (earlier) {
               win = c;
               win_gate = g;
          }
     }
     out = gates_of(win);
     return true;
}

std::vector<int> ClusterScheduler::schedule_replay()
{
     order_.clear();
     top_ = locals_ ? 63 - __builtin_clzll(locals_) : -1;
     {
          // capacity: every subset of the locals with <= cluster_size qubits, plus the seeds
          const int nl = popcount(locals_);
          double bound = static_cast<double>(gate_.size()) + 1;
          double binom = 1;
          for (int s = 0; s <= std::min(cluster_size_, nl); ++s) {
               bound += binom;
               binom = binom * (nl - s) / (s + 1);
          }
          size_t cap = 1024;
          int shift = 54;
          while (static_cast<double>(cap) < 2.0 * bound && shift > 34) {
               cap <<= 1;
               --shift;
          }
          table_.assign(cap, Entry{});
          table_mask_ = cap - 1;
          table_shift_ = shift;
     }
     for (size_t i = 0; i < gate_.size(); ++i) {
          const Mask seed = locals_ & all_[i];
          if (popcount(seed) <= cluster_size_) visit(seed, 0);
     }
     n_candidates_ = order_.size();

     // winner = more gates, then fewer qubits, then earlier visit; score 0 never displaces "nothing"
     struct Best {
          int score = 0;
          int bits = 0;
          size_t index = 0;
          bool better_than(const Best& o) const
          {
               if (score != o.score) return score > o.score;
               if (bits != o.bits) return bits < o.bits;
               return index < o.index;
          }
     };
     auto scan = [&](size_t lo, size_t hi) {
          Best b;
          b.index = static_cast<size_t>(-1);
          for (size_t k = lo; k < hi; ++k) {
               Best c;
               c.score = score(order_[k]);
               if (c.score == 0) continue;
               c.bits = popcount(order_[k]);
               c.index = k;
               if (b.index == static_cast<size_t>(-1) || c.better_than(b)) b = c;
          }
          return b;
     };
     int threads = g_threads.load();
     if (threads <= 0) threads = static_cast<int>(std::min(16u, std::max(1u, std::thread::hardware_concurrency())));
     if (order_.size() * gate_.size() < (1u << 18)) threads = 1;  // not worth the thread start-up
     Best best;
     best.index = static_cast<size_t>(-1);
     if (threads == 1) {
          best = scan(0, order_.size());
     }
     else {
          std::vector<Best> part(threads);
          std::vector<std::thread> pool;
          const size_t per = (order_.size() + threads - 1) / threads;
          for (int t = 0; t < threads; ++t)
               pool.emplace_back([&, t] { part[t] = scan(std::min(order_.size(), t * per), std::min(order_.size(), (t + 1) * per)); });
          for (auto& th: pool) th.join();
          for (auto& p: part)
               if (p.index != static_cast<size_t>(-1) && (best.index == static_cast<size_t>(-1) || p.better_than(best))) best = p;
     }
     if (best.index != static_cast<size_t>(-1)) return gates_of(order_[best.index]);
     return huge_gate();
}

// nothing fits a cluster: the first gate that can run when every local qubit is allowed ("huge gate")
std::vector<int> ClusterScheduler::huge_gate() const
{
     Mask bad = 0;
     for (int i = 0; i < static_cast<int>(gate_.size()); ++i) {
          if (can_take(i, locals_, bad)) return {i};
          bad |= all_[i];
     }
     return {};
}

// ------------------------------------------------------------------------------------ GreedyPlanner
namespace {
double seconds_since(std::chrono::steady_clock::time_point t0)
{
     return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
}
bool contains(const std::vector<Id>& v, Id x) { return std::find(v.begin(), v.end(), x) != v.end(); }
}  // namespace

GreedyPlanner::GreedyPlanner(std::vector<std::vector<Id>> gate, std::vector<std::vector<Id>> gate_ctrl, std::vector<bool> is_z,
                             std::vector<Id> locals, std::vector<Id> globals, int cluster_size, int num_splits, bool first_stage)
    : gate_(std::move(gate)), ctrl_(std::move(gate_ctrl)), is_z_(std::move(is_z)), locals_(std::move(locals)),
      globals_(std::move(globals)), cluster_size_(cluster_size), num_splits_(num_splits), state_(first_stage ? FIRST : STAGE_BEGIN)
{
     if (gate_.size() != ctrl_.size() || gate_.size() != is_z_.size()) throw std::runtime_error("GreedyPlanner: ctor(): size mismatch");
     left_.resize(gate_.size());
     for (size_t i = 0; i < left_.size(); ++i) left_[i] = static_cast<int>(i);
     if (left_.empty()) state_ = FINISHED;
}

void GreedyPlanner::remaining(std::vector<std::vector<Id>>& gate, std::vector<std::vector<Id>>& ctrl) const
{
     gate.clear();
     ctrl.clear();
     gate.reserve(left_.size());
     ctrl.reserve(left_.size());
     for (int i: left_) {
          gate.push_back(gate_[i]);
          ctrl.push_back(ctrl_[i]);
     }
}

// reference: _greedyscheduler.py:95-112 — a controlled-Z whose target is global hands the target role to its first local control
void GreedyPlanner::prepare_ctrlz()
{
     for (int i: left_) {
          if (!is_z_[i]) continue;
          if (gate_[i].size() != 1) throw std::runtime_error("GreedyPlanner: a controlled-Z has exactly one target");
          if (!contains(globals_, gate_[i][0])) continue;
          for (size_t c = 0; c < ctrl_[i].size(); ++c)
               if (contains(locals_, ctrl_[i][c])) {
                    std::swap(ctrl_[i][c], gate_[i][0]);
                    Step s;
                    s.kind = ZSWAP;
                    s.data = {static_cast<Id>(i), static_cast<Id>(c)};
                    queue_.push_back(std::move(s));
                    break;
               }
     }
}

// reference: _greedyscheduler.py:175-193
bool GreedyPlanner::schedule_swap(std::vector<Id>& g_to_l, std::vector<Id>& l_to_g)
{
     const auto t0 = std::chrono::steady_clock::now();
     std::vector<std::vector<Id>> gate, ctrl;
     remaining(gate, ctrl);
     const std::vector<bool> diag(gate.size(), false);
     const int nl = static_cast<int>(locals_.size());
     std::vector<Id> want = SwapScheduler(gate, ctrl, diag, num_splits_, nl, true).ScheduleSwap();
     if (want.empty()) want = SwapScheduler(gate, ctrl, diag, num_splits_, nl, false).ScheduleSwap();
     swap_s_ += seconds_since(t0);
     g_to_l.clear();
     l_to_g.clear();
     for (Id q: want)  // ascending already; keep distinct ids that are not local
          if (!contains(locals_, q) && !contains(g_to_l, q)) g_to_l.push_back(q);
     std::sort(g_to_l.begin(), g_to_l.end());
     if (!g_to_l.empty()) {
          std::vector<Id> out;
          for (Id q: locals_)
               if (!contains(want, q) && !contains(out, q)) out.push_back(q);
          std::sort(out.begin(), out.end());
          if (out.size() < g_to_l.size()) throw std::runtime_error("GreedyPlanner: not enough local qubits to evict");
          l_to_g.assign(out.begin(), out.begin() + g_to_l.size());
     }
     return !g_to_l.empty();
}

GreedyPlanner::Step GreedyPlanner::next()
{
     for (;;) {
          if (queue_pos_ < queue_.size()) return std::move(queue_[queue_pos_++]);
          queue_.clear();
          queue_pos_ = 0;
          switch (state_) {
               case FINISHED: return Step{};
               case FIRST: {
                    // first scheduling of this engine: choose the local set by relabelling, no data motion
                    // (reference: _greedyscheduler.py:211-222)
                    std::vector<Id> g_to_l, l_to_g;
                    schedule_swap(g_to_l, l_to_g);
                    std::vector<Id> perm = locals_;
                    perm.insert(perm.end(), globals_.begin(), globals_.end());
                    for (size_t i = 0; i < l_to_g.size(); ++i) {
                         auto a = std::find(perm.begin(), perm.end(), g_to_l[i]);
                         auto b = std::find(perm.begin(), perm.end(), l_to_g[i]);
                         if (a == perm.end() || b == perm.end()) throw std::runtime_error("GreedyPlanner: unknown qubit id in the swap choice");
                         std::iter_swap(a, b);
                    }
                    locals_.assign(perm.begin(), perm.begin() + locals_.size());
                    globals_.assign(perm.end() - globals_.size(), perm.end());
                    Step s;
                    s.kind = PERM;
                    s.data = std::move(perm);
                    queue_.push_back(std::move(s));
                    state_ = STAGE_BEGIN;
                    break;
               }
               case STAGE_BEGIN:
                    prepare_ctrlz();
                    state_ = CLUSTERS;
                    break;
               case CLUSTERS: {
                    const auto t0 = std::chrono::steady_clock::now();
                    std::vector<std::vector<Id>> gate, ctrl;
                    remaining(gate, ctrl);
                    const std::vector<int> avail =
                        ClusterScheduler(gate, ctrl, std::vector<bool>(gate.size(), false), locals_, globals_, cluster_size_).ScheduleCluster();
                    cluster_s_ += seconds_since(t0);
                    if (!avail.empty()) {
                         Step s;
                         s.kind = CLUSTER;
                         std::vector<char> gone(left_.size(), 0);
                         for (int a: avail) {
                              s.data.push_back(left_[a]);
                              gone[a] = 1;
                         }
                         std::vector<int> keep;
                         keep.reserve(left_.size() - avail.size());
                         for (size_t i = 0; i < left_.size(); ++i)
                              if (!gone[i]) keep.push_back(left_[i]);
                         left_.swap(keep);
                         queue_.push_back(std::move(s));
                         break;
                    }
                    if (left_.empty()) {
                         state_ = FINISHED;
                         break;
                    }
                    // nothing more runs with these locals: next stage (reference: _greedyscheduler.py:224-241)
                    std::vector<Id> g_to_l, l_to_g;
                    if (!schedule_swap(g_to_l, l_to_g)) throw std::runtime_error("GreedyPlanner: the swap scheduler found nothing to bring in");
                    Step s;
                    s.kind = SWAP;
                    for (size_t i = 0; i < g_to_l.size(); ++i) {
                         s.data.push_back(g_to_l[i]);
                         s.data.push_back(l_to_g[i]);
                         auto g = std::find(globals_.begin(), globals_.end(), g_to_l[i]);
                         auto l = std::find(locals_.begin(), locals_.end(), l_to_g[i]);
                         if (g == globals_.end() || l == locals_.end()) throw std::runtime_error("GreedyPlanner: swap choice outside the maps");
                         std::iter_swap(g, l);
                    }
                    queue_.push_back(std::move(s));
                    state_ = STAGE_BEGIN;
                    break;
               }
          }
     }
}

}  // namespace sched
}  // namespace hiq
