// Host-side gate-fusion accumulator: collects the gates of one cluster and multiplies them
// into ONE fused-gate descriptor (dense 2^N x 2^N, N <= 5, or diagonal, or a scalar).
//
// Behavioural spec = the reference's Fusion (reference: src/simulator-mpi/fusion_mpi.hpp:68-234;
// SURVEY.md Appendix D): 1x1 matrices fold into `factor`; a gate's controls that are not yet
// common to the cluster either join the common-control set (first gate) or become the highest
// matrix bits of that gate; common controls a new gate lacks are demoted into every earlier
// gate; the fused qubit list is sorted by qubit id; the fused matrix is item_n ... item_1 * factor.
// Matrices are flat row-major std::vector<complex<double>> here.
#pragma once
#include <algorithm>
#include <complex>
#include <cstdint>
#include <set>
#include <vector>

namespace hiq {

using cplx = std::complex<double>;
using Index = int64_t;

struct GateMatrix {
     int dim = 0;              // rows == cols
     std::vector<cplx> a;      // row-major
     GateMatrix() = default;
     explicit GateMatrix(int d) : dim(d), a(static_cast<size_t>(d) * d, cplx(0.0)) {}
     cplx& at(int r, int c) { return a[static_cast<size_t>(r) * dim + c]; }
     const cplx& at(int r, int c) const { return a[static_cast<size_t>(r) * dim + c]; }
};

// exact-zero test off the diagonal (reference: funcs.hpp:325-337)
inline bool is_diagonal(const GateMatrix& m)
{
     for (int i = 0; i < m.dim; ++i)
          for (int j = 0; j < m.dim; ++j)
               if (i != j && m.at(i, j) != cplx(0.0)) return false;
     return true;
}

class FusionAccumulator {
public:
     struct Item {
          GateMatrix m;
          bool diag;
          std::vector<Index> ids;  // matrix bit l <-> ids[l]
     };

     // identity on the first rows, `m` in the bottom-right block: controls = highest matrix bits
     static void add_controls(GateMatrix& m, std::vector<Index>& ids, const std::vector<Index>& ctrls)
     {
          if (ctrls.empty()) return;
          ids.insert(ids.end(), ctrls.begin(), ctrls.end());
          const int big = m.dim << ctrls.size();
          GateMatrix out(big);
          const int off = big - m.dim;
          for (int i = 0; i < off; ++i) out.at(i, i) = 1.0;
          for (int i = 0; i < m.dim; ++i)
               for (int j = 0; j < m.dim; ++j) out.at(off + i, off + j) = m.at(i, j);
          m = std::move(out);
     }

     void insert(GateMatrix m, bool diag, std::vector<Index> ids, const std::vector<Index>& ctrls)
     {
          if (m.dim == 1) {
               factor_ *= m.at(0, 0);
               return;
          }
          for (Index q: ids) qubits_.insert(q);
          absorb_controls(m, ids, ctrls);
          items_.push_back(Item{std::move(m), diag, std::move(ids)});
     }

     size_t num_qubits() const { return qubits_.size(); }
     bool empty() const { return items_.empty() && factor_ == cplx(1.0); }

     // -> fused matrix over `ids` (ascending qubit id), common controls, all-diagonal flag
     void fuse(GateMatrix& out, std::vector<Index>& ids, std::vector<Index>& ctrls, bool& diag)
     {
          ids.assign(qubits_.begin(), qubits_.end());
          const int N = static_cast<int>(ids.size());
          const int dim = 1 << N;
          out = GateMatrix(dim);
          for (int i = 0; i < dim; ++i) out.at(i, i) = factor_;
          diag = true;
          std::vector<cplx> col(dim);
          for (const Item& it: items_) {
               if (!it.diag) diag = false;
               const int n = static_cast<int>(it.ids.size());
               std::vector<int> bit(n);  // position of the item's l-th qubit in the fused index
               for (int l = 0; l < n; ++l)
                    bit[l] = static_cast<int>(std::lower_bound(ids.begin(), ids.end(), it.ids[l]) - ids.begin());
               for (int k = 0; k < dim; ++k) {
                    for (int i = 0; i < dim; ++i) col[i] = out.at(i, k);
                    for (int i = 0; i < dim; ++i) {
                         int row = 0, cleared = i;
                         for (int l = 0; l < n; ++l) {
                              row |= ((i >> bit[l]) & 1) << l;
                              cleared &= ~(1 << bit[l]);
                         }
                         cplx acc = 0.0;
                         for (int j = 0; j < (1 << n); ++j) {
                              int src = cleared;
                              for (int l = 0; l < n; ++l) src |= ((j >> l) & 1) << bit[l];
                              acc += col[src] * it.m.at(row, j);
                         }
                         out.at(i, k) = acc;
                    }
               }
          }
          ctrls.assign(common_ctrls_.begin(), common_ctrls_.end());
          factor_ = 1.0;
     }

     void reset() { *this = FusionAccumulator(); }

private:
     void absorb_controls(GateMatrix& m, std::vector<Index>& ids, const std::vector<Index>& ctrls)
     {
          std::set<Index> missing = common_ctrls_;  // common controls the new gate does not have
          for (Index c: ctrls) {
               if (common_ctrls_.count(c) == 0) {
                    if (!items_.empty()) {
                         add_controls(m, ids, {c});
                         qubits_.insert(c);
                    }
                    else {
                         common_ctrls_.insert(c);
                    }
               }
               else {
                    missing.erase(c);
               }
          }
          if (!missing.empty()) {
               std::vector<Index> demoted(missing.begin(), missing.end());
               for (Index c: demoted) {
                    common_ctrls_.erase(c);
                    qubits_.insert(c);
               }
               for (Item& it: items_) add_controls(it.m, it.ids, demoted);
          }
     }

     std::set<Index> qubits_;
     std::vector<Item> items_;
     std::set<Index> common_ctrls_;
     cplx factor_ = 1.0;
};

}  // namespace hiq
