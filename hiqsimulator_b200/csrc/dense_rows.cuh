// Row products shared by the dense-gate kernels (apply_dense.cu, tile_program.cu): out[b] = sum_c m[b][c] in[c] over one
// 2^K tuple held in registers, the matrix consumed straight from the kernel-parameter constant bank.
#pragma once
#include "hiq_device.cuh"

namespace hiq {

// out[b] = sum_c m[b][c] in[c]; every row is handed to `store(b, value)` as soon as it is done.
// KS < K: the matrix is block diagonal in its K - KS high index bits ("select" bits: the qubit multiplexes the
// gate, it is never mixed — fused controls and diagonal factors produce these, dense_shape() below finds them
// and moves them to the top).  Row b only meets the 2^KS columns that share its high bits; the skipped terms
// are exact zeros, so the result equals the full product and the pass drops from 8 * 2^K to 8 * 2^KS flops
// per amplitude — a QFT cluster (one or two Hadamards among controlled phases) turns from FP64-bound into HBM-bound.
// Full 16x16 product with three real multiplications per complex one (Re = Ax - By, Im = (A+B)(x+y) - Ax - By):
// 768 DFMA + 64 DADD per tuple instead of 1024 DFMA.  Under the sustained power cap the k = 4 pass is limited by
// FP64 issue, so the pass gets faster; the rounding differs from the four-multiplication form by a few ulp of
// the row norm (amplitudes agree with the reference far inside the 1e-12 tolerance of BASELINE.json).
template <class Store>
__device__ __forceinline__ void apply_rows_3m(const double2 (&in)[16], const double2* __restrict__ m, const double* __restrict__ msum,
                                              Store store)
{
     double s[16];
#pragma unroll
     for (int c = 0; c < 16; ++c) s[c] = in[c].x + in[c].y;
#pragma unroll
     for (int b = 0; b < 16; ++b) {
          double t1 = 0.0, t2 = 0.0, t3 = 0.0;
#pragma unroll
          for (int c = 0; c < 16; ++c) {
               t1 = fma(m[b * 16 + c].x, in[c].x, t1);
               t2 = fma(m[b * 16 + c].y, in[c].y, t2);
               t3 = fma(msum[b * 16 + c], s[c], t3);
          }
          store(b, make_double2(t1 - t2, t3 - t1 - t2));
     }
}

template <int K, int KS = K, class Store>
__device__ __forceinline__ void apply_rows(const double2 (&in)[1 << K], const double2* __restrict__ m,
                                           Store store)
{
     constexpr int D = 1 << K;
     constexpr int DS = 1 << KS;
     if constexpr (K <= 4) {
#pragma unroll
          for (int b = 0; b < D; ++b) {
               double2 acc = make_double2(0.0, 0.0);
#pragma unroll
               for (int c = (b & ~(DS - 1)); c < (b & ~(DS - 1)) + DS; ++c) cmac(acc, m[b * D + c], in[c]);
               store(b, acc);
          }
     }
     else {
          // 32x32: keep the row loop rolled (a full unroll is 64 KB of SASS)
          static_assert(K <= 4 || KS == K, "block form is instantiated for K <= 4 only");
#pragma unroll 2
          for (int b = 0; b < D; ++b) {
               double2 acc = make_double2(0.0, 0.0);
#pragma unroll
               for (int c = 0; c < D; ++c) cmac(acc, m[b * D + c], in[c]);
               store(b, acc);
          }
     }
}

}  // namespace hiq
