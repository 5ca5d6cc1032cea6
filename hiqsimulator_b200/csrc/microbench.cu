// Roofline denominators measured with this library's own kernels, reported by bench.py next to
// the gate kernels: HBM copy GB/s (read+write bytes), FP64 FMA TFLOP/s (DFMA pipe) and FP64
// tensor TFLOP/s (mma.sync.m8n8k4.f64).  The reference has socket-test.cpp / opyt.cpp for the same
// purpose on the CPU (reference: socket-test.cpp:47-66, opyt.cpp:148-212).
#include <algorithm>

#include "hiq_device.cuh"
#include "hiq_host.hpp"

namespace hiq {

__global__ void __launch_bounds__(256) mb_copy_kernel(const double2* __restrict__ src, double2* __restrict__ dst, uint64_t n)
{
     const uint64_t stride = static_cast<uint64_t>(gridDim.x) * 256;
     uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
     for (; i + 3 * stride < n; i += 4 * stride) {
          double2 v[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) v[u] = ldg_stream(src + i + u * stride);
#pragma unroll
          for (int u = 0; u < 4; ++u) dst[i + u * stride] = v[u];
     }
     for (; i < n; i += stride) dst[i] = src[i];
}

__global__ void __launch_bounds__(256) mb_dfma_kernel(double* out, int iters, double x, double y)
{
     double a[8];
#pragma unroll
     for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 1e-9 + i;
     for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int i = 0; i < 8; ++i) a[i] = fma(a[i], x, y);
     }
     double s = 0.0;
#pragma unroll
     for (int i = 0; i < 8; ++i) s += a[i];
     if (s == 12345.678) out[0] = s;  // never true: keeps the chains alive
}

__global__ void __launch_bounds__(256) mb_dmma_kernel(double* out, int iters, double x, double y)
{
     double d[8][2];
#pragma unroll
     for (int i = 0; i < 8; ++i) d[i][0] = d[i][1] = threadIdx.x * 1e-9 + i;
     for (int it = 0; it < iters; ++it) {
#pragma unroll
          for (int i = 0; i < 8; ++i)
               asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                            : "+d"(d[i][0]), "+d"(d[i][1])
                            : "d"(x), "d"(y));
     }
     double s = 0.0;
#pragma unroll
     for (int i = 0; i < 8; ++i) s += d[i][0] + d[i][1];
     if (s == 12345.678) out[0] = s;
}

}  // namespace hiq

using namespace hiq;

extern "C" int hiqk_microbench(int what, int iters, double* out_value)
{
     if (!out_value || iters < 1) return set_error(HIQ_ERR_ARG, "hiqk_microbench: bad argument");
     cudaEvent_t e0, e1;
     HIQ_CUDA(cudaEventCreate(&e0));
     HIQ_CUDA(cudaEventCreate(&e1));
     float best_ms = 1e30f;
     double work = 0.0;  // bytes or flops per launch
     if (what == HIQK_MB_COPY_GBS) {
          const uint64_t n = 1ull << 27;  // 2 GiB source + 2 GiB destination: far beyond the 126 MB L2
          double2 *src = nullptr, *dst = nullptr;
          HIQ_CUDA(cudaMalloc(&src, n * sizeof(double2)));
          HIQ_CUDA(cudaMalloc(&dst, n * sizeof(double2)));
          HIQ_CUDA(cudaMemset(src, 0, n * sizeof(double2)));
          for (int r = 0; r < iters + 2; ++r) {
               cudaEventRecord(e0);
               mb_copy_kernel<<<num_sms() * 32, 256>>>(src, dst, n);
               cudaEventRecord(e1);
               HIQ_CUDA(cudaEventSynchronize(e1));
               float ms;
               cudaEventElapsedTime(&ms, e0, e1);
               if (r >= 2) best_ms = std::min(best_ms, ms);
          }
          count_launch(iters + 2);
          cudaFree(src);
          cudaFree(dst);
          work = 2.0 * n * sizeof(double2);
          *out_value = work / (best_ms * 1e-3) / 1e9;
     }
     else if (what == HIQK_MB_DFMA_TFLOPS || what == HIQK_MB_DMMA_TFLOPS) {
          double* out = nullptr;
          HIQ_CUDA(cudaMalloc(&out, sizeof(double)));
          const int inner = 4096;
          const unsigned grid = num_sms() * 8;
          for (int r = 0; r < iters + 2; ++r) {
               cudaEventRecord(e0);
               if (what == HIQK_MB_DFMA_TFLOPS) mb_dfma_kernel<<<grid, 256>>>(out, inner, 0.999999, 1e-7);
               else mb_dmma_kernel<<<grid, 256>>>(out, inner, 0.999999, 1e-7);
               cudaEventRecord(e1);
               HIQ_CUDA(cudaEventSynchronize(e1));
               float ms;
               cudaEventElapsedTime(&ms, e0, e1);
               if (r >= 2) best_ms = std::min(best_ms, ms);
          }
          count_launch(iters + 2);
          cudaFree(out);
          if (what == HIQK_MB_DFMA_TFLOPS) work = 2.0 * 8 * inner * 256.0 * grid;
          else work = 2.0 * 256 * 8 * inner * (256.0 / 32) * grid;  // m8n8k4 = 256 FMA per warp instruction
          *out_value = work / (best_ms * 1e-3) / 1e12;
     }
     else {
          cudaEventDestroy(e0);
          cudaEventDestroy(e1);
          return set_error(HIQ_ERR_ARG, "hiqk_microbench: unknown benchmark");
     }
     cudaEventDestroy(e0);
     cudaEventDestroy(e1);
     return check_launch("microbench");
}
