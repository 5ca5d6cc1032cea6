// Host-side plumbing shared by the launchers: error reporting for the C ABI and launch accounting.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>
#include <string>

#include "../../include/hiq_b200.h"

namespace hiq {

// Records `msg` as the calling thread's last error and returns `code`.
int set_error(int code, const std::string& msg);
// cudaGetLastError() after a launch; returns HIQ_OK or records and returns HIQ_ERR_CUDA.
int check_launch(const char* what);
int check_cuda(cudaError_t e, const char* what);
void count_launch(unsigned n = 1);

struct DiagBatch;
// hiqk_diag_op[] (host, validated) -> kernel-side batch; defined in stream_kernels.cu
int make_diag_batch(DiagBatch& b, int L, const hiqk_diag_op* ops, int n_ops, uint64_t varying_mask, uint64_t force_lo_mask,
                    const char* who);

#define HIQ_CUDA(call)                                          \
     do {                                                       \
          int _rc = ::hiq::check_cuda((call), #call);           \
          if (_rc != HIQ_OK) return _rc;                        \
     } while (0)

}  // namespace hiq
