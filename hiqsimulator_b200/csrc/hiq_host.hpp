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
// grid-size cap of the persistent kernels: `natural` unless hiqk_debug_set_max_grid() lowered it (tests use
// that to drive the multi-iteration / prefetch paths on small slabs)
uint64_t grid_cap(uint64_t natural);
// multiprocessors of the current device (cached per device; 148 on B200): persistent grids are multiples of it
int num_sms();

struct DiagProg;
// host builders of the batched-diagonal program (defined in stream_kernels.cu)
int choose_u_positions(int L, const hiqk_diag_op* ops, int n_ops, uint64_t exclude, int want, int* out);
int build_diag_prog(DiagProg& p, int L, const hiqk_diag_op* ops, int n_ops, const int* upos, int n_u, uint64_t target_mask,
                    uint64_t tid_mask, int* order_out, const char* who);

#define HIQ_CUDA(call)                                          \
     do {                                                       \
          int _rc = ::hiq::check_cuda((call), #call);           \
          if (_rc != HIQ_OK) return _rc;                        \
     } while (0)

}  // namespace hiq
