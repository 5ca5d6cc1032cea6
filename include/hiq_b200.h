/* hiq_b200.h — C ABI of the B200-native state-vector engine.
 *
 * Drop-in boundary for the hot path behind HiQsimulator's pybind11 class
 * `_cppsim_mpi.SimulatorMPI` (reference: /root/reference/_cppsim_mpi.cpp:61-83,
 * src/simulator-mpi/SimulatorMPI.hpp:43-308).  Two layers:
 *
 *   hiqk_*  device-level launchers: one call = one kernel pass over a local
 *           amplitude slab that is already resident in HBM.  They take plain
 *           device pointers and sizes; `stream` is a cudaStream_t passed as
 *           void* (NULL = default stream).  Each cites the reference kernel
 *           it replaces.
 *   hiq_*   engine-level entry points: an opaque handle that owns the slab,
 *           the qubit->slot maps, the gate-fusion accumulator and the RNG —
 *           one per rank/GPU, exactly the methods of the reference class.
 *
 * Conventions: complex128 values are interleaved (re, im) doubles — the memory
 * layout of std::complex<double> and of CUDA double2.  Every function returns
 * 0 on success and a non-zero code on failure; hiq_last_error() then returns a
 * thread-local message (the reference throws std::runtime_error with free text,
 * SimulatorMPI.cpp:160-165 etc.; the pybind layer re-throws it).
 * There is no CPU fallback: without a CUDA device every compute call fails.
 */
#ifndef HIQ_B200_H
#define HIQ_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HIQ_OK 0
#define HIQ_ERR_RUNTIME 1  /* maps to std::runtime_error / Python RuntimeError */
#define HIQ_ERR_CUDA 2
#define HIQ_ERR_ARG 3

const char* hiq_last_error(void);
/* "hiq_b200 <version> sm_100a" */
const char* hiq_version(void);
/* number of visible CUDA devices, or -1 if the runtime cannot be initialised */
int hiq_device_count(void);

/* ------------------------------------------------------------------------- *
 * Device-level launchers (hiqk_*)
 * ------------------------------------------------------------------------- */

/* Kernel variant selector for hiqk_apply_dense (0 = pick automatically). */
#define HIQK_DENSE_AUTO 0
#define HIQK_DENSE_DIRECT 1 /* one tuple per thread, registers only              */
#define HIQK_DENSE_TILED 2  /* shared-memory tile, for targets in the lowest slots */
#define HIQK_DENSE_DMMA 3   /* FP64 tensor-core path (k >= 2)                    */

/* Dense k-qubit gate, k = 1..5, in place over a slab of 2^L amplitudes:
 * for every base index I with all target bits 0 and (I & ctrl_mask) == ctrl_mask
 *   out[b] = sum_c m[b][c] * in[c],  element c at I + sum_l c_l << slots[l].
 * `matrix` is HOST memory, 2^k x 2^k row-major complex128; matrix bit l <-> slots[l]
 * (slots need not be sorted).
 * Replaces kernelK<V,M,kernel_core> k=1..5 (reference:
 * src/simulator-mpi/kernels/intrin/kernel{1..5}.hpp; call site SimulatorMPI.cpp:470-515). */
int hiqk_apply_dense(void* slab, int L, int k, const int* slots, const double* matrix,
                     uint64_t ctrl_mask, int variant, void* stream);

/* Diagonal k-qubit gate: psi[i] *= diag[d], d = target bits of i gathered in
 * matrix-bit order, where (i & ctrl_mask) == ctrl_mask. `diag` is HOST memory, 2^k complex128.
 * Replaces kernel_core_diag (reference: kernels/intrin/kernels_diag.hpp:35-144). */
int hiqk_apply_diag(void* slab, int L, int k, const int* slots, const double* diag,
                    uint64_t ctrl_mask, void* stream);

/* psi[i] *= (re + i*im) for the whole slab.
 * Replaces kernelK_diag1 (reference: kernels/intrin/kernels_diag.hpp:21-32). */
int hiqk_scale(void* slab, int L, double re, double im, void* stream);

/* Bytes of device scratch the reduction launchers need (`workspace`). */
size_t hiqk_workspace_bytes(void);

/* *d_out = sum over i with (i & mask) == val of |psi[i]|^2   (d_out: device double).
 * Replaces getProbability_internal's loop (reference: SimulatorMPI.cpp:852-860) and
 * norm() (funcs.hpp:402-414) with mask = 0. */
int hiqk_prob_masked(const void* slab, int L, uint64_t mask, uint64_t val, double* d_out,
                     void* workspace, void* stream);

/* d_out[b] = sum_{j < 2^L / n_blocks} |psi[b * 2^L / n_blocks + j]|^2, b < n_blocks
 * (n_blocks a power of two, <= 2^L). Replaces calcLocalApproxDistribution's block sums
 * (reference: SimulatorMPI.cpp:817-836); the prefix sum stays on the host. */
int hiqk_block_norms(const void* slab, int L, uint64_t n_blocks, double* d_out, void* stream);

/* d_out[0..1] = sum of |psi[i]|^2 over i with bit `slot` = 0 / = 1.
 * Replaces the classicality check of DeallocateLocalQubit (reference: SimulatorMPI.cpp:282-286). */
int hiqk_bit_norms(const void* slab, int L, int slot, double* d_out, void* workspace, void* stream);

/* *d_out = sum_i p_i log2 p_i, p_i = |psi_i|^2, p_i > 0 (reference: SimulatorMPI.cpp:681-690). */
int hiqk_entropy(const void* slab, int L, double* d_out, void* workspace, void* stream);

/* psi[i] = ((i & mask) == val) ? psi[i] * scale : 0   (reference: normalize(), SimulatorMPI.cpp:872-890) */
int hiqk_collapse(void* slab, int L, uint64_t mask, uint64_t val, double scale, void* stream);

/* psi[i] = (re, im) for i in [begin, begin+count)   (FillVector, SimulatorMPI.cpp:136-146) */
int hiqk_fill(void* slab, uint64_t begin, uint64_t count, double re, double im, void* stream);

/* Remove bit `slot` from the index space keeping the half where that bit == keep:
 * dst[j] = src[insert_bit(j, slot, keep)], j < 2^(L-1).  dst may alias the start of src
 * (the launcher stages through `scratch`, `scratch_amps` amplitudes, in index order).
 * Replaces the compaction loops of DeallocateLocalQubit (reference: SimulatorMPI.cpp:304-318). */
int hiqk_compact_bit(void* slab, int L, int slot, int keep, void* scratch, uint64_t scratch_amps,
                     void* stream);

/* Swap pack: gather the amplitudes whose swapped local slots spell pattern `pat`
 * (bit j of pat <-> j-th lowest swapped slot) into a contiguous buffer, for the
 * free-index range [begin, begin+count):  dst[f - begin] = psi[deposit(f) | spread(pat)].
 * Unpack is the inverse scatter.  Replace Swapping::doCalc + f_consumer2
 * (reference: src/simulator-mpi/swapping.hpp:33-68, SwapperMT.cpp:48-86). */
int hiqk_swap_pack(const void* slab, int L, int q, const int* slots, uint64_t pat, uint64_t begin,
                   uint64_t count, void* dst, void* stream);
int hiqk_swap_unpack(void* slab, int L, int q, const int* slots, uint64_t pat, uint64_t begin,
                     uint64_t count, const void* src, void* stream);

/* Micro-benchmarks used by bench.py to state the roofline denominators next to the
 * kernels: device copy GB/s, FP64 FMA TFLOP/s (DFMA) and FP64 tensor TFLOP/s (DMMA). */
int hiqk_microbench(int what, int iters, double* out_value);
#define HIQK_MB_COPY_GBS 0
#define HIQK_MB_DFMA_TFLOPS 1
#define HIQK_MB_DMMA_TFLOPS 2

/* Number of kernels launched by this library in the calling process (bench evidence). */
uint64_t hiqk_launch_count(void);

#ifdef __cplusplus
}
#endif
#endif /* HIQ_B200_H */
