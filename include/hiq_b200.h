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
#define HIQK_DENSE_DIRECT_FULL 4 /* DIRECT without the block-structure shortcut (A/B measurements) */

/* Dense k-qubit gate, k = 1..5, in place over a slab of 2^L amplitudes:
 * for every base index I with all target bits 0 and (I & ctrl_mask) == ctrl_mask
 *   out[b] = sum_c m[b][c] * in[c],  element c at I + sum_l c_l << slots[l].
 * `matrix` is HOST memory, 2^k x 2^k row-major complex128; matrix bit l <-> slots[l]
 * (slots need not be sorted).
 * Replaces kernelK<V,M,kernel_core> k=1..5 (reference:
 * src/simulator-mpi/kernels/intrin/kernel{1..5}.hpp; call site SimulatorMPI.cpp:470-515). */
int hiqk_apply_dense(void* slab, int L, int k, const int* slots, const double* matrix,
                     uint64_t ctrl_mask, int variant, void* stream);

/* The variant HIQK_DENSE_AUTO resolves to for these targets (so callers can label timings). */
int hiqk_dense_pick_variant(int L, int k, const int* slots);

/* Block structure the DIRECT kernels exploit (host-only, no device needed): returns the number of MIXING
 * index bits of `matrix` (2^k x 2^k row-major complex128, host memory) and, if `order` is not NULL, the bit
 * order that puts them first (order[i] = original bit at new position i).  A bit is a SELECT bit when every
 * entry coupling two indices that differ in it is exactly zero: the qubit multiplexes the gate on the others
 * (controls and diagonal factors folded into a fused cluster, reference fusion_mpi.hpp:167-187, produce
 * these), and row b of the product only meets the 2^mixing columns that share its select bits — the same
 * result as the full product of the reference kernels at 2^(k - mixing) times fewer flops.  -1 on bad input. */
int hiqk_dense_block_shape(int k, const double* matrix, int* order);
/* Mixing bits the DIRECT kernels will really use for this matrix: the block shape for k = 2..4, k otherwise
 * or when HIQ_DENSE_BLOCKS=0 is set in the environment (so callers can label timings).  Host only. */
int hiqk_dense_direct_mixing_bits(int k, const double* matrix);

/* Diagonal k-qubit gate: psi[i] *= diag[d], d = target bits of i gathered in
 * matrix-bit order, where (i & ctrl_mask) == ctrl_mask. `diag` is HOST memory, 2^k complex128.
 * Replaces kernel_core_diag (reference: kernels/intrin/kernels_diag.hpp:35-144). */
int hiqk_apply_diag(void* slab, int L, int k, const int* slots, const double* diag,
                    uint64_t ctrl_mask, void* stream);

/* One diagonal factor of a batched launch: psi[i] *= lut[d], d = bits of i at `slots`
 * (bit l of d <-> slots[l]); lut holds 2^k interleaved complex128 values. */
typedef struct hiqk_diag_op {
     int k;          /* 0..5 (k = 0: a global scalar, lut[0..1]) */
     int slots[5];
     double lut[64];
} hiqk_diag_op;
#define HIQK_MAX_DIAG_OPS 16

/* Several diagonal gates in ONE pass over the slab: psi[i] *= prod_j lut_j[bits_j(i)], n_ops <=
 * HIQK_MAX_DIAG_OPS, no control masks.  Each op is what one launch of kernel_core_diag does in the
 * reference (kernels/intrin/kernels_diag.hpp:35-144); diagonal factors compose by multiplication,
 * so consecutive diagonal passes of the reference's plan cost one HBM pass here. */
int hiqk_apply_diag_batch(void* slab, int L, const hiqk_diag_op* ops, int n_ops, void* stream);

/* Dense k-qubit gate (k <= 4, no control mask, lowest target slot >= 2 — the DIRECT kernel) preceded by
 * n_pre <= HIQK_MAX_DIAG_OPS diagonal factors applied to the loaded tuple in the same pass:
 * psi <- M * (prod_j D_j) * psi.  Replaces n_pre launches of kernel_core_diag + one of kernelK. */
int hiqk_apply_dense_prediag(void* slab, int L, int k, const int* slots, const double* matrix,
                             const hiqk_diag_op* pre, int n_pre, void* stream);
/* 1 if hiqk_apply_dense_prediag accepts these targets (otherwise apply the diagonals with
 * hiqk_apply_diag_batch first). */
int hiqk_dense_prediag_supported(int L, int k, const int* slots);

/* The kernel parameters hiqk_apply_dense / hiqk_apply_diag_batch / hiqk_apply_dense_prediag would launch with, written to host memory —
 * no device call, no computation: how the index is split (thread | per-thread tuples | chunk), the class of every
 * diagonal factor (one per CTA and chunk / per thread and chunk / per element / on the gate's targets), the partial
 * selector tables and the class-E pattern tables, exactly as the launchers encode them.  Layout: a header of uint32 words
 * (magic, constants and the byte offset of every field; see hiqk_diag_batch_image in csrc/stream_kernels.cu and
 * direct_pre_image in csrc/apply_dense.cu) followed by the parameter structure with a null slab pointer.
 * tests/diag_emulator.py executes the images the way the kernels do.  HIQ_OK or HIQ_ERR_*. */
size_t hiqk_dense_image_bytes(void);
int hiqk_dense_image(int L, int k, const int* slots, const double* matrix, uint64_t ctrl_mask, int variant, void* image,
                     size_t image_bytes); /* hiqk_apply_dense: resolved variant + its parameters (tests/dense_emulator.py) */
size_t hiqk_diag_batch_image_bytes(void);
int hiqk_diag_batch_image(int L, const hiqk_diag_op* ops, int n_ops, void* image, size_t image_bytes);
size_t hiqk_dense_prediag_image_bytes(void);
int hiqk_dense_prediag_image(int L, int k, const int* slots, const double* matrix, const hiqk_diag_op* pre, int n_pre,
                             void* image, size_t image_bytes);

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

/* Tile-resident gate program: up to HIQK_TILE_MAX_STEPS consecutive dense fused gates of the plan (1..4 targets, no
 * control mask), each preceded by up to HIQK_TILE_MAX_OPS diagonal factors, applied in ONE pass over the slab: a CTA
 * keeps a tile of 2^11 or 2^12 amplitudes (>= 4 low slots x every combination of the higher targets of the run) in
 * shared memory while the gates go by.  Same result as hiqk_apply_dense_prediag called once per gate (reference: one
 * kernelK / kernel_core_diag sweep per fused cluster, SimulatorMPI.cpp:470-515).  hiqk_tile_program_fits returns the tile
 * size in bits (11 / 12) when the run can be taken, 0 when it cannot (targets too spread out, too many tables). */
#define HIQK_TILE_MAX_STEPS 4
#define HIQK_TILE_MAX_OPS 16
typedef struct hiqk_tile_step {
     int k;
     int slots[5];               /* matrix index bit l <-> slab slot slots[l] */
     const double* matrix;       /* 2^k x 2^k complex128, row-major, interleaved re/im */
     const hiqk_diag_op* pre;    /* diagonal factors applied before this gate (after the previous gate of the run) */
     int n_pre;
} hiqk_tile_step;
int hiqk_tile_program_fits(int L, int n_steps, const hiqk_tile_step* steps);
/* 1 when the 2^k x 2^k matrix has exactly one nonzero entry per row and column (a permutation with phases: fused X / Y / Z /
 * phase gates) — a tile program applies it without summing products.  Host only. */
int hiqk_dense_is_monomial(int k, const double* matrix);
int hiqk_apply_tile_program(void* slab, int L, int n_steps, const hiqk_tile_step* steps, void* stream);
/* The kernel-parameter image hiqk_apply_tile_program would launch with, written to host memory — no device call, no
 * computation: tile choice, tuple layout, bank swizzle, matrices in planned bit order, diagonal-op classes and table
 * pool exactly as the launcher encodes them.  Layout: 64 x uint32 header (magic 'HQTP', tile bits, structure sizes and
 * the byte offset of every field: see hiqk_tile_program_image in csrc/tile_program.cu), the parameter structure (slab and
 * table pointers null), the table pool (complex128).  Lets the launcher's host logic be checked where no GPU is
 * (tests/tile_emulator.py interprets the image with numpy).  Returns HIQ_OK or HIQ_ERR_* like every other call;
 * hiqk_tile_program_image_bytes() = the buffer size to pass. */
size_t hiqk_tile_program_image_bytes(void);
int hiqk_tile_program_image(int L, int n_steps, const hiqk_tile_step* steps, void* image, size_t image_bytes);

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

/* In-place exchange over peer-mapped slabs (one launch per GPU, no staging): for every free index f in
 * [begin[k], begin[k]+count[k]) and peer k, swap  local[deposit(f) | spread(peer_pats[k])]  with
 * peer_slabs[k][deposit(f) | spread(my_pat)]  (patterns over the sorted swapped slots, as in hiqk_swap_pack).
 * The two GPUs of a pair must be given complementary ranges.  Replaces pack + mpi::all_to_all + unpack
 * (reference: SwapperMT.cpp:30-126) when the peers' slabs are mapped into this process (NVLink P2P). */
int hiqk_swap_p2p(void* local, void* const* peer_slabs, int n_peers, int L, int q, const int* slots,
                  const uint64_t* peer_pats, uint64_t my_pat, const uint64_t* begin, const uint64_t* count, void* stream);

/* Packed exchange, all peers of a swap group in one launch: for every peer k and free index f in [begin, begin+count)
 *   pack != 0:  bufs[k][f - begin] = slab[deposit(f) | spread(peer_pats[k])]      (gather; bufs may be peer memory)
 *   pack == 0:  slab[deposit(f) | spread(peer_pats[k])] = bufs[k][f - begin]      (scatter)
 * Same index convention as hiqk_swap_pack; four independent 128-bit accesses in flight per thread.  The engine uses it
 * when a swapped slot is low: contiguous full-line NVLink traffic instead of 16-64 B runs (reference data movement:
 * swapping.hpp:33-68 pack, SwapperMT.cpp:48-86 unpack). */
int hiqk_swap_move(void* slab, int L, int q, const int* slots, int n_peers, const uint64_t* peer_pats, uint64_t begin,
                   uint64_t count, void* const* bufs, int pack, void* stream);

/* y[j] += (a_re + i a_im) * x[j] for every j < 2^L with (j & mask) == val; x may alias y.  Taylor accumulation of
 * emulate_time_evolution. */
int hiqk_axpy_masked(void* y, const void* x, int L, uint64_t mask, uint64_t val, double a_re, double a_im, void* stream);

/* ---- Pauli-operator passes ------------------------------------------------------------------
 * The reference wrapper calls get_expectation_value / apply_qubit_operator on its C++ simulator
 * (reference: hiq/projectq/backends/_sim/_simulator_mpi.py:180-183, 220-223) although the reference
 * class exports neither (_cppsim_mpi.cpp:63-82); semantics = ProjectQ simulator.hpp.  A Pauli string
 * acts as (P psi)[i ^ x] = i^{#Y} (-1)^{popcount(i & z)} psi[i]; the terms of an operator that share the
 * flip mask x form one group with F(i) = sum_t c_t (-1)^{popcount(i & zmask_t)} (c_t carries i^{#Y}). */
#define HIQK_MAX_PAULI_TERMS 64
typedef struct hiqk_pauli_term {
     uint64_t zmask; /* sign mask over the local slots (Z and Y factors) */
     double re, im;  /* coefficient, times i^{#Y}, times the sign contributed by global qubits */
} hiqk_pauli_term;
/* d_out[0..1] = Re, Im of  sum_{i in [begin, begin+count)} conj(slab[i ^ xmask]) F(i) S[i],  S[i] = src[i - begin]
 * (src != NULL: amplitudes of a partner GPU staged on this one) or slab[i].  src == NULL over the whole slab
 * reads every amplitude once (pairs (i, i ^ xmask) are visited together).  workspace as for hiqk_prob_masked. */
int hiqk_pauli_expect(const void* slab, int L, uint64_t xmask, const hiqk_pauli_term* terms, int n_terms, const void* src,
                      uint64_t begin, uint64_t count, double* d_out, void* workspace, void* stream);
/* acc == NULL: slab <- P slab in place over the whole slab, (P psi)[i ^ xmask] = F(i) psi[i] (src NULL, begin 0,
 * count 2^L).  Otherwise acc[i ^ xmask] (+)= F(i) S[i] for i in [begin, begin+count) (accumulate != 0 adds). */
int hiqk_pauli_apply(void* slab, int L, uint64_t xmask, const hiqk_pauli_term* terms, int n_terms, void* acc, int accumulate,
                     const void* src, uint64_t begin, uint64_t count, void* stream);

/* ---- register permutation (emulate_math) ------------------------------------------------------
 * ProjectQ's emulate_math (simulator.hpp; the reference C++ class throws, SimulatorMPI.hpp:217-225,
 * call site _simulator_mpi.py:459-468): basis states that satisfy the control mask have the value v of a
 * register replaced by f(v).  The device pass is a gather through the inverse map:
 *   dst[j] = slabs[s >> L][s & (2^L - 1)],  G = rank << L | j,  s = G with the register bits replaced by f^-1(v)
 * where register bit b is index bit pos[b] (>= L: a global qubit, read from a peer-mapped slab). */
#define HIQK_PERM_TABLE 0   /* f^-1 given as a device table of 2^n_bits uint32 entries */
#define HIQK_PERM_ADD 1     /* f(v) = (v + a) mod 2^n_bits */
#define HIQK_PERM_ADD_MOD 2 /* f(v) = (v + a) mod N for v < N, v otherwise */
#define HIQK_PERM_MUL_MOD 3 /* f(v) = (a v) mod N for v < N, v otherwise; gcd(a, N) = 1, N <= 2^32 */
typedef struct hiqk_perm {
     int kind;
     int n_bits;
     int pos[40];
     uint64_t ctrl_mask; /* over the global amplitude index */
     uint64_t a, N;      /* the FORWARD constants; the launcher inverts them */
     const uint32_t* table;
} hiqk_perm;
int hiqk_permute_gather(void* dst, const void* const* slabs, int n_slabs, int rank, int L, const hiqk_perm* perm, void* stream);
/* a^-1 mod N (host helper; error when gcd(a, N) != 1) */
int hiq_modinv(uint64_t a, uint64_t N, uint64_t* out);

/* Micro-benchmarks used by bench.py to state the roofline denominators next to the
 * kernels: device copy GB/s, FP64 FMA TFLOP/s (DFMA) and FP64 tensor TFLOP/s (DMMA). */
int hiqk_microbench(int what, int iters, double* out_value);
#define HIQK_MB_COPY_GBS 0
#define HIQK_MB_DFMA_TFLOPS 1
#define HIQK_MB_DMMA_TFLOPS 2

/* Test hook: cap the grid of the persistent kernels at `max_ctas` CTAs (0 = natural size) so that small slabs
 * exercise the multi-iteration and prefetch paths that large slabs take. */
int hiqk_debug_set_max_grid(int max_ctas);

/* Number of kernels launched by this library in the calling process (bench evidence). */
uint64_t hiqk_launch_count(void);

/* ------------------------------------------------------------------------- *
 * Engine-level entry points (hiq_*): one handle per rank / GPU.
 * They are the methods of the reference's pybind class, in the same order and with the
 * same argument meaning (reference: _cppsim_mpi.cpp:63-82).  Qubit ids are int64; bit
 * strings are one byte per bit (0/1); matrices are row-major interleaved complex128.
 * ------------------------------------------------------------------------- */
typedef struct hiq_engine hiq_engine;

#define HIQ_FLAG_DRY_RUN 1 /* no device: host logic only, every device op is recorded as a descriptor */
#define HIQ_FLAG_TRACE 2   /* also record descriptors while executing on the GPU */
#define HIQ_FLAG_TIMING 4  /* bracket every fused pass / swap with CUDA events on the engine stream */
#define HIQ_FLAG_NO_BATCH 8 /* one launch per fused gate of the plan: do not fold diagonal passes into neighbours */

/* descriptor kinds (what the host hands to the device layer) */
#define HIQ_DESC_NONE 0
#define HIQ_DESC_DENSE 1
#define HIQ_DESC_DIAG 2
#define HIQ_DESC_SCALE 3
#define HIQ_DESC_SWAP 4
#define HIQ_DESC_GROW 5
#define HIQ_DESC_FILL 6
/* operator-level passes: aux = [mode, flip mask over local slots, source rank, zmask_0 .. zmask_{k-1}], payload = the k
 * coefficients; mode 0 = in place (APPLY) / sum (EXPECT), 1 = overwrite the accumulator, 2 = add to it */
#define HIQ_DESC_PAULI_EXPECT 7
#define HIQ_DESC_PAULI_APPLY 8
#define HIQ_DESC_PAULI_COMMIT 9 /* slab <- accumulator */
/* aux = [HIQK_PERM_*, a, N, ctrl mask over the global index, this rank takes part (0/1), pos_0 .. pos_{k-1}, inverse table] */
#define HIQ_DESC_PERMUTE 10
#define HIQ_DESC_LOAD 11 /* set_wavefunction: aux = [slice of the host vector this rank copies, -1 = zeros] */
#define HIQ_DESC_TILE 12 /* timing records only: one tile-resident launch that carried k dense gates of the plan */
/* launch trace only (hiq_launch_trace_*): ONE device launch that carries several passes of the plan.  k = steps,
 * slots[0] = form (HIQ_LAUNCH_*); aux = per step [targets k_s (-1: no dense gate), slot_0 .. slot_{k_s-1}, n_ops, then per
 * diagonal factor: k_o, slot_0 .. slot_{k_o-1}]; payload = per step the 4^k_s matrix entries, then per factor its 2^k_o
 * table.  Every step = its factors first, then its dense gate. */
#define HIQ_DESC_LAUNCH 13
#define HIQ_LAUNCH_DIAG_BATCH 0    /* hiqk_apply_diag_batch */
#define HIQ_LAUNCH_DENSE_PREDIAG 1 /* hiqk_apply_dense_prediag */
#define HIQ_LAUNCH_TILE 2          /* hiqk_apply_tile_program */

typedef struct hiq_descriptor {
     int kind;           /* HIQ_DESC_* */
     int k;              /* targets (dense/diag), swapped pairs (swap), new local count (grow) */
     int slots[5];       /* matrix bit l <-> local slot slots[l] */
     uint64_t ctrl_mask; /* over local slots */
     int n_payload;      /* complex128 values in the payload: 4^k dense, 2^k diag, 1 scale/fill */
     int n_aux;          /* swap: 2k ints gpos0, slot0, gpos1, slot1 ... */
} hiq_descriptor;

/* rank 0 obtains the 128-byte NCCL unique id; the launcher hands it to every rank's hiq_create */
int hiq_comm_unique_id(void* out128);

/* SimulatorMPI(seed, max_local, max_cluster_size) (reference: SimulatorMPI.cpp:66-104).
 * rank/world_size replace the MPI communicator (one process per GPU); nccl_id may be NULL when
 * world_size == 1; device is the CUDA ordinal. */
int hiq_create(uint64_t seed, int max_local, int max_cluster_size, int rank, int world_size, const void* nccl_id,
               int device, int flags, hiq_engine** out);
int hiq_destroy(hiq_engine* e);

int hiq_allocate_qubit(hiq_engine* e, int64_t id);                                   /* AllocateQubit  :184-216 */
int hiq_allocate_qureg(hiq_engine* e, const int64_t* ids, int n, double init_re, double init_im); /* AllocateQureg :218-251 */
int hiq_deallocate_qubit(hiq_engine* e, int64_t id);                                 /* DeallocateQubit :369-426 */
int hiq_apply_controlled_gate(hiq_engine* e, const double* matrix, int dim, const int64_t* ids, int n_ids,
                              const int64_t* ctrls, int n_ctrls);                    /* ApplyGate :710-815 */
int hiq_run(hiq_engine* e);                                                          /* Run :441-541 */
int hiq_swap_qubits(hiq_engine* e, const int64_t* pairs, int n);                     /* SwapQubitsWrapper :1060-1138 */
int hiq_measure_qubits(hiq_engine* e, const int64_t* ids, int n, uint8_t* out_bits); /* MeasureQubits :897-1008 */
int hiq_get_probability(hiq_engine* e, const uint8_t* bits, const int64_t* ids, int n, double* out); /* :561-600 */
int hiq_get_amplitude(hiq_engine* e, const uint8_t* bits, const int64_t* ids, int n, double* out_re_im); /* :602-650 */
int hiq_collapse_wavefunction(hiq_engine* e, const int64_t* ids, const uint8_t* values, int n); /* :1010-1058 */
int hiq_entropy(hiq_engine* e, double* out);                                         /* Entropy :681-694 */
/* kind 0: locals ++ globals (get_qubits_ids), 1: locals, 2: globals; empty global slots are -1 */
int hiq_get_qubits_ids(hiq_engine* e, int kind, int64_t* out, int cap, int* n);      /* :652-679 */
int hiq_set_qubits_perm(hiq_engine* e, const int64_t* p, int n);                     /* :661-667 */
/* cheat_local (:543-559): id -> bit position map, and a host copy of the local slab
 * (host_dst may be NULL to query sizes only; *n_amps = 2^L) */
int hiq_cheat_local(hiq_engine* e, int64_t* ids, int* pos, int cap, int* n_map, void* host_dst, uint64_t cap_amps,
                    uint64_t* n_amps);
/* cheat() (reference: _simulator_mpi.py:348-380, an MPI Allgather of the rank slabs): the same map as
 * hiq_cheat_local and world_size * 2^L amplitudes, rank-major, on every rank (host_dst may be NULL to query sizes) */
int hiq_cheat(hiq_engine* e, int64_t* ids, int* pos, int cap, int* n_map, void* host_dst, uint64_t cap_amps, uint64_t* n_amps);

/* Operator-level calls the reference wrapper makes on its simulator object but the reference class never
 * exported (reference: _simulator_mpi.py:180-183, 220-223, 305, 459-468).  Semantics: ProjectQ simulator.hpp.
 * A Pauli operator is a list of n_terms terms; term t has the factors [term_offsets[t], term_offsets[t+1]):
 * factor f acts with factor_pauli[f] in {'X','Y','Z'} on qubit ids[factor_index[f]]; coefs = (re, im) per term. */
int hiq_get_expectation_value(hiq_engine* e, const int* term_offsets, const int* factor_index, const char* factor_pauli,
                              const double* coefs_re_im, int n_terms, const int64_t* ids, int n_ids, double* out);
int hiq_apply_qubit_operator(hiq_engine* e, const int* term_offsets, const int* factor_index, const char* factor_pauli,
                             const double* coefs_re_im, int n_terms, const int64_t* ids, int n_ids);
/* psi <- exp(-i t H) psi on the part of the register whose control qubits are 1, H = the Pauli operator (same encoding as
 * above, real or complex coefficients).  ProjectQ's algorithm: s = |t| * (sum of |coefficients| of the non-identity terms)
 * + 1 slices, each a Taylor series summed until the norm of a term drops below 1e-12; the identity terms enter as a phase.
 * Needs two buffers of the slab's size next to the slab (reference call site: _simulator_mpi.py:469-475; the reference
 * class exports no such method and its only specification is the commented-out test _simulator_mpi_test.py:481-537). */
int hiq_emulate_time_evolution(hiq_engine* e, const int* term_offsets, const int* factor_index, const char* factor_pauli,
                               const double* coefs_re_im, int n_terms, double time, const int64_t* ids, int n_ids,
                               const int64_t* ctrls, int n_ctrls);
/* amps: 2^n_ids complex128 on the host (the whole vector on every rank); index bit i <-> ids[i] */
int hiq_set_wavefunction(hiq_engine* e, const double* amps_re_im, uint64_t n_amps, const int64_t* ids, int n_ids);
/* emulate_math with the function tabulated over the concatenated registers (reg_ids in bit order, register 0
 * lowest): table[v] = f(v), 2^n_reg entries; must be a permutation */
int hiq_emulate_math_table(hiq_engine* e, const uint64_t* table, uint64_t table_len, const int64_t* reg_ids, int n_reg,
                           const int64_t* ctrls, int n_ctrls);
/* closed forms of ProjectQ's math gates on one register: kind = HIQK_PERM_ADD / _ADD_MOD / _MUL_MOD */
int hiq_emulate_math_const(hiq_engine* e, int kind, uint64_t a, uint64_t N, const int64_t* reg_ids, int n_reg,
                           const int64_t* ctrls, int n_ctrls);

/* device pointer of the local slab and L (zero-copy views; valid until the next (de)allocation) */
int hiq_local_slab(hiq_engine* e, void** dev_ptr, int* L);
int hiq_set_local_slab(hiq_engine* e, const void* host_src, uint64_t n_amps);
/* wait until every operation issued so far has completed on the device */
int hiq_synchronize(hiq_engine* e);
int hiq_rank(hiq_engine* e, int* rank, int* world_size);
/* force a dense kernel variant (HIQK_DENSE_*) for every fused pass; 0 = automatic */
int hiq_set_dense_variant(hiq_engine* e, int variant);

/* counters and timers in the spirit of the reference's stage statistics (:106-134, :1148-1157) */
typedef struct hiq_stats {
     uint64_t total_gates, total_runs, total_stages, total_swaps;
     uint64_t dense_passes, diag_passes, scale_passes, skipped_passes;
     double runs_s, swaps_s, measures_s, allocs_s, deallocs_s;
     double swap_bytes_sent;
     uint64_t swaps_p2p, swaps_staged; /* exchanges done in place over peer-mapped slabs / through the staged NCCL pipeline */
     double h2d_bytes, d2h_bytes; /* host<->device traffic issued by the engine (descriptor payloads, results) */
     uint64_t gate_launches;      /* device launches that carried the dense/diag/scale passes (<= their sum) */
     uint64_t swaps_packed;       /* exchanges done by packing + reading the peers' staging buffers (low swapped slots) */
     double ctor_s, slab_grow_s, peer_map_s; /* host seconds: constructor, mapping physical memory into the slab, peer-slab handshakes */
     uint64_t tile_launches, tile_steps;     /* multi-gate tile-resident launches and the dense fused gates they carried */
} hiq_stats;
int hiq_get_stats(hiq_engine* e, hiq_stats* out);

/* per-launch device times (HIQ_FLAG_TIMING): waits for the stream, then returns and clears the
 * records since the last call.  kind = HIQ_DESC_*, variant = HIQK_DENSE_* (dense only), k = targets
 * (dense), ops in the batch (diag) or swapped pairs (swap); n_ref = fused-gate passes of the
 * reference's plan that the launch carried (1 + folded diagonal passes). */
int hiq_collect_timings(hiq_engine* e, double* ms, int* kind, int* k, int* variant, int* n_ref, int cap, int* n);
/* cudaStream_t of the engine (for event timing by the caller) */
int hiq_stream(hiq_engine* e, void** stream);

/* descriptor trace (HIQ_FLAG_DRY_RUN / HIQ_FLAG_TRACE) */
int hiq_trace_count(hiq_engine* e, int* n);
int hiq_trace_get(hiq_engine* e, int i, hiq_descriptor* d, double* payload, int cap_payload, int64_t* aux, int cap_aux);
int hiq_trace_clear(hiq_engine* e);
/* launch trace of a HIQ_FLAG_DRY_RUN engine: what would really be launched, in device order — the plan's descriptors
 * after diagonal folding and tile-run grouping (non-gate descriptors and single gates unchanged, HIQ_DESC_LAUNCH records
 * for the launches that carry several passes).  Same accessors as the descriptor trace; cleared by hiq_trace_clear. */
int hiq_launch_trace_count(hiq_engine* e, int* n);
int hiq_launch_trace_get(hiq_engine* e, int i, hiq_descriptor* d, double* payload, int cap_payload, int64_t* aux, int cap_aux);

#ifdef __cplusplus
}
#endif
#endif /* HIQ_B200_H */
