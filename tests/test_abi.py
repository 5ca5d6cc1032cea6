"""The C-ABI library loads on a machine without a GPU and exports every function that
include/hiq_b200.h declares (no compute calls here)."""
import ctypes
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "hiq_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    names = re.findall(r"\b(hiqk?_[a-z0-9_]+)\s*\(", src)
    return sorted(set(names))


def test_header_declares_both_layers():
    names = declared_functions()
    assert "hiqk_apply_dense" in names and "hiq_create" in names and "hiq_measure_qubits" in names
    assert len(names) >= 40


def test_header_is_plain_c():
    """the boundary is a C ABI: include/hiq_b200.h compiles as C99 and as C++11 on its own (no torch, no CUDA headers)"""
    import shutil
    import subprocess
    hdr = os.path.join(ROOT, "include", "hiq_b200.h")
    for cc, flags in (("gcc", ["-std=c99", "-x", "c"]), ("g++", ["-std=c++11", "-x", "c++"])):
        if shutil.which(cc) is None:
            continue
        res = subprocess.run([cc, *flags, "-Wall", "-Wextra", "-pedantic", "-Werror", "-fsyntax-only", hdr], capture_output=True, text=True)
        assert res.returncode == 0, res.stderr


def test_library_exports_every_declared_symbol():
    from hiqsimulator_b200 import _lib
    lib = ctypes.CDLL(_lib.LIB_PATH)
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_signatures_cover_device_layer():
    from hiqsimulator_b200 import _lib
    declared = [n for n in declared_functions() if n.startswith("hiqk_")]
    assert set(declared) <= set(_lib.exported_symbols())


def test_compute_fails_loudly_without_gpu():
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from hiqsimulator_b200 import _cppsim_mpi as M
    M.init_world(0, 1, b"", 0, 0)
    with pytest.raises(RuntimeError):
        M.SimulatorMPI(1, 10, 4)  # no device, no fallback


def test_dense_block_shape_is_host_only_and_finds_select_bits():
    """hiqk_dense_block_shape needs no device: select bits = index bits no nonzero entry mixes; QFT-like
    clusters (Hadamards among controlled phases, fused as the reference fuses them) have them."""
    import numpy as np
    from hiqsimulator_b200 import kernels as K
    h = np.array([[1, 1], [1, -1]]) / np.sqrt(2)
    cr = np.diag([1, 1, 1, np.exp(0.3j)])
    # bits (0, 1, 2, 3): H on bit 2, controlled phases between the others and bit 2 -> only bit 2 mixes
    m = np.kron(np.eye(2), np.kron(h, np.eye(4)))
    d = np.ones(16, dtype=complex)
    for i in range(16):
        if (i >> 2) & 1 and i & 1:
            d[i] *= np.exp(0.7j)
        if (i >> 2) & 1 and (i >> 3) & 1:
            d[i] *= np.exp(0.2j)
    m = m @ np.diag(d)
    ks, order = K.dense_block_shape(m)
    assert ks == 1 and order[0] == 2 and sorted(order) == [0, 1, 2, 3]
    ks, order = K.dense_block_shape(np.kron(h, np.kron(np.eye(2), h)))
    assert ks == 2 and order[:2] == [0, 2]
    ks, order = K.dense_block_shape(np.kron(h, h))
    assert ks == 2 and order == [0, 1]
    ks, order = K.dense_block_shape(cr)  # diagonal: one nominal mixing bit
    assert ks == 1 and sorted(order) == [0, 1]


def test_reference_import_paths_resolve():
    """the two import statements of the reference's Python layer (reference: _simulator_mpi.py:39, cengines/__init__.py:15)
    resolve to this repository's modules through the hiq/ shim tree"""
    from hiq.projectq.backends._sim._cppsim_mpi import SimulatorMPI as SimulatorBackend
    from hiq.projectq.cengines._sched_cpp import ClusterScheduler, SwapScheduler
    from hiqsimulator_b200 import _cppsim_mpi, _sched_cpp
    assert SimulatorBackend is _cppsim_mpi.SimulatorMPI
    assert SwapScheduler is _sched_cpp.SwapScheduler and ClusterScheduler is _sched_cpp.ClusterScheduler
    for name in ("get_qubits_ids", "get_local_qubits_ids", "get_global_qubits_ids", "set_qubits_perm", "swap_qubits", "allocate_qureg",
                 "allocate_qubit", "deallocate_qubit", "measure_qubits", "apply_controlled_gate", "emulate_math", "get_amplitude",
                 "get_probability", "run", "entropy", "cheat_local", "collapse_wavefunction",   # reference: _cppsim_mpi.cpp:63-82
                 "get_expectation_value", "apply_qubit_operator", "set_wavefunction", "emulate_time_evolution", "cheat"):
        assert hasattr(SimulatorBackend, name), name


NULL_ARGUMENT_PROBE = r'''
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
lib.hiq_last_error.restype = ctypes.c_char_p
HIQ_OK, HIQ_ERR_ARG = 0, 1
e = ctypes.c_void_p()
i64, u64, dbl = ctypes.c_int64, ctypes.c_uint64, ctypes.c_double
rc = lib.hiq_create(u64(1), 10, 4, 0, 1, None, 0, 1, ctypes.byref(e))   # HIQ_FLAG_DRY_RUN: host logic only
assert rc == HIQ_OK, lib.hiq_last_error()
ids = (i64 * 4)(0, 1, 2, 3)
assert lib.hiq_allocate_qureg(e, ids, 4, dbl(0.0), dbl(0.0)) == HIQ_OK
m = (dbl * 8)(1, 0, 0, 0, 0, 0, 1, 0)
n = ctypes.c_int(0)
out = dbl(0.0)
bits = (ctypes.c_uint8 * 4)()
N = None
calls = {
    "create_null_out": lambda: lib.hiq_create(u64(1), 10, 4, 0, 1, N, 0, 1, N),
    "qureg_null_ids": lambda: lib.hiq_allocate_qureg(e, N, 3, dbl(0.0), dbl(0.0)),
    "qureg_negative": lambda: lib.hiq_allocate_qureg(e, ids, -1, dbl(0.0), dbl(0.0)),
    "gate_null_matrix": lambda: lib.hiq_apply_controlled_gate(e, N, 2, ids, 1, N, 0),
    "gate_null_ids": lambda: lib.hiq_apply_controlled_gate(e, m, 2, N, 1, N, 0),
    "gate_null_ctrls": lambda: lib.hiq_apply_controlled_gate(e, m, 2, ids, 1, N, 2),
    "gate_bad_dim": lambda: lib.hiq_apply_controlled_gate(e, m, 64, ids, 1, N, 0),
    "swap_null": lambda: lib.hiq_swap_qubits(e, N, 2),
    "measure_null_ids": lambda: lib.hiq_measure_qubits(e, N, 2, bits),
    "measure_null_out": lambda: lib.hiq_measure_qubits(e, ids, 2, N),
    "prob_null_bits": lambda: lib.hiq_get_probability(e, N, ids, 2, ctypes.byref(out)),
    "prob_null_ids": lambda: lib.hiq_get_probability(e, bits, N, 2, ctypes.byref(out)),
    "prob_null_out": lambda: lib.hiq_get_probability(e, bits, ids, 2, N),
    "amp_null_out": lambda: lib.hiq_get_amplitude(e, bits, ids, 4, N),
    "amp_null_ids": lambda: lib.hiq_get_amplitude(e, bits, N, 4, ctypes.byref(out)),
    "collapse_null": lambda: lib.hiq_collapse_wavefunction(e, ids, N, 2),
    "entropy_null": lambda: lib.hiq_entropy(e, N),
    "ids_null_count": lambda: lib.hiq_get_qubits_ids(e, 1, N, 0, N),
    "perm_null": lambda: lib.hiq_set_qubits_perm(e, N, 4),
    "expect_null_out": lambda: lib.hiq_get_expectation_value(e, N, N, N, N, 0, ids, 4, N),
    "expect_null_terms": lambda: lib.hiq_get_expectation_value(e, N, N, N, N, 2, ids, 4, ctypes.byref(out)),
    "operator_null_ids": lambda: lib.hiq_apply_qubit_operator(e, N, N, N, N, 0, N, 4),
    "evolution_null_ctrls": lambda: lib.hiq_emulate_time_evolution(e, N, N, N, N, 0, dbl(1.0), ids, 4, N, 1),
    "wavefunction_null": lambda: lib.hiq_set_wavefunction(e, N, u64(16), ids, 4),
    "math_table_null": lambda: lib.hiq_emulate_math_table(e, N, u64(4), ids, 2, N, 0),
    "math_const_null_reg": lambda: lib.hiq_emulate_math_const(e, 1, u64(1), u64(0), N, 2, N, 0),
    "slab_null": lambda: lib.hiq_local_slab(e, N, N),
    "set_slab_null": lambda: lib.hiq_set_local_slab(e, N, u64(16)),
    "rank_null": lambda: lib.hiq_rank(e, N, N),
    "stats_null": lambda: lib.hiq_get_stats(e, N),
    "timings_null": lambda: lib.hiq_collect_timings(e, N, N, N, N, N, 4, N),
    "stream_null": lambda: lib.hiq_stream(e, N),
    "trace_count_null": lambda: lib.hiq_trace_count(e, N),
    "launch_trace_count_null": lambda: lib.hiq_launch_trace_count(e, N),
    "trace_get_null": lambda: lib.hiq_trace_get(e, 0, N, N, 0, N, 0),
    "null_engine": lambda: lib.hiq_run(N),
    "modinv_null": lambda: lib.hiq_modinv(u64(3), u64(7), N),
    "unique_id_null": lambda: lib.hiq_comm_unique_id(N),
}
for name, f in calls.items():
    rc = f()
    assert rc != HIQ_OK, name
    assert lib.hiq_last_error(), name
    print("refused", name, rc)
# the engine is still usable afterwards: empty lists may come with null pointers
assert lib.hiq_apply_controlled_gate(e, m, 2, ids, 1, N, 0) == HIQ_OK
assert lib.hiq_run(e) == HIQ_OK
assert lib.hiq_get_qubits_ids(e, 1, N, 0, ctypes.byref(n)) == HIQ_OK and n.value == 4
assert lib.hiq_destroy(e) == HIQ_OK
print("PROBE_OK", len(calls))
'''


def test_entry_points_refuse_null_arguments_with_a_status_code():
    """status codes out, never a crash: every hiq_* entry point handed a null array / null output / negative length
    answers with an error code and a message (own process: a segmentation fault would show as a signal)"""
    import subprocess
    import sys
    from hiqsimulator_b200 import _lib
    res = subprocess.run([sys.executable, "-c", NULL_ARGUMENT_PROBE, _lib.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, (res.returncode, res.stdout[-1500:], res.stderr[-1500:])
    assert "PROBE_OK" in res.stdout


DEVICE_LAYER_NULL_PROBE = r'''
import ctypes, sys
lib = ctypes.CDLL(sys.argv[1])
lib.hiq_last_error.restype = ctypes.c_char_p
N=None; u64=ctypes.c_uint64; dbl=ctypes.c_double
slots=(ctypes.c_int*4)(2,3,4,5)
m=(dbl*512)()
calls = {
 "dense_null_slab": lambda: lib.hiqk_apply_dense(N,10,2,slots,m,u64(0),0,N),
 "dense_null_slots": lambda: lib.hiqk_apply_dense(ctypes.c_void_p(16),10,2,N,m,u64(0),0,N),
 "dense_null_matrix": lambda: lib.hiqk_apply_dense(ctypes.c_void_p(16),10,2,slots,N,u64(0),0,N),
 "diag_null": lambda: lib.hiqk_apply_diag(N,10,2,slots,m,u64(0),N),
 "diag_batch_null": lambda: lib.hiqk_apply_diag_batch(N,10,N,2,N),
 "diag_batch_null_ops": lambda: lib.hiqk_apply_diag_batch(ctypes.c_void_p(16),10,N,2,N),
 "prediag_null": lambda: lib.hiqk_apply_dense_prediag(N,10,2,slots,m,N,1,N),
 "prediag_null_pre": lambda: lib.hiqk_apply_dense_prediag(ctypes.c_void_p(16),12,2,slots,m,N,1,N),
 "scale_null": lambda: lib.hiqk_scale(N,10,dbl(1),dbl(0),N),
 "prob_null": lambda: lib.hiqk_prob_masked(N,10,u64(0),u64(0),N,N,N),
 "block_norms_null": lambda: lib.hiqk_block_norms(N,10,u64(4),N,N),
 "bit_norms_null": lambda: lib.hiqk_bit_norms(N,10,1,N,N,N),
 "entropy_null": lambda: lib.hiqk_entropy(N,10,N,N,N),
 "collapse_null": lambda: lib.hiqk_collapse(N,10,u64(1),u64(1),dbl(1),N),
 "fill_null": lambda: lib.hiqk_fill(N,u64(0),u64(4),dbl(1),dbl(0),N),
 "tile_null": lambda: lib.hiqk_apply_tile_program(N,12,1,N,N),
 "tile_null_steps": lambda: lib.hiqk_apply_tile_program(ctypes.c_void_p(16),12,1,N,N),
 "tile_fits_null": lambda: lib.hiqk_tile_program_fits(12,1,N) == 0 and 1,
 "tile_image_null": lambda: lib.hiqk_tile_program_image(12,1,N,N,ctypes.c_size_t(0)),
 "compact_null": lambda: lib.hiqk_compact_bit(N,10,1,0,N,u64(0),N),
 "pack_null": lambda: lib.hiqk_swap_pack(N,10,1,slots,u64(0),u64(0),u64(4),N,N),
 "unpack_null": lambda: lib.hiqk_swap_unpack(N,10,1,slots,u64(0),u64(0),u64(4),N,N),
 "p2p_null": lambda: lib.hiqk_swap_p2p(N,N,1,10,1,slots,N,u64(0),N,N,N),
 "move_null": lambda: lib.hiqk_swap_move(N,10,1,slots,1,N,u64(0),u64(4),N,1,N),
 "axpy_null": lambda: lib.hiqk_axpy_masked(N,N,10,u64(0),u64(0),dbl(1),dbl(0),N),
 "pauli_expect_null": lambda: lib.hiqk_pauli_expect(N,10,u64(0),N,1,N,u64(0),u64(4),N,N,N),
 "pauli_apply_null": lambda: lib.hiqk_pauli_apply(N,10,u64(0),N,1,N,0,N,u64(0),u64(4),N),
 "permute_null": lambda: lib.hiqk_permute_gather(N,N,1,0,10,N,N),
 "microbench_null": lambda: lib.hiqk_microbench(0,1,N),
 "dense_image_null": lambda: lib.hiqk_dense_image(10,2,N,N,u64(0),0,N,ctypes.c_size_t(0)),
 "diag_image_null": lambda: lib.hiqk_diag_batch_image(10,N,1,N,ctypes.c_size_t(0)),
 "prediag_image_null": lambda: lib.hiqk_dense_prediag_image(12,2,slots,m,N,1,N,ctypes.c_size_t(0)),
 "monomial_null": lambda: lib.hiqk_dense_is_monomial(2,N) == 0 and 1,
 "pick_variant_null": lambda: (lib.hiqk_dense_pick_variant(10,2,N), 1)[1],
 "prediag_supported_null": lambda: lib.hiqk_dense_prediag_supported(10,2,N) == 0 and 1,
 "block_shape_null": lambda: lib.hiqk_dense_block_shape(2,N,N,N),
}
predicates = ("tile_fits_null", "monomial_null", "pick_variant_null", "prediag_supported_null")
for name, f in calls.items():
    print("try", name, flush=True)
    rc = f()
    assert rc != 0, name
    if name not in predicates:
        assert lib.hiq_last_error(), name
print("PROBE_OK", len(calls))
'''


def test_device_layer_refuses_null_arguments_before_any_device_call():
    """hiqk_* launchers and host helpers: null slab / slots / matrix / op arrays are refused with a status code before anything
    touches the device (runs without a GPU; own process, a crash would show as a signal)"""
    import subprocess
    import sys
    from hiqsimulator_b200 import _lib
    res = subprocess.run([sys.executable, "-c", DEVICE_LAYER_NULL_PROBE, _lib.LIB_PATH], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, (res.returncode, res.stdout[-600:], res.stderr[-1500:])
    assert "PROBE_OK" in res.stdout
