"""Edge cases of the engine API on the CPU: the numpy oracle against the unmodified compiled reference on empty and
degenerate inputs (empty registers and id lists, scalar "gates" on zero targets, unknown ids, a full register, release and
re-allocation) — outputs, error texts, slot maps and states must agree.  The GPU engine is compared with this oracle.
(`swap_qubits([])` is left out on purpose: the reference never returns from it — SwapperMT waits for pieces that are never
produced — while this engine and the oracle treat it as the no-op it is; the reference's scheduler never issues one,
_greedyscheduler.py:185-200.)"""
import numpy as np
import pytest

import scripts
from oracle import ref, statevec

H = (np.array([[1, 1], [1, -1]]) / np.sqrt(2)).tolist()
PH = [[complex(np.exp(0.3j))]]
NQ = 5
PRE = [("ctor", 3, 5, 3)]
ALLOC = [("allocate_qureg", list(range(NQ)), 0), ("apply_controlled_gate", H, [0], []), ("run",)]
CASES = {
    "empty_register": PRE + [("allocate_qureg", [], 0), ("get_qubits_ids",), ("cheat_local",)],
    "measure_nothing_before_any_qubit": PRE + [("measure_qubits", [])],
    "probability_of_nothing_before_any_qubit": PRE + [("get_probability", [], [])],
    "run_with_nothing_queued": PRE + ALLOC + [("run",), ("run",), ("cheat_local",)],
    "probability_of_nothing": PRE + ALLOC + [("get_probability", [], [])],
    "measure_nothing": PRE + ALLOC + [("measure_qubits", []), ("cheat_local",)],
    "collapse_nothing": PRE + ALLOC + [("collapse_wavefunction", [], []), ("cheat_local",)],
    "scalar_gate": PRE + ALLOC + [("apply_controlled_gate", PH, [], []), ("run",), ("cheat_local",)],
    "scalar_gate_with_control": PRE + ALLOC + [("apply_controlled_gate", PH, [], [1]), ("run",), ("cheat_local",)],
    "scalar_gate_on_the_single_amplitude": PRE + [("apply_controlled_gate", PH, [], []), ("run",), ("cheat_local",)],
    "scalar_gate_on_one_qubit": PRE + [("allocate_qubit", 0), ("apply_controlled_gate", PH, [], []), ("run",), ("cheat_local",)],
    "unknown_id_probability": PRE + ALLOC + [("get_probability", [True], [9])],
    "unknown_id_measure": PRE + ALLOC + [("measure_qubits", [9])],
    "unknown_id_deallocate": PRE + ALLOC + [("deallocate_qubit", 9)],
    "register_full": PRE + ALLOC + [("allocate_qubit", 7), ("get_qubits_ids",)],
    "entropy_of_one_superposed_qubit": PRE + ALLOC + [("entropy",)],
    "release_and_reallocate": PRE + ALLOC + [("measure_qubits", [0, 1]), ("deallocate_qubit", 1), ("get_qubits_ids",), ("cheat_local",),
                                             ("allocate_qubit", 1), ("get_qubits_ids",), ("cheat_local",)],
    "amplitude_needs_every_qubit": PRE + ALLOC + [("get_amplitude", [False] * 4, [0, 1, 2, 3]), ("get_amplitude", [False] * 5, [0, 1, 2, 3, 3]),
                                                  ("get_amplitude", [False] * 5, [4, 3, 2, 1, 0])],
    "collapse_on_an_impossible_outcome": PRE + ALLOC + [("collapse_wavefunction", [1], [True])],
    "release_a_superposed_qubit": PRE + ALLOC + [("deallocate_qubit", 0)],
}


@pytest.mark.skipif(not ref.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", sorted(CASES))
def test_oracle_matches_the_compiled_reference_on_edge_cases(name):
    script = CASES[name]
    exp = scripts.merge_rank_outputs(ref.run_script(script, 1, 1, timeout=60))
    got = scripts.run_on_oracle(script, 1)
    scripts.assert_outputs_match(script, got, exp)


def test_empty_swap_is_a_no_op_here():
    from hiqsimulator_b200 import _cppsim_mpi as M
    o = statevec.SimulatorMPI(3, 5, 3, 1)
    o.allocate_qureg(list(range(5)), 0)
    before = o.get_qubits_ids()
    o.swap_qubits([])
    assert o.get_qubits_ids() == before
    M.init_world(0, 2, b"", 0, M.FLAG_DRY_RUN)
    e = M.SimulatorMPI(3, 5, 3)
    M.init_world(0, 1, b"", 0, 0)
    e.allocate_qureg(list(range(6)), 0)
    ids = e.get_qubits_ids()
    n = len(e.launch_trace())
    e.swap_qubits([])
    assert e.get_qubits_ids() == ids
    assert all(d["kind"] != scripts.KIND["swap"] or len(d["aux"]) == 0 for d in e.launch_trace()[n:])


@pytest.mark.skipif(not ref.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("name", ["empty_register", "register_full", "unknown_id_probability", "unknown_id_deallocate", "run_with_nothing_queued",
                                  "scalar_gate", "scalar_gate_with_control", "scalar_gate_on_the_single_amplitude", "scalar_gate_on_one_qubit"])
def test_engine_host_logic_on_edge_cases(name):
    """the same scripts on a dry-run engine (host logic of the product, no device): slot maps and error/no-error outcomes equal the
    compiled reference's, and the recorded launches replay to the reference's state"""
    from hiqsimulator_b200 import _cppsim_mpi as M
    script = CASES[name]
    exp = scripts.merge_rank_outputs(ref.run_script(script, 1, 1, timeout=60))
    M.init_world(0, 1, b"", 0, M.FLAG_DRY_RUN)
    e = M.SimulatorMPI(*script[0][1:])
    M.init_world(0, 1, b"", 0, 0)
    for op, want in zip(script[1:], exp[1:]):
        if op[0] == "cheat_local":
            e.synchronize()
            state = scripts.replay_traces([e.launch_trace()], 1, {"tile_emulator": True})
            assert np.abs(state - want[1]).max() <= 1e-12
            assert e.get_qubits_ids() == [q for q, _ in sorted(want[0].items(), key=lambda kv: kv[1])]
            continue
        try:
            got = getattr(e, op[0])(*op[1:])
        except RuntimeError as err:
            got = ("error", str(err))
        if isinstance(want, tuple) and want and want[0] == "error":
            assert isinstance(got, tuple) and got[0] == "error", (op, got)
            assert want[1].split("(")[0] in got[1], (want, got)  # same reference function name in the message
        elif op[0] == "get_qubits_ids":
            assert list(got) == list(want)
        else:
            assert not (isinstance(got, tuple) and got and got[0] == "error"), (op, got)


# ------------------------------------------------------------------ several ranks: global qubits
X = [[0, 1], [1, 0]]
T = [[1, 0], [0, complex(np.exp(0.25j * np.pi))]]


def _multi_rank_cases(R):
    g = R.bit_length() - 1
    nq = 4 + g  # 4 local slots at most, 2 at least: qubits 2 .. 2 + g - 1 are allocated global
    pre = [("ctor", 5, 4, 2), ("allocate_qureg", list(range(nq)), 0), ("get_local_qubits_ids",), ("get_global_qubits_ids",)]
    return {
        "maps_after_alloc": pre + [("get_qubits_ids",)],
        "x_on_local_ctrl_global": pre + [("apply_controlled_gate", X, [0], [2]), ("run",), ("cheat_local",)],
        "diag_on_global": pre + [("apply_controlled_gate", H, [0], []), ("run",), ("apply_controlled_gate", T, [2], [0]), ("run",), ("cheat_local",)],
        "nondiag_on_global_raises": pre + [("apply_controlled_gate", X, [2], []), ("run",)],
        "swap_then_gate": pre + [("apply_controlled_gate", H, [0], []), ("run",), ("swap_qubits", [2, 0]), ("get_qubits_ids",),
                                 ("apply_controlled_gate", H, [2], []), ("run",), ("cheat_local",)],
        "dealloc_global_in_state_1": pre + [("swap_qubits", [2, 0]), ("apply_controlled_gate", X, [2], []), ("run",), ("swap_qubits", [0, 2]),
                                            ("get_qubits_ids",), ("deallocate_qubit", 2), ("get_qubits_ids",), ("cheat_local",)],
        "dealloc_global_in_state_0": pre + [("deallocate_qubit", 2), ("get_qubits_ids",), ("cheat_local",)],
        "dealloc_local_when_few_locals": pre + [("deallocate_qubit", 0), ("get_qubits_ids",), ("deallocate_qubit", 1), ("get_qubits_ids",),
                                                ("cheat_local",)],
        "realloc_after_global_freed": pre + [("deallocate_qubit", 2), ("allocate_qubit", 9), ("get_qubits_ids",), ("cheat_local",)],
        "measure_global": pre + [("apply_controlled_gate", H, [0], []), ("run",), ("swap_qubits", [2, 0]), ("measure_qubits", [0, 1]),
                                 ("get_probability", [False], [0]), ("cheat_local",)],
        "collapse_global": pre + [("apply_controlled_gate", H, [0], []), ("run",), ("swap_qubits", [2, 0]), ("collapse_wavefunction", [0], [True]),
                                  ("cheat_local",)],
        "probability_mixed": pre + [("apply_controlled_gate", H, [0], []), ("apply_controlled_gate", H, [1], []), ("run",), ("swap_qubits", [2, 0]),
                                    ("get_probability", [True, False, False], [0, 1, 2]),
                                    ("get_amplitude", [True] + [False] * (nq - 1), list(range(nq)))],
        "swap_unknown_pair": pre + [("swap_qubits", [0, 2])],
        "swap_duplicate": pre + [("swap_qubits", [2, 0, 2, 1])],
        "entropy": pre + [("apply_controlled_gate", H, [0], []), ("run",), ("swap_qubits", [2, 0]), ("entropy",)],
    }


_MR = [(R, name) for R in (2, 4) for name in sorted(_multi_rank_cases(R))]


@pytest.mark.skipif(not ref.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("R,name", _MR)
def test_oracle_matches_the_compiled_reference_on_global_qubit_cases(R, name):
    script = _multi_rank_cases(R)[name]
    exp = scripts.merge_rank_outputs(ref.run_script(script, R, 1, timeout=60))
    got = scripts.run_on_oracle(script, R)
    scripts.assert_outputs_match(script, got, exp)


@pytest.mark.skipif(not ref.have_ref(), reason="oracle/_ref not built")
@pytest.mark.parametrize("R,name", [(R, n) for R, n in _MR if n in ("maps_after_alloc", "x_on_local_ctrl_global", "diag_on_global", "swap_then_gate",
                                                                      "nondiag_on_global_raises", "swap_unknown_pair", "swap_duplicate")])
def test_engine_host_logic_on_global_qubit_cases(R, name):
    """the product's host logic (one dry-run engine per rank): slot maps, error outcomes, and the launches — global-control filter,
    per-rank slices of diagonal gates on global qubits, swap plans — replayed to the compiled reference's state"""
    from hiqsimulator_b200 import _cppsim_mpi as M
    script = _multi_rank_cases(R)[name]
    exp = scripts.merge_rank_outputs(ref.run_script(script, R, 1, timeout=60))
    engines = []
    for r in range(R):
        M.init_world(r, R, b"", 0, M.FLAG_DRY_RUN)
        engines.append(M.SimulatorMPI(*script[0][1:]))
    M.init_world(0, 1, b"", 0, 0)
    for op, want in zip(script[1:], exp[1:]):
        if op[0] == "cheat_local":
            for e in engines:
                e.synchronize()
            state = scripts.replay_traces([e.launch_trace() for e in engines], R, {"tile_emulator": True})
            assert np.abs(state - want[1]).max() <= 1e-12
            continue
        for e in engines:
            try:
                got = getattr(e, op[0])(*op[1:])
            except RuntimeError as err:
                got = ("error", str(err))
            if isinstance(want, tuple) and want and want[0] == "error":
                assert isinstance(got, tuple) and got[0] == "error", (op, got)
            elif op[0] in ("get_qubits_ids", "get_local_qubits_ids", "get_global_qubits_ids"):
                assert list(got) == list(want)
            else:
                assert not (isinstance(got, tuple) and got and got[0] == "error"), (op, got)
