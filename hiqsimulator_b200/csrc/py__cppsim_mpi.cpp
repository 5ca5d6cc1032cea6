// pybind11 module `_cppsim_mpi`: the operator API that the ProjectQ/HiQ `SimulatorMPI` backend
// drives, re-hosted on the B200 engine through the C ABI in include/hiq_b200.h.
//
// Same class name, method names, argument order and error type as the reference binding
// (reference: _cppsim_mpi.cpp:61-83).  Differences, all additive:
//   * the MPI world is replaced by one process per GPU; `init_world(rank, world_size, nccl_id,
//     device)` plays the role of MPI_Init (world size 1 needs no call);
//   * `cheat_local()` returns the local slab as a numpy complex128 array instead of a Python list;
//   * the methods the reference wrapper calls but the reference binding never exported
//     (`get_expectation_value`, `apply_qubit_operator`, `set_wavefunction`), a working `emulate_math`
//     (the reference's throws) with closed forms of ProjectQ's math gates, and `cheat()`;
//   * extra helpers: `synchronize`, `stats`, `local_slab_ptr`, descriptor trace accessors.
#include <pybind11/complex.h>
#include <pybind11/numpy.h>
#include <pybind11/pybind11.h>
#include <pybind11/stl.h>

#include <complex>
#include <map>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/hiq_b200.h"

namespace py = pybind11;
using cplx = std::complex<double>;

namespace {

struct World {
     int rank = 0, size = 1, device = 0, flags = 0;
     std::string nccl_id;  // 128 bytes when size > 1
} g_world;

void check(int rc)
{
     if (rc != HIQ_OK) throw std::runtime_error(hiq_last_error());
}

class SimulatorB200 {
public:
     SimulatorB200(uint64_t seed, int max_local, int max_cluster)
     {
          check(hiq_create(seed, max_local, max_cluster, g_world.rank, g_world.size,
                           g_world.size > 1 ? g_world.nccl_id.data() : nullptr, g_world.device, g_world.flags, &e_));
     }
     // explicit world: used for dry-run (descriptor trace) engines of any rank without touching the
     // process-wide world set by init_world
     SimulatorB200(uint64_t seed, int max_local, int max_cluster, int rank, int world_size, int flags)
     {
          if (!(flags & HIQ_FLAG_DRY_RUN) && world_size > 1)
               throw std::runtime_error("SimulatorMPI(rank, world_size, flags): only dry-run engines may name their own world");
          check(hiq_create(seed, max_local, max_cluster, rank, world_size, nullptr, g_world.device, flags, &e_));
     }
     ~SimulatorB200() { hiq_destroy(e_); }
     SimulatorB200(const SimulatorB200&) = delete;

     std::vector<int64_t> ids(int kind) const
     {
          int n = 0;
          check(hiq_get_qubits_ids(e_, kind, nullptr, 0, &n));
          std::vector<int64_t> v(n);
          check(hiq_get_qubits_ids(e_, kind, v.data(), n, &n));
          return v;
     }
     std::vector<int64_t> get_qubits_ids() const { return ids(0); }
     std::vector<int64_t> get_local_qubits_ids() const { return ids(1); }
     std::vector<int64_t> get_global_qubits_ids() const { return ids(2); }
     void set_qubits_perm(const std::vector<int64_t>& p) { check(hiq_set_qubits_perm(e_, p.data(), static_cast<int>(p.size()))); }
     void swap_qubits(const std::vector<int64_t>& pairs)
     {
          py::gil_scoped_release nogil;
          check(hiq_swap_qubits(e_, pairs.data(), static_cast<int>(pairs.size())));
     }
     void allocate_qureg(const std::vector<int64_t>& q, cplx init)
     {
          check(hiq_allocate_qureg(e_, q.data(), static_cast<int>(q.size()), init.real(), init.imag()));
     }
     void allocate_qubit(int64_t id) { check(hiq_allocate_qubit(e_, id)); }
     void deallocate_qubit(int64_t id) { check(hiq_deallocate_qubit(e_, id)); }
     std::vector<bool> measure_qubits(const std::vector<int64_t>& q)
     {
          std::vector<uint8_t> out(q.size());
          check(hiq_measure_qubits(e_, q.data(), static_cast<int>(q.size()), out.data()));
          return std::vector<bool>(out.begin(), out.end());
     }
     void apply_controlled_gate(const std::vector<std::vector<cplx>>& m, const std::vector<int64_t>& q,
                                const std::vector<int64_t>& ctrls)
     {
          const int dim = static_cast<int>(m.size());
          std::vector<cplx> flat;
          flat.reserve(static_cast<size_t>(dim) * dim);
          for (auto& row: m) {
               if (static_cast<int>(row.size()) != dim) throw std::runtime_error("apply_controlled_gate: matrix must be square");
               flat.insert(flat.end(), row.begin(), row.end());
          }
          check(hiq_apply_controlled_gate(e_, reinterpret_cast<const double*>(flat.data()), dim, q.data(),
                                          static_cast<int>(q.size()), ctrls.data(), static_cast<int>(ctrls.size())));
     }
     // fast path: numpy matrix, no list-of-lists conversion
     void apply_controlled_matrix(py::array_t<cplx, py::array::c_style | py::array::forcecast> m, const std::vector<int64_t>& q,
                                  const std::vector<int64_t>& ctrls)
     {
          if (m.ndim() != 2 || m.shape(0) != m.shape(1)) throw std::runtime_error("apply_controlled_matrix: matrix must be square");
          check(hiq_apply_controlled_gate(e_, reinterpret_cast<const double*>(m.data()), static_cast<int>(m.shape(0)), q.data(),
                                          static_cast<int>(q.size()), ctrls.data(), static_cast<int>(ctrls.size())));
     }
     // emulate_math(f, quregs, ctrls) (reference binding: _cppsim_mpi.cpp:42-59,75; the reference engine throws,
     // SimulatorMPI.hpp:217-225).  f maps the list of register values to the list of new values (ProjectQ's
     // BasicMathGate.get_math_function); it is tabulated once on the host — the device applies the permutation.
     void emulate_math(py::function f, const std::vector<std::vector<int64_t>>& quregs, const std::vector<int64_t>& ctrls)
     {
          std::vector<int64_t> reg_ids;
          for (auto& qr: quregs) reg_ids.insert(reg_ids.end(), qr.begin(), qr.end());
          if (reg_ids.empty() || reg_ids.size() > 22)
               throw std::runtime_error("emulate_math(): a Python function can be tabulated over at most 22 register qubits; "
                                        "use emulate_math_add_constant / _add_constant_modN / _multiply_by_constant_modN");
          const uint64_t space = 1ull << reg_ids.size();
          std::vector<uint64_t> table(space);
          for (uint64_t v = 0; v < space; ++v) {
               py::list args;
               unsigned shift = 0;
               for (auto& qr: quregs) {
                    args.append(py::int_((v >> shift) & ((1ull << qr.size()) - 1ull)));
                    shift += static_cast<unsigned>(qr.size());
               }
               py::object res = f(args);
               py::sequence out = py::reinterpret_borrow<py::sequence>(res);
               if (static_cast<size_t>(py::len(out)) != quregs.size())
                    throw std::runtime_error("emulate_math(): the function must return one value per register");
               uint64_t w = 0;
               shift = 0;
               for (size_t r = 0; r < quregs.size(); ++r) {
                    // Python integers of any sign and size: keep the low bits of the two's complement
                    const py::int_ mask((1ull << quregs[r].size()) - 1ull);
                    const uint64_t val = py::cast<uint64_t>(py::int_(out[r]).attr("__and__")(mask));
                    w |= val << shift;
                    shift += static_cast<unsigned>(quregs[r].size());
               }
               table[v] = w;
          }
          py::gil_scoped_release nogil;
          check(hiq_emulate_math_table(e_, table.data(), space, reg_ids.data(), static_cast<int>(reg_ids.size()), ctrls.data(),
                                       static_cast<int>(ctrls.size())));
     }
     // ProjectQ's math gates in closed form (projectq.libs.math: AddConstant, AddConstantModN, MultiplyByConstantModN)
     void emulate_math_add_constant(int64_t a, const std::vector<int64_t>& qureg, const std::vector<int64_t>& ctrls)
     {
          check(hiq_emulate_math_const(e_, HIQK_PERM_ADD, static_cast<uint64_t>(a), 0, qureg.data(), static_cast<int>(qureg.size()),
                                       ctrls.data(), static_cast<int>(ctrls.size())));
     }
     void emulate_math_add_constant_modN(int64_t a, uint64_t N, const std::vector<int64_t>& qureg, const std::vector<int64_t>& ctrls)
     {
          if (N == 0) throw std::runtime_error("emulate_math(): the modulus must be positive");
          const int64_t n = static_cast<int64_t>(N);
          const uint64_t ar = static_cast<uint64_t>(((a % n) + n) % n);
          check(hiq_emulate_math_const(e_, HIQK_PERM_ADD_MOD, ar, N, qureg.data(), static_cast<int>(qureg.size()), ctrls.data(),
                                       static_cast<int>(ctrls.size())));
     }
     void emulate_math_multiply_by_constant_modN(uint64_t a, uint64_t N, const std::vector<int64_t>& qureg,
                                                 const std::vector<int64_t>& ctrls)
     {
          check(hiq_emulate_math_const(e_, HIQK_PERM_MUL_MOD, a, N, qureg.data(), static_cast<int>(qureg.size()), ctrls.data(),
                                       static_cast<int>(ctrls.size())));
     }
     using Term = std::vector<std::pair<int, char>>;
     using TermsDict = std::vector<std::pair<Term, cplx>>;
     struct FlatTerms {
          std::vector<int> offsets{0}, index;
          std::vector<char> pauli;
          std::vector<double> coefs;
          explicit FlatTerms(const TermsDict& td)
          {
               for (auto& t: td) {
                    for (auto& f: t.first) {
                         index.push_back(f.first);
                         pauli.push_back(f.second);
                    }
                    offsets.push_back(static_cast<int>(index.size()));
                    coefs.push_back(t.second.real());
                    coefs.push_back(t.second.imag());
               }
          }
     };
     // the wrapper's call shapes (reference: _simulator_mpi.py:180-183, 220-223, 304-305)
     double get_expectation_value(const TermsDict& td, const std::vector<int64_t>& ids)
     {
          FlatTerms f(td);
          double out = 0.0;
          py::gil_scoped_release nogil;
          check(hiq_get_expectation_value(e_, f.offsets.data(), f.index.data(), f.pauli.data(), f.coefs.data(), static_cast<int>(td.size()),
                                          ids.data(), static_cast<int>(ids.size()), &out));
          return out;
     }
     void apply_qubit_operator(const TermsDict& td, const std::vector<int64_t>& ids)
     {
          FlatTerms f(td);
          py::gil_scoped_release nogil;
          check(hiq_apply_qubit_operator(e_, f.offsets.data(), f.index.data(), f.pauli.data(), f.coefs.data(), static_cast<int>(td.size()),
                                         ids.data(), static_cast<int>(ids.size())));
     }
     void emulate_time_evolution(const TermsDict& td, double time, const std::vector<int64_t>& ids, const std::vector<int64_t>& ctrls)
     {
          FlatTerms f(td);
          py::gil_scoped_release nogil;
          check(hiq_emulate_time_evolution(e_, f.offsets.data(), f.index.data(), f.pauli.data(), f.coefs.data(), static_cast<int>(td.size()),
                                           time, ids.data(), static_cast<int>(ids.size()), ctrls.data(), static_cast<int>(ctrls.size())));
     }
     void set_wavefunction(py::array_t<cplx, py::array::c_style | py::array::forcecast> wf, const std::vector<int64_t>& ids)
     {
          check(hiq_set_wavefunction(e_, reinterpret_cast<const double*>(wf.data()), static_cast<uint64_t>(wf.size()), ids.data(),
                                     static_cast<int>(ids.size())));
     }
     // cheat(): (id -> bit position, the concatenation of every rank's slab) on every rank
     // (reference: _simulator_mpi.py:348-380 does this with mpi4py on top of cheat_local)
     py::tuple cheat()
     {
          int n_map = 0;
          uint64_t n_amps = 0;
          check(hiq_cheat(e_, nullptr, nullptr, 0, &n_map, nullptr, 0, &n_amps));
          std::vector<int64_t> idv(n_map);
          std::vector<int> posv(n_map);
          py::array_t<cplx> vec(static_cast<py::ssize_t>(n_amps));
          check(hiq_cheat(e_, idv.data(), posv.data(), n_map, &n_map, vec.mutable_data(), n_amps, &n_amps));
          py::dict d;
          for (int i = 0; i < n_map; ++i) d[py::int_(idv[i])] = posv[i];
          return py::make_tuple(d, vec);
     }
     cplx get_amplitude(const std::vector<bool>& bits, const std::vector<int64_t>& q)
     {
          if (bits.size() != q.size()) throw std::runtime_error("GetAmplitude(): ids.size() != number of qubits");
          std::vector<uint8_t> b(bits.begin(), bits.end());
          double out[2];
          check(hiq_get_amplitude(e_, b.data(), q.data(), static_cast<int>(q.size()), out));
          return cplx(out[0], out[1]);
     }
     double get_probability(const std::vector<bool>& bits, const std::vector<int64_t>& q)
     {
          if (bits.size() != q.size()) throw std::runtime_error("GetProbability(): ids.size() != bit_string.size()");
          std::vector<uint8_t> b(bits.begin(), bits.end());
          double out = 0.0;
          check(hiq_get_probability(e_, b.data(), q.data(), static_cast<int>(q.size()), &out));
          return out;
     }
     void run() { check(hiq_run(e_)); }
     double entropy()
     {
          double out = 0.0;
          check(hiq_entropy(e_, &out));
          return out;
     }
     py::tuple cheat_local()
     {
          int n_map = 0;
          uint64_t n_amps = 0;
          check(hiq_cheat_local(e_, nullptr, nullptr, 0, &n_map, nullptr, 0, &n_amps));
          std::vector<int64_t> idv(n_map);
          std::vector<int> posv(n_map);
          py::array_t<cplx> vec(static_cast<py::ssize_t>(n_amps));
          check(hiq_cheat_local(e_, idv.data(), posv.data(), n_map, &n_map, vec.mutable_data(), n_amps, &n_amps));
          py::dict d;
          for (int i = 0; i < n_map; ++i) d[py::int_(idv[i])] = posv[i];
          return py::make_tuple(d, vec);
     }
     void collapse_wavefunction(const std::vector<int64_t>& q, const std::vector<bool>& values)
     {
          if (values.size() != q.size()) throw std::runtime_error("collapseWaveFunction(): ids.size() != values.size()");
          std::vector<uint8_t> b(values.begin(), values.end());
          check(hiq_collapse_wavefunction(e_, q.data(), b.data(), static_cast<int>(q.size())));
     }

     // ---- additions
     void synchronize() { check(hiq_synchronize(e_)); }
     void set_dense_variant(int v) { check(hiq_set_dense_variant(e_, v)); }
     void set_local_slab(py::array_t<cplx, py::array::c_style | py::array::forcecast> a)
     {
          check(hiq_set_local_slab(e_, a.data(), static_cast<uint64_t>(a.size())));
     }
     py::tuple local_slab_ptr()
     {
          void* p = nullptr;
          int L = 0;
          check(hiq_local_slab(e_, &p, &L));
          return py::make_tuple(reinterpret_cast<uintptr_t>(p), L);
     }
     py::dict stats()
     {
          hiq_stats s;
          check(hiq_get_stats(e_, &s));
          py::dict d;
          d["total_gates"] = s.total_gates;
          d["total_runs"] = s.total_runs;
          d["total_stages"] = s.total_stages;
          d["total_swaps"] = s.total_swaps;
          d["dense_passes"] = s.dense_passes;
          d["diag_passes"] = s.diag_passes;
          d["scale_passes"] = s.scale_passes;
          d["skipped_passes"] = s.skipped_passes;
          d["runs_s"] = s.runs_s;
          d["swaps_s"] = s.swaps_s;
          d["measures_s"] = s.measures_s;
          d["allocs_s"] = s.allocs_s;
          d["deallocs_s"] = s.deallocs_s;
          d["swap_bytes_sent"] = s.swap_bytes_sent;
          d["swaps_p2p"] = s.swaps_p2p;
          d["swaps_staged"] = s.swaps_staged;
          d["swaps_packed"] = s.swaps_packed;
          d["h2d_bytes"] = s.h2d_bytes;
          d["d2h_bytes"] = s.d2h_bytes;
          d["gate_launches"] = s.gate_launches;
          d["ctor_s"] = s.ctor_s;
          d["slab_grow_s"] = s.slab_grow_s;
          d["peer_map_s"] = s.peer_map_s;
          d["tile_launches"] = s.tile_launches;
          d["tile_steps"] = s.tile_steps;
          return d;
     }
     py::list trace() { return trace_list(hiq_trace_count, hiq_trace_get); }
     // dry-run engines: the launches behind the plan (diagonal folding and tile runs applied), see include/hiq_b200.h
     py::list launch_trace() { return trace_list(hiq_launch_trace_count, hiq_launch_trace_get); }
     py::list trace_list(int (*count)(hiq_engine*, int*), int (*get)(hiq_engine*, int, hiq_descriptor*, double*, int, int64_t*, int))
     {
          int n = 0;
          check(count(e_, &n));
          py::list out;
          for (int i = 0; i < n; ++i) {
               hiq_descriptor d;
               check(get(e_, i, &d, nullptr, 0, nullptr, 0));
               py::array_t<cplx> payload(d.n_payload);
               std::vector<int64_t> aux(d.n_aux);
               check(get(e_, i, &d, reinterpret_cast<double*>(payload.mutable_data()), d.n_payload, aux.data(), d.n_aux));
               py::dict r;
               r["kind"] = d.kind;
               r["k"] = d.k;
               r["slots"] = std::vector<int>(d.slots, d.slots + (d.kind == HIQ_DESC_DENSE || d.kind == HIQ_DESC_DIAG ? d.k : 0));
               if (d.kind == HIQ_DESC_LAUNCH) r["form"] = d.slots[0];
               r["ctrl_mask"] = d.ctrl_mask;
               r["payload"] = payload;
               r["aux"] = aux;
               out.append(r);
          }
          return out;
     }
     void clear_trace() { check(hiq_trace_clear(e_)); }
     py::list collect_timings()
     {
          const int cap = 1 << 16;
          std::vector<double> ms(cap);
          std::vector<int> kind(cap), k(cap), variant(cap), n_ref(cap);
          int n = 0;
          check(hiq_collect_timings(e_, ms.data(), kind.data(), k.data(), variant.data(), n_ref.data(), cap, &n));
          py::list out;
          for (int i = 0; i < n; ++i) out.append(py::make_tuple(kind[i], k[i], variant[i], ms[i], n_ref[i]));
          return out;
     }
     uintptr_t stream_ptr()
     {
          void* s = nullptr;
          check(hiq_stream(e_, &s));
          return reinterpret_cast<uintptr_t>(s);
     }

private:
     hiq_engine* e_ = nullptr;
};

}  // namespace

PYBIND11_MODULE(_cppsim_mpi, m)
{
     m.doc() = "B200-native drop-in for HiQsimulator's _cppsim_mpi (state-vector engine on sm_100a)";
     m.def("unique_id", [] {
          std::string id(128, '\0');
          check(hiq_comm_unique_id(id.data()));
          return py::bytes(id);
     }, "NCCL unique id (call on rank 0, hand to every rank's init_world)");
     m.def("init_world", [](int rank, int world_size, py::bytes nccl_id, int device, int flags) {
          g_world.rank = rank;
          g_world.size = world_size;
          g_world.nccl_id = std::string(nccl_id);
          g_world.device = device;
          g_world.flags = flags;
          if (world_size > 1 && !(flags & HIQ_FLAG_DRY_RUN) && g_world.nccl_id.size() != 128)
               throw std::runtime_error("init_world: nccl_id must be the 128 bytes returned by unique_id() on rank 0");
     }, py::arg("rank") = 0, py::arg("world_size") = 1, py::arg("nccl_id") = py::bytes(), py::arg("device") = 0, py::arg("flags") = 0);
     m.def("world", [] { return py::make_tuple(g_world.rank, g_world.size, g_world.device, g_world.flags); });
     m.def("device_count", &hiq_device_count);
     m.def("version", [] { return std::string(hiq_version()); });
     m.def("launch_count", &hiqk_launch_count);
     m.attr("FLAG_DRY_RUN") = HIQ_FLAG_DRY_RUN;
     m.attr("FLAG_TRACE") = HIQ_FLAG_TRACE;
     m.attr("FLAG_TIMING") = HIQ_FLAG_TIMING;
     m.attr("FLAG_NO_BATCH") = HIQ_FLAG_NO_BATCH;

     py::class_<SimulatorB200>(m, "SimulatorMPI")
         .def(py::init<uint64_t, int, int>())
         .def(py::init<uint64_t, int, int, int, int, int>(), py::arg("seed"), py::arg("max_local"), py::arg("max_cluster_size"),
              py::arg("rank"), py::arg("world_size"), py::arg("flags"))
         .def("get_qubits_ids", &SimulatorB200::get_qubits_ids)
         .def("get_local_qubits_ids", &SimulatorB200::get_local_qubits_ids)
         .def("get_global_qubits_ids", &SimulatorB200::get_global_qubits_ids)
         .def("set_qubits_perm", &SimulatorB200::set_qubits_perm)
         .def("swap_qubits", &SimulatorB200::swap_qubits)
         .def("allocate_qureg", &SimulatorB200::allocate_qureg, py::arg("ids"), py::arg("init") = cplx(0.0))
         .def("allocate_qubit", &SimulatorB200::allocate_qubit)
         .def("deallocate_qubit", &SimulatorB200::deallocate_qubit)
         .def("measure_qubits", &SimulatorB200::measure_qubits)
         .def("apply_controlled_gate", &SimulatorB200::apply_controlled_gate)
         .def("apply_controlled_matrix", &SimulatorB200::apply_controlled_matrix)
         .def("emulate_math", &SimulatorB200::emulate_math)
         .def("emulate_math_add_constant", &SimulatorB200::emulate_math_add_constant)
         .def("emulate_math_add_constant_modN", &SimulatorB200::emulate_math_add_constant_modN)
         .def("emulate_math_multiply_by_constant_modN", &SimulatorB200::emulate_math_multiply_by_constant_modN)
         .def("get_expectation_value", &SimulatorB200::get_expectation_value)
         .def("apply_qubit_operator", &SimulatorB200::apply_qubit_operator)
         .def("emulate_time_evolution", &SimulatorB200::emulate_time_evolution)
         .def("set_wavefunction", &SimulatorB200::set_wavefunction)
         .def("cheat", &SimulatorB200::cheat)
         .def("get_amplitude", &SimulatorB200::get_amplitude)
         .def("get_probability", &SimulatorB200::get_probability)
         .def("run", &SimulatorB200::run)
         .def("entropy", &SimulatorB200::entropy)
         .def("cheat_local", &SimulatorB200::cheat_local)
         .def("collapse_wavefunction", &SimulatorB200::collapse_wavefunction)
         .def("synchronize", &SimulatorB200::synchronize)
         .def("set_dense_variant", &SimulatorB200::set_dense_variant)
         .def("set_local_slab", &SimulatorB200::set_local_slab)
         .def("local_slab_ptr", &SimulatorB200::local_slab_ptr)
         .def("stats", &SimulatorB200::stats)
         .def("trace", &SimulatorB200::trace)
         .def("launch_trace", &SimulatorB200::launch_trace)
         .def("clear_trace", &SimulatorB200::clear_trace)
         .def("collect_timings", &SimulatorB200::collect_timings)
         .def("stream_ptr", &SimulatorB200::stream_ptr);
}
