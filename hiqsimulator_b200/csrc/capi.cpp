// extern "C" surface of the engine (include/hiq_b200.h): plain pointers and sizes in, status
// codes out; C++ exceptions never cross the boundary.
#include <cstring>

#include "engine.hpp"
#include "hiq_host.hpp"

using namespace hiq;

struct hiq_engine {
     Engine impl;
     template <class... A>
     explicit hiq_engine(A&&... a) : impl(std::forward<A>(a)...) {}
};

template <class F>
static int guarded(F&& f)
{
     try {
          f();
          return HIQ_OK;
     }
     catch (const EngineError& e) {
          return set_error(e.code, e.what());
     }
     catch (const std::exception& e) {
          return set_error(HIQ_ERR_RUNTIME, e.what());
     }
}

#define NEED(e) \
     if (!(e)) return set_error(HIQ_ERR_ARG, "null engine handle")
// an input array: a null pointer is fine for an empty list only
#define NEED_ARRAY(who, p, n) \
     if ((n) < 0 || ((n) > 0 && !(p))) return set_error(HIQ_ERR_ARG, who ": null array or negative length")
#define NEED_OUT(who, p) \
     if (!(p)) return set_error(HIQ_ERR_ARG, who ": null output")

static std::vector<Index> vec_ids(const int64_t* p, int n) { return std::vector<Index>(p, p + (n > 0 ? n : 0)); }
static std::vector<bool> vec_bits(const uint8_t* p, int n)
{
     std::vector<bool> v(n > 0 ? n : 0);
     for (int i = 0; i < n; ++i) v[i] = p[i] != 0;
     return v;
}

extern "C" {

int hiq_create(uint64_t seed, int max_local, int max_cluster_size, int rank, int world_size, const void* nccl_id, int device,
               int flags, hiq_engine** out)
{
     if (!out) return set_error(HIQ_ERR_ARG, "hiq_create: null output");
     *out = nullptr;
     return guarded([&] { *out = new hiq_engine(seed, max_local, max_cluster_size, rank, world_size, nccl_id, device, flags); });
}

int hiq_destroy(hiq_engine* e)
{
     delete e;
     return HIQ_OK;
}

int hiq_allocate_qubit(hiq_engine* e, int64_t id)
{
     NEED(e);
     return guarded([&] { e->impl.allocate_qubit(id); });
}

int hiq_allocate_qureg(hiq_engine* e, const int64_t* ids, int n, double init_re, double init_im)
{
     NEED(e);
     NEED_ARRAY("hiq_allocate_qureg", ids, n);
     return guarded([&] { e->impl.allocate_qureg(vec_ids(ids, n), cplx(init_re, init_im)); });
}

int hiq_deallocate_qubit(hiq_engine* e, int64_t id)
{
     NEED(e);
     return guarded([&] { e->impl.deallocate_qubit(id); });
}

int hiq_apply_controlled_gate(hiq_engine* e, const double* matrix, int dim, const int64_t* ids, int n_ids, const int64_t* ctrls,
                              int n_ctrls)
{
     NEED(e);
     NEED_ARRAY("hiq_apply_controlled_gate", ids, n_ids);
     NEED_ARRAY("hiq_apply_controlled_gate", ctrls, n_ctrls);
     if (!matrix || dim < 1 || dim > 32) return set_error(HIQ_ERR_ARG, "hiq_apply_controlled_gate: bad matrix");
     return guarded([&] {
          GateMatrix m(dim);
          std::memcpy(static_cast<void*>(m.a.data()), matrix, sizeof(cplx) * dim * dim);
          e->impl.apply_gate(std::move(m), vec_ids(ids, n_ids), vec_ids(ctrls, n_ctrls));
     });
}

int hiq_run(hiq_engine* e)
{
     NEED(e);
     return guarded([&] { e->impl.run(); });
}

int hiq_swap_qubits(hiq_engine* e, const int64_t* pairs, int n)
{
     NEED(e);
     NEED_ARRAY("hiq_swap_qubits", pairs, n);
     return guarded([&] { e->impl.swap_qubits_stage(vec_ids(pairs, n)); });
}

int hiq_measure_qubits(hiq_engine* e, const int64_t* ids, int n, uint8_t* out_bits)
{
     NEED(e);
     NEED_ARRAY("hiq_measure_qubits", ids, n);
     if (n > 0) NEED_OUT("hiq_measure_qubits", out_bits);
     return guarded([&] {
          auto r = e->impl.measure_qubits(vec_ids(ids, n));
          for (int i = 0; i < n; ++i) out_bits[i] = r[i] ? 1 : 0;
     });
}

int hiq_get_probability(hiq_engine* e, const uint8_t* bits, const int64_t* ids, int n, double* out)
{
     NEED(e);
     NEED_ARRAY("hiq_get_probability", ids, n);
     NEED_ARRAY("hiq_get_probability", bits, n);
     NEED_OUT("hiq_get_probability", out);
     return guarded([&] { *out = e->impl.get_probability(vec_bits(bits, n), vec_ids(ids, n)); });
}

int hiq_get_amplitude(hiq_engine* e, const uint8_t* bits, const int64_t* ids, int n, double* out_re_im)
{
     NEED(e);
     NEED_ARRAY("hiq_get_amplitude", ids, n);
     NEED_ARRAY("hiq_get_amplitude", bits, n);
     NEED_OUT("hiq_get_amplitude", out_re_im);
     return guarded([&] {
          const cplx v = e->impl.get_amplitude(vec_bits(bits, n), vec_ids(ids, n));
          out_re_im[0] = v.real();
          out_re_im[1] = v.imag();
     });
}

int hiq_collapse_wavefunction(hiq_engine* e, const int64_t* ids, const uint8_t* values, int n)
{
     NEED(e);
     NEED_ARRAY("hiq_collapse_wavefunction", ids, n);
     NEED_ARRAY("hiq_collapse_wavefunction", values, n);
     return guarded([&] { e->impl.collapse_wavefunction(vec_ids(ids, n), vec_bits(values, n)); });
}

int hiq_entropy(hiq_engine* e, double* out)
{
     NEED(e);
     NEED_OUT("hiq_entropy", out);
     return guarded([&] { *out = e->impl.entropy(); });
}

int hiq_get_qubits_ids(hiq_engine* e, int kind, int64_t* out, int cap, int* n)
{
     NEED(e);
     NEED_OUT("hiq_get_qubits_ids", n);
     return guarded([&] {
          std::vector<Index> v = kind == 0 ? e->impl.qubits_permutation() : (kind == 1 ? e->impl.locals() : e->impl.globals());
          *n = static_cast<int>(v.size());
          if (out) {
               if (cap < *n) throw EngineError(HIQ_ERR_ARG, "hiq_get_qubits_ids: buffer too small");
               std::copy(v.begin(), v.end(), out);
          }
     });
}

int hiq_set_qubits_perm(hiq_engine* e, const int64_t* p, int n)
{
     NEED(e);
     NEED_ARRAY("hiq_set_qubits_perm", p, n);
     return guarded([&] { e->impl.set_qubits_permutation(vec_ids(p, n)); });
}

int hiq_cheat_local(hiq_engine* e, int64_t* ids, int* pos, int cap, int* n_map, void* host_dst, uint64_t cap_amps, uint64_t* n_amps)
{
     NEED(e);
     return guarded([&] {
          auto m = e->impl.id2pos();
          if (n_map) *n_map = static_cast<int>(m.size());
          if (ids && pos) {
               if (cap < static_cast<int>(m.size())) throw EngineError(HIQ_ERR_ARG, "hiq_cheat_local: map buffer too small");
               int i = 0;
               for (auto& kv: m) {
                    ids[i] = kv.first;
                    pos[i] = kv.second;
                    ++i;
               }
          }
          if (n_amps) *n_amps = 1ull << e->impl.local_qubits();
          if (host_dst) e->impl.copy_slab_to_host(host_dst, cap_amps);
     });
}

int hiq_cheat(hiq_engine* e, int64_t* ids, int* pos, int cap, int* n_map, void* host_dst, uint64_t cap_amps, uint64_t* n_amps)
{
     NEED(e);
     return guarded([&] {
          auto m = e->impl.id2pos();
          if (n_map) *n_map = static_cast<int>(m.size());
          if (ids && pos) {
               if (cap < static_cast<int>(m.size())) throw EngineError(HIQ_ERR_ARG, "hiq_cheat: map buffer too small");
               int i = 0;
               for (auto& kv: m) {
                    ids[i] = kv.first;
                    pos[i] = kv.second;
                    ++i;
               }
          }
          if (n_amps) *n_amps = static_cast<uint64_t>(e->impl.world_size()) << e->impl.local_qubits();
          if (host_dst) e->impl.gather_state_to_host(host_dst, cap_amps);
     });
}

static std::vector<Engine::PauliTerm> pauli_terms(const int* term_offsets, const int* factor_index, const char* factor_pauli,
                                                  const double* coefs, int n_terms)
{
     if (n_terms < 0 || (n_terms > 0 && (!term_offsets || !coefs))) throw EngineError(HIQ_ERR_ARG, "qubit operator: null argument");
     std::vector<Engine::PauliTerm> terms(n_terms);
     for (int t = 0; t < n_terms; ++t) {
          if (term_offsets[t + 1] < term_offsets[t]) throw EngineError(HIQ_ERR_ARG, "qubit operator: term offsets must ascend");
          if (term_offsets[t + 1] > term_offsets[t] && (!factor_index || !factor_pauli))
               throw EngineError(HIQ_ERR_ARG, "qubit operator: null factor arrays");
          for (int f = term_offsets[t]; f < term_offsets[t + 1]; ++f) terms[t].factors.emplace_back(factor_index[f], factor_pauli[f]);
          terms[t].coef = cplx(coefs[2 * t], coefs[2 * t + 1]);
     }
     return terms;
}

int hiq_get_expectation_value(hiq_engine* e, const int* term_offsets, const int* factor_index, const char* factor_pauli,
                              const double* coefs_re_im, int n_terms, const int64_t* ids, int n_ids, double* out)
{
     NEED(e);
     NEED_ARRAY("hiq_get_expectation_value", ids, n_ids);
     if (!out) return set_error(HIQ_ERR_ARG, "hiq_get_expectation_value: null output");
     return guarded([&] {
          *out = e->impl.get_expectation_value(pauli_terms(term_offsets, factor_index, factor_pauli, coefs_re_im, n_terms), vec_ids(ids, n_ids));
     });
}

int hiq_apply_qubit_operator(hiq_engine* e, const int* term_offsets, const int* factor_index, const char* factor_pauli,
                             const double* coefs_re_im, int n_terms, const int64_t* ids, int n_ids)
{
     NEED(e);
     NEED_ARRAY("hiq_apply_qubit_operator", ids, n_ids);
     return guarded([&] {
          e->impl.apply_qubit_operator(pauli_terms(term_offsets, factor_index, factor_pauli, coefs_re_im, n_terms), vec_ids(ids, n_ids));
     });
}

int hiq_emulate_time_evolution(hiq_engine* e, const int* term_offsets, const int* factor_index, const char* factor_pauli,
                               const double* coefs_re_im, int n_terms, double time, const int64_t* ids, int n_ids,
                               const int64_t* ctrls, int n_ctrls)
{
     NEED(e);
     NEED_ARRAY("hiq_emulate_time_evolution", ids, n_ids);
     NEED_ARRAY("hiq_emulate_time_evolution", ctrls, n_ctrls);
     return guarded([&] {
          e->impl.emulate_time_evolution(pauli_terms(term_offsets, factor_index, factor_pauli, coefs_re_im, n_terms), time,
                                         vec_ids(ids, n_ids), vec_ids(ctrls, n_ctrls));
     });
}

int hiq_set_wavefunction(hiq_engine* e, const double* amps_re_im, uint64_t n_amps, const int64_t* ids, int n_ids)
{
     NEED(e);
     NEED_ARRAY("hiq_set_wavefunction", ids, n_ids);
     if (n_amps > 0) NEED_OUT("hiq_set_wavefunction (amplitudes)", amps_re_im);
     return guarded([&] { e->impl.set_wavefunction(reinterpret_cast<const cplx*>(amps_re_im), n_amps, vec_ids(ids, n_ids)); });
}

int hiq_emulate_math_table(hiq_engine* e, const uint64_t* table, uint64_t table_len, const int64_t* reg_ids, int n_reg,
                           const int64_t* ctrls, int n_ctrls)
{
     NEED(e);
     NEED_ARRAY("hiq_emulate_math_table", reg_ids, n_reg);
     NEED_ARRAY("hiq_emulate_math_table", ctrls, n_ctrls);
     if (!table) return set_error(HIQ_ERR_ARG, "hiq_emulate_math_table: null table");
     return guarded([&] {
          e->impl.emulate_math(HIQK_PERM_TABLE, 0, 0, std::vector<uint64_t>(table, table + table_len), vec_ids(reg_ids, n_reg),
                               vec_ids(ctrls, n_ctrls));
     });
}

int hiq_emulate_math_const(hiq_engine* e, int kind, uint64_t a, uint64_t N, const int64_t* reg_ids, int n_reg, const int64_t* ctrls,
                           int n_ctrls)
{
     NEED(e);
     NEED_ARRAY("hiq_emulate_math_const", reg_ids, n_reg);
     NEED_ARRAY("hiq_emulate_math_const", ctrls, n_ctrls);
     if (kind != HIQK_PERM_ADD && kind != HIQK_PERM_ADD_MOD && kind != HIQK_PERM_MUL_MOD)
          return set_error(HIQ_ERR_ARG, "hiq_emulate_math_const: kind must be HIQK_PERM_ADD, _ADD_MOD or _MUL_MOD");
     return guarded([&] { e->impl.emulate_math(kind, a, N, {}, vec_ids(reg_ids, n_reg), vec_ids(ctrls, n_ctrls)); });
}

int hiq_local_slab(hiq_engine* e, void** dev_ptr, int* L)
{
     NEED(e);
     NEED_OUT("hiq_local_slab", dev_ptr);
     NEED_OUT("hiq_local_slab", L);
     return guarded([&] {
          if (e->impl.dry_run()) throw EngineError(HIQ_ERR_RUNTIME, "hiq_local_slab: dry-run engine has no slab");
          *dev_ptr = e->impl.slab_ptr();
          *L = e->impl.local_qubits();
     });
}

int hiq_set_local_slab(hiq_engine* e, const void* host_src, uint64_t n_amps)
{
     NEED(e);
     if (n_amps > 0) NEED_OUT("hiq_set_local_slab (source)", host_src);
     return guarded([&] { e->impl.copy_slab_from_host(host_src, n_amps); });
}

int hiq_synchronize(hiq_engine* e)
{
     NEED(e);
     return guarded([&] { e->impl.synchronize(); });
}

int hiq_rank(hiq_engine* e, int* rank, int* world_size)
{
     NEED(e);
     NEED_OUT("hiq_rank", rank);
     NEED_OUT("hiq_rank", world_size);
     *rank = e->impl.rank();
     *world_size = e->impl.world_size();
     return HIQ_OK;
}

int hiq_set_dense_variant(hiq_engine* e, int variant)
{
     NEED(e);
     e->impl.set_dense_variant(variant);
     return HIQ_OK;
}

int hiq_get_stats(hiq_engine* e, hiq_stats* out)
{
     NEED(e);
     NEED_OUT("hiq_get_stats", out);
     const EngineStats& s = e->impl.stats();
     out->total_gates = s.total_gates;
     out->total_runs = s.total_runs;
     out->total_stages = s.total_stages;
     out->total_swaps = s.total_swaps;
     out->dense_passes = s.dense_passes;
     out->diag_passes = s.diag_passes;
     out->scale_passes = s.scale_passes;
     out->skipped_passes = s.skipped_passes;
     out->runs_s = s.runs_s;
     out->swaps_s = s.swaps_s;
     out->measures_s = s.measures_s;
     out->allocs_s = s.allocs_s;
     out->deallocs_s = s.deallocs_s;
     out->swap_bytes_sent = s.swap_bytes_sent;
     out->swaps_p2p = s.swaps_p2p;
     out->swaps_staged = s.swaps_staged;
     out->swaps_packed = s.swaps_packed;
     out->h2d_bytes = s.h2d_bytes;
     out->d2h_bytes = s.d2h_bytes;
     out->gate_launches = s.gate_launches;
     out->ctor_s = s.ctor_s;
     out->slab_grow_s = s.slab_grow_s;
     out->peer_map_s = s.peer_map_s;
     out->tile_launches = s.tile_launches;
     out->tile_steps = s.tile_steps;
     return HIQ_OK;
}

int hiq_collect_timings(hiq_engine* e, double* ms, int* kind, int* k, int* variant, int* n_ref, int cap, int* n)
{
     NEED(e);
     NEED_OUT("hiq_collect_timings", n);
     if (cap > 0 && (!ms || !kind || !k || !variant)) return set_error(HIQ_ERR_ARG, "hiq_collect_timings: null output");
     return guarded([&] {
          auto t = e->impl.collect_timings();
          *n = static_cast<int>(t.size());
          if (cap < *n) throw EngineError(HIQ_ERR_ARG, "hiq_collect_timings: buffer too small");
          for (int i = 0; i < *n; ++i) {
               ms[i] = t[i].ms;
               kind[i] = t[i].kind;
               k[i] = t[i].k;
               variant[i] = t[i].variant;
               if (n_ref) n_ref[i] = t[i].n_ref;
          }
     });
}

int hiq_stream(hiq_engine* e, void** stream)
{
     NEED(e);
     NEED_OUT("hiq_stream", stream);
     *stream = e->impl.stream();
     return HIQ_OK;
}

}  // extern "C"

namespace {
int trace_get(const std::vector<Descriptor>& t, int i, hiq_descriptor* d, double* payload, int cap_payload, int64_t* aux, int cap_aux)
{
     if (!d) return set_error(HIQ_ERR_ARG, "hiq_trace_get: null descriptor");
     if (i < 0 || i >= static_cast<int>(t.size())) return set_error(HIQ_ERR_ARG, "hiq_trace_get: index out of range");
     const Descriptor& s = t[i];
     d->kind = s.kind;
     d->k = s.k;
     for (int l = 0; l < 5; ++l) d->slots[l] = s.slots[l];
     d->ctrl_mask = s.ctrl_mask;
     d->n_payload = static_cast<int>(s.payload.size());
     d->n_aux = static_cast<int>(s.aux.size());
     if (payload) {
          if (cap_payload < d->n_payload) return set_error(HIQ_ERR_ARG, "hiq_trace_get: payload buffer too small");
          if (!s.payload.empty()) std::memcpy(payload, s.payload.data(), sizeof(cplx) * s.payload.size());
     }
     if (aux) {
          if (cap_aux < d->n_aux) return set_error(HIQ_ERR_ARG, "hiq_trace_get: aux buffer too small");
          std::copy(s.aux.begin(), s.aux.end(), aux);
     }
     return HIQ_OK;
}
}  // namespace

extern "C" {

int hiq_trace_count(hiq_engine* e, int* n)
{
     NEED(e);
     NEED_OUT("hiq_trace_count", n);
     *n = static_cast<int>(e->impl.trace().size());
     return HIQ_OK;
}

int hiq_trace_get(hiq_engine* e, int i, hiq_descriptor* d, double* payload, int cap_payload, int64_t* aux, int cap_aux)
{
     NEED(e);
     return trace_get(e->impl.trace(), i, d, payload, cap_payload, aux, cap_aux);
}

int hiq_launch_trace_count(hiq_engine* e, int* n)
{
     NEED(e);
     NEED_OUT("hiq_launch_trace_count", n);
     *n = static_cast<int>(e->impl.launch_trace().size());
     return HIQ_OK;
}

int hiq_launch_trace_get(hiq_engine* e, int i, hiq_descriptor* d, double* payload, int cap_payload, int64_t* aux, int cap_aux)
{
     NEED(e);
     return trace_get(e->impl.launch_trace(), i, d, payload, cap_payload, aux, cap_aux);
}

int hiq_trace_clear(hiq_engine* e)
{
     NEED(e);
     e->impl.clear_trace();
     return HIQ_OK;
}

}  // extern "C"
