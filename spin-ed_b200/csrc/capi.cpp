// capi.cpp -- the extern "C" surface declared in include/sped.h.
//
// Group A mirrors, symbol for symbol, the `foreign import ccall` list of
// /root/reference/src/SpinED/Internal.hs (line numbers in include/sped.h).  Handles are heap cells
// holding a std::shared_ptr, so every object keeps what it depends on alive and the GHC finalizers
// (Internal.hs:116,150,228,362,402) may run in any order.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <new>

#include <nvtx3/nvToolsExt.h>

#include "internal.h"

namespace sped {

bool g_logging = std::getenv("SPED_LOG") != nullptr;  // also ls_enable_logging()

NvtxRange::NvtxRange(char const* name) { nvtxRangePushA(name); }
NvtxRange::~NvtxRange() { nvtxRangePop(); }
std::uint64_t g_launches = 0;

static thread_local int t_last_code = 0;
static thread_local std::string t_last_message;

void fail(int code, std::string what) { throw Error{code, std::move(what)}; }

void cuda_check(cudaError_t e, char const* expr, char const* file, int line) {
  if (e == cudaSuccess) return;
  std::string msg = std::string("CUDA error '") + cudaGetErrorString(e) + "' in " + expr + " at " + file + ":" +
                    std::to_string(line);
  cudaGetLastError();  // clear the sticky error so later calls report their own
  fail(SPED_CUDA_ERROR, msg);
}

int guarded(void (*thunk)(void*), void* ctx) {
  try {
    thunk(ctx);
    return LS_SUCCESS;
  } catch (Error const& e) {
    t_last_code = e.code;
    t_last_message = e.what;
    SPED_LOG("error %d: %s", e.code, e.what.c_str());
    return e.code;
  } catch (std::bad_alloc const&) {
    t_last_code = LS_OUT_OF_MEMORY;
    t_last_message = "host allocation failed";
    return LS_OUT_OF_MEMORY;
  } catch (std::exception const& e) {
    t_last_code = LS_SYSTEM_ERROR;
    t_last_message = e.what();
    return LS_SYSTEM_ERROR;
  }
}

void basis_state_info(Basis& b, u64 count, u64 const* states, u64* reps, double* chars, double* norms);
void small_eigh(int m, std::vector<cplx> A, std::vector<double>& evals, std::vector<cplx>& evecs);
std::string jit_program_source(Basis const& b);
size_t jit_compile_only(Basis& b, int dtype, int nb);

namespace {

char const* code_text(int code) {
  switch (code) {
    case LS_SUCCESS: return "no error";
    case LS_OUT_OF_MEMORY: return "failed to allocate memory";
    case LS_INVALID_ARGUMENT: return "argument is invalid";
    case LS_INVALID_HAMMING_WEIGHT: return "specified Hamming weight is invalid";
    case LS_INVALID_SPIN_INVERSION: return "specified spin_inversion is invalid";
    case LS_INVALID_NUMBER_SPINS: return "specified number of spins is invalid";
    case LS_INVALID_PERMUTATION: return "argument is not a valid permutation";
    case LS_INVALID_SECTOR: return "specified sector exceeds the periodicity of the operator";
    case LS_INVALID_STATE: return "invalid basis state";
    case LS_INVALID_DATATYPE: return "invalid datatype";
    case LS_PERMUTATION_TOO_LONG: return "such long permutations are not supported";
    case LS_INCOMPATIBLE_SYMMETRIES: return "symmetries are incompatible";
    case LS_NOT_A_REPRESENTATIVE: return "spin configuration is not a representative";
    case LS_WRONG_BASIS_TYPE: return "expected a basis of different type";
    case LS_CACHE_NOT_BUILT: return "list of representatives is not yet built";
    case LS_COULD_NOT_OPEN_FILE: return "failed to open file";
    case LS_FILE_IO_FAILED: return "file input/output failed";
    case LS_CACHE_IS_CORRUPT: return "file does not contain a list of representatives";
    case LS_OPERATOR_IS_COMPLEX: return "trying to apply complex operator to real vector";
    case LS_DIMENSION_MISMATCH: return "operator dimension does not match vector length";
    case LS_SYSTEM_ERROR: return "unknown error";
    case SPED_CUDA_ERROR: return "CUDA runtime error";
    case SPED_NCCL_ERROR: return "NCCL error";
    case SPED_NOT_CONVERGED: return "eigensolver did not converge";
    case SPED_INTERNAL_ERROR: return "internal consistency check failed";
  }
  return "unrecognised error code";
}

template <class T>
using Cell = std::shared_ptr<T>;
template <class T>
void* to_handle(std::shared_ptr<T> p) { return new Cell<T>(std::move(p)); }
template <class T>
std::shared_ptr<T>& from_handle(void const* h) { return *static_cast<Cell<T>*>(const_cast<void*>(h)); }
template <class T>
void drop_handle(void* h) { delete static_cast<Cell<T>*>(h); }

struct States {
  std::shared_ptr<std::vector<u64>> data;
};

}  // namespace
}  // namespace sped

using namespace sped;

extern "C" {

char const* ls_error_to_string(int code) {
  std::string s = code_text(code);
  if (code != LS_SUCCESS && code == t_last_code && !t_last_message.empty()) s += ": " + t_last_message;
  char* out = static_cast<char*>(std::malloc(s.size() + 1));
  if (out) std::memcpy(out, s.c_str(), s.size() + 1);
  return out;
}
void ls_destroy_string(char const* s) { std::free(const_cast<char*>(s)); }
void ls_enable_logging(void) { g_logging = true; }
void ls_disable_logging(void) { g_logging = false; }

int ls_create_symmetry(void** out, unsigned length, unsigned const* permutation, unsigned sector) {
  return guard([&] { *out = to_handle(make_symmetry(length, permutation, sector)); });
}
void ls_destroy_symmetry(void* s) { drop_handle<Symmetry>(s); }
unsigned ls_get_sector(void const* s) { return from_handle<Symmetry>(s)->sector; }
double ls_get_phase(void const* s) {
  auto& p = from_handle<Symmetry>(s);
  return (double)p->sector / (double)p->periodicity;
}
unsigned ls_get_periodicity(void const* s) { return from_handle<Symmetry>(s)->periodicity; }

int ls_create_group(void** out, unsigned size, void const* const* generators) {
  return guard([&] {
    std::vector<Symmetry const*> gens;
    for (unsigned i = 0; i < size; ++i) gens.push_back(from_handle<Symmetry>(generators[i]).get());
    *out = to_handle(make_group(gens));
  });
}
void ls_destroy_group(void* g) { drop_handle<Group>(g); }
unsigned ls_get_group_size(void const* g) { return (unsigned)from_handle<Group>(g)->elems.size(); }

int ls_create_spin_basis(void** out, void const* group, unsigned number_spins, int hamming_weight, int spin_inversion) {
  return guard([&] { *out = to_handle(make_basis(from_handle<Group>(group), number_spins, hamming_weight, spin_inversion)); });
}
void ls_destroy_spin_basis(void* b) { drop_handle<Basis>(b); }
int ls_build(void* basis) {
  return guard([&] { from_handle<Basis>(basis)->build(); });
}
int ls_build_unsafe(void* basis, uint64_t size, uint64_t const* representatives) {
  return guard([&] { from_handle<Basis>(basis)->adopt(size, representatives); });
}
int ls_get_number_states(void const* basis, uint64_t* out) {
  return guard([&] {
    auto& b = from_handle<Basis>(basis);
    if (!b->built) fail(LS_CACHE_NOT_BUILT, "basis has not been built");
    *out = b->n_states;
  });
}
int ls_get_states(void** out_states, void const* basis) {
  return guard([&] {
    auto* s = new States{from_handle<Basis>(basis)->states_host()};
    *out_states = s;
  });
}
uint64_t const* ls_states_get_data(void const* states) { return static_cast<States const*>(states)->data->data(); }
uint64_t ls_states_get_size(void const* states) { return static_cast<States const*>(states)->data->size(); }
void ls_destroy_states(void* states) { delete static_cast<States*>(states); }

static int create_interaction(void** out, int k, void const* matrix, unsigned n, uint16_t const* sites) {
  return guard([&] { *out = to_handle(make_interaction(k, matrix, n, sites)); });
}
int ls_create_interaction1(void** out, void const* m, unsigned n, uint16_t const* sites) { return create_interaction(out, 1, m, n, sites); }
int ls_create_interaction2(void** out, void const* m, unsigned n, uint16_t const* sites) { return create_interaction(out, 2, m, n, sites); }
int ls_create_interaction3(void** out, void const* m, unsigned n, uint16_t const* sites) { return create_interaction(out, 3, m, n, sites); }
int ls_create_interaction4(void** out, void const* m, unsigned n, uint16_t const* sites) { return create_interaction(out, 4, m, n, sites); }
bool ls_interaction_is_real(void const* t) { return from_handle<Interaction>(t)->is_real(); }
void ls_destroy_interaction(void* t) { drop_handle<Interaction>(t); }

int ls_create_operator(void** out, void const* basis, unsigned number_terms, void const* const* terms) {
  return guard([&] {
    std::vector<Interaction const*> ts;
    for (unsigned i = 0; i < number_terms; ++i) ts.push_back(from_handle<Interaction>(terms[i]).get());
    *out = to_handle(make_operator(from_handle<Basis>(basis), ts));
  });
}
void ls_destroy_operator(void* op) { drop_handle<Operator>(op); }
bool ls_operator_is_real(void const* op) { return from_handle<Operator>(op)->is_real(); }
int ls_operator_matmat(void const* op, int dtype, uint64_t size, uint64_t block_size, void const* x, uint64_t x_stride,
                       void* y, uint64_t y_stride) {
  return guard([&] { from_handle<Operator>(op)->matmat_host(dtype, size, block_size, x, x_stride, y, y_stride); });
}
int ls_operator_expectation(void const* op, int dtype, uint64_t size, uint64_t block_size, void const* x,
                            uint64_t x_stride, void* out) {
  return guard([&] {
    from_handle<Operator>(op)->expectation_host(dtype, size, block_size, x, x_stride, static_cast<cplx*>(out));
  });
}

/* ------------------------------- group B ------------------------------- */

char const* sped_version(void) { return "sped-b200 0.1.0 (sm_100a)"; }

int sped_device_count(int* out) {
  return guard([&] {
    int n = 0;
    CUDA_CHECK(cudaGetDeviceCount(&n));
    if (n <= 0) fail(SPED_CUDA_ERROR, "no CUDA device available");
    *out = n;
  });
}
int sped_set_device(int device) {
  return guard([&] { CUDA_CHECK(cudaSetDevice(device)); });
}
uint64_t sped_kernel_launches(void) { return g_launches; }

int sped_comm_unique_id(void* out) {
  return guard([&] { comm_unique_id(out); });
}
int sped_comm_init(int world, int rank, void const* id) {
  return guard([&] { comm_init(world, rank, id); });
}
int sped_comm_finalize(void) {
  return guard([&] { comm_finalize(); });
}
int sped_comm_rank(void) { return comm().rank; }
int sped_comm_size(void) { return comm().world; }
static_assert(sizeof(sped_row_dist) == sizeof(RowDist), "sped_row_dist mirrors RowDist");
void sped_row_distribution(uint64_t n, int world, int rank, sped_row_dist* out) {
  RowDist d = make_row_dist(n, world < 1 ? 1 : world, rank);
  std::memcpy(out, &d, sizeof d);
}
uint64_t sped_dist_local_to_global(sped_row_dist const* d, uint64_t local_index) {
  RowDist r;
  std::memcpy(&r, d, sizeof r);
  return dist_local_to_global(r, local_index);
}
uint64_t sped_dist_global_to_position(sped_row_dist const* d, uint64_t global_row) {
  RowDist r;
  std::memcpy(&r, d, sizeof r);
  return dist_global_to_pos(r, global_row);
}

int sped_basis_build_seconds(void const* basis, double* out) {
  return guard([&] { *out = from_handle<Basis>(basis)->build_seconds; });
}
int sped_basis_row_distribution(void const* basis, sped_row_dist* out) {
  return guard([&] {
    auto& b = from_handle<Basis>(basis);
    if (!b->built) fail(LS_CACHE_NOT_BUILT, "basis has not been built");
    RowDist d = b->dist();
    std::memcpy(out, &d, sizeof d);
  });
}
int sped_basis_device_states(void const* basis, uint64_t const** out) {
  return guard([&] {
    auto& b = from_handle<Basis>(basis);
    if (!b->built) fail(LS_CACHE_NOT_BUILT, "basis has not been built");
    *out = b->d_reps.ptr;
  });
}
int sped_basis_norms(void const* basis, double* out) {
  return guard([&] {
    auto& b = from_handle<Basis>(basis);
    if (!b->built) fail(LS_CACHE_NOT_BUILT, "basis has not been built");
    if (b->trivial()) {
      for (u64 i = 0; i < b->n_states; ++i) out[i] = 1.0;
      return;
    }
    auto stab = b->d_stab.download();
    double order = (double)b->group_order();
    for (u64 i = 0; i < b->n_states; ++i) out[i] = std::sqrt((double)stab[i] / order);
  });
}
int sped_basis_state_info(void const* basis, uint64_t count, uint64_t const* states, uint64_t* reps, double* chars,
                          double* norms) {
  return guard([&] { basis_state_info(*from_handle<Basis>(basis), count, states, reps, chars, norms); });
}
int sped_basis_program_stats(void const* basis, unsigned* steps, unsigned* rot_ops, unsigned* benes_ops) {
  return guard([&] {
    auto& b = from_handle<Basis>(basis);
    *steps = (unsigned)b->program.steps.size();
    *rot_ops = b->program.rot_ops + 2 * b->program.fast_steps;
    *benes_ops = b->program.benes_ops;
  });
}

int sped_operator_matmat_device(void const* op, int dtype, uint64_t block_size, void const* x_full, uint64_t x_stride,
                                void* y_local, uint64_t y_stride, void* stream) {
  return guard([&] {
    from_handle<Operator>(op)->matmat_device(dtype, block_size, x_full, x_stride, y_local, y_stride,
                                            static_cast<cudaStream_t>(stream));
  });
}
int sped_operator_matvec_sharded(void const* op, int dtype, void const* x_local, void* y_local, void* x_replicated,
                                 void* stream) {
  return guard([&] {
    from_handle<Operator>(op)->matvec_sharded(dtype, x_local, y_local, x_replicated, static_cast<cudaStream_t>(stream));
  });
}
int sped_operator_matmat_local(void const* op, int dtype, uint64_t block_size, void const* x_local, uint64_t x_stride,
                               void* y_local, uint64_t y_stride) {
  return guard([&] { from_handle<Operator>(op)->matmat_host_local(dtype, block_size, x_local, x_stride, y_local, y_stride); });
}
int sped_operator_count_elements(void const* op, uint64_t* rows, uint64_t* offdiag) {
  return guard([&] {
    u64 r, e;
    from_handle<Operator>(op)->count_elements(r, e);
    *rows = r;
    *offdiag = e;
  });
}
int sped_operator_diagonal(void const* op, double* out) {
  return guard([&] {
    auto& o = from_handle<Operator>(op);
    o->prepare();
    u64 n = o->dist.n_local;
    if (n) CUDA_CHECK(cudaMemcpy(out, o->d_diag.ptr, n * sizeof(double), cudaMemcpyDeviceToHost));
  });
}

int sped_operator_set_cache(void const* op, int mode) {
  return guard([&] {
    auto& o = from_handle<Operator>(op);
    if (mode < -1 || mode > 1) fail(LS_INVALID_ARGUMENT, "cache mode must be -1, 0 or 1");
    o->drop_cache();
    o->cache_mode = mode;
  });
}
int sped_operator_cache_info(void const* op, int* ready, uint64_t* bytes, double* build_seconds) {
  return guard([&] {
    auto& o = from_handle<Operator>(op);
    *ready = o->cache_ready ? 1 : 0;
    *bytes = o->cache_bytes;
    *build_seconds = o->cache_build_seconds;
  });
}

int sped_eigh(void const* op, int dtype, uint64_t n_evals, double eps, int max_basis_size, int max_block_size,
              int min_restart_size, double* evals, void* evecs, double* rnorms, sped_monitor_fn monitor, void* ctx) {
  int status = LS_SUCCESS;
  int rc = guard([&] {
    status = eigh(*from_handle<Operator>(op), dtype, n_evals, eps, max_basis_size, max_block_size, min_restart_size,
                  evals, evecs, rnorms, monitor, ctx);
  });
  return rc != LS_SUCCESS ? rc : status;
}
int sped_operator_release_workspace(void const* op) {
  return guard([&] { from_handle<Operator>(op)->release_workspace(); });
}
int sped_eigh_last_stats(void const* op, sped_eigh_stats* out) {
  return guard([&] {
    auto& s = from_handle<Operator>(op)->last_stats;
    out->matvecs = s.matvecs;
    out->iterations = s.iterations;
    out->restarts = s.restarts;
    out->seconds_total = s.seconds_total;
    out->seconds_matvec = s.seconds_matvec;
    out->seconds_ortho = s.seconds_ortho;
    out->seconds_residual = s.seconds_residual;
    out->seconds_restart = s.seconds_restart;
    out->seconds_project = s.seconds_project;
  });
}

/* Host-only self checks used by the CPU test-suite (no GPU work). */
int sped_selftest_small_eigh(int m, double const* a_re_im, double* evals, double* evecs_re_im) {
  return guard([&] {
    std::vector<cplx> A((size_t)m * m);
    for (size_t i = 0; i < A.size(); ++i) A[i] = cplx(a_re_im[2 * i], a_re_im[2 * i + 1]);
    std::vector<double> ev;
    std::vector<cplx> V;
    small_eigh(m, A, ev, V);
    for (int i = 0; i < m; ++i) evals[i] = ev[i];
    for (size_t i = 0; i < V.size(); ++i) {
      evecs_re_im[2 * i] = V[i].real();
      evecs_re_im[2 * i + 1] = V[i].imag();
    }
  });
}
/* Runs the compiled canonicalisation program of `basis` on the host (verification only: the
 * product path never canonicalises on the CPU). */
int sped_selftest_program(void const* basis, uint64_t count, uint64_t const* states, uint64_t* reps, int* phases,
                          int* stabs) {
  return guard([&] {
    auto& b = from_handle<Basis>(basis);
    auto const& P = b->program;
    std::vector<PermOp<u32>> ops32;
    for (auto const& o : P.ops) ops32.push_back(PermOp<u32>{(u32)o.mask, o.amount});
    std::vector<FastStep<u32>> fast32;
    for (auto const& f : P.fast) fast32.push_back(FastStep<u32>{(u32)f.mask, f.ctl});
    ProgramView<u64> v{P.fast.data(), P.steps.data(), P.ops.data(), P.phase.data(), (u32)P.steps.size(),
                       (u32)P.ops.size(), P.n_spins, P.shift, P.inversion, P.denom};
    ProgramView<u32> v32{fast32.data(), P.steps.data(), ops32.data(), P.phase.data(), (u32)P.steps.size(),
                         (u32)P.ops.size(), P.n_spins, 0, P.inversion, P.denom};
    for (u64 i = 0; i < count; ++i) {
      u32 step, flipped;
      if (b->use32()) {
        u32 r;
        canonicalize<u32>(v32, (u32)states[i], r, step, flipped);
        reps[i] = r;
        phases[i] = element_phase<u32>(v32, step, flipped);
        stabs[i] = stabilizer_scan<u32>(v32, r, false);
      } else {
        u64 r;
        canonicalize<u64>(v, states[i], r, step, flipped);
        reps[i] = r;
        phases[i] = element_phase<u64>(v, step, flipped);
        stabs[i] = stabilizer_scan<u64>(v, r, false);
      }
    }
  });
}
int sped_selftest_burnside(void const* basis, uint64_t* out) {
  return guard([&] { *out = from_handle<Basis>(basis)->expected_dimension(); });
}
int sped_selftest_jit_source(void const* basis, char* out, uint64_t capacity, uint64_t* needed) {
  return guard([&] {
    std::string src = jit_program_source(*from_handle<Basis>(basis));
    *needed = src.size() + 1;
    if (out && capacity) {
      size_t n = std::min<size_t>(capacity - 1, src.size());
      std::memcpy(out, src.data(), n);
      out[n] = 0;
    }
  });
}
int sped_selftest_jit_compile(void const* basis, int dtype, int columns, uint64_t* cubin_bytes) {
  return guard([&] { *cubin_bytes = jit_compile_only(*from_handle<Basis>(basis), dtype, columns); });
}

}  // extern "C"
