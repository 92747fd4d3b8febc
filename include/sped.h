/* sped.h -- C ABI of libsped.so, the B200-native replacement for SpinED's numerical back end.
 *
 * Group A (`ls_*`) is signature-exact with the symbols the reference binds through
 * `foreign import ccall` in /root/reference/src/SpinED/Internal.hs (line numbers cited per entry);
 * upstream they are provided by liblattice_symmetries.  Group B (`sped_*`) is new: it replaces the
 * single PRIMME call `eigh primmeOptions primmeOperator` (/root/reference/src/SpinED.hs:404) with a
 * device-resident eigensolver, and exposes the device-pointer / multi-GPU plumbing a host driver
 * needs.  All pointers in group A are HOST pointers (Internal.hs:233-234,417-419,441-442).
 *
 * Every fallible call returns an `int` status, 0 = success; `ls_error_to_string` turns it into a
 * heap string the caller releases with `ls_destroy_string` (Internal.hs:48-65).  Out-parameters
 * are written only on success (Internal.hs:96-97).  Handles are reference counted internally:
 * a group copies its generators, a basis shares its group, an operator shares its basis and
 * copies its terms, an `ls_states` view keeps the representatives alive on its own
 * (Internal.hs:136-150,238-244,386-402), so destroy order is free (GHC finalizers).
 */
#ifndef SPED_H
#define SPED_H

#include <stdbool.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ---- status codes (values follow liblattice_symmetries' ls_error_code enum order) ---- */
enum {
  LS_SUCCESS = 0,
  LS_OUT_OF_MEMORY = 1,
  LS_INVALID_ARGUMENT = 2,
  LS_INVALID_HAMMING_WEIGHT = 3,
  LS_INVALID_SPIN_INVERSION = 4,
  LS_INVALID_NUMBER_SPINS = 5,
  LS_INVALID_PERMUTATION = 6,
  LS_INVALID_SECTOR = 7,
  LS_INVALID_STATE = 8,
  LS_INVALID_DATATYPE = 9,
  LS_PERMUTATION_TOO_LONG = 10,
  LS_INCOMPATIBLE_SYMMETRIES = 11,
  LS_NOT_A_REPRESENTATIVE = 12,
  LS_WRONG_BASIS_TYPE = 13,
  LS_CACHE_NOT_BUILT = 14,
  LS_COULD_NOT_OPEN_FILE = 15,
  LS_FILE_IO_FAILED = 16,
  LS_CACHE_IS_CORRUPT = 17,
  LS_OPERATOR_IS_COMPLEX = 18,
  LS_DIMENSION_MISMATCH = 19,
  LS_SYSTEM_ERROR = 20,
  /* sped additions */
  SPED_CUDA_ERROR = 100,      /* a CUDA runtime call failed (includes "no CUDA device") */
  SPED_NCCL_ERROR = 101,
  SPED_NOT_CONVERGED = 102,
  SPED_INTERNAL_ERROR = 103
};

/* datatype tags of ls_operator_matmat / ls_operator_expectation (Internal.hs:404-409) */
enum { SPED_F32 = 0, SPED_F64 = 1, SPED_C64 = 2, SPED_C128 = 3 };

/* =============================== group A: ls_* ==================================== */

/* errors / logging -- Internal.hs:39-45 */
char const* ls_error_to_string(int code);
void ls_destroy_string(char const* s);
void ls_enable_logging(void);
void ls_disable_logging(void);

/* symmetry -- Internal.hs:69-79.  Validates that `permutation` is a bijection of 0..length-1,
 * computes its periodicity and requires sector < periodicity (test/Spec.hs:38-43). */
int ls_create_symmetry(void** out, unsigned length, unsigned const* permutation, unsigned sector);
void ls_destroy_symmetry(void* symmetry);
unsigned ls_get_sector(void const* symmetry);
double ls_get_phase(void const* symmetry);
unsigned ls_get_periodicity(void const* symmetry);

/* group -- Internal.hs:120-127.  Closure of the generators; the same permutation reached with
 * two different phases is LS_INCOMPATIBLE_SYMMETRIES (test/Spec.hs:72-79). */
int ls_create_group(void** out, unsigned size, void const* const* generators);
void ls_destroy_group(void* group);
unsigned ls_get_group_size(void const* group);

/* basis -- Internal.hs:172-196.  hamming_weight = -1: unrestricted; spin_inversion = 0: none. */
int ls_create_spin_basis(void** out, void const* group, unsigned number_spins, int hamming_weight,
                         int spin_inversion);
void ls_destroy_spin_basis(void* basis);
int ls_build(void* basis);                                              /* Internal.hs:178 */
int ls_build_unsafe(void* basis, uint64_t size, uint64_t const* representatives); /* :181 */
int ls_get_number_states(void const* basis, uint64_t* out);             /* :184 */
int ls_get_states(void** out_states, void const* basis);                /* :187 */
uint64_t const* ls_states_get_data(void const* states);                 /* :190 */
uint64_t ls_states_get_size(void const* states);                        /* :193 */
void ls_destroy_states(void* states);                                   /* :196 */

/* interactions -- Internal.hs:260-275.  `matrix` is a row-major 2^k x 2^k complex<double>;
 * `sites` holds number_tuples * k site indices; first listed site = most significant bit. */
int ls_create_interaction1(void** out, void const* matrix_2x2, unsigned number_tuples, uint16_t const* sites);
int ls_create_interaction2(void** out, void const* matrix_4x4, unsigned number_tuples, uint16_t const* sites);
int ls_create_interaction3(void** out, void const* matrix_8x8, unsigned number_tuples, uint16_t const* sites);
int ls_create_interaction4(void** out, void const* matrix_16x16, unsigned number_tuples, uint16_t const* sites);
bool ls_interaction_is_real(void const* interaction);
void ls_destroy_interaction(void* interaction);

/* operator -- Internal.hs:371-383.  x and y are column-major [size x block_size] HOST blocks with
 * strides in elements; size must equal the number of representatives. */
int ls_create_operator(void** out, void const* basis, unsigned number_terms, void const* const* terms);
void ls_destroy_operator(void* op);
bool ls_operator_is_real(void const* op);
int ls_operator_matmat(void const* op, int dtype, uint64_t size, uint64_t block_size, void const* x,
                       uint64_t x_stride, void* y, uint64_t y_stride);
int ls_operator_expectation(void const* op, int dtype, uint64_t size, uint64_t block_size, void const* x,
                            uint64_t x_stride, void* out_complex128);

/* =============================== group B: sped_* ================================== */

/* Library / device management. */
char const* sped_version(void);
int sped_device_count(int* out);                  /* SPED_CUDA_ERROR when no usable CUDA device */
int sped_set_device(int device);
uint64_t sped_kernel_launches(void);              /* kernels launched by this library so far */

/* Multi-GPU: one process per GPU.  Rank 0 obtains a 128-byte NCCL unique id, the host driver
 * distributes it (torch.distributed / MPI / a file), every rank calls sped_comm_init once before
 * creating bases.  Rows (and the enumeration work of ls_build) are dealt block-cyclically over the
 * ranks (sped_row_dist below); the Krylov vector is all-gathered each matvec -- overlapped with the
 * part of the product that needs no remote entries -- and dot products are all-reduced (NCCL over
 * NVLink). */
int sped_comm_unique_id(void* out_128_bytes);
int sped_comm_init(int world_size, int rank, void const* unique_id_128_bytes);
int sped_comm_finalize(void);
int sped_comm_rank(void);
int sped_comm_size(void);
/* Row distribution (host logic, no GPU).  Rows are dealt to the ranks round-robin in blocks of
 * 2^log2_block consecutive rows, so that every rank holds the same mix of rows (the sorted
 * representatives are not homogeneous: contiguous blocks leave the ranks unbalanced).  A rank keeps
 * its `n_local` rows compactly in local order; the replicated vector handed to
 * sped_operator_matmat_device is laid out [rank][local index] with every shard padded to `chunk`
 * entries -- what an all-gather of the local shards produces.  With one rank local == global. */
typedef struct sped_row_dist {
  uint64_t n, n_local, chunk;
  uint32_t world, rank, log2_block, reserved;
} sped_row_dist;
void sped_row_distribution(uint64_t n, int world, int rank, sped_row_dist* out);
uint64_t sped_dist_local_to_global(sped_row_dist const* d, uint64_t local_index);
uint64_t sped_dist_global_to_position(sped_row_dist const* d, uint64_t global_row);

/* Basis timing / layout queries. */
int sped_basis_build_seconds(void const* basis, double* out);   /* device time of the last ls_build */
int sped_basis_row_distribution(void const* basis, sped_row_dist* out);   /* rows of this rank */
int sped_basis_device_states(void const* basis, uint64_t const** out_device_ptr);
/* norms of the representatives as the matvec uses them (host copy, length N) */
int sped_basis_norms(void const* basis, double* out);
/* Representative, character and norm of an arbitrary basis word (device canonicalisation). */
int sped_basis_state_info(void const* basis, uint64_t count, uint64_t const* states, uint64_t* representatives,
                          double* characters_re_im, double* norms);
/* Description of the compiled canonicalisation program (steps, rotate-mask ops, delta-swap ops). */
int sped_basis_program_stats(void const* basis, unsigned* steps, unsigned* rot_ops, unsigned* benes_ops);

/* Device-resident operator application: x_full is the replicated vector, [world * chunk x block]
 * in the [rank][local] layout of sped_row_dist, y_local the rows of this rank [n_local x block],
 * both DEVICE pointers, column-major.  With one rank this is plainly y = H x.  `stream` is a
 * cudaStream_t (NULL = default stream). */
int sped_operator_matmat_device(void const* op, int dtype, uint64_t block_size, void const* x_full,
                                uint64_t x_stride, void* y_local, uint64_t y_stride, void* stream);
/* One column, row-sharded over the ranks of the communicator: x_local holds this rank's n_local
 * entries, x_replicated is a [world * chunk] DEVICE work vector in the [rank][local] layout (left
 * holding the gathered vector), y_local receives this rank's rows.  The exchange of the shards -- one
 * NCCL all-gather, or for shards of 64 MB and more two rounds of copy-engine pulls from the peers'
 * IPC-mapped send buffers (SPED_EXCHANGE=ce|nccl forces either) -- runs on the library's own streams
 * and overlaps the part of the product whose source entries this rank owns; only the remote-source
 * parts wait for it.  This is what sped_eigh does per matvec; the per-row summation order (local
 * class, then the remote classes) depends on the number of ranks, so results agree across rank
 * counts to rounding, not bitwise.  Collective: every rank of the communicator must call it. */
int sped_operator_matvec_sharded(void const* op, int dtype, void const* x_local, void* y_local, void* x_replicated,
                                 void* stream);
/* Host-pointer form of the row-sharded product, for a rank-parallel host eigensolver (PRIMME's
 * parallel mode keeps nLocal rows per process): x_local / y_local are HOST blocks holding this
 * rank's n_local rows (local order, see sped_row_distribution), column-major with the given
 * strides.  Per column a rank sends n_local entries over PCIe, the shards are exchanged over NVLink
 * inside the sharded product, and n_local entries come back -- unlike ls_operator_matmat, where
 * every rank is handed the full x and receives the full y (/root/reference/src/SpinED/Internal.hs:411-429,
 * a shared-memory reference has no notion of ranks).  With one rank the two calls coincide. */
int sped_operator_matmat_local(void const* op, int dtype, uint64_t block_size, void const* x_local, uint64_t x_stride,
                               void* y_local, uint64_t y_stride);
/* Number of matrix elements one application touches: rows N and off-diagonal elements E
 * (term applications with non-zero target norm); global counts. */
int sped_operator_count_elements(void const* op, uint64_t* rows, uint64_t* offdiag);
/* Diagonal of the operator on the local row block (host copy). */
int sped_operator_diagonal(void const* op, double* out_local_rows);
/* Operator cache.  By default (mode -1, or the SPED_OPERATOR_CACHE environment variable) the first
 * application of an operator is matrix-free and stores the off-diagonal elements it finds in HBM
 * when they fit in half of the free device memory; later applications stream them (same elements,
 * results equal to rounding).  mode 0: always matrix-free; mode 1: cache whenever it fits at all.  Changing the mode
 * drops an existing cache. */
int sped_operator_set_cache(void const* op, int mode);
int sped_operator_cache_info(void const* op, int* ready, uint64_t* bytes, double* build_seconds);

/* Progress record handed to the monitor callback (mirrors PRIMME's monitor, SpinED.hs:379-382). */
typedef struct sped_eigh_info {
  int iteration;          /* outer iteration */
  int basis_size;         /* current search-space dimension */
  int number_converged;
  int number_evals;
  uint64_t number_matvecs;
  double const* evals;    /* current Ritz values [number_evals] */
  double const* rnorms;   /* their residual norms [number_evals] */
  double elapsed_seconds;
} sped_eigh_info;
typedef int (*sped_monitor_fn)(sped_eigh_info const* info, void* ctx); /* non-zero return aborts */

/* Replaces `eigh` (SpinED.hs:404).  Option fields map to ConfigSpec (SpinED.hs:158-173):
 * number_vectors, precision (0 -> 1e4 * machine epsilon of dtype's real type, PRIMME's default),
 * max_primme_basis_size / max_primme_block_size / min_primme_restart_size (<= 0 -> defaults).
 * evals [n_evals] and rnorms [n_evals] are doubles; evecs (may be NULL) is a column-major HOST
 * block [N x n_evals] of `dtype` elements holding the full eigenvectors on every rank. */
int sped_eigh(void const* op, int dtype, uint64_t n_evals, double eps, int max_basis_size, int max_block_size,
              int min_restart_size, double* evals, void* evecs, double* rnorms, sped_monitor_fn monitor,
              void* ctx);
/* sped_eigh keeps its device workspace (Krylov basis, H V, residuals) on the operator between calls
 * -- freeing and re-allocating gigabytes costs more wall time than a warm 6x6 solve.  This releases
 * it (it is also released with the operator). */
int sped_operator_release_workspace(void const* op);
/* Statistics of the last sped_eigh call on this operator. */
typedef struct sped_eigh_stats {
  uint64_t matvecs;
  int iterations;
  int restarts;
  double seconds_total;
  double seconds_matvec;
  double seconds_ortho;
  double seconds_residual;
  double seconds_restart;
  double seconds_project;
} sped_eigh_stats;
int sped_eigh_last_stats(void const* op, sped_eigh_stats* out);

#ifdef __cplusplus
}
#endif
#endif /* SPED_H */
