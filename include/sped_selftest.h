/* sped_selftest.h -- host-only verification hooks of libsped.so (no GPU work, not a compute path).
 * Used by the CPU test-suite to check the pieces of the product that run on the host: the
 * projected-eigenproblem solver, the Burnside sector dimension and the compiled canonicalisation
 * program (interpreted on the host for verification only -- the product never canonicalises on
 * the CPU; every ls_* / sped_* compute entry point needs a CUDA device and fails loudly without). */
#ifndef SPED_SELFTEST_H
#define SPED_SELFTEST_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* eigen-decomposition of a row-major Hermitian m x m matrix given as (re, im) pairs */
int sped_selftest_small_eigh(int m, double const* a_re_im, double* evals, double* evecs_re_im);
/* representative, phase numerator (over the group's common denominator) and |Stab| of states */
int sped_selftest_program(void const* basis, uint64_t count, uint64_t const* states, uint64_t* reps, int* phases,
                          int* stabs);
/* sector dimension by character-weighted Burnside counting */
int sped_selftest_burnside(void const* basis, uint64_t* out);
/* CUDA C++ source of the run-time specialised canonicalisation of this basis' symmetry group
 * (what jit.cpp hands to NVRTC); writes at most `capacity` bytes, returns the full length. */
int sped_selftest_jit_source(void const* basis, char* out, uint64_t capacity, uint64_t* needed);
/* Compiles the specialised matvec kernel with NVRTC for sm_100a without loading it (no GPU
 * needed); returns the cubin size. */
int sped_selftest_jit_compile(void const* basis, int dtype, int columns, uint64_t* cubin_bytes);
#ifdef __cplusplus
}
#endif
#endif
