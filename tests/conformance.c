/* C conformance driver for libsped.so: the 32 ls_* symbols in the order
 * /root/reference/app/Main.hs:14-27 + src/SpinED.hs:178-248,315-337,370-411 use them, through
 * include/sped.h compiled as plain C, with handles destroyed in an "unhelpful" order (GHC
 * finalizers give no ordering guarantee, src/SpinED/Internal.hs:116,150,228,362,402).
 * Prints "CONFORMANCE_OK <n_states> <e0>" on success. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sped.h"

#define CHECK(call)                                                        \
  do {                                                                     \
    int rc_ = (call);                                                      \
    if (rc_ != 0) {                                                        \
      char const* msg_ = ls_error_to_string(rc_);                          \
      fprintf(stderr, "%s failed: [%d] %s\n", #call, rc_, msg_);           \
      ls_destroy_string(msg_);                                             \
      return 1;                                                            \
    }                                                                      \
  } while (0)

int main(void) {
  enum { N = 10 };
  unsigned t[N], p[N];
  for (unsigned i = 0; i < N; ++i) {
    t[i] = (i + 1) % N;
    p[i] = N - 1 - i;
  }
  /* error path first: an invalid permutation must fail and leave the out-parameter alone */
  void* bad = (void*)0x1;
  unsigned broken[4] = {4, 3, 4, 1};
  int rc = ls_create_symmetry(&bad, 4, broken, 0);
  if (rc == 0 || bad != (void*)0x1) return 2;
  char const* msg = ls_error_to_string(rc);
  if (!msg || !*msg) return 3;
  ls_destroy_string(msg);

  void *s_t, *s_p, *group, *basis, *term, *op, *states;
  CHECK(ls_create_symmetry(&s_t, N, t, 5));
  CHECK(ls_create_symmetry(&s_p, N, p, 1));
  if (ls_get_periodicity(s_t) != N || ls_get_sector(s_t) != 5 || fabs(ls_get_phase(s_t) - 0.5) > 1e-15) return 4;
  void const* gens[2] = {s_t, s_p};
  CHECK(ls_create_group(&group, 2, gens));
  ls_destroy_symmetry(s_t); /* generators are only alive during ls_create_group (Internal.hs:136-149) */
  ls_destroy_symmetry(s_p);
  if (ls_get_group_size(group) != 20) return 5;
  CHECK(ls_create_spin_basis(&basis, group, N, 5, -1));
  ls_destroy_group(group); /* the Haskell SpinBasis wrapper does not retain the group */

  double m[16][2];
  memset(m, 0, sizeof m);
  m[0][0] = 1; m[5][0] = -1; m[6][0] = 2; m[9][0] = 2; m[10][0] = -1; m[15][0] = 1;
  uint16_t sites[2 * N];
  for (unsigned i = 0; i < N; ++i) {
    sites[2 * i] = (uint16_t)i;
    sites[2 * i + 1] = (uint16_t)((i + 1) % N);
  }
  CHECK(ls_create_interaction2(&term, m, N, sites));
  if (!ls_interaction_is_real(term)) return 6;
  void const* terms[1] = {term};
  CHECK(ls_create_operator(&op, basis, 1, terms)); /* before ls_build, like SpinED.hs:243-248 */
  ls_destroy_interaction(term);
  if (!ls_operator_is_real(op)) return 7;

  uint64_t n = 0;
  if (ls_get_number_states(basis, &n) == 0) return 8; /* not built yet: must be an error */
  CHECK(ls_build(basis));
  CHECK(ls_get_number_states(basis, &n));
  CHECK(ls_get_states(&states, basis));
  if (ls_states_get_size(states) != n || n != 13 || ls_states_get_data(states)[0] != 31) return 9;

  double* x = calloc(n * 2, sizeof(double));
  double* y = calloc(n * 2, sizeof(double));
  for (uint64_t i = 0; i < n; ++i) {
    x[i] = 1.0 / (double)(i + 1);
    x[n + i] = (double)(i % 3) - 1.0;
  }
  CHECK(ls_operator_matmat(op, SPED_F64, n, 2, x, n, y, n));
  double expect[2][2];
  CHECK(ls_operator_expectation(op, SPED_F64, n, 2, x, n, expect));
  double dot0 = 0;
  for (uint64_t i = 0; i < n; ++i) dot0 += x[i] * y[i];
  if (fabs(dot0 - expect[0][0]) > 1e-12 * fabs(dot0) || fabs(expect[0][1]) > 1e-14) return 10;
  if (ls_operator_matmat(op, SPED_F64, n + 1, 1, x, n + 1, y, n + 1) != LS_DIMENSION_MISMATCH) return 11;

  ls_destroy_spin_basis(basis); /* the operator must keep the basis alive */
  double e0 = 0, rnorm = 0;
  double* v = calloc(n, sizeof(double));
  CHECK(sped_eigh(op, SPED_F64, 1, 0.0, 0, 0, 0, &e0, v, &rnorm, NULL, NULL));
  if (fabs(e0 + 18.061785418) > 1e-8 || rnorm > 1e-8) return 12;
  ls_destroy_operator(op);
  /* the states view outlives everything else (Internal.hs:238-244) */
  if (ls_states_get_data(states)[12] != 341) return 13;
  ls_destroy_states(states);
  printf("CONFORMANCE_OK %llu %.9f\n", (unsigned long long)n, e0);
  free(x); free(y); free(v);
  return 0;
}
