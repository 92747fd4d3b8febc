/* Exit while NVRTC is still compiling: ls_build starts the specialisation of the canonicalisation on a
 * background thread (jit.cpp); the process must neither crash nor hang when it ends -- with the basis
 * destroyed (the compilation is waited for there) or leaked (the exit-time registry waits).  Runs
 * without a GPU: SPED_JIT_PREFETCH=1 starts the compilation even though ls_build then fails for lack
 * of a device.  Driven by tests/test_conformance.py. */
#include <stdio.h>
#include <stdint.h>
#include "sped.h"
int main(int argc, char** argv) {
  (void)argv;
  enum { N = 36 };
  unsigned t[N], p[N];
  for (int i = 0; i < N; ++i) { t[i] = (i + 1) % N; p[i] = N - 1 - i; }
  void *s_t, *s_p, *group, *basis;
  if (ls_create_symmetry(&s_t, N, t, 0)) return 1;
  if (ls_create_symmetry(&s_p, N, p, 0)) return 2;
  void const* gens[2] = {s_t, s_p};
  if (ls_create_group(&group, 2, gens)) return 3;
  if (ls_create_spin_basis(&basis, group, N, N / 2, 1)) return 4;
  int rc = ls_build(basis);
  printf("build rc %d (no GPU here: an error is expected)\n", rc);
  fflush(stdout);
  if (argc > 1) ls_destroy_spin_basis(basis);
  return 0;
}
