/* Second C conformance driver: every one of the 32 ls_* symbols that
 * /root/reference/src/SpinED/Internal.hs:39-384 imports is called at least once (the table below
 * takes the address of each, so a missing export is a link error), including the entry points the
 * first driver leaves out -- ls_build_unsafe (resume path, src/SpinED.hs:319-328), 1-, 3- and
 * 4-site interactions with a complex matrix (Internal.hs:293-322), logging switches -- and the
 * handles are destroyed in EVERY order of {basis, operator, states}: GHC runs the ForeignPtr
 * finalizers in no particular order (Internal.hs:116,150,228,362,402), so whatever is still alive
 * must keep working.  Prints "CONFORMANCE_FULL_OK" on success. */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "sped.h"

#define CHECK(call)                                                          \
  do {                                                                       \
    int rc_ = (call);                                                        \
    if (rc_ != 0) {                                                          \
      char const* msg_ = ls_error_to_string(rc_);                            \
      fprintf(stderr, "line %d: %s failed: [%d] %s\n", __LINE__, #call, rc_, msg_); \
      ls_destroy_string(msg_);                                               \
      return 100;                                                            \
    }                                                                        \
  } while (0)
#define REQUIRE(cond)                                                \
  do {                                                               \
    if (!(cond)) {                                                   \
      fprintf(stderr, "line %d: %s does not hold\n", __LINE__, #cond); \
      return 101;                                                    \
    }                                                                \
  } while (0)

/* the 32 imports of Internal.hs, by address */
typedef void (*any_fn)(void);
static any_fn const k_all_symbols[] = {
    (any_fn)ls_error_to_string, (any_fn)ls_destroy_string, (any_fn)ls_enable_logging, (any_fn)ls_disable_logging,
    (any_fn)ls_create_symmetry, (any_fn)ls_destroy_symmetry, (any_fn)ls_get_sector, (any_fn)ls_get_phase,
    (any_fn)ls_get_periodicity, (any_fn)ls_create_group, (any_fn)ls_destroy_group, (any_fn)ls_get_group_size,
    (any_fn)ls_create_spin_basis, (any_fn)ls_destroy_spin_basis, (any_fn)ls_build, (any_fn)ls_build_unsafe,
    (any_fn)ls_get_number_states, (any_fn)ls_get_states, (any_fn)ls_states_get_data, (any_fn)ls_states_get_size,
    (any_fn)ls_destroy_states, (any_fn)ls_create_interaction1, (any_fn)ls_create_interaction2,
    (any_fn)ls_create_interaction3, (any_fn)ls_create_interaction4, (any_fn)ls_interaction_is_real,
    (any_fn)ls_destroy_interaction, (any_fn)ls_create_operator, (any_fn)ls_destroy_operator, (any_fn)ls_operator_matmat,
    (any_fn)ls_operator_expectation, (any_fn)ls_operator_is_real};

enum { N = 8 };

/* translation-invariant basis of an 8-site ring at half filling, momentum sector k */
static int make_basis(void** basis, unsigned sector) {
  unsigned t[N];
  for (unsigned i = 0; i < N; ++i) t[i] = (i + 1) % N;
  void *sym, *group;
  CHECK(ls_create_symmetry(&sym, N, t, sector));
  void const* gens[1] = {sym};
  CHECK(ls_create_group(&group, 1, gens));
  ls_destroy_symmetry(sym);
  REQUIRE(ls_get_group_size(group) == N);
  CHECK(ls_create_spin_basis(basis, group, N, N / 2, 0));
  ls_destroy_group(group);
  return 0;
}

/* Heisenberg bonds (2-site, real) + a field (1-site) + a chirality-like 3-site term with a complex
 * matrix + a 4-site cyclic exchange: all four constructors, real and complex */
static int make_operator(void** op, void* basis, int with_complex) {
  double m2[16][2], m1[4][2], m3[64][2], m4[256][2];
  memset(m2, 0, sizeof m2);
  memset(m1, 0, sizeof m1);
  memset(m3, 0, sizeof m3);
  memset(m4, 0, sizeof m4);
  m2[0][0] = 1; m2[5][0] = -1; m2[6][0] = 2; m2[9][0] = 2; m2[10][0] = -1; m2[15][0] = 1;
  m1[0][0] = -0.5; m1[3][0] = 0.5;
  /* 3-site: i (|a b c> -> |b c a>) - i (|a b c> -> |c a b>): Hermitian and purely imaginary */
  for (int a = 0; a < 8; ++a) {
    int b0 = (a >> 2) & 1, b1 = (a >> 1) & 1, b2 = a & 1;
    int fwd = (b1 << 2) | (b2 << 1) | b0, bwd = (b2 << 2) | (b0 << 1) | b1;
    m3[fwd * 8 + a][1] += 0.3;
    m3[bwd * 8 + a][1] -= 0.3;
  }
  /* 4-site: cyclic shift + its inverse (real symmetric) */
  for (int a = 0; a < 16; ++a) {
    int fwd = ((a << 1) & 15) | (a >> 3), bwd = (a >> 1) | ((a & 1) << 3);
    m4[fwd * 16 + a][0] += 0.25;
    m4[bwd * 16 + a][0] += 0.25;
  }
  uint16_t s2[2 * N], s1[N], s3[3 * N], s4[4 * N];
  for (unsigned i = 0; i < N; ++i) {
    s2[2 * i] = (uint16_t)i; s2[2 * i + 1] = (uint16_t)((i + 1) % N);
    s1[i] = (uint16_t)i;
    s3[3 * i] = (uint16_t)i; s3[3 * i + 1] = (uint16_t)((i + 1) % N); s3[3 * i + 2] = (uint16_t)((i + 2) % N);
    for (unsigned j = 0; j < 4; ++j) s4[4 * i + j] = (uint16_t)((i + j) % N); /* every term commutes with the translation */
  }
  void *t1, *t2, *t3, *t4;
  CHECK(ls_create_interaction1(&t1, m1, N, s1));
  CHECK(ls_create_interaction2(&t2, m2, N, s2));
  CHECK(ls_create_interaction3(&t3, m3, N, s3));
  CHECK(ls_create_interaction4(&t4, m4, N, s4));
  REQUIRE(ls_interaction_is_real(t1) && ls_interaction_is_real(t2) && ls_interaction_is_real(t4));
  REQUIRE(!ls_interaction_is_real(t3));
  /* a site outside the lattice is the library's job to reject (Internal.hs:107-108): at the latest
   * when the term meets a basis, and without touching the out-parameter */
  {
    uint16_t bad_site[1] = {N};
    void* tb = (void*)0x2;
    int rc = ls_create_interaction1(&tb, m1, 1, bad_site);
    if (rc == 0) {
      void* ob = (void*)0x3;
      void const* one[1] = {tb};
      REQUIRE(ls_create_operator(&ob, basis, 1, one) != 0 && ob == (void*)0x3);
      ls_destroy_interaction(tb);
    } else {
      REQUIRE(tb == (void*)0x2);
    }
  }
  void const* all[4] = {t1, t2, t4, t3};
  CHECK(ls_create_operator(op, basis, with_complex ? 4 : 3, all));
  ls_destroy_interaction(t3); /* interactions are only alive during ls_create_operator (Internal.hs:386-397) */
  ls_destroy_interaction(t1);
  ls_destroy_interaction(t4);
  ls_destroy_interaction(t2);
  return 0;
}

/* <x|H|x> through matmat and through expectation must agree; Hermitian => real */
static int check_operator(void* op, uint64_t n, int complex_dtype) {
  int const dt = complex_dtype ? SPED_C128 : SPED_F64;
  size_t const w = complex_dtype ? 2 : 1;
  double* x = calloc(n * w, sizeof(double));
  double* y = calloc(n * w, sizeof(double));
  for (uint64_t i = 0; i < n * w; ++i) x[i] = sin(0.7 * (double)(i + 1));
  CHECK(ls_operator_matmat(op, dt, n, 1, x, n, y, n));
  double ex[2] = {0, 0}, re = 0, im = 0;
  CHECK(ls_operator_expectation(op, dt, n, 1, x, n, ex));
  for (uint64_t i = 0; i < n; ++i) {
    if (complex_dtype) {
      re += x[2 * i] * y[2 * i] + x[2 * i + 1] * y[2 * i + 1];
      im += x[2 * i] * y[2 * i + 1] - x[2 * i + 1] * y[2 * i];
    } else {
      re += x[i] * y[i];
    }
  }
  REQUIRE(fabs(re - ex[0]) <= 1e-12 * (1 + fabs(re)) && fabs(im - ex[1]) <= 1e-12 * (1 + fabs(re)) && fabs(ex[1]) <= 1e-12 * (1 + fabs(re)));
  free(x);
  free(y);
  return 0;
}

int main(void) {
  REQUIRE(sizeof k_all_symbols / sizeof k_all_symbols[0] == 32);
  ls_enable_logging();
  ls_disable_logging();
  char const* ok = ls_error_to_string(0);
  REQUIRE(ok != NULL);
  ls_destroy_string(ok);

  /* sector k = 1 has complex characters: a real matrix set still gives a complex operator */
  {
    void *basis, *op;
    if (make_basis(&basis, 1)) return 1;
    if (make_operator(&op, basis, 0)) return 1;
    /* real matrices but complex characters: the operator is complex (drives the dtype choice, Main.hs:23-25) */
    REQUIRE(!ls_operator_is_real(op));
    CHECK(ls_build(basis));
    uint64_t n = 0;
    CHECK(ls_get_number_states(basis, &n));
    double dummy[4] = {0};
    REQUIRE(ls_operator_matmat(op, SPED_F64, n, 1, dummy, n, dummy, n) != 0); /* real dtype on a complex operator */
    if (check_operator(op, n, 1)) return 1;
    ls_destroy_operator(op);
    ls_destroy_spin_basis(basis);
  }

  /* every destroy order of {basis, operator, states}; k = 0, all four interaction kinds */
  static int const orders[6][3] = {{0, 1, 2}, {0, 2, 1}, {1, 0, 2}, {1, 2, 0}, {2, 0, 1}, {2, 1, 0}};
  uint64_t reference_reps[64];
  uint64_t n_ref = 0;
  for (int o = 0; o < 6; ++o) {
    void *basis, *op, *states;
    if (make_basis(&basis, 0)) return 1;
    if (make_operator(&op, basis, 1)) return 1; /* created before the build, like SpinED.hs:243-248 */
    REQUIRE(!ls_operator_is_real(op)); /* the 3-site term has an imaginary matrix */
    if (o == 0) { /* same terms without it, real characters (k = 0): a real operator */
      void* real_op;
      if (make_operator(&real_op, basis, 0)) return 1;
      REQUIRE(ls_operator_is_real(real_op));
      ls_destroy_operator(real_op);
    }
    if (o == 0) {
      CHECK(ls_build(basis));
    } else {
      /* resume path: adopt the representatives of the first round (ls_build_unsafe) */
      uint64_t wrong[64];
      memcpy(wrong, reference_reps, n_ref * sizeof(uint64_t));
      wrong[1] = wrong[0]; /* not strictly increasing */
      REQUIRE(ls_build_unsafe(basis, n_ref, wrong) != 0);
      uint64_t n_bad = 0;
      REQUIRE(ls_get_number_states(basis, &n_bad) != 0); /* a failed adoption leaves the basis unbuilt */
      CHECK(ls_build_unsafe(basis, n_ref, reference_reps));
    }
    uint64_t n = 0;
    CHECK(ls_get_number_states(basis, &n));
    CHECK(ls_get_states(&states, basis));
    REQUIRE(ls_states_get_size(states) == n && n > 0 && n <= 64);
    if (o == 0) {
      n_ref = n;
      memcpy(reference_reps, ls_states_get_data(states), n * sizeof(uint64_t));
      REQUIRE(n == 10); /* C(8,4) = 70 states, 8 translations, k = 0: 10 orbits with non-zero norm */
    } else {
      REQUIRE(n == n_ref && memcmp(reference_reps, ls_states_get_data(states), n * sizeof(uint64_t)) == 0);
    }
    int alive[3] = {1, 1, 1};
    for (int step = 0; step < 3; ++step) {
      switch (orders[o][step]) {
        case 0: ls_destroy_spin_basis(basis); break;
        case 1: ls_destroy_operator(op); break;
        default: ls_destroy_states(states); break;
      }
      alive[orders[o][step]] = 0;
      /* whatever is still alive keeps working */
      if (alive[1] && check_operator(op, n, 1)) return 1;
      if (alive[2]) REQUIRE(ls_states_get_data(states)[n - 1] == reference_reps[n - 1]);
      if (alive[0]) {
        uint64_t again = 0;
        CHECK(ls_get_number_states(basis, &again));
        REQUIRE(again == n);
      }
    }
  }
  printf("CONFORMANCE_FULL_OK %llu\n", (unsigned long long)n_ref);
  return 0;
}
