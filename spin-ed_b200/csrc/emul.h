/* emul.h -- entry point of the TEST-ONLY library libsped_emul.so (csrc/emul.cpp).  Not part of the
 * product: libsped.so neither contains nor calls it, and include/ does not declare it. */
#ifndef SPED_EMUL_H
#define SPED_EMUL_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif
/* Single-threaded HOST emulation of the matvec kernels (csrc/emul.cpp: the device sources compiled
 * by the host compiler with the CUDA built-ins shimmed) on a small problem whose representatives
 * and stabiliser sizes the caller supplies (the test-suite takes them from the oracle): y = H x for
 * the rows rank `rank` of `world` owns, computed by the matrix-free row routine (y_free), by the
 * streaming kernel over all source classes at once (y_all) and class by class (y_phased).
 * dtype: 1 = f64, 3 = c128; x in global row order.  stats[6] = slots, stored elements, elements
 * with the default coefficient, source classes, entries of the compact code stream, bytes per code.  With ncols = 2..4 the block kernel is run as well
 * on the columns x_c[g] = x[(g + c) mod n] (y_block: n_local x ncols, column-major).
 * Verification only: nothing in the product calls it. */
int sped_selftest_emulate_matvec(void const* op, uint64_t n, uint64_t const* reps, uint16_t const* stab, int world, int rank,
                                 int dtype, void const* x_global, void* y_free, void* y_all, void* y_phased,
                                 uint64_t* stats, unsigned ncols, void* y_block);
/* Host emulation of the fused kernels of the single-pair eigensolver iteration (csrc/eigh_kernels.cuh):
 * restart_residual_kernel on V, W (n x m, column-major, leading dimension ld, in place; V needs m
 * columns, column p receives the residual), then axpy_norm_kernel and scale_rel_kernel on that column.
 * C: m x p complex, row-major, (re, im) pairs.  out: (re, im) of |r|^2 and of <V'_q, r> for q < p,
 * then |t|^2, the kept share |t|^2 / |r|^2, and the DGKS flag. */
int sped_selftest_emulate_restart(int dtype, uint64_t n, int m, int p, void* V, void* W, uint64_t ld,
                                  double const* C_re_im, double theta, double* out);
#ifdef __cplusplus
}
#endif
#endif
