/*
 * ndchg.h -- C ABI of the omnibus change detection, part of libndnlm.so.
 *
 * Replaces the reference's Cython extension entry point
 *     nd/_change.pyx:266   cpdef BOOL[:, :, :] change_detection(floating[:, :, :, :] values, double alpha,
 *                                                                unsigned int n=1, unsigned int njobs=1)
 * called from `_omnibus_change_detection` (nd/change.py:67), i.e. per pixel `single_pixel_change_detection`
 * (:224-260) over `single_pixel_omnibus` (:139-160) and `_z` / `_rho` / `_omega2` / `_f` (:19-79): the
 * Conradsen et al. (2015) omnibus test on dual-pol covariance time series [C11, Re C12, Im C12, C22].
 * The reference parallelises over rows with OpenMP `prange` (:280); here one GPU thread owns one pixel.
 *
 * `values` is a DEVICE array (rows, cols, k, 4) of float32 or float64 with arbitrary element strides;
 * `result` is a DEVICE uint8 array (rows, cols, k), C-contiguous, written completely (0 / 1).
 * `floating` locals of the reference keep the data type (float32 data -> float32 statistic), its doubles stay
 * double.  The chi-square CDF (GSL's gsl_cdf_chisq_P in the reference) is evaluated in double from the closed
 * forms of the regularised incomplete gamma function for the integer shape 2 (k - 1) the test always has.
 * Asynchronous on `stream`; returns 0 or a negative NDNLM_E* code (message: ndchg_last_error()).
 */
#ifndef NDCHG_H
#define NDCHG_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDCHG_F32 0
#define NDCHG_F64 1

int ndchg_change_detection(const void* values, const int64_t shape[3], const int64_t strides[4], int dtype,
                           uint8_t* result, double alpha, uint32_t n, void* stream);

/* The test statistic's probability for every pixel over the whole series (`single_pixel_omnibus`,
 * nd/_change.pyx:139-160): `prob` is a DEVICE array (rows, cols) of the data type, C-contiguous. */
int ndchg_omnibus_probability(const void* values, const int64_t shape[3], const int64_t strides[4], int dtype,
                              void* prob, uint32_t n, void* stream);

const char* ndchg_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NDCHG_H */
