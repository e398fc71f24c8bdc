/*
 * ndflt.h -- C ABI of the sibling filters of nd.filters, part of libndnlm.so.
 *
 * The reference's ConvolutionFilter / BoxcarFilter / GaussianFilter (nd/filters.py:205-381) are thin
 * wrappers over scipy.ndimage (`snf.convolve`, nd/filters.py:260-268; `snf.gaussian_filter`,
 * nd/filters.py:370-378).  scipy is a dependency of the reference, not part of its tree; the pinned
 * algorithm restated here is scipy 1.x `ndimage/src/ni_filters.c`:
 *
 *   NI_Correlate     out[p] = sum over the footprint taps k (C order, |w_k| > DBL_EPSILON) of
 *                    in[ext(p + k - size/2 - origin)] * w_k, accumulated in double with a separate
 *                    multiply and add per tap, starting from 0.0;
 *   NI_Correlate1D   the same along one axis; odd symmetric kernels use
 *                    tmp = x[0] w[0] + sum_{j=-size1}^{-1} (x[j] + x[-j]) w[j]   (antisymmetric: x[j] - x[-j]),
 *                    general kernels tmp = x[size2] w[size2] + sum_{j=-size1}^{size2-1} x[j] w[j];
 *   NI_ExtendLine    boundary modes reflect (d c b a | a b c d | d c b a), constant, nearest, mirror, wrap.
 *
 * The CUDA kernels reproduce that arithmetic operation by operation (no FMA contraction), so float64
 * results are bit-identical to scipy's; float32 data is accumulated in double and rounded once, as in scipy.
 *
 * All pointers `in` / `out` are DEVICE pointers; `weights` is a HOST array.  Arrays have up to 4 axes with
 * arbitrary element strides (unused axes: shape 1).  `in` and `out` must not overlap.  Every call is
 * asynchronous on `stream` (a cudaStream_t, may be NULL).  Return 0 or a negative NDNLM_E* code
 * (include/ndnlm.h); the message is in ndflt_last_error().
 */
#ifndef NDFLT_H
#define NDFLT_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NDFLT_F32 0
#define NDFLT_F64 1

/* scipy.ndimage boundary modes ('grid-mirror' = reflect, 'grid-constant' = constant, 'grid-wrap' = wrap) */
#define NDFLT_MODE_REFLECT  0
#define NDFLT_MODE_CONSTANT 1
#define NDFLT_MODE_NEAREST  2
#define NDFLT_MODE_MIRROR   3
#define NDFLT_MODE_WRAP     4

/*
 * N-D correlation (scipy `_nd_image.correlate`, the engine below `snf.convolve` as called from
 * ConvolutionFilter._filter, nd/filters.py:260-268; the caller flips the kernel and negates the origins
 * for a convolution exactly like scipy's `_correlate_or_convolve`).
 *   weights   host, C order, kshape[0..3] doubles
 *   origin    per axis, scipy convention: tap k of an axis reads in[i + k - kshape/2 - origin]
 */
int ndflt_correlate(const void* in, void* out, const int64_t shape[4], const int64_t in_strides[4],
                    const int64_t out_strides[4], int dtype, const double* weights, const int64_t kshape[4],
                    const int64_t origin[4], int mode, double cval, void* stream);

/*
 * 1-D correlation along `axis` (scipy `_nd_image.correlate1d`, the engine below `snf.gaussian_filter`
 * as called from GaussianFilter._filter, nd/filters.py:370-378).
 */
int ndflt_correlate1d(const void* in, void* out, const int64_t shape[4], const int64_t in_strides[4],
                      const int64_t out_strides[4], int dtype, int axis, const double* weights, int64_t nweights,
                      int64_t origin, int mode, double cval, void* stream);

/* Number of kernel launches issued by the ndflt entry points since load. */
int64_t ndflt_launch_count(void);

const char* ndflt_last_error(void);

#ifdef __cplusplus
}
#endif
#endif /* NDFLT_H */
