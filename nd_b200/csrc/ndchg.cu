// ndchg.cu -- C ABI (include/ndchg.h) of the omnibus change detection: one thread per pixel walks the reference's
// control flow (nd/_change.pyx:224-260) with the reference's arithmetic types.  The marginal tests of one starting
// point l grow by one time step at a time, so their sums / determinant product are kept running (the same
// left-to-right accumulation order as recomputing every subset from scratch, hence identical values).
#include "../../include/ndchg.h"
#include "../../include/ndnlm.h"

#include <cstdarg>
#include <cstdio>

#include <cuda_runtime.h>

static thread_local char g_chg_err[512] = "";
static int chg_fail(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_chg_err, sizeof(g_chg_err), fmt, ap);
    va_end(ap);
    return code;
}

namespace ndchg {

// separate, correctly rounded operations in the data type (no FMA contraction: the reference is plain C)
__device__ __forceinline__ float mul_(float a, float b) { return __fmul_rn(a, b); }
__device__ __forceinline__ float add_(float a, float b) { return __fadd_rn(a, b); }
__device__ __forceinline__ float sub_(float a, float b) { return __fsub_rn(a, b); }
__device__ __forceinline__ double mul_(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double add_(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double sub_(double a, double b) { return __dsub_rn(a, b); }

// Regularised lower incomplete gamma P(a, y) for INTEGER a >= 1 and y > 0, in double:
//   y <= a:  P = e^-y y^a / a!  * sum_{m>=0} y^m / ((a+1)...(a+m))
//   y >  a:  P = 1 - e^-y y^(a-1) / (a-1)! * sum_{i=0}^{a-1} (a-1)...(a-i) / y^i
__device__ double gamma_p_int(const int a, const double y) {
    if (y <= double(a)) {
        double term = 1.0, sum = 1.0;
        for (int m = 1; m < 100000; ++m) {
            term *= y / double(a + m);
            sum += term;
            if (term < sum * 1e-17) break;
        }
        return exp(double(a) * log(y) - y - lgamma(double(a) + 1.0)) * sum;
    }
    double term = 1.0, sum = 1.0;
    for (int i = 1; i < a; ++i) {
        term *= double(a - i) / y;
        sum += term;
        if (term < sum * 1e-17) break;
    }
    return 1.0 - exp(double(a - 1) * log(y) - y - lgamma(double(a))) * sum;
}

// gsl_cdf_chisq_P(x, nu) with nu = 2 a
__device__ __forceinline__ double chisq_p(const double x, const int a) {
    if (x != x) return x;
    if (!(x > 0.0)) return 0.0;
    return gamma_p_int(a, 0.5 * x);
}

// The double-precision constants, operation by operation in the reference's order (explicit intrinsics: no FMA
// contraction, `x**2` is pow(x, 2) = the correctly rounded x*x).
__device__ __forceinline__ double rho_(const double p, const double k, const double n) {
    // 1 - (2 * p**2 - 1) / (6 * (k - 1) * p) * (k/n - 1/(n*k))                              nd/_change.pyx:26-30
    const double A = __dsub_rn(__dmul_rn(2.0, __dmul_rn(p, p)), 1.0);
    const double B = __dmul_rn(__dmul_rn(6.0, __dsub_rn(k, 1.0)), p);
    const double C = __dsub_rn(__ddiv_rn(k, n), __ddiv_rn(1.0, __dmul_rn(n, k)));
    return __dsub_rn(1.0, __dmul_rn(__ddiv_rn(A, B), C));
}
__device__ __forceinline__ double omega2_(const double p, const double k, const double n, const double rho) {
    // p**2 * (p**2 - 1) / (24 * rho**2) * (k/(n**2) - 1/((n*k)**2)) - p**2 * (k - 1) / 4 * (1 - 1/rho)**2   :33-39
    const double p2 = __dmul_rn(p, p);
    const double nk = __dmul_rn(n, k);
    const double left = __dmul_rn(__ddiv_rn(__dmul_rn(p2, __dsub_rn(p2, 1.0)), __dmul_rn(24.0, __dmul_rn(rho, rho))),
                                  __dsub_rn(__ddiv_rn(k, __dmul_rn(n, n)), __ddiv_rn(1.0, __dmul_rn(nk, nk))));
    const double x = __dsub_rn(1.0, __ddiv_rn(1.0, rho));
    const double right = __dmul_rn(__ddiv_rn(__dmul_rn(p2, __dsub_rn(k, 1.0)), 4.0), __dmul_rn(x, x));
    return __dsub_rn(left, right);
}

// running state of `_z` (nd/_change.pyx:45-79) over a growing subset
template <typename T>
struct Acc {
    T c11, c12r, c12i, c22;
    double prod;
    __device__ void reset() { c11 = c12r = c12i = c22 = T(0); prod = 1.0; }
    __device__ void push(const T a, const T br, const T bi, const T d) {
        const T det = sub_(mul_(a, d), add_(mul_(br, br), mul_(bi, bi)));
        prod = __dmul_rn(prod, double(det));
        c11 = add_(c11, a);
        c12r = add_(c12r, br);
        c12i = add_(c12i, bi);
        c22 = add_(c22, d);
    }
    // single_pixel_omnibus (nd/_change.pyx:139-160) of the k steps pushed so far
    __device__ T probability(const int k, const unsigned n) const {
        const double kd = double(k), nd = double(n);
        const T det_of_sum = sub_(mul_(c11, c22), add_(mul_(c12r, c12r), mul_(c12i, c12i)));
        const T pk = mul_(T(2), T(k));                                                     // `p*k` in `floating`
        // n * (p*k*log(k) + log(prod_of_dets) - k*log(det_of_sum))                            nd/_change.pyx:74
        const double logQ = __dmul_rn(nd, __dsub_rn(__dadd_rn(__dmul_rn(double(pk), log(kd)), log(prod)),
                                                    __dmul_rn(kd, log(double(det_of_sum)))));
        const double rho = rho_(2.0, kd, nd);
        const T rho_t = T(rho);
        const T z = T(__dmul_rn(double(mul_(T(-2), rho_t)), logQ));
        const double omega2 = omega2_(2.0, kd, nd, rho);
        const int a = 2 * (k - 1);                                                         // f / 2 with f = (k-1) p^2
        const T P1 = T(chisq_p(double(z), a));
        const T P2 = T(chisq_p(double(z), a + 2));
        return T(__dadd_rn(double(P1), __dmul_rn(omega2, double(sub_(P2, P1)))));
    }
};

template <typename T>
__global__ void __launch_bounds__(128)
change_detection_kernel(const T* __restrict__ values, const long long rows, const long long cols, const int k,
                        const long long s0, const long long s1, const long long s2, const long long s3,
                        unsigned char* __restrict__ result, const double alpha, const unsigned n) {
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= rows * cols) return;
    const long long i = pix / cols, j = pix - i * cols;
    const T* ts = values + i * s0 + j * s1;
    unsigned char* res = result + pix * k;
    for (int t = 0; t < k; ++t) res[t] = 0;
    auto push = [&](Acc<T>& acc, const int t) {
        const T* q = ts + (long long)t * s2;
        acc.push(q[0], q[s3], q[2 * s3], q[3 * s3]);
    };
    int l = 0, r = 0;
    Acc<T> acc;
    while (true) {                                                                         // nd/_change.pyx:238-260
        acc.reset();
        for (int t = l; t < k; ++t) push(acc, t);
        if (!(double(acc.probability(k - l, n)) > alpha)) break;                           // global hypothesis H0_l
        acc.reset();
        push(acc, l);
        for (int jj = 2; jj < k - l + 1; ++jj) {                                           // marginal hypotheses
            push(acc, l + jj - 1);
            r = jj - 1;
            if (double(acc.probability(jj, n)) > alpha) {
                res[l + r] = 1;
                break;
            }
        }
        l = l + r;
        if (l >= k - 1) break;
    }
}

template <typename T>
__global__ void __launch_bounds__(128)
omnibus_probability_kernel(const T* __restrict__ values, const long long rows, const long long cols, const int k,
                           const long long s0, const long long s1, const long long s2, const long long s3,
                           T* __restrict__ prob, const unsigned n) {
    const long long pix = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (pix >= rows * cols) return;
    const long long i = pix / cols, j = pix - i * cols;
    const T* ts = values + i * s0 + j * s1;
    Acc<T> acc;
    acc.reset();
    for (int t = 0; t < k; ++t) {
        const T* q = ts + (long long)t * s2;
        acc.push(q[0], q[s3], q[2 * s3], q[3 * s3]);
    }
    prob[pix] = acc.probability(k, n);
}

}  // namespace ndchg

struct ChgDeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit ChgDeviceGuard(const void* ptr) {
        cudaPointerAttributes attr;
        if (cudaPointerGetAttributes(&attr, ptr) != cudaSuccess || attr.type != cudaMemoryTypeDevice) {
            cudaGetLastError();
            ok = false;
            return;
        }
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != attr.device && cudaSetDevice(attr.device) != cudaSuccess) ok = false;
        if (prev == attr.device) prev = -1;
    }
    ~ChgDeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

static int check_args(const void* values, const int64_t* shape, const int64_t* strides, int dtype, const void* out, uint32_t n) {
    if (!values || !shape || !strides || !out) return chg_fail(NDNLM_EINVAL, "null argument");
    if (dtype != NDCHG_F32 && dtype != NDCHG_F64) return chg_fail(NDNLM_EDTYPE, "No matching signature found (only float32 / float64 data is supported)");
    if (shape[0] < 1 || shape[1] < 1 || shape[0] * shape[1] > 0x7fffffffLL * 128) return chg_fail(NDNLM_EINVAL, "shape out of range");
    if (shape[2] < 2 || shape[2] > 0x7fffffffLL) return chg_fail(NDNLM_EINVAL, "the omnibus test needs at least 2 time steps");
    if (n < 1) return chg_fail(NDNLM_EINVAL, "the number of looks n must be >= 1");
    return NDNLM_OK;
}

extern "C" int ndchg_change_detection(const void* values, const int64_t shape[3], const int64_t strides[4], int dtype,
                                      uint8_t* result, double alpha, uint32_t n, void* stream) {
    int rc = check_args(values, shape, strides, dtype, result, n);
    if (rc) return rc;
    ChgDeviceGuard guard(result);
    if (!guard.ok) return chg_fail(NDNLM_EINVAL, "result is not a CUDA device pointer");
    const long long pixels = shape[0] * shape[1];
    const unsigned grid = unsigned((pixels + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NDCHG_F64)
        ndchg::change_detection_kernel<double><<<grid, 128, 0, st>>>((const double*)values, shape[0], shape[1], int(shape[2]),
                                                                    strides[0], strides[1], strides[2], strides[3], result, alpha, n);
    else
        ndchg::change_detection_kernel<float><<<grid, 128, 0, st>>>((const float*)values, shape[0], shape[1], int(shape[2]),
                                                                   strides[0], strides[1], strides[2], strides[3], result, alpha, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return chg_fail(NDNLM_ECUDA, "launch failed: %s", cudaGetErrorString(e));
    return NDNLM_OK;
}

extern "C" int ndchg_omnibus_probability(const void* values, const int64_t shape[3], const int64_t strides[4], int dtype,
                                         void* prob, uint32_t n, void* stream) {
    int rc = check_args(values, shape, strides, dtype, prob, n);
    if (rc) return rc;
    ChgDeviceGuard guard(prob);
    if (!guard.ok) return chg_fail(NDNLM_EINVAL, "prob is not a CUDA device pointer");
    const long long pixels = shape[0] * shape[1];
    const unsigned grid = unsigned((pixels + 127) / 128);
    cudaStream_t st = (cudaStream_t)stream;
    if (dtype == NDCHG_F64)
        ndchg::omnibus_probability_kernel<double><<<grid, 128, 0, st>>>((const double*)values, shape[0], shape[1], int(shape[2]),
                                                                       strides[0], strides[1], strides[2], strides[3], (double*)prob, n);
    else
        ndchg::omnibus_probability_kernel<float><<<grid, 128, 0, st>>>((const float*)values, shape[0], shape[1], int(shape[2]),
                                                                      strides[0], strides[1], strides[2], strides[3], (float*)prob, n);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return chg_fail(NDNLM_ECUDA, "launch failed: %s", cudaGetErrorString(e));
    return NDNLM_OK;
}

extern "C" const char* ndchg_last_error(void) { return g_chg_err; }
