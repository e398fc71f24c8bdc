/*
 * TEST INFRASTRUCTURE ONLY -- not product code.
 *
 * Plain-C restatement of the reference NLM kernel, jnhansen/nd nd/_filters.pyx:317-420
 * (`_pixelwise_nlmeans_3d`) with its helpers `_idx` (:15-40) and `find_weight` (:297-314).
 * Loop order, arithmetic types and rounding points follow the reference statement by statement:
 *   - differences and squares in the data type T, accumulated into a double `dsquare`   (:372-386)
 *   - `dsquare /= dsq_norm` with dsq_norm of type T                                      (:337, :388)
 *   - weight = exp(-max(dsquare - 2 sigma^2, 0) / h^2) in double; `max` is `0 > x ? 0 : x`  (:391)
 *   - total_weight / total_sq_weight / max_weight in double                               (:393-397)
 *   - weighted_sum[v] of type T, rounded after every neighbour                            (:336, :399-403)
 *   - self weight: max weight (1 if 0) or find_weight                                     (:406-413)
 *   - output = weighted_sum / total_weight                                                (:415-420)
 * `compiled_bug != 0` reproduces the LP64 binary: because `f` is `unsigned int`
 * (:323) the patch loops `range(-f[i], f[i]+1)` (:373-375) do not execute when f[i] > 0, so
 * dsquare stays 0.  Pinned against the reference's own compiled kernel by tests/test_oracle.py.
 *
 * Strides are in ELEMENTS.  Returns 0, or 1 if find_weight had no solution (the reference raises
 * ValueError('No solution') at that voxel; here the scan stops there as well).
 * `roi` (may be NULL) restricts the OUTPUT voxels to the box [roi[0],roi[1]) x [roi[2],roi[3]) x [roi[4],roi[5]);
 * the arithmetic of every computed voxel is unchanged (bench.py uses it to check sampled sub-cubes of a large cube).
 * Build: gcc -O3 -fopenmp -shared -fPIC nlm_oracle.c -o libnlm_oracle.so -lm  (OpenMP only splits
 * the outermost voxel loop; every voxel is computed exactly as in the serial reference.)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static inline int64_t idx_reflect(int64_t i, int64_t n) { /* nd/_filters.pyx:34-40 */
    if (i < 0) return -i;
    if (i >= n) return 2 * n - 2 - i;
    return i;
}

#define DEFINE_ORACLE(NAME, T)                                                                          \
    int NAME(const T* arr, T* out, const int64_t shape[4], const int64_t as[4], const int64_t os[4],     \
             const uint32_t r[3], const uint32_t f[3], double sigma, double h, double n_eff,             \
             int compiled_bug, const int64_t* roi) {                                                     \
        const int64_t N0 = shape[0], N1 = shape[1], N2 = shape[2], V = shape[3];                        \
        const T dsq_norm = (T)(V * (2 * (int64_t)f[0] + 1) * (2 * (int64_t)f[1] + 1) * (2 * (int64_t)f[2] + 1)); \
        const int skip_patch = compiled_bug && (f[0] > 0 || f[1] > 0 || f[2] > 0);                       \
        int failed = 0;                                                                                  \
        const int64_t a0 = roi ? roi[0] : 0, a1 = roi ? roi[1] : N0, b0 = roi ? roi[2] : 0,              \
                      b1 = roi ? roi[3] : N1, c0 = roi ? roi[4] : 0, c1 = roi ? roi[5] : N2;             \
        _Pragma("omp parallel for schedule(dynamic, 1)")                                                 \
        for (int64_t p0 = 0; p0 < (a1 - a0) * (b1 - b0); ++p0) {                                         \
            const int64_t pa = a0 + p0 / (b1 - b0), pb = b0 + p0 % (b1 - b0);                            \
            T* wsum = (T*)malloc(sizeof(T) * (size_t)V);                                                 \
            for (int64_t pc = c0; pc < c1 && !failed; ++pc) {                                            \
                double total_w = 0, total_sq = 0, max_w = 0;                                             \
                for (int64_t v = 0; v < V; ++v) wsum[v] = 0;                                             \
                for (int64_t qa = pa - r[0]; qa <= pa + (int64_t)r[0]; ++qa)                             \
                    for (int64_t qb = pb - r[1]; qb <= pb + (int64_t)r[1]; ++qb)                         \
                        for (int64_t qc = pc - r[2]; qc <= pc + (int64_t)r[2]; ++qc) {                   \
                            if (qa == pa && qb == pb && qc == pc) continue;                              \
                            double dsq = 0;                                                              \
                            if (!skip_patch)                                                             \
                                for (int64_t da = -(int64_t)f[0]; da <= (int64_t)f[0]; ++da)             \
                                    for (int64_t db = -(int64_t)f[1]; db <= (int64_t)f[1]; ++db)         \
                                        for (int64_t dc = -(int64_t)f[2]; dc <= (int64_t)f[2]; ++dc) {   \
                                            const int64_t po = idx_reflect(pa + da, N0) * as[0] +        \
                                                               idx_reflect(pb + db, N1) * as[1] +        \
                                                               idx_reflect(pc + dc, N2) * as[2];         \
                                            const int64_t qo = idx_reflect(qa + da, N0) * as[0] +        \
                                                               idx_reflect(qb + db, N1) * as[1] +        \
                                                               idx_reflect(qc + dc, N2) * as[2];         \
                                            for (int64_t v = 0; v < V; ++v) {                            \
                                                const T df = arr[po + v * as[3]] - arr[qo + v * as[3]];  \
                                                dsq += (double)(T)(df * df);                             \
                                            }                                                            \
                                        }                                                                \
                            dsq /= (double)dsq_norm;                                                     \
                            double a = dsq - 2 * sigma * sigma;                                          \
                            a = (0 > a) ? 0 : a;                                                         \
                            const double w = exp(-a / (h * h));                                          \
                            total_w += w;                                                                \
                            total_sq += w * w;                                                           \
                            if (w > max_w) max_w = w;                                                    \
                            const int64_t qo = idx_reflect(qa, N0) * as[0] + idx_reflect(qb, N1) * as[1] + \
                                               idx_reflect(qc, N2) * as[2];                              \
                            for (int64_t v = 0; v < V; ++v)                                              \
                                wsum[v] = (T)((double)wsum[v] + w * (double)arr[qo + v * as[3]]);        \
                        }                                                                                \
                double ws;                                                                               \
                if (n_eff < 0) {                                                                         \
                    if (max_w == 0) max_w = 1;                                                           \
                    ws = max_w;                                                                          \
                } else {                                                                                 \
                    if (n_eff - 1 > total_w * total_w / total_sq) {                                      \
                        failed = 1;                                                                      \
                        break;                                                                           \
                    }                                                                                    \
                    ws = (total_w + sqrt(n_eff * total_w * total_w - n_eff * n_eff * total_sq + n_eff * total_sq)) / \
                         (n_eff - 1);                                                                    \
                }                                                                                        \
                total_w += ws;                                                                           \
                const int64_t pi = pa * as[0] + pb * as[1] + pc * as[2];                                 \
                const int64_t oi = pa * os[0] + pb * os[1] + pc * os[2];                                 \
                for (int64_t v = 0; v < V; ++v) {                                                        \
                    wsum[v] = (T)((double)wsum[v] + ws * (double)arr[pi + v * as[3]]);                   \
                    out[oi + v * os[3]] = (T)((double)wsum[v] / total_w);                                \
                }                                                                                        \
            }                                                                                            \
            free(wsum);                                                                                  \
        }                                                                                                \
        return failed;                                                                                   \
    }

DEFINE_ORACLE(nlm_oracle_f32, float)
DEFINE_ORACLE(nlm_oracle_f64, double)
