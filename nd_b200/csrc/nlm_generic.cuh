// nlm_generic.cuh -- reference-faithful one-thread-per-voxel kernel, any r / f / V, float or double.
//
// A GPU twin of the loops of nd/_filters.pyx:351-420 with the reference's own arithmetic:
// differences and squares in the data type, d^2 / weights / weight sums in float64, the
// per-variable `weighted_sum` in the data type rounded after every neighbour (:336, :400).
// Used for float64 data, for configurations the tiled kernel has no instantiation for, for
// `reference_compiled` semantics (zero_dist) and as the on-device cross-check of the tiled kernel.
#pragma once
#include "nlm_common.cuh"

namespace ndnlm {

constexpr int GENERIC_VC = 8;   // variables accumulated per pass (larger V: several passes)

template <typename T>
__global__ void __launch_bounds__(256)
nlm_generic_kernel(const DevParams P, const T* __restrict__ padded, T* __restrict__ out, int* __restrict__ err) {
    const long long total = (long long)P.n[0] * P.n[1] * P.n[2];
    const long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= total) return;
    const int x = int(idx % P.n[2]);
    const int r = int((idx / P.n[2]) % P.n[1]);
    const int w = int(idx / ((long long)P.n[2] * P.n[1]));
    const int V = P.V;
    const long long sX = V, sR = (long long)P.pd[2] * V, sW = (long long)P.pd[1] * P.pd[2] * V;
    const T* pc = padded + (w + P.pad[0]) * sW + (r + P.pad[1]) * sR + (x + P.pad[2]) * sX;
    const double norm = double(T(double(V) * (2 * P.fr[0] + 1) * (2 * P.fr[1] + 1) * (2 * P.fr[2] + 1)));   // :337 (floating)
    const double h2 = 1.0 / P.inv_h2;

    for (int v0 = 0; v0 < V; v0 += GENERIC_VC) {
        const int nv = min(GENERIC_VC, V - v0);
        T wsum[GENERIC_VC];
#pragma unroll
        for (int v = 0; v < GENERIC_VC; ++v) wsum[v] = T(0);
        double total_w = 0.0, total_sq = 0.0, max_w = 0.0;

        for (int tw = -P.rad[0]; tw <= P.rad[0]; ++tw)
            for (int tr = -P.rad[1]; tr <= P.rad[1]; ++tr)
                for (int tx = -P.rad[2]; tx <= P.rad[2]; ++tx) {
                    if (tw == 0 && tr == 0 && tx == 0) continue;
                    const T* pq = pc + tw * sW + tr * sR + tx * sX;
                    double dsq = 0.0;
                    if (!P.zero_dist) {
                        for (int dw = -P.fr[0]; dw <= P.fr[0]; ++dw)
                            for (int dr = -P.fr[1]; dr <= P.fr[1]; ++dr)
                                for (int dx = -P.fr[2]; dx <= P.fr[2]; ++dx) {
                                    const long long off = dw * sW + dr * sR + dx * sX;
                                    for (int v = 0; v < V; ++v) {
                                        const T df = pc[off + v] - pq[off + v];
                                        dsq += double(T(df * df));
                                    }
                                }
                    }
                    dsq /= norm;
                    double a = dsq - P.two_sigma2;
                    a = (0.0 > a) ? 0.0 : a;           // NaN propagates, as in the reference
                    const double wgt = exp(-a / h2);
                    total_w += wgt;
                    total_sq += wgt * wgt;
                    if (wgt > max_w) max_w = wgt;
#pragma unroll
                    for (int v = 0; v < GENERIC_VC; ++v)
                        if (v < nv) wsum[v] = T(double(wsum[v]) + wgt * double(pq[v0 + v]));
                }

        double ws;
        if (P.n_eff < 0) {
            if (max_w == 0.0) max_w = 1.0;
            ws = max_w;
        } else {
            if (P.n_eff - 1.0 > total_w * total_w / total_sq) atomicOr(err, 1);
            ws = (total_w + sqrt(P.n_eff * total_w * total_w - P.n_eff * P.n_eff * total_sq + P.n_eff * total_sq)) /
                 (P.n_eff - 1.0);
        }
        total_w += ws;
        T* po = out + idx * V + v0;
#pragma unroll
        for (int v = 0; v < GENERIC_VC; ++v)
            if (v < nv) {
                const T s = T(double(wsum[v]) + ws * double(pc[v0 + v]));
                po[v] = T(double(s) / total_w);
            }
    }
}

}  // namespace ndnlm
