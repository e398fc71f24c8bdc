// nlm_boxmean.cuh -- fast path for `reference_compiled` semantics.
//
// The reference as it compiles on LP64 never runs its patch loops when any f_i > 0 (`f` is `unsigned int[:]`,
// nd/_filters.pyx:323; `range(-f[i], f[i]+1)`, :373-375): d^2 == 0, every weight is exp(0) == 1, the weight sum is
// K = prod(2 r_i + 1) - 1 and the self weight is the maximum weight, 1 (:406-409).  The output is therefore the
// reflect BOX MEAN over the (2r+1)^3 search window, (sum_{q != p} a_q + a_p) / (K + 1) (:399-420) -- HBM-bound
// (8V bytes per voxel), not FP32-bound.  Two kernels over the staged reflect-padded cube [q][W][X][R] float4:
//
//   boxmean_xr_kernel  one thread per (W row, X segment, pair of R positions): marches along X with the last
//                      2 r_X + 1 raw values in a register ring (every input element is loaded once, 512-byte
//                      coalesced runs along R), sums the ring, then sums along R across lanes with shuffles;
//   boxmean_w_kernel   one thread per (X, R) element: marches along W with a register ring of 2 r_W + 1 rows,
//                      fully coalesced, and scales by 1 / (K + 1).
//
// Both are one coalesced read + one coalesced write of the cube.  (Running the pair slab by slab so that the
// intermediate stays in the 126 MB L2 -- NDNLM_BOXMEAN_SLAB=rows, a tuning aid -- is slower at every slab height: a
// slab small enough for L2, a cfg3 row being 2 MB, has too few rows to fill 148 SMs)
// The reference accumulates `weighted_sum` in float32 in (y, x, t) loop order; a separable float32 sum differs
// from it by its own rounding noise (~1e-7 scaled, tests allow 1e-5).
#pragma once
#include "nlm_common.cuh"

namespace ndnlm {

__device__ __forceinline__ float4 f4add(const float4 a, const float4 b) {
    return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w);
}
__device__ __forceinline__ float4 f4sub(const float4 a, const float4 b) {
    return make_float4(a.x - b.x, a.y - b.y, a.z - b.z, a.w - b.w);
}
__device__ __forceinline__ float4 f4shfl(const float4 a, const int src_lane) {
    return make_float4(__shfl_sync(0xffffffffu, a.x, src_lane), __shfl_sync(0xffffffffu, a.y, src_lane),
                       __shfl_sync(0xffffffffu, a.z, src_lane), __shfl_sync(0xffffffffu, a.w, src_lane));
}

struct BoxParams {
    int n[3], rad[3], pad[3], pd[3];   // role order (W, R, X)
    int nv4;
    int w_first, w_count;              // rows of the intermediate: padded W rows [w_first, w_first + w_count)
    int seg_len, nseg;                 // X segments of the march (xr kernel)
    int nwr;                           // warps along R (xr kernel)
    int out_first, out_count;          // output rows [out_first, out_first + out_count) (w kernel)
    int w_seg, nwseg;                  // W segments of the march (w kernel)
    float scale;                       // 1 / (K + 1)
};

// intermediate layout: [q][w_count][n2 (X)][n1 (R)] float4
// One R position per lane (a warp covers 32 consecutive padded R positions, 32 - 2 r_R of them produce output), one
// W row and one X segment per warp.  The next RING input values are prefetched into registers while the current
// group is processed: ~100 registers per thread, 18+ warps per SM, ~100 KB of loads in flight per SM.
template <int RING>
__global__ void __launch_bounds__(128)     // (128, 4) = 128 registers costs the prefetch depth: 5.2 instead of 4.2 ms on cfg3
boxmean_xr_kernel(const BoxParams B, const float4* __restrict__ padded, float4* __restrict__ inter) {
    const int lane = threadIdx.x & 31;
    long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int kr = int(wid % B.nwr);
    wid /= B.nwr;
    const int seg = int(wid % B.nseg);
    wid /= B.nseg;
    const int wrow = int(wid % B.w_count);
    const int q = int(wid / B.w_count);
    if (q >= B.nv4) return;
    const int rR = B.rad[1], rX = B.rad[2];
    const int VO = 32 - 2 * rR;                          // R positions produced per warp
    const int r0 = kr * VO + B.pad[1] - rR + lane;       // my padded R position
    const bool ld = r0 >= 0 && r0 < B.pd[1];
    const int ro = r0 - B.pad[1];                        // my output R index
    const bool st = lane >= rR && lane < 32 - rR && ro >= 0 && ro < B.n[1];
    const int xs = seg * B.seg_len, xe = min(xs + B.seg_len, B.n[2]);
    if (xs >= xe) return;
    const float4* src = padded + ((size_t(q) * B.pd[0] + (B.w_first + wrow)) * B.pd[2]) * B.pd[1] + (ld ? r0 : 0);
    float4* dst = inter + ((size_t(q) * B.w_count + wrow) * B.n[2]) * B.n[1] + (st ? ro : 0);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    const size_t rp = size_t(B.pd[1]);

    // sum along R across lanes, then store
    auto finish = [&](const int x, const float4 s) {
        float4 o = s;
        if (rR == 2) {                                   // ((v[i-2] + v[i-1]) + (v[i] + v[i+1])) + v[i+2], 3 shuffles
            const float4 pr = f4add(s, f4shfl(s, lane + 1));
            o = f4add(f4add(f4shfl(pr, lane - 2), pr), f4shfl(s, lane + 2));
        } else {
            for (int d = 1; d <= rR; ++d) o = f4add(o, f4add(f4shfl(s, lane - d), f4shfl(s, lane + d)));
        }
        if (st) dst[size_t(x) * B.n[1]] = o;
    };

    float4 ring[RING];
    // ring slot k holds the padded X position first + k (first = x' - rX of the first output, x' = xs + pad_X)
    const float4* p = src + size_t(xs + B.pad[2] - rX) * rp;
#pragma unroll
    for (int k = 0; k < RING; ++k) {
        ring[k] = zero;
        if (k < 2 * rX && ld) ring[k] = __ldg(p + size_t(k) * rp);
    }
    p += size_t(2 * rX) * rp;                            // next position to load
    const int P = 2 * rX + 1;
    int slot = (2 * rX) % P;
    int x = xs;
    if (P == RING) {
        // groups of RING steps with compile-time slots.  The window sum is recomputed from the ring at the first step
        // of every group and updated incrementally (minus the value that leaves, plus the one that enters) in between,
        // so rounding errors cannot accumulate over more than RING steps.
        float4 nxt[RING];
#pragma unroll
        for (int u = 0; u < RING; ++u) nxt[u] = (ld && x + u < xe) ? __ldg(p + size_t(u) * rp) : zero;
        float4 s = zero;
        for (; x + RING <= xe; x += RING) {
            p += size_t(RING) * rp;
#pragma unroll
            for (int u = 0; u < RING; ++u) {             // slot of step j = x - xs + u is (j + RING - 1) mod RING
                const int sl = (u + RING - 1) % RING;     // compile-time after unrolling (x - xs is a multiple of RING)
                const float4 v = nxt[u];
                nxt[u] = (ld && x + RING + u < xe) ? __ldg(p + size_t(u) * rp) : zero;   // prefetch for the next group
                if (u == 0) {
                    ring[sl] = v;
                    s = zero;
#pragma unroll
                    for (int k = 0; k < RING; ++k) s = f4add(s, ring[k]);
                } else {
                    s = f4add(f4sub(s, ring[sl]), v);
                    ring[sl] = v;
                }
                finish(x + u, s);
            }
        }
        slot = (RING - 1) % RING;                         // x - xs is a multiple of RING again
        // the tail (< RING steps) continues from the prefetched values
#pragma unroll
        for (int u = 0; u < RING; ++u) {
            if (x + u < xe) {
                float4 sum = zero;
#pragma unroll
                for (int k = 0; k < RING; ++k) {
                    if (k == slot) ring[k] = nxt[u];
                    sum = f4add(sum, ring[k]);
                }
                slot = (slot + 1 == P) ? 0 : slot + 1;
                finish(x + u, sum);
            }
        }
        return;
    }
    for (; x < xe; ++x) {
        const float4 v = ld ? __ldg(p) : zero;
        p += rp;
        float4 sum = zero;
#pragma unroll
        for (int k = 0; k < RING; ++k) {
            if (k == slot) ring[k] = v;
            if (k < P) sum = f4add(sum, ring[k]);
        }
        slot = (slot + 1 == P) ? 0 : slot + 1;
        finish(x, sum);
    }
}

// out layout: [q][n0][n2][n1] float4 (the tiled kernels' internal output layout)
template <int RING>
__global__ void __launch_bounds__(256)
boxmean_w_kernel(const BoxParams B, const float4* __restrict__ inter, float4* __restrict__ out) {
    const long long plane = (long long)B.n[1] * B.n[2];
    long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long e = i % plane;
    i /= plane;
    const int seg = int(i % B.nwseg);
    const int q = int(i / B.nwseg);
    if (q >= B.nv4) return;
    const int rW = B.rad[0];
    const int P = 2 * rW + 1;
    const int ys = B.out_first + seg * B.w_seg, ye = min(ys + B.w_seg, B.out_first + B.out_count);
    if (ys >= ye) return;
    // output row y needs padded rows y + pad_W - rW .. y + pad_W + rW, i.e. intermediate rows (.. - w_first)
    const float4* src = inter + (size_t(q) * B.w_count) * plane + e;
    float4* dst = out + (size_t(q) * B.n[0]) * plane + e;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 ring[RING];
    const int first = ys + B.pad[0] - rW - B.w_first;
#pragma unroll
    for (int k = 0; k < RING; ++k) ring[k] = (k < 2 * rW) ? __ldg(src + size_t(first + k) * plane) : zero;
    int slot = (2 * rW) % P;
    int y = ys;
    if (P == RING) {
        float4 s = zero;
        for (; y + RING <= ye; y += RING) {
#pragma unroll
            for (int u = 0; u < RING; ++u) {
                const int sl = (u + RING - 1) % RING;
                const float4 v = __ldg(src + size_t(y + u + B.pad[0] + rW - B.w_first) * plane);
                if (u == 0) {
                    ring[sl] = v;
                    s = zero;
#pragma unroll
                    for (int k = 0; k < RING; ++k) s = f4add(s, ring[k]);
                } else {
                    s = f4add(f4sub(s, ring[sl]), v);
                    ring[sl] = v;
                }
                dst[size_t(y + u) * plane] = make_float4(s.x * B.scale, s.y * B.scale, s.z * B.scale, s.w * B.scale);
            }
        }
        slot = (RING - 1) % RING;
    }
    for (; y < ye; ++y) {
        const float4 v = __ldg(src + size_t(y + B.pad[0] + rW - B.w_first) * plane);
        float4 s = zero;
#pragma unroll
        for (int k = 0; k < RING; ++k) {
            if (k == slot) ring[k] = v;
            if (P == RING || k < P) s = f4add(s, ring[k]);
        }
        slot = (slot + 1 == P) ? 0 : slot + 1;
        dst[size_t(y) * plane] = make_float4(s.x * B.scale, s.y * B.scale, s.z * B.scale, s.w * B.scale);
    }
}

// ---- host-side launcher -------------------------------------------------------------------------------
inline int boxmean_ring_for(int rad) {
    const int sizes[5] = {3, 7, 11, 15, 21};
    for (int k = 0; k < 5; ++k)
        if (2 * rad + 1 <= sizes[k]) return sizes[k];
    return 0;
}
// Can the fast path serve this geometry?  (ring sizes up to 21, R radius up to 4, staged pad wide enough)
inline bool boxmean_supported(const DevParams& P) {
    return boxmean_ring_for(P.rad[0]) && boxmean_ring_for(P.rad[2]) && P.rad[1] <= 10 && P.pad[1] >= P.rad[1];
}


// padded: staged cube [q][pd0][pd2][pd1] float4; out: [q][n0][n2][n1] float4; inter: (n0 + 2 rW) n1 n2 nv4 float4
template <int DUMMY = 0>
inline cudaError_t boxmean_launch_xr(const BoxParams& B, int rX, const float4* padded, float4* inter, cudaStream_t st) {
    const long long warps = (long long)B.nv4 * B.w_count * B.nwr * B.nseg;
    const unsigned grid = unsigned((warps * 32 + 127) / 128);
    switch (boxmean_ring_for(rX)) {
        case 3: boxmean_xr_kernel<3><<<grid, 128, 0, st>>>(B, padded, inter); break;
        case 7: boxmean_xr_kernel<7><<<grid, 128, 0, st>>>(B, padded, inter); break;
        case 11: boxmean_xr_kernel<11><<<grid, 128, 0, st>>>(B, padded, inter); break;
        case 15: boxmean_xr_kernel<15><<<grid, 128, 0, st>>>(B, padded, inter); break;
        default: boxmean_xr_kernel<21><<<grid, 128, 0, st>>>(B, padded, inter); break;
    }
    return cudaGetLastError();
}
template <int DUMMY = 0>
inline cudaError_t boxmean_launch_w(const BoxParams& B, int rW, const float4* inter, float4* out, cudaStream_t st) {
    const long long plane = (long long)B.n[1] * B.n[2];
    const unsigned grid = unsigned((plane * B.nwseg * B.nv4 + 255) / 256);
    switch (boxmean_ring_for(rW)) {
        case 3: boxmean_w_kernel<3><<<grid, 256, 0, st>>>(B, inter, out); break;
        case 7: boxmean_w_kernel<7><<<grid, 256, 0, st>>>(B, inter, out); break;
        case 11: boxmean_w_kernel<11><<<grid, 256, 0, st>>>(B, inter, out); break;
        case 15: boxmean_w_kernel<15><<<grid, 256, 0, st>>>(B, inter, out); break;
        default: boxmean_w_kernel<21><<<grid, 256, 0, st>>>(B, inter, out); break;
    }
    return cudaGetLastError();
}

// Rows of W that one (xr, w) kernel pair handles.  0 = the whole cube in one pair.  A slab whose intermediate
// ((slab + 2 r_W) rows) stays in the 126 MB L2 lets the W pass read it from L2 instead of DRAM.
inline int boxmean_slab_rows(const DevParams& P) {
    const char* env = getenv("NDNLM_BOXMEAN_SLAB");
    if (env) return atoi(env);
    return 0;
}
inline size_t boxmean_scratch_rows(const DevParams& P) {
    const int slab = boxmean_slab_rows(P);
    return size_t((slab > 0 && slab < P.n[0]) ? slab : P.n[0]) + 2 * size_t(P.rad[0]);
}

// padded: staged cube [q][pd0][pd2][pd1] float4; out: [q][n0][n2][n1] float4; inter: boxmean_scratch_rows x n1 n2 nv4 float4
inline cudaError_t boxmean_run(const DevParams& P, const float4* padded, float4* out, float4* inter, cudaStream_t st, int* launches) {
    BoxParams B;
    for (int k = 0; k < 3; ++k) { B.n[k] = P.n[k]; B.rad[k] = P.rad[k]; B.pad[k] = P.pad[k]; B.pd[k] = P.pd[k]; }
    B.nv4 = P.nv4;
    const double K1 = double(2 * P.rad[0] + 1) * (2 * P.rad[1] + 1) * (2 * P.rad[2] + 1);
    B.scale = float(1.0 / K1);
    B.nwr = (P.n[1] + (32 - 2 * P.rad[1]) - 1) / (32 - 2 * P.rad[1]);
    int slab = boxmean_slab_rows(P);
    if (slab <= 0 || slab > P.n[0]) slab = P.n[0];
    const long long plane = (long long)P.n[1] * P.n[2];
    for (int s0 = 0; s0 < P.n[0]; s0 += slab) {
        B.out_first = s0;
        B.out_count = (P.n[0] - s0 < slab) ? P.n[0] - s0 : slab;
        B.w_first = s0 + P.pad[0] - P.rad[0];
        B.w_count = B.out_count + 2 * P.rad[0];
        // enough warps to fill the machine (~8 K), X segments of at least 4 ring lengths
        const long long rows_warps = (long long)B.nv4 * B.w_count * B.nwr;
        int nseg = int((8192 + rows_warps - 1) / rows_warps);
        const int min_seg = 4 * (2 * P.rad[2] + 1);
        nseg = nseg < 1 ? 1 : nseg;
        B.seg_len = (P.n[2] + nseg - 1) / nseg;
        if (B.seg_len < min_seg) B.seg_len = min_seg;
        B.nseg = (P.n[2] + B.seg_len - 1) / B.seg_len;
        cudaError_t e = boxmean_launch_xr(B, P.rad[2], padded, inter, st);
        if (e != cudaSuccess) return e;
        int nwseg = int((262144 + plane * B.nv4 - 1) / (plane * B.nv4));
        nwseg = nwseg < 1 ? 1 : nwseg;
        B.w_seg = (B.out_count + nwseg - 1) / nwseg;
        const int min_wseg = 4 * (2 * P.rad[0] + 1);
        if (B.w_seg < min_wseg) B.w_seg = min_wseg;
        B.nwseg = (B.out_count + B.w_seg - 1) / B.w_seg;
        e = boxmean_launch_w(B, P.rad[0], inter, out, st);
        if (e != cudaSuccess) return e;
        *launches += 2;
    }
    return cudaSuccess;
}

}  // namespace ndnlm
