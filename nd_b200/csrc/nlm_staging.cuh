// nlm_staging.cuh -- HBM-bound data movement around the filter kernels.
//
//   stage    strided caller array (N0,N1,N2,V) -> internal reflect-padded cube in role order.
//            Materialises np.pad(mode='reflect') by r+f, i.e. the reference's `_idx`
//            (nd/_filters.pyx:34-40) applied to p+d / q+d (:378-384), once instead of per access.
//   unstage  internal output -> strided caller array (the reference's in-place write to `output`).
//   halo     pack / unpack of the pad rows along one axis (y-shard neighbour exchange).
//   synth    counter-based synthetic SAR-like cube keyed by GLOBAL voxel index.
//
// Two internal layouts:
//   tiled    [nv4][W][X][R] float4  (R fastest; variables padded with zeros to a multiple of 4)
//   generic  [W][R][X][V]   T
#pragma once
#include "nlm_common.cuh"
#include "nlm_tiled.cuh"   // double4v, mk4

namespace ndnlm {

struct StageParams {
    int n[3], pad[3], pd[3];       // role order
    long long rstride[3];          // user element stride of the axis playing each role
    long long vstride;             // user element stride of the variable axis
    int V, nv4;
    int halo_role;                 // role of the shard axis or -1
    int lo_halo, hi_halo;          // edge mode of the shard axis: 0 reflect, 1 leave those pad rows untouched,
                                   // 2 the source array itself extends over them (slab of a larger array)
};

template <typename TIN, typename V4>
__global__ void stage_tiled_kernel(const StageParams S, const TIN* __restrict__ arr, V4* __restrict__ padded) {
    using TS = decltype(V4().x);     // float or double: the type the kernel computes in
    const long long plane = (long long)S.pd[0] * S.pd[1] * S.pd[2];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane * S.nv4) return;
    const int q = int(i / plane);
    long long rem = i - q * plane;
    int ip[3];                       // tiled layout: R fastest, then X, then W
    ip[1] = int(rem % S.pd[1]);
    rem /= S.pd[1];
    ip[2] = int(rem % S.pd[2]);
    ip[0] = int(rem / S.pd[2]);
    long long src = 0;
#pragma unroll
    for (int role = 0; role < 3; ++role) {
        const int u = ip[role] - S.pad[role];
        if (role == S.halo_role) {
            if (S.lo_halo == 1 && u < 0) return;
            if (S.hi_halo == 1 && u >= S.n[role]) return;
            if ((S.lo_halo == 2 && u < 0) || (S.hi_halo == 2 && u >= S.n[role])) {
                src += (long long)u * S.rstride[role];      // real rows of the enclosing array
                continue;
            }
        }
        src += (long long)reflect_index(u, S.n[role]) * S.rstride[role];
    }
    TS v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int var = 4 * q + k;
        v[k] = (var < S.V) ? TS(arr[src + var * S.vstride]) : TS(0);
    }
    padded[i] = mk4(v[0], v[1], v[2], v[3]);
}

// Source offset contribution of one role for padded index `ip` (the rule of stage_tiled_kernel above): reflection
// (reference `_idx`), or -- on the shard axis -- rows that are left untouched (false is returned) / read from the
// enclosing array.
__host__ __device__ inline bool stage_role_offset(const StageParams& S, const int role, const int ip, long long& off) {
    const int u = ip - S.pad[role];
    if (role == S.halo_role) {
        if ((S.lo_halo == 1 && u < 0) || (S.hi_halo == 1 && u >= S.n[role])) return false;
        if ((S.lo_halo == 2 && u < 0) || (S.hi_halo == 2 && u >= S.n[role])) {
            off += (long long)u * S.rstride[role];
            return true;
        }
    }
    off += (long long)reflect_index(u, S.n[role]) * S.rstride[role];
    return true;
}

// One thread of stage_tiled_rows_kernel (also run on the host by tools/emu_stage.cu, which checks it against
// stage_tiled_kernel's rule element by element): thread `j` of the (X, R) plane stages STAGE_ROWS consecutive W rows.
constexpr int STAGE_ROWS = 4;
template <typename TIN, typename V4, bool VEC>
__host__ __device__ inline void stage_tiled_rows_thread(const StageParams& S, const TIN* __restrict__ arr,
                                                        V4* __restrict__ padded, const unsigned j, const unsigned wblock,
                                                        const int q) {
    using TS = decltype(V4().x);
    const unsigned pdr = unsigned(S.pd[1]), pdx = unsigned(S.pd[2]);
    if (j >= pdr * pdx) return;
    const unsigned x = j / pdr, r = j - x * pdr;            // R is the fastest axis of the staged cube
    long long off_xr = (long long)(4 * q) * S.vstride;
    if (!stage_role_offset(S, ROLE_X, int(x), off_xr)) return;
    if (!stage_role_offset(S, ROLE_R, int(r), off_xr)) return;
    const int w0 = int(wblock) * STAGE_ROWS;
    V4 v[STAGE_ROWS];
    bool ok[STAGE_ROWS];
#pragma unroll
    for (int k = 0; k < STAGE_ROWS; ++k) {
        long long src = off_xr;
        ok[k] = (w0 + k < S.pd[0]) && stage_role_offset(S, ROLE_W, w0 + k, src);
        if (!ok[k]) continue;
        if constexpr (VEC) {
            v[k] = *reinterpret_cast<const V4*>(arr + src);   // 4 consecutive variables, 16 / 32-byte aligned (host check)
        } else {
            TS t[4];
#pragma unroll
            for (int c = 0; c < 4; ++c) t[c] = (4 * q + c < S.V) ? TS(arr[src + c * S.vstride]) : TS(0);
            v[k].x = t[0]; v[k].y = t[1]; v[k].z = t[2]; v[k].w = t[3];
        }
    }
#pragma unroll
    for (int k = 0; k < STAGE_ROWS; ++k)
        if (ok[k]) padded[((size_t(q) * S.pd[0] + (w0 + k)) * pdx + x) * pdr + r] = v[k];
}

// The staging kernel of the tiled layout: the same result as stage_tiled_kernel, with 32-bit index arithmetic done once
// per STAGE_ROWS elements and one vector load per voxel where the caller's array allows it (VEC: the variables are
// contiguous, a multiple of 4, and every voxel starts on a 16- / 32-byte boundary).  grid = ((X R plane) / 256,
// ceil(pd_W / STAGE_ROWS) [folded into y and z], nv4 [z]).
template <typename TIN, typename V4, bool VEC>
__global__ void __launch_bounds__(256) stage_tiled_rows_kernel(const StageParams S, const TIN* __restrict__ arr,
                                                               V4* __restrict__ padded, const unsigned wblocks) {
    const unsigned j = blockIdx.x * 256u + threadIdx.x;
    const int q = int(blockIdx.z);
    for (unsigned wb = blockIdx.y; wb < wblocks; wb += gridDim.y) stage_tiled_rows_thread<TIN, V4, VEC>(S, arr, padded, j, wb, q);
}

template <typename T>
__global__ void stage_generic_kernel(const StageParams S, const T* __restrict__ arr, T* __restrict__ padded) {
    const long long total = (long long)S.pd[0] * S.pd[1] * S.pd[2] * S.V;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int var = int(i % S.V);
    long long rem = i / S.V;
    int ip[3];
    ip[2] = int(rem % S.pd[2]);
    rem /= S.pd[2];
    ip[1] = int(rem % S.pd[1]);
    ip[0] = int(rem / S.pd[1]);
    long long src = (long long)var * S.vstride;
#pragma unroll
    for (int role = 0; role < 3; ++role) {
        const int u = ip[role] - S.pad[role];
        if (role == S.halo_role) {
            if (S.lo_halo == 1 && u < 0) return;
            if (S.hi_halo == 1 && u >= S.n[role]) return;
            if ((S.lo_halo == 2 && u < 0) || (S.hi_halo == 2 && u >= S.n[role])) {
                src += (long long)u * S.rstride[role];      // real rows of the enclosing array
                continue;
            }
        }
        src += (long long)reflect_index(u, S.n[role]) * S.rstride[role];
    }
    padded[i] = arr[src];
}

template <typename TOUT, typename V4>
__global__ void unstage_tiled_kernel(const StageParams S, const V4* __restrict__ internal, TOUT* __restrict__ output) {
    using TS = decltype(V4().x);
    const long long plane = (long long)S.n[0] * S.n[1] * S.n[2];
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= plane * S.nv4) return;
    const int q = int(i / plane);
    long long rem = i - q * plane;
    int ip[3];                       // tiled layout: R fastest, then X, then W
    ip[1] = int(rem % S.n[1]);
    rem /= S.n[1];
    ip[2] = int(rem % S.n[2]);
    ip[0] = int(rem / S.n[2]);
    long long dst = 0;
#pragma unroll
    for (int role = 0; role < 3; ++role) dst += (long long)ip[role] * S.rstride[role];
    const V4 v = internal[i];
    const TS vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const int var = 4 * q + k;
        if (var < S.V) output[dst + var * S.vstride] = TOUT(vv[k]);
    }
}

template <typename T>
__global__ void unstage_generic_kernel(const StageParams S, const T* __restrict__ internal, T* __restrict__ output) {
    const long long total = (long long)S.n[0] * S.n[1] * S.n[2] * S.V;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int var = int(i % S.V);
    long long rem = i / S.V;
    int ip[3];
    ip[2] = int(rem % S.n[2]);
    rem /= S.n[2];
    ip[1] = int(rem % S.n[1]);
    ip[0] = int(rem / S.n[1]);
    long long dst = (long long)var * S.vstride;
#pragma unroll
    for (int role = 0; role < 3; ++role) dst += (long long)ip[role] * S.rstride[role];
    output[dst] = internal[i];
}

// Copy `rows` consecutive rows (starting at row `first`) along role `hr` between the padded cube and a
// dense message buffer.  `elems` = units per voxel row element: the cube is viewed as
// [outer][pd_hr][inner] units of UNIT bytes.
template <typename UNIT, bool PACK>
__global__ void halo_copy_kernel(UNIT* __restrict__ cube, UNIT* __restrict__ msg, long long outer, long long pd_hr,
                                 long long inner, long long first, long long rows) {
    const long long total = outer * rows * inner;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long in = i % inner;
    const long long row = (i / inner) % rows;
    const long long o = i / (inner * rows);
    const long long ci = (o * pd_hr + first + row) * inner + in;
    if (PACK)
        msg[i] = cube[ci];
    else
        cube[ci] = msg[i];
}

// ---- synthetic SAR-like cube -------------------------------------------------------------
__host__ __device__ inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ inline float u01(uint64_t bits) {   // (0, 1]
    return (float((bits >> 40) & 0xFFFFFF) + 1.0f) * (1.0f / 16777216.0f);
}

// Multi-look (L=4) covariance-like values: intensities ~ scene * Gamma(4, 1/4), cross terms
// ~ N(0, 0.3^2) * scene; the scene is piecewise constant (64x64 blocks, 3 levels).
__global__ void synth_cube_kernel(float* __restrict__ out, long long ny, long long nx, long long nt, int V,
                                  long long y_offset, uint64_t seed) {
    const long long total = ny * nx * nt;
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const long long t = i % nt;
    const long long x = (i / nt) % nx;
    const long long y = i / (nt * nx) + y_offset;
    const float scene = 0.5f + 0.75f * float(((y >> 6) + 2 * (x >> 6)) % 3);
    const uint64_t key = splitmix64(seed ^ splitmix64(uint64_t(y) * 0x100000001B3ull + uint64_t(x))) + uint64_t(t) * 0x9E3779B97F4A7C15ull;
    float* o = out + i * V;
    for (int v = 0; v < V; ++v) {
        const uint64_t k = splitmix64(key + 0x632BE59BD9B4E019ull * uint64_t(v + 1));
        const bool intensity = (v == 0) || (v == 3) || (v == 4);
        float val;
        if (intensity) {
            float g = 0.f;
            uint64_t kk = k;
#pragma unroll
            for (int l = 0; l < 4; ++l) {
                kk = splitmix64(kk);
                g -= __logf(u01(kk));
            }
            val = scene * 0.25f * g;
        } else {
            const uint64_t k2 = splitmix64(k);
            const float rad = sqrtf(-2.0f * __logf(u01(k)));
            val = 0.3f * scene * rad * __cosf(6.2831853f * u01(k2));
        }
        o[v] = val;
    }
}

}  // namespace ndnlm
