// ndflt_tiled.cuh -- fast paths of the sibling-filter kernels for C-contiguous arrays.
//
// The array is viewed as [outer][n][inner] around the slowest filtered axis (extent n): every tap moves by a whole
// number of `inner`-sized rows plus an offset inside the row.  A CTA owns a tile of TI contiguous `inner` positions x
// (RP * U) consecutive rows; a thread keeps its `inner` position and walks U rows, so
//   * global loads and stores are coalesced along `inner`,
//   * the rows a CTA touches are re-read from L1, not L2/HBM: traffic ~ (RP*U + taps)/(RP*U) of the array,
//   * the per-thread index arithmetic (32-bit divisions) is done once, not per output,
//   * weights / taps live in shared memory (one broadcast LDS per tap).
// The arithmetic (order of the taps, separate multiply and add, double accumulation) is exactly that of the generic
// kernels in ndflt.cu, i.e. scipy's NI_Correlate1D / NI_Correlate.
#pragma once
#include <type_traits>

namespace ndflt {

constexpr int TILE_THREADS = 256;
constexpr int MAX_SMEM_WEIGHTS = 1024;     // 1-D kernels up to this length use the tiled path
constexpr int MAX_SMEM_TAPS = 384;         // N-D footprints up to this many taps use the tiled path

struct TileGeom {
    long long outer, n, inner;             // contiguous view [outer][n][inner]
    int TI, RP, U;                         // tile: TI inner positions, RP rows per step, U steps per thread
    unsigned nbi, nbn;                     // tiles along inner / along n
    int mode;
    double cval;
};

// ---- 1-D correlation along the axis of extent n, inner > 1 (NI_Correlate1D) -------------------------------
template <typename T, int SYM>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_1d_tiled_kernel(const TileGeom G, const T* __restrict__ in, T* __restrict__ out,
                          const double* __restrict__ wbuf, const int size1, const int size2, const int origin) {
    __shared__ double ws[MAX_SMEM_WEIGHTS];
    const int nw = size1 + size2 + 1;
    for (int k = threadIdx.x; k < nw; k += TILE_THREADS) ws[k] = wbuf[k];
    __syncthreads();
    unsigned b = blockIdx.x;
    const unsigned bi = b % G.nbi;
    b /= G.nbi;
    const unsigned bn = b % G.nbn;
    const long long o = b / G.nbn;
    const int tr = int(threadIdx.x) / G.TI, ti = int(threadIdx.x) - tr * G.TI;
    const long long i = (long long)bi * G.TI + ti;
    if (tr >= G.RP || i >= G.inner) return;
    const long long n = G.n, inner = G.inner;
    const T* __restrict__ line = in + o * n * inner + i;            // element (o, 0, i); row l is line[l * inner]
    T* __restrict__ oline = out + o * n * inner + i;
    const double* w = ws + size1;
    for (int k = 0; k < G.U; ++k) {
        const long long row = ((long long)bn * G.U + k) * G.RP + tr;
        if (row >= n) break;
        const long long l = row - origin;                           // tap j reads row l + j
        double tmp;
        if (l - size1 >= 0 && l + size2 < n) {
            const T* p = line + l * inner;
            if (SYM != 0) {
                tmp = __dmul_rn(double(__ldg(p)), w[0]);
                for (int j = -size1; j < 0; ++j) {
                    const double a = double(__ldg(p + j * inner)), c = double(__ldg(p - j * inner));
                    tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), w[j]));
                }
            } else {
                tmp = __dmul_rn(double(__ldg(p + size2 * inner)), w[size2]);
                for (int j = -size1; j < size2; ++j) tmp = __dadd_rn(tmp, __dmul_rn(double(__ldg(p + j * inner)), w[j]));
            }
        } else {
            auto at = [&](const int j) -> double {
                const long long q = extend_index(l + j, n, G.mode);
                return q >= 0 ? double(__ldg(line + q * inner)) : G.cval;
            };
            if (SYM != 0) {
                tmp = __dmul_rn(at(0), w[0]);
                for (int j = -size1; j < 0; ++j) {
                    const double a = at(j), c = at(-j);
                    tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), w[j]));
                }
            } else {
                tmp = __dmul_rn(at(size2), w[size2]);
                for (int j = -size1; j < size2; ++j) tmp = __dadd_rn(tmp, __dmul_rn(at(j), w[j]));
            }
        }
        oline[row * inner] = T(tmp);
    }
}

// ---- 1-D symmetric / antisymmetric correlation with a register sliding window (inner > 1) -------------------
// A thread keeps its `inner` position and walks U CONSECUTIVE rows; the 2*S1+1 row values the kernel needs live in
// registers and shift by one row per output, so every input element is loaded and converted to double once per
// thread instead of once per tap (F2F runs at only 16 lanes/clk/SM on sm_100a) and the tap loop is straight-line
// code: ~10 + 3*S1 instructions per output.  Rows are warp-uniform, so the boundary extension never diverges.
// The arithmetic is NI_Correlate1D's symmetric form, farthest taps first.
template <typename T, int SYM, int S1, int U>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_1d_slide_kernel(const TileGeom G, const T* __restrict__ in, T* __restrict__ out,
                          const double* __restrict__ wbuf, const int origin) {
    constexpr int NW = 2 * S1 + 1;
    static_assert(SYM != 0, "symmetric or antisymmetric kernels only");
    double w[S1 + 1];                                               // w[d] = weight at distance d on the LEFT (w[0] centre)
#pragma unroll
    for (int d = 0; d <= S1; ++d) w[d] = __ldg(wbuf + S1 - d);
    unsigned b = blockIdx.x;
    const unsigned bi = b % G.nbi;
    b /= G.nbi;
    const unsigned bn = b % G.nbn;
    const long long o = b / G.nbn;
    const int tr = int(threadIdx.x) / G.TI, ti = int(threadIdx.x) - tr * G.TI;
    const long long i = (long long)bi * G.TI + ti;
    if (tr >= G.RP || i >= G.inner) return;
    const long long n = G.n, inner = G.inner;
    const long long row0 = ((long long)bn * G.RP + tr) * U;         // this thread's first row
    if (row0 >= n) return;
    const T* __restrict__ line = in + o * n * inner + i;
    T* __restrict__ oline = out + o * n * inner + i;
    // Raw loads run PF rows ahead of the arithmetic (memory-level parallelism: PF independent loads in flight per
    // thread).  Threads whose rows never leave [0, n) -- almost all of them -- run the INTERIOR variant: no boundary
    // extension, the row pointer advances by `inner` per output.  `cmask` remembers which prefetched values stand
    // for the constant `cval` (kept in double like scipy's line buffer).
    constexpr int PF = 16;
    static_assert(PF <= 32 && U >= PF, "prefetch depth");
    const long long base = row0 - origin - S1;                      // position held by window slot 0 before the first output
    auto run = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        auto raw = [&](const long long pos, bool& is_c) -> T {
            if constexpr (INTERIOR) {
                is_c = false;
                return __ldg(line + pos * inner);
            } else {
                const long long q = extend_index(pos, n, G.mode);
                is_c = q < 0;
                return __ldg(line + (q < 0 ? 0 : q) * inner);
            }
        };
        double win[NW];                                             // win[(pos - base) % NW] holds the value at position pos
        T pre[PF];
        unsigned cmask = 0;
#pragma unroll
        for (int c = 0; c < PF; ++c) {                              // rows entering the window at outputs 0 .. PF-1
            bool is_c;
            pre[c] = raw(base + c + NW - 1, is_c);
            cmask |= unsigned(is_c) << c;
        }
#pragma unroll
        for (int c = 0; c < NW - 1; ++c) {
            bool is_c;
            const T v = raw(base + c, is_c);
            win[c] = is_c ? G.cval : double(v);
        }
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long long row = row0 + k;
            if (INTERIOR || row < n) {
                // the newest row enters the window; its slot in the prefetch ring is refilled PF rows ahead
                win[(k + NW - 1) % NW] = (!INTERIOR && ((cmask >> (k % PF)) & 1u)) ? G.cval : double(pre[k % PF]);
                if (k + PF < U && (INTERIOR || row + PF < n)) {
                    bool is_c;
                    pre[k % PF] = raw(base + k + PF + NW - 1, is_c);
                    cmask = (cmask & ~(1u << (k % PF))) | (unsigned(is_c) << (k % PF));
                }
                double tmp = __dmul_rn(win[(k + S1) % NW], w[0]);
#pragma unroll
                for (int d = S1; d >= 1; --d) {
                    const double a = win[(k + S1 - d) % NW], c = win[(k + S1 + d) % NW];
                    tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), w[d]));
                }
                oline[row * inner] = T(tmp);
            }
        }
    };
    if (base >= 0 && base + U + NW - 2 < n && row0 + U <= n) run(std::true_type{});
    else run(std::false_type{});
}

// ---- 1-D correlation along the FASTEST axis through shared memory (inner == 1) ----------------------------------
// A CTA owns LN consecutive lines x CW consecutive positions.  The input span (CW + taps - 1 values per line,
// boundary extension applied, converted to double ONCE) is staged in shared memory with coalesced loads; every
// output then reads its taps with LDS.64 -- no boundary logic, no divergence in the arithmetic, even for lines as
// short as a warp (the time axis of a (y, x, time) cube).  S1 > 0: symmetric / antisymmetric kernel of half-width S1,
// fully unrolled, weights in registers; S1 == 0: any kernel, weights in shared memory.
struct RowsGeom {
    long long lines;       // number of lines (product of the other axes)
    unsigned n;            // line length
    int CW, LN;            // tile: CW positions x LN lines
    unsigned nbc;          // tiles along a line
    int pitch;             // doubles per staged line = CW + size1 + size2
    int mode;
    double cval;
};

template <typename T, int SYM, int S1>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_1d_rows_smem_kernel(const RowsGeom G, const T* __restrict__ in, T* __restrict__ out,
                              const double* __restrict__ wbuf, const int size1, const int size2, const int origin) {
    extern __shared__ double sm[];                                  // [LN][pitch] values, then the weights (S1 == 0)
    const unsigned bc = blockIdx.x % G.nbc;
    const long long line0 = (long long)(blockIdx.x / G.nbc) * G.LN;
    const int nlines = int(min((long long)G.LN, G.lines - line0));
    const long long c0 = (long long)bc * G.CW;                      // first output position of the tile
    const int cw = int(min((long long)G.CW, (long long)G.n - c0));
    const int span = cw + size1 + size2;
    double* wsm = sm + G.LN * G.pitch;
    if (S1 == 0)
        for (int k = threadIdx.x; k < size1 + size2 + 1; k += TILE_THREADS) wsm[k] = wbuf[k];
    // ---- stage: smem column c of a line holds position c0 - size1 - origin + c ----
    {
        int ln = int(threadIdx.x) / span, c = int(threadIdx.x) - ln * span;
        const int dln = TILE_THREADS / span, dc = TILE_THREADS - dln * span;
        const long long pos0 = c0 - size1 - origin;
        while (ln < nlines) {
            const long long q = extend_index(pos0 + c, G.n, G.mode);
            sm[ln * G.pitch + c] = q >= 0 ? double(__ldg(in + (line0 + ln) * G.n + q)) : G.cval;
            ln += dln;
            c += dc;
            if (c >= span) { c -= span; ++ln; }
        }
    }
    __syncthreads();
    // ---- compute: outputs of the tile in flat order (coalesced stores) ----
    double w[S1 + 1];
    if (S1 > 0) {
#pragma unroll
        for (int d = 0; d <= S1; ++d) w[d] = __ldg(wbuf + S1 - d);
    }
    int ln = int(threadIdx.x) / cw, l = int(threadIdx.x) - ln * cw;
    const int dln = TILE_THREADS / cw, dl = TILE_THREADS - dln * cw;
    while (ln < nlines) {
        const double* p = sm + ln * G.pitch + l + size1;            // the centre tap of this output
        double tmp;
        if (S1 > 0) {
            tmp = __dmul_rn(p[0], w[0]);
#pragma unroll
            for (int d = S1; d >= 1; --d) {
                const double a = p[-d], c = p[d];
                tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), w[d]));
            }
        } else {
            const double* wc = wsm + size1;
            if (SYM != 0) {
                tmp = __dmul_rn(p[0], wc[0]);
                for (int j = -size1; j < 0; ++j) {
                    const double a = p[j], c = p[-j];
                    tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), wc[j]));
                }
            } else {
                tmp = __dmul_rn(p[size2], wc[size2]);
                for (int j = -size1; j < size2; ++j) tmp = __dadd_rn(tmp, __dmul_rn(p[j], wc[j]));
            }
        }
        out[(line0 + ln) * G.n + c0 + l] = T(tmp);
        ln += dln;
        l += dl;
        if (l >= cw) { l -= cw; ++ln; }
    }
}

// ---- 1-D correlation along the FASTEST axis (inner == 1): lines of length n, flat thread order --------------
template <typename T, int SYM, int U>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_1d_rows_kernel(const long long total, const unsigned n, const T* __restrict__ in, T* __restrict__ out,
                         const double* __restrict__ wbuf, const int size1, const int size2, const int origin,
                         const int mode, const double cval) {
    __shared__ double ws[MAX_SMEM_WEIGHTS];
    const int nw = size1 + size2 + 1;
    for (int k = threadIdx.x; k < nw; k += TILE_THREADS) ws[k] = wbuf[k];
    __syncthreads();
    const double* w = ws + size1;
    const long long first = (long long)blockIdx.x * (TILE_THREADS * U) + threadIdx.x;
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const long long lin = first + (long long)k * TILE_THREADS;
        if (lin >= total) break;
        long long row;
        if (total <= 0xffffffffLL) row = (long long)(unsigned(lin) / n);
        else row = lin / n;
        const long long l = lin - row * n - origin;                 // tap j reads position l + j of this line
        const T* __restrict__ line = in + row * n;
        double tmp;
        if (l - size1 >= 0 && l + size2 < (long long)n) {
            const T* p = line + l;
            if (SYM != 0) {
                tmp = __dmul_rn(double(__ldg(p)), w[0]);
                for (int j = -size1; j < 0; ++j) {
                    const double a = double(__ldg(p + j)), c = double(__ldg(p - j));
                    tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), w[j]));
                }
            } else {
                tmp = __dmul_rn(double(__ldg(p + size2)), w[size2]);
                for (int j = -size1; j < size2; ++j) tmp = __dadd_rn(tmp, __dmul_rn(double(__ldg(p + j)), w[j]));
            }
        } else {
            auto at = [&](const int j) -> double {
                const long long q = extend_index(l + j, n, mode);
                return q >= 0 ? double(__ldg(line + q)) : cval;
            };
            if (SYM != 0) {
                tmp = __dmul_rn(at(0), w[0]);
                for (int j = -size1; j < 0; ++j) {
                    const double a = at(j), c = at(-j);
                    tmp = __dadd_rn(tmp, __dmul_rn(SYM > 0 ? __dadd_rn(a, c) : __dsub_rn(a, c), w[j]));
                }
            } else {
                tmp = __dmul_rn(at(size2), w[size2]);
                for (int j = -size1; j < size2; ++j) tmp = __dadd_rn(tmp, __dmul_rn(at(j), w[j]));
            }
        }
        out[lin] = T(tmp);
    }
}

// ---- dense KH x KW footprints with a register sliding window (NI_Correlate) ------------------------------------
// The common ConvolutionFilter / BoxcarFilter case: a 2-D kernel whose rows run along the tiled axis (extent n) and
// whose columns run along ONE trailing axis (extent dimx, element stride xs inside the row).  A thread keeps its
// position inside the row and walks U consecutive rows; the KH x KW input values live in registers (each loaded and
// converted once per thread), the weights in shared memory.  Taps are accumulated in C order from 0.0 with a separate
// multiply and add, as in scipy.
struct Dense2dGeom {
    unsigned dimx;         // extent of the trailing axis the kernel columns run along
    unsigned xs;           // its element stride inside the row (product of the faster extents)
    int oy, ox;            // scipy origins along the two axes
};

template <typename T, int KH, int KW, int U>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_2d_slide_kernel(const TileGeom G, const Dense2dGeom D, const T* __restrict__ in, T* __restrict__ out,
                          const double* __restrict__ wbuf) {
    __shared__ double ws[KH * KW];
    for (int k = threadIdx.x; k < KH * KW; k += TILE_THREADS) ws[k] = wbuf[k];
    __syncthreads();
    unsigned b = blockIdx.x;
    const unsigned bi = b % G.nbi;
    b /= G.nbi;
    const unsigned bn = b % G.nbn;
    const long long o = b / G.nbn;
    const int tr = int(threadIdx.x) / G.TI, ti = int(threadIdx.x) - tr * G.TI;
    const long long i = (long long)bi * G.TI + ti;
    if (tr >= G.RP || i >= G.inner) return;
    const long long n = G.n, inner = G.inner;
    const long long row0 = ((long long)bn * G.RP + tr) * U;
    if (row0 >= n) return;
    const int ix = int((unsigned(i) / D.xs) % D.dimx);              // my index along the column axis
    const long long irest = i - (long long)ix * D.xs;               // my position inside the row without that axis
    const T* __restrict__ plane = in + o * n * inner;
    T* __restrict__ oline = out + o * n * inner + i;
    const int dx0 = -(KW / 2) - D.ox;                               // column c reads ix + dx0 + c
    const long long base = row0 - (KH / 2) - D.oy;                  // window row slot 0 holds position `base` before output 0
    auto run = [&](auto interior_tag) {
        constexpr bool INTERIOR = decltype(interior_tag)::value;
        long long xoff[KW];                                         // element offsets of my KW columns (or -1: constant)
#pragma unroll
        for (int c = 0; c < KW; ++c) {
            if constexpr (INTERIOR) {
                xoff[c] = irest + (long long)(ix + dx0 + c) * D.xs;
            } else {
                const long long q = extend_index(ix + dx0 + c, D.dimx, G.mode);
                xoff[c] = q < 0 ? -1 : irest + q * D.xs;
            }
        }
        auto load_row = [&](const long long pos, double (&dst)[KW]) {
            long long q = pos;
            if constexpr (!INTERIOR) q = extend_index(pos, n, G.mode);
            const T* r = plane + (q < 0 ? 0 : q) * inner;
#pragma unroll
            for (int c = 0; c < KW; ++c) {
                if constexpr (INTERIOR) dst[c] = double(__ldg(r + xoff[c]));
                else dst[c] = (q < 0 || xoff[c] < 0) ? G.cval : double(__ldg(r + xoff[c]));
            }
        };
        double win[KH][KW];
#pragma unroll
        for (int r = 0; r < KH - 1; ++r) load_row(base + r, win[r]);
#pragma unroll
        for (int k = 0; k < U; ++k) {
            const long long row = row0 + k;
            if (INTERIOR || row < n) {
                load_row(base + k + KH - 1, win[(k + KH - 1) % KH]);
                double tmp = 0.0;
#pragma unroll
                for (int r = 0; r < KH; ++r)
#pragma unroll
                    for (int c = 0; c < KW; ++c) tmp = __dadd_rn(tmp, __dmul_rn(win[(k + r) % KH][c], ws[r * KW + c]));
                oline[row * inner] = T(tmp);
            }
        }
    };
    const bool interior = base >= 0 && base + U + KH - 2 < n && row0 + U <= n && ix + dx0 >= 0 &&
                          ix + dx0 + KW - 1 < int(D.dimx);
    if (interior) run(std::true_type{});
    else run(std::false_type{});
}

// ---- dense KH x KW x KT footprints (e.g. BoxcarFilter(dims=('y','x','time'), w=3)) --------------------------------
// Like correlate_2d_slide_kernel with a second trailing axis.  The boundary extension along the two trailing axes is
// turned into DATA once per thread (element offsets of the KW x KT columns, -1 = constant), so warps whose lanes sit
// on the edge of a short fastest axis (a 32-long time axis puts edge lanes into every warp) do not diverge; the
// extension along the walked axis is warp-uniform.
struct Dense3dGeom {
    unsigned dimx, xs;     // first trailing kernel axis: extent, element stride inside the row
    unsigned dimt, ts;     // second trailing kernel axis (faster than the first)
    int oy, ox, ot;        // scipy origins
};

template <typename T, int KH, int KW, int KT, int U>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_3d_slide_kernel(const TileGeom G, const Dense3dGeom D, const T* __restrict__ in, T* __restrict__ out,
                          const double* __restrict__ wbuf) {
    constexpr int NC = KW * KT;
    __shared__ double ws[KH * NC];
    for (int k = threadIdx.x; k < KH * NC; k += TILE_THREADS) ws[k] = wbuf[k];
    __syncthreads();
    unsigned b = blockIdx.x;
    const unsigned bi = b % G.nbi;
    b /= G.nbi;
    const unsigned bn = b % G.nbn;
    const long long o = b / G.nbn;
    const int tr = int(threadIdx.x) / G.TI, ti = int(threadIdx.x) - tr * G.TI;
    const long long i = (long long)bi * G.TI + ti;
    if (tr >= G.RP || i >= G.inner) return;
    const long long n = G.n, inner = G.inner;
    const long long row0 = ((long long)bn * G.RP + tr) * U;
    if (row0 >= n) return;
    const int ix = int((unsigned(i) / D.xs) % D.dimx), it = int((unsigned(i) / D.ts) % D.dimt);
    const long long irest = i - (long long)ix * D.xs - (long long)it * D.ts;
    const T* __restrict__ plane = in + o * n * inner;
    T* __restrict__ oline = out + o * n * inner + i;
    long long coff[NC];                                             // element offsets of my KW x KT columns, -1: constant
#pragma unroll
    for (int c = 0; c < KW; ++c) {
        const long long qx = extend_index(ix - (KW / 2) - D.ox + c, D.dimx, G.mode);
#pragma unroll
        for (int d = 0; d < KT; ++d) {
            const long long qt = extend_index(it - (KT / 2) - D.ot + d, D.dimt, G.mode);
            coff[c * KT + d] = (qx < 0 || qt < 0) ? -1 : irest + qx * D.xs + qt * D.ts;
        }
    }
    auto load_row = [&](const long long pos, double (&dst)[NC]) {
        const long long q = extend_index(pos, n, G.mode);           // warp-uniform
        const T* r = plane + (q < 0 ? 0 : q) * inner;
#pragma unroll
        for (int c = 0; c < NC; ++c) {
            const double v = double(__ldg(r + (coff[c] < 0 ? 0 : coff[c])));
            dst[c] = (q < 0 || coff[c] < 0) ? G.cval : v;
        }
    };
    const long long base = row0 - (KH / 2) - D.oy;
    double win[KH][NC];
#pragma unroll
    for (int r = 0; r < KH - 1; ++r) load_row(base + r, win[r]);
#pragma unroll
    for (int k = 0; k < U; ++k) {
        const long long row = row0 + k;
        if (row < n) {
            load_row(base + k + KH - 1, win[(k + KH - 1) % KH]);
            double tmp = 0.0;
#pragma unroll
            for (int r = 0; r < KH; ++r)
#pragma unroll
                for (int c = 0; c < NC; ++c) tmp = __dadd_rn(tmp, __dmul_rn(win[(k + r) % KH][c], ws[r * NC + c]));
            oline[row * inner] = T(tmp);
        }
    }
}

// ---- N-D correlation (NI_Correlate), tiled around the slowest filtered axis ---------------------------------
struct TiledTap {
    long long lin;         // element offset of the tap: ds * inner + offset inside the row
    double w;
    int ds;                // offset along the tiled axis
    int dt[3];             // offsets along the (up to three) trailing axes
};

struct TrailGeom {
    int ntrail;            // number of trailing axes (0..3)
    unsigned dim[3];       // their extents, slowest first
    long long str[3];      // their element strides
    int lo[3], hi[3];      // smallest / largest tap offset per trailing axis
    int slo, shi;          // the same along the tiled axis
};

template <typename T>
__global__ void __launch_bounds__(TILE_THREADS)
correlate_nd_tiled_kernel(const TileGeom G, const TrailGeom R, const T* __restrict__ in, T* __restrict__ out,
                          const TiledTap* __restrict__ gtaps, const int ntaps) {
    __shared__ TiledTap taps[MAX_SMEM_TAPS];
    for (int k = threadIdx.x; k < ntaps; k += TILE_THREADS) taps[k] = gtaps[k];
    __syncthreads();
    unsigned b = blockIdx.x;
    const unsigned bi = b % G.nbi;
    b /= G.nbi;
    const unsigned bn = b % G.nbn;
    const long long o = b / G.nbn;
    const int tr = int(threadIdx.x) / G.TI, ti = int(threadIdx.x) - tr * G.TI;
    const long long i = (long long)bi * G.TI + ti;
    if (tr >= G.RP || i >= G.inner) return;
    // position inside the row along the trailing axes (once per thread)
    unsigned idx[3] = {0, 0, 0};
    {
        unsigned rem = unsigned(i);
        for (int a = R.ntrail - 1; a >= 0; --a) {
            idx[a] = rem % R.dim[a];
            rem /= R.dim[a];
        }
    }
    bool trail_interior = true;
    for (int a = 0; a < R.ntrail; ++a)
        trail_interior = trail_interior && (int(idx[a]) + R.lo[a] >= 0) && (int(idx[a]) + R.hi[a] < int(R.dim[a]));
    const long long n = G.n, inner = G.inner;
    const T* __restrict__ plane = in + o * n * inner;               // element (o, 0, 0)
    T* __restrict__ oline = out + o * n * inner + i;
    for (int k = 0; k < G.U; ++k) {
        const long long row = ((long long)bn * G.U + k) * G.RP + tr;
        if (row >= n) break;
        double tmp = 0.0;
        if (trail_interior && row + R.slo >= 0 && row + R.shi < n) {
            const T* p = plane + row * inner + i;
#pragma unroll 4
            for (int t = 0; t < ntaps; ++t) tmp = __dadd_rn(tmp, __dmul_rn(double(__ldg(p + taps[t].lin)), taps[t].w));
        } else {
            for (int t = 0; t < ntaps; ++t) {
                long long q = extend_index(row + taps[t].ds, n, G.mode);
                bool inside = q >= 0;
                long long src = q * inner;
                for (int a = 0; a < R.ntrail; ++a) {
                    q = extend_index((long long)idx[a] + taps[t].dt[a], R.dim[a], G.mode);
                    inside = inside && q >= 0;
                    src += q * R.str[a];
                }
                const double v = inside ? double(__ldg(plane + src)) : G.cval;
                tmp = __dadd_rn(tmp, __dmul_rn(v, taps[t].w));
            }
        }
        oline[row * inner] = T(tmp);
    }
}

}  // namespace ndflt
