// Host emulation of stage_tiled_rows_kernel (nlm_staging.cuh): every (block, thread) of its grid is run on the CPU and
// the staged cube is compared element by element with the rule of the original one-thread-per-element kernel
// (reflection / EDGE_HALO / EDGE_SOURCE on the shard axis, strided variable-major input, zero-padded variables).
// Build and run (no GPU needed):  nvcc -std=c++17 -gencode arch=compute_100a,code=sm_100a -o /tmp/emu_stage tools/emu_stage.cu && /tmp/emu_stage
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include "../nd_b200/csrc/nlm_staging.cuh"
using namespace ndnlm;

template <typename TIN, typename V4>
static void legacy(const StageParams& S, const TIN* arr, V4* padded, std::vector<char>& written) {
    using TS = decltype(V4().x);
    const long long plane = (long long)S.pd[0] * S.pd[1] * S.pd[2];
    for (long long i = 0; i < plane * S.nv4; ++i) {
        const int q = int(i / plane);
        long long rem = i - q * plane;
        int ip[3];
        ip[1] = int(rem % S.pd[1]); rem /= S.pd[1];
        ip[2] = int(rem % S.pd[2]); ip[0] = int(rem / S.pd[2]);
        long long src = 0;
        bool skip = false;
        for (int role = 0; role < 3 && !skip; ++role) {
            const int u = ip[role] - S.pad[role];
            if (role == S.halo_role) {
                if (S.lo_halo == 1 && u < 0) { skip = true; break; }
                if (S.hi_halo == 1 && u >= S.n[role]) { skip = true; break; }
                if ((S.lo_halo == 2 && u < 0) || (S.hi_halo == 2 && u >= S.n[role])) { src += (long long)u * S.rstride[role]; continue; }
            }
            src += (long long)reflect_index(u, S.n[role]) * S.rstride[role];
        }
        if (skip) continue;
        TS v[4];
        for (int k = 0; k < 4; ++k) { const int var = 4 * q + k; v[k] = (var < S.V) ? TS(arr[src + var * S.vstride]) : TS(0); }
        padded[i].x = v[0]; padded[i].y = v[1]; padded[i].z = v[2]; padded[i].w = v[3];
        written[i] = 1;
    }
}

template <typename TIN, typename V4, bool VEC>
static long long run_case(const int n[3], const int pad[3], int V, const int order[4], int halo_role, int lo, int hi, int margin) {
    // user array: dims (a0, a1, a2, V) in memory order `order` (slowest first); role k <-> axis k here; the shard axis
    // carries `margin` extra real rows on both sides (EDGE_SOURCE reads them)
    StageParams S;
    int ext[4] = {n[0], n[1], n[2], V};
    if (halo_role >= 0) ext[halo_role] += 2 * margin;
    long long stride[4], acc = 1;
    for (int k = 3; k >= 0; --k) { stride[order[k]] = acc; acc *= ext[order[k]]; }
    std::vector<TIN> arr_store(acc + 8);
    // 32-byte aligned base
    TIN* base = arr_store.data();
    while (reinterpret_cast<uintptr_t>(base) % 32) ++base;
    for (long long i = 0; i < acc; ++i) base[i] = TIN(i % 9973) * TIN(0.25) + TIN(1);
    const TIN* arr = base + (halo_role >= 0 ? margin * stride[halo_role] : 0);
    for (int k = 0; k < 3; ++k) { S.n[k] = n[k]; S.pad[k] = pad[k]; S.pd[k] = n[k] + 2 * pad[k]; S.rstride[k] = stride[k]; }
    S.vstride = stride[3]; S.V = V; S.nv4 = (V + 3) / 4; S.halo_role = halo_role; S.lo_halo = lo; S.hi_halo = hi;
    const long long total = (long long)S.pd[0] * S.pd[1] * S.pd[2] * S.nv4;
    std::vector<V4> a(total), b(total);
    std::vector<char> wa(total, 0), wb(total, 0);
    memset(a.data(), 0xCD, total * sizeof(V4)); memset(b.data(), 0xCD, total * sizeof(V4));
    legacy<TIN, V4>(S, arr, a.data(), wa);
    const unsigned wblocks = unsigned((S.pd[0] + STAGE_ROWS - 1) / STAGE_ROWS);
    const unsigned gx = unsigned(((long long)S.pd[1] * S.pd[2] + 255) / 256), gy = wblocks < 3 ? wblocks : 3;   // small grid.y: exercises the loop
    for (int q = 0; q < S.nv4; ++q)
        for (unsigned by = 0; by < gy; ++by)
            for (unsigned bx = 0; bx < gx; ++bx)
                for (unsigned t = 0; t < 256; ++t)
                    for (unsigned wb_ = by; wb_ < wblocks; wb_ += gy)
                        stage_tiled_rows_thread<TIN, V4, VEC>(S, arr, b.data(), bx * 256u + t, wb_, q);
    long long bad = 0;
    for (long long i = 0; i < total; ++i) bad += memcmp(&a[i], &b[i], sizeof(V4)) != 0;
    return bad;
}

int main() {
    long long bad = 0; int cases = 0;
    const int orders[3][4] = {{0, 1, 2, 3}, {3, 0, 1, 2}, {0, 2, 1, 3}};
    const int shapes[4][3] = {{9, 7, 13}, {1, 20, 17}, {12, 5, 1}, {6, 6, 6}};
    const int pads[4][3] = {{3, 2, 4}, {0, 5, 3}, {4, 2, 0}, {5, 5, 5}};
    for (int s = 0; s < 4; ++s)
        for (int o = 0; o < 3; ++o)
            for (int V = 1; V <= 8; ++V)
                for (int hr = -1; hr < 3; ++hr)
                    for (int mode = 0; mode < (hr < 0 ? 1 : 9); ++mode) {     // every (lo, hi) pair of reflect / halo / source
                        const int lo = mode / 3, hi = mode % 3;
                        const int margin = hr >= 0 ? pads[s][hr] : 0;
                        bad += run_case<float, float4, false>(shapes[s], pads[s], V, orders[o], hr, lo, hi, margin);
                        bad += run_case<double, double4v, false>(shapes[s], pads[s], V, orders[o], hr, lo, hi, margin);
                        bad += run_case<double, float4, false>(shapes[s], pads[s], V, orders[o], hr, lo, hi, margin);
                        cases += 3;
                        if (o != 1 && V % 4 == 0) {     // variables contiguous: the vector-load variants
                            bad += run_case<float, float4, true>(shapes[s], pads[s], V, orders[o], hr, lo, hi, margin);
                            bad += run_case<double, double4v, true>(shapes[s], pads[s], V, orders[o], hr, lo, hi, margin);
                            cases += 2;
                        }
                    }
    printf("emu_stage: %d cases, %lld mismatching elements\n", cases, bad);
    return bad != 0;
}
