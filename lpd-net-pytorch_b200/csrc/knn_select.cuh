// Thread-per-row top-L selection shared by the kNN kernels.
//
// One thread owns one query row.  Its L best candidates live UNSORTED in shared memory, laid out [slot][row] with a row
// stride of SEL_STRIDE = 129 words, so that "all lanes touch the same slot of their own row" (the selection loop) and
// "all lanes touch different slots of one row" (the final cooperative sort / write-out) are both bank-conflict free.
// The list is split into 4 groups of GS = L/4 slots; the thread keeps the minimum (worst) entry of every group and its
// position in registers.  Replacing the globally worst entry then costs one store plus a rescan of ONE group (GS loads)
// instead of a shuffle-serialised sorted insert - no cross-lane dependency, every lane works on its own row in parallel.
//
// Order relation: entry (v, j) is better than (v', j') when v > v', or v == v' and j < j'  (score descending, index
// ascending - the canonical kNN order with v = pd).  With EXACT == false ties on v are not ordered by index (used by the
// tensor-core filter, whose scores are approximate anyway and which keeps L > k candidates).
#pragma once
#include "common.cuh"
#include <limits.h>

namespace lpd {

constexpr int SEL_STRIDE = 129;   // default: words between consecutive slots (128 rows + 1 pad)

template <int GS, bool EXACT, int STRIDE = SEL_STRIDE>
struct RowSelect {
    static constexpr int L = 4 * GS;
    float gmin[4];     // worst score of each group
    int gidx[4];       // its candidate index (EXACT tie-break)
    int gpos[4];       // its slot inside the group
    float tau;         // worst score of the whole list (-inf while the list is not full)
    int tau_idx;       // candidate index of that entry (EXACT only)
    int filled;        // number of valid entries (<= L)

    __device__ __forceinline__ void reset() {
        filled = 0;
        tau = -INFINITY;
        tau_idx = INT_MAX;
#pragma unroll
        for (int g = 0; g < 4; ++g) { gmin[g] = -INFINITY; gidx[g] = INT_MAX; gpos[g] = 0; }
    }

    // is (v, j) worse than (w, i)?
    __device__ __forceinline__ static bool worse(float v, int j, float w, int i) {
        return EXACT ? (v < w || (v == w && j > i)) : (v < w);
    }

    // would (v, j) enter the list?  (true for everything while the list is still filling, except -inf scores)
    __device__ __forceinline__ bool passes(float v, int j) const {
        return EXACT ? (v > tau || (v == tau && j < tau_idx)) : (v > tau);
    }

    __device__ __forceinline__ void rescan_group(const float* __restrict__ lv, const int* __restrict__ li, int row, int g) {
        float m = INFINITY;
        int mi = -1, mp = 0;
#pragma unroll
        for (int s = 0; s < GS; ++s) {
            const float x = lv[(g * GS + s) * STRIDE + row];
            const int xi = EXACT ? li[(g * GS + s) * STRIDE + row] : 0;
            if (s == 0 || worse(x, xi, m, mi)) { m = x; mi = xi; mp = s; }
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
            if (q == g) { gmin[q] = m; gidx[q] = mi; gpos[q] = mp; }
    }

    __device__ __forceinline__ void refresh_tau() {
        float m = gmin[0];
        int mi = gidx[0];
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (worse(gmin[q], gidx[q], m, mi)) { m = gmin[q]; mi = gidx[q]; }
        tau = m;
        tau_idx = mi;
    }

    // insert candidate (v, j) of row `row` (caller guarantees v > -inf)
    __device__ __forceinline__ void insert(float* __restrict__ lv, int* __restrict__ li, int row, float v, int j) {
        if (filled < L) {
            lv[filled * STRIDE + row] = v;
            li[filled * STRIDE + row] = j;
            ++filled;
            if (filled == L) {
#pragma unroll
                for (int g = 0; g < 4; ++g) rescan_group(lv, li, row, g);
                refresh_tau();
            }
            return;
        }
        if (!passes(v, j)) return;
        // the globally worst entry sits in the group whose minimum equals (tau, tau_idx)
        int g = 0;
        {
            float m = gmin[0];
            int mi = gidx[0];
#pragma unroll
            for (int q = 1; q < 4; ++q)
                if (worse(gmin[q], gidx[q], m, mi)) { m = gmin[q]; mi = gidx[q]; g = q; }
        }
        int p = gpos[0];
#pragma unroll
        for (int q = 1; q < 4; ++q)
            if (q == g) p = gpos[q];
        const int slot = g * GS + p;
        lv[slot * STRIDE + row] = v;
        li[slot * STRIDE + row] = j;
        rescan_group(lv, li, row, g);
        refresh_tau();
    }
};

// bitonic sort of 32*E entries spread over a warp (entry e*32 + lane) by (value descending, index ascending)
template <int E>
__device__ __forceinline__ void warp_sort_desc(float (&v)[E], int (&id)[E], int lane) {
    auto before = [](float a, int ia, float b, int ib) { return (a > b) || (a == b && ia < ib); };
    constexpr int TOTAL = 32 * E;
    for (int size = 2; size <= TOTAL; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {   // E == 2 and size == 64: partner is the other register of the same lane
                const bool swap = before(v[E - 1], id[E - 1], v[0], id[0]);
                if (swap) { float t = v[0]; v[0] = v[E - 1]; v[E - 1] = t; int ti = id[0]; id[0] = id[E - 1]; id[E - 1] = ti; }
            } else {
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int g = e * 32 + lane;
                    const float pv = __shfl_xor_sync(kFull, v[e], stride);
                    const int pi = __shfl_xor_sync(kFull, id[e], stride);
                    const bool dir_desc = ((g & size) == 0) || (size == TOTAL);
                    const bool lower = ((g & stride) == 0);
                    const bool mine_better = before(v[e], id[e], pv, pi);
                    const bool keep_mine = (mine_better == (lower == dir_desc));
                    if (!keep_mine) { v[e] = pv; id[e] = pi; }
                }
            }
        }
    }
}

// ---- in-register bitonic networks over 32 values per THREAD (every index is a compile-time constant after unrolling) ----
__device__ __forceinline__ void cex_desc(float& a, float& b) {   // a >= b afterwards
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    a = hi; b = lo;
}
__device__ __forceinline__ void merge32_desc(float (&v)[32]) {    // bitonic sequence -> descending
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((i & stride) == 0) cex_desc(v[i], v[i | stride]);
    }
}
__device__ __forceinline__ void sort32_desc(float (&v)[32]) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if ((i & stride) == 0) {
                    const bool desc = ((i & size) == 0) || (size == 32);
                    if (desc) cex_desc(v[i], v[i | stride]); else cex_desc(v[i | stride], v[i]);
                }
            }
        }
    }
}
// the same with an index payload and the canonical total order: (a, ia) before (b, ib) when a > b, or a == b and ia < ib
__device__ __forceinline__ bool kv_before(float a, int ia, float b, int ib) { return (a > b) || (a == b && ia < ib); }
__device__ __forceinline__ void cex_desc_kv(float& a, int& ia, float& b, int& ib) {   // (a, ia) before (b, ib) afterwards
    const bool swap = kv_before(b, ib, a, ia);
    const float ta = swap ? b : a, tb = swap ? a : b;
    const int tia = swap ? ib : ia, tib = swap ? ia : ib;
    a = ta; b = tb; ia = tia; ib = tib;
}
__device__ __forceinline__ void merge32_desc_kv(float (&v)[32], int (&id)[32]) {
#pragma unroll
    for (int stride = 16; stride > 0; stride >>= 1) {
#pragma unroll
        for (int i = 0; i < 32; ++i)
            if ((i & stride) == 0) cex_desc_kv(v[i], id[i], v[i | stride], id[i | stride]);
    }
}
__device__ __forceinline__ void sort32_desc_kv(float (&v)[32], int (&id)[32]) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if ((i & stride) == 0) {
                    const bool desc = ((i & size) == 0) || (size == 32);
                    if (desc) cex_desc_kv(v[i], id[i], v[i | stride], id[i | stride]);
                    else cex_desc_kv(v[i | stride], id[i | stride], v[i], id[i]);
                }
            }
        }
    }
}

}  // namespace lpd
