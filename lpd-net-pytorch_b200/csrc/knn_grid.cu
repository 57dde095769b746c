// xyz kNN (C == 3) through a uniform grid: same CANONICAL result as lpd_knn (bit-identical index lists), ~16x fewer
// candidate evaluations than the brute-force scan.  Replaces knn() on the Cartesian coordinates, lpdnet_model.py:255 (and
// :317-326), where the reference builds the full N x N distance matrix.
//
// Canonical score:  pd_ij = ((-xx_j) - (-2 dot_ij)) - xx_i  in fp32 with fmaf chains; order = pd descending,
// index ascending.  The grid only decides WHICH candidates are scored; every scored candidate uses exactly that arithmetic
// and the exact order relation, so the result is independent of the grid as long as no true member of the top-k is skipped.
// Skipping is safe because |pd_fp32 + d^2_true| <= 19 u X (u = 2^-24, X = max_j xx_j): a cell / shell is skipped only when
// its lower-bound squared distance exceeds the current k-th squared distance by more than PRUNE_ULPS u X plus a slack for
// the rounding of the cell boundaries.
//
//   kernel 1 (one CTA per cloud): bounding box, max xx, cell of every point, counting sort -> cell-ordered float4
//             (x, y, z, xx) + original index + cell offsets.
//   kernel 2 (one thread per query, queries in cell order so a warp's rows are neighbours in space): visit the shells of
//             cells around the query's cell in growing Chebyshev radius, keep the k best in a per-thread unsorted
//             replace-worst list in shared memory (knn_select.cuh), stop when the list is full and the next shell cannot
//             contain anything better; selection-sort the list into the output row.
#include "common.cuh"
#include <stdlib.h>
#include "knn_select.cuh"

namespace lpd {

constexpr int GRID_MAX_G = 16;                       // cells per axis (<= 4096 cells)
constexpr int GRID_BUILD_THREADS = 1024;
constexpr int GRID_Q_THREADS = 128;
constexpr float PRUNE_ULPS = 64.f;

struct GridHeader {      // per cloud, 16 floats
    float minx, miny, minz, invhx, invhy, invhz, hx, hy, hz, margin, slack;
    int G;
    float pad[4];
};

__device__ __forceinline__ int cell_coord(float v, float mn, float invh, int G) {
    int c = (int)((v - mn) * invh);
    return c < 0 ? 0 : (c >= G ? G - 1 : c);
}

__global__ void __launch_bounds__(GRID_BUILD_THREADS)
knn_grid_build_kernel(const float* __restrict__ x, int N, int G, float4* __restrict__ sorted, int* __restrict__ sidx,
                      int* __restrict__ cell_start, GridHeader* __restrict__ hdr,
                      int* __restrict__ perm_out, int* __restrict__ inv_out, float* __restrict__ xyz_out, int identity_sidx) {
    __shared__ float red[7][32];
    __shared__ GridHeader h;
    __shared__ int counts[GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1];
    const int b = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const float* xb = x + (size_t)b * N * 3;
    const int cells = G * G * G;
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY}, xxm = 0.f;
    for (int i = tid; i < N; i += GRID_BUILD_THREADS) {
        const float a = xb[i * 3], c = xb[i * 3 + 1], d = xb[i * 3 + 2];
        mn[0] = fminf(mn[0], a); mn[1] = fminf(mn[1], c); mn[2] = fminf(mn[2], d);
        mx[0] = fmaxf(mx[0], a); mx[1] = fmaxf(mx[1], c); mx[2] = fmaxf(mx[2], d);
        xxm = fmaxf(xxm, __fmaf_rn(d, d, __fmaf_rn(c, c, __fmaf_rn(a, a, 0.f))));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            mn[a] = fminf(mn[a], __shfl_xor_sync(kFull, mn[a], o));
            mx[a] = fmaxf(mx[a], __shfl_xor_sync(kFull, mx[a], o));
        }
        xxm = fmaxf(xxm, __shfl_xor_sync(kFull, xxm, o));
    }
    if (lane == 0) {
#pragma unroll
        for (int a = 0; a < 3; ++a) { red[a][warp] = mn[a]; red[3 + a][warp] = mx[a]; }
        red[6][warp] = xxm;
    }
    for (int i = tid; i <= cells; i += GRID_BUILD_THREADS) counts[i] = 0;
    __syncthreads();
    if (tid == 0) {
        float lo[3], hi[3], X = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) { lo[a] = INFINITY; hi[a] = -INFINITY; }
        for (int w = 0; w < GRID_BUILD_THREADS / 32; ++w) {
#pragma unroll
            for (int a = 0; a < 3; ++a) { lo[a] = fminf(lo[a], red[a][w]); hi[a] = fmaxf(hi[a], red[3 + a][w]); }
            X = fmaxf(X, red[6][w]);
        }
        float ext[3], amax = 0.f;
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            ext[a] = fmaxf(hi[a] - lo[a], 1e-30f);
            amax = fmaxf(amax, fmaxf(fabsf(lo[a]), fabsf(hi[a])));
        }
        h.minx = lo[0]; h.miny = lo[1]; h.minz = lo[2];
        h.hx = ext[0] / G; h.hy = ext[1] / G; h.hz = ext[2] / G;
        h.invhx = G / ext[0]; h.invhy = G / ext[1]; h.invhz = G / ext[2];
        h.margin = PRUNE_ULPS * 5.9604645e-8f * X;              // score rounding (see header comment)
        h.slack = 16.f * 5.9604645e-8f * (amax + ext[0] + ext[1] + ext[2]) * G;   // cell-boundary rounding, in length units
        h.G = G;
        hdr[b] = h;
    }
    __syncthreads();
    // histogram
    for (int i = tid; i < N; i += GRID_BUILD_THREADS) {
        const int cx = cell_coord(xb[i * 3], h.minx, h.invhx, G), cy = cell_coord(xb[i * 3 + 1], h.miny, h.invhy, G),
                  cz = cell_coord(xb[i * 3 + 2], h.minz, h.invhz, G);
        atomicAdd(&counts[(cz * G + cy) * G + cx], 1);
    }
    __syncthreads();
    // exclusive prefix sum over <= 4096 cells by one warp (128 cells per lane at most)
    if (warp == 0) {
        const int per = (cells + 31) / 32;
        const int c0 = lane * per, c1 = min(cells, c0 + per);
        int s = 0;
        for (int c = c0; c < c1; ++c) s += counts[c];
        int incl = s;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(kFull, incl, o);
            if (lane >= o) incl += t;
        }
        int run = incl - s;
        for (int c = c0; c < c1; ++c) { const int t = counts[c]; counts[c] = run; run += t; }
        if (lane == 31) counts[cells] = N;
    }
    __syncthreads();
    int* cs = cell_start + (size_t)b * (GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1);
    for (int i = tid; i <= cells; i += GRID_BUILD_THREADS) cs[i] = counts[i];
    __syncthreads();
    // stable scatter: inside a cell the points keep their original index order, so the cell order of a cloud is a pure
    // function of the cloud (deterministic, batch-invariant).  Rounds of 1024 points; inside a round the 32 warps reserve
    // their slots one after the other, lanes of a warp rank themselves among their same-cell peers with match.any.
    float4* so = sorted + (size_t)b * N;
    int* si = sidx + (size_t)b * N;
    for (int base = 0; base < N; base += GRID_BUILD_THREADS) {
        const int i = base + tid;
        const bool valid = i < N;
        float a = 0.f, c = 0.f, d = 0.f;
        int cell = cells;                                  // invalid lanes share a dummy cell
        if (valid) {
            a = xb[i * 3]; c = xb[i * 3 + 1]; d = xb[i * 3 + 2];
            cell = (cell_coord(d, h.minz, h.invhz, G) * G + cell_coord(c, h.miny, h.invhy, G)) * G + cell_coord(a, h.minx, h.invhx, G);
        }
        const unsigned peers = __match_any_sync(kFull, cell);
        const int rank = __popc(peers & ((1u << lane) - 1u));
        const int leader = __ffs(peers) - 1;
        int slot = 0;
        for (int w = 0; w < GRID_BUILD_THREADS / 32; ++w) {
            if (warp == w && valid && lane == leader) {
                slot = counts[cell];
                counts[cell] = slot + __popc(peers);
            }
            __syncthreads();
        }
        slot = __shfl_sync(kFull, slot, leader);
        if (valid) {
            const int pos = slot + rank;
            so[pos] = make_float4(a, c, d, __fmaf_rn(d, d, __fmaf_rn(c, c, __fmaf_rn(a, a, 0.f))));
            si[pos] = identity_sidx ? pos : i;             // identity: the grid is handed on as the grid of the RE-ORDERED cloud
            if (perm_out) perm_out[(size_t)b * N + pos] = i;
            if (inv_out) inv_out[(size_t)b * N + i] = pos;
            if (xyz_out) { float* o = xyz_out + ((size_t)b * N + pos) * 3; o[0] = a; o[1] = c; o[2] = d; }
        }
    }
}

template <int GS>
__device__ __forceinline__ void grid_search_one(const float4* __restrict__ sorted, const int* __restrict__ sidx,
                                                const int* __restrict__ cell_start, const GridHeader* __restrict__ hdr, int N, int k,
                                                void* __restrict__ idx_out, int idx_i64, float* __restrict__ lv, int* __restrict__ li,
                                                const int b, const int t) {
    constexpr int L = 4 * GS;
    constexpr int STRIDE = GRID_Q_THREADS + 1;
    const int row = threadIdx.x;
    const GridHeader h = hdr[b];
    const int G = h.G;
    const float4* so = sorted + (size_t)b * N;
    const int* si = sidx + (size_t)b * N;
    const int* cs = cell_start + (size_t)b * (GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1);
    const float4 q = so[t];
    const int self = si[t];
    const int cx = cell_coord(q.x, h.minx, h.invhx, G), cy = cell_coord(q.y, h.miny, h.invhy, G), cz = cell_coord(q.z, h.minz, h.invhz, G);

    RowSelect<GS, true, STRIDE> sel;
    sel.reset();
    const int need = L < N ? L : N;          // the list can only fill up to N entries

    for (int r = 0; r < G; ++r) {
        // ---- scan the shell of Chebyshev radius r ----
        const int z0 = max(cz - r, 0), z1 = min(cz + r, G - 1);
        const int y0 = max(cy - r, 0), y1 = min(cy + r, G - 1);
        const int x0 = max(cx - r, 0), x1 = min(cx + r, G - 1);
        for (int zz = z0; zz <= z1; ++zz) {
            const bool zface = (zz == cz - r) || (zz == cz + r);
            const float dz = fmaxf(0.f, fmaxf(h.minz + zz * h.hz - q.z, q.z - (h.minz + (zz + 1) * h.hz)) - h.slack);
            for (int yy = y0; yy <= y1; ++yy) {
                const bool yface = (yy == cy - r) || (yy == cy + r);
                const float dy = fmaxf(0.f, fmaxf(h.miny + yy * h.hy - q.y, q.y - (h.miny + (yy + 1) * h.hy)) - h.slack);
                const float dzy = dz * dz + dy * dy;
                // inside the shell only the two x-faces belong to it unless this (z, y) line lies on a z- or y-face
                const int step = (zface || yface) ? 1 : max(x1 - x0, 1);
                for (int xx = x0; xx <= x1; xx += step) {
                    if (!(zface || yface) && !((xx == cx - r) || (xx == cx + r))) continue;
                    if (sel.filled >= need) {
                        const float dx = fmaxf(0.f, fmaxf(h.minx + xx * h.hx - q.x, q.x - (h.minx + (xx + 1) * h.hx)) - h.slack);
                        if (dzy + dx * dx > -sel.tau + h.margin) continue;       // nothing in this cell can enter the list
                    }
                    const int cell = (zz * G + yy) * G + xx;
                    const int p0 = cs[cell], p1 = cs[cell + 1];
                    for (int p = p0; p < p1; ++p) {
                        const float4 c = __ldg(so + p);
                        const float dot = __fmaf_rn(q.z, c.z, __fmaf_rn(q.y, c.y, __fmaf_rn(q.x, c.x, 0.f)));
                        const float tt = -2.0f * dot;
                        const float u = __fsub_rn(-c.w, tt);
                        const float pd = __fsub_rn(u, q.w);
                        if (pd >= sel.tau) {                                      // cheap filter; exact order inside
                            const int j = __ldg(si + p);
                            if (sel.filled < L || sel.passes(pd, j)) sel.insert(lv, li, row, pd, j);
                        }
                    }
                }
            }
        }
        // ---- can anything outside the (2r+1)^3 block still enter? ----
        if (sel.filled >= need) {
            float dmin = INFINITY;
            if (cx - r > 0) dmin = fminf(dmin, q.x - (h.minx + (cx - r) * h.hx));
            if (cx + r < G - 1) dmin = fminf(dmin, (h.minx + (cx + r + 1) * h.hx) - q.x);
            if (cy - r > 0) dmin = fminf(dmin, q.y - (h.miny + (cy - r) * h.hy));
            if (cy + r < G - 1) dmin = fminf(dmin, (h.miny + (cy + r + 1) * h.hy) - q.y);
            if (cz - r > 0) dmin = fminf(dmin, q.z - (h.minz + (cz - r) * h.hz));
            if (cz + r < G - 1) dmin = fminf(dmin, (h.minz + (cz + r + 1) * h.hz) - q.z);
            dmin = fmaxf(0.f, dmin - h.slack);
            if (dmin == INFINITY || dmin * dmin > -sel.tau + h.margin) break;
        }
    }

    // ---- selection sort of the list into the output row (pd descending, index ascending) ----
    const int have = sel.filled;
    const size_t o = ((size_t)b * N + self) * k;
    float pv = INFINITY;
    int pi = -1;
    for (int outp = 0; outp < k; ++outp) {
        float bv = -INFINITY;
        int bi = INT_MAX;
        for (int s = 0; s < have; ++s) {
            const float v = lv[s * STRIDE + row];
            const int j = li[s * STRIDE + row];
            const bool after_prev = (v < pv) || (v == pv && j > pi);               // strictly after the previous output
            const bool better = (v > bv) || (v == bv && j < bi);
            if (after_prev && better) { bv = v; bi = j; }
        }
        pv = bv; pi = bi;
        if (idx_i64) reinterpret_cast<long long*>(idx_out)[o + outp] = bi;
        else reinterpret_cast<int*>(idx_out)[o + outp] = bi;
    }
}

// rlist == nullptr: one thread per query of the whole batch (grid = [ceil(N / 128), B]).
// rlist != nullptr: rescue mode of the lock-step kernel below: rlist[0] queries, rlist[1 + i] = cloud * N + cell-order position;
//                   a fixed grid walks the list.  Only lists longer than rlimit (degenerate clouds: most queries flagged) are
//                   taken here; the usual short list goes to knn_xyz_rescue_warp_kernel.
template <int GS>
__global__ void __launch_bounds__(GRID_Q_THREADS)
knn_grid_search_kernel(const float4* __restrict__ sorted, const int* __restrict__ sidx, const int* __restrict__ cell_start,
                       const GridHeader* __restrict__ hdr, int N, int k, void* __restrict__ idx_out, int idx_i64,
                       const int* __restrict__ rlist, int rlimit) {
    constexpr int L = 4 * GS;
    constexpr int STRIDE = GRID_Q_THREADS + 1;
    __shared__ float lv[L * STRIDE];
    __shared__ int li[L * STRIDE];
    if (rlist == nullptr) {
        const int t = blockIdx.x * GRID_Q_THREADS + threadIdx.x;
        if (t < N) grid_search_one<GS>(sorted, sidx, cell_start, hdr, N, k, idx_out, idx_i64, lv, li, blockIdx.y, t);
        return;
    }
    // entry i goes to warp i mod (number of warps), lane i / (number of warps): a short list is spread over all warps of the grid
    // (a thread-per-query search is a long serial job; lanes of one warp diverge anyway)
    const int count = rlist[0];
    if (count <= rlimit) return;                           // short list: the warp-per-query kernel below has finished it
    const int nwarps = gridDim.x * (GRID_Q_THREADS / 32), gw = blockIdx.x * (GRID_Q_THREADS / 32) + (threadIdx.x >> 5);
    for (long long i = (long long)(threadIdx.x & 31) * nwarps + gw; i < count; i += 32ll * nwarps) {
        const int g = rlist[1 + i];
        grid_search_one<GS>(sorted, sidx, cell_start, hdr, N, k, idx_out, idx_i64, lv, li, g / N, g % N);
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// Lock-step search (the default for k <= 20).  The thread-per-query kernel above executes 39 k warp instructions per 32 queries:
// its lanes sit in ~4 different cells, walk different candidate ranges and insert into their lists at different times (12.9 of
// 32 lanes active on average), and two thirds of the per-thread work is the upkeep of the replace-worst list.  Here the 32
// queries of a warp (consecutive in cell order: a run of ~4 cells of one cell row) share ONE candidate set, never maintain a list
// and finish cooperatively:
//   segments  the lanes are grouped by cell row (1.45 groups per warp on average); a segment's candidate box is its cell range
//             +- 1 cell in x, y and z, i.e. 9 contiguous ranges of the cell-ordered cloud, staged into shared memory as PAIRS
//             (x0 x1 y0 y1 | z0 z1 -xx0 -xx1) so that one packed FFMA2 chain scores two candidates (same IEEE results);
//   sweep 1   every lane scores every staged candidate (broadcast LDS.128) and keeps the running maximum of 32 strided groups
//             (candidate i -> group i mod 32) in registers: no branches.  The k-th largest group maximum tau (in-register bitonic
//             network) is a LOWER bound of the lane's k-th best score: k distinct candidates reach it;
//   sweep 2   the same candidates again: the staged position of everything with score >= tau (27 per query on average) is
//             appended to the lane's list (two predicated instructions per candidate);
//   finish    query by query, the whole warp: lane = list entry; canonical score and original index of the entry, shuffle
//             bitonic sort by (score descending, index ascending), sufficiency test (the k-th distance must not reach the nearest
//             interior face of the box, same rounding margins as above), coalesced write of the k indices.
// Lanes of the other segments ride along with q.w = +inf (every score -inf).  Queries whose list overflows, whose box does not
// suffice or overflows the staging buffer, or whose warp spans more than 4 cell rows are appended to a rescue list that the
// thread-per-query kernel finishes (~2 % of the queries on uniform clouds).  Bit-identical output.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int GL_WARPS = 2;          // warps per CTA
constexpr int GL_CAND = 576;         // staged candidates per segment (a multiple of 64)
constexpr int GL_MAXSEG = 4;

__device__ __forceinline__ float canonical_pd(const float qx, const float qy, const float qz, const float qw, const float cx,
                                              const float cy, const float cz, const float negxx) {
    const float dot = __fmaf_rn(qz, cz, __fmaf_rn(qy, cy, __fmaf_rn(qx, cx, 0.f)));
    const float tt = -2.0f * dot;
    return __fsub_rn(__fsub_rn(negxx, tt), qw);
}
// the same arithmetic on two candidates at once (packed fp32 pairs: every lane of FFMA2 / FADD2 rounds like the scalar instruction)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float v) { f32x2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 canonical_pd2(f32x2 qx, f32x2 qy, f32x2 qz, f32x2 qw, f32x2 zero, f32x2 m2, const ulonglong2& xy, const ulonglong2& zw) {
    const f32x2 dot = fma2(qz, zw.x, fma2(qy, xy.y, fma2(qx, xy.x, zero)));
    const f32x2 tt = mul2(m2, dot);
    return sub2(sub2(zw.y, tt), qw);
}

struct GridLockSmem {
    static constexpr size_t off_idx = (size_t)GL_WARPS * GL_CAND * 16;
    static constexpr size_t off_mask = off_idx + (size_t)GL_WARPS * GL_CAND * 4;
    static constexpr size_t total = off_mask + (size_t)GL_WARPS * (GL_CAND / 32) * 32 * 4;
};

// in-register bitonic sort of 32 (score, index) entries per thread by score descending; equal scores keep an arbitrary order
__device__ __forceinline__ void cex_desc_fv(float& a, int& ia, float& b, int& ib) {   // a >= b afterwards
    const bool sw = a < b;
    const float hi = fmaxf(a, b), lo = fminf(a, b);
    const int ihi = sw ? ib : ia, ilo = sw ? ia : ib;
    a = hi; b = lo; ia = ihi; ib = ilo;
}
__device__ __forceinline__ void sort32_desc_fv(float (&v)[32], int (&id)[32]) {
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
#pragma unroll
            for (int i = 0; i < 32; ++i) {
                if ((i & stride) == 0) {
                    const bool desc = ((i & size) == 0) || (size == 32);
                    if (desc) cex_desc_fv(v[i], id[i], v[i | stride], id[i | stride]);
                    else cex_desc_fv(v[i | stride], id[i | stride], v[i], id[i]);
                }
            }
        }
    }
}

struct GridBox { int x0, x1, y0, y1, z0, z1; };

// Stages the candidate box of segment s of a warp (the lanes with myseg == s: one cell row, cells xmin .. xmax): the cell range
// +- 1 cell in x, y and z = up to 9 contiguous ranges of the cell-ordered cloud, as pairs [x0 x1 y0 y1 | z0 z1 -xx0 -xx1] plus the
// original indices, padded with sentinels (score -inf) to a multiple of 64.  Returns the number of candidates, -1 = the box does
// not fit the buffer.  Not inlined: the kernel calls it from two places and has to fit the instruction cache.
__device__ __noinline__ int gl_stage(int s, int myseg, int rowid, int cx, int G, const int* __restrict__ cs,
                                     const float4* __restrict__ so, const int* __restrict__ si, float* __restrict__ candf,
                                     int* __restrict__ cidx, int2* __restrict__ rng, GridBox* box) {
    const int lane = threadIdx.x & 31;
    const bool inseg = myseg == s;
    const int r = __shfl_sync(kFull, rowid, __ffs(__ballot_sync(kFull, inseg)) - 1);
    const int xmin = __reduce_min_sync(kFull, inseg ? cx : INT_MAX), xmax = __reduce_max_sync(kFull, inseg ? cx : -1);
    const int rz = r / G, ry = r % G;
    GridBox bx;
    bx.x0 = max(xmin - 1, 0); bx.x1 = min(xmax + 1, G - 1);
    bx.y0 = max(ry - 1, 0); bx.y1 = min(ry + 1, G - 1);
    bx.z0 = max(rz - 1, 0); bx.z1 = min(rz + 1, G - 1);
    *box = bx;
    // the <= 9 ranges laid end to end on one flat index (lane = range): rng[r] = (start - first flat index, one past the last flat
    // index), so that the copy below keeps four independent loads per lane in flight whatever the range lengths (walking the
    // ranges one by one made the staging a chain of dependent L2 round trips: a third of the kernel's stall samples)
    const int ny = bx.y1 - bx.y0 + 1, nr = ny * (bx.z1 - bx.z0 + 1);
    int start = 0, len = 0;
    if (lane < nr) {
        const int row = ((bx.z0 + lane / ny) * G + bx.y0 + lane % ny) * G;
        start = __ldg(cs + row + bx.x0);
        len = __ldg(cs + row + bx.x1 + 1) - start;
    }
    int incl = len;
#pragma unroll
    for (int sft = 1; sft < 16; sft <<= 1) {
        const int up = __shfl_up_sync(kFull, incl, sft);
        if (lane >= sft) incl += up;
    }
    const int T = __shfl_sync(kFull, incl, 15);
    __syncwarp();                                          // the previous contents have been consumed
    if (T > GL_CAND) return -1;
    if (lane < nr) rng[lane] = make_int2(start - (incl - len), incl);
    __syncwarp();
    int rc = 0;
#pragma unroll 1
    for (int f0 = 0; f0 < T; f0 += 128) {
        float4 c[4];
        int id[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int f = f0 + 32 * u + lane;
            if (f < T) {
                while (f >= rng[rc].y) ++rc;
                const int p = f + rng[rc].x;
                c[u] = __ldg(so + p);
                id[u] = __ldg(si + p);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = f0 + 32 * u + lane;
            if (i < T) {
                float* d = candf + (i >> 1) * 8 + (i & 1);
                d[0] = c[u].x; d[2] = c[u].y; d[4] = c[u].z; d[6] = -c[u].w;
                cidx[i] = id[u];
            }
        }
    }
    const int Tp = (T + 63) & ~63;
#pragma unroll 1
    for (int i = T + lane; i < Tp; i += 32) {
        float* d = candf + (i >> 1) * 8 + (i & 1);
        d[0] = 0.f; d[2] = 0.f; d[4] = 0.f; d[6] = -INFINITY;
    }
    __syncwarp();
    return T;
}

__global__ void __launch_bounds__(GL_WARPS * 32, 8)
knn_grid_lockstep_kernel(const float4* __restrict__ sorted, const int* __restrict__ sidx, const int* __restrict__ cell_start,
                         const GridHeader* __restrict__ hdr, int N, int k, void* __restrict__ idx_out, int idx_i64,
                         int* __restrict__ rlist) {
    using S = GridLockSmem;
    __shared__ __align__(16) uint8_t gsm[S::total];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* candf = reinterpret_cast<float*>(gsm) + (size_t)w * GL_CAND * 4;             // pairs: [x0 x1 y0 y1 | z0 z1 -xx0 -xx1]
    const ulonglong2* candp = reinterpret_cast<const ulonglong2*>(candf);
    int* cidx = reinterpret_cast<int*>(gsm + S::off_idx) + w * GL_CAND;
    uint32_t* maskbuf = reinterpret_cast<uint32_t*>(gsm + S::off_mask) + w * (GL_CAND / 32) * 32;   // [32-candidate block][lane]
    __shared__ int2 rng_s[GL_WARPS][9];
    int2* rng = rng_s[w];
    const int b = blockIdx.y;
    const int t0 = (blockIdx.x * GL_WARPS + w) * 32;
    if (t0 >= N) return;                                   // the whole warp is out of range
    const int t = t0 + lane;
    const bool live = t < N;
    const GridHeader h = hdr[b];
    const int G = h.G;
    const float4* so = sorted + (size_t)b * N;
    const int* si = sidx + (size_t)b * N;
    const int* cs = cell_start + (size_t)b * (GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1);
    const float4 q = so[live ? t : N - 1];
    const int cx = cell_coord(q.x, h.minx, h.invhx, G), cy = cell_coord(q.y, h.miny, h.invhy, G), cz = cell_coord(q.z, h.minz, h.invhz, G);
    const int rowid = live ? cz * G + cy : -1;
    const f32x2 QX = pack2(q.x), QY = pack2(q.y), QZ = pack2(q.z), QWT = pack2(q.w), ZERO = pack2(0.f), M2 = pack2(-2.0f);

    // ---- segments: lanes of one cell row ----
    int myseg = -1, nseg = 0;
    {
        unsigned todo = __ballot_sync(kFull, live);
        while (todo && nseg < GL_MAXSEG) {
            const int r = __shfl_sync(kFull, rowid, __ffs(todo) - 1);
            const unsigned seg = __ballot_sync(kFull, rowid == r);
            if (rowid == r) myseg = nseg;
            todo &= ~seg;
            ++nseg;
        }
    }
    bool rescue = live && myseg < 0;                       // more than GL_MAXSEG cell rows in this warp
    GridBox box = {0, 0, 0, 0, 0, 0};                      // the box of the staged segment (warp-uniform)

    // ---- sweep 1: group maxima ----
    float tau;
    int Tlast = 0;
    {
        float m[64];
#pragma unroll
        for (int j = 0; j < 64; ++j) m[j] = -INFINITY;
#pragma unroll 1
        for (int s = 0; s < nseg; ++s) {
            const int T = gl_stage(s, myseg, rowid, cx, G, cs, so, si, candf, cidx, rng, &box);
            Tlast = T;
            if (T < 0) { if (myseg == s) rescue = true; continue; }
            const f32x2 QW = (myseg == s) ? QWT : pack2(INFINITY);       // lanes of the other segments: every score -inf
#pragma unroll 1
            for (int base = 0; base < T; base += 64) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const ulonglong2 xy = candp[base + 2 * j], zw = candp[base + 2 * j + 1];
                    float p0, p1;
                    unpack2(canonical_pd2(QX, QY, QZ, QW, ZERO, M2, xy, zw), p0, p1);
                    m[2 * j] = fmaxf(m[2 * j], p0);
                    m[2 * j + 1] = fmaxf(m[2 * j + 1], p1);
                }
            }
        }
        // ---- tau: a lower bound of the k-th largest of the 64 group maxima, by bisection on the value (a rolled loop of 64
        //      compare-and-count steps: any tau with at least k group maxima at or above it is valid, and 9 halvings of the
        //      range of the maxima leave it a small fraction of a neighbour rank below the exact value) ----
        float lo = INFINITY, hi = -INFINITY;
        int nfin = 0;
#pragma unroll
        for (int j = 0; j < 64; ++j) {
            const bool fin = m[j] > -INFINITY;
            lo = fminf(lo, fin ? m[j] : INFINITY);
            hi = fmaxf(hi, m[j]);
            nfin += fin;
        }
        if (nfin < k) lo = -INFINITY;                      // fewer than k non-empty groups: everything is collected
        else {
#pragma unroll 1
            for (int it = 0; it < 9; ++it) {
                const float mid = 0.5f * lo + 0.5f * hi;
                int c = 0;
#pragma unroll
                for (int j = 0; j < 64; ++j) c += m[j] >= mid;
                if (c >= k) lo = mid; else hi = mid;
            }
        }
        tau = lo;
    }
    // ---- sweep 2, segment by segment: bit mask of the staged candidates at or above tau, then every lane of the segment pulls
    //      its entries (64-bit sort keys) into key[] before the buffer is restaged ----
    float kv[32];
    int ki[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) { kv[e] = -INFINITY; ki[e] = INT_MAX; }   // padding: worse than anything
    int cnt = 0;
    float dmin2 = INFINITY;                                // squared distance to the nearest interior face of the lane's box
#pragma unroll 1
    for (int s = 0; s < nseg; ++s) {
        const int T = (nseg == 1) ? Tlast : gl_stage(s, myseg, rowid, cx, G, cs, so, si, candf, cidx, rng, &box);   // a single segment is still staged
        if (T < 0) continue;
        const bool inseg = myseg == s;
        const float tq = inseg ? tau : INFINITY;           // lanes of the other segments collect nothing (a score is never +inf)
#pragma unroll 1
        for (int base = 0; base < T; base += 32) {
            uint32_t mask = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const ulonglong2 xy = candp[base + 2 * j], zw = candp[base + 2 * j + 1];
                float p0, p1;
                unpack2(canonical_pd2(QX, QY, QZ, QWT, ZERO, M2, xy, zw), p0, p1);
                mask |= (p0 >= tq) ? (1u << (2 * j)) : 0u;
                mask |= (p1 >= tq) ? (2u << (2 * j)) : 0u;
            }
            maskbuf[(base >> 5) * 32 + lane] = mask;
            cnt += __popc(mask);
        }
        if (inseg) {
            float dmin = INFINITY;
            if (box.x0 > 0) dmin = fminf(dmin, q.x - (h.minx + box.x0 * h.hx));
            if (box.x1 < G - 1) dmin = fminf(dmin, (h.minx + (box.x1 + 1) * h.hx) - q.x);
            if (box.y0 > 0) dmin = fminf(dmin, q.y - (h.miny + box.y0 * h.hy));
            if (box.y1 < G - 1) dmin = fminf(dmin, (h.miny + (box.y1 + 1) * h.hy) - q.y);
            if (box.z0 > 0) dmin = fminf(dmin, q.z - (h.minz + box.z0 * h.hz));
            if (box.z1 < G - 1) dmin = fminf(dmin, (h.minz + (box.z1 + 1) * h.hz) - q.z);
            dmin = fmaxf(0.f, dmin - h.slack);
            dmin2 = dmin * dmin;                           // (+inf when the box is the whole grid)
        }
        // entries of this lane: the set bits of its mask words, in order (own words: no synchronisation needed)
        int wi = -1;
        uint32_t word = 0;
        const int take = (inseg && cnt <= 32) ? cnt : 0;
#pragma unroll 1
        for (int e = 0; e < take; ++e) {                   // (rolled: kv[] / ki[] sit in local memory here)
            while (word == 0) word = maskbuf[(++wi) * 32 + lane];
            const int slot = wi * 32 + __ffs(word) - 1;
            word &= word - 1;
            const float* c = candf + (slot >> 1) * 8 + (slot & 1);
            kv[e] = __fadd_rn(canonical_pd(q.x, q.y, q.z, q.w, c[0], c[2], c[4], c[6]), 0.0f);   // (+ 0: one representation of zero)
            ki[e] = cidx[slot];
        }
    }
    if (live && (cnt > 32 || cnt < k)) rescue = true;      // list overflow (masses of ties) / fewer than k points in the box
    // ---- rank: in-register bitonic sort by score (two FMNMX + one compare + two selects per exchange); entries with EQUAL scores
    //      keep an arbitrary order, so a lane with a tie among its first k + 1 entries goes to the rescue list (the canonical
    //      order breaks ties by the original index; exact ties only occur for duplicated points and lattices) ----
    sort32_desc_fv(kv, ki);
    {
        bool tie = false;
#pragma unroll
        for (int j = 0; j < 31; ++j) tie |= (j < k) && (kv[j] == kv[j + 1]);
        const float kth = kv[k - 1];
        if (tie || !(dmin2 > -kth + h.margin)) rescue = rescue || live;   // (second test: something outside the box could still be closer)
    }
    if (live && !rescue) {
        const size_t o = ((size_t)b * N + __ldg(si + t)) * k;
#pragma unroll 1
        for (int j = 0; j < k; ++j) {
            if (idx_i64) reinterpret_cast<long long*>(idx_out)[o + j] = ki[j];
            else reinterpret_cast<int*>(idx_out)[o + j] = ki[j];
        }
    }
    if (live && rescue) rlist[1 + atomicAdd(rlist, 1)] = b * N + t;
}

// ---------------------------------------------------------------------------------------------------------------------
// Rescue of a SHORT list (the usual case: ~5 % of the queries, mostly in boundary cells where the k-th neighbour lies beyond the
// +-1 box): one warp per listed query.  The thread-per-query kernel is a ~50 us serial job per query and a short list leaves 30
// of its 32 lanes idle (0.14 ms for 13 k queries).  The warp first searches the +-2 cell box of the query (25 contiguous ranges
// of the cell-ordered cloud) with the same sufficiency test as above and, if that fails, the whole cloud (no box argument at all):
//   pass 1   lane l scores candidates l, l + 32, ... of every range and keeps the maxima of its even and odd rounds: 64 strided
//            group maxima; tau = the largest group maximum that at least k group maxima reach (a lower bound of the k-th best);
//   pass 2   the same scores again; everything >= tau is appended to a shared-memory list by ballot compaction;
//   rank     entry e's output position = the number of listed entries before it in the canonical order (keys are distinct:
//            the original index breaks ties); positions < k are written.
// A whole-cloud list that overflows (masses of exact ties) is replaced by k rounds of "best entry after the previous output":
// slow, always correct.
// ---------------------------------------------------------------------------------------------------------------------
constexpr int RW_WARPS = 4;
constexpr int RW_LIST = 128;
constexpr int RW_R = 2;                                   // box radius in cells
constexpr int RW_RANGES = (2 * RW_R + 1) * (2 * RW_R + 1);

__global__ void __launch_bounds__(RW_WARPS * 32)
knn_xyz_rescue_warp_kernel(const float4* __restrict__ sorted, const int* __restrict__ sidx, const int* __restrict__ cell_start,
                           const GridHeader* __restrict__ hdr, int N, int k, void* __restrict__ idx_out, int idx_i64,
                           const int* __restrict__ rlist, int rlimit) {
    __shared__ float lvs[RW_WARPS][RW_LIST];
    __shared__ int lis[RW_WARPS][RW_LIST];
    __shared__ int2 rngs[RW_WARPS][RW_RANGES];
    const int count = rlist[0];
    if (count > rlimit) return;                            // long list: the thread-per-query kernel takes it
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    float* lv = lvs[w];
    int* li = lis[w];
    int2* rng = rngs[w];
    const unsigned lt = (1u << lane) - 1u;
    const int nwarps = gridDim.x * RW_WARPS;
#pragma unroll 1
    for (int i = blockIdx.x * RW_WARPS + w; i < count; i += nwarps) {
        const int g = rlist[1 + i];
        const int b = g / N, t = g - b * N;
        const float4* so = sorted + (size_t)b * N;
        const int* si = sidx + (size_t)b * N;
        const int* cs = cell_start + (size_t)b * (GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1);
        const GridHeader h = hdr[b];
        const int G = h.G;
        const float4 q = so[t];
        const size_t o = ((size_t)b * N + si[t]) * k;
        const int cx = cell_coord(q.x, h.minx, h.invhx, G), cy = cell_coord(q.y, h.miny, h.invhy, G), cz = cell_coord(q.z, h.minz, h.invhz, G);
        const int x0 = max(cx - RW_R, 0), x1 = min(cx + RW_R, G - 1), y0 = max(cy - RW_R, 0), y1 = min(cy + RW_R, G - 1),
                  z0 = max(cz - RW_R, 0), z1 = min(cz + RW_R, G - 1);
        float dmin2;
        {
            float dmin = INFINITY;
            if (x0 > 0) dmin = fminf(dmin, q.x - (h.minx + x0 * h.hx));
            if (x1 < G - 1) dmin = fminf(dmin, (h.minx + (x1 + 1) * h.hx) - q.x);
            if (y0 > 0) dmin = fminf(dmin, q.y - (h.miny + y0 * h.hy));
            if (y1 < G - 1) dmin = fminf(dmin, (h.miny + (y1 + 1) * h.hy) - q.y);
            if (z0 > 0) dmin = fminf(dmin, q.z - (h.minz + z0 * h.hz));
            if (z1 < G - 1) dmin = fminf(dmin, (h.minz + (z1 + 1) * h.hz) - q.z);
            dmin = fmaxf(0.f, dmin - h.slack);
            dmin2 = dmin * dmin;                           // (+inf when the box is the whole grid)
        }
        bool done = false;
#pragma unroll 1
        for (int attempt = 0; attempt < 2 && !done; ++attempt) {
            // ---- the candidate ranges: the rows of the box, or the whole cloud.  The ranges are laid end to end on one flat index
            //      f (rs[r] = start - first flat index, pe[r] = one past the last flat index of range r), so that a warp iteration
            //      covers 128 consecutive candidates with four independent loads in flight per lane whatever the range lengths
            //      (the search is a latency chain of L2 loads: 25 short ranges walked one by one took 13 us per query) ----
            int nr, total;
            __syncwarp();
            {
                int start = 0, len = 0;
                if (attempt == 0) {
                    const int ny = y1 - y0 + 1;
                    nr = ny * (z1 - z0 + 1);
                    if (lane < nr) {
                        const int row = ((z0 + lane / ny) * G + y0 + lane % ny) * G;
                        start = cs[row + x0];
                        len = cs[row + x1 + 1] - start;
                    }
                } else {
                    nr = 1;
                    if (lane == 0) len = N;
                }
                int incl = len;
#pragma unroll
                for (int sft = 1; sft < 32; sft <<= 1) {
                    const int up = __shfl_up_sync(kFull, incl, sft);
                    if (lane >= sft) incl += up;
                }
                total = __shfl_sync(kFull, incl, 31);
                if (lane < nr) rng[lane] = make_int2(start - (incl - len), incl);      // (rs, pe)
            }
            __syncwarp();
            if (total < k && attempt == 0) continue;       // fewer than k points in the box
            // ---- pass 1: 64 strided group maxima ----
            float m0 = -INFINITY, m1 = -INFINITY;
            {
                int rc = 0;
#pragma unroll 1
                for (int f0 = 0; f0 < total; f0 += 128) {
                    float4 c[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int f = f0 + 32 * u + lane;
                        c[u] = make_float4(0.f, 0.f, 0.f, INFINITY);                  // (xx = +inf: score -inf)
                        if (f < total) {
                            while (f >= rng[rc].y) ++rc;
                            c[u] = __ldg(so + f + rng[rc].x);
                        }
                    }
                    m0 = fmaxf(m0, fmaxf(canonical_pd(q.x, q.y, q.z, q.w, c[0].x, c[0].y, c[0].z, -c[0].w),
                                         canonical_pd(q.x, q.y, q.z, q.w, c[2].x, c[2].y, c[2].z, -c[2].w)));
                    m1 = fmaxf(m1, fmaxf(canonical_pd(q.x, q.y, q.z, q.w, c[1].x, c[1].y, c[1].z, -c[1].w),
                                         canonical_pd(q.x, q.y, q.z, q.w, c[3].x, c[3].y, c[3].z, -c[3].w)));
                }
            }
            float tau = -INFINITY;
#pragma unroll 1
            for (int src = 0; src < 32; ++src) {
                const float v0 = __shfl_sync(kFull, m0, src), v1 = __shfl_sync(kFull, m1, src);
                const int c0 = __popc(__ballot_sync(kFull, m0 >= v0)) + __popc(__ballot_sync(kFull, m1 >= v0));
                const int c1 = __popc(__ballot_sync(kFull, m0 >= v1)) + __popc(__ballot_sync(kFull, m1 >= v1));
                if (c0 >= k) tau = fmaxf(tau, v0);
                if (c1 >= k) tau = fmaxf(tau, v1);
            }
            // ---- pass 2: collect everything at or above tau ----
            int cnt = 0;
            {
                int rc = 0;
#pragma unroll 1
                for (int f0 = 0; f0 < total; f0 += 128) {
                    float4 c[4];
                    int pp[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const int f = f0 + 32 * u + lane;
                        c[u] = make_float4(0.f, 0.f, 0.f, INFINITY);
                        pp[u] = -1;
                        if (f < total) {
                            while (f >= rng[rc].y) ++rc;
                            pp[u] = f + rng[rc].x;
                            c[u] = __ldg(so + pp[u]);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float pd = canonical_pd(q.x, q.y, q.z, q.w, c[u].x, c[u].y, c[u].z, -c[u].w);
                        const bool pass = pp[u] >= 0 && pd >= tau;
                        const unsigned bal = __ballot_sync(kFull, pass);
                        const int pos = cnt + __popc(bal & lt);
                        if (pass && pos < RW_LIST) { lv[pos] = pd; li[pos] = __ldg(si + pp[u]); }
                        cnt += __popc(bal);
                    }
                }
            }
            __syncwarp();
            if (cnt <= RW_LIST) {
                // ---- rank by counting; in the box attempt the k-th score must not reach the nearest interior face ----
                int rk[RW_LIST / 32], jj[RW_LIST / 32];
                float kth = -INFINITY;
#pragma unroll
                for (int u = 0; u < RW_LIST / 32; ++u) {
                    const int e = lane + 32 * u;
                    rk[u] = INT_MAX; jj[u] = 0;
                    if (e < cnt) {
                        const float v = lv[e];
                        const int j = li[e];
                        int rank = 0;
#pragma unroll 4
                        for (int f = 0; f < cnt; ++f) rank += kv_before(lv[f], li[f], v, j) ? 1 : 0;
                        rk[u] = rank; jj[u] = j;
                        if (rank == k - 1) kth = v;            // the k-th best score of the list
                    }
                }
                bool ok = true;
                if (attempt == 0) {
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) kth = fmaxf(kth, __shfl_xor_sync(kFull, kth, sft));
                    ok = dmin2 > -kth + h.margin;              // (false when cnt < k: kth stays -inf)
                }
                if (ok) {
#pragma unroll
                    for (int u = 0; u < RW_LIST / 32; ++u)
                        if (rk[u] < k) {
                            if (idx_i64) reinterpret_cast<long long*>(idx_out)[o + rk[u]] = jj[u];
                            else reinterpret_cast<int*>(idx_out)[o + rk[u]] = jj[u];
                        }
                    done = true;
                }
            } else if (attempt == 1) {
                // ---- overflow on the whole cloud: k rounds of "best entry strictly after the previous output" ----
                float pv = INFINITY;
                int pi = -1;
#pragma unroll 1
                for (int outp = 0; outp < k; ++outp) {
                    float bv = -INFINITY;
                    int bi = INT_MAX;
#pragma unroll 1
                    for (int p = lane; p < N; p += 32) {
                        const float4 c = __ldg(so + p);
                        const float v = canonical_pd(q.x, q.y, q.z, q.w, c.x, c.y, c.z, -c.w);
                        const int j = __ldg(si + p);
                        const bool after_prev = (v < pv) || (v == pv && j > pi);
                        if (after_prev && kv_before(v, j, bv, bi)) { bv = v; bi = j; }
                    }
#pragma unroll
                    for (int sft = 16; sft > 0; sft >>= 1) {
                        const float ov = __shfl_xor_sync(kFull, bv, sft);
                        const int oi = __shfl_xor_sync(kFull, bi, sft);
                        if (kv_before(ov, oi, bv, bi)) { bv = ov; bi = oi; }
                    }
                    pv = bv; pi = bi;
                    if (lane == 0) {
                        if (idx_i64) reinterpret_cast<long long*>(idx_out)[o + outp] = bi;
                        else reinterpret_cast<int*>(idx_out)[o + outp] = bi;
                    }
                }
                done = true;
            }
        }
        __syncwarp();
    }
}

// 1: lock-step search + rescue kernels (default); 0: thread-per-query search only (LPD_KNN_GRID_LOCKSTEP=0)
static int g_grid_lockstep = [] { const char* e = getenv("LPD_KNN_GRID_LOCKSTEP"); return (e && atoi(e) == 0) ? 0 : 1; }();

static int grid_cells_per_axis(int N) {
    int G = (int)lroundf(cbrtf((float)N / 8.f));
    return G < 1 ? 1 : (G > GRID_MAX_G ? GRID_MAX_G : G);
}

}  // namespace lpd

using namespace lpd;

extern "C" size_t lpd_knn_xyz_workspace_bytes(int B, int N) {
    if (B < 1 || N < 1) return 0;
    const size_t cells1 = GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1;
    return (size_t)B * N * (sizeof(float4) + sizeof(int)) + (size_t)B * cells1 * sizeof(int) + (size_t)B * sizeof(GridHeader) + 256
           + ((size_t)B * N + 4) * sizeof(int);                // + the rescue list of the lock-step search
}

static int knn_xyz_impl(const float* x, int B, int N, int k, void* idx, int idx_i64, void* workspace, size_t workspace_bytes,
                        void* stream) {
    LPD_REQUIRE(idx && workspace);
    LPD_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && k >= 1 && k <= 32 && k <= N);
    LPD_REQUIRE(((uintptr_t)workspace & 15) == 0);
    if (workspace_bytes < lpd_knn_xyz_workspace_bytes(B, N)) return LPD_EWORKSPACE;
    const size_t cells1 = GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1;
    float4* sorted = reinterpret_cast<float4*>(workspace);
    int* sidx = reinterpret_cast<int*>(sorted + (size_t)B * N);
    int* cell_start = sidx + (size_t)B * N;
    GridHeader* hdr = reinterpret_cast<GridHeader*>(cell_start + (size_t)B * cells1);
    // header must be 16-byte aligned for the struct copy: B*N*20 + B*cells1*4 is a multiple of 4 only -> round up
    hdr = reinterpret_cast<GridHeader*>((reinterpret_cast<uintptr_t>(hdr) + 15) & ~(uintptr_t)15);
    const int G = grid_cells_per_axis(N);
    cudaStream_t st = as_stream(stream);
    if (x) {                                               // x == nullptr: the workspace already holds the grid (lpd_knn_xyz_ordered)
        knn_grid_build_kernel<<<B, GRID_BUILD_THREADS, 0, st>>>(x, N, G, sorted, sidx, cell_start, hdr, nullptr, nullptr, nullptr, 0);
        LPD_LAUNCH_CHECK();
    }
    dim3 grid(ceil_div(N, GRID_Q_THREADS), B);
    const int gs = (k + 3) / 4;
    // lock-step search first; the thread-per-query kernel then only finishes the queries on its rescue list
    int* rlist = nullptr;
    int rlimit = 0;
    if (g_grid_lockstep && k <= 20) {                      // (the list of the lock-step kernel holds 32 entries: k <= 20)
        rlist = reinterpret_cast<int*>(reinterpret_cast<uint8_t*>(hdr) + (((size_t)B * sizeof(GridHeader) + 15) & ~(size_t)15));
        LPD_CUDA_CHECK(cudaMemsetAsync(rlist, 0, sizeof(int), st));
        dim3 lgrid(ceil_div(N, GL_WARPS * 32), B);
        knn_grid_lockstep_kernel<<<lgrid, GL_WARPS * 32, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist);
        LPD_LAUNCH_CHECK();
        rlimit = (int)(((long long)B * N) / 8);
        knn_xyz_rescue_warp_kernel<<<148 * 16, RW_WARPS * 32, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit);
        LPD_LAUNCH_CHECK();
        const long long all = ((long long)B * N + GRID_Q_THREADS - 1) / GRID_Q_THREADS;
        grid = dim3((unsigned)(all < 592 ? all : 592), 1);   // a fixed grid walks the list (it returns at once when the list is short)
    }
    switch (gs) {
        case 1: knn_grid_search_kernel<1><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        case 2: knn_grid_search_kernel<2><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        case 3: knn_grid_search_kernel<3><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        case 4: knn_grid_search_kernel<4><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        case 5: knn_grid_search_kernel<5><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        case 6: knn_grid_search_kernel<6><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        case 7: knn_grid_search_kernel<7><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
        default: knn_grid_search_kernel<8><<<grid, GRID_Q_THREADS, 0, st>>>(sorted, sidx, cell_start, hdr, N, k, idx, idx_i64, rlist, rlimit); break;
    }
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_knn_xyz(const float* x, int B, int N, int k, void* idx, int idx_i64, void* workspace, size_t workspace_bytes,
                           void* stream) {
    LPD_REQUIRE(x);
    return knn_xyz_impl(x, B, N, k, idx, idx_i64, workspace, workspace_bytes, stream);
}

// kNN of the cloud lpd_cell_order_grid wrote to xyz_sorted, on the grid that call left in `workspace`: a cloud in cell order is
// its own counting sort (same bounding box, same cells, stable order), so the second build is skipped.
extern "C" int lpd_knn_xyz_ordered(int B, int N, int k, void* idx, int idx_i64, void* workspace, size_t workspace_bytes, void* stream) {
    return knn_xyz_impl(nullptr, B, N, k, idx, idx_i64, workspace, workspace_bytes, stream);
}

// Spatial (grid-cell) order of every cloud: perm[b][t] = original index of the t-th point in cell order (stable inside a
// cell), inv = its inverse, xyz_sorted = the coordinates in that order.  The hot path is permutation-equivariant per point
// and NetVLAD sums over the points, so the host modules may run a cloud in this order: neighbours in space become
// neighbours in memory (gather locality) and the kNN candidate lists converge after the first few tiles.
static int cell_order_impl(const float* x, int B, int N, int32_t* perm, int32_t* inv, float* xyz_sorted,
                           void* workspace, size_t workspace_bytes, void* stream, int identity) {
    LPD_REQUIRE(x && perm && workspace && B >= 1 && B <= 65535 && N >= 1);
    LPD_REQUIRE(((uintptr_t)workspace & 15) == 0);
    if (workspace_bytes < lpd_knn_xyz_workspace_bytes(B, N)) return LPD_EWORKSPACE;
    const size_t cells1 = GRID_MAX_G * GRID_MAX_G * GRID_MAX_G + 1;
    float4* sorted = reinterpret_cast<float4*>(workspace);
    int* sidx = reinterpret_cast<int*>(sorted + (size_t)B * N);
    int* cell_start = sidx + (size_t)B * N;
    GridHeader* hdr = reinterpret_cast<GridHeader*>(cell_start + (size_t)B * cells1);
    hdr = reinterpret_cast<GridHeader*>((reinterpret_cast<uintptr_t>(hdr) + 15) & ~(uintptr_t)15);
    knn_grid_build_kernel<<<B, GRID_BUILD_THREADS, 0, as_stream(stream)>>>(x, N, grid_cells_per_axis(N), sorted, sidx, cell_start, hdr,
                                                                        perm, inv, xyz_sorted, identity);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_cell_order(const float* x, int B, int N, int32_t* perm, int32_t* inv, float* xyz_sorted,
                              void* workspace, size_t workspace_bytes, void* stream) {
    return cell_order_impl(x, B, N, perm, inv, xyz_sorted, workspace, workspace_bytes, stream, 0);
}

extern "C" int lpd_cell_order_grid(const float* x, int B, int N, int32_t* perm, int32_t* inv, float* xyz_sorted,
                                   void* workspace, size_t workspace_bytes, void* stream) {
    LPD_REQUIRE(xyz_sorted);
    return cell_order_impl(x, B, N, perm, inv, xyz_sorted, workspace, workspace_bytes, stream, 1);
}
