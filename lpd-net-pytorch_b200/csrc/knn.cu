// Fused pairwise distance + top-k (kNN).  Replaces knn(), reference util/lpdnet_model.py:317-326.
//
// One CTA owns 64 query points of one cloud and streams the cloud's N candidate points through
// shared memory in tiles of 128.  A 64x128 tile of canonical distances is produced by a
// register-tiled FFMA loop (4x8 per thread), parked in shared memory, and consumed by a
// warp-resident selection network: every warp keeps the running top-k of 8 query rows as a
// sorted list spread over its lanes (lane l = l-th best), filters the tile against the current
// k-th value with one compare per element, and inserts survivors with shuffles.  The N x N
// distance matrix (64 MiB per cloud in the reference) never exists outside the SM.
//
// Canonical arithmetic (the CPU oracle oracle/knn_canonical.c is the same code):
//   dot_ij = fmaf(x_i[C-1], x_j[C-1], ... fmaf(x_i[0], x_j[0], +0))     ascending c
//   xx_j   = the same chain on (x_j, x_j)
//   pd_ij  = ((-xx_j) - (-2 * dot_ij)) - xx_i
//   order  = pd descending, then j ascending.
// Zero padding of the channel axis leaves the chain bit-identical (fmaf(0,0,a) == a).
#include "common.cuh"
#include <limits.h>

namespace lpd {

constexpr int KNN_QT = 64;        // query rows per CTA
constexpr int KNN_CT = 128;       // candidate columns per tile
constexpr int KNN_THREADS = 256;  // 8 warps
constexpr int KNN_RPW = KNN_QT / (KNN_THREADS / 32);  // rows per warp = 8

// global point-major [rows][C]  ->  shared channel-major dst[CP][ROWS], zero padded
template <int CP, int ROWS>
__device__ __forceinline__ void knn_load_tile(const float* __restrict__ xb, int row0, int N, int C,
                                              float* __restrict__ dst, int tid) {
    if ((C & 3) == 0) {
        constexpr int G = CP / 4;
        for (int e = tid; e < ROWS * G; e += KNN_THREADS) {
            const int p = e % ROWS, g = e / ROWS;
            const int row = row0 + p;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (row < N && 4 * g < C) v = __ldg(reinterpret_cast<const float4*>(xb + (size_t)row * C + 4 * g));
            dst[(4 * g + 0) * ROWS + p] = v.x;
            dst[(4 * g + 1) * ROWS + p] = v.y;
            dst[(4 * g + 2) * ROWS + p] = v.z;
            dst[(4 * g + 3) * ROWS + p] = v.w;
        }
    } else {
        for (int e = tid; e < ROWS * CP; e += KNN_THREADS) {
            const int p = e % ROWS, c = e / ROWS;
            const int row = row0 + p;
            float v = 0.f;
            if (row < N && c < C) v = __ldg(xb + (size_t)row * C + c);
            dst[c * ROWS + p] = v;
        }
    }
}

template <int CP, int ROWS>
__device__ __forceinline__ float knn_sqnorm(const float* __restrict__ t, int p) {
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < CP; ++c) {
        const float v = t[c * ROWS + p];
        acc = __fmaf_rn(v, v, acc);
    }
    return acc;
}

template <int CP>
__device__ __forceinline__ void knn_tile(const float* __restrict__ x, int N, int C, int k, void* __restrict__ idx_out, int idx_i64,
                                         const int b, const int tile) {
    extern __shared__ __align__(16) float smem[];
    float* Qs = smem;                      // [CP][QT]
    float* Cs = Qs + CP * KNN_QT;          // [CP][CT]
    float* Ds = Cs + CP * KNN_CT;          // [QT][CT]
    float* xxq = Ds + KNN_QT * KNN_CT;     // [QT]
    float* xxc = xxq + KNN_QT;             // [CT]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 thread grid: 4 rows x 8 cols each
    const int q0 = tile * KNN_QT;
    const float* xb = x + (size_t)b * N * C;

    knn_load_tile<CP, KNN_QT>(xb, q0, N, C, Qs, tid);

    float lv[KNN_RPW];
    int li[KNN_RPW];
#pragma unroll
    for (int r = 0; r < KNN_RPW; ++r) { lv[r] = -INFINITY; li[r] = INT_MAX; }

    for (int c0 = 0; c0 < N; c0 += KNN_CT) {
        knn_load_tile<CP, KNN_CT>(xb, c0, N, C, Cs, tid);
        __syncthreads();  // S1: tile visible; previous selection finished before Ds is rewritten

        if (tid < KNN_CT) xxc[tid] = knn_sqnorm<CP, KNN_CT>(Cs, tid);
        else if (c0 == 0 && tid < KNN_CT + KNN_QT) xxq[tid - KNN_CT] = knn_sqnorm<CP, KNN_QT>(Qs, tid - KNN_CT);

        float acc[4][8];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

#pragma unroll
        for (int c = 0; c < CP; ++c) {
            const float4 a = *reinterpret_cast<const float4*>(Qs + c * KNN_QT + ty * 4);
            const float4 b0 = *reinterpret_cast<const float4*>(Cs + c * KNN_CT + tx * 4);
            const float4 b1 = *reinterpret_cast<const float4*>(Cs + c * KNN_CT + 64 + tx * 4);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = __fmaf_rn(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();  // S2: norms visible, everyone done reading Cs

#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int row = ty * 4 + i;
            const float xi = xxq[row];
            float pd[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int col = (j < 4) ? (tx * 4 + j) : (64 + tx * 4 + (j - 4));
                const float t = -2.0f * acc[i][j];
                const float u = __fsub_rn(-xxc[col], t);
                pd[j] = __fsub_rn(u, xi);
            }
            *reinterpret_cast<float4*>(Ds + row * KNN_CT + tx * 4) = make_float4(pd[0], pd[1], pd[2], pd[3]);
            *reinterpret_cast<float4*>(Ds + row * KNN_CT + 64 + tx * 4) = make_float4(pd[4], pd[5], pd[6], pd[7]);
        }
        __syncthreads();  // S3: distance tile complete

        // ---- selection: warp w owns rows w*8 .. w*8+7 ----
#pragma unroll
        for (int r = 0; r < KNN_RPW; ++r) {
            const int row = warp * KNN_RPW + r;
            float tv = __shfl_sync(kFull, lv[r], k - 1);
            int ti = __shfl_sync(kFull, li[r], k - 1);
#pragma unroll
            for (int t = 0; t < KNN_CT / 32; ++t) {
                const float v = Ds[row * KNN_CT + t * 32 + lane];
                const int j = c0 + t * 32 + lane;
                const bool pass = (j < N) && (v > tv || (v == tv && j < ti));
                unsigned mask = __ballot_sync(kFull, pass);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float cv = __shfl_sync(kFull, v, src);
                    const int cj = __shfl_sync(kFull, j, src);
                    const bool better = (lv[r] > cv) || (lv[r] == cv && li[r] < cj);
                    const int pos = __popc(__ballot_sync(kFull, better));
                    if (pos < k) {  // warp-uniform
                        const float upv = __shfl_up_sync(kFull, lv[r], 1);
                        const int upi = __shfl_up_sync(kFull, li[r], 1);
                        if (lane < k) {
                            if (lane > pos) { lv[r] = upv; li[r] = upi; }
                            else if (lane == pos) { lv[r] = cv; li[r] = cj; }
                        }
                        tv = __shfl_sync(kFull, lv[r], k - 1);
                        ti = __shfl_sync(kFull, li[r], k - 1);
                    }
                }
            }
        }
    }

#pragma unroll
    for (int r = 0; r < KNN_RPW; ++r) {
        const int row = q0 + warp * KNN_RPW + r;
        if (row < N && lane < k) {
            const size_t o = ((size_t)b * N + row) * k + lane;
            if (idx_i64) reinterpret_cast<long long*>(idx_out)[o] = li[r];
            else reinterpret_cast<int*>(idx_out)[o] = li[r];
        }
    }
}

template <int CP>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_kernel(const float* __restrict__ x, int N, int C, int k, void* __restrict__ idx_out, int idx_i64,
           const int* __restrict__ tile_flags) {
    // fallback mode of the tensor-core path: only 64-row tiles flagged as "needs exact recompute" do any work
    if (tile_flags != nullptr && tile_flags[blockIdx.y * gridDim.x + blockIdx.x] == 0) return;
    knn_tile<CP>(x, N, C, k, idx_out, idx_i64, blockIdx.y, blockIdx.x);
}

// work-list mode: list[0] = number of flagged tiles, list[1 + i] = cloud * tiles_per_cloud + tile.  A small fixed grid walks the
// list (the usual case is an empty one: a launch of one CTA per tile cost 11 us just to find that out).
template <int CP>
__global__ void __launch_bounds__(KNN_THREADS, 2)
knn_list_kernel(const float* __restrict__ x, int N, int C, int k, void* __restrict__ idx_out, int idx_i64,
                const int* __restrict__ list, int tiles_per_cloud) {
    const int count = list[0];
    for (int i = blockIdx.x; i < count; i += gridDim.x) {
        const int t = list[1 + i];
        knn_tile<CP>(x, N, C, k, idx_out, idx_i64, t / tiles_per_cloud, t % tiles_per_cloud);
        __syncthreads();                   // the shared-memory tiles are rewritten by the next entry
    }
}

template <int CP>
static int knn_launch(const float* x, int B, int N, int C, int k, void* idx, int idx_i64, cudaStream_t st,
                      const int* tile_flags = nullptr) {
    const size_t smem = (size_t)(CP * (KNN_QT + KNN_CT) + KNN_QT * KNN_CT + KNN_QT + KNN_CT) * sizeof(float);
    LPD_CUDA_CHECK(allow_smem(knn_kernel<CP>, smem));
    dim3 grid(ceil_div(N, KNN_QT), B);
    knn_kernel<CP><<<grid, KNN_THREADS, smem, st>>>(x, N, C, k, idx, idx_i64, tile_flags);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

// ------------------------------------------------------------------------------------------------
// xyz kNN (C <= 3): no distance GEMM worth tiling, so the kernel is organised around the selection.
// The cloud is packed once per CTA into shared memory as float4 (x, y, z, xx).  Every warp owns 4 query
// rows; each lane evaluates one candidate per step against the 4 rows (6 FP32 ops per pair, canonical
// order), one compare + ballot per row filters against the row's current k-th value, and the rare
// survivors are inserted into the warp-resident sorted list with shuffles.  No block-level barrier
// inside the scan.
// ------------------------------------------------------------------------------------------------
constexpr int KNN3_WARPS = 8;
constexpr int KNN3_RPW = 4;                       // query rows per warp
constexpr int KNN3_ROWS = KNN3_WARPS * KNN3_RPW;  // 32 query rows per CTA
constexpr int KNN3_CHUNK = 4096;                  // candidates resident in shared memory at a time (64 KB)

__global__ void __launch_bounds__(KNN3_WARPS * 32)
knn3_kernel(const float* __restrict__ x, int N, int C, int k, void* __restrict__ idx_out, int idx_i64) {
    extern __shared__ __align__(16) float4 cand[];   // [min(N, KNN3_CHUNK)]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int b = blockIdx.y;
    const float* xb = x + (size_t)b * N * C;
    const int row0 = blockIdx.x * KNN3_ROWS + warp * KNN3_RPW;

    float qx[KNN3_RPW], qy[KNN3_RPW], qz[KNN3_RPW], qq[KNN3_RPW];
    float lv[KNN3_RPW], tv[KNN3_RPW];
    int li[KNN3_RPW];
#pragma unroll
    for (int r = 0; r < KNN3_RPW; ++r) {
        const int row = min(row0 + r, N - 1);
        const float* p = xb + (size_t)row * C;
        qx[r] = __ldg(p);
        qy[r] = C > 1 ? __ldg(p + 1) : 0.f;
        qz[r] = C > 2 ? __ldg(p + 2) : 0.f;
        qq[r] = __fmaf_rn(qz[r], qz[r], __fmaf_rn(qy[r], qy[r], __fmaf_rn(qx[r], qx[r], 0.f)));
        lv[r] = -INFINITY; li[r] = INT_MAX; tv[r] = -INFINITY;
    }

    for (int c0 = 0; c0 < N; c0 += KNN3_CHUNK) {
        const int cn = min(KNN3_CHUNK, N - c0);
        __syncthreads();                              // previous chunk fully scanned
        for (int i = tid; i < cn; i += KNN3_WARPS * 32) {
            const float* p = xb + (size_t)(c0 + i) * C;
            float4 v;
            v.x = __ldg(p);
            v.y = C > 1 ? __ldg(p + 1) : 0.f;
            v.z = C > 2 ? __ldg(p + 2) : 0.f;
            v.w = __fmaf_rn(v.z, v.z, __fmaf_rn(v.y, v.y, __fmaf_rn(v.x, v.x, 0.f)));
            cand[i] = v;
        }
        __syncthreads();
        for (int j0 = 0; j0 < cn; j0 += 32) {
            const int jl = j0 + lane;
            const bool valid = jl < cn;
            const float4 cv = cand[valid ? jl : 0];
            const int j = c0 + jl;
#pragma unroll
            for (int r = 0; r < KNN3_RPW; ++r) {
                const float dot = __fmaf_rn(qz[r], cv.z, __fmaf_rn(qy[r], cv.y, __fmaf_rn(qx[r], cv.x, 0.f)));
                const float t = -2.0f * dot;
                const float u = __fsub_rn(-cv.w, t);
                const float pd = __fsub_rn(u, qq[r]);
                unsigned mask = __ballot_sync(kFull, valid && pd >= tv[r]);   // cheap filter; exact order below
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const float cvv = __shfl_sync(kFull, pd, src);
                    const int cj = __shfl_sync(kFull, j, src);
                    const bool better = (lv[r] > cvv) || (lv[r] == cvv && li[r] < cj);
                    const int pos = __popc(__ballot_sync(kFull, better));
                    if (pos < k) {
                        const float upv = __shfl_up_sync(kFull, lv[r], 1);
                        const int upi = __shfl_up_sync(kFull, li[r], 1);
                        if (lane < k) {
                            if (lane > pos) { lv[r] = upv; li[r] = upi; }
                            else if (lane == pos) { lv[r] = cvv; li[r] = cj; }
                        }
                        tv[r] = __shfl_sync(kFull, lv[r], k - 1);
                    }
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < KNN3_RPW; ++r) {
        const int row = row0 + r;
        if (row < N && lane < k) {
            const size_t o = ((size_t)b * N + row) * k + lane;
            if (idx_i64) reinterpret_cast<long long*>(idx_out)[o] = li[r];
            else reinterpret_cast<int*>(idx_out)[o] = li[r];
        }
    }
}

int knn_simt64_flagged(const float* x, int B, int N, int k, void* idx, int idx_i64, const int* flags, cudaStream_t st) {
    return knn_launch<64>(x, B, N, 64, k, idx, idx_i64, st, flags);
}

int knn_simt64_list(const float* x, int B, int N, int k, void* idx, int idx_i64, const int* list, cudaStream_t st) {
    const size_t smem = (size_t)(64 * (KNN_QT + KNN_CT) + KNN_QT * KNN_CT + KNN_QT + KNN_CT) * sizeof(float);
    LPD_CUDA_CHECK(allow_smem(knn_list_kernel<64>, smem));
    const int tiles = ceil_div(N, KNN_QT);
    const long long all = (long long)B * tiles;
    knn_list_kernel<64><<<(unsigned)(all < 296 ? all : 296), KNN_THREADS, smem, st>>>(x, N, 64, k, idx, idx_i64, list, tiles);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

static int knn3_launch(const float* x, int B, int N, int C, int k, void* idx, int idx_i64, cudaStream_t st) {
    const size_t smem = (size_t)(N < KNN3_CHUNK ? N : KNN3_CHUNK) * sizeof(float4);
    LPD_CUDA_CHECK(allow_smem(knn3_kernel, smem > 48 * 1024 ? smem : 48 * 1024));
    dim3 grid(ceil_div(N, KNN3_ROWS), B);
    knn3_kernel<<<grid, KNN3_WARPS * 32, smem, st>>>(x, N, C, k, idx, idx_i64);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

}  // namespace lpd

extern "C" int lpd_knn(const float* x, int B, int N, int C, int k, void* idx, int idx_i64, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(x != nullptr && idx != nullptr);
    LPD_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && C >= 1 && C <= 64);
    LPD_REQUIRE(k >= 1 && k <= 32 && k <= N);
    cudaStream_t st = as_stream(stream);
    if (C <= 3) return knn3_launch(x, B, N, C, k, idx, idx_i64, st);
    if (C <= 4) return knn_launch<4>(x, B, N, C, k, idx, idx_i64, st);
    if (C <= 8) return knn_launch<8>(x, B, N, C, k, idx, idx_i64, st);
    if (C <= 16) return knn_launch<16>(x, B, N, C, k, idx, idx_i64, st);
    if (C <= 32) return knn_launch<32>(x, B, N, C, k, idx, idx_i64, st);
    return knn_launch<64>(x, B, N, C, k, idx, idx_i64, st);
}
