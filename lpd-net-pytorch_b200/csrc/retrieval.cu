// Retrieval: exact brute-force k-NN of descriptor queries against a descriptor database.
// Replaces KDTree(database).query(q, k=25), reference evaluate.py:168,186-187 (and the KDTree in
// util/data.py:111-112).  sklearn's KDTree promotes float32 input to float64 and ranks by
// sum_d (q_d - x_d)^2; this kernel evaluates exactly that in fp64 (d ascending, fma accumulate),
// so ranks agree except on exact fp64 ties, which are broken towards the lower database index.
//
// Grid = (query tiles of 64) x (database splits).  Each CTA streams its database slice through
// shared memory in 64-row tiles (D in chunks of 32, converted to double once per load), produces a
// 64x64 tile of squared distances with a 4x4 register tile per thread, and feeds the same
// warp-resident sorted-list selection as the kNN kernel.  Splits are merged by a second tiny kernel.
#include "common.cuh"
#include <limits.h>

namespace lpd {

constexpr int RT_Q = 64, RT_C = 64, RT_D = 32, RT_THREADS = 256, RT_RPW = 8;

// insert candidate (cv, cj) into the warp-distributed ascending list (lane l = l-th smallest)
__device__ __forceinline__ void list_insert_asc(double& lv, int& li, double cv, int cj, int k, int lane) {
    const bool better = (lv < cv) || (lv == cv && li < cj);
    const int pos = __popc(__ballot_sync(kFull, better));
    if (pos < k) {
        const double upv = __shfl_up_sync(kFull, lv, 1);
        const int upi = __shfl_up_sync(kFull, li, 1);
        if (lane < k) {
            if (lane > pos) { lv = upv; li = upi; }
            else if (lane == pos) { lv = cv; li = cj; }
        }
    }
}

template <int ROWS>
__device__ __forceinline__ void rt_load(const float* __restrict__ base, int row0, int nrows, int d0, int D,
                                        double* __restrict__ dst, int tid) {
    const bool vec = (D & 3) == 0;
    for (int e = tid; e < ROWS * (RT_D / 4); e += RT_THREADS) {
        const int p = e % ROWS, g = e / ROWS;
        const int row = row0 + p, d = d0 + 4 * g;
        float v[4] = {0.f, 0.f, 0.f, 0.f};
        if (row < nrows) {
            const float* src = base + (size_t)row * D + d;
            if (vec && d + 3 < D) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(src));
                v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
            } else {
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    if (d + u < D) v[u] = __ldg(src + u);
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) dst[(4 * g + u) * ROWS + p] = (double)v[u];
    }
}

__global__ void __launch_bounds__(RT_THREADS, 2)
retrieval_kernel(const float* __restrict__ db, int Ndb, const float* __restrict__ q, int Nq, int D, int k,
                 int rows_per_split, double* __restrict__ part_val, int* __restrict__ part_idx) {
    extern __shared__ __align__(16) double smd[];
    double* Qs = smd;                   // [RT_D][RT_Q]
    double* Cs = Qs + RT_D * RT_Q;      // [RT_D][RT_C]
    double* Ds = Cs + RT_D * RT_C;      // [RT_Q][RT_C]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int tx = tid & 15, ty = tid >> 4;
    const int q0 = blockIdx.x * RT_Q;
    const int split = blockIdx.y;
    const int r_begin = split * rows_per_split;
    const int r_end = min(Ndb, r_begin + rows_per_split);

    double lv[RT_RPW];
    int li[RT_RPW];
#pragma unroll
    for (int r = 0; r < RT_RPW; ++r) { lv[r] = INFINITY; li[r] = INT_MAX; }

    for (int c0 = r_begin; c0 < r_end; c0 += RT_C) {
        double acc[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = 0.0;
        for (int d0 = 0; d0 < D; d0 += RT_D) {
            __syncthreads();  // previous chunk consumed / previous selection done
            rt_load<RT_Q>(q, q0, Nq, d0, D, Qs, tid);
            rt_load<RT_C>(db, c0, r_end, d0, D, Cs, tid);
            __syncthreads();
#pragma unroll 8
            for (int dd = 0; dd < RT_D; ++dd) {
                double a[4], b[4];
#pragma unroll
                for (int i = 0; i < 4; ++i) a[i] = Qs[dd * RT_Q + ty * 4 + i];
#pragma unroll
                for (int j = 0; j < 4; ++j) b[j] = Cs[dd * RT_C + tx * 4 + j];
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const double t = a[i] - b[j];
                        acc[i][j] = fma(t, t, acc[i][j]);
                    }
            }
        }
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) Ds[(ty * 4 + i) * RT_C + tx * 4 + j] = acc[i][j];
        __syncthreads();
#pragma unroll
        for (int r = 0; r < RT_RPW; ++r) {
            const int row = warp * RT_RPW + r;
            double tv = __shfl_sync(kFull, lv[r], k - 1);
            int ti = __shfl_sync(kFull, li[r], k - 1);
#pragma unroll
            for (int t = 0; t < RT_C / 32; ++t) {
                const double v = Ds[row * RT_C + t * 32 + lane];
                const int j = c0 + t * 32 + lane;
                const bool pass = (j < r_end) && (v < tv || (v == tv && j < ti));
                unsigned mask = __ballot_sync(kFull, pass);
                while (mask) {
                    const int src = __ffs(mask) - 1;
                    mask &= mask - 1;
                    const double cv = __shfl_sync(kFull, v, src);
                    const int cj = __shfl_sync(kFull, j, src);
                    list_insert_asc(lv[r], li[r], cv, cj, k, lane);
                    tv = __shfl_sync(kFull, lv[r], k - 1);
                    ti = __shfl_sync(kFull, li[r], k - 1);
                }
            }
        }
    }
#pragma unroll
    for (int r = 0; r < RT_RPW; ++r) {
        const int row = q0 + warp * RT_RPW + r;
        if (row < Nq && lane < k) {
            const size_t o = ((size_t)split * Nq + row) * k + lane;
            part_val[o] = lv[r];
            part_idx[o] = li[r];
        }
    }
}

// one warp per query: merge `splits` sorted partial lists, write final idx (+offset) and dist
__global__ void __launch_bounds__(256)
retrieval_merge_kernel(const double* __restrict__ part_val, const int* __restrict__ part_idx, int splits, int Nq, int k,
                       int idx_offset, int* __restrict__ idx, double* __restrict__ dist) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= Nq) return;
    double lv = INFINITY;
    int li = INT_MAX;
    for (int s = 0; s < splits; ++s) {
        const size_t o = ((size_t)s * Nq + row) * k + lane;
        const double v = lane < k ? part_val[o] : INFINITY;
        const int j = lane < k ? part_idx[o] : INT_MAX;
        unsigned mask = __ballot_sync(kFull, lane < k && j != INT_MAX && j >= 0);   // INT_MAX / -1 = empty slot
        while (mask) {
            const int src = __ffs(mask) - 1;
            mask &= mask - 1;
            const double cv = __shfl_sync(kFull, v, src);
            const int cj = __shfl_sync(kFull, j, src);
            list_insert_asc(lv, li, cv, cj, k, lane);
        }
    }
    if (lane < k) {
        idx[(size_t)row * k + lane] = (li == INT_MAX) ? -1 : li + idx_offset;
        if (dist) dist[(size_t)row * k + lane] = lv;
    }
}

static int rt_splits(int Ndb, int Nq) {
    int dev = 0, sms = 148;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const int qtiles = ceil_div(Nq, RT_Q);
    const int max_splits = ceil_div(Ndb, RT_C);
    int s = ceil_div(2 * sms, qtiles);
    if (s > max_splits) s = max_splits;
    if (s < 1) s = 1;
    if (s > 1024) s = 1024;
    return s;
}

}  // namespace lpd

extern "C" size_t lpd_retrieval_workspace_bytes(int Ndb, int Nq, int k) {
    using namespace lpd;
    if (Ndb < 1 || Nq < 1 || k < 1) return 0;
    const size_t n = (size_t)rt_splits(Ndb, Nq) * Nq * k;
    return n * (sizeof(double) + sizeof(int)) + 256;
}

extern "C" int lpd_retrieval_topk(const float* db, int Ndb, const float* q, int Nq, int D, int k,
                                  int idx_offset, int32_t* idx, double* dist,
                                  void* workspace, size_t workspace_bytes, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(db && q && idx && workspace);
    LPD_REQUIRE(Ndb >= 1 && Nq >= 1 && D >= 1 && k >= 1 && k <= 32);
    LPD_REQUIRE(((uintptr_t)db & 15) == 0 && ((uintptr_t)q & 15) == 0 && ((uintptr_t)workspace & 7) == 0);
    if (workspace_bytes < lpd_retrieval_workspace_bytes(Ndb, Nq, k)) return LPD_EWORKSPACE;
    const int splits = rt_splits(Ndb, Nq);
    int rows_per_split = ceil_div(Ndb, splits);
    rows_per_split = ceil_div(rows_per_split, RT_C) * RT_C;
    const int used_splits = ceil_div(Ndb, rows_per_split);
    double* part_val = reinterpret_cast<double*>(workspace);
    int* part_idx = reinterpret_cast<int*>(part_val + (size_t)splits * Nq * k);
    const size_t smem = (size_t)(RT_D * RT_Q + RT_D * RT_C + RT_Q * RT_C) * sizeof(double);
    LPD_CUDA_CHECK(allow_smem(retrieval_kernel, smem));
    cudaStream_t st = as_stream(stream);
    LPD_REQUIRE(used_splits <= 65535);
    retrieval_kernel<<<dim3(ceil_div(Nq, RT_Q), used_splits), RT_THREADS, smem, st>>>(db, Ndb, q, Nq, D, k, rows_per_split,
                                                                                   part_val, part_idx);
    LPD_LAUNCH_CHECK();
    retrieval_merge_kernel<<<ceil_div(Nq, 8), 256, 0, st>>>(part_val, part_idx, used_splits, Nq, k, idx_offset, idx, dist);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_topk_merge(const double* part_dist, const int32_t* part_idx, int lists, int Nq, int k,
                              int32_t* idx, double* dist, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(part_dist && part_idx && idx && lists >= 1 && Nq >= 1 && k >= 1 && k <= 32);
    retrieval_merge_kernel<<<ceil_div(Nq, 8), 256, 0, as_stream(stream)>>>(part_dist, part_idx, lists, Nq, k, 0, idx, dist);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
