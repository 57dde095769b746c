// Retrieval on the tensor cores: exact k nearest database descriptors of every query, per database SEGMENT (one segment =
// the database of one run, reference evaluate.py:59-70,162-206: KDTree(DATABASE_VECTORS[m]).query(q, k=25) for every
// ordered run pair), and the recall bookkeeping of get_recall (:176-206) on the device.
//
// Filter and refine, same argument as the feature-space kNN:
//   1. split      every fp32 coordinate x = hi + lo with hi = x with the low 13 mantissa bits cleared (exactly representable
//                 in TF32) and lo = x - hi (exact in fp32).  Queries are stored as [hi | hi | lo], database rows as
//                 [hi | lo | hi] (K = 3 D), so ONE kind::tf32 GEMM accumulates hi.hi + hi.lo + lo.hi in fp32.
//   2. GEMM       lpd_gemm_tf32 (TMA + tcgen05 + TMEM, gemm_tc.cu) with the epilogue  A_ij = -2 * dot_ij + |d_j|^2 :
//                 approximate squared distances minus the per-query constant |q_i|^2, fp32, [Nq][Ndb] in HBM.
//   3. select     one warp per (query, segment): the k-th smallest approximate score T of the segment (bisection on the
//                 order-preserving integer image of the floats), then every row with A_ij <= T + 2 eps_i is a candidate:
//                 |A - E| <= eps for the exact E, so k rows have E <= T + eps, the exact k-th best is <= T + eps, and every
//                 member of the exact top-k (ties included) has A <= T + 2 eps.
//   4. refine     candidates are re-scored exactly as lpd_retrieval_topk does — sum_d ((double)q_d - (double)x_d)^2, d ascending,
//                 fma — and ranked by (distance, index): the result is bit-identical to the brute-force fp64 kernel.
//
// eps_i: dropped lo.lo terms (2^-20), TF32 truncation of the lo operands (2 * 2^-20), fp32 accumulation of 3 D <= 3072
// products (<= 3072 * 2^-24 < 2^-12 worst case, all relative to sum |q_d||x_d| <= |q||x|), the fp32 rounding of |d_j|^2 and of
// the epilogue fma (2^-22 (|x|^2 + 2|q||x|)):  eps_i = 2 * 2^-11 |q_i| xmax + 2^-20 (xmax^2 + 2 |q_i| xmax), twice the sum of
// those terms.  For unit descriptors that is 1e-3 of the typical neighbour distance gap: a handful of extra candidates.
#include "common.cuh"
#include <limits.h>

extern "C" int lpd_gemm_tf32_ex(const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                                int M, int N, int K, int batch, int accumulate, const float* scale, const float* shift, int act,
                                float slope, void* stream);

namespace lpd {
namespace rtc {

constexpr int CHUNK = 1024;     // scores per warp pass: 32 per lane

__device__ __forceinline__ unsigned fkey(float v) {            // order-preserving float -> uint
    const unsigned b = __float_as_uint(v);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float fkey_inv(unsigned k) {
    return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// rows [n][D] -> out [n_pad][3 D] in the role's column order, norm2[n] = fp32(sum x^2 in fp64), xmax = max |x| (atomic on
// the bit pattern of a non-negative float), fill[n_pad] = -2 (the GEMM's per-column scale), rows n..n_pad zeroed
__global__ void __launch_bounds__(256)
split_kernel(const float* __restrict__ x, int n, int n_pad, int D, int is_db, float* __restrict__ out,
             float* __restrict__ norm2, float* __restrict__ fill, unsigned* __restrict__ xmax_bits) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= n_pad) return;
    float* o = out + (size_t)row * 3 * D;
    if (row >= n) {
        for (int d = lane; d < 3 * D; d += 32) o[d] = 0.f;
        if (lane == 0) { if (norm2) norm2[row] = 0.f; if (fill) fill[row] = -2.f; }
        return;
    }
    double s = 0.0;
    for (int d = lane; d < D; d += 32) {
        const float v = __ldg(x + (size_t)row * D + d);
        const float hi = __uint_as_float(__float_as_uint(v) & 0xffffe000u);
        const float lo = v - hi;
        o[d] = hi;
        o[D + d] = is_db ? lo : hi;
        o[2 * D + d] = is_db ? hi : lo;
        s = fma((double)v, (double)v, s);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(kFull, s, off);
    if (lane == 0) {
        if (norm2) norm2[row] = (float)s;
        if (fill) fill[row] = -2.f;
        atomicMax(xmax_bits, __float_as_uint((float)sqrt(s) * 1.0000002f));
    }
}

// one warp per (query, segment)
__global__ void __launch_bounds__(256)
select_refine_kernel(const float* __restrict__ A, int lda, const float* __restrict__ db, const float* __restrict__ q, int Nq, int D,
                     int k, const int* __restrict__ seg_off, int S, const float* __restrict__ qnorm2,
                     const unsigned* __restrict__ dmax_bits, int global_idx, int idx_offset,
                     int* __restrict__ idx, double* __restrict__ dist) {
    extern __shared__ __align__(16) float qsm[];                  // [8 warps][D]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * 8 + warp;
    if (unit >= (long long)Nq * S) return;
    const int s = (int)(unit / Nq), i = (int)(unit % Nq);         // consecutive warps: consecutive queries of one segment
    const int r0 = seg_off[s], r1 = seg_off[s + 1];
    double* qs = reinterpret_cast<double*>(qsm) + warp * D;       // the query row, converted once
    int* cjs = reinterpret_cast<int*>(qsm + 16 * D) + warp * CHUNK;                   // compacted candidate rows of a chunk
    double* mvs = reinterpret_cast<double*>(reinterpret_cast<int*>(qsm + 16 * D) + 8 * CHUNK) + warp * 32;   // merge buffer
    int* mis = reinterpret_cast<int*>(reinterpret_cast<double*>(reinterpret_cast<int*>(qsm + 16 * D) + 8 * CHUNK) + 8 * 32) + warp * 32;
    for (int d = lane; d < D; d += 32) qs[d] = (double)__ldg(q + (size_t)i * D + d);
    __syncwarp();
    const float qn = sqrtf(qnorm2[i]) * 1.0000002f, xm = __uint_as_float(*dmax_bits);
    const float eps = 2.f * (qn * xm * 4.8828125e-4f) + 9.5367431640625e-7f * (xm * xm + 2.f * qn * xm);
    const float* Arow = A + (size_t)i * lda;

    double lv = INFINITY;
    int li = INT_MAX;
    for (int c0 = r0; c0 < r1; c0 += CHUNK) {
        const int len = min(CHUNK, r1 - c0);
        float v[CHUNK / 32];
        unsigned key[CHUNK / 32];
#pragma unroll
        for (int t = 0; t < CHUNK / 32; ++t) {
            const int j = t * 32 + lane;
            v[t] = j < len ? __ldg(Arow + c0 + j) : INFINITY;
            key[t] = fkey(v[t]);
        }
        const int kk = min(k, len);
        // a threshold T with count(key <= T) >= kk, as tight as cheap: (1) the kk-th smallest of the 32 per-lane minima is an upper
        // bound of the kk-th smallest score (kk distinct scores reach it); (2) bisection on the order-preserving integer image between
        // the smallest score and that bound, stopped as soon as the count lands in [kk, kk + 3] (any such T keeps the exact top-k in
        // the candidate set; three extra candidates cost less than the ~20 further halvings down to a single key)
        unsigned lo, hi;
        {
            unsigned mn = 0xffffffffu;
#pragma unroll
            for (int t = 0; t < CHUNK / 32; ++t) mn = min(mn, key[t]);
            lo = __reduce_min_sync(kFull, mn);
            int rank = 0;
#pragma unroll 8
            for (int src = 0; src < 32; ++src) {
                const unsigned o = __shfl_sync(kFull, mn, src);
                rank += (o < mn) || (o == mn && src < lane);
            }
            const unsigned pick = __ballot_sync(kFull, rank == min(kk, 32) - 1);
            hi = __shfl_sync(kFull, mn, __ffs(pick) - 1);
        }
        while (lo < hi) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            int c = 0;
#pragma unroll
            for (int t = 0; t < CHUNK / 32; ++t) c += key[t] <= mid;
            c = __reduce_add_sync(kFull, c);
            if (c >= kk) { hi = mid; if (c <= kk + 3) break; } else lo = mid + 1;
        }
        lo = hi;
        float T = fkey_inv(lo);
        // later chunks of a long segment: nothing worse than the running k-th exact distance can enter the list
        const double kth = __shfl_sync(kFull, lv, k - 1);
        float thr = T + 2.f * eps;
        if (kth < (double)INFINITY) {
            const float cap = (float)(kth - (double)qnorm2[i]) + 2.f * eps + 1e-6f * fabsf((float)kth);
            thr = fminf(thr, cap);
        }
        // ---- the candidates of this chunk, compacted (they are scattered over the 32 x 32 positions of the chunk: re-scoring them
        //      where they lie ran the 256-step fp64 chain ~20 times per warp with one or two live lanes: 0.75 ms of the 1.0 ms pass) ----
        int ncand = 0;
#pragma unroll
        for (int t = 0; t < CHUNK / 32; ++t) {
            const bool pass = v[t] <= thr;
            const unsigned bal = __ballot_sync(kFull, pass);
            if (pass) cjs[ncand + __popc(bal & ((1u << lane) - 1u))] = c0 + t * 32 + lane;
            ncand += __popc(bal);
        }
        __syncwarp();
        for (int base = 0; base < ncand; base += 32) {
            // exact distance: one candidate per lane, d ascending, fma — the arithmetic of lpd_retrieval_topk
            const int j = base + lane < ncand ? cjs[base + lane] : INT_MAX;
            double acc = INFINITY;
            if (j != INT_MAX) {
                acc = 0.0;
                const float4* row = reinterpret_cast<const float4*>(db + (size_t)j * D);
                if ((D & 3) == 0) {
#pragma unroll 4
                    for (int d4 = 0; d4 < D / 4; ++d4) {
                        const float4 x = __ldg(row + d4);
                        const double2 qa = *reinterpret_cast<const double2*>(qs + 4 * d4), qb = *reinterpret_cast<const double2*>(qs + 4 * d4 + 2);
                        double tt = qa.x - (double)x.x; acc = fma(tt, tt, acc);
                        tt = qa.y - (double)x.y; acc = fma(tt, tt, acc);
                        tt = qb.x - (double)x.z; acc = fma(tt, tt, acc);
                        tt = qb.y - (double)x.w; acc = fma(tt, tt, acc);
                    }
                } else {
                    for (int d = 0; d < D; ++d) {
                        const double tt = qs[d] - (double)__ldg(db + (size_t)j * D + d);
                        acc = fma(tt, tt, acc);
                    }
                }
            }
            // merge the <= 32 new entries into the running ascending list (lane l = l-th smallest; ties -> lower index) by
            // counting: an entry's new position = the number of entries of both sets that come before it
            int rank_new = 0, rank_old = lane;
#pragma unroll 4
            for (int src = 0; src < 32; ++src) {
                const double ov = __shfl_sync(kFull, lv, src), nv = __shfl_sync(kFull, acc, src);
                const int oi = __shfl_sync(kFull, li, src), ni = __shfl_sync(kFull, j, src);
                rank_new += ((ov < acc) || (ov == acc && oi < j)) + ((nv < acc) || (nv == acc && ni < j));
                rank_old += (nv < lv) || (nv == lv && ni < li);
            }
            __syncwarp();
            mvs[lane] = INFINITY; mis[lane] = INT_MAX;
            __syncwarp();
            if (li != INT_MAX && rank_old < 32) { mvs[rank_old] = lv; mis[rank_old] = li; }
            if (j != INT_MAX && rank_new < 32) { mvs[rank_new] = acc; mis[rank_new] = j; }
            __syncwarp();
            lv = lane < k ? mvs[lane] : (double)INFINITY;
            li = lane < k ? mis[lane] : INT_MAX;
        }
        __syncwarp();
    }
    if (lane < k) {
        const size_t o = ((size_t)s * Nq + i) * k + lane;
        idx[o] = (li == INT_MAX) ? -1 : (global_idx ? li + idx_offset : li - r0);
        if (dist) dist[o] = lv;
    }
}

// get_recall's bookkeeping (reference evaluate.py:176-206), one warp per (query, segment):
//   truth list empty -> skipped; first rank j whose index is a true neighbour -> hist[pair][j]++ (and, for j == 0, the
//   dot-product similarity of the pair); any hit among the first thresh[segment] ranks -> one_pct[pair]++.
__global__ void __launch_bounds__(256)
recall_kernel(const int* __restrict__ idx, int S, int Nq, int k, const int* __restrict__ q_run, int R,
              const int* __restrict__ seg_run, const int* __restrict__ truth_off, const int* __restrict__ truth_idx, const int* __restrict__ seg_thresh,
              const float* __restrict__ db, const int* __restrict__ seg_off, const float* __restrict__ q, int D,
              int* __restrict__ hist, int* __restrict__ n_eval, int* __restrict__ n_onepct, float* __restrict__ sim) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const long long unit = (long long)blockIdx.x * 8 + warp;
    if (unit >= (long long)Nq * S) return;
    const int s = (int)(unit / Nq), i = (int)(unit % Nq);
    const int n = q_run[i];
    const int m = seg_run ? seg_run[s] : s;                        // the run this segment is the database of
    if (lane == 0 && sim) sim[(size_t)i * S + s] = __int_as_float(0x7fc00000);
    if (n == m) return;                                            // the reference's pair loop skips m == n (:61-62)
    const int t0 = truth_off[(size_t)i * R + m], t1 = truth_off[(size_t)i * R + m + 1];
    if (t1 == t0) return;                                          // :181-182
    const int pair = (n < 0 ? 0 : n) * S + s;
    const int mine = lane < k ? idx[((size_t)s * Nq + i) * k + lane] : -1;
    bool hit = false;
    for (int t = t0; t < t1; ++t) hit |= (mine >= 0 && mine == __ldg(truth_idx + t));
    const unsigned mask = __ballot_sync(kFull, hit);
    const int first = mask ? __ffs(mask) - 1 : -1;
    const int thresh = seg_thresh[s];
    const unsigned low = thresh >= 32 ? 0xffffffffu : ((1u << thresh) - 1u);
    if (lane == 0) {
        atomicAdd(n_eval + pair, 1);
        if (first >= 0) atomicAdd(hist + (size_t)pair * 25 + first, 1);
        if (mask & low) atomicAdd(n_onepct + pair, 1);
    }
    if (first == 0 && sim) {
        const int j = seg_off[s] + __shfl_sync(kFull, mine, 0);
        double acc = 0.0;
        for (int d = lane; d < D; d += 32) acc = fma((double)__ldg(q + (size_t)i * D + d), (double)__ldg(db + (size_t)j * D + d), acc);
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(kFull, acc, off);
        if (lane == 0) sim[(size_t)i * S + s] = (float)acc;
    }
}

struct Layout {
    size_t qs, dbs, A, dn, fill, qn, xmax, total;
    int ndb_pad;
};
static Layout layout(int Ndb, int Nq, int D) {
    Layout L;
    L.ndb_pad = (Ndb + 3) / 4 * 4;
    size_t o = 0;
    auto take = [&](size_t bytes) { size_t r = o; o += (bytes + 255) / 256 * 256; return r; };
    L.qs = take((size_t)Nq * 3 * D * 4);
    L.dbs = take((size_t)L.ndb_pad * 3 * D * 4);
    L.A = take((size_t)Nq * L.ndb_pad * 4);
    L.dn = take((size_t)L.ndb_pad * 4);
    L.fill = take((size_t)L.ndb_pad * 4);
    L.qn = take((size_t)Nq * 4);
    L.xmax = take(256);
    L.total = o;
    return L;
}

}  // namespace rtc
}  // namespace lpd

extern "C" size_t lpd_retrieval_tc_workspace_bytes(int Ndb, int Nq, int D) {
    if (Ndb < 1 || Nq < 1 || D < 1) return 0;
    return lpd::rtc::layout(Ndb, Nq, D).total;
}

extern "C" int lpd_retrieval_tc(const float* db, int Ndb, const float* q, int Nq, int D, int k,
                                const int32_t* seg_off, int S, int global_idx, int idx_offset,
                                int32_t* idx, double* dist, void* workspace, size_t workspace_bytes, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(db && q && idx && seg_off && workspace);
    LPD_REQUIRE(Ndb >= 1 && Nq >= 1 && S >= 1 && k >= 1 && k <= 32);
    LPD_REQUIRE(D >= 4 && (D % 4) == 0 && D <= 1024);            // 3 D is the GEMM's K; rows are read as float4
    LPD_REQUIRE(((uintptr_t)db & 15) == 0 && ((uintptr_t)q & 15) == 0 && ((uintptr_t)workspace & 255) == 0);
    const rtc::Layout L = rtc::layout(Ndb, Nq, D);
    if (workspace_bytes < L.total) return LPD_EWORKSPACE;
    uint8_t* ws = reinterpret_cast<uint8_t*>(workspace);
    float* qs = reinterpret_cast<float*>(ws + L.qs);
    float* dbs = reinterpret_cast<float*>(ws + L.dbs);
    float* A = reinterpret_cast<float*>(ws + L.A);
    float* dn = reinterpret_cast<float*>(ws + L.dn);
    float* fill = reinterpret_cast<float*>(ws + L.fill);
    float* qn = reinterpret_cast<float*>(ws + L.qn);
    unsigned* xmax = reinterpret_cast<unsigned*>(ws + L.xmax);
    cudaStream_t st = as_stream(stream);
    LPD_CUDA_CHECK(cudaMemsetAsync(xmax, 0, 4, st));
    rtc::split_kernel<<<ceil_div(L.ndb_pad, 8), 256, 0, st>>>(db, Ndb, L.ndb_pad, D, 1, dbs, dn, fill, xmax);
    LPD_LAUNCH_CHECK();
    unsigned* qmax = xmax + 1;                                      // the queries' own maximum is not needed; scratch word
    LPD_CUDA_CHECK(cudaMemsetAsync(qmax, 0, 4, st));
    rtc::split_kernel<<<ceil_div(Nq, 8), 256, 0, st>>>(q, Nq, Nq, D, 0, qs, qn, nullptr, qmax);
    LPD_LAUNCH_CHECK();
    int rc = lpd_gemm_tf32_ex(qs, 3 * D, dbs, 3 * D, A, L.ndb_pad, Nq, L.ndb_pad, 3 * D, 1, 0, fill, dn, LPD_ACT_NONE, 0.f, stream);
    if (rc != LPD_OK) return rc;
    const long long units = (long long)Nq * S;
    const size_t smem = (size_t)8 * D * sizeof(double) + (size_t)8 * rtc::CHUNK * sizeof(int) + 8 * 32 * (sizeof(double) + sizeof(int));
    LPD_REQUIRE((D % 2) == 0 && smem <= 200 * 1024);
    LPD_CUDA_CHECK(allow_smem(rtc::select_refine_kernel, smem));
    rtc::select_refine_kernel<<<(unsigned)ceil_div_ll(units, 8), 256, smem, st>>>(A, L.ndb_pad, db, q, Nq, D, k, seg_off, S, qn, xmax,
                                                                                 global_idx, idx_offset, idx, dist);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_recall_count(const int32_t* idx, int S, int Nq, int k, const int32_t* q_run, int R, const int32_t* seg_run,
                                const int32_t* truth_off, const int32_t* truth_idx, const int32_t* seg_thresh,
                                const float* db, const int32_t* seg_off, const float* q, int D,
                                int32_t* hist, int32_t* n_eval, int32_t* n_onepct, float* sim, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(idx && q_run && truth_off && truth_idx && seg_thresh && hist && n_eval && n_onepct);
    LPD_REQUIRE(S >= 1 && Nq >= 1 && k >= 1 && k <= 25 && R >= 1);
    LPD_REQUIRE(!sim || (db && seg_off && q && D >= 1));
    const long long units = (long long)Nq * S;
    rtc::recall_kernel<<<(unsigned)ceil_div_ll(units, 8), 256, 0, as_stream(stream)>>>(
        idx, S, Nq, k, q_run, R, seg_run, truth_off, truth_idx, seg_thresh, db, seg_off, q, D, hist, n_eval, n_onepct, sim);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
