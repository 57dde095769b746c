// Feature-space kNN (C = 64) with the distance GEMM on the tensor cores and an EXACT result.
// Replaces knn() on the 64-channel feature map, reference util/lpdnet_model.py:246 -> :336 -> :317-326.
//
// Filter and refine:
//   1. tcgen05.mma kind::tf32 produces approximate gram tiles  dot~[i][j]  (queries = TMEM lanes, candidates = columns).
//   2. Each epilogue thread owns one query row: a[j] = 2 * dot~ - xx_j is compared against the row's current L-th best
//      approximate score; the rare survivors go to a per-row queue in shared memory and are merged, warp-cooperatively,
//      into the row's sorted list of L = 32 (k <= 24) or 64 candidates.
//   3. After the scan the L candidates of a row are re-scored with the CANONICAL fp32 arithmetic of lpd_knn (fmaf chain,
//      c ascending; pd = ((-xx_j) - (-2 dot)) - xx_i), sorted by (pd descending, index ascending) and the first k are
//      written.  The list provably contains the canonical top-k when  a(L) < a(k) - 2 * eps_i  with
//      eps_i = 2^-7 |x_i| max_j |x_j|  (>= 2x the worst-case TF32 truncation error of 2 * dot); rows that fail this test
//      (masses of near-ties, e.g. duplicated points) are flagged and recomputed by the exact CUDA-core kernel.
// The N x N matrix never leaves TMEM; HBM traffic is the feature map once per 128-query tile (L2 resident).
#include "tc_common.cuh"
#include <limits.h>

namespace lpd {

int knn_simt64_flagged(const float* x, int B, int N, int k, void* idx, int idx_i64, const int* flags, cudaStream_t st);

namespace tc {

constexpr int KT_Q = 128;          // queries per work item (TMEM lanes)
constexpr int KT_C = 128;          // candidates per tile (TMEM columns)
constexpr int KT_THREADS = 192;    // warps 0-3 scan/select, warp 4 TMA, warp 5 MMA
constexpr int KT_QSTRIDE = 33;     // queue row stride in entries (bank spread)

struct KnnTcParams {
    const float* x;        // [B*N][64]
    const float* xxpad;    // [B][Npad] canonical squared norms, +inf padded
    const float* r2;       // [B] max squared norm of the cloud
    int* flags;            // [B][ceil(N/64)] rows needing the exact fallback
    void* idx; int idx_i64;
    int B, N, Npad, k;
    int qtiles, ctiles;    // per cloud
};

__global__ void knn_prep_kernel(const float* __restrict__ x, int N, int Npad, float* __restrict__ xxpad, float* __restrict__ r2) {
    const int b = blockIdx.y;
    const int n = blockIdx.x * blockDim.x + threadIdx.x;
    if (n >= Npad) return;
    float v = INFINITY;
    if (n < N) {
        const float4* p = reinterpret_cast<const float4*>(x + ((size_t)b * N + n) * 64);
        float acc = 0.f;
#pragma unroll
        for (int g = 0; g < 16; ++g) {
            const float4 t = __ldg(p + g);
            acc = __fmaf_rn(t.x, t.x, acc); acc = __fmaf_rn(t.y, t.y, acc);
            acc = __fmaf_rn(t.z, t.z, acc); acc = __fmaf_rn(t.w, t.w, acc);
        }
        v = acc;
        atomicMax(reinterpret_cast<int*>(r2 + b), __float_as_int(acc));   // acc >= 0: int order == float order
    }
    xxpad[(size_t)b * Npad + n] = v;
}

// sorted insert of (cv, cj) into a warp-distributed list: entry l in lane l (L == 32) or entries l and 32 + l (L == 64)
template <int L>
__device__ __forceinline__ void list_insert(float (&lv)[L / 32], int (&li)[L / 32], float cv, int cj, int lane) {
    bool b0 = (lv[0] > cv) || (lv[0] == cv && li[0] < cj);
    int pos = __popc(__ballot_sync(kFull, b0));
    if (L == 64) {
        bool b1 = (lv[L / 32 - 1] > cv) || (lv[L / 32 - 1] == cv && li[L / 32 - 1] < cj);
        pos += __popc(__ballot_sync(kFull, b1));
    }
    if (pos >= L) return;
    if (L == 64) {
        // second half first: its lane 0 receives the element shifted out of the first half
        const float carry_v = __shfl_sync(kFull, lv[0], 31);
        const int carry_i = __shfl_sync(kFull, li[0], 31);
        float upv = __shfl_up_sync(kFull, lv[L / 32 - 1], 1);
        int upi = __shfl_up_sync(kFull, li[L / 32 - 1], 1);
        if (lane == 0) { upv = carry_v; upi = carry_i; }
        const int g = 32 + lane;
        if (g > pos) { lv[L / 32 - 1] = upv; li[L / 32 - 1] = upi; }
        else if (g == pos) { lv[L / 32 - 1] = cv; li[L / 32 - 1] = cj; }
    }
    const float upv = __shfl_up_sync(kFull, lv[0], 1);
    const int upi = __shfl_up_sync(kFull, li[0], 1);
    if (lane > pos) { lv[0] = upv; li[0] = upi; }
    else if (lane == pos) { lv[0] = cv; li[0] = cj; }
}

// bitonic sort of the warp-distributed list by (value descending, index ascending)
template <int L>
__device__ __forceinline__ void list_sort(float (&v)[L / 32], int (&id)[L / 32], int lane) {
    auto before = [](float a, int ia, float b, int ib) { return (a > b) || (a == b && ia < ib); };
    for (int size = 2; size <= L; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            if (stride >= 32) {   // L == 64, partner is the other register of the same lane (global index = e*32 + lane)
                const bool up = true;  // size == 64 only: whole list sorted in one direction
                const bool swap = before(v[1], id[1], v[0], id[0]) == up;
                if (swap) { float t = v[0]; v[0] = v[1]; v[1] = t; int ti = id[0]; id[0] = id[1]; id[1] = ti; }
            } else {
#pragma unroll
                for (int e = 0; e < L / 32; ++e) {
                    const int g = e * 32 + lane;
                    const float pv = __shfl_xor_sync(kFull, v[e], stride);
                    const int pi = __shfl_xor_sync(kFull, id[e], stride);
                    const bool dir_desc = ((g & size) == 0);           // this block sorted "best first"
                    const bool lower = ((g & stride) == 0);            // I keep the better element if lower half
                    const bool mine_better = before(v[e], id[e], pv, pi);
                    const bool keep_mine = (mine_better == (lower == dir_desc));
                    if (!keep_mine) { v[e] = pv; id[e] = pi; }
                }
            }
        }
    }
}

template <int L>
__global__ void __launch_bounds__(KT_THREADS, 1)
knn_tc_kernel(const __grid_constant__ CUtensorMap tmap_x, KnnTcParams P) {
    constexpr uint32_t KB_BYTES = 128 * 128;              // one 32-channel k-block of a 128-row tile
    constexpr uint32_t OP_BYTES = 2 * KB_BYTES;           // 64 channels
    constexpr int E = L / 32;
    extern __shared__ __align__(1024) uint8_t smem[];
    if ((smem_u32(smem) & 1023u) != 0) __trap();
    uint8_t* a_s = smem;                                   // queries
    uint8_t* b_s = smem + OP_BYTES;                        // [2] candidate stages
    // candidate squared norms: 4 slots, because the scan of tile t still reads slot t%4 while the operand stage t%2 is
    // already being refilled for tile t+2 (slot reuse at t+4 is ordered behind tempty of tile t via MMA(t+2))
    float* xs = reinterpret_cast<float*>(smem + 3 * OP_BYTES);           // [4][KT_C]
    float* list_v = xs + 4 * KT_C;                                         // [128][L]
    int* list_i = reinterpret_cast<int*>(list_v + KT_Q * L);               // [128][L]
    float2* queue = reinterpret_cast<float2*>(list_i + KT_Q * L);          // [128][KT_QSTRIDE] (score, index bits)
    uint64_t* afull = reinterpret_cast<uint64_t*>(queue + KT_Q * KT_QSTRIDE);
    uint64_t* aempty = afull + 1;
    uint64_t* bfull = aempty + 1;
    uint64_t* bempty = bfull + 2;
    uint64_t* tfull = bempty + 2;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int items = P.B * P.qtiles;

    if (warp == 5) {
        if (lane == 0) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(&tmap_x)) : "memory");
            mbar_init(afull, 1); mbar_init(aempty, 1);
            for (int s = 0; s < 2; ++s) { mbar_init(&bfull[s], 1); mbar_init(&bempty[s], 1); mbar_init(&tfull[s], 1); mbar_init(&tempty[s], 4); }
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 256;" ::"r"(smem_u32(tmem_slot)) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 4) {
        // ------------------------------ TMA producer ------------------------------
        if (lane == 0) {
            uint32_t tcount = 0, icount = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++icount) {
                const int b = item / P.qtiles, q0 = (item % P.qtiles) * KT_Q;
                mbar_wait(aempty, (icount & 1) ^ 1);
                mbar_expect_tx(afull, OP_BYTES);
                tma_load_2d(a_s, &tmap_x, afull, 0, b * P.N + q0);
                tma_load_2d(a_s + KB_BYTES, &tmap_x, afull, 32, b * P.N + q0);
                for (int ct = 0; ct < P.ctiles; ++ct, ++tcount) {
                    const uint32_t s = tcount & 1, ph = (tcount >> 1) & 1;
                    mbar_wait(&bempty[s], ph ^ 1);
                    mbar_expect_tx(&bfull[s], OP_BYTES + KT_C * 4);
                    uint8_t* bs = b_s + s * OP_BYTES;
                    tma_load_2d(bs, &tmap_x, &bfull[s], 0, b * P.N + ct * KT_C);
                    tma_load_2d(bs + KB_BYTES, &tmap_x, &bfull[s], 32, b * P.N + ct * KT_C);
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                                 ::"r"(smem_u32(xs + (tcount & 3) * KT_C)), "l"(P.xxpad + (size_t)b * P.Npad + ct * KT_C), "r"(KT_C * 4),
                                   "r"(smem_u32(&bfull[s])) : "memory");
                }
            }
        }
    } else if (warp == 5) {
        // ------------------------------ MMA issuer ------------------------------
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc(KT_Q, KT_C);
            const uint32_t a_addr = smem_u32(a_s);
            uint32_t tcount = 0, icount = 0;
            for (int item = blockIdx.x; item < items; item += gridDim.x, ++icount) {
                mbar_wait(afull, icount & 1);
                for (int ct = 0; ct < P.ctiles; ++ct, ++tcount) {
                    const uint32_t s = tcount & 1, ph = (tcount >> 1) & 1;
                    mbar_wait(&tempty[s], ph ^ 1);
                    mbar_wait(&bfull[s], ph);
                    tc_fence_after();
                    const uint32_t b_addr = smem_u32(b_s + s * OP_BYTES);
#pragma unroll
                    for (int kb = 0; kb < 2; ++kb) {
                        const uint64_t da = make_smem_desc(a_addr + kb * KB_BYTES), db = make_smem_desc(b_addr + kb * KB_BYTES);
#pragma unroll
                        for (int ks = 0; ks < 4; ++ks)
                            tc_mma_tf32(tmem_base + s * KT_C, da + (uint64_t)(ks * 2), db + (uint64_t)(ks * 2), idesc, (kb | ks) != 0 ? 1u : 0u);
                    }
                    tc_commit(&bempty[s]);
                    tc_commit(&tfull[s]);
                }
                tc_commit(aempty);   // all MMAs that read this query tile have completed when this arrives
            }
        }
    } else {
        // ------------------------------ scan + select (one query row per thread) ------------------------------
        const int row = warp * 32 + lane;                 // TMEM lane == row of the query tile
        float2* my_q = queue + row * KT_QSTRIDE;
        uint32_t tcount = 0;
        for (int item = blockIdx.x; item < items; item += gridDim.x) {
            const int b = item / P.qtiles, q0 = (item % P.qtiles) * KT_Q;
            for (int e = lane; e < 32 * L; e += 32) {      // reset this warp's 32 lists
                list_v[warp * 32 * L + e] = -INFINITY;
                list_i[warp * 32 * L + e] = INT_MAX;
            }
            __syncwarp();
            float tau = -INFINITY;
            for (int ct = 0; ct < P.ctiles; ++ct, ++tcount) {
                const uint32_t s = tcount & 1, ph = (tcount >> 1) & 1;
                mbar_wait(&tfull[s], ph);
                tc_fence_after();
                const float* xsj = xs + (tcount & 3) * KT_C;
#pragma unroll 1
                for (int c = 0; c < KT_C / 32; ++c) {
                    uint32_t r[32];
                    tc_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + s * KT_C + c * 32, r);
                    float a[32];
                    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 xx = *reinterpret_cast<const float4*>(xsj + c * 32 + j);   // warp broadcast
                        a[j + 0] = fmaf(2.f, __uint_as_float(r[j + 0]), -xx.x);
                        a[j + 1] = fmaf(2.f, __uint_as_float(r[j + 1]), -xx.y);
                        a[j + 2] = fmaf(2.f, __uint_as_float(r[j + 2]), -xx.z);
                        a[j + 3] = fmaf(2.f, __uint_as_float(r[j + 3]), -xx.w);
                        m4[0] = fmaxf(m4[0], a[j]); m4[1] = fmaxf(m4[1], a[j + 1]);
                        m4[2] = fmaxf(m4[2], a[j + 2]); m4[3] = fmaxf(m4[3], a[j + 3]);
                    }
                    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
                    int cnt = 0;
                    if (mx > tau) {
                        const int jbase = ct * KT_C + c * 32;
#pragma unroll
                        for (int j = 0; j < 32; ++j) {
                            if (a[j] > tau) { my_q[cnt] = make_float2(a[j], __int_as_float(jbase + j)); ++cnt; }
                        }
                    }
                    unsigned pend = __ballot_sync(kFull, cnt > 0);
                    if (pend) {
                        __syncwarp();
                        while (pend) {
                            const int src = __ffs(pend) - 1;
                            pend &= pend - 1;
                            const int n = __shfl_sync(kFull, cnt, src);
                            const int rr = warp * 32 + src;
                            float lv[E]; int li[E];
#pragma unroll
                            for (int e = 0; e < E; ++e) { lv[e] = list_v[rr * L + e * 32 + lane]; li[e] = list_i[rr * L + e * 32 + lane]; }
                            const float2* qq = queue + rr * KT_QSTRIDE;
                            for (int e = 0; e < n; ++e) {
                                const float2 ent = qq[e];
                                list_insert<L>(lv, li, ent.x, __float_as_int(ent.y), lane);
                            }
#pragma unroll
                            for (int e = 0; e < E; ++e) { list_v[rr * L + e * 32 + lane] = lv[e]; list_i[rr * L + e * 32 + lane] = li[e]; }
                            const float nt = __shfl_sync(kFull, lv[E - 1], 31);
                            if (lane == src) tau = nt;
                        }
                        __syncwarp();
                    }
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&tempty[s]);
            }
            // ---------------- refine: canonical re-score of the L candidates of every row, sort, write ----------------
            __syncwarp();
            const float r2 = __ldg(P.r2 + b);
            for (int rl = 0; rl < 32; ++rl) {
                const int qi = q0 + warp * 32 + rl;          // query index inside the cloud
                if (qi >= P.N) break;
                const int rr = warp * 32 + rl;
                float av[E]; int id[E];
#pragma unroll
                for (int e = 0; e < E; ++e) { av[e] = list_v[rr * L + e * 32 + lane]; id[e] = list_i[rr * L + e * 32 + lane]; }
                const float xxi = __ldg(P.xxpad + (size_t)b * P.Npad + qi);
                // guarantee test on the approximate scores (the list is sorted by them)
                const float a_k = __shfl_sync(kFull, av[0], P.k - 1);   // k <= 32: always in the first register
                const float a_L = __shfl_sync(kFull, av[E - 1], 31);
                const float eps = 0.0078125f * sqrtf(xxi) * sqrtf(r2) + 1e-6f * (xxi + r2);
                const bool ok = (a_L == -INFINITY) || (a_L < a_k - 2.f * eps);
                const float4* xi = reinterpret_cast<const float4*>(P.x + ((size_t)b * P.N + qi) * 64);
                float pd[E];
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    pd[e] = -INFINITY;
                    if (id[e] != INT_MAX && id[e] < P.N) {
                        const float4* xj = reinterpret_cast<const float4*>(P.x + ((size_t)b * P.N + id[e]) * 64);
                        float dot = 0.f;
#pragma unroll
                        for (int g = 0; g < 16; ++g) {
                            const float4 u = __ldg(xi + g), w = __ldg(xj + g);
                            dot = __fmaf_rn(u.x, w.x, dot); dot = __fmaf_rn(u.y, w.y, dot);
                            dot = __fmaf_rn(u.z, w.z, dot); dot = __fmaf_rn(u.w, w.w, dot);
                        }
                        const float xxj = __ldg(P.xxpad + (size_t)b * P.Npad + id[e]);
                        const float t = -2.0f * dot;
                        pd[e] = __fsub_rn(__fsub_rn(-xxj, t), xxi);
                    } else id[e] = INT_MAX;
                }
                list_sort<L>(pd, id, lane);
                const size_t o = ((size_t)b * P.N + qi) * P.k;
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int g = e * 32 + lane;
                    if (g < P.k) {
                        if (P.idx_i64) reinterpret_cast<long long*>(P.idx)[o + g] = id[e];
                        else reinterpret_cast<int*>(P.idx)[o + g] = id[e];
                    }
                }
                if (!ok && lane == 0) P.flags[(size_t)b * ((P.N + 63) / 64) + qi / 64] = 1;
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 5) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 256;" ::"r"(tmem_base) : "memory");
    }
}

template <int L>
static int knn_tc_launch(const CUtensorMap& tm, const KnnTcParams& P, cudaStream_t st) {
    const size_t smem = 3 * 2 * 128 * 128 + 4 * KT_C * 4 + (size_t)KT_Q * L * 8 + (size_t)KT_Q * KT_QSTRIDE * 8 + 256;
    LPD_CUDA_CHECK(allow_smem(knn_tc_kernel<L>, smem));
    int dev = 0, sms = 0;
    LPD_CUDA_CHECK(cudaGetDevice(&dev));
    LPD_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const int items = P.B * P.qtiles;
    knn_tc_kernel<L><<<items < sms ? items : sms, KT_THREADS, smem, st>>>(tm, P);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

}  // namespace tc
}  // namespace lpd

extern "C" size_t lpd_knn_workspace_bytes(int B, int N, int C, int k) {
    if (B < 1 || N < 1 || C != 64 || k < 1) return 0;
    const size_t npad = ((size_t)N + 127) / 128 * 128;
    return ((size_t)B * npad + (size_t)B) * sizeof(float) + (size_t)B * ((N + 63) / 64) * sizeof(int) + 256;
}

extern "C" int lpd_knn_tc(const float* x, int B, int N, int C, int k, void* idx, int idx_i64,
                          void* workspace, size_t workspace_bytes, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(x && idx && workspace);
    LPD_REQUIRE(B >= 1 && B <= 65535 && N >= 1 && C == 64 && k >= 1 && k <= 32 && k <= N);
    LPD_REQUIRE(((uintptr_t)x & 15) == 0 && ((uintptr_t)workspace & 15) == 0);
    if (workspace_bytes < lpd_knn_workspace_bytes(B, N, C, k)) return LPD_EWORKSPACE;
    int dev = 0, major = 0;
    LPD_CUDA_CHECK(cudaGetDevice(&dev));
    LPD_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    if (major != 10) return LPD_EUNSUPPORTED;
    cudaStream_t st = as_stream(stream);
    const int npad = (N + 127) / 128 * 128;
    float* xxpad = reinterpret_cast<float*>(workspace);
    float* r2 = xxpad + (size_t)B * npad;
    int* flags = reinterpret_cast<int*>(r2 + B);
    const int ftiles = (N + 63) / 64;
    LPD_CUDA_CHECK(cudaMemsetAsync(r2, 0, (size_t)B * sizeof(float) + (size_t)B * ftiles * sizeof(int), st));
    tc::knn_prep_kernel<<<dim3(ceil_div(npad, 256), B), 256, 0, st>>>(x, N, npad, xxpad, r2);
    LPD_LAUNCH_CHECK();
    CUtensorMap tm;
    int rc = tc::make_tmap(&tm, x, (long long)B * N, 64, 64, 128);
    if (rc != LPD_OK) return rc;
    tc::KnnTcParams P;
    P.x = x; P.xxpad = xxpad; P.r2 = r2; P.flags = flags; P.idx = idx; P.idx_i64 = idx_i64;
    P.B = B; P.N = N; P.Npad = npad; P.k = k; P.qtiles = ceil_div(N, tc::KT_Q); P.ctiles = ceil_div(N, tc::KT_C);
    rc = (k <= 24) ? tc::knn_tc_launch<32>(tm, P, st) : tc::knn_tc_launch<64>(tm, P, st);
    if (rc != LPD_OK) return rc;
    return knn_simt64_flagged(x, B, N, k, idx, idx_i64, flags, st);   // exact recompute of flagged 64-row tiles only
}
