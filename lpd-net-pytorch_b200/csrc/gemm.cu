// General fp32 GEMM on the CUDA cores with a fused per-column affine + activation epilogue.
// This is the "strict" arithmetic path (plain FFMA, k ascending) used for every 1x1 conv /
// linear / matmul of the reference that is not covered by a dedicated fused kernel:
//   lpdnet_model.py:231-232,262,297-305 ; PointNetVlad.py:48,64,76,104,155-170,209-230.
//
// Tiling: CTA tile BM x BN x 16, 256 threads as a 16 x 16 grid, (BM/16) x (BN/16) outputs per
// thread, double-buffered shared memory with register prefetch of the next k-slab.  Operands may
// be stored either way round (see LPD_A_* / LPD_B_*), batched with element strides.
#include "common.cuh"

namespace lpd {

constexpr int GEMM_BK = 16;
constexpr int GEMM_THREADS = 256;

// One "group" = 4 consecutive elements along the operand's contiguous axis.
// KCONTIG: operand stored [MN][K] (K contiguous); otherwise stored [K][MN] (MN contiguous).
// Shared layout is always S[kk][mn] with row stride (BMN + 4).
template <int BMN, bool KCONTIG>
struct TileLoader {
    static constexpr int GROUPS = BMN * GEMM_BK / 4;              // float4 groups per tile
    static constexpr int PER_THREAD = GROUPS / GEMM_THREADS;      // 1 (BMN=64) or 2 (BMN=128)
    static constexpr int STRIDE = BMN + 4;
    static_assert(GROUPS % GEMM_THREADS == 0, "tile/threads mismatch");

    __device__ __forceinline__ static void fetch(const float* __restrict__ base, int ld, int mn0, int k0,
                                                 int MN, int K, bool vec, int tid, float4 (&r)[PER_THREAD]) {
#pragma unroll
        for (int g = 0; g < PER_THREAD; ++g) {
            const int e = tid + g * GEMM_THREADS;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (KCONTIG) {
                const int mn = mn0 + e / (GEMM_BK / 4);
                const int k = k0 + (e % (GEMM_BK / 4)) * 4;
                if (mn < MN) {
                    const float* p = base + (size_t)mn * ld + k;
                    if (vec && k + 3 < K) v = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (k + 0 < K) v.x = __ldg(p + 0);
                        if (k + 1 < K) v.y = __ldg(p + 1);
                        if (k + 2 < K) v.z = __ldg(p + 2);
                        if (k + 3 < K) v.w = __ldg(p + 3);
                    }
                }
            } else {
                const int k = k0 + e / (BMN / 4);
                const int mn = mn0 + (e % (BMN / 4)) * 4;
                if (k < K) {
                    const float* p = base + (size_t)k * ld + mn;
                    if (vec && mn + 3 < MN) v = __ldg(reinterpret_cast<const float4*>(p));
                    else {
                        if (mn + 0 < MN) v.x = __ldg(p + 0);
                        if (mn + 1 < MN) v.y = __ldg(p + 1);
                        if (mn + 2 < MN) v.z = __ldg(p + 2);
                        if (mn + 3 < MN) v.w = __ldg(p + 3);
                    }
                }
            }
            r[g] = v;
        }
    }

    __device__ __forceinline__ static void stash(float* __restrict__ S, int tid, const float4 (&r)[PER_THREAD]) {
#pragma unroll
        for (int g = 0; g < PER_THREAD; ++g) {
            const int e = tid + g * GEMM_THREADS;
            if (KCONTIG) {
                const int mn = e / (GEMM_BK / 4);
                const int kk = (e % (GEMM_BK / 4)) * 4;
                S[(kk + 0) * STRIDE + mn] = r[g].x;
                S[(kk + 1) * STRIDE + mn] = r[g].y;
                S[(kk + 2) * STRIDE + mn] = r[g].z;
                S[(kk + 3) * STRIDE + mn] = r[g].w;
            } else {
                const int kk = e / (BMN / 4);
                const int mn = (e % (BMN / 4)) * 4;
                *reinterpret_cast<float4*>(S + kk * STRIDE + mn) = r[g];
            }
        }
    }
};

struct GemmParams {
    const float* A; const float* B; float* C;
    int lda, ldb, ldc;
    long long sA, sB, sC;
    int M, N, K;
    const float* scale; const float* shift; const float* aux;
    int act; float slope;
    int vecA, vecB, vecC;
};

template <int BM, int BN, bool A_KCONTIG, bool B_KCONTIG>
__global__ void __launch_bounds__(GEMM_THREADS, 2) gemm_kernel(GemmParams p) {
    constexpr int TM = BM / 16, TN = BN / 16;
    using LA = TileLoader<BM, A_KCONTIG>;
    using LB = TileLoader<BN, B_KCONTIG>;
    __shared__ __align__(16) float As[2][GEMM_BK * LA::STRIDE];
    __shared__ __align__(16) float Bs[2][GEMM_BK * LB::STRIDE];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM;
    const float* A = p.A + (size_t)blockIdx.z * p.sA;
    const float* B = p.B + (size_t)blockIdx.z * p.sB;
    float* C = p.C + (size_t)blockIdx.z * p.sC;

    float acc[TM][TN];
#pragma unroll
    for (int i = 0; i < TM; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[LA::PER_THREAD], rb[LB::PER_THREAD];
    LA::fetch(A, p.lda, m0, 0, p.M, p.K, p.vecA, tid, ra);
    LB::fetch(B, p.ldb, n0, 0, p.N, p.K, p.vecB, tid, rb);
    LA::stash(As[0], tid, ra);
    LB::stash(Bs[0], tid, rb);
    __syncthreads();

    const int ktiles = (p.K + GEMM_BK - 1) / GEMM_BK;
    for (int kt = 0; kt < ktiles; ++kt) {
        const int cur = kt & 1;
        if (kt + 1 < ktiles) {
            LA::fetch(A, p.lda, m0, (kt + 1) * GEMM_BK, p.M, p.K, p.vecA, tid, ra);
            LB::fetch(B, p.ldb, n0, (kt + 1) * GEMM_BK, p.N, p.K, p.vecB, tid, rb);
        }
        const float* as = As[cur];
        const float* bs = Bs[cur];
#pragma unroll
        for (int kk = 0; kk < GEMM_BK; ++kk) {
            float a[TM], b[TN];
#pragma unroll
            for (int h = 0; h < TM / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(as + kk * LA::STRIDE + h * (BM / 2) + ty * 4);
                a[h * 4 + 0] = v.x; a[h * 4 + 1] = v.y; a[h * 4 + 2] = v.z; a[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int h = 0; h < TN / 4; ++h) {
                const float4 v = *reinterpret_cast<const float4*>(bs + kk * LB::STRIDE + h * (BN / 2) + tx * 4);
                b[h * 4 + 0] = v.x; b[h * 4 + 1] = v.y; b[h * 4 + 2] = v.z; b[h * 4 + 3] = v.w;
            }
#pragma unroll
            for (int i = 0; i < TM; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < ktiles) {
            LA::stash(As[cur ^ 1], tid, ra);
            LB::stash(Bs[cur ^ 1], tid, rb);
        }
        __syncthreads();
    }

    // ---- epilogue: per-column affine + activation ----
#pragma unroll
    for (int i = 0; i < TM; ++i) {
        const int m = m0 + (i / 4) * (BM / 2) + ty * 4 + (i % 4);
        if (m >= p.M) continue;
#pragma unroll
        for (int h = 0; h < TN / 4; ++h) {
            const int n = n0 + h * (BN / 2) + tx * 4;
            if (n >= p.N) continue;
            float v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                float o = acc[i][h * 4 + j];
                if (n + j < p.N) {
                    if (p.scale) o *= __ldg(p.scale + n + j);
                    if (p.shift) o += __ldg(p.shift + n + j);
                    if (p.act == LPD_ACT_GATE) o = __ldg(p.aux + (size_t)blockIdx.z * p.sC + (size_t)m * p.ldc + n + j) * (1.f / (1.f + expf(-o)));
                    else if (p.act == LPD_ACT_ADD) o += p.aux[(size_t)blockIdx.z * p.sC + (size_t)m * p.ldc + n + j];  // aux may alias C
                    else o = apply_act(o, p.act, p.slope);
                }
                v[j] = o;
            }
            float* dst = C + (size_t)m * p.ldc + n;
            if (p.vecC && n + 3 < p.N) *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
            else {
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    if (n + j < p.N) dst[j] = v[j];
            }
        }
    }
}

template <int BM, int BN>
static int gemm_dispatch_layout(const GemmParams& p, int a_layout, int b_layout, int batch, cudaStream_t st) {
    dim3 grid(ceil_div(p.N, BN), ceil_div(p.M, BM), batch);
    if (a_layout == LPD_A_MK && b_layout == LPD_B_NK) gemm_kernel<BM, BN, true, true><<<grid, GEMM_THREADS, 0, st>>>(p);
    else if (a_layout == LPD_A_MK && b_layout == LPD_B_KN) gemm_kernel<BM, BN, true, false><<<grid, GEMM_THREADS, 0, st>>>(p);
    else if (a_layout == LPD_A_KM && b_layout == LPD_B_NK) gemm_kernel<BM, BN, false, true><<<grid, GEMM_THREADS, 0, st>>>(p);
    else gemm_kernel<BM, BN, false, false><<<grid, GEMM_THREADS, 0, st>>>(p);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

static inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace lpd

extern "C" int lpd_gemm(const float* A, int a_layout, int lda, long long strideA,
                        const float* B, int b_layout, int ldb, long long strideB,
                        float* C, int ldc, long long strideC,
                        int M, int N, int K, int batch,
                        const float* scale, const float* shift, int act, float slope, const float* aux,
                        void* stream) {
    using namespace lpd;
    LPD_REQUIRE(A && B && C);
    LPD_REQUIRE(M >= 1 && N >= 1 && K >= 1 && batch >= 1 && batch <= 65535);
    LPD_REQUIRE(a_layout == LPD_A_MK || a_layout == LPD_A_KM);
    LPD_REQUIRE(b_layout == LPD_B_NK || b_layout == LPD_B_KN);
    LPD_REQUIRE(lda >= (a_layout == LPD_A_MK ? K : M));
    LPD_REQUIRE(ldb >= (b_layout == LPD_B_NK ? K : N));
    LPD_REQUIRE(ldc >= N);
    LPD_REQUIRE(act >= LPD_ACT_NONE && act <= LPD_ACT_ADD);
    LPD_REQUIRE((act != LPD_ACT_GATE && act != LPD_ACT_ADD) || aux != nullptr);
    GemmParams p;
    p.A = A; p.B = B; p.C = C; p.lda = lda; p.ldb = ldb; p.ldc = ldc;
    p.sA = strideA; p.sB = strideB; p.sC = strideC; p.M = M; p.N = N; p.K = K;
    p.scale = scale; p.shift = shift; p.aux = aux; p.act = act; p.slope = slope;
    p.vecA = aligned16(A) && (lda % 4 == 0) && (strideA % 4 == 0);
    p.vecB = aligned16(B) && (ldb % 4 == 0) && (strideB % 4 == 0);
    p.vecC = aligned16(C) && (ldc % 4 == 0) && (strideC % 4 == 0);
    LPD_REQUIRE(ceil_div(M, 64) <= 65535 * 2);
    cudaStream_t st = as_stream(stream);
    const bool smallM = M <= 64, smallN = N <= 64;
    if (smallM && smallN) return gemm_dispatch_layout<64, 64>(p, a_layout, b_layout, batch, st);
    if (smallM) return gemm_dispatch_layout<64, 128>(p, a_layout, b_layout, batch, st);
    if (smallN) return gemm_dispatch_layout<128, 64>(p, a_layout, b_layout, batch, st);
    return gemm_dispatch_layout<128, 128>(p, a_layout, b_layout, batch, st);
}
