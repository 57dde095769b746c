// NetVLAD pieces that are not plain GEMMs (reference util/PointNetVlad.py:45-83):
//   softmax over the K=64 clusters (after the assignment GEMM + folded BatchNorm1d(K)),
//   the a_sum * cluster_weights2 residual, intra-normalisation over D, global L2 normalisation.
// The contractions themselves (X.Wc, A^T.X, v.W_h, h.W_g) go through lpd_gemm.
#include "common.cuh"
#include <cuda_fp16.h>

namespace lpd {

// in place: a[m][0..63] = softmax(a[m][0..63]); one warp per row, 2 values per lane
__global__ void __launch_bounds__(256) softmax64_kernel(float* __restrict__ a, long long M, __half2* __restrict__ a_h = nullptr) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    float2 v = *reinterpret_cast<const float2*>(a + row * 64 + lane * 2);
    float mx = fmaxf(v.x, v.y);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(kFull, mx, o));
    v.x = expf(v.x - mx);
    v.y = expf(v.y - mx);
    float s = v.x + v.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    v.x = v.x / s;
    v.y = v.y / s;
    *reinterpret_cast<float2*>(a + row * 64 + lane * 2) = v;
    if (a_h) a_h[row * 32 + lane] = __floats2half2_rn(v.x, v.y);       // fp16 copy: the B operand of the f16-mode aggregate
}

constexpr int ASUM_SPLITS = 8;

// part[b][s][k] = sum over the s-th slice of n of a[b][n][k]      (K == 64, 256 threads)
__global__ void __launch_bounds__(256) asum_kernel(const float* __restrict__ a, int N, float* __restrict__ part) {
    __shared__ float red[4][64];
    const int b = blockIdx.x, s = blockIdx.y;
    const int k = threadIdx.x & 63, g = threadIdx.x >> 6;
    const int per = (N + ASUM_SPLITS - 1) / ASUM_SPLITS;
    const int n0 = s * per, n1 = min(N, n0 + per);
    const float* ab = a + (size_t)b * N * 64;
    float acc = 0.f;
    for (int n = n0 + g; n < n1; n += 4) acc += __ldg(ab + (size_t)n * 64 + k);
    red[g][k] = acc;
    __syncthreads();
    if (g == 0) part[((size_t)b * ASUM_SPLITS + s) * 64 + k] = (red[0][k] + red[1][k]) + (red[2][k] + red[3][k]);
}

// in place on v[b][d][k]: residual, intra-norm over d, global L2.  One CTA (1024 threads) per cloud.
__global__ void __launch_bounds__(1024) vlad_finish_kernel(float* __restrict__ v, const float* __restrict__ part,
                                                           const float* __restrict__ wc2, int D) {
    __shared__ float asum[64];
    __shared__ float red[16][64];
    __shared__ float inv[64];
    __shared__ float ginv_s;
    const int b = blockIdx.x, t = threadIdx.x;
    const int k = t & 63, g = t >> 6;  // 16 row groups
    if (t < 64) {
        float s = 0.f;
        for (int i = 0; i < ASUM_SPLITS; ++i) s += part[((size_t)b * ASUM_SPLITS + i) * 64 + t];
        asum[t] = s;
    }
    __syncthreads();
    float* vb = v + (size_t)b * D * 64;
    const float as = asum[k];
    float sq = 0.f;
    for (int d = g; d < D; d += 16) {
        const size_t e = (size_t)d * 64 + k;
        const float x = vb[e] - as * __ldg(wc2 + e);
        vb[e] = x;
        sq = fmaf(x, x, sq);
    }
    red[g][k] = sq;
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[i][t];
        const float nrm = fmaxf(sqrtf(s), 1e-12f);
        inv[t] = 1.f / nrm;
        red[0][t] = s / (nrm * nrm);  // squared norm of the normalised column
    }
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int i = 0; i < 64; ++i) s += red[0][i];
        ginv_s = 1.f / fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();
    const float sc = inv[k] * ginv_s;
    for (int d = g; d < D; d += 16) {
        const size_t e = (size_t)d * 64 + k;
        vb[e] *= sc;
    }
}

}  // namespace lpd

extern "C" int lpd_netvlad_assign(const float* x, int M, int D, const float* wc, const float* scale,
                                  const float* shift, int K, float* a, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(x && wc && a && M >= 1 && D >= 1);
    LPD_REQUIRE(K == 64);
    int rc = lpd_gemm(x, LPD_A_MK, D, 0, wc, LPD_B_KN, K, 0, a, K, 0, M, K, D, 1, scale, shift, LPD_ACT_NONE, 0.f,
                      nullptr, stream);
    if (rc != LPD_OK) return rc;
    const long long blocks = ((long long)M + 7) / 8;
    softmax64_kernel<<<(unsigned)blocks, 256, 0, as_stream(stream)>>>(a, M);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_softmax64(float* a, long long M, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(a && M >= 1);
    softmax64_kernel<<<(unsigned)((M + 7) / 8), 256, 0, as_stream(stream)>>>(a, M);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_softmax64_f16(float* a, long long M, void* a_h, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(a && a_h && M >= 1 && ((uintptr_t)a_h & 3) == 0);
    softmax64_kernel<<<(unsigned)((M + 7) / 8), 256, 0, as_stream(stream)>>>(a, M, reinterpret_cast<__half2*>(a_h));
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_netvlad_finish(float* vlad, const float* a, const float* wc2, int B, int N, int D, int K,
                                  float* asum_ws, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(vlad && a && wc2 && asum_ws && B >= 1 && B <= 65535 && N >= 1 && D >= 1);
    LPD_REQUIRE(K == 64);
    asum_kernel<<<dim3(B, ASUM_SPLITS), 256, 0, as_stream(stream)>>>(a, N, asum_ws);
    LPD_LAUNCH_CHECK();
    vlad_finish_kernel<<<B, 1024, 0, as_stream(stream)>>>(vlad, asum_ws, wc2, D);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
