// NetVLAD pieces that are not plain GEMMs (reference util/PointNetVlad.py:45-83):
//   softmax over the K=64 clusters (after the assignment GEMM + folded BatchNorm1d(K)),
//   the a_sum * cluster_weights2 residual, intra-normalisation over D, global L2 normalisation.
// The contractions themselves (X.Wc, A^T.X, v.W_h, h.W_g) go through lpd_gemm.
#include "common.cuh"
#include <cuda_fp16.h>
#include <cooperative_groups.h>

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

// The same arithmetic spread over a thread-block cluster of 8 CTAs per cloud (D % 128 == 0, D <= 1024): each CTA owns D / 8 rows
// of the cloud's [D][64] aggregate and keeps its 8 values per thread in registers (one read, one write of the 256 KB matrix
// instead of two each; 512 CTAs instead of 64); the per-column squared norms are combined through distributed shared memory in
// fixed rank order, so every CTA of the cluster derives bit-identical scales.  apart = [B][nparts][64] partial column sums of
// the assignment (any split over the points: the 8 slices of asum_kernel or the 32-row blocks of the softmax epilogue).
constexpr int VF_CLUSTER = 8;
constexpr int VF_RPT = 8;        // rows per thread at D = 1024

__global__ void __cluster_dims__(VF_CLUSTER, 1, 1) __launch_bounds__(1024)
vlad_finish_cluster_kernel(float* __restrict__ v, const float* __restrict__ apart, int nparts, const float* __restrict__ wc2, int D) {
    namespace cg = cooperative_groups;
    cg::cluster_group cluster = cg::this_cluster();
    __shared__ float red[16][64];
    __shared__ float asum[64];
    __shared__ float csq[64];
    __shared__ float inv[64];
    __shared__ float nsq[64];
    __shared__ float ginv_s;
    const int b = blockIdx.y, t = threadIdx.x;
    const int k = t & 63, g = t >> 6;
    const int rows = D / VF_CLUSTER, rpt = rows / 16;     // rows of this CTA, rows per thread (<= VF_RPT)
    const int d0 = blockIdx.x * rows;
    {   // asum[k]: the partials in fixed order (16 interleaved chains, then the chains)
        const float* ap = apart + (size_t)b * nparts * 64;
        float s = 0.f;
        for (int i = g; i < nparts; i += 16) s += __ldg(ap + (size_t)i * 64 + k);
        red[g][k] = s;
    }
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[i][t];
        asum[t] = s;
    }
    __syncthreads();
    float* vb = v + ((size_t)b * D + d0) * 64;
    const float* wb = wc2 + (size_t)d0 * 64;
    const float as = asum[k];
    float x[VF_RPT];
    float sq = 0.f;
#pragma unroll
    for (int i = 0; i < VF_RPT; ++i) {
        x[i] = 0.f;
        if (i < rpt) {
            const size_t e = (size_t)(g + 16 * i) * 64 + k;
            x[i] = vb[e] - as * __ldg(wb + e);
            sq = fmaf(x[i], x[i], sq);
        }
    }
    __syncthreads();                                       // red is reused
    red[g][k] = sq;
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[i][t];
        csq[t] = s;
    }
    cluster.sync();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int r = 0; r < VF_CLUSTER; ++r) s += *cluster.map_shared_rank(&csq[t], r);
        const float nrm = fmaxf(sqrtf(s), 1e-12f);
        inv[t] = 1.f / nrm;
        nsq[t] = s / (nrm * nrm);                          // squared norm of the normalised column
    }
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int i = 0; i < 64; ++i) s += nsq[i];
        ginv_s = 1.f / fmaxf(sqrtf(s), 1e-12f);
    }
    __syncthreads();
    const float sc = inv[k] * ginv_s;
#pragma unroll
    for (int i = 0; i < VF_RPT; ++i)
        if (i < rpt) vb[(size_t)(g + 16 * i) * 64 + k] = x[i] * sc;
    cluster.sync();                                        // nobody leaves while its csq may still be read
}

// Tail of NetVLAD + context gating in one launch (PointNetVlad.py:76-81, 103-115), one CTA per descriptor, G groups of O threads:
//   h[o] = s2[o] * sum_s part[s][b][o] + t2[o]                   (fixed-order split-K reduce of the hidden projection + bn2)
//   out[o] = h[o] * sigmoid(sg[o] * sum_i h[i] wg[i][o] + tg[o])  (gating_weights [O][O] as stored, bn1 folded or gating_biases)
// Both sums are latency chains of L2 loads: group g takes every G-th split / the g-th slice of i, eight loads in flight per thread;
// the group partials are combined in fixed order (deterministic).
__global__ void __launch_bounds__(1024)
hidden_gate_kernel(const float* __restrict__ part, int splits, int B, int O, int G, const float* __restrict__ s2,
                   const float* __restrict__ t2, const float* __restrict__ wg, const float* __restrict__ sg,
                   const float* __restrict__ tg, float* __restrict__ out) {
    extern __shared__ float sm[];
    float* hs = sm;                 // [O]
    float* ps = sm + O;             // [G][O]
    const int b = blockIdx.x, o = threadIdx.x % O, g = threadIdx.x / O;
    const bool on = g < G;
    if (on) {
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
        const float* pp = part + (size_t)b * O + o;
        int s = g;
        for (; s + 7 * G < splits; s += 8 * G) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] += __ldg(pp + (size_t)(s + u * G) * B * O);
        }
        for (; s < splits; s += G) acc[0] += __ldg(pp + (size_t)s * B * O);
        ps[g * O + o] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
    __syncthreads();
    if (g == 0) {
        float h = 0.f;
        for (int q = 0; q < G; ++q) h += ps[q * O + o];
        if (s2) h *= s2[o];
        if (t2) h += t2[o];
        hs[o] = h;
    }
    __syncthreads();
    if (on) {
        const int per = (O + G - 1) / G;
        const int i0 = g * per, i1 = min(O, i0 + per);
        float acc[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc[u] = 0.f;
        int i = i0;
        for (; i + 7 < i1; i += 8) {
#pragma unroll
            for (int u = 0; u < 8; ++u) acc[u] = fmaf(hs[i + u], __ldg(wg + (size_t)(i + u) * O + o), acc[u]);
        }
        for (; i < i1; ++i) acc[0] = fmaf(hs[i], __ldg(wg + (size_t)i * O + o), acc[0]);
        ps[g * O + o] = ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
    }
    __syncthreads();
    if (g == 0) {
        float z = 0.f;
        for (int q = 0; q < G; ++q) z += ps[q * O + o];
        if (sg) z *= sg[o];
        if (tg) z += tg[o];
        out[(size_t)b * O + o] = hs[o] * (1.f / (1.f + expf(-z)));
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
    if (D % 128 == 0 && D <= 128 * VF_RPT) vlad_finish_cluster_kernel<<<dim3(VF_CLUSTER, B), 1024, 0, as_stream(stream)>>>(vlad, asum_ws, ASUM_SPLITS, wc2, D);
    else vlad_finish_kernel<<<B, 1024, 0, as_stream(stream)>>>(vlad, asum_ws, wc2, D);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_netvlad_finish_parts(float* vlad, const float* apart, int nparts, const float* wc2, int B, int D, int K,
                                        void* stream) {
    using namespace lpd;
    LPD_REQUIRE(vlad && apart && wc2 && B >= 1 && B <= 65535 && nparts >= 1 && D >= 1);
    LPD_REQUIRE(K == 64);
    LPD_REQUIRE(D % 128 == 0 && D <= 128 * VF_RPT);
    vlad_finish_cluster_kernel<<<dim3(VF_CLUSTER, B), 1024, 0, as_stream(stream)>>>(vlad, apart, nparts, wc2, D);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_hidden_gate(const float* part, int splits, int B, int O, const float* s2, const float* t2, const float* wg,
                               const float* sg, const float* tg, float* out, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(part && wg && out && splits >= 1 && B >= 1 && O >= 1 && O <= 1024);
    const int G = 1024 / O;                                 // thread groups per descriptor (4 at O = 256)
    const int threads = (G * O + 31) / 32 * 32;
    hidden_gate_kernel<<<B, threads, (size_t)(1 + G) * O * sizeof(float), as_stream(stream)>>>(part, splits, B, O, G, s2, t2, wg, sg, tg, out);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
