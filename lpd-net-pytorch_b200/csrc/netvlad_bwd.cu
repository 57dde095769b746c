// NetVLAD train-mode pieces that are not plain GEMMs: the residual / intra-norm / global-L2 stage with its saved
// norms (forward, PointNetVlad.py:61-74) and its backward, the softmax backward of the soft-assignment
// (PointNetVlad.py:58) with the a_sum gradient folded in, and the cluster_weights2 gradient.
// Also: column max over the points of a cloud WITH argmax and its backward scatter (torch.max / MaxPool2d over N in
// TranformNet / STN3d, lpdnet_model.py:300, PointNetVlad.py:137,169).
#include "common.cuh"

namespace lpd {

// block-wide sum of one float per thread (1024 threads), result broadcast
__device__ __forceinline__ float block_sum_1024(float v, float* red /* [32] */) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(kFull, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    float t = red[threadIdx.x & 31];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) t += __shfl_xor_sync(kFull, t, o);
    return t;
}

// forward, one CTA (1024 threads) per cloud, in place on v[b][d][k] (K == 64):
//   u = vraw - asum[k] * wc2[d][k] ; n1[k] = max(||u[:,k]||, 1e-12) ; w = u / n1 ; n2 = max(||w||, 1e-12) ; v = w / n2
// saves asum [B][64], n1 [B][64], n2 [B]
__global__ void __launch_bounds__(1024)
vlad_finish_train_kernel(float* __restrict__ v, const float* __restrict__ a, const float* __restrict__ wc2, int N, int D,
                         float* __restrict__ asum_out, float* __restrict__ n1_out, float* __restrict__ n2_out) {
    __shared__ float red[16][64];
    __shared__ float asum[64];
    __shared__ float inv[64];
    __shared__ float ginv_s;
    const int b = blockIdx.x, t = threadIdx.x;
    const int k = t & 63, g = t >> 6;
    {   // asum[k] = sum_n a[b][n][k]  (fixed order: 16 row groups, then serial)
        const float* ab = a + (size_t)b * N * 64;
        float acc = 0.f;
        for (int n = g; n < N; n += 16) acc += __ldg(ab + (size_t)n * 64 + k);
        red[g][k] = acc;
        __syncthreads();
        if (t < 64) {
            float s = 0.f;
#pragma unroll
            for (int i = 0; i < 16; ++i) s += red[i][t];
            asum[t] = s;
            asum_out[(size_t)b * 64 + t] = s;
        }
        __syncthreads();
    }
    float* vb = v + (size_t)b * D * 64;
    const float as = asum[k];
    float sq = 0.f;
    for (int d = g; d < D; d += 16) {
        const size_t e = (size_t)d * 64 + k;
        const float x = vb[e] - as * __ldg(wc2 + e);
        vb[e] = x;
        sq = fmaf(x, x, sq);
    }
    __syncthreads();
    red[g][k] = sq;
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[i][t];
        const float nrm = fmaxf(sqrtf(s), 1e-12f);
        inv[t] = 1.f / nrm;
        n1_out[(size_t)b * 64 + t] = nrm;
        red[0][t] = s / (nrm * nrm);
    }
    __syncthreads();
    if (t == 0) {
        float s = 0.f;
        for (int i = 0; i < 64; ++i) s += red[0][i];
        const float nrm = fmaxf(sqrtf(s), 1e-12f);
        ginv_s = 1.f / nrm;
        n2_out[b] = nrm;
    }
    __syncthreads();
    const float sc = inv[k] * ginv_s;
    for (int d = g; d < D; d += 16) vb[(size_t)d * 64 + k] *= sc;
}

// backward of the stage above, one CTA per cloud, in place on dv[b][d][k] (-> du = d loss / d vraw):
//   dw = (dv - v * <v, dv>) / n2                 (global L2; if the norm was clamped: dw = dv / n2)
//   du[:,k] = (dw[:,k] - w[:,k] * <w[:,k], dw[:,k]>) / n1[k],   w = v * n2
//   dasum[k] = - sum_d du[d][k] * wc2[d][k]
__global__ void __launch_bounds__(1024)
vlad_finish_bwd_kernel(float* __restrict__ dv, const float* __restrict__ v, const float* __restrict__ wc2,
                       const float* __restrict__ n1, const float* __restrict__ n2, int D, float* __restrict__ dasum) {
    __shared__ float red32[32];
    __shared__ float red[16][64];
    __shared__ float coldot[64];
    const int b = blockIdx.x, t = threadIdx.x;
    const int k = t & 63, g = t >> 6;
    float* dvb = dv + (size_t)b * D * 64;
    const float* vb = v + (size_t)b * D * 64;
    const float nn2 = n2[b], nn1 = n1[(size_t)b * 64 + k];
    float dot = 0.f;
    for (int d = g; d < D; d += 16) dot = fmaf(vb[(size_t)d * 64 + k], dvb[(size_t)d * 64 + k], dot);
    const float vdv = nn2 > 1e-12f ? block_sum_1024(dot, red32) : 0.f;
    // per column <w, dw> with dw = (dv - v*vdv)/n2, w = v*n2  ->  sum_d v*(dv - v*vdv)
    float cd = 0.f;
    for (int d = g; d < D; d += 16) {
        const float vv = vb[(size_t)d * 64 + k];
        cd = fmaf(vv, dvb[(size_t)d * 64 + k] - vv * vdv, cd);
    }
    __syncthreads();
    red[g][k] = cd;
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[i][t];
        coldot[t] = n1[(size_t)b * 64 + t] > 1e-12f ? s : 0.f;
    }
    __syncthreads();
    const float cdk = coldot[k];
    float da = 0.f;
    for (int d = g; d < D; d += 16) {
        const size_t e = (size_t)d * 64 + k;
        const float vv = vb[e];
        const float dw = (dvb[e] - vv * vdv) / nn2;
        const float w = vv * nn2;
        const float du = (dw - w * cdk) / nn1;
        dvb[e] = du;
        da = fmaf(du, __ldg(wc2 + e), da);
    }
    __syncthreads();
    red[g][k] = da;
    __syncthreads();
    if (t < 64) {
        float s = 0.f;
#pragma unroll
        for (int i = 0; i < 16; ++i) s += red[i][t];
        dasum[(size_t)b * 64 + t] = -s;
    }
}

// dwc2[d][k] = - sum_b du[b][d][k] * asum[b][k]
__global__ void __launch_bounds__(256)
vlad_dwc2_kernel(const float* __restrict__ du, const float* __restrict__ asum, int B, int DK, float* __restrict__ dwc2) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= DK) return;
    const int k = e & 63;
    float s = 0.f;
    for (int b = 0; b < B; ++b) s = fmaf(du[(size_t)b * DK + e], __ldg(asum + b * 64 + k), s);
    dwc2[e] = -s;
}

// softmax backward with the a_sum gradient: ds = a * (g - sum_k a*g), g = da + dasum[cloud]; in place on da. warp per row.
__global__ void __launch_bounds__(256)
softmax64_bwd_kernel(float* __restrict__ da, const float* __restrict__ a, const float* __restrict__ dasum, long long M, int N) {
    const int lane = threadIdx.x & 31;
    const long long row = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= M) return;
    const long long b = row / N;
    float2 g = *reinterpret_cast<const float2*>(da + row * 64 + lane * 2);
    const float2 av = *reinterpret_cast<const float2*>(a + row * 64 + lane * 2);
    if (dasum) {
        const float2 ds = __ldg(reinterpret_cast<const float2*>(dasum + b * 64 + lane * 2));
        g.x += ds.x; g.y += ds.y;
    }
    float s = av.x * g.x + av.y * g.y;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    g.x = av.x * (g.x - s);
    g.y = av.y * (g.y - s);
    *reinterpret_cast<float2*>(da + row * 64 + lane * 2) = g;
}

// out[b][c] = max_n x[b][n][c], arg[b][c] = first n reaching it.   grid (ceil(C/32), B), block (32, 32)
__global__ void colmax_arg_kernel(const float* __restrict__ x, int N, int C, int ldx, float* __restrict__ out, int* __restrict__ arg) {
    __shared__ float red[32][33];
    __shared__ int redi[32][33];
    const int b = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;
    float m = -INFINITY;
    int mi = 0;
    if (c < C) {
        const float* xb = x + (size_t)b * N * ldx + c;
        for (int n = threadIdx.y; n < N; n += 32) {
            const float v = __ldg(xb + (size_t)n * ldx);
            if (v > m) { m = v; mi = n; }
        }
    }
    red[threadIdx.y][threadIdx.x] = m;
    redi[threadIdx.y][threadIdx.x] = mi;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float r = red[0][threadIdx.x];
        int ri = redi[0][threadIdx.x];
        for (int i = 1; i < 32; ++i) {
            const float v = red[i][threadIdx.x];
            const int vi = redi[i][threadIdx.x];
            if (v > r || (v == r && vi < ri)) { r = v; ri = vi; }
        }
        out[(size_t)b * C + c] = r;
        arg[(size_t)b * C + c] = ri;
    }
}

// dx[b][arg[b][c]][c] += dout[b][c]   (dx must be pre-zeroed by the caller if it is a fresh buffer)
__global__ void colmax_bwd_kernel(const float* __restrict__ dout, const int* __restrict__ arg, int N, int C, int lddx,
                                  float* __restrict__ dx, int total) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= total) return;
    const int b = e / C, c = e % C;
    dx[((size_t)b * N + arg[e]) * lddx + c] += dout[e];
}

}  // namespace lpd

using namespace lpd;

extern "C" int lpd_netvlad_finish_train(float* vlad, const float* a, const float* wc2, int B, int N, int D, int K,
                                        float* asum, float* n1, float* n2, void* stream) {
    LPD_REQUIRE(vlad && a && wc2 && asum && n1 && n2 && B >= 1 && B <= 65535 && N >= 1 && D >= 1);
    LPD_REQUIRE(K == 64);
    vlad_finish_train_kernel<<<B, 1024, 0, as_stream(stream)>>>(vlad, a, wc2, N, D, asum, n1, n2);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_netvlad_finish_bwd(float* dv, const float* v, const float* wc2, const float* asum, const float* n1,
                                      const float* n2, int B, int D, int K, float* dasum, float* dwc2, void* stream) {
    LPD_REQUIRE(dv && v && wc2 && asum && n1 && n2 && dasum && dwc2 && B >= 1 && B <= 65535 && D >= 1);
    LPD_REQUIRE(K == 64);
    vlad_finish_bwd_kernel<<<B, 1024, 0, as_stream(stream)>>>(dv, v, wc2, n1, n2, D, dasum);
    LPD_LAUNCH_CHECK();
    vlad_dwc2_kernel<<<ceil_div(D * 64, 256), 256, 0, as_stream(stream)>>>(dv, asum, B, D * 64, dwc2);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_softmax64_bwd(float* da, const float* a, const float* dasum, long long M, int N, void* stream) {
    LPD_REQUIRE(da && a && M >= 1 && N >= 1);
    softmax64_bwd_kernel<<<(unsigned)((M + 7) / 8), 256, 0, as_stream(stream)>>>(da, a, dasum, M, N);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_colmax_arg(const float* x, int B, int N, int C, int ldx, float* out, int32_t* arg, void* stream) {
    LPD_REQUIRE(x && out && arg && B >= 1 && B <= 65535 && N >= 1 && C >= 1 && ldx >= C);
    dim3 grid(ceil_div(C, 32), B), block(32, 32);
    colmax_arg_kernel<<<grid, block, 0, as_stream(stream)>>>(x, N, C, ldx, out, arg);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_colmax_bwd(const float* dout, const int32_t* arg, int B, int N, int C, float* dx, int lddx, void* stream) {
    LPD_REQUIRE(dout && arg && dx && B >= 1 && N >= 1 && C >= 1 && lddx >= C);
    const int total = B * C;
    colmax_bwd_kernel<<<ceil_div(total, 256), 256, 0, as_stream(stream)>>>(dout, arg, N, C, lddx, dx, total);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
