// Fused input layers of the LPD-Net feature nets: conv1 (D -> 64) + BN + act and conv2 (64 -> 64) + BN + act per point in
// ONE pass (reference util/lpdnet_model.py:231-232 for LPDNet, :86-87 for LPDNetOrign).  Strict fp32: the result feeds the
// feature-space kNN.  One thread = one point: the 64 hidden values never leave registers, the second-layer weights are
// broadcast from shared memory, the 128 x 64 output tile of a block is staged so that the stores are coalesced.
// (Two separate FFMA GEMMs moved the 64-wide hidden map through HBM and ran at 22 TFLOP/s: 0.16 ms per 64 clouds.)
#include "common.cuh"

namespace lpd {

constexpr int PW_THREADS = 128;
constexpr int PW_MAXD = 8;

__global__ void __launch_bounds__(PW_THREADS)
pointwise_mlp2_kernel(const float* __restrict__ x, int ldx, int D, long long M, const float* __restrict__ w1,
                      const float* __restrict__ s1, const float* __restrict__ t1, const float* __restrict__ w2,
                      const float* __restrict__ s2, const float* __restrict__ t2, float neg_slope, float* __restrict__ out, int ldo) {
    __shared__ __align__(16) float w2s[64 * 64];
    __shared__ float w1s[64 * PW_MAXD], s1s[64], t1s[64], s2s[64], t2s[64];
    extern __shared__ __align__(16) float stage[];                 // [PW_THREADS][65]
    for (int i = threadIdx.x; i < 64 * 64; i += PW_THREADS) w2s[i] = __ldg(w2 + i);
    for (int i = threadIdx.x; i < 64 * D; i += PW_THREADS) w1s[i] = __ldg(w1 + i);
    if (threadIdx.x < 64) {
        s1s[threadIdx.x] = __ldg(s1 + threadIdx.x); t1s[threadIdx.x] = __ldg(t1 + threadIdx.x);
        s2s[threadIdx.x] = __ldg(s2 + threadIdx.x); t2s[threadIdx.x] = __ldg(t2 + threadIdx.x);
    }
    __syncthreads();
    for (long long p0 = (long long)blockIdx.x * PW_THREADS; p0 < M; p0 += (long long)gridDim.x * PW_THREADS) {
        const long long pt = p0 + threadIdx.x;
        if (pt < M) {
            float xin[PW_MAXD];
#pragma unroll
            for (int d = 0; d < PW_MAXD; ++d) xin[d] = d < D ? __ldg(x + pt * ldx + d) : 0.f;
            float h1[64];
#pragma unroll
            for (int c = 0; c < 64; ++c) {
                float acc = 0.f;
#pragma unroll
                for (int d = 0; d < PW_MAXD; ++d)
                    if (d < D) acc = fmaf(xin[d], w1s[c * D + d], acc);
                const float v = fmaf(s1s[c], acc, t1s[c]);
                h1[c] = fmaxf(v, v * neg_slope);
            }
            for (int o = 0; o < 64; o += 4) {                                        // four independent accumulation chains
                const float4* wr = reinterpret_cast<const float4*>(w2s + o * 64);
                float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
                for (int g = 0; g < 16; ++g) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        const float4 w = wr[q * 16 + g];                             // warp broadcast
                        acc[q] = fmaf(h1[4 * g + 0], w.x, acc[q]); acc[q] = fmaf(h1[4 * g + 1], w.y, acc[q]);
                        acc[q] = fmaf(h1[4 * g + 2], w.z, acc[q]); acc[q] = fmaf(h1[4 * g + 3], w.w, acc[q]);
                    }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float v = fmaf(s2s[o + q], acc[q], t2s[o + q]);
                    stage[threadIdx.x * 65 + o + q] = fmaxf(v, v * neg_slope);
                }
            }
        }
        __syncthreads();
        // the block's rows are contiguous in the output: 64 consecutive floats per row
        const int rows_here = (int)((M - p0) < PW_THREADS ? (M - p0) : PW_THREADS);
        for (int i = threadIdx.x; i < rows_here * 64; i += PW_THREADS) {
            const int r = i >> 6, c = i & 63;
            out[(p0 + r) * ldo + c] = stage[r * 65 + c];
        }
        __syncthreads();
    }
}

}  // namespace lpd

extern "C" int lpd_pointwise_mlp2(const float* x, int ldx, int D, long long M, const float* w1, const float* s1, const float* t1,
                                  const float* w2, const float* s2, const float* t2, int act, float slope, float* out, int ldo,
                                  void* stream) {
    using namespace lpd;
    LPD_REQUIRE(x && w1 && s1 && t1 && w2 && s2 && t2 && out && M >= 1);
    LPD_REQUIRE(D >= 1 && D <= PW_MAXD && ldx >= D && ldo >= 64);
    LPD_REQUIRE(act == LPD_ACT_NONE || act == LPD_ACT_RELU || (act == LPD_ACT_LEAKY && slope >= 0.f && slope <= 1.f));
    const float neg_slope = act == LPD_ACT_NONE ? 1.f : (act == LPD_ACT_RELU ? 0.f : slope);
    long long blocks = (M + PW_THREADS - 1) / PW_THREADS;
    if (blocks > 148 * 8) blocks = 148 * 8;                   // 4 resident blocks per SM (51 KB shared memory each), two waves
    const size_t smem = (size_t)PW_THREADS * 65 * sizeof(float);
    LPD_CUDA_CHECK(allow_smem(pointwise_mlp2_kernel, smem + 20 * 1024));
    pointwise_mlp2_kernel<<<(int)blocks, PW_THREADS, smem, as_stream(stream)>>>(x, ldx, D, M, w1, s1, t1, w2, s2, t2, neg_slope, out, ldo);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
