// Fused input layers of the LPD-Net feature nets: conv1 (D -> 64) + BN + act and conv2 (64 -> 64) + BN + act per point in
// ONE pass (reference util/lpdnet_model.py:231-232 for LPDNet, :86-87 for LPDNetOrign).  Strict fp32: the result feeds the
// feature-space kNN.  One thread = one point: the hidden values never leave registers, the second-layer weights are
// broadcast from shared memory, the 128 x 64 output tile of a block is staged so that the stores are coalesced.
// (Two separate FFMA GEMMs moved the 64-wide hidden map through HBM and ran at 22 TFLOP/s: 0.16 ms per 64 clouds.)
#include "common.cuh"

namespace lpd {

constexpr int PW_THREADS = 128;
constexpr int PW_MAXD = 8;
constexpr int PW_STRIDE = 68;        // staged row stride in floats: 16-byte aligned rows, conflict-free STS.128 / LDS.128

// packed fp32 pairs: both halves of fma.rn.f32x2 round exactly like fmaf
typedef unsigned long long pw_f32x2;
__device__ __forceinline__ pw_f32x2 pw_pack2(float v) { pw_f32x2 r; asm("mov.b64 %0, {%1, %1};" : "=l"(r) : "f"(v)); return r; }
__device__ __forceinline__ void pw_unpack2(pw_f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ pw_f32x2 pw_fma2(pw_f32x2 a, pw_f32x2 b, pw_f32x2 c) {
    pw_f32x2 d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

// One thread = one point.  All 64 second-layer accumulators are live (32 packed pairs); the hidden value of channel c is computed
// when the c-loop reaches it (never stored), broadcast into both halves of a pair and multiplied into four outputs per LDS.128 of
// the TRANSPOSED second-layer weights (w2t[c][o]) with two FFMA2: per point 2048 FFMA2 + 1024 LDS.128 instead of 4096 FFMA + 1024
// LDS.128.  Every output is still the fmaf chain over c ascending: results are bit-identical to the scalar form.
__global__ void __launch_bounds__(PW_THREADS)
pointwise_mlp2_kernel(const float* __restrict__ x, int ldx, int D, long long M, const float* __restrict__ w1,
                      const float* __restrict__ s1, const float* __restrict__ t1, const float* __restrict__ w2,
                      const float* __restrict__ s2, const float* __restrict__ t2, float neg_slope, float* __restrict__ out, int ldo) {
    __shared__ __align__(16) float w2t[64 * 64];                   // [c][o]
    __shared__ __align__(16) float w1s[64 * PW_MAXD];
    __shared__ __align__(16) float s1s[64], t1s[64], s2s[64], t2s[64];
    extern __shared__ __align__(16) float stage[];                 // [PW_THREADS][PW_STRIDE]
    for (int i = threadIdx.x; i < 64 * 64; i += PW_THREADS) w2t[(i & 63) * 64 + (i >> 6)] = __ldg(w2 + i);   // w2 [o][c]
    for (int i = threadIdx.x; i < 64 * PW_MAXD; i += PW_THREADS) {
        const int c = i / PW_MAXD, d = i % PW_MAXD;
        w1s[i] = d < D ? __ldg(w1 + c * D + d) : 0.f;              // rows padded to PW_MAXD (the padding multiplies zeros)
    }
    if (threadIdx.x < 64) {
        s1s[threadIdx.x] = __ldg(s1 + threadIdx.x); t1s[threadIdx.x] = __ldg(t1 + threadIdx.x);
        s2s[threadIdx.x] = __ldg(s2 + threadIdx.x); t2s[threadIdx.x] = __ldg(t2 + threadIdx.x);
    }
    __syncthreads();
    const bool vec_out = ((ldo & 3) == 0) && ((reinterpret_cast<uintptr_t>(out) & 15) == 0);
    for (long long p0 = (long long)blockIdx.x * PW_THREADS; p0 < M; p0 += (long long)gridDim.x * PW_THREADS) {
        const long long pt = p0 + threadIdx.x;
        if (pt < M) {
            float xin[PW_MAXD];
#pragma unroll
            for (int d = 0; d < PW_MAXD; ++d) xin[d] = d < D ? __ldg(x + pt * ldx + d) : 0.f;
            pw_f32x2 acc[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) acc[j] = 0ull;
#pragma unroll 8
            for (int c = 0; c < 64; ++c) {
                float a = 0.f;
                if (D <= 4) {                                              // (the usual 3-d input: half the padded row)
                    const float4 wa = *reinterpret_cast<const float4*>(w1s + c * PW_MAXD);
                    a = fmaf(xin[0], wa.x, a);
                    if (D > 1) a = fmaf(xin[1], wa.y, a);
                    if (D > 2) a = fmaf(xin[2], wa.z, a);
                    if (D > 3) a = fmaf(xin[3], wa.w, a);
                } else {
                    const float4 wa = *reinterpret_cast<const float4*>(w1s + c * PW_MAXD), wb = *reinterpret_cast<const float4*>(w1s + c * PW_MAXD + 4);
                    a = fmaf(xin[0], wa.x, a); a = fmaf(xin[1], wa.y, a); a = fmaf(xin[2], wa.z, a); a = fmaf(xin[3], wa.w, a);
                    a = fmaf(xin[4], wb.x, a);
                    if (D > 5) a = fmaf(xin[5], wb.y, a);
                    if (D > 6) a = fmaf(xin[6], wb.z, a);
                    if (D > 7) a = fmaf(xin[7], wb.w, a);
                }
                const float v = fmaf(s1s[c], a, t1s[c]);
                const pw_f32x2 hp = pw_pack2(fmaxf(v, v * neg_slope));
                const ulonglong2* wr = reinterpret_cast<const ulonglong2*>(w2t + c * 64);
#pragma unroll
                for (int g = 0; g < 16; ++g) {
                    const ulonglong2 w = wr[g];                            // warp broadcast: outputs 4g .. 4g+3 of channel c
                    acc[2 * g] = pw_fma2(hp, w.x, acc[2 * g]);
                    acc[2 * g + 1] = pw_fma2(hp, w.y, acc[2 * g + 1]);
                }
            }
            float4* srow = reinterpret_cast<float4*>(stage + threadIdx.x * PW_STRIDE);
#pragma unroll
            for (int g = 0; g < 16; ++g) {
                float r[4];
                pw_unpack2(acc[2 * g], r[0], r[1]);
                pw_unpack2(acc[2 * g + 1], r[2], r[3]);
                const float4 sc = *reinterpret_cast<const float4*>(s2s + 4 * g), sh = *reinterpret_cast<const float4*>(t2s + 4 * g);
                float4 o4;
                o4.x = fmaf(sc.x, r[0], sh.x); o4.y = fmaf(sc.y, r[1], sh.y); o4.z = fmaf(sc.z, r[2], sh.z); o4.w = fmaf(sc.w, r[3], sh.w);
                o4.x = fmaxf(o4.x, o4.x * neg_slope); o4.y = fmaxf(o4.y, o4.y * neg_slope);
                o4.z = fmaxf(o4.z, o4.z * neg_slope); o4.w = fmaxf(o4.w, o4.w * neg_slope);
                srow[g] = o4;
            }
        }
        __syncthreads();
        // the block's rows are contiguous in the output: 64 consecutive floats per row
        const int rows_here = (int)((M - p0) < PW_THREADS ? (M - p0) : PW_THREADS);
        if (vec_out) {
            for (int i = threadIdx.x; i < rows_here * 16; i += PW_THREADS) {
                const int r = i >> 4, c4 = i & 15;
                *reinterpret_cast<float4*>(out + (p0 + r) * ldo + c4 * 4) = *reinterpret_cast<const float4*>(stage + r * PW_STRIDE + c4 * 4);
            }
        } else {
            for (int i = threadIdx.x; i < rows_here * 64; i += PW_THREADS) {
                const int r = i >> 6, c = i & 63;
                out[(p0 + r) * ldo + c] = stage[r * PW_STRIDE + c];
            }
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
    const size_t smem = (size_t)PW_THREADS * PW_STRIDE * sizeof(float);
    LPD_CUDA_CHECK(allow_smem(pointwise_mlp2_kernel, smem + 20 * 1024));
    pointwise_mlp2_kernel<<<(int)blocks, PW_THREADS, smem, as_stream(stream)>>>(x, ldx, D, M, w1, s1, t1, w2, s2, t2, neg_slope, out, ldo);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
