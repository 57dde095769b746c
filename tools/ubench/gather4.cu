// Micro-benchmark: row gather into shared memory, three ways (TMA tile::gather4, per-row cp.async.bulk, LDG+STS),
// on the shape of the kNN refine: R query rows x CPR candidate rows of 64 floats out of a [NR][64] fp32 table.
// build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o gather4 gather4.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__); exit(1); } } while (0)
constexpr int WARPS = 8;
constexpr int CPR = 28;          // candidates per row (7 x gather4)
constexpr int ROWB = 256;        // bytes per table row

__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_wait(uint32_t b, uint32_t ph) {
    uint32_t ok = 0;
    while (!ok) asm volatile("{.reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p;}" : "=r"(ok) : "r"(b), "r"(ph) : "memory");
}

// mode 0: gather4, mode 1: per-lane cp.async.bulk of one row, mode 2: LDG.128 + STS.128 (quarter-warp per half row)
template <int MODE, int NBUF>
__global__ void __launch_bounds__(WARPS * 32)
gather_kernel(const __grid_constant__ CUtensorMap tm, const float* __restrict__ x, const int* __restrict__ cand, int R, float* __restrict__ out) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bars[WARPS][NBUF];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    uint8_t* buf = smem + (size_t)w * NBUF * 32 * ROWB;
    if (lane == 0) for (int i = 0; i < NBUF; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(&bars[w][i])));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncthreads();
    const int gw = blockIdx.x * WARPS + w, nw = gridDim.x * WARPS;
    auto issue = [&](int row, int b) {
        const int* c = cand + (size_t)row * CPR;
        uint8_t* dst = buf + (size_t)b * 32 * ROWB;
        if (MODE == 0) {
            if (lane == 0) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[w][b])), "r"(CPR * ROWB) : "memory");
                for (int g = 0; g < CPR / 4; ++g) {
                    const int4 r = *reinterpret_cast<const int4*>(c + 4 * g);
                    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile::gather4.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];"
                                 ::"r"(s32(dst + g * 4 * ROWB)), "l"(reinterpret_cast<uint64_t>(&tm)), "r"(0), "r"(r.x), "r"(r.y), "r"(r.z), "r"(r.w), "r"(s32(&bars[w][b])) : "memory");
                }
            }
        } else if (MODE == 1) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(&bars[w][b])), "r"(CPR * ROWB) : "memory");
            __syncwarp();
            if (lane < CPR) {
                const int j = c[lane];
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(s32(dst + lane * ROWB)), "l"(x + (size_t)j * 64), "r"(ROWB), "r"(s32(&bars[w][b])) : "memory");
            }
        } else {
            const int qw = lane >> 3, ql = lane & 7;
            for (int cc = qw; cc < CPR; cc += 4) {
                const int j = c[cc];
                const float4 v0 = __ldg(reinterpret_cast<const float4*>(x + (size_t)j * 64) + ql);
                const float4 v1 = __ldg(reinterpret_cast<const float4*>(x + (size_t)j * 64) + 8 + ql);
                *reinterpret_cast<float4*>(dst + cc * ROWB + ql * 16) = v0;
                *reinterpret_cast<float4*>(dst + cc * ROWB + 128 + ql * 16) = v1;
            }
        }
    };
    float acc = 0.f;
    int n = 0;
    for (int row = gw; row < R && n < NBUF - 1; row += nw, ++n) issue(row, n);      // prologue: NBUF - 1 rows in flight
    int it = 0;
    for (int row = gw; row < R; row += nw, ++it) {
        const int b = it % NBUF;
        const int ahead = row + (NBUF - 1) * nw;
        if (NBUF > 1 && ahead < R) issue(ahead, (it + NBUF - 1) % NBUF);
        if (NBUF == 1) issue(row, 0);
        if (MODE != 2) mbar_wait(s32(&bars[w][b]), (it / NBUF) & 1); else __syncwarp();
        // consume: lane = candidate, read its row (16 x LDS.128) -> checksum
        if (lane < CPR) {
            const float4* rj = reinterpret_cast<const float4*>(buf + (size_t)b * 32 * ROWB + lane * ROWB);
#pragma unroll
            for (int g = 0; g < 16; ++g) { const float4 v = rj[(g + lane) & 15]; acc += v.x + v.y + v.z + v.w; }
        }
        __syncwarp();
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                             const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

template <int MODE, int NBUF>
static void run(const char* name, const CUtensorMap& tm, const float* x, const int* cand, int R, float* out, double want, int blocks) {
    const size_t smem = (size_t)WARPS * NBUF * 32 * ROWB;
    CK(cudaFuncSetAttribute(gather_kernel<MODE, NBUF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int i = 0; i < 2; ++i) gather_kernel<MODE, NBUF><<<blocks, WARPS * 32, smem>>>(tm, x, cand, R, out);
    CK(cudaDeviceSynchronize());
    CK(cudaEventRecord(e0));
    for (int i = 0; i < 5; ++i) gather_kernel<MODE, NBUF><<<blocks, WARPS * 32, smem>>>(tm, x, cand, R, out);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
    float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
    std::vector<float> h((size_t)blocks * WARPS * 32);
    CK(cudaMemcpy(h.data(), out, h.size() * 4, cudaMemcpyDeviceToHost));
    double s = 0; for (float v : h) s += v;
    printf("%-34s blocks %4d smem %6zu  %.3f ms   checksum %.6e (want %.6e) %s\n", name, blocks, smem, ms / 5, s, want, fabs(s - want) <= 1e-3 * fabs(want) ? "OK" : "MISMATCH");
}

int main(int argc, char** argv) {
    const int boxrows = argc > 1 ? atoi(argv[1]) : 1;
    const int local = argc > 2 ? atoi(argv[2]) : 1;          // 1: candidates near the row (like the kNN), 0: uniform random
    const int NR = 262144, R = 262144;
    std::vector<float> hx((size_t)NR * 64);
    for (int r = 0; r < NR; ++r) for (int c = 0; c < 64; ++c) hx[(size_t)r * 64 + c] = (float)((r % 4096) * 0.001 + c * 0.01);
    std::vector<int> hc((size_t)R * CPR);
    double want = 0;
    srand(1);
    for (int r = 0; r < R; ++r)
        for (int c = 0; c < CPR; ++c) {
            int j = local ? (r / 4096) * 4096 + ((r % 4096) + (rand() % 601) - 300 + 4096) % 4096 : rand() % NR;
            hc[(size_t)r * CPR + c] = j;
            for (int q = 0; q < 64; ++q) want += hx[(size_t)j * 64 + q];
        }
    float *x, *out; int* cand;
    CK(cudaMalloc(&x, hx.size() * 4)); CK(cudaMalloc(&cand, hc.size() * 4)); CK(cudaMalloc(&out, 148 * 8 * WARPS * 32 * 4));
    CK(cudaMemcpy(x, hx.data(), hx.size() * 4, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(cand, hc.data(), hc.size() * 4, cudaMemcpyHostToDevice));
    EncodeFn enc = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&enc, cudaEnableDefault, &qres));
    CUtensorMap tm;
    cuuint64_t dims[2] = {64, (cuuint64_t)NR};
    cuuint64_t strides[1] = {256};
    cuuint32_t box[2] = {64, (cuuint32_t)boxrows};
    cuuint32_t estr[2] = {1, 1};
    CUresult cr = enc(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, x, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                      CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode (box rows %d): CUresult %d; candidates %s\n", boxrows, (int)cr, local ? "local (+-300 rows)" : "uniform random");
    if (cr != CUDA_SUCCESS) return 1;
    run<2, 1>("LDG.128 + STS.128, 1 buffer", tm, x, cand, R, out, want, 148 * 3);
    run<1, 1>("per-lane cp.async.bulk, 1 buffer", tm, x, cand, R, out, want, 148 * 3);
    run<1, 2>("per-lane cp.async.bulk, 2 buffers", tm, x, cand, R, out, want, 148 * 1);
    run<0, 1>("TMA gather4, 1 buffer", tm, x, cand, R, out, want, 148 * 3);
    run<0, 2>("TMA gather4, 2 buffers", tm, x, cand, R, out, want, 148 * 1);
    run<0, 3>("TMA gather4, 3 buffers", tm, x, cand, R, out, want, 148 * 1);
    return 0;
}
