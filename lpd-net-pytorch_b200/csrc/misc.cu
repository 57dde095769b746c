// Small memory-bound helpers: library info, BN folding, transpose, column max, split-K reduce.
#include "common.cuh"

namespace lpd {

thread_local char g_last_error[512] = {0};

__global__ void bn_fold_kernel(const float* __restrict__ gamma, const float* __restrict__ beta,
                               const float* __restrict__ mean, const float* __restrict__ var,
                               const float* __restrict__ bias, float eps, int C,
                               float* __restrict__ scale, float* __restrict__ shift) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= C) return;
    // y = (x - mean) / sqrt(var + eps) * gamma + beta      (torch batch_norm, eval mode)
    const float inv = 1.0f / sqrtf(var[c] + eps);
    const float s = (gamma ? gamma[c] : 1.f) * inv;
    float t = (beta ? beta[c] : 0.f) - mean[c] * s;
    if (bias) t += bias[c] * s;
    scale[c] = s;
    shift[c] = t;
}

// 32x32 tiled transpose through shared memory, in[b][R][C] -> out[b][C][R]
__global__ void transpose_kernel(const float* __restrict__ in, float* __restrict__ out, int R, int C) {
    __shared__ float tile[32][33];
    const size_t boff = (size_t)blockIdx.z * R * C;
    const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = r0 + i, c = c0 + threadIdx.x;
        if (r < R && c < C) tile[i][threadIdx.x] = in[boff + (size_t)r * C + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int c = c0 + i, r = r0 + threadIdx.x;
        if (r < R && c < C) out[boff + (size_t)c * R + r] = tile[threadIdx.x][i];
    }
}

// out[b][c] = max_n x[b][n][c]; grid (ceil(C/32), B), block (32, 32): coalesced rows, smem reduce
__global__ void colmax_kernel(const float* __restrict__ x, int N, int C, int ldx, float* __restrict__ out) {
    __shared__ float red[32][33];
    const int b = blockIdx.y;
    const int c = blockIdx.x * 32 + threadIdx.x;
    float m = -INFINITY;
    if (c < C) {
        const float* xb = x + (size_t)b * N * ldx + c;
        for (int n = threadIdx.y; n < N; n += 32) m = fmaxf(m, __ldg(xb + (size_t)n * ldx));
    }
    red[threadIdx.y][threadIdx.x] = m;
    __syncthreads();
    if (threadIdx.y == 0 && c < C) {
        float r = red[0][threadIdx.x];
#pragma unroll
        for (int i = 1; i < 32; ++i) r = fmaxf(r, red[i][threadIdx.x]);
        out[(size_t)b * C + c] = r;
    }
}

__global__ void splitk_reduce_kernel(const float* __restrict__ part, int splits, int MN, int N,
                                     const float* __restrict__ scale, const float* __restrict__ shift,
                                     float* __restrict__ out) {
    const int e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= MN) return;
    float acc = 0.f;
    for (int s = 0; s < splits; ++s) acc += __ldg(part + (size_t)s * MN + e);  // fixed order: deterministic
    const int n = e % N;
    if (scale) acc *= scale[n];
    if (shift) acc += shift[n];
    out[e] = acc;
}

}  // namespace lpd

extern "C" int lpd_abi_version(void) { return LPD_ABI_VERSION; }

extern "C" const char* lpd_status_str(int status) {
    switch (status) {
        case LPD_OK: return "ok";
        case LPD_EINVAL: return "invalid argument";
        case LPD_EWORKSPACE: return "workspace too small";
        case LPD_ECUDA: return "CUDA error";
        case LPD_EUNSUPPORTED: return "unsupported device or feature";
        default: return "unknown status";
    }
}

extern "C" const char* lpd_last_cuda_error(void) { return lpd::g_last_error; }

extern "C" int lpd_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* smem_optin_bytes) {
    using namespace lpd;
    int dev = 0;
    LPD_CUDA_CHECK(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    LPD_CUDA_CHECK(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    if (smem_optin_bytes) *smem_optin_bytes = prop.sharedMemPerBlockOptin;
    return LPD_OK;
}

extern "C" int lpd_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var,
                           const float* bias, float eps, int C, float* scale, float* shift, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(mean && var && scale && shift && C >= 1);
    bn_fold_kernel<<<ceil_div(C, 128), 128, 0, as_stream(stream)>>>(gamma, beta, mean, var, bias, eps, C, scale, shift);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_transpose(const float* in, float* out, int batch, int rows, int cols, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(in && out && batch >= 1 && batch <= 65535 && rows >= 1 && cols >= 1);
    LPD_REQUIRE(ceil_div(rows, 32) <= 65535);
    dim3 grid(ceil_div(cols, 32), ceil_div(rows, 32), batch), block(32, 8);
    transpose_kernel<<<grid, block, 0, as_stream(stream)>>>(in, out, rows, cols);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_colmax(const float* x, int B, int N, int C, int ldx, float* out, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(x && out && B >= 1 && B <= 65535 && N >= 1 && C >= 1 && ldx >= C);
    dim3 grid(ceil_div(C, 32), B), block(32, 32);
    colmax_kernel<<<grid, block, 0, as_stream(stream)>>>(x, N, C, ldx, out);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}

extern "C" int lpd_splitk_reduce(const float* part, int splits, int M, int N, const float* scale,
                                 const float* shift, float* out, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(part && out && splits >= 1 && M >= 1 && N >= 1);
    const int MN = M * N;
    splitk_reduce_kernel<<<ceil_div(MN, 256), 256, 0, as_stream(stream)>>>(part, splits, MN, N, scale, shift, out);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
