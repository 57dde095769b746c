// Lazy / non-lazy triplet and quadruplet hinge loss, forward (+ optional backward) in one launch.
// Replaces best_pos_distance / triplet_loss / quadruplet_loss, reference loss/pointnetvlad_loss.py:6-97
// (about ten tiny torch kernels plus .repeat copies per call).  The problem is a few KB: one CTA.
//
//   dpos[b][p] = ||pos[b][p] - q[b]||^2        positive[b] = min_p or max_p        (:6-12, :52-57)
//   l1[b][n] = max(0, m1 + positive[b] - ||neg[b][n] - q[b]||^2)                     (:65-66)
//   l2[b][n] = max(0, m2 + positive[b] - ||neg[b][n] - other[b]||^2)                 (:80-82)
//   t1[b] = max_n l1 (lazy) or sum_n l1 ; T1 = mean_b t1, or sum_b t1 / (#(t1 > 1e-16) + 1e-16)   (:68-78)
//   loss = T1 + T2                                                                   (:96)
// Backward follows autograd: gradients of min/max go to the selected element (first on ties),
// clamp passes the gradient where the pre-clamp value is >= 0.
#include "common.cuh"

namespace lpd {

struct LossParams {
    const float* q; const float* pos; const float* neg; const float* other;
    int Bq, P, Nn, D; float m1, m2; int use_min, lazy, ignore_zero;
    float* loss; float* gq; float* gpos; float* gneg; float* gother; const float* grad_out;
};

__device__ __forceinline__ float warp_sqdist(const float* __restrict__ a, const float* __restrict__ b, int D, int lane) {
    float s = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float t = a[d] - b[d];
        s = fmaf(t, t, s);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(kFull, s, o);
    return s;
}

__global__ void __launch_bounds__(256) quadruplet_loss_kernel(LossParams L) {
    extern __shared__ float sm[];
    const int Bq = L.Bq, P = L.P, Nn = L.Nn, D = L.D;
    float* dpos = sm;                 // [Bq][P]
    float* d1 = dpos + Bq * P;        // [Bq][Nn]   ||neg - q||^2  -> later a1 (dL/dl1)
    float* d2 = d1 + Bq * Nn;         // [Bq][Nn]   ||neg - other||^2 -> later a2
    float* t1 = d2 + Bq * Nn;         // [Bq]
    float* t2 = t1 + Bq;              // [Bq]
    float* ap = t2 + Bq;              // [Bq] dL/dpositive
    int* pstar = reinterpret_cast<int*>(ap + Bq);  // [Bq]
    __shared__ float coef[2];

    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
    const bool quad = L.other != nullptr;
    const int per_b = P + Nn + (quad ? Nn : 0);
    for (int task = warp; task < Bq * per_b; task += nwarps) {
        const int b = task / per_b, r = task % per_b;
        const float* qb = L.q + (size_t)b * D;
        if (r < P) {
            const float s = warp_sqdist(L.pos + ((size_t)b * P + r) * D, qb, D, lane);
            if (lane == 0) dpos[b * P + r] = s;
        } else if (r < P + Nn) {
            const int n = r - P;
            const float s = warp_sqdist(L.neg + ((size_t)b * Nn + n) * D, qb, D, lane);
            if (lane == 0) d1[b * Nn + n] = s;
        } else {
            const int n = r - P - Nn;
            const float s = warp_sqdist(L.neg + ((size_t)b * Nn + n) * D, L.other + (size_t)b * D, D, lane);
            if (lane == 0) d2[b * Nn + n] = s;
        }
    }
    __syncthreads();

    // per query tuple: hinge terms; d1/d2 are overwritten by 0/1 activity masks
    for (int b = threadIdx.x; b < Bq; b += blockDim.x) {
        int ps = 0;
        float pv = dpos[b * P];
        for (int p = 1; p < P; ++p) {
            const float v = dpos[b * P + p];
            if (L.use_min ? (v < pv) : (v > pv)) { pv = v; ps = p; }
        }
        pstar[b] = ps;
        for (int pass = 0; pass < (quad ? 2 : 1); ++pass) {
            float* dd = pass == 0 ? d1 : d2;
            const float margin = pass == 0 ? L.m1 : L.m2;
            float acc = 0.f;
            if (L.lazy) {
                float best = -1.f, pre_best = 0.f;
                int bi = 0;
                for (int n = 0; n < Nn; ++n) {
                    const float pre = margin + pv - dd[b * Nn + n];
                    const float l = fmaxf(pre, 0.f);
                    if (l > best) { best = l; bi = n; pre_best = pre; }  // first maximum wins
                }
                for (int n = 0; n < Nn; ++n) dd[b * Nn + n] = 0.f;
                dd[b * Nn + bi] = pre_best >= 0.f ? 1.f : 0.f;
                acc = best;
            } else {
                for (int n = 0; n < Nn; ++n) {
                    const float pre = margin + pv - dd[b * Nn + n];
                    acc += fmaxf(pre, 0.f);
                    dd[b * Nn + n] = pre >= 0.f ? 1.f : 0.f;
                }
            }
            (pass == 0 ? t1 : t2)[b] = acc;
        }
    }
    __syncthreads();

    if (threadIdx.x == 0) {
        float T[2] = {0.f, 0.f};
        for (int pass = 0; pass < (quad ? 2 : 1); ++pass) {
            const float* tt = pass == 0 ? t1 : t2;
            float s = 0.f, cnt = 0.f;
            for (int b = 0; b < Bq; ++b) { s += tt[b]; cnt += tt[b] > 1e-16f ? 1.f : 0.f; }
            if (L.ignore_zero) { T[pass] = s / (cnt + 1e-16f); coef[pass] = 1.f / (cnt + 1e-16f); }
            else { T[pass] = s / (float)Bq; coef[pass] = 1.f / (float)Bq; }
        }
        if (!quad) coef[1] = 0.f;
        L.loss[0] = T[0] + T[1];
    }
    if (L.gq == nullptr) return;
    __syncthreads();

    const float go = L.grad_out ? L.grad_out[0] : 1.f;
    const float c1 = coef[0] * go, c2 = coef[1] * go;
    // scale masks into dL/dl and accumulate dL/dpositive
    for (int b = threadIdx.x; b < Bq; b += blockDim.x) {
        float s = 0.f;
        for (int n = 0; n < Nn; ++n) {
            d1[b * Nn + n] *= c1;
            s += d1[b * Nn + n];
            if (quad) { d2[b * Nn + n] *= c2; s += d2[b * Nn + n]; }
        }
        ap[b] = s;
    }
    __syncthreads();

    for (int e = threadIdx.x; e < Bq * D; e += blockDim.x) {
        const int b = e / D, d = e % D;
        const float qv = L.q[e];
        const float ov = quad ? L.other[e] : 0.f;
        float gqv = 0.f, gov = 0.f;
        for (int n = 0; n < Nn; ++n) {
            const size_t ne = ((size_t)b * Nn + n) * D + d;
            const float nv = L.neg[ne];
            const float a1 = d1[b * Nn + n];
            const float a2 = quad ? d2[b * Nn + n] : 0.f;
            const float dq = 2.f * (nv - qv);
            const float dob = 2.f * (nv - ov);
            L.gneg[ne] = -a1 * dq - a2 * dob;
            gqv += a1 * dq;
            gov += a2 * dob;
        }
        const int ps = pstar[b];
        for (int p = 0; p < P; ++p) {
            const size_t pe = ((size_t)b * P + p) * D + d;
            const float dp = 2.f * (L.pos[pe] - qv);
            const float g = (p == ps) ? ap[b] * dp : 0.f;
            L.gpos[pe] = g;
            gqv -= g;
        }
        L.gq[e] = gqv;
        if (quad && L.gother) L.gother[e] = gov;
    }
}

}  // namespace lpd

extern "C" int lpd_quadruplet_loss(const float* q, const float* pos, const float* neg, const float* other,
                                   int Bq, int P, int Nn, int D, float m1, float m2, int flags,
                                   float* loss, float* gq, float* gpos, float* gneg, float* gother,
                                   const float* grad_out, void* stream) {
    using namespace lpd;
    LPD_REQUIRE(q && pos && neg && loss);
    LPD_REQUIRE(Bq >= 1 && P >= 1 && Nn >= 1 && D >= 1);
    LPD_REQUIRE((gq == nullptr) == (gpos == nullptr) && (gq == nullptr) == (gneg == nullptr));
    LPD_REQUIRE(!(gq && other && !gother));
    const size_t smem = ((size_t)Bq * (P + 2 * Nn) + 4 * (size_t)Bq) * sizeof(float);
    LPD_REQUIRE(smem <= 200 * 1024);
    LossParams L;
    L.q = q; L.pos = pos; L.neg = neg; L.other = other; L.Bq = Bq; L.P = P; L.Nn = Nn; L.D = D;
    L.m1 = m1; L.m2 = m2; L.use_min = flags & 1; L.lazy = (flags >> 1) & 1; L.ignore_zero = (flags >> 2) & 1;
    L.loss = loss; L.gq = gq; L.gpos = gpos; L.gneg = gneg; L.gother = gother; L.grad_out = grad_out;
    LPD_CUDA_CHECK(allow_smem(quadruplet_loss_kernel, smem > 48 * 1024 ? smem : 48 * 1024));
    quadruplet_loss_kernel<<<1, 256, smem, as_stream(stream)>>>(L);
    LPD_LAUNCH_CHECK();
    return LPD_OK;
}
