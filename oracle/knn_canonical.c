/*
 * TEST INFRASTRUCTURE ONLY — CPU oracle, never linked or called by the product path.
 *
 * Canonical restatement of knn(), reference util/lpdnet_model.py:317-326:
 *     inner = -2 * matmul(x^T, x)                 (:318)
 *     xx    = sum(x ** 2, dim=1)                  (:320)
 *     pd    = -xx - inner ; pd = pd - xx^T        (:322, :324)   == -||x_i - x_j||^2
 *     idx   = pd.topk(k, dim=-1)[1]               (:325)
 * torch leaves the SGEMM accumulation order and the topk tie order unspecified, so the bit-exact
 * target is the canonical form of SURVEY.md App. A.1:
 *     dot_ij = fmaf chain over c = 0..C-1 starting from +0
 *     xx_j   = the same chain on (x_j, x_j)
 *     pd_ij  = ((-xx_j) - (-2 * dot_ij)) - xx_i      fp32, in that order
 *     rank   = pd descending, ties by j ascending
 * Compile with -ffp-contract=off so only the explicit fmaf() calls fuse.
 *
 * Also: the fp64 brute-force restatement of KDTree(db).query(q, k) (reference evaluate.py:168,186-187;
 * scikit-learn promotes float32 to float64 and ranks by sum_d (q_d - x_d)^2), ties by lower index.
 *
 * Layouts: x is POINT-MAJOR [B][N][C] (the kernels' layout); idx is [B][N][k] int32.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

static void insert_desc(float* v, int32_t* id, int k, int* len, float cv, int32_t cj) {
    /* list sorted by (v desc, id asc); insert candidate if it ranks among the first k */
    int n = *len;
    if (n == k) {
        if (!(cv > v[k - 1] || (cv == v[k - 1] && cj < id[k - 1]))) return;
        n = k - 1;
    }
    int pos = n;
    while (pos > 0 && (cv > v[pos - 1] || (cv == v[pos - 1] && cj < id[pos - 1]))) {
        v[pos] = v[pos - 1];
        id[pos] = id[pos - 1];
        --pos;
    }
    v[pos] = cv;
    id[pos] = cj;
    *len = n + 1;
}

void lpd_oracle_knn(const float* x, int B, int N, int C, int k, int32_t* idx, float* pd_out /* nullable [B][N][k] */) {
    for (int b = 0; b < B; ++b) {
        const float* xb = x + (size_t)b * N * C;
        float* xx = (float*)malloc(sizeof(float) * (size_t)N);
        for (int j = 0; j < N; ++j) {
            float acc = 0.0f;
            for (int c = 0; c < C; ++c) acc = fmaf(xb[(size_t)j * C + c], xb[(size_t)j * C + c], acc);
            xx[j] = acc;
        }
        float* v = (float*)malloc(sizeof(float) * (size_t)k);
        int32_t* id = (int32_t*)malloc(sizeof(int32_t) * (size_t)k);
        for (int i = 0; i < N; ++i) {
            int len = 0;
            const float* xi = xb + (size_t)i * C;
            for (int j = 0; j < N; ++j) {
                const float* xj = xb + (size_t)j * C;
                float dot = 0.0f;
                for (int c = 0; c < C; ++c) dot = fmaf(xi[c], xj[c], dot);
                const float t = -2.0f * dot;
                const float u = (-xx[j]) - t;
                const float pd = u - xx[i];
                insert_desc(v, id, k, &len, pd, j);
            }
            memcpy(idx + ((size_t)b * N + i) * k, id, sizeof(int32_t) * (size_t)k);
            if (pd_out) memcpy(pd_out + ((size_t)b * N + i) * k, v, sizeof(float) * (size_t)k);
        }
        free(xx); free(v); free(id);
    }
}

static void insert_asc_d(double* v, int32_t* id, int k, int* len, double cv, int32_t cj) {
    int n = *len;
    if (n == k) {
        if (!(cv < v[k - 1] || (cv == v[k - 1] && cj < id[k - 1]))) return;
        n = k - 1;
    }
    int pos = n;
    while (pos > 0 && (cv < v[pos - 1] || (cv == v[pos - 1] && cj < id[pos - 1]))) {
        v[pos] = v[pos - 1];
        id[pos] = id[pos - 1];
        --pos;
    }
    v[pos] = cv;
    id[pos] = cj;
    *len = n + 1;
}

/* idx [Nq][k] (-1 beyond Ndb), dist [Nq][k] squared distances (nullable) */
void lpd_oracle_retrieval(const float* db, int Ndb, const float* q, int Nq, int D, int k, int32_t* idx, double* dist) {
    for (int i = 0; i < Nq; ++i) {
        double v[64];
        int32_t id[64];
        int len = 0;
        for (int j = 0; j < Ndb; ++j) {
            double acc = 0.0;
            for (int d = 0; d < D; ++d) {
                const double t = (double)q[(size_t)i * D + d] - (double)db[(size_t)j * D + d];
                acc = fma(t, t, acc);
            }
            insert_asc_d(v, id, k, &len, acc, j);
        }
        for (int r = 0; r < k; ++r) {
            idx[(size_t)i * k + r] = r < len ? id[r] : -1;
            if (dist) dist[(size_t)i * k + r] = r < len ? v[r] : INFINITY;
        }
    }
}
