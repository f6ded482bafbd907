/*
 * tt_oracle.c -- TEST INFRASTRUCTURE ONLY (never linked, imported or called by the product path).
 *
 * Plain-C CPU restatement of the arithmetic the reference delegates to TensorFlow on its hot path,
 * with ONE canonical fp32 evaluation order so that "bit-exact" is well defined:
 *
 *   dot products are accumulated sequentially, k ascending, with one fused multiply-add per term
 *   (C99 fmaf: a single rounding per step).  The CUDA exact paths use the same order and the same
 *   fused operation, so their scores / tower outputs can be compared bit for bit.
 *
 * Reference call sites restated here (all paths relative to /root/reference):
 *   - pkg/modelling/models/tower.py:41-49,72-75        Dense(u, relu): relu(x.W + b)
 *   - pkg/modelling/models/two_tower_model.py:90-92    logits = matmul(q, c, transpose_b=True)
 *   - pkg/modelling/indices/brute_force.py:75-83       scores = matmul(q, C^T); top_k(k); gather ids
 *   - tf.math.top_k contract (TF 2.16.2, not vendored): sorted descending, lower index first on ties
 *
 * Parity: the brute-force top-2 fixture of tests/test_indices.py:63-129 and the tie-break KAT-C of
 * SURVEY.md section 9 are checked in tests/test_oracle.py.  TensorFlow's own CPU summation order
 * (Eigen blocking) is unspecified, so for arbitrary fp32 inputs bit-exactness is defined against
 * THIS order ("parity unpinned" beyond the reference fixtures).
 *
 * Build: see oracle/Makefile  (gcc -O2 -mfma -ffp-contract=off -shared -fPIC -lpthread; this image
 * has no libgomp, so row-parallel loops use a small pthread parallel-for).
 */
#define _GNU_SOURCE
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

int tto_version(void) { return 1; }

static int g_threads = 0;

int tto_max_threads(void) {
    if (g_threads > 0) return g_threads;
    long n = sysconf(_SC_NPROCESSORS_ONLN);
    return n > 0 ? (int)n : 1;
}

void tto_set_threads(int n) { g_threads = n > 0 ? n : 0; }

/* ---- minimal parallel-for: fn(ctx, begin, end) over [0, n) in contiguous chunks ---- */
typedef void (*range_fn)(void* ctx, int64_t begin, int64_t end);
typedef struct { range_fn fn; void* ctx; int64_t begin, end; } range_job;
static void* range_tramp(void* p) { range_job* j = (range_job*)p; j->fn(j->ctx, j->begin, j->end); return NULL; }

static void parallel_for(int64_t n, range_fn fn, void* ctx) {
    int nt = tto_max_threads();
    if (nt > 64) nt = 64;
    if ((int64_t)nt > n) nt = (int)(n > 0 ? n : 1);
    if (nt <= 1) { fn(ctx, 0, n); return; }
    pthread_t th[64]; range_job jobs[64];
    int64_t chunk = (n + nt - 1) / nt;
    int started = 0;
    for (int t = 0; t < nt; ++t) {
        int64_t b = t * chunk, e = b + chunk < n ? b + chunk : n;
        if (b >= e) break;
        jobs[t].fn = fn; jobs[t].ctx = ctx; jobs[t].begin = b; jobs[t].end = e;
        if (pthread_create(&th[t], NULL, range_tramp, &jobs[t]) != 0) { fn(ctx, b, e); th[t] = 0; }
        started = t + 1;
    }
    for (int t = 0; t < started; ++t) if (th[t]) pthread_join(th[t], NULL);
}

/* C[i][j] = sum_k A[i][k] * B[j][k]   (both operands K-contiguous), canonical order. */
void tto_gemm_nt_fmaf(const float* A, const float* B, float* C, int64_t M, int64_t N, int64_t K,
                      int64_t lda, int64_t ldb, int64_t ldc) {
    for (int64_t i = 0; i < M; ++i) {
        const float* a = A + i * lda;
        for (int64_t j = 0; j < N; ++j) {
            const float* b = B + j * ldb;
            float acc = 0.0f;
            for (int64_t k = 0; k < K; ++k) acc = fmaf(a[k], b[k], acc);
            C[i * ldc + j] = acc;
        }
    }
}

/* Y = act(X.W + b): X (M,K) row-major ld=ldx, W (K,N) row-major (Keras kernel layout), b (N) or NULL.
 * acc starts at 0, k ascending fmaf, then one rounded add of the bias, then max(.,0) if relu. */
void tto_dense_fmaf(const float* X, const float* W, const float* b, float* Y, int64_t M, int64_t N,
                    int64_t K, int64_t ldx, int64_t ldy, int relu) {
    for (int64_t i = 0; i < M; ++i) {
        const float* x = X + i * ldx;
        for (int64_t j = 0; j < N; ++j) {
            float acc = 0.0f;
            for (int64_t k = 0; k < K; ++k) acc = fmaf(x[k], W[k * N + j], acc);
            if (b) acc = acc + b[j];
            if (relu && !(acc > 0.0f)) acc = 0.0f;
            Y[i * ldy + j] = acc;
        }
    }
}

/* strict "a ranks before b": higher score first, then lower index (tf.math.top_k tie rule). */
static inline int ranks_before(float sa, int64_t ia, float sb, int64_t ib) {
    return (sa > sb) || (sa == sb && ia < ib);
}

/* Binary-heap top-k over one row of n scores.  heap[0] is the WORST kept entry. */
static void topk_row(const float* s, int64_t n, int64_t k, int64_t idx_base, float* hs, int64_t* hi,
                     float* out_s, int64_t* out_i) {
    int64_t m = 0;
    for (int64_t j = 0; j < n; ++j) {
        float v = s[j];
        int64_t id = idx_base + j;
        if (m < k) {
            int64_t c = m++;
            hs[c] = v; hi[c] = id;
            while (c > 0) {
                int64_t p = (c - 1) / 2;
                /* parent must be worse-or-equal than child: swap if parent ranks before child */
                if (ranks_before(hs[p], hi[p], hs[c], hi[c])) {
                    float ts = hs[p]; hs[p] = hs[c]; hs[c] = ts;
                    int64_t ti = hi[p]; hi[p] = hi[c]; hi[c] = ti;
                    c = p;
                } else break;
            }
        } else if (ranks_before(v, id, hs[0], hi[0])) {
            hs[0] = v; hi[0] = id;
            int64_t c = 0;
            for (;;) {
                int64_t l = 2 * c + 1, r = l + 1, w = c;
                if (l < m && ranks_before(hs[w], hi[w], hs[l], hi[l])) w = l;
                if (r < m && ranks_before(hs[w], hi[w], hs[r], hi[r])) w = r;
                if (w == c) break;
                float ts = hs[w]; hs[w] = hs[c]; hs[c] = ts;
                int64_t ti = hi[w]; hi[w] = hi[c]; hi[c] = ti;
                c = w;
            }
        }
    }
    /* pop worst-first into the tail => output sorted best-first */
    for (int64_t e = m; e > 0; --e) {
        out_s[e - 1] = hs[0]; out_i[e - 1] = hi[0];
        hs[0] = hs[e - 1]; hi[0] = hi[e - 1];
        int64_t mm = e - 1, c = 0;
        for (;;) {
            int64_t l = 2 * c + 1, r = l + 1, w = c;
            if (l < mm && ranks_before(hs[w], hi[w], hs[l], hi[l])) w = l;
            if (r < mm && ranks_before(hs[w], hi[w], hs[r], hi[r])) w = r;
            if (w == c) break;
            float ts = hs[w]; hs[w] = hs[c]; hs[c] = ts;
            int64_t ti = hi[w]; hi[w] = hi[c]; hi[c] = ti;
            c = w;
        }
    }
    for (int64_t e = m; e < k; ++e) { out_s[e] = -INFINITY; out_i[e] = -1; }
}

/* top-k of a materialised score matrix (nq, n).  out_* are (nq, k). */
int tto_topk(const float* scores, int64_t nq, int64_t n, int64_t k, float* out_scores, int64_t* out_idx) {
    if (k <= 0) return -1;
    float* hs = (float*)malloc(sizeof(float) * (size_t)k);
    int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)k);
    if (!hs || !hi) { free(hs); free(hi); return -2; }
    for (int64_t q = 0; q < nq; ++q)
        topk_row(scores + q * n, n, k, 0, hs, hi, out_scores + q * k, out_idx + q * k);
    free(hs); free(hi);
    return 0;
}

/* Fused brute force: canonical scores of nq queries against n corpus rows (dim E), top-k each,
 * never materialising more than one row of scores per thread.  brute_force.py:75-83. */
typedef struct {
    const float* Q; const float* C; int64_t n, E, k, idx_base; float* out_scores; int64_t* out_idx; int fail;
} index_ctx;

static void index_range(void* p, int64_t begin, int64_t end) {
    index_ctx* c = (index_ctx*)p;
    float* row = (float*)malloc(sizeof(float) * (size_t)(c->n > 0 ? c->n : 1));
    float* hs = (float*)malloc(sizeof(float) * (size_t)c->k);
    int64_t* hi = (int64_t*)malloc(sizeof(int64_t) * (size_t)c->k);
    if (!row || !hs || !hi) { c->fail = 1; free(row); free(hs); free(hi); return; }
    for (int64_t q = begin; q < end; ++q) {
        const float* a = c->Q + q * c->E;
        for (int64_t j = 0; j < c->n; ++j) {
            const float* b = c->C + j * c->E;
            float acc = 0.0f;
            for (int64_t kk = 0; kk < c->E; ++kk) acc = fmaf(a[kk], b[kk], acc);
            row[j] = acc;
        }
        topk_row(row, c->n, c->k, c->idx_base, hs, hi, c->out_scores + q * c->k, c->out_idx + q * c->k);
    }
    free(row); free(hs); free(hi);
}

int tto_index_topk(const float* Q, const float* C, int64_t nq, int64_t n, int64_t E, int64_t k,
                   int64_t idx_base, float* out_scores, int64_t* out_idx) {
    if (k <= 0) return -1;
    index_ctx c = {Q, C, n, E, k, idx_base, out_scores, out_idx, 0};
    parallel_for(nq, index_range, &c);
    return c.fail ? -2 : 0;
}

/* hits[t] += sum_{b, j<ks[t]} [true_idx[b] == cand[b][j]]   (index_recall.py:54-58, int32 exact) */
void tto_recall_hits(const int64_t* cand, const int64_t* true_idx, int64_t nq, int64_t k_stride,
                     const int32_t* ks, int32_t nk, int32_t* hits) {
    for (int32_t t = 0; t < nk; ++t) {
        int32_t h = 0;
        int64_t kk = ks[t] < k_stride ? ks[t] : k_stride;
        for (int64_t b = 0; b < nq; ++b)
            for (int64_t j = 0; j < kk; ++j) h += (cand[b * k_stride + j] == true_idx[b]);
        hits[t] += h;
    }
}
