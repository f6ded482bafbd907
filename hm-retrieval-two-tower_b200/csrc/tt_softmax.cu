// tt_softmax.cu -- C-ABI dispatch for the in-batch sampled softmax (tt.h).
#include "tt_common.cuh"

namespace tt {
// tt_softmax_simt.cu
int softmax_fwd_simt(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse,
                     float* loss, float* rowloss, cudaStream_t st);
int softmax_bwd_pass_simt(const float* R, int ldr, const float* T, int ldt, const float* rowv, const float* colv, int nR, int nT, int E,
                          int d, float* G, int ldg, cudaStream_t st);
int logits_simt(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, float* Z, int ldz,
                cudaStream_t st);
// tt_softmax_tc.cu
bool softmax_tc_supported(int ldq, int ldc, int E, const void* Q, const void* C);
int softmax_fwd_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse,
                   float* loss, float* ws, cudaStream_t st);
int softmax_bwd_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, const float* lse, int Bq, int Bc, int E, int off, int which,
                   float* G0, int ldg0, float* G1, int ldg1, float* ws, cudaStream_t st);
int softmax_step_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, int off, float* lse, float* loss,
                    float* dQ, int lddq, float* dC, int lddc, float* ws, cudaStream_t st);
int logits_tc(const float* Q, int ldq, const float* C, int ldc, const float* bias, int Bq, int Bc, int E, float* Z, int ldz, cudaStream_t st);
size_t softmax_tc_workspace_bytes(int Bq, int Bc, int E);
void debug_tc(void* trace, int max_splits);
void debug_flash(void* trace, int mn_lbo, int mn_sbo);
void debug_index_cap(int cap);
void debug_index_stages(float* host_ms8);

static int pick_impl(int impl, int ldq, int ldc, int E, const void* Q, const void* C, const char* who) {
    if (impl == TT_IMPL_SIMT) return TT_IMPL_SIMT;
    bool ok = softmax_tc_supported(ldq, ldc, E, Q, C);
    if (impl == TT_IMPL_TC) {
        if (!ok) {
            set_error("%s: TT_IMPL_TC needs an sm_100 device, E in {32,64,128}, 16-byte aligned rows (ld %% 4 == 0)", who);
            return TT_ERR_UNSUPPORTED;
        }
        return TT_IMPL_TC;
    }
    return ok ? TT_IMPL_TC : TT_IMPL_SIMT;
}
}  // namespace tt

using namespace tt;

namespace tt {
bool index_tc_supported(int ldq, int ldc, int E, int K, int64_t n, const void* Q, const void* C);
}

extern "C" {

int tt_debug_tc(void* trace, int max_splits) {
    debug_tc(trace, max_splits);
    return TT_OK;
}

int tt_debug_flash(void* trace, int mn_lbo, int mn_sbo) {
    debug_flash(trace, mn_lbo, mn_sbo);
    return TT_OK;
}

int tt_debug_index_stages(float* host_ms8) {
    debug_index_stages(host_ms8);
    return TT_OK;
}

int tt_debug_index_cap(int cap) {
    debug_index_cap(cap);
    return TT_OK;
}

int tt_tc_available(int kind, int E) {
    if (kind == 0) return softmax_tc_supported(E, E, E, nullptr, nullptr) ? 1 : 0;
    if (kind == 1) return index_tc_supported(E, E, E, 100, 1 << 20, nullptr, nullptr) ? 1 : 0;
    return 0;
}

size_t tt_softmax_workspace_bytes(int Bq, int Bc, int E) {
    size_t rows = (size_t)(Bq > Bc ? Bq : Bc);
    size_t simt = align_up(rows * sizeof(float), 256) + 512;
    size_t tcb = (E == 32 || E == 64 || E == 128) ? softmax_tc_workspace_bytes(Bq, Bc, E) : 0;
    return simt > tcb ? simt : tcb;
}

int tt_inbatch_softmax_fwd(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, int Bq, int Bc, int E,
                           int diag_offset, float* lse, float* loss, void* ws, size_t ws_bytes, int impl, void* stream) {
    TT_REQUIRE(Q && C && lse && loss, "tt_inbatch_softmax_fwd: null pointer");
    TT_REQUIRE(Bq >= 0 && Bc >= 0 && E >= 1 && ldq >= E && ldc >= E, "tt_inbatch_softmax_fwd: bad shape");
    TT_REQUIRE(diag_offset >= 0 && (Bq == 0 || diag_offset + Bq <= Bc), "tt_inbatch_softmax_fwd: positives (row i -> column i+%d) fall outside Bc=%d", diag_offset, Bc);
    TT_REQUIRE(ws && ws_bytes >= tt_softmax_workspace_bytes(Bq, Bc, E), "tt_inbatch_softmax_fwd: workspace too small");
    cudaStream_t st = as_stream(stream);
    if (Bq == 0) {
        TT_CUDA_OK(cudaMemsetAsync(loss, 0, sizeof(float), st));
        return TT_OK;
    }
    int use = pick_impl(impl, ldq, ldc, E, Q, C, "tt_inbatch_softmax_fwd");
    if (use < 0) return use;
    float* rowloss = reinterpret_cast<float*>(ws);
    if (use == TT_IMPL_TC) return softmax_fwd_tc(Q, ldq, C, ldc, col_bias, Bq, Bc, E, diag_offset, lse, loss, rowloss, st);  // rowloss == ws base
    return softmax_fwd_simt(Q, ldq, C, ldc, col_bias, Bq, Bc, E, diag_offset, lse, loss, rowloss, st);
}

int tt_inbatch_softmax_bwd(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, const float* lse, int Bq, int Bc,
                           int E, int diag_offset, float* dQ, int lddq, float* dC, int lddc, void* ws, size_t ws_bytes, int impl,
                           void* stream) {
    TT_REQUIRE(Q && C && lse && dQ && dC, "tt_inbatch_softmax_bwd: null pointer");
    TT_REQUIRE(ws && ws_bytes >= tt_softmax_workspace_bytes(Bq, Bc, E), "tt_inbatch_softmax_bwd: workspace too small");
    TT_REQUIRE(Bq >= 0 && Bc >= 0 && E >= 1 && ldq >= E && ldc >= E && lddq >= E && lddc >= E, "tt_inbatch_softmax_bwd: bad shape");
    TT_REQUIRE(diag_offset >= 0 && (Bq == 0 || diag_offset + Bq <= Bc), "tt_inbatch_softmax_bwd: bad diag_offset");
    cudaStream_t st = as_stream(stream);
    int use = pick_impl(impl, ldq, ldc, E, Q, C, "tt_inbatch_softmax_bwd");
    if (use < 0) return use;
    if (Bq == 0) {
        if (Bc > 0) TT_CUDA_OK(cudaMemset2DAsync(dC, sizeof(float) * lddc, 0, sizeof(float) * E, Bc, st));
        return TT_OK;
    }
    int rc;
    if (use == TT_IMPL_TC) {
        float* wsf = reinterpret_cast<float*>(ws);
        return softmax_bwd_tc(Q, ldq, C, ldc, col_bias, lse, Bq, Bc, E, diag_offset, 2, dQ, lddq, dC, lddc, wsf, st);
    }
    rc = softmax_bwd_pass_simt(Q, ldq, C, ldc, lse, col_bias, Bq, Bc, E, diag_offset, dQ, lddq, st);
    if (rc) return rc;
    return softmax_bwd_pass_simt(C, ldc, Q, ldq, col_bias, lse, Bc, Bq, E, -diag_offset, dC, lddc, st);
}

/* forward + backward of one training step in one call (loss, lse, dQ, dC): the tensor-core path prepares its operand
 * copies once and runs 5 launches; the exact path is the forward followed by the backward. */
int tt_inbatch_softmax_step(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, int Bq, int Bc, int E, int diag_offset,
                            float* lse, float* loss, float* dQ, int lddq, float* dC, int lddc, void* ws, size_t ws_bytes, int impl, void* stream) {
    TT_REQUIRE(Q && C && lse && loss && dQ && dC, "tt_inbatch_softmax_step: null pointer");
    TT_REQUIRE(Bq >= 0 && Bc >= 0 && E >= 1 && ldq >= E && ldc >= E && lddq >= E && lddc >= E, "tt_inbatch_softmax_step: bad shape");
    TT_REQUIRE(diag_offset >= 0 && (Bq == 0 || diag_offset + Bq <= Bc), "tt_inbatch_softmax_step: bad diag_offset");
    TT_REQUIRE(ws && ws_bytes >= tt_softmax_workspace_bytes(Bq, Bc, E), "tt_inbatch_softmax_step: workspace too small");
    int use = Bq == 0 ? TT_IMPL_SIMT : pick_impl(impl, ldq, ldc, E, Q, C, "tt_inbatch_softmax_step");
    if (use < 0) return use;
    if (use == TT_IMPL_TC)
        return softmax_step_tc(Q, ldq, C, ldc, col_bias, Bq, Bc, E, diag_offset, lse, loss, dQ, lddq, dC, lddc, reinterpret_cast<float*>(ws),
                               as_stream(stream));
    int rc = tt_inbatch_softmax_fwd(Q, ldq, C, ldc, col_bias, Bq, Bc, E, diag_offset, lse, loss, ws, ws_bytes, use, stream);
    if (rc) return rc;
    return tt_inbatch_softmax_bwd(Q, ldq, C, ldc, col_bias, lse, Bq, Bc, E, diag_offset, dQ, lddq, dC, lddc, ws, ws_bytes, use, stream);
}

/* one half of the backward: which = 0 -> dQ (G is (Bq,E)), which = 1 -> dC (G is (Bc,E)).  The halves are
 * independent; with separate workspaces they may run concurrently on two streams. */
int tt_inbatch_softmax_bwd_one(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, const float* lse, int Bq, int Bc,
                               int E, int diag_offset, int which, float* G, int ldg, void* ws, size_t ws_bytes, int impl, void* stream) {
    TT_REQUIRE(Q && C && lse && G, "tt_inbatch_softmax_bwd_one: null pointer");
    TT_REQUIRE(Bq >= 0 && Bc >= 0 && E >= 1 && ldq >= E && ldc >= E && ldg >= E, "tt_inbatch_softmax_bwd_one: bad shape");
    TT_REQUIRE(which == 0 || which == 1, "tt_inbatch_softmax_bwd_one: which must be 0 (dQ) or 1 (dC)");
    TT_REQUIRE(diag_offset >= 0 && (Bq == 0 || diag_offset + Bq <= Bc), "tt_inbatch_softmax_bwd_one: bad diag_offset");
    TT_REQUIRE(ws && ws_bytes >= tt_softmax_workspace_bytes(Bq, Bc, E), "tt_inbatch_softmax_bwd_one: workspace too small");
    cudaStream_t st = as_stream(stream);
    int use = pick_impl(impl, ldq, ldc, E, Q, C, "tt_inbatch_softmax_bwd_one");
    if (use < 0) return use;
    if (Bq == 0) {
        if (which == 1 && Bc > 0) TT_CUDA_OK(cudaMemset2DAsync(G, sizeof(float) * ldg, 0, sizeof(float) * E, Bc, st));
        return TT_OK;
    }
    float* wsf = reinterpret_cast<float*>(ws);
    if (use == TT_IMPL_TC) {
        return softmax_bwd_tc(Q, ldq, C, ldc, col_bias, lse, Bq, Bc, E, diag_offset, which, G, ldg, nullptr, 0, wsf, st);
    }
    return which == 0 ? softmax_bwd_pass_simt(Q, ldq, C, ldc, lse, col_bias, Bq, Bc, E, diag_offset, G, ldg, st)
                      : softmax_bwd_pass_simt(C, ldc, Q, ldq, col_bias, lse, Bc, Bq, E, -diag_offset, G, ldg, st);
}

int tt_logits(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, int Bq, int Bc, int E, float* Z, int ldz,
              int impl, void* stream) {
    // the materialised matrix is an API/test convenience: exact path unless TT_IMPL_TC is asked for explicitly
    TT_REQUIRE(Q && C && Z, "tt_logits: null pointer");
    TT_REQUIRE(Bq >= 0 && Bc >= 0 && E >= 1 && ldq >= E && ldc >= E && ldz >= Bc, "tt_logits: bad shape");
    if (Bq == 0 || Bc == 0) return TT_OK;
    if (impl == TT_IMPL_TC) {
        int use = pick_impl(impl, ldq, ldc, E, Q, C, "tt_logits");
        if (use < 0) return use;
        return logits_tc(Q, ldq, C, ldc, col_bias, Bq, Bc, E, Z, ldz, as_stream(stream));
    }
    return logits_simt(Q, ldq, C, ldc, col_bias, Bq, Bc, E, Z, ldz, as_stream(stream));
}

}  // extern "C"
