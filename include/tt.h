/*
 * tt.h -- C ABI of libtt.so: the B200 (sm_100a) implementation of the two-tower hot path.
 *
 * The reference (SelvinSelbaraju/hm-retrieval-two-tower) has no FFI / operator layer of its own: its
 * hot path is ~15 TensorFlow op call sites inside pkg/modelling.  Each entry
 * point below replaces one group of those call sites; the reference file:line it stands in for is
 * given per function.  The Python classes in hm-retrieval-two-tower_b200/pkg/modelling bind these
 * with ctypes (INTEGRATION.md shows the binding a maintainer of the reference would add).
 *
 * Conventions
 *   - every function returns 0 on success, <0 on error; tt_last_error() gives a thread-local message;
 *   - all tensor arguments are BORROWED raw device pointers (caller-owned; never freed or retained);
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); no hidden synchronisation,
 *     no cudaMalloc in any call; scratch comes from the caller via *_workspace_bytes();
 *   - matrices are row-major fp32 with an explicit leading dimension in ELEMENTS;
 *   - row ids are int32 with 0 = OOV (StringLookup(num_oov_indices=1), input_layer.py:33-36);
 *   - there is no CPU fallback: on a machine without an sm_100 GPU the calls fail with TT_ERR_CUDA.
 */
#ifndef TT_H_
#define TT_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define TT_OK 0
#define TT_ERR_ARG (-1)
#define TT_ERR_CUDA (-2)
#define TT_ERR_WORKSPACE (-3)
#define TT_ERR_UNSUPPORTED (-4)
#define TT_ERR_INVALID (-5)   /* malformed input data (TFRecord framing / CRC / Example message) */

#define TT_MAX_FEATURES 16   /* features per tower */
#define TT_MAX_SRC 16        /* gradient sources feeding one table: same-named features x data-parallel ranks */
#define TT_MAX_JOBS 32       /* tables updated by one sparse-optimizer call */
#define TT_MAX_KS 8

/* which implementation of a contraction to run */
#define TT_IMPL_AUTO 0
#define TT_IMPL_SIMT 1   /* exact fp32 FMA on CUDA cores, canonical k-ascending order */
#define TT_IMPL_TC 2     /* tcgen05 tensor cores (TF32 operands, fp32 accumulate in TMEM), TMA-fed */

int tt_version(void);
const char* tt_last_error(void);
/* 1 when the current device is compute capability 10.x and the tcgen05 kernels can launch. */
int tt_device_supports_tc(void);
/* 1 when TT_IMPL_AUTO resolves to the tensor-core kernels for the in-batch softmax (kind 0) or the index
 * (kind 1) at joint dimension E on the current device; callers use it to pick the TF32-rounded operands. */
int tt_tc_available(int kind, int E);
/* Debug/profiling knobs of the tensor-core kernels: `trace` (device, u64[ctas][16], or NULL) receives
 * globaltimer stamps of subsequent launches; `max_splits` caps the column splits (0 = default). */
int tt_debug_tc(void* trace, int max_splits);
/* Debug knobs of the two-pass softmax kernels (E = 64 / 128): `trace` (device, u64[ctas][64][8] or NULL) receives globaltimer
 * stamps; mn_lbo / mn_sbo override the byte offsets in the MN-major shared-memory descriptor of the second product (0 = default). */
int tt_debug_flash(void* trace, int mn_lbo, int mn_sbo);
/* Test knob: cap the candidate lists of the tensor-core index filter (0 = default 4K+512) to force the
 * on-device exact fallback. */
int tt_debug_index_cap(int cap);
/* Profiling knob: while `host_ms8` (8 host floats, or NULL to stop) is set, every tensor-core tt_index_topk call adds
 * the device time of its stages (prep, filter, select, collect, rescore, fallback) to it and synchronises. */
int tt_debug_index_stages(float* host_ms8);
/* Where a tensor-core tt_index_topk call keeps its per-query diagnostics inside the caller's workspace: out8 = {byte offset of the
 * int32 fallback flags [nq], of the int32 listed-column counts [nq], of the int32 hit-log fill counts [n_logs], list capacity, log
 * capacity, n_logs, the rank the threshold sits near, filter groups per query}.  Read after the call (diagnostics only). */
int tt_debug_index_layout(int nq, int64_t n, int E, int K, int have_corpus_prepared, int64_t* out8);
/* ------------------------------------------------------------------------------------------------
 * Peer-shareable device memory (one process per GPU): row-sharded embedding tables and the towers' dX
 * blocks are read by the other GPUs' kernels directly over NVLink.  tt_peer_alloc: cudaMalloc (zero-filled) on the
 * current device + a 64-byte CUDA IPC handle to hand to the other processes (any transport; the host layer uses
 * torch.distributed.all_gather_object).  tt_peer_open maps another process's allocation into this one (peer access
 * enabled lazily); the pointer is valid in kernels launched on the current device.  Set-up calls, never on the hot path.
 * ---------------------------------------------------------------------------------------------- */
#define TT_PEER_HANDLE_BYTES 64
int tt_peer_alloc(size_t bytes, void** ptr, void* handle /* TT_PEER_HANDLE_BYTES */);
int tt_peer_open(const void* handle, void** ptr);
int tt_peer_close(void* ptr);
int tt_peer_free(void* ptr);
/* Barrier across the G ranks of a peer group, executed by a kernel on `stream` (capturable in a CUDA graph; no NCCL).
 * `flag_blocks`: device array of G pointers; entry r = rank r's flag block, TT_PEER_FLAG_WORDS(G) zero-initialised uint32 in
 * peer-shareable memory.  `slot` < TT_PEER_SLOTS names the barrier (each call site of a step uses its own slot).  On return of the
 * kernel every rank's work queued on its stream before its own call is complete and visible to peer reads.  All ranks must call
 * the same slots in the same order; a rank that waits longer than TT_PEER_TIMEOUT_S seconds (environment, default 120) traps:
 * the launch fails loudly instead of hanging the GPUs. */
#define TT_PEER_SLOTS 4
#define TT_PEER_FLAG_WORDS(G) (TT_PEER_SLOTS * (1 + (G)))
int tt_peer_barrier(const void* flag_blocks, int rank, int world, int slot, void* stream);
/* Phase timestamps INSIDE a (captured) step: stamp k of the current step writes %globaltimer (ns) to
 * ring[1 + (ring[0] % ring_len) * nk + k]; the last stamp (k == nk - 1) increments the step counter ring[0].  `ring` is
 * 1 + ring_len * nk zero-initialised uint64 on the device.  Measurement aid of the data-parallel step (bench.py dp_phases_ms). */
int tt_stamp(uint64_t* ring, int ring_len, int k, int nk, void* stream);
/* out[i] = sum over ranks r = 0..G-1 (in that order) of src_r[i]; `src_ptrs`: device array of G (peer-mapped) float pointers.
 * Dense-gradient all-reduce of the data-parallel step, read in place. */
int tt_peer_sum_f32(const void* src_ptrs, int world, int64_t n, float* out, void* stream);
/* out[r * n + i] = src_r[i]: all-gather by peer reads (n a multiple of 4, buffers 16-byte aligned).  Candidate embeddings and their
 * logQ terms for cross-GPU in-batch negatives (BASELINE configs[4]); the matching reduce-scatter of dC is tt_peer_sum_f32 over the
 * ranks' partial-gradient buffers offset to this rank's slice. */
int tt_peer_gather_f32(const void* src_ptrs, int world, int64_t n, float* out, void* stream);

/* Number of kernels this library has launched (or captured into a CUDA graph) in this process so far. */
int64_t tt_launch_count(void);

/* ------------------------------------------------------------------------------------------------
 * InputLayer (pkg/modelling/layers/input_layer.py:45-69): one column block per feature.
 * Numeric feature: table == NULL, src = float[B], e = 1.  Categorical: src = int32 ids[B].
 * Ids outside [0, rows) are treated as OOV (row 0).
 * ---------------------------------------------------------------------------------------------- */
typedef struct tt_feature {
    const float* table; /* (rows, e) row-major fp32, or NULL for a numeric feature */
    const void* src;    /* int32 ids[B] (categorical) or float[B] (numeric) */
    int32_t rows;
    int32_t e;
    int32_t col;        /* first output column of this block */
    int32_t shards;     /* 0 or 1: `table` is the whole table.  G > 1: the table is row-sharded over G GPUs and `table` is really a
                         * device array of G base pointers (const float* const*): row i lives in shard i % G at local row i / G.
                         * Shards of other GPUs are peer-mapped HBM (CUDA IPC over NVLink); rows are read where they live. */
} tt_feature;

/* X[b, col_f : col_f+e_f] = table_f[ids_f[b]]  (or the numeric value); columns [D, ldx) are zeroed.
 * Replaces LookupTableFind->ResourceGather->Reshape->ConcatV2 (input_layer.py:37-41,61-68). */
int tt_gather_concat(const tt_feature* feats, int nfeat, int B, int D, float* X, int ldx, void* stream);

/* Y = relu?(X.W + b).  X (B,K) ld ldx; W (K,N) row-major (Keras kernel layout); b (N) or NULL.
 * Optional Y_tf32: a copy of Y rounded to TF32 (round-to-nearest) for the tensor-core logits kernels.
 * Replaces MatMul+BiasAdd+Relu of tf.keras.layers.Dense (tower.py:41-49,72-75).  Canonical
 * k-ascending fmaf accumulation (bit-exact against oracle/tt_oracle.c:tto_dense_fmaf). */
int tt_dense_fwd(const float* X, int ldx, const float* W, const float* b, float* Y, int ldy, float* Y_tf32,
                 int B, int K, int N, int relu, void* stream);

/* Fused InputLayer + first Dense: gathers the A operand straight from the embedding tables.
 * If X_out != NULL the concatenated input is also written (needed by tt_dense_bwd for dW). */
int tt_input_dense_fwd(const tt_feature* feats, int nfeat, int D, const float* W, const float* b, float* X_out,
                       int ldx, float* Y, int ldy, float* Y_tf32, int B, int N, int relu, void* stream);

/* Backward of Y = relu(X.W + b) given dY (autodiff of tower.py:72-75):
 *   dpre = dY * (Y > 0);  dW = X^T.dpre;  db = sum_b dpre;  dX = dpre.W^T (skipped when dX == NULL).
 * dW/db are reduced over the batch in a fixed chunk order (deterministic).  */
size_t tt_dense_bwd_workspace_bytes(int B, int K, int N);
int tt_dense_bwd(const float* X, int ldx, const float* W, const float* Y, int ldy, const float* dY, int lddy,
                 float* dX, int lddx, float* dW, float* db, int B, int K, int N, int relu, void* ws,
                 size_t ws_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * In-batch sampled softmax (two_tower_model.py:90-92,110-124; logq_correction.py:66-71;
 * runner.py:78-83):   S = Q.C^T ; Z = S - col_bias[None,:] ; labels = eye (row i <-> column
 * i + diag_offset) ; loss = sum_i (logsumexp_j Z_ij - Z_i,i+off) ; dZ = softmax_row(Z) - eye.
 * Q (Bq,E), C (Bc,E); col_bias = ln p(candidate j) (NULL = no logQ correction).  S is never
 * written to HBM.  Bc != Bq with diag_offset serves all-gathered negatives.
 *
 * Precision contract.  TT_IMPL_SIMT: exact fp32 (canonical k-ascending fmaf logits, fp32 exponentials and sums).
 * TT_IMPL_TC / AUTO on sm_100 with E in {64, 128}: the two contractions per logit run on fp16 operand tiles with fp32
 * accumulation.  The operands may have ANY finite magnitude: each call scales Q and C by a power of two chosen from their
 * largest |element| (no overflow; an element is lost to fp16 subnormals only below 2^-28 of the largest), and the softmax
 * weights are stored with an exponent offset (fp16 normals down to p = 2^-28).  The positive's logit, its loss term and its
 * gradient term are formed in fp32 from the fp32 operands.  Resulting accuracy against the exact path: loss <= 1e-3
 * relative (typically 1e-6), every row of dQ / dC within 1e-3 of its norm (tests/test_gpu_tc.py).  E = 32: TF32 tiles, same
 * bounds.  Non-finite operands give non-finite results, as in the reference.
 * ---------------------------------------------------------------------------------------------- */
size_t tt_softmax_workspace_bytes(int Bq, int Bc, int E);
int tt_inbatch_softmax_fwd(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, int Bq, int Bc,
                           int E, int diag_offset, float* lse, float* loss, void* ws, size_t ws_bytes, int impl,
                           void* stream);
/* dQ = dZ.C (Bq,E);  dC = dZ^T.Q (Bc,E).  Recomputes S tile by tile from Q, C and lse. */
int tt_inbatch_softmax_bwd(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, const float* lse,
                           int Bq, int Bc, int E, int diag_offset, float* dQ, int lddq, float* dC, int lddc,
                           void* ws, size_t ws_bytes, int impl, void* stream);
/* Forward + backward of one training step in one call: loss, lse, dQ and dC (two_tower_model.py:113-124 minus the optimizer).
 * Same arguments as the two calls above; the tensor-core path prepares its operand copies once (5 launches in all). */
int tt_inbatch_softmax_step(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, int Bq, int Bc, int E,
                            int diag_offset, float* lse, float* loss, float* dQ, int lddq, float* dC, int lddc, void* ws,
                            size_t ws_bytes, int impl, void* stream);
/* One half of the backward (which = 0: dQ into G (Bq,E); which = 1: dC into G (Bc,E)).  The halves are
 * independent: given separate workspaces they may run concurrently on two streams. */
int tt_inbatch_softmax_bwd_one(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, const float* lse,
                               int Bq, int Bc, int E, int diag_offset, int which, float* G, int ldg, void* ws,
                               size_t ws_bytes, int impl, void* stream);
/* Materialised Z (Bq,Bc) for TwoTowerModel.call / LogQCorrection parity tests only. */
int tt_logits(const float* Q, int ldq, const float* C, int ldc, const float* col_bias, int Bq, int Bc, int E,
              float* Z, int ldz, int impl, void* stream);
/* Z = logits - col_bias[None,:]   (LogQCorrection.__call__ on a materialised matrix). */
int tt_logq_apply(const float* logits, int ldl, const float* col_bias, int Bq, int Bc, float* Z, int ldz, void* stream);
/* out[i] = ln(p[i])  (fp32 log, logq_correction.py:69). */
int tt_log_f32(const float* p, float* out, int64_t n, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Optimizers (optimizer_factory.py:15-18 -> tf-keras 2.16 legacy Adagrad / Adam).
 * Sparse path: duplicate ids are summed FIRST (OptimizerV2._deduplicate_indexed_slices), in
 * ascending position order (deterministic, no floating-point atomics), then the update touches
 * each unique row once.
 * ---------------------------------------------------------------------------------------------- */
int tt_dense_adagrad(float* w, float* acc, const float* g, int64_t n, float lr, float eps, void* stream);
int tt_dense_adam(float* w, float* m, float* v, const float* g, int64_t n, float lr_t, float beta1, float beta2,
                  float eps, void* stream);
int tt_fill_f32(float* p, float value, int64_t n, void* stream);

typedef struct tt_sparse_job {
    float* table;  /* (rows, e) */
    float* slot0;  /* Adagrad accumulator, or Adam m */
    float* slot1;  /* Adam v (unused for Adagrad) */
    int32_t rows;
    int32_t e;
    int32_t nsrc;  /* number of (ids, grad) sources feeding this table */
    int32_t n_per_src; /* rows per source (the batch size) */
    int32_t shard_rank;  /* row-sharded table (shard_world > 1): `table`/`slot*` are THIS rank's shard ((rows + G - 1) / G local rows); */
    int32_t shard_world; /* only ids with id % shard_world == shard_rank are applied, at local row id / shard_world.  0/1: whole table */
    const int32_t* ids[TT_MAX_SRC];
    const float* grad[TT_MAX_SRC]; /* first column of this feature's slice of dX (may be peer-mapped memory of another GPU) */
    int32_t grad_ld[TT_MAX_SRC];
} tt_sparse_job;

size_t tt_sparse_workspace_bytes(int njobs, int max_n, int max_e);
/* Stable LSD radix sort of (id, position) for every job; depends on ids only, so it can run on a
 * side stream concurrently with forward/backward. */
int tt_sparse_sort(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, void* stream);
/* The same sort, radix passes [first_pass, end_pass) only (8 bits per pass, as many passes as the largest table's row count needs;
 * end_pass beyond that is clipped).  The persistent softmax kernels hold every SM and are statically partitioned: a sort kernel
 * that runs beside them slows some SMs and with them the whole pass, so the train step runs pass 0 under the tower forward and
 * the remaining passes under the tower backward (same stream or events in between: the passes must run in order). */
int tt_sparse_sort_passes(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, int first_pass, int end_pass,
                          void* stream);
int tt_sparse_adagrad(const tt_sparse_job* jobs, int njobs, float lr, float eps, void* ws, size_t ws_bytes,
                      void* stream);
/* Non-lazy legacy Adam: whole-table decay + update every step. `touched` is a
 * per-job bitmap workspace inside ws. */
int tt_sparse_adam(const tt_sparse_job* jobs, int njobs, float lr_t, float beta1, float beta2, float eps, void* ws,
                   size_t ws_bytes, void* stream);
/* Host-only (no CUDA call; tests): the access width the segmented reduce will use for every job -- vec[j] = floats per
 * lane access (1, 2 or 4: the widest that divides e and the alignment of the table, slots and workspace), gvec[j] = 1 when
 * the gradient sources allow the same width (a feature's dX slice may start at any column), 0 for scalar gradient loads. */
int tt_debug_sparse_plan(const tt_sparse_job* jobs, int njobs, void* ws, size_t ws_bytes, int32_t* vec, int32_t* gvec);

/* ------------------------------------------------------------------------------------------------
 * Brute-force index (pkg/modelling/indices/brute_force.py:75-83): scores = Q.corpus^T, top_k
 * sorted descending with the LOWER index first on equal scores, int32 indices (+ idx_base for a
 * row-sharded corpus).  The (nq, n) score matrix never reaches HBM.
 * Scores returned are the canonical fp32 values (k-ascending fmaf), bit-exact against the oracle.
 * ---------------------------------------------------------------------------------------------- */
/* Operand preparation for the tensor-core filter, done once at index-build time (tt_index_prepare) and passed
 * to every query call:
 *   corpus_prepared  tt_index_prepared_bytes(n, E) bytes, opaque: the rows as tensor-core operand tiles -- fp16 of the rows scaled by
 *                    one power of two (E >= 64) or TF32-rounded fp32 (E = 32) -- stored under a fixed pseudo-random permutation
 *                    (so that neighbouring -- e.g. equally popular -- rows do not share a filter group);
 *   corpus_norms     TT_INDEX_NORM_PAD(n) floats: ||row||_2 in the same order, zero padded to TT_INDEX_ROWS_PAD(n),
 *                    followed by the maximum norm of every chunk of 32 rows, followed by TT_INDEX_ROWS_PAD(n) int32: the
 *                    original row of every permuted position, followed by 32 floats ([0] = the operand scale).
 * Pass both or neither; when NULL they are rebuilt in the workspace on every call (size the workspace with
 * have_corpus_prepared = 0).  The exact fp32 `corpus` stays authoritative: results never depend on the copy. */
#define TT_INDEX_ROWS_PAD(n) ((((n) + 255) / 256 + 1) * 256)
#define TT_INDEX_NORM_PAD(n) (2 * TT_INDEX_ROWS_PAD(n) + TT_INDEX_ROWS_PAD(n) / 32 + 32)
size_t tt_index_prepared_bytes(int64_t n, int E);
int tt_index_prepare(const float* corpus, int ldc, int64_t n, int E, void* corpus_prepared, float* corpus_norms,
                     void* stream);
size_t tt_index_workspace_bytes(int nq, int64_t n, int E, int K, int impl, int have_corpus_prepared);
int tt_index_topk(const float* Q, int ldq, const float* corpus, int ldc, const void* corpus_prepared,
                  const float* corpus_norms, int nq, int64_t n, int E, int K, int64_t idx_base, float* out_scores,
                  int32_t* out_idx, void* ws, size_t ws_bytes, int impl, void* stream);
/* Round-to-nearest TF32 copy of a matrix (operand preparation for TT_IMPL_TC). */
int tt_round_tf32(const float* src, int lds, float* dst, int ldd, int64_t rows, int cols, void* stream);

/* K-way merge of per-shard results laid out (G, nq, K) by (score desc, idx asc) -> (nq, K). */
int tt_topk_merge(const float* scores, const int32_t* idx, int G, int nq, int K, float* out_scores,
                  int32_t* out_idx, void* stream);

/* out[i] = table[max(idx[i], 0)]: row indices -> identifiers on the device (tf.gather(identifiers, indices), brute_force.py:83, for
 * integer identifiers; absent entries (-1) read row 0). */
int tt_take_i32(const int32_t* table, const int32_t* idx, int64_t n, int32_t* out, void* stream);

/* Stage one batch: copy up to TT_MAX_STAGE_COLS feature columns of `rows` 4-byte elements each (int32 row ids, fp32 values) from
 * where the caller holds them -- device memory, or PINNED host memory, which the kernel reads in place over PCIe (no separate
 * cudaMemcpy per feature) -- into the towers' staging buffers, in ONE launch.  kind 0: 4-byte elements copied as they are;
 * kind 1: int64 source elements narrowed to int32.  Replaces one copy per feature of the reference's per-feature tensors
 * (input_layer.py:55-66 consumes a dict of (B, 1) columns). */
#define TT_MAX_STAGE_COLS 32
typedef struct tt_stage_col {
    const void* src;
    void* dst;
    int32_t kind;
    int32_t reserved;
} tt_stage_col;
int tt_stage_columns(const tt_stage_col* cols, int n_cols, int64_t rows, void* stream);

/* Embedding-table initialiser (tf-keras RandomUniform(lo, hi), reference input_layer.py:33-38 builds Embedding() with the default):
 * out[i, c] = lo + (hi - lo) * u(seed, (row0 + i * row_stride) * e + c), u a counter-based hash -> the value of a table cell depends
 * on (seed, global row, column) only, so a row shard initialised by its owner holds what the whole table would hold in those rows. */
int tt_fill_uniform(float* out, int64_t rows_local, int e, int64_t row0, int64_t row_stride, uint64_t seed, float lo, float hi, void* stream);

/* hits[t] += sum_{b, j < ks[t]} [true_idx[b] == cand[b, j]]   (index_recall.py:54-58; int32 exact).
 * cand is (nq, k_stride) int32; hits is int32[nk] on the device and is accumulated into. */
int tt_recall_hits(const int32_t* cand, int k_stride, const int32_t* true_idx, int nq, const int32_t* ks, int nk,
                   int32_t* hits, void* stream);

/* ------------------------------------------------------------------------------------------------------------
 * Host-side helpers on either side of the GPU path (no device work; csrc/tt_host.cu).
 *
 * Vocabulary map -- replaces tf.keras.layers.StringLookup(num_oov_indices=1, vocabulary=v) of input_layer.py:33-36
 * (vocabulary order: features.py:119-127): string -> row id, 0 = out of vocabulary, v[i] -> i + 1 (the first
 * occurrence wins when v holds duplicates).  Strings are byte strings; `blob`/`offsets` is the usual CSR form
 * (string i = blob[offsets[i] .. offsets[i+1])).  Lookups are read-only and may run concurrently. */
void* tt_vocab_create(const char* blob, const int64_t* offsets, int64_t n);
void tt_vocab_destroy(void* vocab);
int64_t tt_vocab_size(void* vocab);
int tt_vocab_lookup(void* vocab, const char* blob, const int64_t* offsets, int64_t n, int32_t* out_rows, int nthreads);
/* n fixed-width cells of `width` bytes, NUL padded on the right (numpy dtype 'S<width>') */
int tt_vocab_lookup_fixed(void* vocab, const char* cells, int64_t n, int width, int32_t* out_rows, int nthreads);

/* TFRecord files -- replace tf.io.TFRecordWriter (tfrecord_writer.py:122-126) and tf.data.TFRecordDataset +
 * tf.io.parse_single_example (tfrecord_dataset.py:48-50,86-88).  A record is {u64 length, u32 masked CRC32C(length),
 * payload, u32 masked CRC32C(payload)}, little endian; masked(c) = rotr(c, 15) + 0xa282ead8. */
uint32_t tt_crc32c(const void* data, size_t n);           /* CRC-32C (Castagnoli), SSE4.2 when available */
uint32_t tt_crc32c_portable(const void* data, size_t n);  /* table-driven twin */
uint32_t tt_crc32c_masked(const void* data, size_t n);
/* Returns the number of records in the file image (writing at most max_records (payload offset, payload length)
 * pairs), or TT_ERR_INVALID for a truncated file / CRC mismatch. */
int64_t tt_tfrecord_scan(const void* file, size_t nbytes, int verify_crc, int64_t* rec_offset, int64_t* rec_len, int64_t max_records);
int tt_tfrecord_frame(const void* payload, uint64_t len, void* out /* len + 16 bytes */);
/* Batch parser of serialized tf.train.Example payloads holding ONE value per requested feature (FixedLenFeature([1])):
 * kind[f] 0 = bytes -> (str_off, str_len)[f * nrec + i] into `file`; 1 = float -> fvals[f * nrec + i]. */
int tt_example_parse(const void* file, const int64_t* rec_offset, const int64_t* rec_len, int64_t nrec, const char* const* names,
                     const int32_t* kind, int nfeat, int64_t* str_off, int64_t* str_len, float* fvals, int nthreads);
/* Copies the n byte strings file[str_off[i] .. str_off[i] + str_len[i]) into fixed-width cells out[i * width ..], NUL padded on the
 * right (a numpy 'S<width>' column); strings longer than `width` or lying outside the nbytes-long file image are an error. */
int tt_gather_cells(const void* file, size_t nbytes, const int64_t* str_off, const int64_t* str_len, int64_t n, int width, char* out, int nthreads);

#ifdef __cplusplus
}
#endif
#endif /* TT_H_ */
