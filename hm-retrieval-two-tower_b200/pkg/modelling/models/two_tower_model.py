"""
TwoTowerModel: two Towers, dot-product logits, in-batch sampled softmax with optional logQ correction
(reference pkg/modelling/models/two_tower_model.py:13-205).

train_step (reference :94-130) = tape -> logQ -> labels=eye(B) -> CE(from_logits, SUM) -> minimize.
Here it is a fixed sequence of CUDA launches on pre-allocated buffers (optionally one CUDA graph):

    (eager)     : every id / numeric column of the batch into the staging buffers, one launch    tt_stage_columns
    side stream : ln p(candidate) per column; radix pass 0 of the (id, position) sort          tt_gather_concat, tt_sparse_sort_passes
    main / cand : gather + Dense per tower, the two towers on two streams (fp32)                tt_input_dense_fwd / tt_dense_fwd
    main stream : S = Q.C^T, Z = S - ln p, LSE, loss, dQ, dC in ONE call -- two tensor-core      tt_inbatch_softmax_step
                  passes on power-of-two scaled fp16 operand tiles, S never leaves the SM
    side stream : the remaining radix pass (training batches sort in two 11-bit passes)          tt_sparse_sort_passes
    main / cand : Dense backward per tower (dW, db, dX)                                          tt_dense_bwd
                  [data parallel: device barriers; dense gradients summed, gradient rows         tt_peer_barrier, tt_peer_sum_f32
                   pulled over NVLink peer memory -- pkg/modelling/distributed.py]
                  Adagrad | Adam on the flat Dense buffer                                        tt_dense_adagrad | tt_dense_adam
                  join side stream; de-duplicated row update of every table                      tt_sparse_adagrad | tt_sparse_adam
The sort passes are placed so that no kernel runs beside the persistent softmax passes.
"""
from __future__ import annotations

import logging
import os
from typing import Dict, List, Optional

import numpy as np

from pkg import _native as N
from pkg.modelling import _device as D
from pkg.modelling.layers.logq_correction import LogQCorrection
from pkg.modelling.losses import CategoricalCrossentropy
from pkg.modelling.models.abstract_keras_model import AbstractKerasModel, TensorSpec
from pkg.modelling.models.tower import Tower
from pkg.modelling.optimizer_factory import Adagrad, Adam
from pkg.schema.features import Feature
from pkg.schema.schema import Schema

logger = logging.getLogger(__name__)


class _StepWorkspace:
    """Everything one train step of batch size B touches, allocated once."""

    def __init__(self, model: "TwoTowerModel", batch: int):
        torch = N.require_cuda()
        lib = N.load()
        self.batch = batch
        self.q = model.query_tower.workspace(batch)
        self.c = model.candidate_tower.workspace(batch)
        e = model.joint_embedding_size
        f32 = dict(dtype=torch.float32, device="cuda")
        self.col_bias = torch.zeros(batch, **f32) if model.logq_correction else None
        self.col_prob = torch.ones(batch, **f32) if model.logq_correction else None
        self.lse = torch.zeros(batch, **f32)
        self.loss = torch.zeros(1, **f32)
        self.bias_ready = torch.cuda.Event()
        self.dq = torch.zeros((batch, e), **f32)
        self.dc = torch.zeros((batch, e), **f32)
        self.sm_ws = torch.empty(int(lib.tt_softmax_workspace_bytes(batch, batch, e)), dtype=torch.uint8, device="cuda")
        self.bias_feat = None
        if model.logq_correction:
            cid_buf = None
            for (f, t, _, _), buf in zip(model.candidate_tower.input_layer.blocks, self.c.bufs):
                if f.name == model.candidate_id_col and t is not None:
                    cid_buf = buf
            self.cid_buf = cid_buf
            self.bias_feat = D.feature_array([{"table": model._logq_rows.data_ptr(), "src": cid_buf.data_ptr(),
                                               "rows": model._logq_rows.shape[0], "e": 1, "col": 0}])
        self.graph = None
        self.side = torch.cuda.Stream()
        self.cand = torch.cuda.Stream()               # candidate tower / dC pass, concurrent with the query side
        self.jobs = None
        self.sp_ws = None
        self.loss_host = torch.zeros(1, dtype=torch.float32).pin_memory()


class TwoTowerModel(AbstractKerasModel):
    def __init__(self, query_features: List[Feature], candidate_features: List[Feature], candidate_id_col: str,
                 joint_embedding_size: int, query_tower_units: Optional[List[int]] = None,
                 candidate_tower_units: Optional[List[int]] = None, candidate_prob_lookup: Optional[Dict[str, float]] = None):
        super().__init__()
        self.query_features = query_features
        self.candidate_features = candidate_features
        if candidate_id_col not in [f.name for f in candidate_features]:
            raise ValueError(f"candidate_id_col {candidate_id_col} not a candidate feature")
        self.candidate_id_col = candidate_id_col
        self.joint_embedding_size = int(joint_embedding_size)
        torch = N.require_cuda()

        def dense_count(feats, units):
            dim = 0
            last_e = {}
            for f in feats:
                if f.embedding_size and f.dtype.value == "string":
                    last_e[f.name] = int(f.embedding_size)
            for f in feats:
                dim += last_e[f.name] if f.dtype.value == "string" else 1
            return Tower.dense_param_count(dim, Tower.layer_sizes(units, joint_embedding_size))

        self._store = D.ParamStore(dense_count(query_features, query_tower_units) +
                                   dense_count(candidate_features, candidate_tower_units))
        self.query_tower = Tower(query_features, joint_embedding_size, query_tower_units, _store=self._store)
        self.candidate_tower = Tower(candidate_features, joint_embedding_size, candidate_tower_units, _store=self._store)
        if candidate_prob_lookup:
            self.logq_correction = LogQCorrection(candidate_prob_lookup)
            table = self.candidate_tower.input_layer.embedding_layers[candidate_id_col]
            rows_p = self.logq_correction.row_probabilities(table.vocab)  # probabilities per row
            self._logq_rows = torch.log(torch.from_numpy(rows_p).cuda())     # fp32 ln, as logq_correction.py:69
        else:
            self.logq_correction = None
            self._logq_rows = None
        self.optimizer = None
        self.loss_fn = None
        self.impl = N.TT_IMPL_AUTO           # contraction implementation for the B x B softmax
        self.use_cuda_graph = os.environ.get("TT_CUDA_GRAPH", "0") == "1"
        self.dist = None                     # set by pkg.modelling.distributed.DataParallel
        self._steps: Dict[int, _StepWorkspace] = {}
        self.phase_stamps = None
        self._opt_state = None
        self._loss_sum = 0.0
        self._loss_count = 0
        self.initialise_model()

    # ---- inference-style call ------------------------------------------------------------------------
    def call(self, x, training: bool = True):
        """(B x B) score of query i with candidate j -- materialised, for API parity (reference :65-92).
        No logQ correction here; train_step applies it (fused) exactly as the reference does."""
        torch = N.require_cuda()
        lib = N.load()
        qf = {f.name: x[f.name] for f in self.query_features}
        cf = {f.name: x[f.name] for f in self.candidate_features}
        q = self.query_tower(qf)
        c = self.candidate_tower(cf)
        e = self.joint_embedding_size
        out = torch.empty((q.shape[0], c.shape[0]), dtype=torch.float32, device="cuda")
        N.check(lib.tt_logits(q.data_ptr(), e, c.data_ptr(), e, None, q.shape[0], c.shape[0], e, out.data_ptr(), c.shape[0],
                              N.TT_IMPL_SIMT, N.stream_ptr()), "tt_logits")
        return out

    # ---- training ----------------------------------------------------------------------------------
    def compile(self, loss=None, optimizer=None, **_ignored) -> None:
        if loss is not None:
            if not isinstance(loss, CategoricalCrossentropy):
                raise NotImplementedError("loss must be pkg.modelling.losses.CategoricalCrossentropy(from_logits=True, reduction=SUM)")
            loss.validate()
        if optimizer is None or not isinstance(optimizer, (Adagrad, Adam)):
            raise ValueError("optimizer must come from OptimizerFactory.get_optimizer ('adagrad' or 'adam')")
        self.loss_fn = loss
        self.optimizer = optimizer
        self._build_optimizer_state()

    def _tables(self):
        out = []
        for tower in (self.query_tower, self.candidate_tower):
            for name, t in tower.input_layer.embedding_layers.items():
                out.append((tower, name, t))
        return out

    def _build_optimizer_state(self) -> None:
        torch = N.require_cuda()
        opt = self.optimizer
        init = opt.slot_init()
        # cached step workspaces (their sparse-optimizer job lists and captured CUDA graphs) hold raw pointers to the OLD slot tensors:
        # drop them, or a compile() / load() after training would update freed memory and never touch the new accumulators
        self._steps.clear()
        self._opt_state = {
            "dense": [torch.full_like(self._store.params, v) for v in init],
            # table slots take the shape of the table's local storage WITHOUT materialising a table nobody has used yet (a table
            # that is about to be row-sharded is initialised shard by shard, see EmbeddingTable)
            "tables": {id(t): [torch.full(t.local_shape, v, dtype=torch.float32, device="cuda") for v in init] for _, _, t in self._tables()},
        }

    def _step_ws(self, batch: int) -> _StepWorkspace:
        sw = self._steps.get(batch)
        if sw is None:
            if len(self._steps) >= 4:
                old = self._steps.pop(next(iter(self._steps)))
                for v in (getattr(old, "dp", None) or {}).values():     # data parallel: peer-shared buffers are released collectively
                    if hasattr(v, "close") and hasattr(v, "ptr_table"):
                        v.close()
            sw = self._steps[batch] = _StepWorkspace(self, batch)
            self._build_jobs(sw)
        return sw

    def _build_jobs(self, sw: _StepWorkspace) -> None:
        """Sparse-optimizer job list: one job per embedding table, one source per feature using it
        (or, data-parallel, per all-gathered block)."""
        torch = N.require_cuda()
        lib = N.load()
        srcs = []
        if self.dist is not None:   # data parallel: sources are the all-gathered (ids, gradient rows) of every rank
            srcs = self.dist.build_sparse_sources(self, sw)
        else:
            for tower, tws in ((self.query_tower, sw.q), (self.candidate_tower, sw.c)):
                for name, (table, lst) in tower.sparse_sources(tws).items():
                    srcs.append((table, lst))
        if len(srcs) > N.TT_MAX_JOBS:
            raise ValueError(f"at most {N.TT_MAX_JOBS} embedding tables per model")
        jobs = (N.TTSparseJob * len(srcs))()
        max_n = 0
        for j, (table, lst) in enumerate(srcs):
            if len(lst) > N.TT_MAX_SRC:
                raise ValueError(f"table {table.name}: more than {N.TT_MAX_SRC} features share it")
            slots = self._opt_state["tables"][id(table)]
            jobs[j].table = table.weight.data_ptr()
            jobs[j].slot0 = slots[0].data_ptr()
            jobs[j].slot1 = slots[1].data_ptr() if len(slots) > 1 else None
            jobs[j].rows, jobs[j].e, jobs[j].nsrc = table.rows, table.e, len(lst)
            jobs[j].shard_rank, jobs[j].shard_world = (table.shard_rank, table.shard_world) if table.shard_world > 1 else (0, 0)
            n_per = lst[0][0].numel()
            jobs[j].n_per_src = n_per
            for s, (ids, gptr, gld) in enumerate(lst):
                jobs[j].ids[s] = ids.data_ptr()
                jobs[j].grad[s] = gptr
                jobs[j].grad_ld[s] = gld
            max_n = max(max_n, n_per * len(lst))
        sw.jobs, sw.njobs = jobs, len(srcs)
        max_e = max(t.e for t, _ in srcs)
        sw.sp_ws = torch.empty(int(lib.tt_sparse_workspace_bytes(len(srcs), max_n, max_e)), dtype=torch.uint8, device="cuda")

    def _stage(self, sw: _StepWorkspace, data) -> None:
        qf = {f.name: data[f.name] for f in self.query_features}
        cf = {f.name: data[f.name] for f in self.candidate_features}
        stager = D.Stager(sw.batch, sw)       # both towers' columns: ONE staging launch (and one pinned block for host-resident columns)
        self.query_tower.input_layer.stage(qf, sw.q.bufs, stager)
        self.candidate_tower.input_layer.stage(cf, sw.c.bufs, stager)
        stager.flush()
        if self.logq_correction is not None and D.is_string_like(data[self.candidate_id_col]):
            # exact reference semantics for string ids: probability looked up by the STRING (an id outside
            # the vocabulary may still have a sampling probability); ln taken on the device below
            torch = N.require_cuda()
            p = self.logq_correction.probabilities(data[self.candidate_id_col])
            sw.col_prob.copy_(torch.from_numpy(p), non_blocking=True)
            sw.bias_from_strings = True
        else:
            sw.bias_from_strings = False

    # A step is four phases; data parallel runs the two exchanges eagerly and replays one CUDA graph for each compute phase:
    #   pre   (dist) all-gather of the batch ids            -- NCCL
    #   A     id sort || towers fwd, softmax fwd+bwd, towers bwd (three streams, forked and joined inside)
    #   mid   (dist) dense-gradient sum + all-gather of the embedding-gradient rows -- NCCL
    #   B     dense optimizer, de-duplicated sparse optimizer
    def _phase_pre(self, sw: _StepWorkspace) -> None:
        if self.dist is not None:
            self.dist.gather_ids(self, sw)

    def _phase_a(self, sw: _StepWorkspace) -> None:
        torch = N.require_cuda()
        lib = N.load()
        b, e = sw.batch, self.joint_embedding_size
        main = torch.cuda.current_stream()
        # fork: the ln p(candidate) gather and the id sort depend on the (gathered) ids only -- both leave the critical path
        sw.side.wait_stream(main)
        bias = None
        with torch.cuda.stream(sw.side):
            sts = N.stream_ptr()
            if self.logq_correction is not None:
                if sw.bias_from_strings:
                    N.check(lib.tt_log_f32(sw.col_prob.data_ptr(), sw.col_bias.data_ptr(), b, sts), "tt_log_f32")
                else:
                    N.check(lib.tt_gather_concat(sw.bias_feat, 1, b, 1, sw.col_bias.data_ptr(), 1, sts), "tt_gather_concat(logq)")
                bias = sw.col_bias.data_ptr()
                sw.bias_ready.record(sw.side)
            # The persistent softmax passes hold every SM and are statically partitioned: a kernel running beside them slows a few
            # SMs and with them the whole pass (measured: the id sort beside the softmax cost 25 us of a 206 us step).  A training
            # batch sorts in two radix passes (11-bit digits): pass 0 runs now, under the tower forward; the second one is queued
            # behind the softmax (below) and runs under the tower backward.
            # Larger sorts (batch x ranks > 16384 ids per table: 8-bit digits, three longer passes) stay here as a whole: beside a
            # multi-millisecond softmax the interference is a few per cent, and at 8 GPUs x 8192 the split was measured slower
            # (0.304 vs 0.294 ms per step) because two exposed passes cost more than the interference they avoid.
            split_sort = sw.batch * (1 if self.dist is None else self.dist.world) <= int(os.environ.get("TT_SPLIT_SORT_MAX", "16384"))
            N.check(lib.tt_sparse_sort_passes(sw.jobs, sw.njobs, sw.sp_ws.data_ptr(), sw.sp_ws.numel(), 0, 1 if split_sort else 1 << 30, sts),
                    "tt_sparse_sort_passes(0)")
        st = N.stream_ptr()
        # the two towers are independent until the logits: candidate side on its own stream
        sw.cand.wait_stream(main)
        q, q32 = self.query_tower.forward_ws(sw.q)
        with torch.cuda.stream(sw.cand):
            c, c32 = self.candidate_tower.forward_ws(sw.c)
        main.wait_stream(sw.cand)
        if bias is not None:
            main.wait_event(sw.bias_ready)
        use_tc = self.impl != N.TT_IMPL_SIMT and self._tc_ok()
        qa, ca = (q32, c32) if use_tc else (q, c)
        impl = N.TT_IMPL_TC if use_tc else N.TT_IMPL_SIMT
        # loss, lse, dQ and dC in one call: one prep launch, forward, combine, ONE persistent launch for the dQ and the dC
        # pass, reduction; then dQ -> query tower on the main stream, dC -> candidate tower on the candidate stream
        if self.dist is not None and self.dist.global_negatives:
            # cross-GPU in-batch negatives: my B query rows against the G.B candidates of all ranks (diagonal at rank.B); the
            # partial dC over all candidates is reduce-scattered back to the candidates' owners
            c_all, bias_all, off, bc = self.dist.gather_candidates(self, sw)
            dc_all = sw.dp["peer_dc"].local
            N.check(lib.tt_inbatch_softmax_step(qa.data_ptr(), e, c_all.data_ptr(), e, bias_all.data_ptr() if bias_all is not None else None,
                                                b, bc, e, off, sw.lse.data_ptr(), sw.loss.data_ptr(), sw.dq.data_ptr(), e, dc_all.data_ptr(), e,
                                                sw.sm_ws.data_ptr(), sw.sm_ws.numel(), impl, st), "tt_inbatch_softmax_step(global negatives)")
            self.dist.reduce_dc(self, sw)
        else:
            N.check(lib.tt_inbatch_softmax_step(qa.data_ptr(), e, ca.data_ptr(), e, bias, b, b, e, 0, sw.lse.data_ptr(), sw.loss.data_ptr(),
                                                sw.dq.data_ptr(), e, sw.dc.data_ptr(), e, sw.sm_ws.data_ptr(), sw.sm_ws.numel(), impl, st),
                    "tt_inbatch_softmax_step")
        sw.cand.wait_stream(main)
        if split_sort:
            sw.side.wait_stream(main)   # the softmax is done (stream order): the rest of the id sort overlaps the tower backward
            with torch.cuda.stream(sw.side):
                N.check(lib.tt_sparse_sort_passes(sw.jobs, sw.njobs, sw.sp_ws.data_ptr(), sw.sp_ws.numel(), 1, 1 << 30, N.stream_ptr()),
                        "tt_sparse_sort_passes(1..)")
        self.query_tower.backward_ws(sw.q, sw.dq)
        with torch.cuda.stream(sw.cand):
            self.candidate_tower.backward_ws(sw.c, sw.dc)
        main.wait_stream(sw.cand)
        main.wait_stream(sw.side)  # join
        if self.dist is not None:
            self.dist.pack_dense(self, sw)

    def _phase_mid(self, sw: _StepWorkspace) -> None:
        if self.dist is not None:
            self.dist.gather_rows(self, sw)

    def _phase_b(self, sw: _StepWorkspace) -> None:
        torch = N.require_cuda()
        lib = N.load()
        opt = self.optimizer
        n_dense = self._store.used
        # the dense half (sum over ranks + Dense-layer update) and the sparse half (embedding rows) touch different memory: two streams
        # when data parallel (on one GPU the dense half is one 2 us kernel: the fork / join costs more than it hides, measured)
        main = torch.cuda.current_stream()
        dense_stream = sw.cand if self.dist is not None else main
        if dense_stream is not main:
            dense_stream.wait_stream(main)
        with torch.cuda.stream(dense_stream):
            stc = N.stream_ptr()
            if self.dist is not None:
                self.dist.sum_dense(self, sw)
            if n_dense:
                ds = self._opt_state["dense"]
                if isinstance(opt, Adagrad):
                    N.check(lib.tt_dense_adagrad(self._store.params.data_ptr(), ds[0].data_ptr(), self._store.grads.data_ptr(), n_dense,
                                                 opt.learning_rate, opt.epsilon, stc), "tt_dense_adagrad")
                else:
                    N.check(lib.tt_dense_adam(self._store.params.data_ptr(), ds[0].data_ptr(), ds[1].data_ptr(),
                                              self._store.grads.data_ptr(), n_dense, sw.lr_t, opt.beta_1, opt.beta_2, opt.epsilon, stc),
                            "tt_dense_adam")
        st = N.stream_ptr()
        if isinstance(opt, Adagrad):
            N.check(lib.tt_sparse_adagrad(sw.jobs, sw.njobs, opt.learning_rate, opt.epsilon, sw.sp_ws.data_ptr(), sw.sp_ws.numel(), st),
                    "tt_sparse_adagrad")
        else:
            N.check(lib.tt_sparse_adam(sw.jobs, sw.njobs, sw.lr_t, opt.beta_1, opt.beta_2, opt.epsilon, sw.sp_ws.data_ptr(),
                                       sw.sp_ws.numel(), st), "tt_sparse_adam")
        if dense_stream is not main:
            main.wait_stream(dense_stream)

    def _launch_step(self, sw: _StepWorkspace) -> None:
        ring = self.phase_stamps            # measurement aid: (1 + ring_len * 5,) int64 device tensor, see tt_stamp
        if ring is None:
            self._phase_pre(sw)
            self._phase_a(sw)
            self._phase_mid(sw)
            self._phase_b(sw)
            return
        lib = N.load()
        n = (ring.numel() - 1) // 5
        for k, phase in enumerate((self._phase_pre, self._phase_a, self._phase_mid, self._phase_b)):
            N.check(lib.tt_stamp(ring.data_ptr(), n, k, 5, N.stream_ptr()), "tt_stamp")
            phase(sw)
        N.check(lib.tt_stamp(ring.data_ptr(), n, 4, 5, N.stream_ptr()), "tt_stamp")

    def _tc_ok(self) -> bool:
        return bool(N.load().tt_tc_available(0, self.joint_embedding_size))

    def train_step(self, data) -> Dict[str, object]:
        """One optimisation step on a batch {feature name: (B,1) column}.  Returns {"loss": 0-d device
        tensor holding this batch's SUM loss} (``float()`` it to synchronise)."""
        if self.optimizer is None:
            raise RuntimeError("call compile(loss=..., optimizer=...) before training")
        torch = N.require_cuda()
        batch = D.batch_size_of(data[self.candidate_id_col])
        sw = self._step_ws(batch)
        self._stage(sw, data)
        self.optimizer.iterations += 1
        graph_ok = self.use_cuda_graph and not isinstance(self.optimizer, Adam) and not sw.bias_from_strings
        if isinstance(self.optimizer, Adam):
            sw.lr_t = self.optimizer.lr_t(self.optimizer.iterations)
        if graph_ok:
            if sw.graph is None:
                self._launch_step(sw)  # eager warm-up (sets kernel attributes), also a real step
                torch.cuda.current_stream().synchronize()
                sw.graph = "pending"
            elif sw.graph == "pending":
                if self.dist is None or self.dist.peer_sync:   # the whole step is one graph (peer_sync: the ranks are ordered by device barriers inside it)
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        self._launch_step(sw)
                    sw.graph = (g,)
                    g.replay()
                else:                   # data parallel: the NCCL exchanges stay eager between two captured compute phases
                    self._phase_pre(sw)
                    ga = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(ga):
                        self._phase_a(sw)
                    ga.replay()
                    self._phase_mid(sw)
                    gb = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(gb):
                        self._phase_b(sw)
                    gb.replay()
                    sw.graph = (ga, gb)
            elif len(sw.graph) == 1:
                sw.graph[0].replay()
            else:
                self._phase_pre(sw)
                sw.graph[0].replay()
                self._phase_mid(sw)
                sw.graph[1].replay()
        else:
            self._launch_step(sw)
        return {"loss": sw.loss[0]}

    def fit(self, ds, epochs: int = 1, callbacks=None, verbose: int = 1):
        """Minimal Keras-like loop: ``ds`` is any iterable of batches (re-iterable for epochs > 1)."""
        history = {"loss": []}
        for epoch in range(epochs):
            total, steps, last = None, 0, None
            for batch in ds:
                last = self.train_step(batch)["loss"]
                total = last.clone() if total is None else total + last
                steps += 1
            mean = float(total) / steps if steps else float("nan")
            history["loss"].append(mean)
            if verbose:
                logger.info(f"epoch {epoch + 1}/{epochs}: {steps} steps, mean batch-sum loss {mean:.6f}")
            for cb in callbacks or []:
                if hasattr(cb, "on_epoch_end"):
                    cb.on_epoch_end(epoch, {"loss": mean})
        return history

    # ---- factory / signature / save ------------------------------------------------------------------
    @classmethod
    def create_from_schema(cls, schema: Schema, candidate_id_col: str) -> "TwoTowerModel":
        return TwoTowerModel(
            query_features=schema.query_features,
            candidate_features=schema.candidate_features,
            candidate_id_col=candidate_id_col,
            joint_embedding_size=schema.model_config.joint_embedding_size,
            query_tower_units=schema.model_config.query_tower_units,
            candidate_tower_units=schema.model_config.candidate_tower_units,
            candidate_prob_lookup=schema.training_config.candidate_prob_lookup,
        )

    def get_input_signature(self) -> Dict[str, TensorSpec]:
        return {f.name: TensorSpec((None, 1), f.dtype, f.name) for f in self.candidate_features + self.query_features}

    def state_arrays(self) -> Dict[str, np.ndarray]:
        arrs = self.query_tower.state_arrays("query_tower/")
        arrs.update(self.candidate_tower.state_arrays("candidate_tower/"))
        return arrs

    def load(self, model_path: str) -> "TwoTowerModel":
        """Restore the weights ``save(model_path)`` wrote (the two_tower/ directory next to ``model_path``, or a directory /
        variables.npz given directly).  The reference has no load path; optimizer slots restart fresh."""
        base = os.path.dirname(model_path)
        cand = [os.path.join(base, "two_tower"), model_path]
        path = next((c for c in cand if os.path.exists(os.path.join(c, "variables.npz")) or (c.endswith(".npz") and os.path.exists(c))), None)
        if path is None:
            raise FileNotFoundError(f"no saved two-tower variables under {cand}")
        arrs = Tower._read(path)
        self.query_tower.load_state_arrays(arrs, "query_tower/")
        self.candidate_tower.load_state_arrays(arrs, "candidate_tower/")
        if self.optimizer is not None:
            self._build_optimizer_state()
        return self

    def save(self, model_path: str) -> None:
        """two_tower/, query_tower/, candidate_tower/ next to ``model_path`` (reference :176-205), each holding
        variables.npz instead of a SavedModel."""
        base = os.path.dirname(model_path)
        os.makedirs(base or ".", exist_ok=True)
        for sub, arrs in (("two_tower", self.state_arrays()), ("query_tower", self.query_tower.state_arrays()),
                          ("candidate_tower", self.candidate_tower.state_arrays())):
            path = os.path.join(base, sub)
            logger.info(f"Saving {sub} at path: {path}")
            os.makedirs(path, exist_ok=True)
            np.savez(os.path.join(path, "variables.npz"), **arrs)
