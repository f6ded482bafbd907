"""
InputLayer: embed every string feature and concatenate, numeric features first, each group in schema
order (reference pkg/modelling/layers/input_layer.py:16-69).

Per string feature the reference builds StringLookup(num_oov_indices=1) -> Embedding(len(vocab)+1, e)
-> Reshape.  Here the lookup is a host dictionary (strings never reach the GPU; integer inputs are taken
as row ids directly), the tables live in HBM as (V+1, e) fp32 row-major, and gather + concat is one CUDA
kernel (tt_gather_concat) -- or is fused into the first Dense layer by Tower (tt_input_dense_fwd).
As in the reference, the table dict is keyed by feature *name*: two features with the same name share
the table built last and are both concatenated (input_layer.py:30-43).
"""
from __future__ import annotations

from typing import Dict, List

import numpy as np

from pkg import _native as N
from pkg.modelling import _device as D
from pkg.schema.dtypes import DType
from pkg.schema.features import Feature


class EmbeddingTable:
    """Embedding(len(vocab)+1, e) with tf-keras' default RandomUniform(-0.05, 0.05) initialiser.

    The rows are materialised on first use, by a counter-based generator (tt_fill_uniform): a cell's value depends on
    (seed, row, column) only.  A table that is row-sharded BEFORE its first use (DataParallel(shard_tables=True) attached to a
    fresh model) is therefore initialised shard by shard on the owning GPUs -- no rank ever allocates, fills or broadcasts the
    whole table -- and holds exactly the values the unsharded table would."""

    def __init__(self, feature: Feature):
        torch = N.require_cuda()
        if feature.vocab is None:
            raise ValueError(f"feature {feature.name} has no vocabulary; build the schema first")
        if not feature.embedding_size:
            raise ValueError(f"string feature {feature.name} needs an embedding_size")
        self.name = feature.name
        self.vocab = D.Vocab(feature.vocab)
        self.e = int(feature.embedding_size)
        self.rows = self.vocab.rows
        self._weight = None
        self.seed = D.next_seed()
        # row sharding over the ranks of a process group (pkg.modelling.distributed.DataParallel(shard_tables=True))
        self.shard_rank, self.shard_world, self._peer, self._group = 0, 1, None, None

    # ---- storage -------------------------------------------------------------------------------------
    @property
    def materialised(self) -> bool:
        return self._weight is not None

    @property
    def weight(self):
        """(rows, e) fp32 rows in HBM -- this rank's (local_rows, e) shard once the table is row-sharded."""
        if self._weight is None:
            torch = N.require_cuda()
            self._weight = torch.empty((self.rows, self.e), dtype=torch.float32, device="cuda")
            self._fill(self._weight, self.rows, 0, 1)
        return self._weight

    @weight.setter
    def weight(self, value) -> None:
        self._weight = value

    @property
    def local_shape(self):
        """Shape of ``weight`` without materialising it (optimizer slots take this shape)."""
        return (self.rows, self.e) if self.shard_world == 1 else (self.local_rows, self.e)

    def _fill(self, out, n_rows: int, row0: int, row_stride: int) -> None:
        N.check(N.load().tt_fill_uniform(out.data_ptr(), n_rows, self.e, row0, row_stride, self.seed, -0.05, 0.05, N.stream_ptr()),
                "tt_fill_uniform")

    # ---- row sharding (one process per GPU) ---------------------------------------------------------
    @property
    def local_rows(self) -> int:
        return (self.rows + self.shard_world - 1) // self.shard_world

    def shard_rows(self, group=None) -> None:
        """Collective.  Keep rows {i : i % G == rank} (local row i // G) in a peer-shareable buffer; every rank's kernels read
        the other shards directly over NVLink.  The interleaved split spreads the hot head of a frequency-ordered vocabulary
        (features.py:119-127) evenly over the GPUs."""
        import torch.distributed as dist

        from pkg.modelling._peer import PeerBuffer

        if self.shard_world > 1:
            return
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        if world == 1:
            return
        full = self._weight
        self.shard_rank, self.shard_world, self._group = rank, world, group
        self._peer = PeerBuffer((self.local_rows, self.e), "float32", group)
        if full is None:       # never used: the owner initialises its rows {rank, rank + world, ...} in place
            self._fill(self._peer.local, (self.rows - rank + world - 1) // world, rank, world)
        else:
            mine = full[rank::world]
            self._peer.local[: mine.shape[0]].copy_(mine)
        self._weight = self._peer.local

    def table_pointer(self) -> int:
        """What tt_feature.table holds: the table itself, or the device array of shard base pointers."""
        return self.weight.data_ptr() if self.shard_world == 1 else self._peer.ptr_table.data_ptr()

    def full_weight(self):
        """(rows, e) table; collective when the table is row-sharded (all-gather of the shards)."""
        if self.shard_world == 1:
            return self.weight
        import torch.distributed as dist

        torch = N.require_cuda()
        g = self.shard_world
        shards = torch.empty((g,) + tuple(self.weight.shape), dtype=torch.float32, device="cuda")
        dist.all_gather_into_tensor(shards, self.weight.contiguous(), group=self._group)
        full = torch.empty((self.rows, self.e), dtype=torch.float32, device="cuda")
        for r in range(g):
            n_r = (self.rows - r + g - 1) // g
            full[r::g] = shards[r, :n_r]
        return full


class InputLayer:
    def __init__(self, features: List[Feature]):
        self.numerical_features = [f for f in features if f.dtype != DType.string]
        self.categorical_features = [f for f in features if f.dtype == DType.string]
        if len(features) > N.TT_MAX_FEATURES:
            raise ValueError(f"at most {N.TT_MAX_FEATURES} features per tower")
        self._init_embedding_layers()
        # column layout of the concatenated output
        self.blocks = []  # (feature, table or None, first column, width)
        col = 0
        for f in self.numerical_features:
            self.blocks.append((f, None, col, 1))
            col += 1
        for f in self.categorical_features:
            t = self.embedding_layers[f.name]
            self.blocks.append((f, t, col, t.e))
            col += t.e
        self.output_dim = col
        self.ld = D.ParamStore.padded(col)

    def _init_embedding_layers(self) -> None:
        self.embedding_layers: Dict[str, EmbeddingTable] = {}
        for f in self.categorical_features:
            self.embedding_layers[f.name] = EmbeddingTable(f)

    # ---- staging / descriptors -------------------------------------------------------------------
    def new_buffers(self, batch: int):
        """Device buffers for one batch: int32 ids per string feature *occurrence*, fp32 per numeric."""
        torch = N.require_cuda()
        bufs = []
        for f, t, _, _ in self.blocks:
            dt = torch.float32 if t is None else torch.int32
            bufs.append(torch.zeros(batch, dtype=dt, device="cuda"))
        return bufs

    def stage(self, x: Dict[str, object], bufs, stager=None) -> None:
        """Stage one batch into ``bufs``: every column in one tt_stage_columns launch.  ``stager``: a caller's Stager that collects
        the columns of several input layers (the train step stages both towers with one launch) and is flushed by the caller."""
        own = stager is None
        if own:
            stager = D.Stager(int(bufs[0].shape[0]), self)
        for (f, t, _, _), buf in zip(self.blocks, bufs):
            if f.name not in x:
                raise KeyError(f"input is missing feature {f.name!r}")
            if t is None:
                stager.add_floats(x[f.name], buf)
            else:
                stager.add_ids(x[f.name], t.vocab, buf)
        if own:
            stager.flush()

    def descriptors(self, bufs):
        return D.feature_array(
            {"table": None if t is None else t.table_pointer(), "src": buf.data_ptr(), "rows": 0 if t is None else t.rows,
             "e": w, "col": c, "shards": 0 if t is None or t.shard_world == 1 else t.shard_world}
            for (f, t, c, w), buf in zip(self.blocks, bufs))

    def sharded(self) -> bool:
        return any(t.shard_world > 1 for t in self.embedding_layers.values())

    def batch_size(self, x: Dict[str, object]) -> int:
        return D.batch_size_of(x[self.blocks[0][0].name])

    def call(self, x: Dict[str, object]):
        """{name: (B,1) strings | row ids | floats} -> (B, D) fp32 device tensor."""
        torch = N.require_cuda()
        lib = N.load()
        batch = self.batch_size(x)
        bufs = self.new_buffers(batch)
        self.stage(x, bufs)
        out = torch.empty((batch, self.ld), dtype=torch.float32, device="cuda")
        feats = self.descriptors(bufs)
        N.check(lib.tt_gather_concat(feats, len(self.blocks), batch, self.output_dim, out.data_ptr(), self.ld, N.stream_ptr()),
                "tt_gather_concat")
        return out[:, :self.output_dim]

    __call__ = call

    def tables(self) -> Dict[str, EmbeddingTable]:
        return self.embedding_layers

    def state_arrays(self, prefix: str = "") -> Dict[str, np.ndarray]:
        return {f"{prefix}embedding/{n}": t.full_weight().detach().cpu().numpy() for n, t in self.embedding_layers.items()}

    def load_state_arrays(self, arrs, prefix: str = "") -> None:
        """Inverse of state_arrays (row-sharded tables keep their own rows)."""
        torch = N.require_cuda()
        for n, t in self.embedding_layers.items():
            a = np.asarray(arrs[f"{prefix}embedding/{n}"], dtype=np.float32)
            if a.shape != (t.rows, t.e):
                raise ValueError(f"embedding {n}: saved shape {a.shape} != {(t.rows, t.e)}")
            if t.shard_world > 1:
                a = a[t.shard_rank::t.shard_world]
            t.weight[: a.shape[0]].copy_(torch.from_numpy(np.ascontiguousarray(a)))
