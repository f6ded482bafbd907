"""
Tower: InputLayer -> [Dense(u, relu)]* -> Dense(joint, relu) (reference pkg/modelling/models/tower.py:9-91;
note the ReLU on the LAST layer too, so tower outputs are non-negative).

Forward: the first Dense is fused with the embedding gather (tt_input_dense_fwd), later layers use
tt_dense_fwd; all run as exact fp32 FMA in canonical k order, so tower outputs match the oracle bit for
bit.  The final activation is also emitted rounded to TF32 for the tensor-core logits / index kernels.
Backward (tt_dense_bwd): dW, db reduced over the batch in a fixed order; dX of the first layer is the
row gradient of the embedding tables, consumed in place by the sparse optimizer.
"""
from __future__ import annotations

import math
from typing import Dict, List, Optional

import numpy as np

from pkg import _native as N
from pkg.modelling import _device as D
from pkg.modelling.layers.input_layer import InputLayer
from pkg.modelling.models.abstract_keras_model import AbstractKerasModel, TensorSpec
from pkg.schema.features import Feature


class _TowerWorkspace:
    """Device buffers of one tower for one batch size."""

    def __init__(self, tower: "Tower", batch: int):
        torch = N.require_cuda()
        lib = N.load()
        il = tower.input_layer
        self.batch = batch
        self.bufs = il.new_buffers(batch)
        self.feats = il.descriptors(self.bufs)
        f32 = dict(dtype=torch.float32, device="cuda")
        self.x = torch.zeros((batch, il.ld), **f32)
        self.dx = torch.zeros((batch, il.ld), **f32)
        self.acts = [torch.zeros((batch, n), **f32) for n in tower.layer_units]
        self.d_acts = [torch.zeros((batch, n), **f32) for n in tower.layer_units[:-1]]
        self.out_tf32 = torch.zeros((batch, tower.layer_units[-1]), **f32)
        need = 256
        k = il.output_dim
        for n in tower.layer_units:
            need = max(need, int(lib.tt_dense_bwd_workspace_bytes(batch, k, n)))
            k = n
        self.bwd_ws = torch.empty(need, dtype=torch.uint8, device="cuda")


class Tower(AbstractKerasModel):
    def __init__(self, features: List[Feature], joint_embedding_size: int, hidden_units: Optional[List[int]] = None,
                 _store: Optional[D.ParamStore] = None):
        super().__init__()
        self.features = features
        self.joint_embedding_size = int(joint_embedding_size)
        self.hidden_units = hidden_units
        self._store = _store
        self._ws: Dict[int, _TowerWorkspace] = {}
        self._init_layers()
        self.initialise_model()

    # ---- construction -----------------------------------------------------------------------------
    @staticmethod
    def layer_sizes(hidden_units: Optional[List[int]], joint: int) -> List[int]:
        return [int(u) for u in (hidden_units or [])] + [int(joint)]

    @staticmethod
    def dense_param_count(input_dim: int, units: List[int]) -> int:
        total, k = 0, input_dim
        for n in units:
            total += D.ParamStore.padded(k * n) + D.ParamStore.padded(n)
            k = n
        return total

    def _init_layers(self) -> None:
        torch = N.require_cuda()
        self.input_layer = InputLayer(self.features)
        self.layer_units = self.layer_sizes(self.hidden_units, self.joint_embedding_size)
        if self._store is None:
            self._store = D.ParamStore(self.dense_param_count(self.input_layer.output_dim, self.layer_units))
        self.kernels, self.biases, self.kernel_grads, self.bias_grads = [], [], [], []
        k = self.input_layer.output_dim
        for n in self.layer_units:
            w, gw = self._store.alloc((k, n))
            b, gb = self._store.alloc((n,))
            limit = math.sqrt(6.0 / (k + n))  # glorot_uniform, tf-keras Dense default
            w.uniform_(-limit, limit, generator=D.next_generator())
            b.zero_()
            self.kernels.append(w); self.biases.append(b); self.kernel_grads.append(gw); self.bias_grads.append(gb)
            k = n
        # model_layers mirrors the reference attribute: [InputLayer, Dense, ..., Dense]
        self.model_layers = [self.input_layer] + [("dense", i) for i in range(len(self.layer_units))]

    # ---- engine-facing ----------------------------------------------------------------------------
    def workspace(self, batch: int) -> _TowerWorkspace:
        ws = self._ws.get(batch)
        if ws is None:
            if len(self._ws) >= 4:  # keep at most a few batch shapes alive (ragged last batch etc.)
                self._ws.pop(next(iter(self._ws)))
            ws = self._ws[batch] = _TowerWorkspace(self, batch)
        return ws

    def forward_ws(self, ws: _TowerWorkspace, keep_input: bool = True):
        """Runs the tower on already staged inputs; returns (out fp32, out rounded to TF32)."""
        lib = N.load()
        st = N.stream_ptr()
        il = self.input_layer
        last = len(self.layer_units) - 1
        n0 = self.layer_units[0]
        # the 64-unit first layer has a fused gather + Dense kernel whose gather issues all loads of a round before the first store
        # (csrc/tt_tower_panel.cu dense_fwd_fused64_kernel): one NVLink round trip per round also for rows in other GPUs' HBM
        fused64 = n0 == 64 and il.output_dim <= 96 and all(t is None or t.e % 4 == 0 for _, t, _, _ in il.blocks)
        if il.sharded() and not fused64:
            # row-sharded tables: most rows live in other GPUs' HBM.  A dedicated gather with every 16-byte load independent pays
            # the NVLink latency once; the first Dense then reads the local copy (bit-identical to the fused kernel).
            N.check(lib.tt_gather_concat(ws.feats, len(il.blocks), ws.batch, il.output_dim, ws.x.data_ptr(), il.ld, st), "tt_gather_concat")
            N.check(lib.tt_dense_fwd(ws.x.data_ptr(), il.ld, self.kernels[0].data_ptr(), self.biases[0].data_ptr(), ws.acts[0].data_ptr(), n0,
                                     ws.out_tf32.data_ptr() if last == 0 else None, ws.batch, il.output_dim, n0, 1, st), "tt_dense_fwd")
        else:
            N.check(lib.tt_input_dense_fwd(ws.feats, len(il.blocks), il.output_dim, self.kernels[0].data_ptr(),
                                           self.biases[0].data_ptr(), ws.x.data_ptr() if keep_input else None, il.ld,
                                           ws.acts[0].data_ptr(), n0, ws.out_tf32.data_ptr() if last == 0 else None,
                                           ws.batch, n0, 1, st), "tt_input_dense_fwd")
        k = n0
        for i in range(1, len(self.layer_units)):
            n = self.layer_units[i]
            N.check(lib.tt_dense_fwd(ws.acts[i - 1].data_ptr(), k, self.kernels[i].data_ptr(), self.biases[i].data_ptr(),
                                     ws.acts[i].data_ptr(), n, ws.out_tf32.data_ptr() if i == last else None, ws.batch, k, n, 1,
                                     st), "tt_dense_fwd")
            k = n
        return ws.acts[-1], ws.out_tf32

    def backward_ws(self, ws: _TowerWorkspace, d_out) -> None:
        """Fills kernel_grads / bias_grads and ws.dx from d(loss)/d(tower output)."""
        lib = N.load()
        st = N.stream_ptr()
        il = self.input_layer
        dy, lddy = d_out, d_out.stride(0)
        for i in range(len(self.layer_units) - 1, -1, -1):
            n = self.layer_units[i]
            if i == 0:
                xin, ldx, k, dx, lddx = ws.x, il.ld, il.output_dim, ws.dx, il.ld
            else:
                k = self.layer_units[i - 1]
                xin, ldx, dx, lddx = ws.acts[i - 1], k, ws.d_acts[i - 1], k
            N.check(lib.tt_dense_bwd(xin.data_ptr(), ldx, self.kernels[i].data_ptr(), ws.acts[i].data_ptr(), n, dy.data_ptr(),
                                     lddy, dx.data_ptr(), lddx, self.kernel_grads[i].data_ptr(), self.bias_grads[i].data_ptr(),
                                     ws.batch, k, n, 1, ws.bwd_ws.data_ptr(), ws.bwd_ws.numel(), st), "tt_dense_bwd")
            dy, lddy = dx, lddx

    def sparse_sources(self, ws: _TowerWorkspace):
        """table name -> (EmbeddingTable, [(ids buffer, grad pointer, grad ld)]) for the sparse optimizer."""
        out = {}
        il = self.input_layer
        for (f, t, col, w), buf in zip(il.blocks, ws.bufs):
            if t is None:
                continue
            out.setdefault(f.name, (t, []))[1].append((buf, ws.dx.data_ptr() + 4 * col, il.ld))
        return out

    # ---- public API -------------------------------------------------------------------------------
    def call(self, x, training: bool = True):
        """{name: (B,1) column} -> (B, E) fp32 device tensor (a fresh tensor; ``training`` is ignored as in
        the reference, tower.py:51-75)."""
        batch = self.input_layer.batch_size(x)
        ws = self.workspace(batch)
        self.input_layer.stage(x, ws.bufs)
        out, _ = self.forward_ws(ws, keep_input=False)
        return out.clone()

    def embed_tf32(self, x):
        """(fp32 output, TF32-rounded copy) without cloning -- used by BruteForceIndex."""
        batch = self.input_layer.batch_size(x)
        ws = self.workspace(batch)
        self.input_layer.stage(x, ws.bufs)
        return self.forward_ws(ws, keep_input=False)

    def get_input_signature(self) -> Dict[str, TensorSpec]:
        return {f.name: TensorSpec((None, 1), f.dtype, f.name) for f in self.features}

    def load_state_arrays(self, arrs, prefix: str = "") -> None:
        """Inverse of state_arrays: restores tables, kernels and biases (shapes must match this tower's schema)."""
        torch = N.require_cuda()
        self.input_layer.load_state_arrays(arrs, prefix)
        for i, (w, b) in enumerate(zip(self.kernels, self.biases)):
            kw, kb = np.asarray(arrs[f"{prefix}dense_{i}/kernel"], np.float32), np.asarray(arrs[f"{prefix}dense_{i}/bias"], np.float32)
            if kw.shape != tuple(w.shape) or kb.shape != tuple(b.shape):
                raise ValueError(f"dense_{i}: saved shapes {kw.shape}/{kb.shape} != {tuple(w.shape)}/{tuple(b.shape)}")
            w.copy_(torch.from_numpy(kw)); b.copy_(torch.from_numpy(kb))

    @staticmethod
    def _read(path: str):
        import os

        f = path if path.endswith(".npz") else os.path.join(path, "variables.npz")
        with np.load(f) as z:
            return {k: z[k] for k in z.files}

    def load(self, model_path: str) -> "Tower":
        """Restore what ``save`` wrote (query_tower/ or candidate_tower/ directory, or its variables.npz)."""
        self.load_state_arrays(self._read(model_path))
        return self

    def state_arrays(self, prefix: str = "") -> Dict[str, np.ndarray]:
        arrs = self.input_layer.state_arrays(prefix)
        for i, (w, b) in enumerate(zip(self.kernels, self.biases)):
            arrs[f"{prefix}dense_{i}/kernel"] = w.detach().cpu().numpy()
            arrs[f"{prefix}dense_{i}/bias"] = b.detach().cpu().numpy()
        return arrs
