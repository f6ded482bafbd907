"""
LogQCorrection: logits - ln p(candidate_j), broadcast over rows (reference
pkg/modelling/layers/logq_correction.py:16-71).  Unknown ids get p = 1.0, i.e. no correction.
The string -> probability lookup is host-side; ln and the subtraction run on the GPU (tt_log_f32,
tt_logq_apply).  Inside TwoTowerModel.train_step the correction is fused into the softmax kernel's
epilogue instead and the B x B matrix is never formed.
"""
from __future__ import annotations

from typing import Dict

import numpy as np

from pkg import _native as N
from pkg.modelling import _device as D


class LogQCorrection:
    def __init__(self, candidate_prob_lookup: Dict[str, float]):
        self._init_lookup(candidate_prob_lookup)

    def _init_lookup(self, candidate_prob_lookup: Dict[str, float]) -> None:
        # StaticHashTable[str -> float32], default 1.0
        self.lookup = {D._as_str(k): np.float32(v) for k, v in candidate_prob_lookup.items()}
        self.default_value = np.float32(1.0)

    def probabilities(self, candidate_ids) -> np.ndarray:
        flat = np.asarray(D.unwrap(candidate_ids)).reshape(-1)
        return np.fromiter((self.lookup.get(D._as_str(v), self.default_value) for v in flat), dtype=np.float32,
                           count=flat.shape[0])

    def row_probabilities(self, vocab: "D.Vocab") -> np.ndarray:
        """p for every embedding row of the candidate-id table (row 0 = OOV -> 1.0, i.e. ln p = 0); the fused training path,
        where ids are already row numbers, takes the fp32 ln of this table once on the device."""
        p = np.ones(vocab.rows, dtype=np.float32)
        if getattr(vocab, "_range", False):
            for k, v in self.lookup.items():
                if k.isdigit() and 1 <= int(k) <= vocab.size:
                    p[int(k)] = v
        else:
            keys = list(self.lookup.keys())
            rows = vocab.encode(np.array(keys, dtype=object)) if keys else []
            for k, row in zip(keys, rows):
                if row > 0:                      # ids with a probability but outside the vocabulary share the OOV row (ln 1 = 0)
                    p[row] = self.lookup[k]
        return p

    def __call__(self, logits, candidate_ids):
        """logits: (B, B) device tensor (or array); candidate_ids: (B, 1) strings.  Returns a device tensor."""
        torch = N.require_cuda()
        lib = N.load()
        z_in = logits if isinstance(logits, torch.Tensor) else torch.from_numpy(np.asarray(logits, dtype=np.float32))
        z_in = z_in.to(device="cuda", dtype=torch.float32).contiguous()
        bq, bc = z_in.shape
        probs = torch.from_numpy(self.probabilities(candidate_ids)).cuda()
        if probs.numel() != bc:
            raise ValueError(f"candidate_ids has {probs.numel()} entries, logits has {bc} columns")
        corr = torch.empty_like(probs)
        out = torch.empty_like(z_in)
        st = N.stream_ptr()
        N.check(lib.tt_log_f32(probs.data_ptr(), corr.data_ptr(), bc, st), "tt_log_f32")
        N.check(lib.tt_logq_apply(z_in.data_ptr(), bc, corr.data_ptr(), bq, bc, out.data_ptr(), bc, st), "tt_logq_apply")
        return out
