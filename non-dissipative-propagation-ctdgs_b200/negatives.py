"""Negative supply for TGB one-vs-many evaluation.

The reference asks `NegativeEdgeSampler.query_batch(src, dst, t, split_mode)` once per 200-event batch: a
Python loop over a dict keyed by (src, dst, t) that returns a list of lists, which the evaluation loop then
turns into tensors one query at a time (TGB/tgb/linkproppred/negative_sampler.py:83-136, train_link.py:117-132).
Here the dict is packed ONCE into a CSR tensor indexed by event position, resident on the device; a batch of
negatives is then a slice (`rows(lo, hi)`), and the evaluation engine consumes the dense [n_events, n_neg]
view directly (`TGBEvalEngine(stream={..., "neg": packed.dense()})`)."""
from __future__ import annotations

import numpy as np
import torch


class PackedNegatives:
    def __init__(self, rowptr: torch.Tensor, neg: torch.Tensor):
        self.rowptr, self.neg = rowptr, neg          # [n+1] int64, [nnz] int64 (same device)

    @classmethod
    def from_eval_set(cls, eval_set: dict, src, dst, t, device=None) -> "PackedNegatives":
        """eval_set: {(src, dst, t): array of negative destinations}; src/dst/t: the split's events in order.
        Same error behaviour as query_batch: a positive edge missing from the set raises ValueError."""
        if eval_set is None:
            raise ValueError("Evaluation set is None! You should load the evaluation set first!")
        s, d, tt = (np.asarray(v.detach().cpu().numpy() if torch.is_tensor(v) else v) for v in (src, dst, t))
        counts = np.empty(s.shape[0] + 1, dtype=np.int64)
        counts[0] = 0
        chunks = []
        for k, key in enumerate(zip(s.tolist(), d.tolist(), tt.tolist())):
            row = eval_set.get(key)
            if row is None:
                raise ValueError(f"The edge ({key[0]}, {key[1]}, {key[2]}) is not in the evaluation set! "
                                 "Please check the implementation.")
            row = np.asarray(row, dtype=np.int64).reshape(-1)
            chunks.append(row)
            counts[k + 1] = row.shape[0]
        rowptr = torch.from_numpy(np.cumsum(counts))
        neg = torch.from_numpy(np.concatenate(chunks) if chunks else np.empty(0, dtype=np.int64))
        return cls(rowptr.to(device), neg.to(device))

    @property
    def num_events(self) -> int:
        return self.rowptr.numel() - 1

    def is_uniform(self) -> bool:
        c = self.rowptr[1:] - self.rowptr[:-1]
        return bool(c.numel() == 0 or (c == c[0]).all())

    def dense(self) -> torch.Tensor:
        """[n_events, n_neg] view for the evaluation engine (every query must have the same number of negatives)."""
        if not self.is_uniform():
            raise ValueError("ragged negative sets: the evaluation engine needs one negative count per split")
        n = self.num_events
        return self.neg.view(n, -1) if n else self.neg.view(0, 0)

    def rows(self, lo: int, hi: int):
        """(rowptr slice rebased to 0, negatives) of events [lo, hi) -- no host round trip."""
        a, b = int(self.rowptr[lo]), int(self.rowptr[hi])
        return self.rowptr[lo:hi + 1] - a, self.neg[a:b]

    def query_batch(self, lo: int, hi: int) -> list:
        """list-of-lists result of the reference's query_batch for events [lo, hi) (drop-in for its loop)."""
        rp = self.rowptr[lo:hi + 1].tolist()
        flat = self.neg[rp[0]:rp[-1]].tolist()
        return [flat[rp[i] - rp[0]:rp[i + 1] - rp[0]] for i in range(hi - lo)]
