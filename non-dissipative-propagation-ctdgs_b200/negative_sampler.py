"""Drop-in `NegativeSampler` (negative_sampler.py:8-48) that draws on the device (csrc/sampler.cu).

Same constructor and `sample(src, eval=False, eval_seed=9)` contract: one negative destination per source, uniform
over the known destination nodes, redrawn (at most `nattemps` more times) while (src, dst) is an edge given at
construction.  `eval=True` is deterministic in `eval_seed` and -- like upstream, which re-seeds a fresh generator on
every call -- returns the same draws for the same `src` length on every call.  The numpy PCG64 stream of the
reference is not reproduced; the distribution and the no-collision guarantee are."""
from __future__ import annotations

from typing import Iterable

import torch

from . import _lib
from ._lib import call, ptr

neg_sampler_names = ["NegativeSampler"]


class NegativeSampler:
    def __init__(self, src_nodes: Iterable, dst_nodes: Iterable, name: str, seed: int = 9,
                 check_link_existence: bool = True, device=None) -> None:
        src_nodes, dst_nodes = torch.as_tensor(src_nodes), torch.as_tensor(dst_nodes)
        device = torch.device(device if device is not None else (src_nodes.device if src_nodes.is_cuda else "cuda"))
        if device.type != "cuda":
            raise _lib.CtanLibraryError("ctan_b200.NegativeSampler draws on a CUDA device (no CPU path)")
        _lib.lib()
        s, d = src_nodes.to(device, torch.long), dst_nodes.to(device, torch.long)
        self.src_nodes, self.dst_nodes = s.unique(), d.unique()         # sorted unique, as upstream (kept on the device)
        self.seen = torch.unique((s << 32) | d) if check_link_existence else None
        self.seed, self.name, self.check_link_existence, self.nattemps = seed, name, check_link_existence, 500
        self.device = device
        self._calls = 0
        self._failed = torch.zeros(1, dtype=torch.int32, device=device)

    def sample(self, src: torch.Tensor, eval: bool = False, eval_seed: int = 9, *args, **kwargs) -> torch.Tensor:
        src_d = src.to(self.device, torch.long).contiguous()
        n = src_d.numel()
        out = torch.empty(n, dtype=torch.long, device=self.device)
        if n == 0:
            return out.to(src.device)
        seed, ncall = (int(eval_seed), 0) if eval else (int(self.seed), self._calls + 1)
        if not eval:
            self._calls += 1
        with torch.cuda.device(self.device):
            call("ctan_neg_sample", ptr(self.dst_nodes), self.dst_nodes.numel(),
                 ptr(self.seen) if self.seen is not None else None, self.seen.numel() if self.seen is not None else 0,
                 ptr(src_d), n, seed & (2 ** 64 - 1), ncall, self.nattemps, ptr(out), ptr(self._failed))
        return out.to(src.device)

    def failed_draws(self) -> int:
        """Elements for which no unseen destination was found within the attempt cap (the reference prints a line
        per such element); reading it synchronises."""
        return int(self._failed.item())

    def __repr__(self) -> str:
        return f"{self.__class__.__name__}({self.name}, seed={self.seed})"
