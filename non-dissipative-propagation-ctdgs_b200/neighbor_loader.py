"""Drop-in `LastNeighborLoader` backed by the sm_100a ring-buffer kernels (csrc/loader.cu).

Mirrors PyG's / TGB's class (reference: TGB/modules/neighbor_loader.py:15-85; imported by
train_link.py:4): same constructor, `__call__(n_id) -> (n_id, edge_index, e_id)`, `insert(src, dst)`,
`reset_state()`, and the attributes callers read (`e_id`, `neighbors`, `cur_e_id`, `size`).

Differences that a caller can observe:
  * results follow the reference's CPU/stable-sort semantics (the reference itself is
    implementation-defined when one node gets more than `size` events in one batch);
  * `neighbors` slots whose `e_id` is -1 hold 0 / stale ids (the reference leaves them uninitialised);
  * `__call__` expects a sorted-unique `n_id` (what every call site passes: `.unique()` output or a
    previous call's result);
  * there is no CPU path: tensors must live on a CUDA device.
Extra, non-reference API: `lookup_khop(seeds, K)` does the K chained calls of train_link.py:258-262 in
one device pass with a single host synchronisation, and also returns the CSR row pointer."""
from __future__ import annotations

import weakref

import torch

from . import _lib
from ._lib import call, ptr

_CSR_CACHE: list = []          # [(weakref(edge_index), data_ptr, E, N, rowptr)] most recent last


def _remember_csr(edge_index: torch.Tensor, n_nodes: int, rowptr: torch.Tensor):
    _CSR_CACHE.append((weakref.ref(edge_index), edge_index.data_ptr(), edge_index.size(1), n_nodes, rowptr))
    del _CSR_CACHE[:-8]


def cached_rowptr(edge_index: torch.Tensor, n_nodes: int):
    """rowptr of an edge_index tensor produced by our loader (None if unknown)."""
    for ref, dp, e, n, rowptr in reversed(_CSR_CACHE):
        if ref() is edge_index and dp == edge_index.data_ptr() and e == edge_index.size(1) and n == n_nodes:
            return rowptr
    return None


class LastNeighborLoader:
    def __init__(self, num_nodes: int, size: int, device=None):
        device = torch.device(device if device is not None else "cuda")
        if device.type != "cuda":
            raise _lib.CtanLibraryError("ctan_b200.LastNeighborLoader needs a CUDA device (no CPU path)")
        _lib.lib()
        self.size = int(size)
        self.num_nodes = int(num_nodes)
        self.device = device
        self.neighbors = torch.empty((num_nodes, size), dtype=torch.long, device=device)
        self.e_id = torch.empty((num_nodes, size), dtype=torch.long, device=device)
        self._assoc = torch.empty(num_nodes, dtype=torch.long, device=device)
        self._deg = torch.zeros(num_nodes, dtype=torch.int32, device=device)
        n_words = (num_nodes + 31) // 32
        self._bitmap = torch.zeros(n_words, dtype=torch.int32, device=device)
        self._cbitmap = torch.zeros(n_words, dtype=torch.int32, device=device)
        self._counts = torch.zeros(8, dtype=torch.int32, device=device)
        self._counts_host = torch.zeros(8, dtype=torch.int32).pin_memory()
        self._frontier_cnt = torch.zeros(8, dtype=torch.int32, device=device)
        self.reset_state()

    # ------------------------------------------------------------------ reference API
    def reset_state(self):
        self.cur_e_id = 0
        with torch.cuda.device(self.device):
            call("ctan_nbr_reset", ptr(self.neighbors), ptr(self.e_id), ptr(self._deg), self.num_nodes, self.size)

    def insert(self, src: torch.Tensor, dst: torch.Tensor):
        B = src.numel()
        if B == 0:
            return
        src = src.contiguous()
        dst = dst.contiguous()
        ws = torch.empty(4 * B, dtype=torch.int32, device=self.device)
        with torch.cuda.device(self.device):
            call("ctan_nbr_insert", ptr(self.neighbors), ptr(self.e_id), ptr(self._deg), ptr(src, torch.long),
                 ptr(dst, torch.long), B, self.cur_e_id, None, self.size, ptr(ws))
        self.cur_e_id += B

    def __call__(self, n_id: torch.Tensor):
        n_out, edge_index, e_id, _ = self.lookup_khop(n_id, 1)
        return n_out, edge_index, e_id

    # ------------------------------------------------------------------ fused K-hop lookup
    def lookup_khop(self, seeds: torch.Tensor, num_hops: int):
        """Equivalent of `for _ in range(K): n_id, edge_index, e_id = loader(n_id)` starting from
        `seeds` (duplicates allowed: the seed set is de-duplicated on the device).  Returns
        (n_id [N] sorted unique, edge_index [2,E], e_id [E], rowptr [N+1] int32)."""
        if num_hops < 1:
            raise ValueError("num_hops must be >= 1")
        if num_hops > 8:
            raise ValueError("num_hops > 8 is not supported")
        S = self.size
        seeds = seeds.contiguous()
        n_seeds = seeds.numel()
        dev = self.device
        if n_seeds == 0:
            z = torch.empty(0, dtype=torch.long, device=dev)
            return z, torch.empty((2, 0), dtype=torch.long, device=dev), z.clone(), torch.zeros(1, dtype=torch.int32, device=dev)
        n_cap = min(self.num_nodes, n_seeds * (S + 1) ** num_hops)
        c_cap = min(self.num_nodes, n_seeds * (S + 1) ** (num_hops - 1))
        e_cap = c_cap * S
        n_out = torch.empty(n_cap, dtype=torch.long, device=dev)
        rowptr = torch.empty(n_cap + 1, dtype=torch.int32, device=dev)
        frontier = torch.empty(max(num_hops - 1, 1) * n_cap, dtype=torch.long, device=dev)
        with torch.cuda.device(dev):
            call("ctan_nbr_lookup", ptr(self.neighbors), ptr(self._deg), self.num_nodes, S, ptr(seeds, torch.long),
                 n_seeds, None, num_hops, ptr(self._bitmap), ptr(self._cbitmap), ptr(frontier),
                 ptr(self._frontier_cnt), n_cap, ptr(n_out), n_cap, ptr(self._assoc), ptr(rowptr), e_cap,
                 ptr(self._counts))
            self._counts_host.copy_(self._counts, non_blocking=True)
            torch.cuda.current_stream().synchronize()          # the one sync: exact output shapes
            N, E, _, err = (int(v) for v in self._counts_host[:4])
            if err != 0:
                self._bitmap.zero_(); self._cbitmap.zero_(); self._counts.zero_(); self._frontier_cnt.zero_()
                raise _lib.CtanLibraryError(f"neighbor lookup overflowed its workspace (code {err})")
            edge_index = torch.empty((2, E), dtype=torch.long, device=dev)
            e_id = torch.empty(E, dtype=torch.long, device=dev)
            call("ctan_nbr_fill", ptr(self.neighbors), ptr(self.e_id), S, ptr(n_out), ptr(rowptr), ptr(self._assoc),
                 N, None, ptr(edge_index), ptr(edge_index) + 8 * E if E else ptr(edge_index), ptr(e_id),
                 ptr(self._bitmap), ptr(self._cbitmap))
        n_out = n_out[:N]
        rowptr = rowptr[: N + 1]
        _remember_csr(edge_index, N, rowptr)
        return n_out, edge_index, e_id, rowptr
