"""Symmetric peer memory for the sharded TGB evaluation (csrc/sym.cu): one cudaMalloc'd block per rank, exported with
CUDA IPC, mapped by every rank of the node, so that the per-batch exchange is plain stores over NVLink from inside
the captured graph.  torch.distributed is used ONCE, at construction, to pass the 64-byte IPC handles around."""
from __future__ import annotations

import ctypes

import torch

from . import _lib
from ._lib import CtanLibraryError

_HOST_SIGS = {
    "ctan_sym_alloc": [ctypes.c_size_t, ctypes.POINTER(ctypes.c_void_p)],
    "ctan_sym_free": [ctypes.c_void_p],
    "ctan_ipc_export": [ctypes.c_void_p, ctypes.c_char_p],
    "ctan_ipc_import": [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)],
    "ctan_ipc_close": [ctypes.c_void_p],
}


def _host(name, *args):
    fn = getattr(_lib.lib(), name)
    fn.argtypes, fn.restype = _HOST_SIGS[name], ctypes.c_int
    rc = fn(*args)
    if rc != 0:
        raise CtanLibraryError(f"{name} failed with code {rc}")


class _RawCuda:
    """Minimal __cuda_array_interface__ carrier so torch can view memory it did not allocate."""

    def __init__(self, ptr, shape, typestr):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 3, "strides": None}


class SymmetricExchange:
    """[flags: 2 x world x u64 | pad | rows: 2 x B x row f32] on every rank; `peers` (device int64 [world]) holds this
    process's mapping of every rank's block (own block included)."""

    def __init__(self, B: int, row: int, rank: int, world: int, dev, group=None):
        import torch.distributed as dist
        self.B, self.row, self.rank, self.world, self.dev = B, row, rank, world, dev
        self.rows_off = ((2 * world * 8 + 255) // 256) * 256
        self.bytes = self.rows_off + 2 * B * row * 4
        with torch.cuda.device(dev):
            base = ctypes.c_void_p()
            _host("ctan_sym_alloc", self.bytes, ctypes.byref(base))
            self.base = base.value
            handle = ctypes.create_string_buffer(64)
            _host("ctan_ipc_export", self.base, handle)
            handles = [None] * world
            dist.all_gather_object(handles, bytes(handle.raw), group=group)
            self._mapped = []
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(self.base)
                    continue
                p = ctypes.c_void_p()
                _host("ctan_ipc_import", ctypes.c_char_p(handles[r]), ctypes.byref(p))
                self._mapped.append(p.value)
                ptrs.append(p.value)
            self.peers = torch.tensor(ptrs, dtype=torch.int64, device=dev)
            self.epoch = torch.zeros(1, dtype=torch.int64, device=dev)
            self.ticket = torch.zeros(1, dtype=torch.int32, device=dev)
            self.spins = torch.zeros(1, dtype=torch.int64, device=dev)
            self._rows = torch.as_tensor(_RawCuda(self.base + self.rows_off, (2, B, row), "<f4"), device=dev)
        dist.barrier(group=group)                               # every rank has mapped every block
        self._group = group
        self.count = 0                                          # exchanges issued from the host (parity = count & 1)

    def rows(self, parity: int) -> torch.Tensor:
        return self._rows[parity]

    def rows_ptr(self, parity: int) -> int:
        return self.base + self.rows_off + parity * self.B * self.row * 4

    def flags_ptr(self, parity: int) -> int:
        return self.base + parity * self.world * 8

    def push(self, send: torch.Tensor, n_rows: int, q0: int, parity: int):
        _lib.call("ctan_xchg_push", _lib.ptr(send), n_rows, self.row, q0, self.B, _lib.ptr(self.peers), self.world,
                  self.rank, parity, self.rows_off, _lib.ptr(self.epoch), _lib.ptr(self.ticket))

    def wait(self, parity: int):
        _lib.call("ctan_xchg_wait", self.flags_ptr(parity), self.world, _lib.ptr(self.epoch), _lib.ptr(self.spins))

    def close(self):
        import torch.distributed as dist
        if self.base is None:
            return
        torch.cuda.synchronize(self.dev)
        if dist.is_initialized():
            dist.barrier(group=self._group)                    # nobody still stores into a block that is going away
        with torch.cuda.device(self.dev):
            for p in self._mapped:
                _host("ctan_ipc_close", p)
            self._rows = None
            _host("ctan_sym_free", self.base)
        self.base, self._mapped = None, []
