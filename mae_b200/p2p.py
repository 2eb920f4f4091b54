"""NVLink peer-memory all-gather for the packed posterior rows (host side of csrc/p2p_gather.cu).

Each rank allocates one landing buffer with the library (``cudaMalloc``), publishes its CUDA IPC handle through
``torch.distributed`` and maps every peer's buffer; afterwards ``gather`` is a single kernel launch per rank (no
library collective, CUDA-graph replayable).  Spins inside the kernel are bounded; ``check()`` raises if one timed out.
"""
from __future__ import annotations

import ctypes

import torch
import torch.distributed as dist

from . import _lib


class PeerGather:
    def __init__(self, group, slot_floats: int, device: torch.device):
        lib = _lib.load()
        self.lib = lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.slot_floats = (int(slot_floats) + 3) // 4 * 4
        nbytes = int(lib.mae_comm_buffer_bytes(self.world, self.slot_floats))
        own = ctypes.c_void_p()
        _lib.check(lib.mae_comm_alloc(nbytes, ctypes.byref(own)), "mae_comm_alloc")
        self.own = own.value
        handle = ctypes.create_string_buffer(64)
        _lib.check(lib.mae_comm_get_handle(ctypes.c_void_p(self.own), handle), "mae_comm_get_handle")
        handles = [None] * self.world
        dist.all_gather_object(handles, handle.raw, group=group)
        ptrs, self._opened = [], []
        for r, h in enumerate(handles):
            if r == self.rank:
                ptrs.append(self.own)
                continue
            p = ctypes.c_void_p()
            _lib.check(lib.mae_comm_open_handle(ctypes.create_string_buffer(h, 64), ctypes.byref(p)), "mae_comm_open_handle")
            ptrs.append(p.value)
            self._opened.append(p.value)
        self.peer_table = torch.tensor(ptrs, dtype=torch.int64, device=device)
        torch.cuda.synchronize(device)
        dist.barrier(group=group)

    def usable(self, packed: torch.Tensor) -> bool:
        n = packed.numel()
        return (packed.is_cuda and packed.dtype == torch.float32 and packed.is_contiguous() and n % 4 == 0
                and n <= self.slot_floats and packed.data_ptr() % 16 == 0)

    def gather(self, packed: torch.Tensor) -> torch.Tensor:
        """[n, c] contiguous fp32 -> [world*n, c], rank-major; launched on the current stream."""
        n = packed.numel()
        out = torch.empty(self.world * packed.size(0), packed.size(1), device=packed.device, dtype=torch.float32)
        _lib.check(self.lib.mae_allgather_rows(packed.data_ptr(), n, self.peer_table.data_ptr(), self.rank, self.world,
                                               self.slot_floats, out.data_ptr(),
                                               torch.cuda.current_stream(packed.device).cuda_stream), "mae_allgather_rows")
        from . import ops
        ops._count()
        return out

    def check(self) -> None:
        err = ctypes.c_int(0)
        _lib.check(self.lib.mae_comm_error(ctypes.c_void_p(self.own), ctypes.byref(err)), "mae_comm_error")
        if err.value:
            raise RuntimeError("peer all-gather timed out waiting for a rank (ranks out of step or a peer died)")

    def close(self) -> None:
        for p in self._opened:
            self.lib.mae_comm_close_handle(ctypes.c_void_p(p))
        self._opened = []
        if self.own:
            self.lib.mae_comm_free(ctypes.c_void_p(self.own))
            self.own = None
