"""NVLink peer-memory all-gather for the packed posterior rows (host side of csrc/p2p_gather.cu).

Each rank allocates one landing buffer with the library (``cudaMalloc``), publishes its CUDA IPC handle through
``torch.distributed`` and maps every peer's buffer; afterwards ``gather`` is a single kernel launch per rank (no
library collective, CUDA-graph replayable).  Spins inside the kernel are bounded: on a timeout the kernel NaN-fills the
slot it was waiting for and raises a sticky error flag that ``check()`` turns into an exception.

Construction is COLLECTIVE and its outcome is agreed on by all ranks (``PeerGather.create``): a rank whose allocation or
IPC mapping fails still takes part in every exchange, and if any rank failed, every rank gets ``None`` and uses the
library all-gather, so ranks can never end up in different collectives.
"""
from __future__ import annotations

import ctypes
from typing import Optional

import torch
import torch.distributed as dist

from . import _lib


class PeerGather:
    def __init__(self):
        raise RuntimeError("use PeerGather.create(group, slot_floats, device)")

    @classmethod
    def create(cls, group, slot_floats: int, device: torch.device, layout: str = "gather") -> Optional["PeerGather"]:
        """Collective over ``group``.  Returns a ready object on EVERY rank, or None on every rank (with the reason in
        ``PeerGather.last_failure``) when any rank could not allocate / export / map the buffers."""
        self = object.__new__(cls)
        lib = _lib.load()
        self.lib = lib
        self.group = group
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self.device = device
        self.slot_floats = (int(slot_floats) + 3) // 4 * 4
        self.own = None
        self._opened = []
        self.calls = 0
        why = None
        handle_raw = None
        with torch.cuda.device(device):
            try:
                # "gather": landing buffers of mae_allgather_rows; "rows": those of mae_kl_mpd_rows_fused (the exchange
                # fused into the regulariser's cluster kernel: its own header, flags and partial-sum slots)
                self.layout = layout
                nbytes = int((lib.mae_rows_comm_bytes if layout == "rows" else lib.mae_comm_buffer_bytes)(self.world, self.slot_floats))
                if nbytes <= 0:
                    raise RuntimeError("unsupported world size / slot size for the %s buffers" % layout)
                own = ctypes.c_void_p()
                _lib.check(lib.mae_comm_alloc(nbytes, ctypes.byref(own)), "mae_comm_alloc")
                self.own = own.value
                handle = ctypes.create_string_buffer(64)
                _lib.check(lib.mae_comm_get_handle(ctypes.c_void_p(self.own), handle), "mae_comm_get_handle")
                handle_raw = handle.raw
            except Exception as exc:      # keep going: the peers are waiting in the exchange below
                why = "rank %d: %s" % (self.rank, exc)
            handles = [None] * self.world
            dist.all_gather_object(handles, handle_raw, group=group)
            ptrs = []
            if why is None and all(h is not None for h in handles):
                try:
                    for r, h in enumerate(handles):
                        if r == self.rank:
                            ptrs.append(self.own)
                            continue
                        p = ctypes.c_void_p()
                        _lib.check(lib.mae_comm_open_handle(ctypes.create_string_buffer(h, 64), ctypes.byref(p)),
                                   "mae_comm_open_handle")
                        ptrs.append(p.value)
                        self._opened.append(p.value)
                except Exception as exc:
                    why = "rank %d: %s" % (self.rank, exc)
            elif why is None:
                why = "a peer could not export its buffer"
            ok = torch.tensor([0 if why is not None else 1], device=device, dtype=torch.int32)
            dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)      # every rank takes the same decision
            if int(ok) == 0:
                cls.last_failure = why or "a peer failed to map the buffers"
                self.close()
                return None
            self.peer_table = torch.tensor(ptrs, dtype=torch.int64, device=device)
            torch.cuda.synchronize(device)
            dist.barrier(group=group)
        return self

    last_failure: Optional[str] = None

    def usable(self, packed: torch.Tensor) -> bool:
        n = packed.numel()
        return (packed.is_cuda and packed.dtype == torch.float32 and packed.is_contiguous() and n % 4 == 0
                and n <= self.slot_floats and packed.data_ptr() % 16 == 0)

    def gather(self, packed: torch.Tensor) -> torch.Tensor:
        """[n, c] contiguous fp32 -> [world*n, c], rank-major; launched on the current stream."""
        n = packed.numel()
        out = torch.empty(self.world * packed.size(0), packed.size(1), device=packed.device, dtype=torch.float32)
        with torch.cuda.device(packed.device):
            _lib.check(self.lib.mae_allgather_rows(packed.data_ptr(), n, self.peer_table.data_ptr(), self.rank, self.world,
                                                   self.slot_floats, out.data_ptr(),
                                                   torch.cuda.current_stream(packed.device).cuda_stream), "mae_allgather_rows")
        from . import ops
        ops._count()
        self.calls += 1
        return out

    def check(self) -> None:
        """Raises if a gather timed out waiting for a peer (synchronises with the device: do not call while capturing)."""
        err = ctypes.c_int(0)
        with torch.cuda.device(self.device):
            _lib.check(self.lib.mae_comm_error(ctypes.c_void_p(self.own), ctypes.byref(err)), "mae_comm_error")
        if err.value:
            raise RuntimeError("peer all-gather timed out waiting for a rank (ranks out of step or a peer died); the rows of "
                               "that rank were NaN-filled, so losses and gradients since then are NaN")

    def close(self) -> None:
        for p in self._opened:
            self.lib.mae_comm_close_handle(ctypes.c_void_p(p))
        self._opened = []
        if self.own:
            self.lib.mae_comm_free(ctypes.c_void_p(self.own))
            self.own = None
