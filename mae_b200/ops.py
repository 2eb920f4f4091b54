"""Autograd-aware PyTorch operators over the C ABI of libmae_b200.so.

PyTorch is plumbing here (device memory, streams, autograd bookkeeping); all arithmetic of the hot path
runs in the hand-written sm_100a kernels.  Every op raises on non-CUDA tensors: there is no CPU or eager
fallback.  Launches go to ``torch.cuda.current_stream()`` so the ops compose with CUDA-graph capture.
"""
from __future__ import annotations

import os
import threading
from typing import Dict, List, Optional, Tuple

import torch

from . import _lib

__all__ = [
    "reparam_kl", "mpd", "mpd_forward_only", "bernoulli_recon_error", "dmol_recon_error", "log_prob_prior",
    "log_prob_posterior", "iw_nll", "assemble_loss", "kl_mpd", "loss_tail", "loss_head", "encode_head", "kl_mpd_reg_fused", "kl_mpd_reg", "mpd_rows_loss_backward", "reg_rows_fused", "reg_rows_cluster", "rows_cluster_supported", "reparam_fwd",
    "reparam_reg_bwd", "fused_reg_supported", "set_pdl", "set_unit_loss_grad", "set_dmol_head", "dmol_head_mode", "dmol_head_supported", "dmol_head_recon_error", "join_side", "launch_count", "reset_launch_count",
]

_launches = 0  # number of kernels of ours enqueued since the last reset (bench.py reports it)


def launch_count() -> int:
    return _launches


def reset_launch_count() -> None:
    global _launches
    _launches = 0


def _count(n: int = 1) -> None:
    global _launches
    _launches += n


def _lib_():
    return _lib.load()


def set_pdl(enabled: bool) -> bool:
    """Process-wide switch of programmatic dependent launch between the kernels of a chain (mae_set_pdl); returns the
    previous setting.  On by default; the batch-sharded head turns it off."""
    return bool(_lib_().mae_set_pdl(1 if enabled else 0))


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _on:
    """Device guard: the library launches on the CURRENT device and ``_stream()`` is that device's current stream, so every
    op body runs with the device of its tensors current (a model on cuda:1 while cuda:0 is current, MAE.load(path,
    'cuda:1'), DataParallel replicas).  A no-op when the device already is current."""
    __slots__ = ("idx", "prev")

    def __init__(self, dev):
        idx = dev.index if isinstance(dev, torch.device) else dev
        self.idx = idx if (idx is not None and idx != torch.cuda.current_device()) else None

    def __enter__(self):
        if self.idx is not None:
            self.prev = torch.cuda.current_device()
            torch.cuda.set_device(self.idx)
        return self

    def __exit__(self, *exc):
        if self.idx is not None:
            torch.cuda.set_device(self.prev)
        return False


def _guarded(fn):
    """Decorator for autograd.Function.forward / backward and plain ops: runs the body under the device of the first CUDA
    tensor among the arguments (backward: of the tensors the forward saw, remembered on ctx)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kwargs):
        dev = None
        for a in args:
            if isinstance(a, torch.Tensor) and a.is_cuda:
                dev = a.device
                break
        ctx = args[0] if args and hasattr(args[0], "save_for_backward") else None
        if ctx is not None:
            if dev is None:
                dev = getattr(ctx, "_mae_dev", None)
            else:
                ctx._mae_dev = dev
        if dev is None:
            return fn(*args, **kwargs)
        with _on(dev):
            return fn(*args, **kwargs)

    return wrapper


def _req_cuda(*tensors: Optional[torch.Tensor]) -> None:
    dev = None
    for t in tensors:
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("mae_b200 ops run on CUDA tensors only (got %s): there is no CPU fallback" % t.device)
        if t.dtype != torch.float32:
            raise RuntimeError("mae_b200 ops compute in float32 (got %s)" % t.dtype)
        if dev is None:
            dev = t.device
        elif t.device != dev:
            raise RuntimeError("mae_b200 ops need all tensors on one device (got %s and %s)" % (dev, t.device))


def _ptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def _rows(t: torch.Tensor) -> Tuple[torch.Tensor, int]:
    """View ``t`` [B,*] as rows of a strided matrix without copying when possible (the reference's mu/logvar
    are ``chunk(2,1)`` views, encoder_cores/resnet.py:43).  Returns (2-D tensor, row stride in elements)."""
    b = t.size(0)
    d = t.numel() // max(b, 1)
    try:
        v = t.view(b, d)
    except RuntimeError:
        v = t.reshape(b, d)
    if d > 1 and v.stride(1) != 1:
        v = v.contiguous()
    if v.data_ptr() % 4 != 0:
        v = v.contiguous()
    stride = v.stride(0) if b > 1 else d
    if stride < d:
        v = v.contiguous()
        stride = d
    return v, stride


# ------------------------------------------------------------------------------------------------
# workspace pool: zero-filled once (tickets are self-resetting), reused across steps per stream
# ------------------------------------------------------------------------------------------------
class _Pool:
    def __init__(self) -> None:
        self._free: Dict[tuple, List[torch.Tensor]] = {}
        self._lock = threading.Lock()

    def take(self, nbytes: int, device: torch.device) -> torch.Tensor:
        key = (device.index, _stream(), int(nbytes))
        with self._lock:
            lst = self._free.get(key)
            if lst:
                return lst.pop()
        return torch.zeros(int(nbytes), dtype=torch.uint8, device=device)

    def give(self, ws: torch.Tensor) -> None:
        key = (ws.device.index, _stream(), ws.numel())
        with self._lock:
            self._free.setdefault(key, []).append(ws)


_pool = _Pool()


# ------------------------------------------------------------------------------------------------
# side stream: the MPD kernels only depend on (mu, logvar), so inside MAE.loss they run concurrently with
# the decoder conv stack and the likelihood kernels (forward) and with the decoder's backward (backward).
# ------------------------------------------------------------------------------------------------
_side_streams: Dict[int, "torch.cuda.Stream"] = {}
_pending_join: Dict[int, "torch.cuda.Event"] = {}     # forward: MPD results not yet joined into the main stream
_grads_ready: Dict[int, "torch.cuda.Event"] = {}      # backward: recorded right after the assemble backward kernel


def _side_stream(device: torch.device) -> "torch.cuda.Stream":
    idx = device.index if device.index is not None else torch.cuda.current_device()
    st = _side_streams.get(idx)
    if st is None:
        # high priority: the side-stream kernels are tiny (a handful of CTAs) but sit on a dependent chain; they must
        # get SM slots as soon as CTAs of the big likelihood kernel retire instead of queueing behind its whole grid
        st = torch.cuda.Stream(device=device, priority=-1)
        _side_streams[idx] = st
    return st


def join_side(device: torch.device) -> None:
    """Make the current stream wait for MPD work that was launched on the side stream (no-op if none)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    pending = _pending_join.pop(idx, None)
    if pending is not None:
        torch.cuda.current_stream(device).wait_event(pending[0])


# ------------------------------------------------------------------------------------------------
# A1 + A2
# ------------------------------------------------------------------------------------------------
class _ReparamKL(torch.autograd.Function):
    @staticmethod
    @_guarded
    def forward(ctx, mu, logvar, eps, want_kl):
        ctx.set_materialize_grads(False)
        _req_cuda(mu, logvar, eps)
        b, ns = eps.size(0), eps.size(1)
        mu2, ms = _rows(mu)
        lv2, ls = _rows(logvar)
        d = mu2.size(1)
        eps_c = eps.contiguous()
        z = torch.empty_like(eps_c)
        kl = torch.empty(b, device=mu.device, dtype=torch.float32) if want_kl else None
        _lib.check(_lib_().mae_reparam_kl_fwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, eps_c.data_ptr(), b, ns, d,
                                              z.data_ptr(), _ptr(kl), _stream()), "mae_reparam_kl_fwd")
        _count()
        ctx.save_for_backward(mu2, lv2, eps_c)
        ctx.meta = (ms, ls, b, ns, d, mu.shape, logvar.shape)
        if kl is None:
            kl = torch.zeros(0, device=mu.device)
            ctx.mark_non_differentiable(kl)
        return z, kl

    @staticmethod
    @_guarded
    def backward(ctx, gz, gkl):
        mu2, lv2, eps_c = ctx.saved_tensors
        ms, ls, b, ns, d, mu_shape, lv_shape = ctx.meta
        if gz is None and gkl is None:
            return None, None, None, None
        gz = gz.contiguous() if gz is not None else None
        gkl = gkl.contiguous() if (gkl is not None and gkl.numel() == b) else None
        dmu = torch.empty(b, d, device=mu2.device, dtype=torch.float32)
        dlv = torch.empty_like(dmu)
        _lib.check(_lib_().mae_reparam_kl_bwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, eps_c.data_ptr(), _ptr(gz), _ptr(gkl),
                                              b, ns, d, dmu.data_ptr(), dlv.data_ptr(), 0, _stream()), "mae_reparam_kl_bwd")
        _count()
        return dmu.view(mu_shape), dlv.view(lv_shape), None, None


def reparam_kl(mu: torch.Tensor, logvar: torch.Tensor, eps: torch.Tensor, want_kl: bool = True):
    """z = eps*exp(0.5*logvar)+mu  [B,ns,*z]  and the analytic KL [B] (or None).
    Reference: gaussian_encoder.py:56-62 and :219."""
    z, kl = _ReparamKL.apply(mu, logvar, eps, want_kl)
    return z, (kl if want_kl else None)


# ------------------------------------------------------------------------------------------------
# A3
# ------------------------------------------------------------------------------------------------
def _mpd_alloc_outputs(d: int, dev: torch.device):
    return (torch.empty(d, device=dev, dtype=torch.float32), torch.empty((), device=dev, dtype=torch.float32),
            torch.empty(1, device=dev, dtype=torch.float64))


def _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, store_kl, outs, row0=0, nrows=None, kl=None):
    lib = _lib_()
    m_d, s, sumsq = outs
    nrows = b if nrows is None else nrows
    _lib.check(lib.mae_mpd_prepare(mu2.data_ptr(), ms, lv2.data_ptr(), ls, b, d, ws.data_ptr(), ws.numel(), m_d.data_ptr(),
                                   _ptr(kl), _stream()), "mae_mpd_prepare")
    _lib.check(lib.mae_mpd_fwd(ws.data_ptr(), ws.numel(), b, d, row0, nrows, int(store_kl), sumsq.data_ptr(), s.data_ptr(),
                               _stream()), "mae_mpd_fwd")
    _count(2)
    return m_d, s, sumsq


# Largest pair of stored B x B KL matrices (KL and its transpose, 8*Bp^2 bytes) the backward may keep between forward
# and backward; above it the tiles are recomputed (path 2).  B200: 180 GB of HBM; B = 32768 needs 8.6 GB, B = 65536 34.4 GB.
MPD_KL_BUDGET_BYTES = int(os.environ.get("MAE_MPD_KL_BUDGET_BYTES", 48 << 30))


def _mpd_pick_path(b: int, d: int, requested: int) -> int:
    """Backward path of mae_mpd_bwd: 1 stored-KL per-row kernel (B <= 2048), 3 stored-KL tile GEMMs, 2 recompute."""
    if requested in (2, 3):
        return requested
    if requested == 1 or b <= 2048:
        if b > 2048:
            raise ValueError("bwd_path=1 (per-row stored-KL kernel) supports B <= 2048")
        return 1
    lib = _lib_()
    extra = lib.mae_mpd_workspace_bytes(b, d, 1) - lib.mae_mpd_workspace_bytes(b, d, 0)
    return 3 if extra <= MPD_KL_BUDGET_BYTES else 2


class _MPD(torch.autograd.Function):
    """(mu, logvar) -> (m_d, S, KL).  KL is the analytic KL(q || N(0,I)) when ``with_kl`` (else an empty tensor): it
    shares the O(B*D) pre-pass with the MPD, and its gradient is folded into the MPD backward kernel."""

    @staticmethod
    @_guarded
    def forward(ctx, mu, logvar, bwd_path, side, with_kl):
        ctx.set_materialize_grads(False)
        _req_cuda(mu, logvar)
        b = mu.size(0)
        if b < 2:
            raise RuntimeError("the MPD regulariser needs a batch of at least 2 posteriors (gaussian_encoder.py:181)")
        mu2, ms = _rows(mu)
        lv2, ls = _rows(logvar)
        d = mu2.size(1)
        lib = _lib_()
        dev = mu.device
        need_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        path = _mpd_pick_path(b, d, bwd_path) if need_grad else 2
        store = path != 2
        nbytes = lib.mae_mpd_workspace_bytes(b, d, int(store))
        outs = _mpd_alloc_outputs(d, dev)   # owned by the main stream; the side stream only writes them
        kl = torch.empty(b if with_kl else 0, device=dev, dtype=torch.float32)
        klp = kl if with_kl else None
        if side:
            cs = torch.cuda.current_stream(dev)
            ss = _side_stream(dev)
            ss.wait_stream(cs)
            with torch.cuda.stream(ss):
                ws = _pool.take(nbytes, dev)
                m_d, s, _ = _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, store, outs, kl=klp)
                ev = torch.cuda.Event()
                ev.record(ss)
            # everything the side-stream kernels touch stays referenced until the join
            _pending_join[dev.index if dev.index is not None else torch.cuda.current_device()] = (ev, (outs, kl, mu2, lv2))
        else:
            ws = _pool.take(nbytes, dev)
            m_d, s, _ = _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, store, outs, kl=klp)
        if need_grad:
            ctx.ws = ws
            ctx.side = bool(side)
            ctx.save_for_backward(mu2, lv2, s)
            ctx.meta = (ms, ls, b, d, mu.shape, logvar.shape, path)
        else:
            _MPD._release(ws, dev, side)
        if not with_kl:
            ctx.mark_non_differentiable(kl)
        return m_d.view(mu.shape[1:]), s, kl

    @staticmethod
    def _release(ws, dev, side):
        if side:
            with torch.cuda.stream(_side_stream(dev)):
                _pool.give(ws)
        else:
            _pool.give(ws)

    @staticmethod
    @_guarded
    def backward(ctx, g_m, g_s, g_kl):
        mu2, lv2, s = ctx.saved_tensors
        ms, ls, b, d, mu_shape, lv_shape, path = ctx.meta
        ws, side = ctx.ws, ctx.side
        dev = mu2.device
        ctx.ws = None
        if g_kl is not None and g_kl.numel() != b:
            g_kl = None
        if g_m is None and g_s is None and g_kl is None:
            _MPD._release(ws, dev, side)
            _grads_ready.pop(dev.index if dev.index is not None else torch.cuda.current_device(), None)
            return None, None, None, None, None
        g_m = g_m.contiguous().view(-1) if g_m is not None else None
        g_s = g_s.contiguous() if g_s is not None else None
        g_kl = g_kl.contiguous() if g_kl is not None else None
        dmu = dlv = None

        def launch():
            # allocated on the stream that writes them (on the side-stream path autograd has already queued the decoder's
            # backward on the main stream: a block freshly freed there may still be in use by a pending main-stream kernel)
            nonlocal dmu, dlv
            dmu = torch.empty(b, d, device=dev, dtype=torch.float32)
            dlv = torch.empty_like(dmu)
            _lib.check(_lib_().mae_mpd_bwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, ws.data_ptr(), ws.numel(), b, d, 0, b,
                                           _ptr(g_m), _ptr(g_s), s.data_ptr(), _ptr(g_kl), dmu.data_ptr(), dlv.data_ptr(), 0, path,
                                           _stream()), "mae_mpd_bwd")
            _count()

        if side:
            # run on the side stream as soon as the upstream scalars exist (event recorded by the assemble backward),
            # concurrently with whatever the main stream already has queued (likelihood / decoder backward)
            cs = torch.cuda.current_stream(dev)
            ss = _side_stream(dev)
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            ready = _grads_ready.pop(idx, None)
            ptrs = {t.data_ptr() for t in (g_m, g_s, g_kl) if t is not None}
            if ready is not None and ptrs <= ready[1]:
                ss.wait_event(ready[0])     # produced by the assemble backward of this pass
            else:
                ss.wait_stream(cs)
            with torch.cuda.stream(ss):
                launch()
                _pool.give(ws)
                ev = torch.cuda.Event()
                ev.record(ss)
            cs.wait_event(ev)   # later main-stream work (and any reuse of g_m / g_s memory) is ordered after the kernel
            dmu.record_stream(cs)   # consumed and freed on the main stream
            dlv.record_stream(cs)
        else:
            launch()
            _pool.give(ws)
        return dmu.view(mu_shape), dlv.view(lv_shape), None, None, None


def kl_mpd(mu: torch.Tensor, logvar: torch.Tensor, side_stream: bool = False, bwd_path: int = 0):
    """Analytic KL [B] and MPD (m_d, S) from one pre-pass: returns (kl, m_d, S).  One autograd node whose backward is a
    single kernel.  Reference: gaussian_encoder.py:219 and :162-188.  See :func:`mpd` for ``side_stream``."""
    m_d, s, kl = _MPD.apply(mu, logvar, bwd_path, side_stream, True)
    return kl, m_d, s


def mpd(mu: torch.Tensor, logvar: torch.Tensor, bwd_path: int = 0, side_stream: bool = False):
    """Mutual posterior divergence: (m_d with the shape of one latent, S scalar).
    Reference: GaussianEncoder._postKL, gaussian_encoder.py:162-188.  ``bwd_path``: 0 auto, 1 stored-KL per-row kernel (B <= 2048), 2 recompute tiles, 3 stored-KL tile GEMMs.

    ``side_stream=True`` launches the kernels on a side stream; the results may only be consumed after
    :func:`join_side` (``assemble_loss`` does it).  Used by ``MAE.loss`` / ``heads`` to overlap the regulariser with
    the decoder; the public ``encode()`` keeps the synchronous default."""
    m_d, s, _ = _MPD.apply(mu, logvar, bwd_path, side_stream, False)
    return m_d, s


@_guarded
def mpd_forward_only(mu: torch.Tensor, logvar: torch.Tensor, row0: int = 0, nrows: Optional[int] = None):
    """Forward on a row block: returns (m_d, S_of_block_as_if_complete, sumsq[1] fp64).  For multi-GPU
    row-blocked evaluation: all-reduce ``sumsq`` and call :func:`mpd_finalize`."""
    _req_cuda(mu, logvar)
    mu2, ms = _rows(mu)
    lv2, ls = _rows(logvar)
    b, d = mu2.shape
    ws = _pool.take(_lib_().mae_mpd_workspace_bytes(b, d, 0), mu.device)
    m_d, s, sumsq = _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, False, _mpd_alloc_outputs(d, mu.device), row0, nrows)
    _pool.give(ws)
    return m_d.view(mu.shape[1:]), s, sumsq


class MpdRowsState:
    """What the row-blocked MPD backward needs from the forward (workspace with panels / stored KL, S, geometry)."""

    def __init__(self, ws, mu2, ms, lv2, ls, b, d, s, path):
        self.ws, self.mu2, self.ms, self.lv2, self.ls, self.b, self.d, self.s, self.path = ws, mu2, ms, lv2, ls, b, d, s, path


@_guarded
def mpd_rows_forward(mu_all: torch.Tensor, logvar_all: torch.Tensor, row0: int, nrows: int, full: bool):
    """Row-blocked MPD forward for batch-sharded runs (mu_all / logvar_all are the ALL-GATHERED posteriors [B,D]).

    ``full=True`` (small global batch): every rank evaluates the whole B x B forward redundantly and keeps the KL
    tiles -> (m_d, S, kl_all, state), no collective needed.  ``full=False``: only the tiles of rows
    [row0,row0+nrows) -> (m_d, sumsq_partial[1] fp64, kl_all, state); the caller all-reduces ``sumsq`` and calls
    :func:`mpd_finalize`, then sets ``state.s``.  kl_all is the analytic KL of every gathered row [B]."""
    _req_cuda(mu_all, logvar_all)
    mu2, ms = _rows(mu_all)
    lv2, ls = _rows(logvar_all)
    b, d = mu2.shape
    if b < 2:
        raise RuntimeError("the MPD regulariser needs a global batch of at least 2 posteriors")
    dev = mu2.device
    lib = _lib_()
    small = full and b <= 2048
    if full:
        store, path = small, (1 if small else 2)
        nbytes = lib.mae_mpd_workspace_bytes(b, d, int(small))
    else:
        # row block: keep KL_ij and KL_ji of the own rows (two tile passes) when they fit the budget, so that the backward
        # is the stored-KL tile kernel (8 B_loc B D executed flop) instead of the recompute path (16, times the K chunks)
        base = lib.mae_mpd_workspace_bytes(b, d, 0)
        with_rows = lib.mae_mpd_workspace_bytes_rows(b, d, row0, nrows)
        store = with_rows - base <= MPD_KL_BUDGET_BYTES
        path = 3 if store else 2
        nbytes = with_rows if store else base
    ws = _pool.take(nbytes, dev)
    kl_all = torch.empty(b, device=dev, dtype=torch.float32)
    outs = _mpd_alloc_outputs(d, dev)
    if full:
        m_d, s, sumsq = _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, store, outs, kl=kl_all)
    else:
        m_d, s, sumsq = _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, store, outs, row0, nrows, kl=kl_all)
        if store:
            _count()      # the transposed tile pass
    state = MpdRowsState(ws, mu2, ms, lv2, ls, b, d, s if full else None, path)
    state.outs = outs       # the tile kernel writes its fp64 sum into outs[2]: alive as long as the state (side-stream callers)
    return m_d, (s if full else sumsq), kl_all, state


@_guarded
def mpd_rows_backward(state: "MpdRowsState", row0: int, nrows: int, g_m, g_s, g_kl_rows):
    """Gradient w.r.t. mu / logvar of rows [row0,row0+nrows) only (complete: as the i and as the j argument), so a
    batch-sharded run needs no reduce-scatter.  g_kl_rows [nrows] is the upstream gradient of those rows' KL."""
    dev = state.mu2.device
    dmu = torch.empty(nrows, state.d, device=dev, dtype=torch.float32)
    dlv = torch.empty_like(dmu)
    g_m = g_m.contiguous().view(-1) if g_m is not None else None
    g_s = g_s.contiguous() if g_s is not None else None
    gkl_ptr = None
    if g_kl_rows is not None:
        g_kl_rows = g_kl_rows.contiguous()
        gkl_ptr = g_kl_rows.data_ptr() - 4 * row0      # the kernel indexes gkl by GLOBAL row and only touches own rows
    _lib.check(_lib_().mae_mpd_bwd(state.mu2.data_ptr(), state.ms, state.lv2.data_ptr(), state.ls, state.ws.data_ptr(),
                                   state.ws.numel(), state.b, state.d, row0, nrows, _ptr(g_m), _ptr(g_s), state.s.data_ptr(),
                                   gkl_ptr, dmu.data_ptr(), dlv.data_ptr(), 0, state.path, _stream()), "mae_mpd_bwd")
    _count()
    _pool.give(state.ws)
    state.ws = None
    return dmu, dlv


@_guarded
def mpd_rows_loss_backward(state: "MpdRowsState", row0: int, nrows: int, eta: float, gamma: float, free_bits: float,
                           kl_weight: float, m_d: torch.Tensor, kl_all: Optional[torch.Tensor], want_grad: bool = True):
    """Regulariser value and its gradient w.r.t. rows [row0,row0+nrows), enqueued right behind the forward kernels
    (mae_mpd_loss_bwd): -> (grads [2,nrows,D] or None, reg []).  ``state.s`` must hold S (finalized in row-sharded runs).
    Consumes the state (the workspace goes back to the pool)."""
    dev = state.mu2.device
    grads = torch.empty(2, nrows, state.d, device=dev, dtype=torch.float32) if want_grad else None
    reg = torch.empty((), device=dev, dtype=torch.float32)
    m_c = m_d.contiguous().view(-1)
    _lib.check(_lib_().mae_mpd_loss_bwd(state.mu2.data_ptr(), state.ms, state.lv2.data_ptr(), state.ls, state.ws.data_ptr(),
                                        state.ws.numel(), state.b, state.d, row0, nrows, float(eta), float(gamma),
                                        float(free_bits), float(kl_weight), m_c.data_ptr(), state.s.data_ptr(), _ptr(kl_all),
                                        reg.data_ptr(), grads[0].data_ptr() if want_grad else None,
                                        grads[1].data_ptr() if want_grad else None, state.path, _stream()), "mae_mpd_loss_bwd")
    _count(2 if want_grad else 1)
    _pool.give(state.ws)
    state.ws = None
    return grads, reg


def reg_rows_fused(mu_all: torch.Tensor, logvar_all: torch.Tensor, row0: int, nrows: int, eta: float, gamma: float,
                   free_bits: float, want_grad: bool = True, kl_weight: Optional[float] = None):
    """Batch-sharded runs whose GLOBAL batch fits the cluster kernel (B_global <= 128, D <= 64): the whole regulariser of
    the gathered rows in one launch instead of prepare -> tiles -> gradients -> backward.  Every rank evaluates all rows
    (a few microseconds) and keeps the gradient of its own.  -> (m_d, S, kl_all, grads [2,nrows,D] or None, reg), or None
    when the shape is outside the kernel."""
    mu2, _ = _rows(mu_all)
    b, d = mu2.shape
    if not fused_reg_supported(b, d, mu2.device):
        return None
    kl_all, m_d, s_, reg, grads, _ = kl_mpd_reg_fused(mu_all, logvar_all, eta, gamma, free_bits, kl_weight, want_grad)
    own = grads[:, row0:row0 + nrows] if grads is not None else None      # [2, nrows, D]: rows are contiguous
    return m_d, s_, kl_all, own, reg


_rows_ok: Dict[Tuple[int, int, int, int], bool] = {}


def rows_cluster_supported(world: int, b_loc: int, d: int, device: torch.device) -> bool:
    """True when (world, local batch, D) fits the one-launch sharded regulariser (mae_kl_mpd_rows_fused)."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, int(world), int(b_loc), int(d))
    ok = _rows_ok.get(key)
    if ok is None:
        with torch.cuda.device(idx):
            ok = bool(_lib_().mae_kl_mpd_rows_fused_supported(int(world), int(b_loc), int(d)))
        _rows_ok[key] = ok
    return ok


@_guarded
def reg_rows_cluster(packed: torch.Tensor, peer, eta: float, gamma: float, free_bits: float,
                     kl_weight: Optional[float] = None, want_grad: bool = True, scratch: Optional[torch.Tensor] = None):
    """Batch-sharded runs: exchange of the packed rows over NVLink peer memory AND the regulariser of the global batch,
    forward and backward for the own rows, in ONE cluster launch (csrc/mpd_rows_fused.cu).  ``packed`` [b_loc, 2D] =
    (mu | logvar) of this rank, contiguous; ``peer`` a ``p2p.PeerGather`` created with ``layout="rows"``.
    -> (m_d [D], S [], kl_all [world*b_loc], grads [2,b_loc,D] or None, reg [])."""
    _req_cuda(packed)
    b_loc, w2 = packed.shape
    d = w2 // 2
    dev = packed.device
    world, rank = peer.world, peer.rank
    kw = 1.0 / float(world * b_loc) if kl_weight is None else float(kl_weight)
    lib = _lib_()
    if scratch is None:
        scratch = torch.empty(int(lib.mae_rows_scratch_bytes()), dtype=torch.uint8, device=dev)
    m_d = torch.empty(d, device=dev, dtype=torch.float32)
    s_ = torch.empty((), device=dev, dtype=torch.float32)
    reg = torch.empty((), device=dev, dtype=torch.float32)
    kl_all = torch.empty(world * b_loc, device=dev, dtype=torch.float32)
    grads = torch.empty(2, b_loc, d, device=dev, dtype=torch.float32) if want_grad else None
    _lib.check(lib.mae_kl_mpd_rows_fused(packed.data_ptr(), peer.peer_table.data_ptr(), rank, world, peer.slot_floats, b_loc, d,
                                         float(eta), float(gamma), float(free_bits), kw, m_d.data_ptr(), s_.data_ptr(),
                                         kl_all.data_ptr(), reg.data_ptr(), grads[0].data_ptr() if want_grad else None,
                                         grads[1].data_ptr() if want_grad else None, scratch.data_ptr(), scratch.numel(),
                                         _stream()), "mae_kl_mpd_rows_fused")
    _count()
    peer.calls += 1
    return m_d, s_, kl_all, grads, reg


@_guarded
def reparam_fwd(mu: torch.Tensor, logvar: torch.Tensor, eps: torch.Tensor):
    """z = eps * exp(0.5 logvar) + mu without an autograd node (building block of the fused encoder nodes):
    -> (z, saved) with ``saved`` = what :func:`reparam_reg_bwd` needs."""
    _req_cuda(mu, logvar, eps)
    mu2, ms = _rows(mu)
    lv2, ls = _rows(logvar)
    b, ns, d = eps.size(0), eps.size(1), mu2.size(1)
    eps_c = eps.contiguous()
    z = torch.empty_like(eps_c)
    _lib.check(_lib_().mae_reparam_kl_fwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, eps_c.data_ptr(), b, ns, d, z.data_ptr(), None,
                                          _stream()), "mae_reparam_kl_fwd")
    _count()
    return z, (mu2, ms, lv2, ls, eps_c, b, ns, d)


@_guarded
def reparam_reg_bwd(saved, gz: Optional[torch.Tensor], grads: Optional[torch.Tensor], g_reg: Optional[torch.Tensor]):
    """d(mu, logvar) = reparameterisation backward of d z  +  g_reg * grads  in one launch (mae_reparam_reg_bwd);
    grads = [2,B,D] from the regulariser kernels, g_reg a device scalar.  -> (dmu, dlv) dense [B,D]."""
    mu2, ms, lv2, ls, eps_c, b, ns, d = saved
    dev = mu2.device
    gz = gz.contiguous() if gz is not None else None
    use = grads is not None and g_reg is not None
    g_reg = g_reg.contiguous() if use else None
    dmu = torch.empty(b, d, device=dev, dtype=torch.float32)
    dlv = torch.empty_like(dmu)
    _lib.check(_lib_().mae_reparam_reg_bwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, eps_c.data_ptr(), _ptr(gz), None,
                                           grads[0].data_ptr() if use else None, grads[1].data_ptr() if use else None,
                                           _ptr(g_reg), b, ns, d, dmu.data_ptr(), dlv.data_ptr(), 0, _stream()),
               "mae_reparam_reg_bwd")
    _count()
    return dmu, dlv


@_guarded
def mpd_finalize(sumsq: torch.Tensor, batch: int) -> torch.Tensor:
    s = torch.empty((), device=sumsq.device, dtype=torch.float32)
    _lib.check(_lib_().mae_mpd_finalize(sumsq.data_ptr(), batch, s.data_ptr(), _stream()), "mae_mpd_finalize")
    _count()
    return s


# ------------------------------------------------------------------------------------------------
# A4a
# ------------------------------------------------------------------------------------------------
class _Bernoulli(torch.autograd.Function):
    @staticmethod
    @_guarded
    def forward(ctx, probs, x):
        _req_cuda(probs, x)
        b, ns, p = probs.shape
        probs_c = probs.contiguous()
        x_c = x.reshape(b, -1).contiguous()
        if x_c.size(1) != p:
            raise RuntimeError("x has %d pixels per image, probabilities have %d" % (x_c.size(1), p))
        err = torch.empty(b, ns, device=probs.device, dtype=torch.float32)
        _lib.check(_lib_().mae_bernoulli_fwd(probs_c.data_ptr(), x_c.data_ptr(), b, ns, p, err.data_ptr(), _stream()),
                   "mae_bernoulli_fwd")
        _count()
        ctx.save_for_backward(probs_c, x_c)
        return err

    @staticmethod
    @_guarded
    def backward(ctx, g):
        probs_c, x_c = ctx.saved_tensors
        b, ns, p = probs_c.shape
        g = g.contiguous()
        d = torch.empty_like(probs_c)
        _lib.check(_lib_().mae_bernoulli_bwd(probs_c.data_ptr(), x_c.data_ptr(), g.data_ptr(), 1, 1.0, b, ns, p, d.data_ptr(), _stream()),
                   "mae_bernoulli_bwd")
        _count()
        return d, None


def bernoulli_recon_error(probs: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """probs [B,ns,P] post-sigmoid, x [B,*] -> reconstruction error [B,ns].
    Reference: binary_image_decoder.py:84-93 / pixelcnn.py:136-139."""
    return _Bernoulli.apply(probs, x)


# ------------------------------------------------------------------------------------------------
# A4b + A4c
# ------------------------------------------------------------------------------------------------
class _Dmol(torch.autograd.Function):
    @staticmethod
    @_guarded
    def forward(ctx, raw, x, ns, nmix):
        _req_cuda(raw, x)
        n, ch = raw.size(0), raw.size(1)
        hw = raw[0, 0].numel()
        b = x.size(0)
        if ch != 10 * nmix or n != b * ns or x.size(1) != 3 or x[0, 0].numel() != hw:
            raise RuntimeError("dmol: raw %s / x %s inconsistent with ns=%d nmix=%d" % (tuple(raw.shape), tuple(x.shape), ns, nmix))
        raw_c = raw.contiguous()
        x_c = x.contiguous()
        lib = _lib_()
        ws = _pool.take(max(int(lib.mae_dmol_workspace_bytes(n, hw)), 256), raw.device)
        err = torch.empty(b, ns, device=raw.device, dtype=torch.float32)
        _lib.check(lib.mae_dmol_fwd(raw_c.data_ptr(), x_c.data_ptr(), b, ns, nmix, hw, err.data_ptr(), ws.data_ptr(), ws.numel(),
                                    _stream()), "mae_dmol_fwd")
        _count()
        _pool.give(ws)
        ctx.save_for_backward(raw_c, x_c)
        ctx.meta = (b, ns, nmix, hw)
        return err

    @staticmethod
    @_guarded
    def backward(ctx, g):
        raw_c, x_c = ctx.saved_tensors
        b, ns, nmix, hw = ctx.meta
        g = g.contiguous()
        d = torch.empty_like(raw_c)
        _lib.check(_lib_().mae_dmol_bwd(raw_c.data_ptr(), x_c.data_ptr(), g.data_ptr(), 1, 1.0, b, ns, nmix, hw, d.data_ptr(), _stream()),
                   "mae_dmol_bwd")
        _count()
        return d, None, None, None


def dmol_recon_error(raw: torch.Tensor, x: torch.Tensor, ns: int, nmix: int) -> torch.Tensor:
    """raw [B*ns,10*nmix,H,W] decoder output, x [B,3,H,W] in [-1,1] -> reconstruction error [B,ns].
    Reference: color_image_decoder.py:38-52,91-125 + utils.py:61-123."""
    return _Dmol.apply(raw, x, int(ns), int(nmix))


# ------------------------------------------------------------------------------------------------
# N2: final 1x1 convolution + A4b + A4c in one tcgen05 kernel (csrc/dmol_head.cu)
# ------------------------------------------------------------------------------------------------
_dmol_head_mode = os.environ.get("MAE_DMOL_HEAD", "eval")


def set_dmol_head(mode: str) -> str:
    """When the colour decoders fuse their last 1x1 convolution into the likelihood kernel (csrc/dmol_head.cu):
    ``"eval"`` (default) - whenever no gradient is needed (``MAE.nll``, evaluation under ``no_grad``): the K-sample raw tensor
    (210 MB per CIFAR image at K = 512) is never written; ``"always"`` - in training too (the kernel then also emits
    d err / d raw and the convolution's backward is three library contractions); ``"off"`` - never.  Returns the previous mode."""
    global _dmol_head_mode
    if mode not in ("off", "eval", "always"):
        raise ValueError("mode must be 'off', 'eval' or 'always'")
    prev, _dmol_head_mode = _dmol_head_mode, mode
    return prev


def dmol_head_mode() -> str:
    return _dmol_head_mode


def dmol_head_supported(hidden: torch.Tensor, nmix: int) -> bool:
    """True when the fused kernel covers this shape (hidden [N, C, H, W] fp32 on a CUDA device)."""
    if not (hidden.is_cuda and hidden.dtype == torch.float32 and hidden.dim() == 4):
        return False
    return bool(_lib_().mae_dmol_head_supported(hidden.size(1), int(nmix), hidden.size(2) * hidden.size(3)))


@_guarded
def _dmol_head_launch(hidden_c, weight_c, bias_c, x_c, b, ns, nmix, act, g_mul, want_grad):
    n, c = hidden_c.size(0), hidden_c.size(1)
    hw = hidden_c.size(2) * hidden_c.size(3)
    lib = _lib_()
    ws = _pool.take(max(int(lib.mae_dmol_head_workspace_bytes(n, hw)), 256), hidden_c.device)
    err = torch.empty(b, ns, device=hidden_c.device, dtype=torch.float32)
    draw = torch.empty(n, 10 * nmix, hidden_c.size(2), hidden_c.size(3), device=hidden_c.device, dtype=torch.float32) if want_grad else None
    _lib.check(lib.mae_dmol_head(hidden_c.data_ptr(), weight_c.data_ptr(), _ptr(bias_c), x_c.data_ptr(), b, ns, nmix, c, hw,
                                 1 if act else 0, float(g_mul), err.data_ptr(), _ptr(draw), ws.data_ptr(), ws.numel(), _stream()),
               "mae_dmol_head")
    _count()
    _pool.give(ws)
    return err, draw


class _DmolHead(torch.autograd.Function):
    """(hidden, weight, bias, x) -> err [B, ns].  When any input needs a gradient the same launch also emits d err / d raw
    (unit upstream gradient); the convolution's own backward (d hidden, d weight, d bias) is then three library GEMM-shaped
    contractions on it - the raw tensor is never materialised in either direction."""

    @staticmethod
    def forward(ctx, hidden, weight, bias, x, ns, nmix, act):
        _req_cuda(hidden, weight, bias, x)
        n, c = hidden.size(0), hidden.size(1)
        b = x.size(0)
        if hidden.dim() != 4 or weight.shape != (10 * nmix, c) or n != b * ns or x.size(1) != 3 or x.shape[2:] != hidden.shape[2:]:
            raise RuntimeError("dmol_head: hidden %s / weight %s / x %s inconsistent with ns=%d nmix=%d"
                               % (tuple(hidden.shape), tuple(weight.shape), tuple(x.shape), ns, nmix))
        if not dmol_head_supported(hidden, nmix):
            raise RuntimeError("dmol_head: unsupported shape C=%d nmix=%d HW=%d (see mae_dmol_head_supported)"
                               % (c, nmix, hidden.size(2) * hidden.size(3)))
        need = any(ctx.needs_input_grad[:3])
        hidden_c, weight_c, x_c = hidden.contiguous(), weight.contiguous(), x.contiguous()
        bias_c = bias.contiguous() if bias is not None else None
        err, draw = _dmol_head_launch(hidden_c, weight_c, bias_c, x_c, b, ns, nmix, act, 1.0, need)
        if need:
            ctx.save_for_backward(hidden_c, weight_c, draw)
        ctx.meta = (b, ns, nmix, act, bias is not None)
        return err

    @staticmethod
    def backward(ctx, g):
        hidden_c, weight_c, draw = ctx.saved_tensors
        b, ns, nmix, act, has_bias = ctx.meta
        n, c = hidden_c.size(0), hidden_c.size(1)
        graw = (draw.view(n, 10 * nmix, -1) * g.reshape(n, 1, 1)).view_as(draw) if g is not None else draw
        g3 = graw.view(n, 10 * nmix, -1)
        h3 = hidden_c.view(n, c, -1)
        a3 = torch.nn.functional.elu(h3) if act else h3
        gw = gb = gh = None
        if ctx.needs_input_grad[1]:
            gw = torch.einsum("nop,ncp->oc", g3, a3)
        if has_bias and ctx.needs_input_grad[2]:
            gb = g3.sum(dim=(0, 2))
        if ctx.needs_input_grad[0]:
            ga = torch.matmul(weight_c.t().unsqueeze(0), g3)            # [n, c, hw]
            if act:
                ga = ga * torch.where(h3 > 0, torch.ones_like(h3), a3 + 1.0)   # ELU'(h) = 1 (h > 0), exp(h) = ELU(h) + 1
            gh = ga.view_as(hidden_c)
        return gh, gw, gb, None, None, None, None


def dmol_head_recon_error(hidden: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor], x: torch.Tensor,
                          ns: int, nmix: int, act: bool) -> torch.Tensor:
    """Reconstruction error [B, ns] straight from the INPUT of the decoder's last 1x1 convolution.

    hidden [B*ns, C, H, W] (before the ELU in front of that convolution when ``act``), weight [10*nmix, C] / bias [10*nmix]
    the effective (weight-normalised) parameters of the convolution, x [B, 3, H, W] in [-1, 1].
    Reference: image_decoders/pixelcnnpp.py:32-34 / resnet.py:85 followed by color_image_decoder.py:91-125."""
    return _DmolHead.apply(hidden, weight, bias, x, int(ns), int(nmix), bool(act))


# ------------------------------------------------------------------------------------------------
# A5 / A6
# ------------------------------------------------------------------------------------------------
class _LogPrior(torch.autograd.Function):
    @staticmethod
    @_guarded
    def forward(ctx, z, logdet):
        _req_cuda(z, logdet)
        n = z.size(0)
        z_c = z.reshape(n, -1).contiguous()
        d = z_c.size(1)
        ld = logdet.contiguous() if logdet is not None else None
        out = torch.empty(n, device=z.device, dtype=torch.float32)
        _lib.check(_lib_().mae_log_prob_prior_fwd(z_c.data_ptr(), _ptr(ld), n, d, out.data_ptr(), _stream()),
                   "mae_log_prob_prior_fwd")
        _count()
        ctx.save_for_backward(z_c)
        ctx.meta = (z.shape, logdet is not None)
        return out

    @staticmethod
    @_guarded
    def backward(ctx, g):
        (z_c,) = ctx.saved_tensors
        shape, has_ld = ctx.meta
        g = g.contiguous()
        dz = None
        if ctx.needs_input_grad[0]:
            dz = torch.empty_like(z_c)
            _lib.check(_lib_().mae_log_prob_prior_bwd(z_c.data_ptr(), g.data_ptr(), z_c.size(0), z_c.size(1), dz.data_ptr(),
                                                      _stream()), "mae_log_prob_prior_bwd")
            _count()
            dz = dz.view(shape)
        return dz, (g if has_ld else None)


def log_prob_prior(z_normal: torch.Tensor, logdet: Optional[torch.Tensor] = None) -> torch.Tensor:
    """log N(z_normal; 0, I) + logdet, [N].  Reference: gaussian_encoder.py:256-258."""
    return _LogPrior.apply(z_normal, logdet)


class _LogPosterior(torch.autograd.Function):
    @staticmethod
    @_guarded
    def forward(ctx, z, mu, logvar, logdet):
        _req_cuda(z, mu, logvar, logdet)
        b, k = z.size(0), z.size(1)
        z_c = z.reshape(b, k, -1).contiguous()
        d = z_c.size(2)
        mu2, ms = _rows(mu)
        lv2, ls = _rows(logvar)
        ld = logdet.contiguous() if logdet is not None else None
        out = torch.empty(b, k, device=z.device, dtype=torch.float32)
        _lib.check(_lib_().mae_log_prob_posterior_fwd(z_c.data_ptr(), _ptr(ld), mu2.data_ptr(), ms, lv2.data_ptr(), ls, b, k, d,
                                                      out.data_ptr(), _stream()), "mae_log_prob_posterior_fwd")
        _count()
        ctx.save_for_backward(z_c, mu2, lv2)
        ctx.meta = (ms, ls, z.shape, mu.shape, logvar.shape, logdet is not None)
        return out

    @staticmethod
    @_guarded
    def backward(ctx, g):
        z_c, mu2, lv2 = ctx.saved_tensors
        ms, ls, z_shape, mu_shape, lv_shape, has_ld = ctx.meta
        b, k, d = z_c.shape
        g = g.contiguous()
        dz = torch.empty_like(z_c)
        dmu = torch.empty(b, d, device=z_c.device, dtype=torch.float32)
        dlv = torch.empty_like(dmu)
        _lib.check(_lib_().mae_log_prob_posterior_bwd(z_c.data_ptr(), g.data_ptr(), mu2.data_ptr(), ms, lv2.data_ptr(), ls, b, k, d,
                                                      dz.data_ptr(), dmu.data_ptr(), dlv.data_ptr(), _stream()),
                   "mae_log_prob_posterior_bwd")
        _count()
        return dz.view(z_shape), dmu.view(mu_shape), dlv.view(lv_shape), (g if has_ld else None)


def log_prob_posterior(z_normal: torch.Tensor, mu: torch.Tensor, logvar: torch.Tensor,
                       logdet: Optional[torch.Tensor] = None) -> torch.Tensor:
    """log q(z|x) for z_normal [B,K,*z], [B,K].  Reference: gaussian_encoder.py:297-299."""
    return _LogPosterior.apply(z_normal, mu, logvar, logdet)


# ------------------------------------------------------------------------------------------------
# A5 + A6 + A7 fused (evaluation only: outputs carry no autograd graph, like calc_nll's no_grad use)
# ------------------------------------------------------------------------------------------------
@_guarded
@torch.no_grad()
def iw_nll(log_gen: torch.Tensor, z_prior_normal: torch.Tensor, logdet_prior: Optional[torch.Tensor],
           z_post_normal: torch.Tensor, logdet_post: Optional[torch.Tensor], mu: torch.Tensor, logvar: torch.Tensor,
           gen_scale: float = 1.0):
    """Returns ((nll_elbo, recon, kl), nll_iw), each [B].  ``gen_scale=-1`` takes the reconstruction ERROR [B,K] in place of
    log p(x|z) (no negation pass).  Reference: mae.py:203-215 + utils.py:23-38."""
    _req_cuda(log_gen, z_prior_normal, logdet_prior, z_post_normal, logdet_post, mu, logvar)
    b, k = log_gen.shape
    zq = z_post_normal.reshape(b, k, -1).contiguous()
    zp = z_prior_normal.reshape(b, k, -1).contiguous()
    d = zq.size(2)
    mu2, ms = _rows(mu)
    lv2, ls = _rows(logvar)
    lg = log_gen.contiguous()
    ldp = logdet_prior.reshape(b, k).contiguous() if logdet_prior is not None else None
    ldq = logdet_post.reshape(b, k).contiguous() if logdet_post is not None else None
    out = torch.empty(4, b, device=lg.device, dtype=torch.float32)
    lib = _lib_()
    ws = _pool.take(max(int(lib.mae_iw_nll_workspace_bytes(b, k)), 256), lg.device)
    _lib.check(lib.mae_iw_nll(lg.data_ptr(), zp.data_ptr(), _ptr(ldp), zq.data_ptr(), _ptr(ldq), mu2.data_ptr(), ms,
                              lv2.data_ptr(), ls, b, k, d, float(gen_scale), out[0].data_ptr(), out[1].data_ptr(), out[2].data_ptr(),
                              out[3].data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "mae_iw_nll")
    _count()
    _pool.give(ws)
    return (out[0], out[1], out[2]), out[3]


# ------------------------------------------------------------------------------------------------
# A3' + A8
# ------------------------------------------------------------------------------------------------
class _Assemble(torch.autograd.Function):
    @staticmethod
    @_guarded
    def forward(ctx, err, kl, m_d, s, eta, gamma, free_bits, inv_batch, ext_sums):
        ctx.set_materialize_grads(False)
        _req_cuda(err, kl, m_d, s, ext_sums)
        join_side(err.device)  # MPD results may still be in flight on the side stream
        b, ns = err.shape
        err_c, kl_c = err.contiguous(), kl.contiguous()
        m_c = m_d.contiguous().view(-1) if m_d is not None else None
        d = m_c.numel() if m_c is not None else 0
        scalars = torch.empty(6, device=err.device, dtype=torch.float32)
        recon = torch.empty(b, device=err.device, dtype=torch.float32)
        _lib.check(_lib_().mae_loss_assemble_fwd(err_c.data_ptr(), kl_c.data_ptr(), _ptr(m_c), _ptr(s), _ptr(ext_sums), b, ns, d, 0, 1, eta, gamma,
                                                 free_bits, inv_batch, scalars.data_ptr(), recon.data_ptr(), _stream()),
                   "mae_loss_assemble_fwd")
        _count()
        ctx.save_for_backward(scalars, m_c if m_c is not None else scalars)
        ctx.meta = (b, ns, d, eta, gamma, free_bits, inv_batch, m_d.shape if m_d is not None else None, s is not None)
        loss = scalars[0]
        stats = scalars[1:]
        ctx.mark_non_differentiable(stats, recon)
        return loss, stats, recon

    @staticmethod
    @_guarded
    def backward(ctx, gloss, _gstats, _grecon):
        scalars, m_c = ctx.saved_tensors
        b, ns, d, eta, gamma, free_bits, inv_batch, m_shape, has_s = ctx.meta
        dev = scalars.device
        if gloss is None:
            return (None,) * 9
        gloss = gloss.contiguous()
        derr = torch.empty(b, ns, device=dev, dtype=torch.float32)
        dkl = torch.empty(b, device=dev, dtype=torch.float32)
        dm = torch.empty(d, device=dev, dtype=torch.float32) if m_shape is not None else None
        ds = torch.empty((), device=dev, dtype=torch.float32) if has_s else None
        _lib.check(_lib_().mae_loss_assemble_bwd(gloss.data_ptr(), scalars.data_ptr(), _ptr(m_c) if m_shape is not None else None,
                                                 b, ns, d, eta, gamma, free_bits, inv_batch, derr.data_ptr(), dkl.data_ptr(),
                                                 _ptr(dm), _ptr(ds), _stream()), "mae_loss_assemble_bwd")
        _count()
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(dev))
        _grads_ready[dev.index if dev.index is not None else torch.cuda.current_device()] = (
            ev, {t.data_ptr() for t in (dm, ds, dkl) if t is not None})
        return derr, dkl, (dm.view(m_shape) if dm is not None else None), ds, None, None, None, None, None


def _likelihood_fwd(dec_out, x, kind, ns, nmix, inv_batch, need_grad):
    """Likelihood kernel of the fused loss nodes on the current stream.  Returns (out_c, x_c, err, err_chunks,
    fused_grad, geom); ``fused_grad`` is (1/(B ns)) d err / d dec_out when the one-pass forward+backward DMOL kernel ran."""
    lib = _lib_()
    dev = dec_out.device
    b = x.size(0)
    out_c = dec_out.contiguous()
    fused_grad = None
    if kind == "dmol":
        hw = out_c[0, 0].numel()
        if out_c.size(1) != 10 * nmix or out_c.size(0) != b * ns or x.size(1) != 3 or x[0, 0].numel() != hw:
            raise RuntimeError("dmol: raw %s / x %s inconsistent with ns=%d nmix=%d" % (tuple(out_c.shape), tuple(x.shape), ns, nmix))
        x_c = x.contiguous()
        err_chunks = 1
        if need_grad and nmix == 10 and hw in (1024, 4096):
            # training: one pass produces the per-CTA partial sums of err AND (1/(B ns)) d err/d raw; the objective
            # kernel adds the partials, the backward only applies d loss
            err_chunks = int(lib.mae_dmol_err_chunks(b, ns, nmix, hw))
            err = torch.empty(b, ns, err_chunks, device=dev, dtype=torch.float32)
            fused_grad = torch.empty_like(out_c)
            gmul = (inv_batch if inv_batch > 0.0 else 1.0 / float(b)) / float(ns)
            _lib.check(lib.mae_dmol_fwd_bwd(out_c.data_ptr(), x_c.data_ptr(), gmul, b, ns, nmix, hw, err.data_ptr(),
                                            fused_grad.data_ptr(), _stream()), "mae_dmol_fwd_bwd")
        else:
            ws = _pool.take(max(int(lib.mae_dmol_workspace_bytes(b * ns, hw)), 256), dev)
            err = torch.empty(b, ns, device=dev, dtype=torch.float32)
            _lib.check(lib.mae_dmol_fwd(out_c.data_ptr(), x_c.data_ptr(), b, ns, nmix, hw, err.data_ptr(), ws.data_ptr(),
                                        ws.numel(), _stream()), "mae_dmol_fwd")
            _pool.give(ws)
        geom = (hw,)
    else:
        out_c = out_c.reshape(b, ns, -1)
        p = out_c.size(2)
        x_c = x.reshape(b, -1).contiguous()
        if x_c.size(1) != p:
            raise RuntimeError("x has %d pixels per image, probabilities have %d" % (x_c.size(1), p))
        err = torch.empty(b, ns, device=dev, dtype=torch.float32)
        err_chunks = 1
        _lib.check(lib.mae_bernoulli_fwd(out_c.data_ptr(), x_c.data_ptr(), b, ns, p, err.data_ptr(), _stream()),
                   "mae_bernoulli_fwd")
        geom = (p,)
    _count()
    return out_c, x_c, err, err_chunks, fused_grad, geom


_unit_loss_grad = False


def set_unit_loss_grad(enabled: bool) -> bool:
    """Promise that every ``backward()`` through the fused loss nodes starts from d loss = 1 (``loss.backward()`` without an
    explicit gradient and without scaling the loss first, as in experiments/color_image.py:158).  The fused DMOL node
    produced (1/B) d err / d raw in the forward pass already; with the promise its backward is a no-op instead of a
    conditional rescale launch that would find the scale to be 1.  Used by ``mae_b200.graph.capture_step`` and bench.py,
    which create that gradient themselves.  Returns the previous setting.  NOT safe with loss scaling (AMP) or when the
    loss enters a larger expression before ``backward()``."""
    global _unit_loss_grad
    prev = _unit_loss_grad
    _unit_loss_grad = bool(enabled)
    return prev


def _likelihood_bwd(ctx, out_c, x_c, gloss, kind, b, ns, nmix, geom, inv_batch, out_shape):
    """d loss / d dec_out on the current stream (gloss: device scalar d loss)."""
    lib = _lib_()
    if ctx.fused_grad is not None:
        if ctx.fused_grad is False:
            raise RuntimeError("fused loss node: the likelihood gradient was already consumed (retain_graph with a second "
                               "backward is not supported on this path)")
        dout = ctx.fused_grad
        ctx.fused_grad = False
        if _unit_loss_grad:
            return dout.view(out_shape)       # d loss == 1 by the caller's promise: nothing to launch
        _lib.check(lib.mae_scale_inplace(dout.data_ptr(), dout.numel(), gloss.data_ptr(), _stream()), "mae_scale_inplace")
    else:
        dout = torch.empty_like(out_c)
        gmul = (inv_batch if inv_batch > 0.0 else 1.0 / float(b)) / float(ns)
        if kind == "dmol":
            _lib.check(lib.mae_dmol_bwd(out_c.data_ptr(), x_c.data_ptr(), gloss.data_ptr(), 0, gmul, b, ns, nmix, geom[0],
                                        dout.data_ptr(), _stream()), "mae_dmol_bwd")
        else:
            _lib.check(lib.mae_bernoulli_bwd(out_c.data_ptr(), x_c.data_ptr(), gloss.data_ptr(), 0, gmul, b, ns, geom[0],
                                             dout.data_ptr(), _stream()), "mae_bernoulli_bwd")
    _count()
    return dout.view(out_shape)


class _LossTail(torch.autograd.Function):
    """likelihood + objective assembly as ONE autograd node (the part of MAE.loss after the decoder):
    forward = likelihood kernel, join of the MPD side stream, assembly kernel; backward = likelihood backward on the
    main stream (taking d loss as a broadcast device scalar, so it does not wait for anything else) and, concurrently
    on the side stream, the tiny assembly backward that feeds the KL / MPD backward.

    ``defer``: the assembly kernels run on the side stream behind the MPD chain, so the main stream goes straight
    from the likelihood forward to the likelihood backward; the returned values are valid on the main stream only
    after ``backward()`` (which joins) or ``ops.join_side()``."""

    @staticmethod
    @_guarded
    def forward(ctx, dec_out, x, kl, m_d, s, ext_sums, kl_all, kind, ns, nmix, eta, gamma, free_bits, inv_batch, defer):
        ctx.set_materialize_grads(False)
        _req_cuda(dec_out, x, kl, m_d, s, ext_sums, kl_all)
        lib = _lib_()
        dev = dec_out.device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        b = x.size(0)
        out_c, x_c, err, err_chunks, fused_grad, geom = _likelihood_fwd(dec_out, x, kind, ns, nmix, inv_batch,
                                                                         ctx.needs_input_grad[0])
        kl_c = (kl_all if kl_all is not None else kl).contiguous()   # sharded runs sum the all-gathered KL vector
        kl_count = kl_c.numel() if kl_all is not None else 0
        m_c = m_d.contiguous().view(-1) if m_d is not None else None
        d = m_c.numel() if m_c is not None else 0
        scalars = torch.empty(6, device=dev, dtype=torch.float32)
        recon = torch.empty(b, device=dev, dtype=torch.float32)

        def assemble():
            _lib.check(lib.mae_loss_assemble_fwd(err.data_ptr(), kl_c.data_ptr(), _ptr(m_c), _ptr(s), _ptr(ext_sums), b, ns, d,
                                                 kl_count, err_chunks, eta, gamma, free_bits, inv_batch, scalars.data_ptr(),
                                                 recon.data_ptr(), _stream()), "mae_loss_assemble_fwd")

        if defer:
            cs, ss = torch.cuda.current_stream(dev), _side_stream(dev)
            ev_err = torch.cuda.Event()
            ev_err.record(cs)
            prev = _pending_join.pop(idx, None)      # the MPD chain already sits on the side stream: ordered by it
            with torch.cuda.stream(ss):
                ss.wait_event(ev_err)
                assemble()
                ev = torch.cuda.Event()
                ev.record(ss)
            _pending_join[idx] = (ev, (prev, err, kl_c, m_c, s, ext_sums, scalars, recon))
        else:
            join_side(dev)
            assemble()
        _count()
        ctx.save_for_backward(out_c, x_c, scalars, m_c if m_c is not None else scalars)
        ctx.fused_grad = fused_grad
        ctx.meta = (kind, b, ns, nmix, geom, d, eta, gamma, free_bits, inv_batch, dec_out.shape,
                    m_d.shape if m_d is not None else None, s is not None, defer)
        loss, stats = scalars[0], scalars[1:]
        ctx.mark_non_differentiable(stats, recon)
        return loss, stats, recon

    @staticmethod
    @_guarded
    def backward(ctx, gloss, _gs, _gr):
        if gloss is None:
            return (None,) * 15
        out_c, x_c, scalars, m_c = ctx.saved_tensors
        kind, b, ns, nmix, geom, d, eta, gamma, free_bits, inv_batch, out_shape, m_shape, has_s, defer = ctx.meta
        lib = _lib_()
        dev = out_c.device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        gloss = gloss.contiguous()
        cs = torch.cuda.current_stream(dev)
        ss = _side_stream(dev)
        if defer:
            ev_g = torch.cuda.Event()
            ev_g.record(cs)
            ss.wait_event(ev_g)                 # d loss exists; the side stream still holds the deferred forward
            _pending_join.pop(idx, None)        # joined below through the assembly backward's event
        else:
            ss.wait_stream(cs)                  # d loss is ready on the main stream
        dout = None
        if ctx.needs_input_grad[0]:
            dout = _likelihood_bwd(ctx, out_c, x_c, gloss, kind, b, ns, nmix, geom, inv_batch, out_shape)
        dkl = torch.empty(b, device=dev, dtype=torch.float32)
        dm = torch.empty(d, device=dev, dtype=torch.float32) if m_shape is not None else None
        ds = torch.empty((), device=dev, dtype=torch.float32) if has_s else None
        with torch.cuda.stream(ss):
            _lib.check(lib.mae_loss_assemble_bwd(gloss.data_ptr(), scalars.data_ptr(), _ptr(m_c) if m_shape is not None else None,
                                                 b, ns, d, eta, gamma, free_bits, inv_batch, None, dkl.data_ptr(), _ptr(dm),
                                                 _ptr(ds), _stream()), "mae_loss_assemble_bwd")
            _count()
            ev = torch.cuda.Event()
            ev.record(ss)
        _grads_ready[idx] = (ev, {t.data_ptr() for t in (dm, ds, dkl) if t is not None})
        cs.wait_event(ev)                       # later main-stream consumers see dkl / dm / dS (and, when deferred, the
        #                                         forward's loss); the likelihood backward enqueued above is not delayed
        return (dout, None, dkl, (dm.view(m_shape) if dm is not None else None), ds, None, None, None, None, None, None, None,
                None, None, None)


def loss_tail(dec_out: torch.Tensor, x: torch.Tensor, kl: torch.Tensor, m_d: Optional[torch.Tensor], s: Optional[torch.Tensor],
              kind: str, ns: int, nmix: int = 10, eta: float = 0.0, gamma: float = 0.0, free_bits: float = 0.0,
              inv_batch: float = 0.0, ext_sums: Optional[torch.Tensor] = None, defer_join: bool = False,
              kl_all: Optional[torch.Tensor] = None):
    """Decoder output -> objective of MAE.loss in one autograd node.  ``kind`` is "dmol" (dec_out = raw
    [B*ns,10*nmix,H,W]) or "bernoulli" (dec_out = probabilities [B,ns,P]).  Returns (loss [], stats [5], recon [B])
    like :func:`assemble_loss`.  ``inv_batch`` / ``ext_sums`` serve batch-sharded runs (see assemble_loss).
    ``kl_all`` (sharded runs): the all-gathered KL vector whose mean enters the free-bits clamp (``kl`` stays the local,
    differentiable slice).  ``defer_join=True`` keeps the objective's tiny kernels on the side stream (see _LossTail); the outputs may then be
    read only after ``backward()`` or :func:`join_side`.
    Reference: color_image_decoder.py:91-125 / binary_image_decoder.py:84-93 + mae.py:132-140."""
    if kind not in ("dmol", "bernoulli"):
        raise ValueError("kind must be 'dmol' or 'bernoulli'")
    return _LossTail.apply(dec_out, x, kl, m_d, s, ext_sums, kl_all, kind, int(ns), int(nmix), float(eta), float(gamma),
                           float(free_bits), float(inv_batch), bool(defer_join))


_fused_reg_ok: Dict[Tuple[int, int, int], bool] = {}


def fused_reg_supported(b: int, d: int, device: torch.device) -> bool:
    """True when (B, D) can take the one-launch KL + MPD forward/backward (mae_kl_mpd_loss_fused) on this device."""
    idx = device.index if device.index is not None else torch.cuda.current_device()
    key = (idx, int(b), int(d))
    ok = _fused_reg_ok.get(key)
    if ok is None:
        with torch.cuda.device(idx):
            ok = bool(_lib_().mae_kl_mpd_loss_fused_supported(int(b), int(d)))
        _fused_reg_ok[key] = ok
    return ok


@_guarded
def kl_mpd_reg_fused(mu: torch.Tensor, logvar: torch.Tensor, eta: float, gamma: float, free_bits: float,
                     kl_weight: Optional[float] = None, want_grad: bool = True, launch_stream=None):
    """Raw (non-autograd) call of the one-launch regulariser kernel on the current stream (or on ``launch_stream``, which
    must already be ordered after the producers of mu / logvar; the outputs stay owned by the current stream).

    Returns (kl [B], m_d [*z], S [], reg [], grads, keep) with
    reg = clamp(kl_weight * sum KL, min=free_bits) - eta * sum logsigmoid(m_d) + gamma * S   (mae.py:132-139)
    and grads = [2,B,D] holding d reg / d mu and d reg / d logvar (None when ``want_grad`` is False).
    ``kl_weight`` defaults to 1/B (KL.mean())."""
    _req_cuda(mu, logvar)
    mu2, ms = _rows(mu)
    lv2, ls = _rows(logvar)
    b, d = mu2.shape
    dev = mu2.device
    if not fused_reg_supported(b, d, dev):
        raise RuntimeError("kl_mpd_reg_fused: B=%d, D=%d is outside the fused path (B <= 128, D <= 64)" % (b, d))
    kw = 1.0 / float(b) if kl_weight is None else float(kl_weight)
    m_d = torch.empty(d, device=dev, dtype=torch.float32)
    s_ = torch.empty((), device=dev, dtype=torch.float32)
    reg = torch.empty((), device=dev, dtype=torch.float32)
    kl = torch.empty(b, device=dev, dtype=torch.float32)
    grads = torch.empty(2, b, d, device=dev, dtype=torch.float32) if want_grad else None
    with torch.cuda.stream(launch_stream if launch_stream is not None else torch.cuda.current_stream(dev)):
        _lib.check(_lib_().mae_kl_mpd_loss_fused(mu2.data_ptr(), ms, lv2.data_ptr(), ls, b, d, float(eta), float(gamma),
                                                 float(free_bits), kw, m_d.data_ptr(), s_.data_ptr(), kl.data_ptr(),
                                                 reg.data_ptr(), grads[0].data_ptr() if want_grad else None,
                                                 grads[1].data_ptr() if want_grad else None, _stream()),
                   "mae_kl_mpd_loss_fused")
    _count()
    return kl, m_d.view(mu.shape[1:]), s_, reg, grads, (mu2, lv2)


@_guarded
def kl_mpd_reg(mu: torch.Tensor, logvar: torch.Tensor, eta: float, gamma: float, free_bits: float,
               kl_weight: Optional[float] = None, want_grad: bool = True, launch_stream=None):
    """Same contract as :func:`kl_mpd_reg_fused` for ANY batch / latent size: the one-launch cluster kernel where it
    applies, else the general chain prepare -> tiles -> (device-side upstream gradients -> backward kernel), all enqueued
    back to back (on ``launch_stream``), so that nothing of the regulariser is left for the backward pass."""
    _req_cuda(mu, logvar)
    mu2, ms = _rows(mu)
    lv2, ls = _rows(logvar)
    b, d = mu2.shape
    dev = mu2.device
    if b < 2:
        raise RuntimeError("the MPD regulariser needs a batch of at least 2 posteriors (gaussian_encoder.py:181)")
    if fused_reg_supported(b, d, dev):
        return kl_mpd_reg_fused(mu, logvar, eta, gamma, free_bits, kl_weight, want_grad, launch_stream)
    kw = 1.0 / float(b) if kl_weight is None else float(kl_weight)
    lib = _lib_()
    path = _mpd_pick_path(b, d, 0) if want_grad else 2
    store = want_grad and path != 2
    outs = _mpd_alloc_outputs(d, dev)
    kl = torch.empty(b, device=dev, dtype=torch.float32)
    with torch.cuda.stream(launch_stream if launch_stream is not None else torch.cuda.current_stream(dev)):
        ws = _pool.take(lib.mae_mpd_workspace_bytes(b, d, int(store)), dev)
        m_d, s_, _ = _mpd_launch_fwd(mu2, ms, lv2, ls, b, d, ws, store, outs, kl=kl)
    state = MpdRowsState(ws, mu2, ms, lv2, ls, b, d, s_, path)
    # outputs are allocated on the current stream, the kernels go to the launch stream
    grads = torch.empty(2, b, d, device=dev, dtype=torch.float32) if want_grad else None
    reg = torch.empty((), device=dev, dtype=torch.float32)
    with torch.cuda.stream(launch_stream if launch_stream is not None else torch.cuda.current_stream(dev)):
        _lib.check(lib.mae_mpd_loss_bwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, ws.data_ptr(), ws.numel(), b, d, 0, b,
                                        float(eta), float(gamma), float(free_bits), kw, m_d.data_ptr(), s_.data_ptr(),
                                        kl.data_ptr(), reg.data_ptr(), grads[0].data_ptr() if want_grad else None,
                                        grads[1].data_ptr() if want_grad else None, state.path, _stream()), "mae_mpd_loss_bwd")
        _count(2 if want_grad else 1)
        _pool.give(ws)
    # ``keep``: every buffer the launch stream's kernels touch that is not among the results (the tile kernel writes its fp64
    # sum into outs[2]).  The caller must hold it until the launch stream has been joined: a block freed earlier goes back to
    # the CURRENT stream's free list and is handed to the next small allocation (e.g. a weight-norm temporary of the decoder)
    # while the side-stream kernel that writes it is still pending.
    return kl, m_d.view(mu.shape[1:]), s_, reg, grads, (mu2, lv2, outs)


class _EncodeHead(torch.autograd.Function):
    """Everything between the encoder core and the decoder as ONE autograd node for training-sized batches:
    (mu, logvar, eps) -> (z, reg, KL, m_d, S), reg = clamp(KL.mean(), free_bits) - eta sum logsigmoid(m_d) + gamma S.

    forward  main stream: the reparameterisation (z is what the decoder waits for); side stream: mae_kl_mpd_loss_fused,
             i.e. KL, MPD, reg AND d reg / d(mu, logvar) in one launch, overlapping with the decoder.
    backward ONE launch: d(mu, logvar) = reparam backward of d z  +  d reg * (stored regulariser gradient)."""

    @staticmethod
    @_guarded
    def forward(ctx, mu, logvar, eps, eta, gamma, free_bits, with_kl, holder=None):
        ctx.set_materialize_grads(False)
        _req_cuda(mu, logvar, eps)
        dev = mu.device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        b, ns = eps.size(0), eps.size(1)
        want_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        cs, ss = torch.cuda.current_stream(dev), _side_stream(dev)
        ss.wait_stream(cs)
        prev = _pending_join.pop(idx, None)
        mu2, ms = _rows(mu)
        lv2, ls = _rows(logvar)
        d = mu2.size(1)
        eps_c = eps.contiguous()
        z = torch.empty_like(eps_c)
        z_side = holder is not None and holder.get("z_on_side", False)

        def reparam():
            _lib.check(_lib_().mae_reparam_kl_fwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, eps_c.data_ptr(), b, ns, d,
                                                  z.data_ptr(), None, _stream()), "mae_reparam_kl_fwd")
            _count()

        kl, m_d, s, reg, grads, keep = kl_mpd_reg(mu2, lv2, eta, gamma, free_bits, None if with_kl else 0.0, want_grad,
                                                  launch_stream=ss)
        m_d = m_d.view(mu.shape[1:])
        if z_side:
            # the caller's next main-stream kernel does not consume z (heads.*: the decoder output is an input): z is
            # produced on the side stream BEHIND the regulariser (whose cluster must be placed before the likelihood kernel
            # fills the SMs: with the reparameterisation first the step measured 30.8 us instead of 28.6) and joined by
            # loss_head after it has enqueued the likelihood kernel
            with torch.cuda.stream(ss):
                reparam()
            holder["z_event"] = True
        ev = torch.cuda.Event()
        ev.record(ss)
        if z_side:
            holder["z_event"] = ev
        _pending_join[idx] = (ev, (prev, kl, m_d, s, reg, grads, mu2, lv2, z, eps_c, keep))
        if not z_side:
            reparam()
        ctx.save_for_backward(mu2, lv2, eps_c)
        ctx.reg_grads = grads
        ctx.reg_event = ev
        ctx.meta = (ms, ls, b, ns, d, mu.shape, logvar.shape)
        ctx.mark_non_differentiable(kl, m_d, s)
        return z, reg, kl, m_d, s

    @staticmethod
    @_guarded
    def backward(ctx, gz, g_reg, *_unused):
        if gz is None and g_reg is None:
            return (None,) * 8
        mu2, lv2, eps_c = ctx.saved_tensors
        ms, ls, b, ns, d, mu_shape, lv_shape = ctx.meta
        dev = mu2.device
        grads = ctx.reg_grads if g_reg is not None else None
        if g_reg is not None and grads is None:
            raise RuntimeError("encode_head: reg has a gradient but mu / logvar did not require one in the forward")
        gz = gz.contiguous() if gz is not None else None
        g_reg = g_reg.contiguous() if g_reg is not None else None
        torch.cuda.current_stream(dev).wait_event(ctx.reg_event)   # the regulariser kernel ran on the side stream
        dmu = torch.empty(b, d, device=dev, dtype=torch.float32)
        dlv = torch.empty_like(dmu)
        _lib.check(_lib_().mae_reparam_reg_bwd(mu2.data_ptr(), ms, lv2.data_ptr(), ls, eps_c.data_ptr(), _ptr(gz), None,
                                               grads[0].data_ptr() if grads is not None else None,
                                               grads[1].data_ptr() if grads is not None else None, _ptr(g_reg), b, ns, d,
                                               dmu.data_ptr(), dlv.data_ptr(), 0, _stream()), "mae_reparam_reg_bwd")
        _count()
        return dmu.view(mu_shape), dlv.view(lv_shape), None, None, None, None, None, None


class EncodedBatch:
    """What :func:`encode_head` hands to :func:`loss_head`: reg (the only differentiable entry), KL [B], m_d, S and the
    hyper-parameters they were computed with.  Valid on the main stream after :func:`join_side` (loss_head does it)."""

    def __init__(self, reg, kl, m_d, s, params, z_event=None):
        self.reg, self.kl, self.m_d, self.s, self.params = reg, kl, m_d, s, params
        self.z_event = z_event      # z was produced on the side stream: loss_head joins it behind the likelihood kernel


def encode_head(mu: torch.Tensor, logvar: torch.Tensor, eps: torch.Tensor, eta: float = 0.0, gamma: float = 0.0,
                free_bits: float = 0.0, with_kl: bool = True, z_on_side: bool = False):
    """z [B,ns,*z] and the encoder-side regulariser of MAE.loss in one autograd node; -> (z, EncodedBatch).
    ``with_kl=False`` (encoders with flows: KL is a Monte-Carlo estimate computed by the caller, gaussian_encoder.py:220-225):
    reg holds the MPD terms only and :func:`loss_head` takes the KL through its ``kl`` argument.
    ``z_on_side=True`` (only for callers whose next main-stream work does NOT consume z, i.e. ``heads.*`` where the decoder
    output is an input): z is produced on the side stream too and becomes valid on the main stream when :func:`loss_head`
    returns (or after :func:`join_side`).
    Reference: gaussian_encoder.py:56-62, :162-188, :219 and mae.py:132-139."""
    holder = {"z_on_side": bool(z_on_side)}
    z, reg, kl, m_d, s = _EncodeHead.apply(mu, logvar, eps, float(eta), float(gamma), float(free_bits), bool(with_kl), holder)
    return z, EncodedBatch(reg, kl if with_kl else None, m_d, s, (float(eta), float(gamma), float(free_bits)),
                           holder.get("z_event"))


class _LossTailReg(torch.autograd.Function):
    """(decoder output, x, reg) -> objective of MAE.loss (mae.py:100-140) when the encoder side came from
    :func:`encode_head`: likelihood kernel on the main stream (for the DMOL head the one-pass forward+backward kernel),
    assembly kernel once both the likelihood's partial sums and the side stream's KL / m_d / S exist.
    backward: d dec_out (a conditional rescale by d loss) and d reg = d loss.  ``defer``: see _LossTail."""

    @staticmethod
    @_guarded
    def forward(ctx, dec_out, x, reg, kl, m_d, s, kind, ns, nmix, eta, gamma, free_bits, defer, inv_batch, ext_sums=None):
        ctx.set_materialize_grads(False)
        _req_cuda(dec_out, x, reg, kl, m_d, s, ext_sums)
        kl_ext = ctx.needs_input_grad[3]      # a differentiable KL from the caller (Monte-Carlo estimate with flows)
        kl = kl.contiguous()
        lib = _lib_()
        dev = dec_out.device
        idx = dev.index if dev.index is not None else torch.cuda.current_device()
        b = x.size(0)
        out_c, x_c, err, err_chunks, fused_grad, geom = _likelihood_fwd(dec_out, x, kind, ns, nmix, inv_batch,
                                                                         ctx.needs_input_grad[0])
        kl_count = kl.numel() if inv_batch > 0.0 else 0     # sharded runs: kl is the all-gathered vector of the global batch
        m_c = m_d.reshape(-1)
        d = m_c.numel()
        scalars = torch.empty(6, device=dev, dtype=torch.float32)
        recon = torch.empty(b, device=dev, dtype=torch.float32)

        def assemble():
            _lib.check(lib.mae_loss_assemble_fwd(err.data_ptr(), kl.data_ptr(), m_c.data_ptr(), s.data_ptr(), _ptr(ext_sums), b, ns, d,
                                                 kl_count, err_chunks, eta, gamma, free_bits, inv_batch, scalars.data_ptr(),
                                                 recon.data_ptr(), _stream()), "mae_loss_assemble_fwd")
            _count()

        if defer:
            cs, ss = torch.cuda.current_stream(dev), _side_stream(dev)
            ev_err = torch.cuda.Event()
            ev_err.record(cs)
            prev = _pending_join.pop(idx, None)      # the regulariser kernel already sits on the side stream
            with torch.cuda.stream(ss):
                ss.wait_event(ev_err)
                assemble()
                ev = torch.cuda.Event()
                ev.record(ss)
            _pending_join[idx] = (ev, (prev, err, kl, m_c, s, scalars, recon, ext_sums))
        else:
            join_side(dev)
            assemble()
        ctx.save_for_backward(out_c, x_c, scalars)
        ctx.fused_grad = fused_grad
        ctx.meta = (kind, b, ns, nmix, geom, dec_out.shape, defer, inv_batch, kl_ext, (eta, gamma, free_bits))
        loss, stats = scalars[0], scalars[1:]
        ctx.mark_non_differentiable(stats, recon)
        return loss, stats, recon

    @staticmethod
    @_guarded
    def backward(ctx, gloss, *_unused):
        if gloss is None:
            return (None,) * 15
        out_c, x_c, scalars = ctx.saved_tensors
        kind, b, ns, nmix, geom, out_shape, defer, inv_batch, kl_ext, (eta, gamma, free_bits) = ctx.meta
        gloss = gloss.contiguous()
        dout = None
        if ctx.needs_input_grad[0]:
            dout = _likelihood_bwd(ctx, out_c, x_c, gloss, kind, b, ns, nmix, geom, inv_batch, out_shape)
        dkl = None
        if kl_ext:
            # d loss / d KL_b = d loss * [mean KL >= free_bits] / B: needs the forward's scalars (side stream when deferred)
            dev = out_c.device
            if defer:
                join_side(dev)
            dkl = torch.empty(b, device=dev, dtype=torch.float32)
            _lib.check(_lib_().mae_loss_assemble_bwd(gloss.data_ptr(), scalars.data_ptr(), None, b, ns, 0, eta, gamma, free_bits,
                                                     inv_batch, None, dkl.data_ptr(), None, None, _stream()),
                       "mae_loss_assemble_bwd")
            _count()
        if defer:
            # nothing in the backward pass consumes the deferred objective kernel: join once the whole pass has been
            # enqueued, so that the encoder-side backward is not ordered behind it
            dev = out_c.device
            torch.autograd.Variable._execution_engine.queue_callback(lambda: join_side(dev))
        return (dout, None, gloss if ctx.needs_input_grad[2] else None, dkl) + (None,) * 11


def loss_head(dec_out: torch.Tensor, x: torch.Tensor, enc: EncodedBatch, kind: str, ns: int, nmix: int = 10,
              eta: float = 0.0, gamma: float = 0.0, free_bits: float = 0.0, defer_join: bool = False,
              inv_batch: float = 0.0, kl: Optional[torch.Tensor] = None, ext_sums: Optional[torch.Tensor] = None):
    """Decoder output + :func:`encode_head` result -> objective of MAE.loss.  Returns (loss [], stats [5], recon [B]) like
    :func:`loss_tail`; KL / m_d / S are in ``enc``.  ``defer_join``: see :func:`loss_tail`.  ``inv_batch`` (batch-sharded
    runs): 1 / global batch; ``enc.kl`` is then the all-gathered KL vector and the returned loss holds this rank's share of
    the reconstruction term plus the global regulariser.  ``kl`` [B] (differentiable): the caller's KL estimate, required
    when ``enc`` was made with ``with_kl=False``; it enters the free-bits clamp and receives its gradient from this node.
    ``ext_sums`` (sharded runs with such a KL): device float[2] = (recon, KL) sums of the OTHER ranks, added before the means.
    Reference: mae.py:100-140 with the two reconstruct_error heads."""
    if kind not in ("dmol", "bernoulli"):
        raise ValueError("kind must be 'dmol' or 'bernoulli'")
    if enc.params != (float(eta), float(gamma), float(free_bits)):
        raise RuntimeError("loss_head: encode_head was called with different eta / gamma / free_bits")
    if (kl is None) == (enc.kl is None):
        raise RuntimeError("loss_head: pass kl exactly when encode_head was called with with_kl=False")
    out = _LossTailReg.apply(dec_out, x, enc.reg, enc.kl if kl is None else kl, enc.m_d, enc.s, kind, int(ns), int(nmix),
                             float(eta), float(gamma), float(free_bits), bool(defer_join), float(inv_batch), ext_sums)
    if enc.z_event is not None:      # z came from the side stream: valid on the main stream from here on
        torch.cuda.current_stream(dec_out.device).wait_event(enc.z_event)
        enc.z_event = None
    return out


def assemble_loss(err: torch.Tensor, kl: torch.Tensor, m_d: Optional[torch.Tensor], s: Optional[torch.Tensor],
                  eta: float = 0.0, gamma: float = 0.0, free_bits: float = 0.0, inv_batch: float = 0.0,
                  ext_sums: Optional[torch.Tensor] = None):
    """Objective of MAE.loss (mae.py:132-140) in one launch.

    Returns (loss [], stats [5] = (postKL_mean, loss_postKL_mean, loss_postKL_std, mean KL, mean recon), recon [B]).
    Only ``loss`` carries a gradient; ``stats`` and ``recon`` are logging outputs.  ``inv_batch`` (1/global batch)
    and ``ext_sums`` (device float[2]: recon and KL sums of the other ranks) serve batch-sharded runs."""
    return _Assemble.apply(err, kl, m_d, s, float(eta), float(gamma), float(free_bits), float(inv_batch), ext_sums)
