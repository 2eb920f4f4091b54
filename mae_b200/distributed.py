"""Batch-sharded (one process per GPU) loss head and IW-NLL evaluation.

Partitioning: every rank owns ``B_loc`` samples.  Per-sample work (reparameterisation, KL,
likelihood, IW estimator) needs no communication.  The MPD regulariser is a function of the GLOBAL batch:

  * forward: ONE collective, an all-gather of the packed rows [B_loc, 2D] = (mu | logvar).  Every rank then runs the
    O(B*D) pre-pass on the gathered panel (identical m_d everywhere, plus the analytic KL of every row, so the
    free-bits clamp sees the global KL mean without another collective).  For global batches <= 2048 every rank also
    evaluates the full B x B tile pass redundantly (a few microseconds; cheaper than an all-reduce's latency); above
    that each rank evaluates its own row block and one fp64 scalar is all-reduced.
  * backward: each rank computes the COMPLETE gradient of its own rows (their role as the i and as the j argument
    of KL_ij), so there is no reduce-scatter of column gradients.

Semantics of the returned loss: gradients w.r.t. the rank's local tensors are the gradients of the GLOBAL objective
(mean over the global batch), so parameter gradients of the conv stacks must be SUM-reduced across ranks (plain
``all_reduce(SUM)``, or DDP with the loss multiplied by the world size).  The loss VALUE a rank returns contains its
share of the reconstruction term plus the global KL / MPD terms; ``global_loss`` all-reduces the share for logging.

The collective plumbing is independent of the compute engine: ``ShardedHead(engine=...)`` takes any object with the
``mae_b200.ops`` functions used below, which is how the world_size-2 gloo tests drive this file on CPU with the
oracle standing in for the CUDA kernels.
"""
from __future__ import annotations

import os

from typing import Optional

import torch
import torch.distributed as dist

SMALL_GLOBAL_BATCH = 2048   # <= : every rank evaluates the whole MPD forward redundantly (no all-reduce)
# Rows per rank up to this many floats (1 MB) take the one-kernel NVLink push gather (pure latency regime); larger panels
# (C5: up to 16 MB per rank) are bandwidth-bound and go through the library all-gather.
PEER_SLOT_FLOATS = 1 << 18
ROWS_SLOT_FLOATS = 64 * 128      # one-launch sharded regulariser: b_loc <= 64 rows of 2 D <= 128 floats
# Largest world size that takes the one-launch exchange + regulariser kernel (mae_kl_mpd_rows_fused).  Measured on one box
# (C3 step, batch 64 per GPU): world 2 43.6-45.7 us vs 46-48 us for gather -> cluster kernel, world 4 54.5 vs 56 us, but
# world 8 68.5 us vs 60.5 us for the gather -> prepare -> tiles -> backward chain: its 16 CTAs then serve 512 global rows.
ROWS_CLUSTER_MAX_WORLD = int(os.environ.get("MAE_ROWS_CLUSTER_MAX_WORLD", "4"))


def _packed_parent(mu, logvar, b_loc, d):
    """mu / logvar are usually the two chunk(2,1) halves of ONE contiguous [B, 2D] encoder output
    (encoder_cores/resnet.py:43): return that parent as a tensor without copying, else None."""
    try:
        m2, l2 = mu.view(b_loc, d), logvar.view(b_loc, d)
    except RuntimeError:
        return None
    if b_loc > 1 and (m2.stride() != (2 * d, 1) or l2.stride() != (2 * d, 1)):
        return None
    if l2.data_ptr() != m2.data_ptr() + 4 * d or m2.dtype != torch.float32:
        return None
    if b_loc == 1 and not (m2.is_contiguous() and l2.is_contiguous()):
        return None
    return torch.as_strided(m2.detach(), (b_loc, 2 * d), (2 * d, 1))


def _all_gather_rows(t: torch.Tensor, group) -> torch.Tensor:
    """[n, c] -> [world*n, c], rank-major (equal n on every rank)."""
    world = dist.get_world_size(group)
    out = t.new_empty(world * t.size(0), t.size(1))
    dist.all_gather_into_tensor(out, t.contiguous(), group=group)
    return out


class _ShardedKLMPD(torch.autograd.Function):
    """(mu_loc, logvar_loc) -> (kl_loc [B_loc], m_d [D], S [], kl_all [B_global]).

    With the CUDA engine the whole chain (pack, all-gather, pre-pass, tiles) runs on the engine's side stream, so the
    collective's latency and the MPD kernels overlap with the likelihood kernel on the main stream; the results are
    joined by ``assemble_loss``.  The backward runs on the side stream too (it needs no collective)."""

    @staticmethod
    def _compute(mu, logvar, head, gathered=None):
        group, eng = head.group, head.engine
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        b_loc = mu.size(0)
        d = mu[0].numel()
        packed = None
        if gathered is None:
            packed = _packed_parent(mu, logvar, b_loc, d)
            if packed is None:
                packed = torch.cat([mu.reshape(b_loc, d), logvar.reshape(b_loc, d)], dim=1)
        if head.check_equal_batches:
            lo = torch.tensor([b_loc], device=mu.device, dtype=torch.int64)
            hi = lo.clone()
            dist.all_reduce(lo, op=dist.ReduceOp.MIN, group=group)
            dist.all_reduce(hi, op=dist.ReduceOp.MAX, group=group)
            if int(lo) != int(hi):
                raise RuntimeError("ShardedHead needs the same local batch on every rank (got %d..%d): pad or drop the "
                                   "ragged tail batch" % (int(lo), int(hi)))
        if gathered is None:
            gathered = head._gather(packed)                     # the one exchange of the forward
        b = world * b_loc
        mu_all, lv_all = gathered[:, :d], gathered[:, d:]       # strided views, no copy
        row0 = rank * b_loc
        full = b <= SMALL_GLOBAL_BATCH
        m_d, s_or_sumsq, kl_all, state = eng.mpd_rows_forward(mu_all, lv_all, row0, b_loc, full)
        if full:
            s = s_or_sumsq
        else:
            dist.all_reduce(s_or_sumsq, op=dist.ReduceOp.SUM, group=group)
            s = eng.mpd_finalize(s_or_sumsq, b)
            state.s = s
        kl_loc = kl_all[row0:row0 + b_loc]
        return kl_loc, m_d, s, kl_all, state, gathered, row0, b_loc

    @staticmethod
    def forward(ctx, mu, logvar, head):
        ctx.set_materialize_grads(False)
        eng = head.engine
        side = mu.is_cuda and hasattr(eng, "_side_stream") and head.overlap
        if side:
            dev = mu.device
            cs, ss = torch.cuda.current_stream(dev), eng._side_stream(dev)
            ss.wait_stream(cs)
            with torch.cuda.stream(ss):
                kl_loc, m_d, s, kl_all, state, gathered, row0, b_loc = _ShardedKLMPD._compute(mu, logvar, head)
                ev = torch.cuda.Event()
                ev.record(ss)
            eng._pending_join[dev.index if dev.index is not None else torch.cuda.current_device()] = (
                ev, (kl_loc, m_d, s, kl_all, gathered, mu, logvar))
        else:
            kl_loc, m_d, s, kl_all, state, gathered, row0, b_loc = _ShardedKLMPD._compute(mu, logvar, head)
        ctx.state, ctx.eng, ctx.rows, ctx.shapes, ctx.side = state, eng, (row0, b_loc), (mu.shape, logvar.shape), side
        ctx.keep = gathered
        ctx.mark_non_differentiable(kl_all)
        return kl_loc, m_d.view(mu.shape[1:]), s, kl_all

    @staticmethod
    def backward(ctx, g_kl, g_m, g_s, _g_others):
        row0, b_loc = ctx.rows
        eng = ctx.eng
        if ctx.side:
            dev = ctx.keep.device
            cs, ss = torch.cuda.current_stream(dev), eng._side_stream(dev)
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            ready = eng._grads_ready.pop(idx, None)
            ptrs = {t.data_ptr() for t in (g_m, g_s, g_kl) if t is not None}
            if ready is not None and ptrs <= ready[1]:
                ss.wait_event(ready[0])
            else:
                ss.wait_stream(cs)
            with torch.cuda.stream(ss):
                dmu, dlv = eng.mpd_rows_backward(ctx.state, row0, b_loc, g_m, g_s, g_kl)
                ev = torch.cuda.Event()
                ev.record(ss)
            cs.wait_event(ev)
            dmu.record_stream(cs)      # allocated on the side stream, consumed (and freed) on the main stream
            dlv.record_stream(cs)
        else:
            dmu, dlv = eng.mpd_rows_backward(ctx.state, row0, b_loc, g_m, g_s, g_kl)
        ctx.keep = None
        return dmu.view(ctx.shapes[0]), dlv.view(ctx.shapes[1]), None


class _ShardedEncodeHead(torch.autograd.Function):
    """(mu_loc, logvar_loc, eps_loc) -> (z_loc, reg, kl_all [B_global], m_d, S): the sharded counterpart of
    ``ops.encode_head``.  Side stream: pack, all-gather, pre-pass, tiles, (all-reduce of the fp64 scalar for large global
    batches) AND the regulariser's backward for the own rows, all enqueued back to back in the forward pass.  Main stream:
    the reparameterisation.  The backward pass is ONE launch (reparam backward + d reg * stored gradient): nothing of
    the regulariser, and no collective, is left on it."""

    @staticmethod
    def forward(ctx, mu, logvar, eps, head, eta, gamma, free_bits, with_kl=True, holder=None):
        ctx.set_materialize_grads(False)
        eng = head.engine
        want_grad = ctx.needs_input_grad[0] or ctx.needs_input_grad[1]
        world = head.world
        z_side = holder is not None and holder.get("z_on_side", False)

        def chain():
            b_loc0, d0 = mu.size(0), mu[0].numel()
            if head._rows_cluster_ok(b_loc0, d0, mu):
                # training batches: exchange + regulariser of the global batch (forward, and backward of the own rows) in
                # ONE cluster launch; the landed rows are consumed in place, nothing of the B x B work is replicated
                packed = _packed_parent(mu, logvar, b_loc0, d0)
                if packed is None:
                    packed = torch.cat([mu.reshape(b_loc0, d0), logvar.reshape(b_loc0, d0)], dim=1)
                packed = packed.detach()
                m_d, s, kl_all, grads, reg = eng.reg_rows_cluster(packed, head.peer_rows, eta, gamma, free_bits,
                                                                  None if with_kl else 0.0, want_grad, head._rows_scratch)
                return m_d, s, kl_all, packed, grads, reg
            if hasattr(eng, "reg_rows_fused"):
                # small global batch: gather, then ONE cluster launch for the whole regulariser of the gathered rows
                b_loc, d = mu.size(0), mu[0].numel()
                packed = _packed_parent(mu, logvar, b_loc, d)
                if packed is None:
                    packed = torch.cat([mu.reshape(b_loc, d), logvar.reshape(b_loc, d)], dim=1)
                gathered = head._gather(packed)
                row0 = head.rank * b_loc
                out = eng.reg_rows_fused(gathered[:, :d], gathered[:, d:], row0, b_loc, eta, gamma, free_bits, want_grad,
                                         kl_weight=None if with_kl else 0.0)
                if out is not None:
                    m_d, s, kl_all, grads, reg = out
                    return m_d, s, kl_all, gathered, grads, reg
                state_in = gathered
            else:
                state_in = None
            kl_loc, m_d, s, kl_all, state, gathered, row0, b_loc = _ShardedKLMPD._compute(mu, logvar, head, state_in)
            grads, reg = eng.mpd_rows_loss_backward(state, row0, b_loc, eta, gamma, free_bits,
                                                    1.0 / (world * b_loc) if with_kl else 0.0, m_d, kl_all, want_grad)
            return m_d, s, kl_all, gathered, grads, reg

        side = mu.is_cuda and hasattr(eng, "_side_stream") and head.overlap
        ev = None
        if side:
            dev = mu.device
            idx = dev.index if dev.index is not None else torch.cuda.current_device()
            cs, ss = torch.cuda.current_stream(dev), eng._side_stream(dev)
            ss.wait_stream(cs)
            prev = eng._pending_join.pop(idx, None)
            z = None
            with torch.cuda.stream(ss):
                m_d, s, kl_all, gathered, grads, reg = chain()
                if z_side:      # see ops.encode_head(z_on_side=True): joined by loss_head behind the likelihood kernel
                    z, saved = eng.reparam_fwd(mu, logvar, eps)
                    z.record_stream(cs)
                ev = torch.cuda.Event()
                ev.record(ss)
                if z_side:
                    holder["z_event"] = ev
            eng._pending_join[idx] = (ev, (prev, m_d, s, kl_all, gathered, grads, reg, mu, logvar, z))
            if z is None:
                z, saved = eng.reparam_fwd(mu, logvar, eps)
        else:
            m_d, s, kl_all, gathered, grads, reg = chain()
            z, saved = eng.reparam_fwd(mu, logvar, eps)
        ctx.saved, ctx.grads, ctx.event, ctx.eng = saved, grads, ev, eng
        ctx.shapes = (mu.shape, logvar.shape)
        ctx.mark_non_differentiable(kl_all, m_d, s)
        return z, reg, kl_all, m_d.view(mu.shape[1:]), s

    @staticmethod
    def backward(ctx, gz, g_reg, *_unused):
        if gz is None and g_reg is None:
            return (None,) * 9
        if g_reg is not None and ctx.grads is None:
            raise RuntimeError("sharded encode head: reg has a gradient but mu / logvar did not require one in the forward")
        if ctx.event is not None:
            torch.cuda.current_stream(ctx.saved[0].device).wait_event(ctx.event)
        dmu, dlv = ctx.eng.reparam_reg_bwd(ctx.saved, gz, ctx.grads, g_reg)
        return dmu.view(ctx.shapes[0]), dlv.view(ctx.shapes[1]), None, None, None, None, None, None, None


class ShardedHead:
    """Batch-sharded MAE loss head / evaluator.  ``engine`` defaults to ``mae_b200.ops`` (the CUDA kernels)."""

    def __init__(self, group=None, engine=None, check_equal_batches: bool = False, overlap: bool = True,
                 use_p2p: bool = True, defer_join: bool = False, check_every: int = 256):
        if engine is None:
            from . import ops as engine
        self.engine = engine
        self.group = group
        self.check_equal_batches = check_equal_batches   # costs two tiny all-reduces and a host sync per step
        self.overlap = overlap                           # all-gather + MPD on the side stream (CUDA engine only)
        # defer_join=True keeps the objective's scalar on the side stream: loss / stats are then valid only after
        # backward() or ops.join_side() (what a CUDA-graph-captured training step wants; bench.py sets it)
        self.defer_join = defer_join
        self.two_node = True                             # regulariser backward in the forward pass (engine permitting)
        self.peer = None                                 # NVLink peer-memory gather (created collectively, CUDA only)
        self.peer_rows = None                            # buffers of the one-launch exchange + regulariser kernel
        self._rows_failed = False
        self._rows_scratch = None
        self.rows_cluster = True                         # use mae_kl_mpd_rows_fused where the shape fits
        self.use_p2p = use_p2p
        self.check_every = check_every                   # poll the gather's error flag every N eager steps (0: never)
        self.world = dist.get_world_size(group)
        self.rank = dist.get_rank(group)
        self._pdl_before = None
        if self.world > 1 and hasattr(self.engine, "set_pdl"):
            # measured at 8 GPUs (same box, A/B): 60.5 us/step without programmatic dependent launch vs 68.3 us with it.
            # The switch is process-wide in the library; close() restores the previous setting.
            self._pdl_before = self.engine.set_pdl(False)

    def _rows_cluster_ok(self, b_loc: int, d: int, like: torch.Tensor) -> bool:
        """Whether this step can take the one-launch exchange + regulariser kernel (decided identically on every rank: it
        depends on shapes, on the collectively created buffers and on the ``rows_cluster`` switch only)."""
        eng = self.engine
        if not (self.rows_cluster and self.use_p2p and like.is_cuda and self.world > 1 and hasattr(eng, "reg_rows_cluster")):
            return False
        if self.world > ROWS_CLUSTER_MAX_WORLD:
            return False
        if (b_loc * 2 * d) % 4 != 0 or b_loc * 2 * d > ROWS_SLOT_FLOATS or not eng.rows_cluster_supported(self.world, b_loc, d, like.device):
            return False
        if self.peer_rows is None and not self._rows_failed:
            from .p2p import PeerGather
            self.peer_rows = PeerGather.create(self.group, ROWS_SLOT_FLOATS, like.device, layout="rows")
            if self.peer_rows is None:
                self._rows_failed = True
            else:
                import ctypes
                from . import _lib
                self._rows_scratch = torch.empty(int(_lib.load().mae_rows_scratch_bytes()), dtype=torch.uint8, device=like.device)
        return self.peer_rows is not None

    def close(self) -> None:
        """Release the peer buffers and restore the process-wide PDL setting this head changed."""
        if self.peer is not None:
            self.peer.close()
            self.peer = None
        if self.peer_rows is not None:
            self.peer_rows.close()
            self.peer_rows = None
        if self._pdl_before is not None:
            self.engine.set_pdl(self._pdl_before)
            self._pdl_before = None

    def check(self) -> None:
        """Raise if the peer-memory gather ever timed out (its rows were NaN-filled).  Synchronises with the device."""
        if self.peer is not None:
            self.peer.check()
        if self.peer_rows is not None:
            self.peer_rows.check()

    def _gather(self, packed: torch.Tensor) -> torch.Tensor:
        """All-gather of the packed rows: one push-kernel over NVLink peer memory when available (tens of KB, pure
        latency), else NCCL / gloo ``all_gather_into_tensor``.  Whether the peer path is used is decided collectively."""
        if self.use_p2p and packed.is_cuda and self.world > 1 and hasattr(self.engine, "_side_stream"):
            if self.peer is None:
                from .p2p import PeerGather
                self.peer = PeerGather.create(self.group, PEER_SLOT_FLOATS, packed.device)
                if self.peer is None:     # same outcome on every rank
                    import warnings
                    warnings.warn("peer-memory all-gather unavailable (%s); using the library all-gather"
                                  % (PeerGather.last_failure,))
                    self.use_p2p = False
            if self.peer is not None and self.peer.usable(packed):
                if (self.check_every and self.peer.calls and self.peer.calls % self.check_every == 0
                        and not torch.cuda.is_current_stream_capturing()):
                    self.peer.check()
                return self.peer.gather(packed)
        return _all_gather_rows(packed, self.group)

    # ------------------------------------------------------------------------------------------ training
    def kl_mpd(self, mu, logvar):
        """Analytic KL of the local rows, global MPD (m_d, S) and the KL of every gathered row [B_global]."""
        return _ShardedKLMPD.apply(mu, logvar, self)

    def _tail(self, dec_out, x, kl, m_d, s, kl_all, kind, ns, nmix, eta, gamma, free_bits):
        eng = self.engine
        b_glob = self.world * x.size(0)
        if hasattr(eng, "loss_tail") and dec_out.is_cuda:
            # fused likelihood + objective node; with overlap the objective's tiny kernels stay on the side stream
            # behind the all-gather / MPD chain, so the main stream runs likelihood fwd -> likelihood bwd back to back
            return eng.loss_tail(dec_out, x, kl, m_d, s, kind, ns, nmix, eta, gamma, free_bits, 1.0 / b_glob, None,
                                 defer_join=self.overlap and self.defer_join, kl_all=kl_all)
        kl_others = kl_all.sum() - kl.detach().sum()
        if kind == "dmol":
            err = eng.dmol_recon_error(dec_out, x, ns, nmix)
        else:
            err = eng.bernoulli_recon_error(dec_out, x)
        ext = torch.stack([kl_others.new_zeros(()), kl_others]).to(kl.dtype)
        return eng.assemble_loss(err, kl, m_d, s, eta, gamma, free_bits, 1.0 / b_glob, ext)

    def encode_head(self, mu, logvar, eps, eta=0.0, gamma=0.0, free_bits=0.0, with_kl: bool = True, z_on_side: bool = False):
        """Sharded counterpart of ``ops.encode_head``: -> (z_loc, EncodedBatch).  ``enc.kl`` is the all-gathered analytic KL
        vector [B_global] (None when ``with_kl`` is False: encoders with flows pass their own KL to ``loss_head``)."""
        eng = self.engine
        holder = {"z_on_side": bool(z_on_side) and self.overlap}
        z, reg, kl_all, m_d, s = _ShardedEncodeHead.apply(mu, logvar, eps, self, float(eta), float(gamma), float(free_bits),
                                                          bool(with_kl), holder)
        enc = eng.EncodedBatch(reg, kl_all if with_kl else None, m_d, s, (float(eta), float(gamma), float(free_bits)))
        enc.z_event = holder.get("z_event")
        return z, enc

    def loss_head(self, dec_out, x, enc, kind, ns, nmix=10, eta=0.0, gamma=0.0, free_bits=0.0, kl=None):
        """Sharded counterpart of ``ops.loss_head``: -> (loss, stats, recon [B_loc]).  ``kl`` [B_loc] (differentiable): the
        caller's Monte-Carlo KL when ``enc`` was made with ``with_kl=False``; its global mean enters the free-bits clamp
        through one scalar all-reduce."""
        b_loc = x.size(0)
        ext = None
        if kl is not None:
            own = kl.detach().sum()
            tot = own.clone()
            dist.all_reduce(tot, op=dist.ReduceOp.SUM, group=self.group)
            ext = torch.stack([torch.zeros_like(own), tot - own]).float()
        return self.engine.loss_head(dec_out, x, enc, kind, ns, nmix, eta, gamma, free_bits,
                                     defer_join=self.overlap and self.defer_join, inv_batch=1.0 / (self.world * b_loc),
                                     kl=kl, ext_sums=ext)

    def _two_node(self, mu, logvar, eps, dec_out, x, kind, nmix, eta, gamma, free_bits):
        """encode head (regulariser forward + backward behind the all-gather, side stream) + loss head."""
        z, enc = self.encode_head(mu, logvar, eps, eta, gamma, free_bits, z_on_side=True)   # dec_out is an input here
        b_loc = x.size(0)
        loss, stats, recon = self.loss_head(dec_out, x, enc, kind, eps.size(1), nmix, eta, gamma, free_bits)
        row0 = self.rank * b_loc
        return loss, recon, enc.kl[row0:row0 + b_loc], stats, z

    def _has_two_node(self) -> bool:
        return self.two_node and all(hasattr(self.engine, n) for n in ("mpd_rows_loss_backward", "reparam_fwd",
                                                                       "reparam_reg_bwd", "loss_head", "EncodedBatch"))

    def loss_head_color(self, mu, logvar, eps, raw, x, nmix=10, eta=0.0, gamma=0.0, free_bits=0.0):
        """C3 on a batch shard.  Returns (loss, recon [B_loc], KL [B_loc], stats, z) like ``heads.loss_head_color``."""
        if self._has_two_node():
            return self._two_node(mu, logvar, eps, raw, x, "dmol", nmix, eta, gamma, free_bits)
        z, _ = self.engine.reparam_kl(mu, logvar, eps, want_kl=False)
        kl, m_d, s, kl_all = self.kl_mpd(mu, logvar)
        loss, stats, recon = self._tail(raw, x, kl, m_d, s, kl_all, "dmol", eps.size(1), nmix, eta, gamma, free_bits)
        return loss, recon, kl, stats, z

    def loss_head_binary(self, mu, logvar, eps, probs, x, eta=0.0, gamma=0.0, free_bits=0.0):
        """C1/C2 on a batch shard."""
        if self._has_two_node():
            return self._two_node(mu, logvar, eps, probs, x, "bernoulli", 0, eta, gamma, free_bits)
        z, _ = self.engine.reparam_kl(mu, logvar, eps, want_kl=False)
        kl, m_d, s, kl_all = self.kl_mpd(mu, logvar)
        loss, stats, recon = self._tail(probs, x, kl, m_d, s, kl_all, "bernoulli", eps.size(1), 0, eta, gamma, free_bits)
        return loss, recon, kl, stats, z

    def global_loss(self, loss: torch.Tensor, stats: torch.Tensor) -> torch.Tensor:
        """The global objective from a rank's loss: all-reduce of the reconstruction share (stats[4] = mean recon share)."""
        if hasattr(self.engine, "join_side") and loss.is_cuda:
            self.engine.join_side(loss.device)      # deferred objective kernels may still sit on the side stream
        share = stats[4].detach().clone()
        total = share.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return loss.detach() - share + total

    def global_loss_value(self, loss: torch.Tensor) -> torch.Tensor:
        """Global objective from the per-rank losses returned by a sharded ``MAE.loss``: every rank's loss holds its share
        of the reconstruction term plus the (identical) global KL / MPD terms, so
        global = sum_r loss_r - (world - 1) * common, with common = loss_r - share_r; computed as
        all_reduce(share) + (loss - share) using the share recorded by the last sharded loss call."""
        share = self._last_share
        if share is None:
            raise RuntimeError("global_loss_value: no sharded MAE.loss call recorded on this head")
        if hasattr(self.engine, "join_side") and loss.is_cuda:
            self.engine.join_side(loss.device)
        share = share.detach().clone()
        total = share.clone()
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=self.group)
        return loss.detach() - share + total

    _last_share = None

    # ----------------------------------------------------------------------------------------- evaluation
    def shard(self, n_items: int):
        """Contiguous block of ``n_items`` test images owned by this rank: (start, stop)."""
        per, rem = divmod(n_items, self.world)
        start = self.rank * per + min(self.rank, rem)
        return start, start + per + (1 if self.rank < rem else 0)

    def reduce_eval(self, sums: torch.Tensor) -> torch.Tensor:
        """Final reduction of an evaluation: element-wise SUM of per-rank accumulators
        (e.g. [sum nll_elbo, sum recon, sum kl, sum nll_iw, n_images]); the only communication of IW-NLL evaluation."""
        out = sums.clone()
        dist.all_reduce(out, op=dist.ReduceOp.SUM, group=self.group)
        return out


# ------------------------------------------------------------------------------------------------ whole models
def shard_model(model, head: Optional[ShardedHead]):
    """Make ``model.loss`` (mae_b200.modules.MAE) treat its batch as this rank's shard of the global batch: the MPD
    regulariser and the KL mean are taken over the all-gathered posteriors, and the returned gradients are those of the
    GLOBAL objective, so ``all_reduce_gradients`` (SUM) afterwards reproduces the single-process parameter gradients.
    Replaces the reference's ``nn.DataParallel`` replicas of the conv cores (gaussian_encoder.py:41-42).  ``None`` undoes it."""
    model._sharded_head = head
    return model


def all_reduce_gradients(model, group=None, bucket_bytes: int = 64 << 20) -> None:
    """SUM-all-reduce of every parameter gradient in flat buckets (north_star: 'gradients are all-reduced').  SUM, not
    MEAN: the sharded loss already divides by the global batch.  Parameters without a gradient on some rank are treated as
    zero there (every rank must own the same parameter set)."""
    params = [p for p in model.parameters() if p.requires_grad]
    bucket, size = [], 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in bucket])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        off = 0
        for p in bucket:
            n = p.numel()
            g = flat[off:off + n].view_as(p)
            if p.grad is None:
                p.grad = g.clone()
            else:
                p.grad.copy_(g)
            off += n
        bucket, size = [], 0

    for p in params:
        if bucket and (bucket[0].dtype != p.dtype or size + p.numel() * p.element_size() > bucket_bytes):
            flush()
        bucket.append(p)
        size += p.numel() * p.element_size()
    flush()
