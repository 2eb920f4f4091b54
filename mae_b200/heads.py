"""The loss / evaluation heads as single calls on synthetic encoder/decoder outputs.

These compose the operators exactly the way ``MAE.loss`` / ``MAE.nll`` do (mae/modules/mae.py:100-140,
179-216) but take the conv stacks' outputs as arguments, which is what benchmarks and the smoke test need
(BASELINE.json: the conv stacks stay in PyTorch and are not part of the measured path)."""
from __future__ import annotations

from typing import Optional

import torch

from . import ops


ONE_NODE = True   # False forces the three-node composition (reparam, kl_mpd, loss_tail) that flows / sharded runs use


def _one_node(mu: torch.Tensor) -> bool:
    """Two autograd nodes in total: ops.encode_head (z + the whole regulariser forward AND backward on the side stream: one
    cluster launch for training-sized batches, the general kernel chain otherwise) and ops.loss_head."""
    return ONE_NODE and mu.size(0) >= 2


def loss_head_binary(mu, logvar, eps, probs, x, eta=0.0, gamma=0.0, free_bits=0.0, defer_join=False):
    """C1/C2: reparam + analytic KL, MPD, Bernoulli likelihood, assembly.  Returns (loss, recon, KL, stats, z).
    ``defer_join``: see ops.loss_tail (outputs valid after backward() / ops.join_side())."""
    if _one_node(mu):
        # regulariser AND z on the side stream: the likelihood kernel below reads the given decoder output, not z
        z, enc = ops.encode_head(mu, logvar, eps, eta, gamma, free_bits, z_on_side=True)
        loss, stats, recon = ops.loss_head(probs, x, enc, "bernoulli", eps.size(1), 0, eta, gamma, free_bits,
                                           defer_join=defer_join)
        return loss, recon, enc.kl, stats, z
    z, _ = ops.reparam_kl(mu, logvar, eps, want_kl=False)
    kl, m_d, s = ops.kl_mpd(mu, logvar, side_stream=True)   # overlaps with the likelihood; joined inside loss_tail
    loss, stats, recon = ops.loss_tail(probs, x, kl, m_d, s, "bernoulli", eps.size(1), 0, eta, gamma, free_bits,
                                       defer_join=defer_join)
    return loss, recon, kl, stats, z


def loss_head_color(mu, logvar, eps, raw, x, nmix=10, eta=0.0, gamma=0.0, free_bits=0.0, defer_join=False):
    """C3: reparam + analytic KL, MPD, discretized-logistic-mixture likelihood, assembly.
    ``defer_join``: see ops.loss_tail (outputs valid after backward() / ops.join_side())."""
    if _one_node(mu):
        # regulariser AND z on the side stream: the likelihood kernel below reads the given decoder output, not z
        z, enc = ops.encode_head(mu, logvar, eps, eta, gamma, free_bits, z_on_side=True)
        loss, stats, recon = ops.loss_head(raw, x, enc, "dmol", eps.size(1), nmix, eta, gamma, free_bits,
                                           defer_join=defer_join)
        return loss, recon, enc.kl, stats, z
    z, _ = ops.reparam_kl(mu, logvar, eps, want_kl=False)
    kl, m_d, s = ops.kl_mpd(mu, logvar, side_stream=True)   # overlaps with the likelihood; joined inside loss_tail
    loss, stats, recon = ops.loss_tail(raw, x, kl, m_d, s, "dmol", eps.size(1), nmix, eta, gamma, free_bits,
                                       defer_join=defer_join)
    return loss, recon, kl, stats, z


def iw_eval_color(raw, x, z, mu, logvar, nmix=10, z_prior_normal: Optional[torch.Tensor] = None,
                  logdet_prior: Optional[torch.Tensor] = None, logdet_post: Optional[torch.Tensor] = None):
    """C4: importance-weighted NLL of a batch of images from K posterior samples each.
    raw [B*K,10*nmix,H,W], x [B,3,H,W], z [B,K,D] -> ((nll_elbo, recon, kl), nll_iw), each [B]."""
    with torch.no_grad():
        k = z.size(1)
        err = ops.dmol_recon_error(raw, x, k, nmix)          # log p(x|z) = -err: the IW kernel applies the sign
        zp = z if z_prior_normal is None else z_prior_normal
        return ops.iw_nll(err, zp, logdet_prior, z, logdet_post, mu, logvar, gen_scale=-1.0)
