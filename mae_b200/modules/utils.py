"""Host-side helpers of mae.modules.utils that are not on the GPU hot path.

``logsumexp`` and ``exponentialMovingAverage`` keep the reference's names and semantics
(mae/modules/utils.py:23-38, :163-172).  The discretized-logistic likelihood itself lives in the CUDA
library (``mae_b200.ops.dmol_recon_error``) and is reached through ``ColorImageDecoder.reconstruct_error``.
"""
import torch


def logsumexp(x, dim=None):
    """log(sum(exp(x))) over ``dim`` (all axes when None).  Reference: mae/modules/utils.py:23-38."""
    if dim is None:
        return torch.logsumexp(x.reshape(-1), dim=0)
    return torch.logsumexp(x, dim=dim)


def exponentialMovingAverage(original, shadow, decay_rate, init=False):
    """Polyak average of parameters into ``shadow``.  Reference: mae/modules/utils.py:163-172."""
    src = dict(original.named_parameters())
    with torch.no_grad():
        for name, sp in shadow.named_parameters():
            p = src[name]
            if init:
                sp.copy_(p)
            else:
                sp.add_((1.0 - decay_rate) * (p - sp))


def norm(p: torch.Tensor, dim):
    """L2 norm over all dimensions except ``dim`` (kept for broadcasting); ``dim=None``: the full norm.
    Reference: mae/modules/utils.py:9-20."""
    if dim is None:
        return p.norm()
    if dim < 0:
        dim += p.dim()
    dims = [i for i in range(p.dim()) if i != dim]
    return torch.linalg.vector_norm(p, dim=dims, keepdim=True) if dims else p.abs()


def logavgexp(x, dim=None):
    """log(mean(exp(x))) over ``dim`` (all axes when None).  Reference: mae/modules/utils.py:41-58."""
    import math
    n = x.numel() if dim is None else x.size(dim)
    return logsumexp(x, dim) - math.log(n)


def sample_from_discretized_mix_logistic(means, logscales, coeffs, logit_probs, random_sample):
    """Most probable mixture component per pixel, optional logistic noise, colour coupling, clamp to [-1, 1]:
    [B,nmix,3,H,W] x3, [B,nmix,H,W] -> [B,3,H,W].  Reference: mae/modules/utils.py:124-160 (sampling is not on the hot path;
    the likelihood itself is ``ColorImageDecoder.reconstruct_error`` on the CUDA kernels)."""
    from .decoder import ColorImageDecoder
    return ColorImageDecoder.sample_pixels(means, logscales, logit_probs, coeffs, random_sample)
