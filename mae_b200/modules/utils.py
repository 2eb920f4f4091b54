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
