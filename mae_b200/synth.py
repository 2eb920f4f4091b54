"""Deterministic synthetic encoder/decoder outputs of the shapes SURVEY.md §8(d) names.

All tensors are drawn on the CPU with a seeded ``torch.Generator`` (default seed 524287, the
reference's default ``--seed``, experiments/binary_image.py:27) so that the CUDA path, the oracle and
the committed golden fixtures consume bit-identical inputs; callers move them to the device.
There is no dataset or checkpoint access in this environment, so every benchmark says
``"data": "synthetic"``.
"""
from __future__ import annotations

from typing import Dict, Tuple

import torch

SEED = 524287


def _gen(seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(int(seed))
    return g


def posterior_params(batch: int, z_shape: Tuple[int, ...], seed: int = SEED, clamp: bool = False
                     ) -> Tuple[torch.Tensor, torch.Tensor]:
    """mu ~ N(0,1), logvar ~ N(0,1) (hard-clamped to [-7,7] for the colour cores,
    encoder_cores/resnet.py:96,103)."""
    g = _gen(seed)
    mu = torch.randn(batch, *z_shape, generator=g)
    logvar = torch.randn(batch, *z_shape, generator=g)
    if clamp:
        logvar = logvar.clamp(-7.0, 7.0)
    return mu, logvar


def noise(batch: int, nsamples: int, z_shape: Tuple[int, ...], seed: int = SEED) -> torch.Tensor:
    return torch.randn(batch, nsamples, *z_shape, generator=_gen(seed + 1))


def binary_head(batch: int = 100, nsamples: int = 1, pixels: int = 784, seed: int = SEED
                ) -> Dict[str, torch.Tensor]:
    """C1/C2: post-sigmoid pixel probabilities [B,ns,P] and a binary image [B,1,28,28]-like x."""
    g = _gen(seed + 2)
    probs = torch.sigmoid(2.0 * torch.randn(batch, nsamples, pixels, generator=g))
    x = (torch.rand(batch, pixels, generator=g) < 0.15).float()
    return {"probs": probs, "x": x}


def color_head(nimages: int = 64, nsamples: int = 1, nmix: int = 10, height: int = 32, width: int = 32,
               seed: int = SEED) -> Dict[str, torch.Tensor]:
    """C3/C4: raw decoder output [B*ns,10*nmix,H,W] and 8-bit images mapped to [-1,1].

    The 3*nmix log-scale channels are shifted by -3 on a random 10 % of pixels so that the
    ``clamp(min=-7)`` gate and very sharp components occur; pixel values 0 and 255 occur, so both
    tail branches of the discretized logistic run.
    """
    g = _gen(seed + 3)
    n = nimages * nsamples
    raw = torch.randn(n, 10 * nmix, height, width, generator=g)
    sharp = (torch.rand(n, 1, height, width, generator=g) < 0.10).float()
    raw[:, 4 * nmix:7 * nmix] -= 3.0 * sharp
    x = torch.randint(0, 256, (nimages, 3, height, width), generator=g).float() / 127.5 - 1.0
    return {"raw": raw, "x": x}


def color_head_hard(nimages: int = 3, nsamples: int = 2, nmix: int = 10, height: int = 8, width: int = 8,
                    seed: int = SEED) -> Dict[str, torch.Tensor]:
    """A small adversarial DMOL case: wide log-scales (delta below 1e-5 -> the ``approx`` branch),
    log-scales below -7 (clamp gate), saturated coefficients and many edge pixels (0 / 255)."""
    g = _gen(seed + 4)
    n = nimages * nsamples
    raw = torch.randn(n, 10 * nmix, height, width, generator=g)
    sel = torch.rand(n, 3 * nmix, height, width, generator=g)
    ls = raw[:, 4 * nmix:7 * nmix]
    ls[sel < 0.25] += 6.0          # very wide components: cdf_delta < 1e-5
    ls[sel > 0.85] -= 9.0          # below the -7 clamp
    raw[:, 7 * nmix:] *= 3.0       # tanh saturation
    raw[:, nmix:4 * nmix] *= 2.0   # means far outside [-1,1]
    pix = torch.randint(0, 256, (nimages, 3, height, width), generator=g)
    edge = torch.rand(nimages, 3, height, width, generator=g)
    pix[edge < 0.2] = 0
    pix[edge > 0.8] = 255
    x = pix.float() / 127.5 - 1.0
    return {"raw": raw, "x": x}


def fill_state(module, seed: int = SEED, gain: float = 0.7):
    """Deterministic parameter values keyed by parameter NAME (crc32), for model-level parity tests without stored
    weights: the same call on the reference's model and on ours (identical state_dict keys) yields identical weights.
    weight_v ~ N(0, 0.05); weight_g ~ |gain (1 + 0.1 N)| for weight-normalised layers (row norm of the weight), 0.1 N +
    log(gain / 3) for the masked layers (log-scale gain, recognised by their ``mask`` buffer; kept small so that stacked
    autoregressive flows stay in range); bias ~ 0.1 N.  Mask buffers are left as constructed."""
    import math
    import zlib
    sd = module.state_dict()
    masked = {n[:-len("mask")] for n in sd if n.endswith("mask")}
    with torch.no_grad():
        for name, t in sd.items():
            if name.endswith("mask"):
                continue
            g = torch.Generator().manual_seed((seed ^ zlib.crc32(name.encode())) & 0x7FFFFFFF)
            r = torch.randn(tuple(t.shape), generator=g, dtype=torch.float32)
            prefix = name[:name.rfind(".") + 1]
            if name.endswith("weight_v"):
                v = 0.05 * r
            elif name.endswith("weight_g"):
                v = 0.1 * r + math.log(gain / 3.0) if prefix in masked else (gain * (1.0 + 0.1 * r)).abs()
            elif name.endswith("bias"):
                v = 0.1 * r
            else:
                v = 0.05 * r
            t.copy_(v.to(t.dtype))
    return module
