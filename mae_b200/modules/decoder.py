"""Decoder side of the drop-in: ``Decoder`` interface + registry (mae/modules/decoders/decoder.py:7-101) and the
two likelihood heads, ``BinaryImageDecoder`` (image_decoders/binary_image_decoder.py) and ``ColorImageDecoder``
(image_decoders/color_image_decoder.py).  Sub-classes supply the conv stack as ``forward``; the likelihood
itself runs on the sm_100a kernels."""
from __future__ import annotations

from typing import Dict, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from .. import ops


class Decoder(nn.Module):
    _registry: Dict[str, type] = dict()

    def z_shape(self) -> Tuple:
        raise NotImplementedError

    def output_size(self) -> Tuple:
        raise NotImplementedError

    def initialize(self, x, z, init_scale=1.0):
        raise NotImplementedError

    def decode(self, z, random_sample):
        raise NotImplementedError

    def reconstruct_error(self, x, z):
        raise NotImplementedError

    def log_probability(self, x, z):
        raise NotImplementedError

    @classmethod
    def register(cls, name: str):
        Decoder._registry[name] = cls

    @classmethod
    def by_name(cls, name: str):
        return Decoder._registry[name]

    @classmethod
    def from_params(cls, params: Dict) -> "Decoder":
        raise NotImplementedError


class BinaryImageDecoder(Decoder):
    """Pixel-wise independent Bernoulli.  ``forward(z)`` of a sub-class returns post-sigmoid probabilities."""

    def __init__(self, nz, ngpu=1):
        super().__init__()
        self.nz = nz
        self.ngpu = ngpu

    def z_shape(self) -> Tuple:
        return self.nz, 1, 1

    def decode(self, z, random_sample):
        """Reference: binary_image_decoder.py:52-68."""
        img_probs = self(z)
        img = torch.bernoulli(img_probs) if random_sample else torch.ge(img_probs, 0.5).float()
        return img, img_probs

    def bernoulli_error(self, recon_x, x):
        """probabilities [B,ns,P] (any trailing shape) and x [B,*] -> [B,ns] on the CUDA kernel."""
        b, ns = recon_x.size(0), recon_x.size(1)
        return ops.bernoulli_recon_error(recon_x.reshape(b, ns, -1), x.reshape(b, -1))

    likelihood_kind = "bernoulli"

    def likelihood_input(self, x, z):
        """The conv stack's output in the layout the likelihood kernel reads: probabilities [B,ns,P].
        Sub-classes whose network also conditions on x (PixelCNN, pixelcnn.py:119-134) override this method instead of
        ``reconstruct_error``, so that ``MAE.loss`` can keep likelihood + objective in one fused autograd node."""
        size = z.size()
        b, ns = size[:2]
        return self(z.view(b * ns, *size[2:])).view(b, ns, -1)

    def reconstruct_error(self, x, z):
        """x [B,*], z [B,ns,*z] -> [B,ns].  Reference: binary_image_decoder.py:71-93."""
        return self.bernoulli_error(self.likelihood_input(x, z), x)

    def log_probability(self, x, z):
        return self.reconstruct_error(x, z) * -1.


class ColorImageDecoder(Decoder):
    """Discretized mixture of logistics for 3-channel images.  ``forward(x, z)`` of a sub-class returns the raw
    conv output [N, 10*nmix, H, W]."""

    def __init__(self, nmix, ngpu=1):
        super().__init__()
        self.nc = 3
        self.nmix = nmix
        self.ngpu = ngpu

    def execute(self, z, x):
        """Post-processed mixture parameters (used by ``decode``; the likelihood kernel does this itself).
        Reference: color_image_decoder.py:38-52."""
        output = self(x, z)
        nmix, nc = self.nmix, self.nc
        logit_probs = output[:, :nmix]
        batch, _, H, W = logit_probs.size()
        mu = output[:, nmix:(nc + 1) * nmix].view(batch, nmix, nc, H, W)
        log_scale = output[:, (nc + 1) * nmix:(nc * 2 + 1) * nmix].view(batch, nmix, nc, H, W)
        coeffs = output[:, (nc * 2 + 1) * nmix:(nc * 2 + 4) * nmix].view(batch, nmix, nc, H, W)
        return mu, log_scale.clamp(min=-7.), F.log_softmax(logit_probs, dim=1), coeffs.tanh()

    @staticmethod
    def sample_pixels(mu, log_scale, logit_probs, coeffs, random_sample):
        """Most probable component per pixel, logistic noise if ``random_sample``, colour coupling, clamp to [-1, 1]
        -> [B, 3, H, W].  Reference: utils.py:127-160 (sampling; not on the hot path)."""
        index = logit_probs.argmax(dim=1, keepdim=True).unsqueeze(2).expand(-1, 1, mu.size(2), -1, -1)
        mu = mu.gather(1, index).squeeze(1)
        log_scale = log_scale.gather(1, index).squeeze(1)
        coeffs = coeffs.gather(1, index).squeeze(1)
        x = mu
        if random_sample:
            u = torch.empty_like(mu).uniform_(1e-5, 1 - 1e-5)
            x = x + log_scale.exp() * (torch.log(u) - torch.log(1.0 - u))
        x0 = x[:, 0].clamp(min=-1., max=1.)
        x1 = (x[:, 1] + coeffs[:, 0] * x0).clamp(min=-1., max=1.)
        x2 = (x[:, 2] + coeffs[:, 1] * x0 + coeffs[:, 2] * x1).clamp(min=-1., max=1.)
        return torch.stack([x0, x1, x2], dim=1)

    def decode(self, z, random_sample):
        """Reference: color_image_decoder.py:71-88 (decoders that do not condition on x)."""
        return self.sample_pixels(*self.execute(z, None), random_sample), None

    likelihood_kind = "dmol"

    def likelihood_input(self, x, z):
        """The conv stack's raw output [B*ns, 10*nmix, H, W] (x repeated per sample, color_image_decoder.py:108-113)."""
        size = z.size()
        b, ns = size[:2]
        x_expand = x.repeat_interleave(ns, dim=0) if ns > 1 else x
        return self(x_expand, z.view(b * ns, *size[2:]))

    def head_parts(self):
        """Sub-classes whose conv stack ends in a 1x1 ``Conv2dWeightNorm`` return ``(trunk, last, act)``: ``trunk(x, z)`` is
        the stack up to the INPUT of that convolution (before the ELU in front of it when ``act``) and ``last`` the
        convolution module; the likelihood kernel then contracts with ``last``'s weights itself and the raw parameter tensor
        is never materialised.  ``None``: no such split (the default)."""
        return None

    def reconstruct_error(self, x, z):
        """x [B,3,H,W] in [-1,1], z [B,ns,*z] -> [B,ns].  Reference: color_image_decoder.py:91-125."""
        ns = z.size(1)
        mode = ops.dmol_head_mode()
        parts = self.head_parts() if mode != "off" else None
        if parts is not None and (mode == "always" or not self._needs_grad(z)):
            trunk, last, act = parts
            size = z.size()
            x_expand = x.repeat_interleave(ns, dim=0) if ns > 1 else x
            hidden = trunk(x_expand, z.view(size[0] * ns, *size[2:]))
            if ops.dmol_head_supported(hidden, self.nmix):
                return ops.dmol_head_recon_error(hidden, last.conv.weight().flatten(1), last.conv.bias, x, ns, self.nmix, act)
            raw = last(F.elu(hidden) if act else hidden)
            return ops.dmol_recon_error(raw, x, ns, self.nmix)
        return ops.dmol_recon_error(self.likelihood_input(x, z), x, ns, self.nmix)

    def _needs_grad(self, z) -> bool:
        return torch.is_grad_enabled() and (z.requires_grad or any(p.requires_grad for p in self.parameters()))

    def log_probability(self, x, z):
        return self.reconstruct_error(x, z) * -1.
