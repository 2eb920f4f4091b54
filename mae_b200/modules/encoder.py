"""Encoder side of the drop-in: ``Encoder`` / ``EncoderCore`` interfaces with their string registries
(mae/modules/encoders/encoder.py:8-121, encoder_cores/encoder_core.py:7-48) and ``GaussianEncoder``
(mae/modules/encoders/gaussian_encoder.py) whose hot methods run on the sm_100a kernels.

Conv cores and flows are ordinary ``nn.Module``s registered by name; they are not part of the hot path.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .flow import Flow


class EncoderCore(nn.Module):
    """Network producing ``(mu, logvar)``; mirror of encoder_cores/encoder_core.py:7-48."""
    _registry: Dict[str, type] = dict()

    def output_size(self) -> Tuple:
        raise NotImplementedError

    def initialize(self, x, init_scale=1.0):
        raise NotImplementedError

    @classmethod
    def register(cls, name: str):
        EncoderCore._registry[name] = cls

    @classmethod
    def by_name(cls, name: str):
        return EncoderCore._registry[name]

    @classmethod
    def from_params(cls, params: Dict):
        raise NotImplementedError


class Encoder(nn.Module):
    """Mirror of mae/modules/encoders/encoder.py:8-121."""
    _registry: Dict[str, type] = dict()

    def z_shape(self) -> Tuple:
        raise NotImplementedError

    def initialize(self, x, init_scale=1.0):
        raise NotImplementedError

    def sample_from_posterior(self, x, nsamples=1):
        raise NotImplementedError

    def sample_from_proir(self, nsamples=1, device=torch.device('cpu')):
        raise NotImplementedError

    def encode(self, x, nsamples):
        raise NotImplementedError

    def log_probability_prior(self, z):
        raise NotImplementedError

    def log_probability_posterior(self, x, z, distr_params=None):
        raise NotImplementedError

    @classmethod
    def register(cls, name: str):
        Encoder._registry[name] = cls

    @classmethod
    def by_name(cls, name: str):
        return Encoder._registry[name]

    @classmethod
    def from_params(cls, params: Dict):
        raise NotImplementedError


def _same_numel(a: Tuple, b: Tuple, what: str) -> None:
    if int(np.prod(a)) != int(np.prod(b)):
        raise ValueError(what % (a, b))


class GaussianEncoder(Encoder):
    """Diagonal-Gaussian posterior.  Same constructor, registry key ('gaussian') and method signatures as
    the reference class; no parameters or buffers of its own, so ``state_dict`` keys are unchanged."""

    def __init__(self, core: EncoderCore, prior_flow: Flow = None, posterior_flow: Flow = None, ngpu=1):
        super().__init__()
        self.ngpu = ngpu
        if prior_flow is not None:
            _same_numel(core.output_size(), prior_flow.input_size(), 'core output size %s does not match prior input size %s')
        if posterior_flow is not None:
            _same_numel(core.output_size(), posterior_flow.input_size(),
                        'core output size %s does not match posterior input size %s')
        if ngpu > 1:
            core = nn.DataParallel(core, device_ids=list(range(ngpu)))
        self.core = core
        self.prior_flow = prior_flow
        self.posterior_flow = posterior_flow

    def _core(self) -> EncoderCore:
        return self.core.module if isinstance(self.core, nn.DataParallel) else self.core

    def z_shape(self) -> Tuple:
        return self._core().output_size()

    # -- noise --------------------------------------------------------------------------------------
    @staticmethod
    def _draw_eps(mu: torch.Tensor, nsamples: int) -> torch.Tensor:
        """eps ~ N(0,1) of [B,ns,*z], drawn by ATen exactly like gaussian_encoder.py:61 (RNG parity)."""
        return mu.new_empty(mu.size(0), nsamples, *mu.size()[1:]).normal_()

    def reparameterize(self, mu, logvar, nsamples=1, eps=None):
        """z = eps*exp(0.5*logvar) + mu, [B,ns,*z].  Reference: gaussian_encoder.py:56-62."""
        if eps is None:
            eps = self._draw_eps(mu, nsamples)
        z, _ = ops.reparam_kl(mu, logvar, eps, want_kl=False)
        return z

    # -- data-dependent init (delegates to the PyTorch networks) --------------------------------------
    def initialize(self, x, init_scale=1.0):
        """Reference: gaussian_encoder.py:64-97."""
        mu, logvar = self._core().initialize(x, init_scale=init_scale)
        if self.posterior_flow is not None:
            z = mu.view(mu.size(0), *self.posterior_flow.input_size())
            z, _ = self.posterior_flow.forward(z, init=True, init_scale=init_scale * 0.1)
            z = z.view(mu.size())
        else:
            z = mu
        if self.prior_flow is not None:
            self.prior_flow.backward(z.view(mu.size(0), *self.prior_flow.input_size()), init=True, init_scale=init_scale * 0.1)
        return z

    # -- sampling -------------------------------------------------------------------------------------
    def _posterior(self, x, nsamples, want_kl):
        mu, logvar = self.core(x)
        eps = self._draw_eps(mu, nsamples)
        z_normal, kl = ops.reparam_kl(mu, logvar, eps, want_kl=want_kl)
        size = z_normal.size()
        if self.posterior_flow is not None:
            z, logdet = self.posterior_flow.forward(z_normal.view(size[0] * size[1], *self.posterior_flow.input_size()))
            z = z.view(size)
            logdet = logdet.view(size[0], size[1])
        else:
            z = z_normal
            logdet = z.new_zeros(size[0], size[1])
        return z, (mu, logvar, z_normal, logdet), kl

    def sample_from_posterior(self, x, nsamples=1):
        """-> z [B,ns,*z], (mu, logvar, z_normal, logdet).  Reference: gaussian_encoder.py:99-130."""
        z, params, _ = self._posterior(x, nsamples, want_kl=False)
        return z, params

    def sample_from_proir(self, nsamples=1, device=torch.device('cpu')):
        """Reference: gaussian_encoder.py:132-160 (spelling of the method name kept)."""
        z_size = self.z_shape()
        z_normal = torch.randn(nsamples, *z_size).to(device)
        if self.prior_flow is not None:
            z, logdet = self.prior_flow.forward(z_normal.view(nsamples, *self.prior_flow.input_size()))
            z = z.view(nsamples, *z_size)
        else:
            z = z_normal
            logdet = z.new_zeros(nsamples)
        return z, (z_normal, logdet)

    # -- MPD --------------------------------------------------------------------------------------------
    def _postKL(self, mu, logvar, side_stream=False):
        """(m_d [*z], S []).  Reference: gaussian_encoder.py:162-188; tiled CUDA kernels, never [B,B,D]."""
        return ops.mpd(mu, logvar, side_stream=side_stream)

    def encode(self, x, nsamples, _side_stream=False):
        """-> z [B,ns,*z], KL [B], (m_d [*z], S []).  Reference: gaussian_encoder.py:190-229.

        ``_side_stream`` (internal, used by ``MAE.loss``): run the MPD kernels on a side stream so that they overlap
        with the decoder; the caller must consume (m_d, S) through ``ops.assemble_loss`` / ``ops.join_side``."""
        no_flows = self.prior_flow is None and self.posterior_flow is None
        z, distr_params, kl = self._posterior(x, nsamples, want_kl=no_flows)
        mu, logvar = distr_params[0], distr_params[1]
        if not no_flows:
            size = z.size()
            log_p = self.log_probability_prior(z.view(size[0] * size[1], *size[2:])).view(size[0], size[1])
            log_q = self.log_probability_posterior(x, z, distr_params=distr_params)
            kl = (log_q - log_p).mean(dim=1)
        return z, kl, self._postKL(mu, logvar, side_stream=_side_stream)

    # -- log densities -----------------------------------------------------------------------------------
    def log_probability_prior(self, z, distr_params=None):
        """z [N,*z] -> [N].  Reference: gaussian_encoder.py:231-260."""
        if distr_params is not None:
            z_normal, logdet = distr_params
        elif self.prior_flow is None:
            z_normal, logdet = z, None
        else:
            z_normal, logdet = self.prior_flow.backward(z.view(z.size(0), *self.prior_flow.input_size()))
        if logdet is not None and not torch.is_tensor(logdet):
            logdet = None if float(logdet) == 0.0 else z_normal.new_full((z_normal.size(0),), float(logdet))
        return ops.log_prob_prior(z_normal, logdet)

    def log_probability_posterior(self, x, z, distr_params=None):
        """z [B,K,*z] -> [B,K].  Reference: gaussian_encoder.py:262-301."""
        size = z.size()
        if distr_params is None:
            mu, logvar = self.core(x)
            if self.posterior_flow is None:
                z_normal, logdet = z, None
            else:
                z_normal, logdet = self.posterior_flow.backward(z.view(size[0] * size[1], *self.posterior_flow.input_size()))
                z_normal = z_normal.view(size)
                logdet = logdet.view(size[0], size[1])
        else:
            mu, logvar, z_normal, logdet = distr_params
        return ops.log_prob_posterior(z_normal, mu, logvar, logdet)

    @classmethod
    def from_params(cls, params: Dict) -> "GaussianEncoder":
        core_params = params.pop('core')
        core = EncoderCore.by_name(core_params.pop('type')).from_params(core_params)
        posterior_flow = None
        pf = params.pop('posterior_flow', None)
        if pf is not None:
            posterior_flow = Flow.by_name(pf.pop('type')).from_params(pf)
        prior_flow = None
        pf = params.pop('prior_flow', None)
        if pf is not None:
            prior_flow = Flow.by_name(pf.pop('type')).from_params(pf)
        return GaussianEncoder(core, prior_flow=prior_flow, posterior_flow=posterior_flow, **params)


GaussianEncoder.register('gaussian')
