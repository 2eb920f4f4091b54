"""``MAE`` facade: same constructor, methods and return types as mae/modules/mae.py, with the bodies of
``loss`` (:100-140) and ``nll`` (:179-216) running on the fused kernels."""
from __future__ import annotations

import json
import os
from typing import Dict, Tuple

import numpy as np
import torch
import torch.nn as nn

from .. import ops
from .decoder import BinaryImageDecoder, ColorImageDecoder, Decoder
from .encoder import Encoder, GaussianEncoder


class MAE(nn.Module):
    def __init__(self, encoder: Encoder, decoder: Decoder):
        super().__init__()
        if int(np.prod(encoder.z_shape())) != int(np.prod(decoder.z_shape())):
            raise ValueError('encoder latent z shape %s does not match decoder input shape %s'
                             % (encoder.z_shape(), decoder.z_shape()))
        self.encoder = encoder
        self.decoder = decoder
        self._sharded_head = None     # mae_b200.distributed.shard_model

    def z_shape(self) -> Tuple:
        return self.encoder.z_shape()

    def sample_from_proir(self, nsamples=1, device=torch.device('cpu')):
        return self.encoder.sample_from_proir(nsamples, device=device)

    def sample_from_posterior(self, x, nsamples=1):
        return self.encoder.sample_from_posterior(x, nsamples)

    def encode(self, x, nsamples):
        return self.encoder.encode(x, nsamples)

    def decode(self, z, random_sample):
        return self.decoder.decode(z.view(z.size(0), *self.decoder.z_shape()), random_sample)

    def loss(self, x, nsamples, eta=0.0, gamma=0.0, free_bits=0.0):
        """-> (loss [], recon [B], KL [B], postKL_mean [], postKL_std [], loss_postKL_mean [] | 0., loss_postKL_std [] | 0.)

        Reference: mae/modules/mae.py:100-140.  ``loss`` carries the gradient; the other entries are logging
        values (the reference's training loops only ``.sum()``/print them)."""
        enc, dec = self.encoder, self.decoder
        z_shape_dec = dec.z_shape()
        if self._fusable():
            # two autograd nodes of ours around the decoder: (1) ops.encode_head = z plus the regulariser (KL + MPD, or MPD
            # only when flows make the KL a Monte-Carlo estimate) forward AND backward on the side stream, beside the
            # decoder (one cluster launch for B <= 128, D <= 64, else prepare -> tiles -> backward kernels back to back);
            # (2) ops.loss_head = likelihood (DMOL: forward+backward in one pass over the raw output) + objective
            mu, logvar = enc.core(x)
            eps = enc._draw_eps(mu, nsamples)
            flows = enc.prior_flow is not None or enc.posterior_flow is not None
            head = getattr(self, "_sharded_head", None)     # set by mae_b200.distributed.shard_model: x is this rank's shard
            if head is None:
                z, eb = ops.encode_head(mu, logvar, eps, eta, gamma, free_bits, with_kl=not flows)
            else:
                z, eb = head.encode_head(mu, logvar, eps, eta, gamma, free_bits, with_kl=not flows)
            KL = None
            if flows:
                # gaussian_encoder.py:220-225: KL = mean_s(log q(z|x) - log p(z)) through the flows (PyTorch) and the
                # log-density kernels
                size = z.size()
                z_normal = z
                if enc.posterior_flow is not None:
                    z, logdet = enc.posterior_flow.forward(z_normal.view(size[0] * size[1], *enc.posterior_flow.input_size()))
                    z, logdet = z.view(size), logdet.view(size[0], size[1])
                else:
                    logdet = None
                log_p = enc.log_probability_prior(z.reshape(size[0] * size[1], *size[2:])).view(size[0], size[1])
                log_q = ops.log_prob_posterior(z_normal, mu, logvar, logdet)
                KL = (log_q - log_p).mean(dim=1)
            dec_out = dec.likelihood_input(x, z.view(z.size(0), z.size(1), *z_shape_dec))
            if head is None:
                loss, stats, recon = ops.loss_head(dec_out, x, eb, dec.likelihood_kind, nsamples, getattr(dec, "nmix", 0), eta,
                                                   gamma, free_bits, kl=KL)
                if KL is None:
                    KL = eb.kl
            else:
                loss, stats, recon = head.loss_head(dec_out, x, eb, dec.likelihood_kind, nsamples, getattr(dec, "nmix", 0), eta,
                                                    gamma, free_bits, kl=KL)
                if KL is None:
                    row0 = head.rank * x.size(0)
                    KL = eb.kl[row0:row0 + x.size(0)]          # eb.kl is the all-gathered vector of the global batch
                head._last_share = stats[4]                    # this rank's share of the reconstruction mean (global_loss_value)
            m_d, S = eb.m_d, eb.s
        else:
            if getattr(self, "_sharded_head", None) is not None:
                raise RuntimeError("shard_model: batch-sharded MAE.loss needs the stock GaussianEncoder.encode and a stock "
                                   "Bernoulli / DMOL reconstruct_error (custom encoders / decoders are single-process only)")
            if isinstance(enc, GaussianEncoder) and type(enc).encode is GaussianEncoder.encode:
                z, KL, (m_d, S) = enc.encode(x, nsamples, _side_stream=True)   # MPD overlaps with the decoder
            else:
                z, KL, (m_d, S) = self.encode(x, nsamples)                     # subclass with the reference's signature
            err = dec.reconstruct_error(x, z.view(z.size(0), z.size(1), *z_shape_dec))
            loss, stats, recon = ops.assemble_loss(err, KL, m_d, S, eta, gamma, free_bits)
        loss_mean = stats[1] if eta > 0. else 0.
        loss_std = stats[2] if gamma > 0. else 0.
        return loss, recon, KL, stats[0], S, loss_mean, loss_std

    def _fusable(self) -> bool:
        """True when loss() can use the fused nodes: a stock Gaussian encoder (with or without flows) and a decoder that
        keeps the stock ``reconstruct_error`` of one of the two likelihood heads."""
        enc, dec = self.encoder, self.decoder
        if not isinstance(enc, GaussianEncoder):
            return False
        if type(enc).encode is not GaussianEncoder.encode:
            return False
        if isinstance(dec, ColorImageDecoder) and ops.dmol_head_mode() == "always" and dec.head_parts() is not None:
            return False      # the decoder's last convolution runs inside the likelihood kernel: reconstruct_error path
        for base in (BinaryImageDecoder, ColorImageDecoder):
            if isinstance(dec, base):
                return type(dec).reconstruct_error is base.reconstruct_error
        return False

    def reconstruct(self, x, k=50, random_sample=False):
        z, _ = self.sample_from_posterior(x, nsamples=k)
        return self.decode(z.mean(dim=1), random_sample=random_sample)

    def initialize(self, x, init_scale=1.0):
        z = self.encoder.initialize(x, init_scale=init_scale)
        return self.decoder.initialize(x, z.view(z.size(0), *self.decoder.z_shape()), init_scale=init_scale)

    def nll(self, x, k):
        """Importance-weighted NLL: -> ((nll_elbo [B], recon [B], kl [B]), nll_iw [B]).

        Reference: mae/modules/mae.py:179-216.  Evaluation-only (the reference calls it under no_grad)."""
        batch = x.size(0)
        with torch.no_grad():
            z, distr_params = self.sample_from_posterior(x, nsamples=k)
            enc = self.encoder
            if not (isinstance(distr_params, (tuple, list)) and len(distr_params) == 4):
                # an encoder whose posterior is not (mu, logvar, z_normal, logdet)-shaped: the reference's own composition
                # (mae.py:203-215) through its log_probability_* methods
                zd = z.view(batch, k, *self.decoder.z_shape())
                log_gen = self.decoder.log_probability(x, zd)
                log_prior = enc.log_probability_prior(z.view(batch * k, *z.size()[2:])).view(batch, k)
                log_post = enc.log_probability_posterior(x, z, distr_params=distr_params)
                log_iw = log_gen + log_prior - log_post
                nll_iw = -(torch.logsumexp(log_iw, dim=1) - float(np.log(k)))
                return (log_iw.mean(dim=1) * -1., log_gen.mean(dim=1) * -1., (log_post - log_prior).mean(dim=1)), nll_iw
            mu, logvar, z_post_normal, logdet_post = distr_params
            prior_flow = getattr(enc, 'prior_flow', None)
            if prior_flow is None:
                z_prior_normal, logdet_prior = z, None
            else:
                z_prior_normal, logdet_prior = prior_flow.backward(z.view(batch * k, *prior_flow.input_size()))
            post_flow = getattr(enc, 'posterior_flow', None)
            dec = self.decoder
            zd = z.view(batch, k, *dec.z_shape())
            stock = any(isinstance(dec, base) and type(dec).log_probability is base.log_probability
                        for base in (BinaryImageDecoder, ColorImageDecoder))
            if stock:   # log p(x|z) = -reconstruct_error: the IW kernel applies the sign, no negation pass
                log_gen, scale = dec.reconstruct_error(x, zd), -1.0
            else:
                log_gen, scale = dec.log_probability(x, zd), 1.0
            return ops.iw_nll(log_gen, z_prior_normal, logdet_prior, z_post_normal,
                              logdet_post if post_flow is not None else None, mu, logvar, gen_scale=scale)

    @classmethod
    def from_params(cls, params: Dict) -> "MAE":
        encoder_params = params.pop('encoder')
        encoder = Encoder.by_name(encoder_params.pop('type')).from_params(encoder_params)
        decoder_params = params.pop('decoder')
        decoder = Decoder.by_name(decoder_params.pop('type')).from_params(decoder_params)
        return MAE(encoder, decoder)

    @classmethod
    def load(cls, model_path, device) -> "MAE":
        params = json.load(open(os.path.join(model_path, 'config.json'), 'r'))
        mae = MAE.from_params(params)
        mae.load_state_dict(torch.load(os.path.join(model_path, 'model.pt'), map_location=device))
        return mae.to(device)
