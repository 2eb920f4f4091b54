"""CPU oracle for the MAE loss-and-evaluation hot path.  TEST INFRASTRUCTURE ONLY.

This file is a CPU restatement (torch CPU tensors, fp32 or fp64 depending on the dtype of the
inputs) of the arithmetic the reference performs on the hot path.  It is the *checker* for the
CUDA kernels in ``mae_b200/csrc``; it is never the thing shipped or measured.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference`` legs may import
it.  Nothing under ``mae_b200/`` imports this module (``tests/test_no_oracle_in_product.py`` checks).

Pinning: the reference has no tests, golden vectors or fixtures of its own, so
the oracle is pinned against outputs of the *live reference* (``/root/reference`` imported in the build
container by ``tests/golden/make_golden.py``); the resulting fixtures are committed under
``tests/golden/`` and ``tests/test_oracle_golden.py`` checks every function here against them
(bit-exact in fp32 on CPU for most, <=1e-6 relative where op fusion differs).

The arithmetic library underneath is the same as the reference's: PyTorch ATen CPU ops (torch 2.11
in this image; reference pinned ``torch>=0.4.1,<0.5.0`` in ``requirements.txt:8`` but runs unmodified).

Every function cites the reference file:line it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import math
from typing import Optional, Tuple

import torch
import torch.nn.functional as F

LOG_2PI = math.log(math.pi * 2.0)
EPS = 1e-12


# --------------------------------------------------------------------------------------------
# A1  reparameterised posterior sample            mae/modules/encoders/gaussian_encoder.py:56-62
# --------------------------------------------------------------------------------------------
def reparameterize(mu: torch.Tensor, logvar: torch.Tensor, eps: torch.Tensor) -> torch.Tensor:
    """z[b,s,...] = eps[b,s,...] * exp(0.5*logvar[b,...]) + mu[b,...].

    The reference draws ``eps`` itself with ``normal_()`` (gaussian_encoder.py:61); here it is an
    argument so that both sides of a parity test consume the same noise.
    """
    sigma = torch.exp(logvar * 0.5)
    return eps * sigma.unsqueeze(1) + mu.unsqueeze(1)


# --------------------------------------------------------------------------------------------
# A2  analytic KL(q(z|x) || N(0,I))                mae/modules/encoders/gaussian_encoder.py:219
# --------------------------------------------------------------------------------------------
def analytic_kl(mu: torch.Tensor, logvar: torch.Tensor) -> torch.Tensor:
    per_dim = mu.pow(2) + logvar.exp() - logvar - 1
    return 0.5 * per_dim.reshape(mu.size(0), -1).sum(dim=1)


# --------------------------------------------------------------------------------------------
# A3  mutual posterior divergence (MPD)           mae/modules/encoders/gaussian_encoder.py:162-188
# --------------------------------------------------------------------------------------------
def post_kl(mu: torch.Tensor, logvar: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """Returns (m_d with the shape of one latent, S scalar).

    Materialises the [B,B,*z] pairwise tensor exactly like the reference does, with the same
    operation order (ratio of variances, squared mean gap over the column variance, log-variance
    gap, diagonal zeroed by a mask multiply, ``cc``/``dd`` bias corrections).
    """
    nb = mu.size(0)
    zdims = mu.dim() - 1
    var = logvar.exp()
    denom = var.unsqueeze(0) + EPS                                   # indexed by column j
    ratio = var.unsqueeze(1) / denom                                 # :169
    gap = (mu.unsqueeze(1) - mu.unsqueeze(0)).pow(2) / denom         # :171
    lgap = logvar.unsqueeze(1) - logvar.unsqueeze(0)                 # :173
    ident = torch.eye(nb, device=mu.device)                          # fp32 like the reference (:176)
    pair = (ratio + gap - lgap - 1) * 0.5                            # :178
    pair = pair * (1.0 - ident.view(nb, nb, *([1] * zdims)))         # :179
    cc = nb / (nb - 1.0)
    m_d = pair.reshape(nb * nb, *mu.shape[1:]).mean(dim=0) * cc      # :183
    dd = math.sqrt((nb ** 2 - 1) / (nb ** 2 - nb - 1.0))
    total = pair.reshape(nb, nb, -1).sum(dim=2) + ident * m_d.sum()  # :187
    return m_d, total.std() * dd


def post_kl_chunked(mu: torch.Tensor, logvar: torch.Tensor, rows_per_chunk: int = 64
                    ) -> Tuple[torch.Tensor, torch.Tensor]:
    """Same quantity as :func:`post_kl`, evaluated in fp64 without the [B,B,D] tensor.

    Used where the reference cannot run (B >= ~4096).  Follows gaussian_encoder.py:162-188 row block
    by row block: identical per-pair formula and eps; the ``Eye*mean`` fill + unbiased ``std`` over B^2
    entries + ``dd`` factor of :185-187 is evaluated as written (sum of squares about the grand mean of
    the filled matrix), not through the off-diagonal shortcut the kernels use, so that the shortcut
    is itself under test.  Validated against :func:`post_kl` for B<=512 in tests/test_oracle_golden.py.
    """
    mu = mu.double().reshape(mu.size(0), -1)
    logvar = logvar.double().reshape(logvar.size(0), -1)
    nb, nd = mu.shape
    var = logvar.exp()
    denom = var + EPS
    acc_d = torch.zeros(nd, dtype=torch.float64, device=mu.device)
    row_totals = []
    for r0 in range(0, nb, rows_per_chunk):
        r1 = min(nb, r0 + rows_per_chunk)
        ratio = var[r0:r1, None, :] / denom[None, :, :]
        gap = (mu[r0:r1, None, :] - mu[None, :, :]).pow(2) / denom[None, :, :]
        lgap = logvar[r0:r1, None, :] - logvar[None, :, :]
        pair = (ratio + gap - lgap - 1) * 0.5
        idx = torch.arange(r0, r1, device=mu.device)
        pair[idx - r0, idx, :] = 0.0
        acc_d += pair.sum(dim=(0, 1))
        row_totals.append(pair.sum(dim=2))
    cc = nb / (nb - 1.0)
    m_d = acc_d / float(nb * nb) * cc
    total = torch.cat(row_totals, dim=0)
    total = total + torch.eye(nb, dtype=torch.float64, device=mu.device) * m_d.sum()
    dd = math.sqrt((nb ** 2 - 1) / (nb ** 2 - nb - 1.0))
    return m_d, total.std() * dd


# --------------------------------------------------------------------------------------------
# A3' MPD loss terms                                               mae/modules/mae.py:132-135
# --------------------------------------------------------------------------------------------
def mpd_loss_terms(m_d: torch.Tensor, s: torch.Tensor, eta: float, gamma: float):
    """(postKL_mean, postKL_std, loss_postKL_mean | 0., loss_postKL_std | 0.)."""
    loss_mean = -eta * F.logsigmoid(m_d).sum() if eta > 0.0 else 0.0
    loss_std = gamma * s if gamma > 0.0 else 0.0
    return m_d.sum(), s, loss_mean, loss_std


# --------------------------------------------------------------------------------------------
# A4a Bernoulli reconstruction error
#     mae/modules/decoders/image_decoders/binary_image_decoder.py:84-93 (dup. pixelcnn.py:136-139)
# --------------------------------------------------------------------------------------------
def bernoulli_recon_error(probs: torch.Tensor, x: torch.Tensor) -> torch.Tensor:
    """probs [B,ns,P] are post-sigmoid pixel probabilities, x [B,...] in {0,1}; returns [B,ns]."""
    nb = probs.size(0)
    xf = x.reshape(nb, -1).unsqueeze(1)
    ll = (probs + EPS).log() * xf + (1.0 - probs + EPS).log() * (1.0 - xf)
    return ll.sum(dim=2) * -1.0


# --------------------------------------------------------------------------------------------
# A4b DMOL parameter post-processing
#     mae/modules/decoders/image_decoders/color_image_decoder.py:38-52
# --------------------------------------------------------------------------------------------
def dmol_split(raw: torch.Tensor, nmix: int, nc: int = 3):
    """raw [N,(3*nc+1)*nmix,H,W] -> (means, log_scales, log_pi, coeffs), mixture-major/colour-minor."""
    n, _, h, w = raw.shape
    logits = raw[:, :nmix]
    means = raw[:, nmix:(nc + 1) * nmix].reshape(n, nmix, nc, h, w)
    lsc = raw[:, (nc + 1) * nmix:(2 * nc + 1) * nmix].reshape(n, nmix, nc, h, w)
    co = raw[:, (2 * nc + 1) * nmix:(3 * nc + 1) * nmix].reshape(n, nmix, nc, h, w)
    return means, lsc.clamp(min=-7.0), F.log_softmax(logits, dim=1), co.tanh()


# --------------------------------------------------------------------------------------------
# A7 helper: logsumexp                                          mae/modules/utils.py:23-38
# --------------------------------------------------------------------------------------------
def logsumexp(x: torch.Tensor, dim: int) -> torch.Tensor:
    peak, _ = x.max(dim, keepdim=True)
    return peak.squeeze(dim) + torch.log(torch.exp(x - peak).sum(dim))


# --------------------------------------------------------------------------------------------
# A4c discretized mixture of logistics                         mae/modules/utils.py:61-123
# --------------------------------------------------------------------------------------------
def dmol_loss(x: torch.Tensor, means: torch.Tensor, log_scales: torch.Tensor, coeffs: torch.Tensor,
              log_pi: torch.Tensor, bin_size: float = 1.0 / 255.0,
              lower: float = 1.0 / 255.0 - 1.0, upper: float = 1.0 - 1.0 / 255.0) -> torch.Tensor:
    """x [B,1,1,3,H,W]; means/log_scales/coeffs [B,ns,nmix,3,H,W]; log_pi [B,ns,nmix,H,W] -> [B,ns]."""
    x0, x1 = x[:, :, :, 0], x[:, :, :, 1]
    # colour coupling through the *true* sub-pixels (:81-85)
    c_mean = torch.stack([
        means[:, :, :, 0],
        means[:, :, :, 1] + coeffs[:, :, :, 0] * x0,
        means[:, :, :, 2] + coeffs[:, :, :, 1] * x0 + coeffs[:, :, :, 2] * x1,
    ], dim=3)
    cx = x - c_mean
    inv_s = torch.exp(-log_scales)
    lo_in = inv_s * (cx - bin_size)       # :94
    hi_in = inv_s * (cx + bin_size)       # :95
    mid_in = inv_s * cx                   # :96
    cdf_lo = torch.sigmoid(lo_in)
    cdf_hi = torch.sigmoid(hi_in)
    delta = cdf_hi - cdf_lo
    log_mid = torch.log(delta.clamp(min=EPS))                                              # :103
    log_approx = mid_in - log_scales - 2.0 * F.softplus(mid_in) + math.log(2 * bin_size)   # :104
    log_low = hi_in - F.softplus(hi_in)                                                    # :107
    log_up = -F.softplus(lo_in)                                                            # :110
    use_mid = delta.gt(1e-5).to(x.dtype)
    lc = log_mid * use_mid + log_approx * (1.0 - use_mid)
    ge_lower = x.ge(lower).to(x.dtype)
    le_upper = x.le(upper).to(x.dtype)
    lc = log_low * (1.0 - ge_lower) + lc * ge_lower
    lc = log_up * (1.0 - le_upper) + lc * le_upper
    per_pixel = logsumexp(lc.sum(dim=3) + log_pi, dim=2)        # :121
    return per_pixel.reshape(per_pixel.size(0), per_pixel.size(1), -1).sum(dim=2) * -1.0


def dmol_recon_error(raw: torch.Tensor, x: torch.Tensor, ns: int, nmix: int) -> torch.Tensor:
    """A4b+A4c from the raw conv output.  color_image_decoder.py:91-125.

    raw [B*ns,10*nmix,H,W] (sample index fastest inside a datum), x [B,3,H,W] in [-1,1] -> [B,ns].
    """
    nb = x.size(0)
    means, lsc, log_pi, co = dmol_split(raw, nmix)
    shp = lambda t: t.reshape(nb, ns, *t.shape[1:])
    return dmol_loss(x.unsqueeze(1).unsqueeze(1), shp(means), shp(lsc), shp(co), shp(log_pi))


# --------------------------------------------------------------------------------------------
# A5 / A6 log densities               mae/modules/encoders/gaussian_encoder.py:231-260, 262-301
# --------------------------------------------------------------------------------------------
def log_prob_prior(z_normal: torch.Tensor, logdet=0.0) -> torch.Tensor:
    """z_normal [N,...] -> [N]:  -0.5*sum(z^2 + log 2pi) + logdet   (:256-258)."""
    t = z_normal.pow(2) + LOG_2PI
    return t.reshape(z_normal.size(0), -1).sum(dim=1) * -0.5 + logdet


def log_prob_posterior(z_normal: torch.Tensor, mu: torch.Tensor, logvar: torch.Tensor,
                       logdet: torch.Tensor) -> torch.Tensor:
    """z_normal [B,K,...], mu/logvar [B,...], logdet [B,K] -> [B,K]   (:297-299)."""
    t = logvar.unsqueeze(1) + (z_normal - mu.unsqueeze(1)).pow(2).div(logvar.exp().unsqueeze(1) + EPS) + LOG_2PI
    return t.reshape(z_normal.size(0), z_normal.size(1), -1).sum(dim=2) * -0.5 + logdet


# --------------------------------------------------------------------------------------------
# A7 importance-weighted NLL                                     mae/modules/mae.py:209-215
# --------------------------------------------------------------------------------------------
def iw_nll(log_gen: torch.Tensor, log_prior: torch.Tensor, log_post: torch.Tensor):
    """all [B,K] -> ((nll_elbo, recon, kl), nll_iw), each [B]."""
    k = log_gen.size(1)
    log_iw = log_gen + log_prior - log_post
    nll_elbo = log_iw.mean(dim=1) * -1.0
    recon = log_gen.mean(dim=1) * -1.0
    kl = (log_post - log_prior).mean(dim=1)
    nll_iw = logsumexp(log_iw, dim=1) * -1.0 + math.log(k)
    return (nll_elbo, recon, kl), nll_iw


# --------------------------------------------------------------------------------------------
# A8 objective assembly                                          mae/modules/mae.py:126-140
# --------------------------------------------------------------------------------------------
def assemble_loss(recon_err: torch.Tensor, kl: torch.Tensor, m_d: torch.Tensor, s: torch.Tensor,
                  eta: float = 0.0, gamma: float = 0.0, free_bits: float = 0.0):
    """recon_err [B,ns], kl [B] -> the 7-tuple MAE.loss returns."""
    pk_mean, pk_std, l_mean, l_std = mpd_loss_terms(m_d, s, eta, gamma)
    recon = recon_err.mean(dim=1)
    loss = recon.mean() + kl.mean().clamp(min=free_bits) + l_mean + l_std
    return loss, recon, kl, pk_mean, pk_std, l_mean, l_std


# --------------------------------------------------------------------------------------------
# whole heads, as the experiment scripts drive them (synthetic encoder/decoder outputs)
# --------------------------------------------------------------------------------------------
def loss_head_binary(mu, logvar, eps, probs_fn, x, eta, gamma, free_bits=0.0):
    """Config C1/C2: reparam + analytic KL + MPD + Bernoulli + assembly.  ``probs_fn(z)`` stands in
    for the decoder conv stack and returns post-sigmoid probabilities [B,ns,P]."""
    z = reparameterize(mu, logvar, eps)
    kl = analytic_kl(mu, logvar)
    m_d, s = post_kl(mu, logvar)
    err = bernoulli_recon_error(probs_fn(z), x)
    return assemble_loss(err, kl, m_d, s, eta, gamma, free_bits), z


def loss_head_color(mu, logvar, eps, raw_fn, x, nmix, eta, gamma, free_bits=0.0):
    """Config C3: reparam + analytic KL + MPD + DMOL + assembly.  ``raw_fn(z)`` -> [B*ns,10*nmix,H,W]."""
    z = reparameterize(mu, logvar, eps)
    kl = analytic_kl(mu, logvar)
    m_d, s = post_kl(mu, logvar)
    err = dmol_recon_error(raw_fn(z), x, eps.size(1), nmix)
    return assemble_loss(err, kl, m_d, s, eta, gamma, free_bits), z


def iw_eval_color(raw, x, z, mu, logvar, nmix):
    """Config C4 for one batch of images, no flows: raw [B*K,10*nmix,H,W], z [B,K,D] -> as iw_nll."""
    nb, k = z.shape[:2]
    log_gen = dmol_recon_error(raw, x, k, nmix) * -1.0
    log_prior = log_prob_prior(z.reshape(nb * k, -1)).reshape(nb, k)
    log_post = log_prob_posterior(z, mu, logvar, z.new_zeros(nb, k))
    return iw_nll(log_gen, log_prior, log_post)
