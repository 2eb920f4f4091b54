"""Generate the golden fixtures in this directory from the LIVE reference (XuezheMax/mae).

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

The reference is imported unmodified from /root/reference; the only missing import on the
``mae.modules`` path is the ``overrides`` decorator package, for which a two-line shim is created in a
temporary directory.  Each fixture holds the inputs, the reference's fp32 outputs, and the outputs
and gradients of the same reference code run on ``.double()`` inputs (the fp64 ground truth used by the
fp64-anchored tolerance, SURVEY.md §7 "Hard parts" 3).  Fixtures are ``torch.save`` dicts of tensors.
"""
from __future__ import annotations

import math
import os
import sys
import tempfile

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

_shim = tempfile.mkdtemp(prefix="overrides_shim_")
os.makedirs(os.path.join(_shim, "overrides"))
with open(os.path.join(_shim, "overrides", "__init__.py"), "w") as f:
    f.write("def overrides(f):\n    return f\n")
sys.path.insert(0, _shim)
sys.path.insert(0, "/root/reference")

import warnings  # noqa: E402

warnings.filterwarnings("ignore")

from mae.modules import MAE as RefMAE  # noqa: E402
from mae.modules.encoders.gaussian_encoder import GaussianEncoder as RefGaussianEncoder  # noqa: E402
from mae.modules.encoders.encoder_cores.encoder_core import EncoderCore as RefEncoderCore  # noqa: E402
from mae.modules.flows.flow import Flow as RefFlow  # noqa: E402
from mae.modules.decoders.image_decoders.binary_image_decoder import BinaryImageDecoder as RefBinaryDecoder  # noqa: E402
from mae.modules.decoders.image_decoders.color_image_decoder import ColorImageDecoder as RefColorDecoder  # noqa: E402
from mae.modules import utils as ref_utils  # noqa: E402

from mae_b200 import synth  # noqa: E402

torch.set_num_threads(8)


class _ZShape:
    """Minimal ``self`` for the reference's unbound GaussianEncoder methods (they only use z_shape())."""

    def __init__(self, z_shape):
        self._z = tuple(z_shape)

    def z_shape(self):
        return self._z


def save(name, obj):
    path = os.path.join(HERE, name + ".pt")
    torch.save(obj, path)
    print("wrote %-28s %8.1f KB" % (name + ".pt", os.path.getsize(path) / 1024.0))


# ------------------------------------------------------------------------------------------- MPD
def ref_postkl(mu, logvar, eta=1.0, gamma=1.0, want_grad=True):
    mu = mu.clone().requires_grad_(want_grad)
    logvar = logvar.clone().requires_grad_(want_grad)
    stub = _ZShape(mu.shape[1:])
    m_d, s = RefGaussianEncoder._postKL(stub, mu, logvar)
    out = {"m_d": m_d.detach(), "S": s.detach()}
    if want_grad:
        # mae/modules/mae.py:134-135
        loss = -eta * torch.nn.functional.logsigmoid(m_d).sum() + gamma * s
        loss.backward()
        out.update(loss=loss.detach(), dmu=mu.grad, dlogvar=logvar.grad)
    return out


def golden_mpd():
    cases = {
        "b2_d3": (2, (3,), False), "b3_d1": (3, (1,), False), "b7_d5": (7, (5,), False),
        "b16_z4x2x2": (16, (4, 2, 2), True), "b64_d64": (64, (64,), True), "b100_d64": (100, (64,), False),
        "b100_d32": (100, (32,), False), "b130_d33": (130, (33,), False), "b64_z16x8x8": (64, (16, 8, 8), True),
        "b512_d64": (512, (64,), False),
    }
    out = {}
    for i, (name, (b, z, clamp)) in enumerate(cases.items()):
        mu, lv = synth.posterior_params(b, z, seed=synth.SEED + 100 + i, clamp=clamp)
        big = b * b * math.prod(z) > 4e6
        r32 = ref_postkl(mu, lv, want_grad=not big)
        r64 = ref_postkl(mu.double(), lv.double())
        out[name] = {"mu": mu, "logvar": lv, "eta": 1.0, "gamma": 1.0,
                     "ref32": r32, "ref64": r64}
    # posterior-collapse regime: all posteriors nearly identical (KL -> 0, cancellation in the reference)
    mu, lv = synth.posterior_params(32, (16,), seed=synth.SEED + 150)
    mu, lv = 0.3 + 1e-3 * mu, -0.2 + 1e-3 * lv
    out["collapse_b32_d16"] = {"mu": mu, "logvar": lv, "eta": 1.0, "gamma": 1.0,
                               "ref32": ref_postkl(mu, lv), "ref64": ref_postkl(mu.double(), lv.double())}
    # a large common offset on mu (translation invariance / centring)
    mu, lv = synth.posterior_params(48, (24,), seed=synth.SEED + 151)
    mu = mu + 50.0
    out["offset_b48_d24"] = {"mu": mu, "logvar": lv, "eta": 0.5, "gamma": 2.0,
                             "ref32": ref_postkl(mu, lv, 0.5, 2.0), "ref64": ref_postkl(mu.double(), lv.double(), 0.5, 2.0)}
    save("mpd", out)


# ------------------------------------------------------------------------------- reparam + KL
def golden_reparam_kl():
    out = {}
    for i, (b, ns, z) in enumerate([(5, 1, (7,)), (100, 1, (64,)), (8, 5, (4, 2, 2)), (64, 1, (64,)), (2, 128, (64,))]):
        mu, lv = synth.posterior_params(b, z, seed=synth.SEED + 200 + i, clamp=True)
        eps = synth.noise(b, ns, z, seed=synth.SEED + 200 + i)
        g = torch.Generator().manual_seed(77 + i)
        gz = torch.randn(b, ns, *z, generator=g)
        gkl = torch.randn(b, generator=g)
        res = {}
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            m, l, e = cast(mu).requires_grad_(), cast(lv).requires_grad_(), cast(eps)
            # gaussian_encoder.py:56-62 with eps injected: same two ops as the reference's last line
            std = l.mul(0.5).exp()
            zz = e.mul(std.unsqueeze(1)).add(m.unsqueeze(1))
            # check against the reference's own reparameterize by replaying its RNG draw
            torch.manual_seed(1234 + i)
            z_ref = RefGaussianEncoder.reparameterize(_ZShape(z), m, l, ns)
            torch.manual_seed(1234 + i)
            e_ref = std.new_empty(b, ns, *z).normal_()
            assert torch.equal(z_ref, e_ref.mul(std.unsqueeze(1)).add(m.unsqueeze(1)))
            kl = 0.5 * (m.pow(2) + l.exp() - l - 1).view(b, -1).sum(dim=1)  # gaussian_encoder.py:219
            (zz * cast(gz)).sum().add((kl * cast(gkl)).sum()).backward()
            res[tag] = {"z": zz.detach(), "KL": kl.detach(), "dmu": m.grad, "dlogvar": l.grad}
        out["b%d_ns%d_%s" % (b, ns, "x".join(map(str, z)))] = {"mu": mu, "logvar": lv, "eps": eps, "gz": gz, "gKL": gkl, **res}
    save("reparam_kl", out)


# ------------------------------------------------------------------------------------ Bernoulli
class _RefBinaryStub(RefBinaryDecoder):
    """reference BinaryImageDecoder whose conv stack is replaced by precomputed probabilities."""

    def __init__(self, probs):
        super().__init__(nz=1)
        self._probs = probs

    def forward(self, z):
        return self._probs.view(z.size(0), -1)


def golden_bernoulli():
    out = {}
    for i, (b, ns, p) in enumerate([(4, 1, 10), (100, 1, 784), (6, 5, 784), (2, 64, 100), (3, 2, 1)]):
        d = synth.binary_head(b, ns, p, seed=synth.SEED + 300 + i)
        probs, x = d["probs"], d["x"]
        if i == 0:  # exact 0 and 1 probabilities hit the eps guards
            probs[0, 0, :3] = torch.tensor([0.0, 1.0, 0.5])
        g = torch.randn(b, ns, generator=torch.Generator().manual_seed(5 + i))
        res = {}
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            pr = cast(probs).requires_grad_()
            dec = _RefBinaryStub(pr)
            z = pr.new_zeros(b, ns, 1, 1, 1)
            err = dec.reconstruct_error(cast(x).view(b, 1, p, 1), z)
            (err * cast(g)).sum().backward()
            res[tag] = {"err": err.detach(), "dprobs": pr.grad}
        out["b%d_ns%d_p%d" % (b, ns, p)] = {"probs": probs, "x": x, "g": g, **res}
    save("bernoulli", out)


# ----------------------------------------------------------------------------------------- DMOL
class _RefColorStub(RefColorDecoder):
    """reference ColorImageDecoder whose conv stack is replaced by a precomputed raw output."""

    def __init__(self, raw, nmix):
        super().__init__(nmix)
        self._raw = raw

    def forward(self, x, z):
        return self._raw


def ref_dmol(raw, x, ns, nmix, g, grad=True):
    raw = raw.clone().requires_grad_(grad)
    dec = _RefColorStub(raw, nmix)
    b = x.size(0)
    z = raw.new_zeros(b, ns, 1)
    err = dec.reconstruct_error(x, z)   # color_image_decoder.py:91-125 -> utils.py:61-123
    res = {"err": err.detach()}
    if grad:
        (err * g.to(err.dtype)).sum().backward()
        res["draw"] = raw.grad
    return res


def golden_dmol():
    out = {}
    small = {
        "b2_ns3_m10_8x8": dict(fn=synth.color_head, kw=dict(nimages=2, nsamples=3, nmix=10, height=8, width=8)),
        "b3_ns1_m5_4x4": dict(fn=synth.color_head, kw=dict(nimages=3, nsamples=1, nmix=5, height=4, width=4)),
        "b1_ns4_m1_5x3": dict(fn=synth.color_head, kw=dict(nimages=1, nsamples=4, nmix=1, height=5, width=3)),
        "hard_b3_ns2_m10_8x8": dict(fn=synth.color_head_hard, kw=dict(nimages=3, nsamples=2, nmix=10, height=8, width=8)),
        "hard_b2_ns1_m3_6x6": dict(fn=synth.color_head_hard, kw=dict(nimages=2, nsamples=1, nmix=3, height=6, width=6)),
    }
    for i, (name, spec) in enumerate(small.items()):
        d = spec["fn"](seed=synth.SEED + 400 + i, **spec["kw"])
        raw, x = d["raw"], d["x"]
        ns, nmix = spec["kw"]["nsamples"], spec["kw"]["nmix"]
        g = torch.randn(x.size(0), ns, generator=torch.Generator().manual_seed(9 + i))
        out[name] = {"raw": raw, "x": x, "ns": ns, "nmix": nmix, "g": g,
                     "ref32": ref_dmol(raw, x, ns, nmix, g), "ref64": ref_dmol(raw.double(), x.double(), ns, nmix, g)}
    save("dmol", out)

    # full-size C3 (B=64, 32x32, nmix=10): inputs are regenerated from the seed by the test;
    # only outputs and a gradient digest are stored.
    kw = dict(nimages=64, nsamples=1, nmix=10, height=32, width=32)
    d = synth.color_head(seed=synth.SEED, **kw)
    g = torch.full((64, 1), 1.0 / 64.0)
    r32 = ref_dmol(d["raw"], d["x"], 1, 10, g)
    r64 = ref_dmol(d["raw"].double(), d["x"].double(), 1, 10, g)
    pix = [(0, 0), (5, 17), (31, 31), (16, 2)]
    digest = {
        "input_checksum": d["raw"].double().sum(), "x_checksum": d["x"].double().sum(),
        "err32": r32["err"], "err64": r64["err"],
        "draw64_channel_sums": r64["draw"].sum(dim=(0, 2, 3)),
        "draw64_abs_sum": r64["draw"].abs().sum(),
        "draw64_pixels": torch.stack([r64["draw"][:, :, h, w] for h, w in pix], dim=0),
        "draw32_pixels": torch.stack([r32["draw"][:, :, h, w] for h, w in pix], dim=0),
        "pixels": torch.tensor(pix), "seed": synth.SEED,
    }
    save("dmol_c3_digest", digest)


# ---------------------------------------------------------------- log densities + IW estimator
def golden_iw():
    out = {}
    for i, (b, k, z) in enumerate([(1, 512, (64,)), (3, 17, (5,)), (2, 1024, (32,)), (4, 50, (4, 2, 2))]):
        mu, lv = synth.posterior_params(b, z, seed=synth.SEED + 500 + i, clamp=True)
        eps = synth.noise(b, k, z, seed=synth.SEED + 500 + i)
        g = torch.Generator().manual_seed(31 + i)
        res = {}
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            m, l, e = cast(mu), cast(lv), cast(eps)
            z_normal = e.mul(l.mul(0.5).exp().unsqueeze(1)).add(m.unsqueeze(1))
            logdet_post = cast(0.1 * torch.randn(b, k, generator=torch.Generator().manual_seed(41 + i)))
            # a stand-in prior flow output: any (z_prior_normal, logdet) pair is valid input for A5
            zp = cast(torch.randn(b * k, *z, generator=torch.Generator().manual_seed(51 + i)))
            logdet_prior = cast(0.1 * torch.randn(b * k, generator=torch.Generator().manual_seed(61 + i)))
            stub = _ZShape(z)
            stub.prior_flow = None
            lp_noflow = RefGaussianEncoder.log_probability_prior(stub, z_normal.view(b * k, *z))
            lp_flow = RefGaussianEncoder.log_probability_prior(stub, z_normal.view(b * k, *z), distr_params=(zp, logdet_prior))
            lq = RefGaussianEncoder.log_probability_posterior(stub, None, z_normal, distr_params=(m, l, z_normal, logdet_post))
            log_gen = cast(-3000.0 + 40.0 * torch.randn(b, k, generator=torch.Generator().manual_seed(71 + i)))
            # mae/modules/mae.py:209-215 with the reference's own logsumexp
            log_iw = log_gen + lp_flow.view(b, k) - lq
            nll_elbo = log_iw.mean(dim=1) * -1.
            recon = log_gen.mean(dim=1) * -1.
            kl = (lq - lp_flow.view(b, k)).mean(dim=1)
            nll_iw = ref_utils.logsumexp(log_iw, dim=1) * -1. + math.log(k)
            res[tag] = {"log_prior_noflow": lp_noflow, "log_prior_flow": lp_flow, "log_post": lq,
                        "nll_elbo": nll_elbo, "recon": recon, "kl": kl, "nll_iw": nll_iw}
            if tag == "ref32":
                inputs = {"mu": mu, "logvar": lv, "z_normal": z_normal, "logdet_post": logdet_post,
                          "z_prior_normal": zp, "logdet_prior": logdet_prior, "log_gen": log_gen}
        out["b%d_k%d_%s" % (b, k, "x".join(map(str, z)))] = {**inputs, **res}
    save("iw", out)


# ----------------------------------------------------------------- whole MAE with stub conv stacks
class _RefLinearCore(RefEncoderCore):
    """Stand-in for the conv encoder cores: one linear layer, ``chunk(2,1)`` like
    encoder_cores/resnet.py:43 (so mu is a strided view)."""

    def __init__(self, nin, nz, clamp):
        super().__init__()
        self.nz, self.clamp = nz, clamp
        self.linear = torch.nn.Linear(nin, 2 * nz)

    def forward(self, x):
        mu, logvar = self.linear(x.view(x.size(0), -1)).chunk(2, 1)
        if self.clamp:
            logvar = torch.nn.functional.hardtanh(logvar, min_val=-7., max_val=7.)
        return mu, logvar

    def output_size(self):
        return self.nz,


class _RefAffineFlow(RefFlow):
    """Stand-in for the MADE flows: elementwise affine, x = y*exp(a)+b going backward."""

    def __init__(self, nz):
        super().__init__()
        self.nz = nz
        self.a = torch.nn.Parameter(0.1 * torch.randn(nz))
        self.b = torch.nn.Parameter(0.1 * torch.randn(nz))

    def input_size(self):
        return self.nz,

    def forward(self, x, init=False, init_scale=1.0):
        return (x - self.b) * torch.exp(-self.a), -self.a.sum().expand(x.size(0))

    def backward(self, y, init=False, init_scale=1.0):
        return y * torch.exp(self.a) + self.b, self.a.sum().expand(y.size(0))


class _RefBinaryLinearDecoder(RefBinaryDecoder):
    def __init__(self, nz, npix):
        super().__init__(nz)
        self.linear = torch.nn.Linear(nz, npix)

    def z_shape(self):
        return self.nz,

    def forward(self, z):
        return torch.sigmoid(self.linear(z))


class _RefColorLinearDecoder(RefColorDecoder):
    def __init__(self, nz, nmix, h, w):
        super().__init__(nmix)
        self.nz, self.h, self.w = nz, h, w
        self.linear = torch.nn.Linear(nz, 10 * nmix * h * w)

    def z_shape(self):
        return self.nz,

    def forward(self, x, z):
        return self.linear(z).view(z.size(0), 10 * self.nmix, self.h, self.w)


class _InjectNoise:
    """Feeds a fixed fp32 noise tensor to the reference's ``new_empty(...).normal_()`` draw
    (gaussian_encoder.py:61) so that the fp32 and the fp64 runs of the UNMODIFIED reference consume the same eps
    (CPU normal_() yields different streams for float and double)."""

    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self._orig = torch.Tensor.normal_
        eps = self.eps

        def fake(t, *a, **k):
            assert t.shape == eps.shape, (t.shape, eps.shape)
            return t.copy_(eps)

        torch.Tensor.normal_ = fake
        return self

    def __exit__(self, *exc):
        torch.Tensor.normal_ = self._orig


def _run_ref_mae(mae, x, ns, eta, gamma, free_bits, eps_loss, eps_nll, k_nll):
    res = {}
    for tag in ("ref32", "ref64"):
        m = mae if tag == "ref32" else mae.double()
        xx = x if tag == "ref32" else x.double()
        m.zero_grad()
        with _InjectNoise(eps_loss):
            out = m.loss(xx, ns, eta=eta, gamma=gamma, free_bits=free_bits)
        out[0].backward()
        grads = {n: p.grad.clone() for n, p in m.named_parameters()}
        with _InjectNoise(eps_nll), torch.no_grad():
            (nll_elbo, recon, kl), nll_iw = m.nll(xx, k_nll)
        res[tag] = {"loss7": [o.detach() if torch.is_tensor(o) else torch.tensor(o) for o in out],
                    "grads": grads, "nll_elbo": nll_elbo, "nll_recon": recon, "nll_kl": kl, "nll_iw": nll_iw}
    mae.float()
    return res


def golden_mae():
    out = {}
    torch.manual_seed(4242)
    # binary, no flows (analytic KL branch, gaussian_encoder.py:213-219)
    b, npix, nz, ns, k = 12, 36, 8, 3, 20
    enc = RefGaussianEncoder(_RefLinearCore(npix, nz, False))
    mae = RefMAE(enc, _RefBinaryLinearDecoder(nz, npix))
    x = (torch.rand(b, 1, 6, 6) < 0.3).float()
    state = {n: p.detach().clone() for n, p in mae.state_dict().items()}
    seed = 99
    torch.manual_seed(seed); eps_loss = torch.empty(b, ns, nz).normal_()
    torch.manual_seed(seed); eps_nll = torch.empty(b, k, nz).normal_()
    out["binary_noflow"] = {"x": x, "state": state, "ns": ns, "k": k, "eta": 1.0, "gamma": 0.5, "free_bits": 0.0,
                            "eps_loss": eps_loss, "eps_nll": eps_nll, "nz": nz, "npix": npix,
                            **_run_ref_mae(mae, x, ns, 1.0, 0.5, 0.0, eps_loss, eps_nll, k)}
    # binary, prior flow (MC KL branch, :220-225) + free bits above the KL mean (clamp active)
    enc = RefGaussianEncoder(_RefLinearCore(npix, nz, False), prior_flow=_RefAffineFlow(nz))
    mae = RefMAE(enc, _RefBinaryLinearDecoder(nz, npix))
    state = {n: p.detach().clone() for n, p in mae.state_dict().items()}
    out["binary_priorflow"] = {"x": x, "state": state, "ns": ns, "k": k, "eta": 0.0, "gamma": 1.0, "free_bits": 1e4,
                               "eps_loss": eps_loss, "eps_nll": eps_nll, "nz": nz, "npix": npix,
                               **_run_ref_mae(mae, x, ns, 0.0, 1.0, 1e4, eps_loss, eps_nll, k)}
    # colour, no flows
    b, nz, nmix, h, w, ns, k = 5, 6, 4, 4, 4, 2, 16
    enc = RefGaussianEncoder(_RefLinearCore(3 * h * w, nz, True))
    mae = RefMAE(enc, _RefColorLinearDecoder(nz, nmix, h, w))
    x = torch.randint(0, 256, (b, 3, h, w)).float() / 127.5 - 1.0
    state = {n: p.detach().clone() for n, p in mae.state_dict().items()}
    torch.manual_seed(seed); eps_loss = torch.empty(b, ns, nz).normal_()
    torch.manual_seed(seed); eps_nll = torch.empty(b, k, nz).normal_()
    out["color_noflow"] = {"x": x, "state": state, "ns": ns, "k": k, "eta": 2.0, "gamma": 1.0, "free_bits": 0.0,
                           "eps_loss": eps_loss, "eps_nll": eps_nll, "nz": nz, "nmix": nmix, "h": h, "w": w,
                           **_run_ref_mae(mae, x, ns, 2.0, 1.0, 0.0, eps_loss, eps_nll, k)}
    save("mae_stub", out)


if __name__ == "__main__":
    golden_mpd()
    golden_reparam_kl()
    golden_bernoulli()
    golden_dmol()
    golden_iw()
    golden_mae()
