"""Pins oracle/mae_oracle.py against fixtures generated from the live reference
(tests/golden/make_golden.py).  CPU only."""
import math

import pytest
import torch

from oracle import mae_oracle as O
from tolerance import check_close

torch.set_num_threads(4)


def _both(case, fn):
    """run fn on fp32 and fp64 casts -> dict tag -> result"""
    return {"ref32": fn(lambda t: t.clone()), "ref64": fn(lambda t: t.double())}


def test_mpd_matches_reference(golden):
    for name, c in golden("mpd").items():
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            mu, lv = cast(c["mu"]).requires_grad_(), cast(c["logvar"]).requires_grad_()
            m_d, s = O.post_kl(mu, lv)
            ref = c[tag]
            assert m_d.shape == ref["m_d"].shape
            # same ATen ops in the same order: bit-exact on the same CPU, tiny slack for thread splits
            check_close(m_d, ref["m_d"], rel=2e-6 if tag == "ref32" else 1e-13, what="%s %s m_d" % (name, tag))
            check_close(s, ref["S"], rel=2e-6 if tag == "ref32" else 1e-13, what="%s %s S" % (name, tag))
            if "dmu" in ref:
                _, _, l_mean, l_std = O.mpd_loss_terms(m_d, s, c["eta"], c["gamma"])
                (l_mean + l_std).backward()
                check_close(mu.grad, ref["dmu"], rel=5e-6 if tag == "ref32" else 1e-12, what="%s %s dmu" % (name, tag))
                check_close(lv.grad, ref["dlogvar"], rel=5e-6 if tag == "ref32" else 1e-12, what="%s %s dlv" % (name, tag))


def test_mpd_chunked_restatement(golden):
    for name, c in golden("mpd").items():
        m_d, s = O.post_kl_chunked(c["mu"], c["logvar"], rows_per_chunk=37)
        ref = c["ref64"]
        check_close(m_d, ref["m_d"].reshape(-1), rel=1e-12, what=name + " chunked m_d")
        check_close(s, ref["S"], rel=1e-12, what=name + " chunked S")


def test_mpd_offdiagonal_identity(golden):
    """S == sqrt(sum_{i!=j}(KL_ij-m)^2 / (B(B-1)-1)) — the shortcut the kernels use."""
    for name, c in golden("mpd").items():
        mu, lv = c["mu"].double().flatten(1), c["logvar"].double().flatten(1)
        b = mu.size(0)
        v = lv.exp()
        r = 1.0 / (v + 1e-12)
        kl = 0.5 * ((v[:, None] + (mu[:, None] - mu[None]) ** 2) * r[None] - (lv[:, None] - lv[None]) - 1).sum(-1)
        off = ~torch.eye(b, dtype=torch.bool)
        m = kl[off].mean()
        s = torch.sqrt(((kl[off] - m) ** 2).sum() / (b * (b - 1) - 1))
        check_close(s, c["ref64"]["S"], rel=1e-11, what=name)
        check_close(m, c["ref64"]["m_d"].sum(), rel=1e-11, what=name)


def test_reparam_kl(golden):
    for name, c in golden("reparam_kl").items():
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            mu, lv = cast(c["mu"]).requires_grad_(), cast(c["logvar"]).requires_grad_()
            z = O.reparameterize(mu, lv, cast(c["eps"]))
            kl = O.analytic_kl(mu, lv)
            ((z * cast(c["gz"])).sum() + (kl * cast(c["gKL"])).sum()).backward()
            ref = c[tag]
            tol = 1e-6 if tag == "ref32" else 1e-13
            check_close(z, ref["z"], rel=tol, what=name + " z")
            check_close(kl, ref["KL"], rel=tol, what=name + " KL")
            check_close(mu.grad, ref["dmu"], rel=tol, what=name + " dmu")
            check_close(lv.grad, ref["dlogvar"], rel=tol, what=name + " dlv")


def test_bernoulli(golden):
    for name, c in golden("bernoulli").items():
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            pr = cast(c["probs"]).requires_grad_()
            err = O.bernoulli_recon_error(pr, cast(c["x"]))
            (err * cast(c["g"])).sum().backward()
            tol = 1e-6 if tag == "ref32" else 1e-13
            check_close(err, c[tag]["err"], rel=tol, what=name + " err")
            check_close(pr.grad, c[tag]["dprobs"], rel=tol, what=name + " dprobs")


def test_dmol(golden):
    for name, c in golden("dmol").items():
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            raw = cast(c["raw"]).requires_grad_()
            err = O.dmol_recon_error(raw, cast(c["x"]), c["ns"], c["nmix"])
            (err * cast(c["g"])).sum().backward()
            tol = 1e-6 if tag == "ref32" else 1e-13
            check_close(err, c[tag]["err"], rel=tol, what=name + " err")
            check_close(raw.grad, c[tag]["draw"], rel=tol, what=name + " draw")


def test_dmol_c3_digest(golden):
    from mae_b200 import synth
    d = golden("dmol_c3_digest")
    inp = synth.color_head(64, 1, 10, 32, 32, seed=int(d["seed"]))
    assert float(inp["raw"].double().sum()) == float(d["input_checksum"]), "synthetic generator drifted"
    assert float(inp["x"].double().sum()) == float(d["x_checksum"])
    raw = inp["raw"].double().requires_grad_()
    err = O.dmol_recon_error(raw, inp["x"].double(), 1, 10)
    (err / 64.0).sum().backward()
    check_close(err, d["err64"], rel=1e-13, what="C3 err64")
    check_close(raw.grad.sum(dim=(0, 2, 3)), d["draw64_channel_sums"], rel=1e-10, what="C3 channel sums")
    pix = d["pixels"]
    got = torch.stack([raw.grad[:, :, int(h), int(w)] for h, w in pix], dim=0)
    check_close(got, d["draw64_pixels"], rel=1e-12, what="C3 grad pixels")
    err32 = O.dmol_recon_error(inp["raw"], inp["x"], 1, 10)
    check_close(err32, d["err32"], rel=1e-6, what="C3 err32")


def test_log_probs_and_iw(golden):
    for name, c in golden("iw").items():
        for tag, cast in (("ref32", lambda t: t.clone()), ("ref64", lambda t: t.double())):
            ref = c[tag]
            zn = ref_z = cast(c["z_normal"]) if tag == "ref32" else None
            if tag == "ref64":
                # the fp64 reference recomputed z_normal in double; rebuild it the same way is not
                # possible from fp32 z, so compare fp64 oracle on fp32->fp64 inputs with a looser tol
                zn = c["z_normal"].double()
            b, k = zn.shape[:2]
            lp0 = O.log_prob_prior(zn.reshape(b * k, *zn.shape[2:]))
            lp1 = O.log_prob_prior(cast(c["z_prior_normal"]), cast(c["logdet_prior"]))
            lq = O.log_prob_posterior(zn, cast(c["mu"]), cast(c["logvar"]), cast(c["logdet_post"]))
            tol = 1e-6 if tag == "ref32" else 1e-6
            check_close(lp0, ref["log_prior_noflow"], rel=tol, what=name + " lp0 " + tag)
            check_close(lp1, ref["log_prior_flow"], rel=1e-6 if tag == "ref32" else 1e-13, what=name + " lp1 " + tag)
            check_close(lq, ref["log_post"], rel=tol, what=name + " lq " + tag)
            (e, r, kl), iw = O.iw_nll(cast(c["log_gen"]), ref["log_prior_flow"].view(b, k), ref["log_post"])
            t2 = 1e-6 if tag == "ref32" else 1e-13
            check_close(e, ref["nll_elbo"], rel=t2, what=name + " elbo")
            check_close(r, ref["recon"], rel=t2, what=name + " recon")
            check_close(kl, ref["kl"], rel=t2, what=name + " kl")
            check_close(iw, ref["nll_iw"], rel=t2, what=name + " iw")


def test_assemble_loss_types():
    m_d = torch.rand(4)
    s = torch.tensor(2.0)
    out = O.assemble_loss(torch.rand(3, 2), torch.rand(3), m_d, s, eta=0.0, gamma=0.0)
    assert out[5] == 0.0 and out[6] == 0.0 and isinstance(out[5], float)  # mae.py:134-135
    out = O.assemble_loss(torch.rand(3, 2), torch.rand(3), m_d, s, eta=1.0, gamma=1.0, free_bits=100.0)
    assert torch.is_tensor(out[5]) and torch.is_tensor(out[6])
