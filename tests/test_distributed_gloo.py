"""world_size-2 gloo tests (CPU) of the batch-sharded head's host logic (mae_b200/distributed.py): the all-gather
plumbing, row-block bookkeeping, own-row gradients, the global KL mean for the free-bits clamp, the row-blocked
sum-of-squares all-reduce and the evaluation reduction.  The CUDA kernels cannot run here, so the collective layer is
driven with an engine built from the oracle (test infrastructure); the result must equal the single-process oracle on
the concatenated batch."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


class OracleEngine:
    """Same call surface as mae_b200.ops for the functions ShardedHead uses, computed by the oracle in fp64."""

    def __init__(self):
        from oracle import mae_oracle as O
        self.O = O

    def reparam_kl(self, mu, logvar, eps, want_kl=True):
        return self.O.reparameterize(mu, logvar, eps), (self.O.analytic_kl(mu, logvar) if want_kl else None)

    def _pairs(self, mu, lv):
        v = lv.exp()
        r = 1.0 / (v + 1e-12)
        return 0.5 * ((v[:, None] + (mu[:, None] - mu[None]) ** 2) * r[None] - (lv[:, None] - lv[None]) - 1).sum(-1)

    def mpd_rows_forward(self, mu_all, lv_all, row0, nrows, full):
        mu_all, lv_all = mu_all.detach(), lv_all.detach()
        m_d, s = self.O.post_kl(mu_all, lv_all)
        kl_all = self.O.analytic_kl(mu_all, lv_all)
        state = {"mu": mu_all, "lv": lv_all, "s": s}
        if full:
            return m_d, s, kl_all, _State(state)
        kl = self._pairs(mu_all, lv_all)
        b = mu_all.size(0)
        off = ~torch.eye(b, dtype=torch.bool)
        dev = ((kl - m_d.sum()) ** 2) * off
        return m_d, dev[row0:row0 + nrows].sum().reshape(1), kl_all, _State(state)

    def mpd_finalize(self, sumsq, batch):
        return torch.sqrt(sumsq[0] / (batch * (batch - 1) - 1))

    def mpd_rows_backward(self, state, row0, nrows, g_m, g_s, g_kl_rows):
        with torch.enable_grad():
            mu = state.d["mu"].clone().requires_grad_()
            lv = state.d["lv"].clone().requires_grad_()
            m_d, s = self.O.post_kl(mu, lv)
            f = mu.sum() * 0.0
            if g_m is not None:
                f = f + (m_d * g_m.reshape(m_d.shape)).sum()
            if g_s is not None:
                f = f + g_s * s
            if g_kl_rows is not None:
                f = f + (self.O.analytic_kl(mu, lv)[row0:row0 + nrows] * g_kl_rows).sum()
            f.backward()
        return mu.grad[row0:row0 + nrows], lv.grad[row0:row0 + nrows]

    def bernoulli_recon_error(self, probs, x):
        return self.O.bernoulli_recon_error(probs, x)

    def dmol_recon_error(self, raw, x, ns, nmix):
        return self.O.dmol_recon_error(raw, x, ns, nmix)

    def assemble_loss(self, err, kl, m_d, s, eta, gamma, free_bits, inv_batch, ext):
        recon = err.mean(dim=1)
        mean_rec = (recon.sum() + ext[0]) * inv_batch
        mean_kl = (kl.sum() + ext[1]) * inv_batch
        _, _, l_mean, l_std = self.O.mpd_loss_terms(m_d, s, eta, gamma)
        loss = mean_rec + mean_kl.clamp(min=free_bits) + l_mean + l_std
        stats = torch.stack([m_d.sum().detach(), torch.as_tensor(l_mean).detach().double(), torch.as_tensor(l_std).detach().double(),
                             mean_kl.detach(), mean_rec.detach()])
        return loss, stats, recon.detach()


class _State:
    def __init__(self, d):
        self.d = d
        self.s = d["s"]


class _Enc:
    def __init__(self, reg, kl, m_d, s, params):
        self.reg, self.kl, self.m_d, self.s, self.params = reg, kl, m_d, s, params


class OracleEngine2(OracleEngine):
    """adds the call surface of the two-node path (regulariser backward in the forward pass, loss head on reg)"""
    EncodedBatch = _Enc

    def mpd_rows_loss_backward(self, state, row0, nrows, eta, gamma, free_bits, kl_weight, m_d, kl_all, want_grad=True):
        O = self.O
        with torch.enable_grad():
            mu = state.d["mu"].clone().requires_grad_()
            lv = state.d["lv"].clone().requires_grad_()
            md, s = O.post_kl(mu, lv)
            reg = (O.analytic_kl(mu, lv).sum() * kl_weight).clamp(min=free_bits)
            _, _, l_mean, l_std = O.mpd_loss_terms(md, s, eta, gamma)
            reg = reg + l_mean + l_std
            reg.backward()
        grads = torch.stack([mu.grad[row0:row0 + nrows], lv.grad[row0:row0 + nrows]])
        return (grads if want_grad else None), reg.detach()

    def reparam_fwd(self, mu, logvar, eps):
        return self.O.reparameterize(mu.detach(), logvar.detach(), eps), (mu.detach(), logvar.detach(), eps)

    def reparam_reg_bwd(self, saved, gz, grads, g_reg):
        mu, lv, eps = saved
        dmu, dlv = torch.zeros_like(mu), torch.zeros_like(lv)
        if gz is not None:
            dmu = dmu + gz.sum(dim=1)
            dlv = dlv + 0.5 * (0.5 * lv).exp() * (gz * eps).sum(dim=1)
        if grads is not None and g_reg is not None:
            dmu, dlv = dmu + g_reg * grads[0], dlv + g_reg * grads[1]
        return dmu, dlv

    def loss_head(self, dec_out, x, enc, kind, ns, nmix, eta, gamma, free_bits, defer_join=False, inv_batch=0.0, kl=None,
                  ext_sums=None):
        err = self.dmol_recon_error(dec_out, x, ns, nmix) if kind == "dmol" else self.bernoulli_recon_error(dec_out, x)
        recon = err.mean(dim=1)
        mean_rec = recon.sum() * inv_batch
        loss = mean_rec + enc.reg
        _, _, l_mean, l_std = self.O.mpd_loss_terms(enc.m_d, enc.s, eta, gamma)
        stats = torch.stack([enc.m_d.sum().detach(), torch.as_tensor(l_mean).detach().double(),
                             torch.as_tensor(l_std).detach().double(), (enc.kl.sum() * inv_batch).detach(), mean_rec.detach()])
        return loss, stats, recon.detach()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, small_limit, kind, free_bits, out, two_node=False):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(1)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from mae_b200 import distributed as D, synth
        D.SMALL_GLOBAL_BATCH = small_limit
        head = D.ShardedHead(engine=OracleEngine2() if two_node else OracleEngine(), check_equal_batches=True)
        assert head._has_two_node() == two_node
        b_loc, ns, d = 5, 2, 6
        b = world * b_loc
        mu_g, lv_g = synth.posterior_params(b, (d,), seed=1, clamp=True)
        eps_g = synth.noise(b, ns, (d,), seed=1)
        sl = slice(rank * b_loc, (rank + 1) * b_loc)
        mu = mu_g[sl].double().requires_grad_()
        lv = lv_g[sl].double().requires_grad_()
        if kind == "bernoulli":
            inp = synth.binary_head(b, ns, 12, seed=2)
            dec_g, x_g = inp["probs"].double(), inp["x"].double()
            dec = dec_g[sl].clone().requires_grad_()
            loss, recon, kl, stats, z = head.loss_head_binary(mu, lv, eps_g[sl].double(), dec, x_g[sl], 1.0, 0.7, free_bits)
        else:
            inp = synth.color_head(b, ns, 3, 4, 4, seed=2)
            dec_g, x_g = inp["raw"].double(), inp["x"].double()
            dec = dec_g[rank * b_loc * ns:(rank + 1) * b_loc * ns].clone().requires_grad_()
            loss, recon, kl, stats, z = head.loss_head_color(mu, lv, eps_g[sl].double(), dec, x_g[sl], 3, 1.0, 0.7, free_bits)
        loss.backward()
        g_loss = head.global_loss(loss, stats)
        ev = head.reduce_eval(torch.tensor([float(rank + 1), 1.0]))
        assert head.shard(11) == ((0, 6) if rank == 0 else (6, 11))
        torch.save({"loss": g_loss, "dmu": mu.grad, "dlv": lv.grad, "ddec": dec.grad, "kl": kl.detach(), "ev": ev,
                    "mean_kl": stats[3]}, out % rank)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("small_limit,kind,free_bits,two_node", [
    (2048, "bernoulli", 0.0, False), (4, "bernoulli", 0.0, False), (2048, "dmol", 1e4, False), (4, "dmol", 0.0, False),
    (2048, "dmol", 0.0, True), (4, "bernoulli", 1e4, True)])
def test_sharded_head_matches_single_process(tmp_path, small_limit, kind, free_bits, two_node):
    from mae_b200 import synth
    from oracle import mae_oracle as O
    world = 2
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(world, _free_port(), small_limit, kind, free_bits, out, two_node), nprocs=world, join=True)
    res = [torch.load(out % r) for r in range(world)]
    b_loc, ns, d = 5, 2, 6
    b = world * b_loc
    mu_g, lv_g = synth.posterior_params(b, (d,), seed=1, clamp=True)
    eps_g = synth.noise(b, ns, (d,), seed=1).double()
    mu, lv = mu_g.double().requires_grad_(), lv_g.double().requires_grad_()
    if kind == "bernoulli":
        inp = synth.binary_head(b, ns, 12, seed=2)
        dec = inp["probs"].double().requires_grad_()
        ref, _ = O.loss_head_binary(mu, lv, eps_g, lambda z: dec, inp["x"].double(), 1.0, 0.7, free_bits)
    else:
        inp = synth.color_head(b, ns, 3, 4, 4, seed=2)
        dec = inp["raw"].double().requires_grad_()
        ref, _ = O.loss_head_color(mu, lv, eps_g, lambda z: dec, inp["x"].double(), 3, 1.0, 0.7, free_bits)
    ref[0].backward()
    for r in range(world):
        sl = slice(r * b_loc, (r + 1) * b_loc)
        assert torch.allclose(res[r]["loss"], ref[0].detach(), rtol=1e-10), (res[r]["loss"], ref[0])
        zero = lambda t, like: torch.zeros_like(like) if t is None else t
        assert torch.allclose(zero(res[r]["dmu"], mu_g[sl].double()), zero(mu.grad, mu)[sl], rtol=1e-8, atol=1e-12)
        assert torch.allclose(zero(res[r]["dlv"], mu_g[sl].double()), zero(lv.grad, lv)[sl], rtol=1e-8, atol=1e-12)
        n = b_loc * (ns if kind == "dmol" else 1)
        assert torch.allclose(res[r]["ddec"], dec.grad[r * n:(r + 1) * n], rtol=1e-8, atol=1e-14)
        assert torch.allclose(res[r]["kl"], ref[2].detach()[sl], rtol=1e-10)
        assert torch.allclose(res[r]["mean_kl"], ref[2].detach().mean(), rtol=1e-10)
        assert torch.equal(res[r]["ev"], torch.tensor([3.0, 2.0]))
