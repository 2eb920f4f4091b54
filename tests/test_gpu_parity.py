"""Parity of the CUDA path (through the C ABI, via mae_b200.ops / mae_b200.modules) against

  * the committed golden fixtures generated from the live reference (tests/golden/*.pt), and
  * the CPU oracle (oracle/mae_oracle.py) on seeded inputs at sizes it finishes in seconds, and
  * size-independent properties at BASELINE.json's full sizes.

Tolerances (tests/tolerance.py): 1e-5 relative to the tensor scale for loss terms and gradients, with the
fp64-anchored escape max(1e-5*scale, 4*|ref32-ref64|); 1e-4 bits/dim for the evaluation NLL.
Run on the B200 box:  python -m pytest tests -m gpu -q
"""
import math

import pytest
import torch

from tolerance import check_bpd, check_close

pytestmark = pytest.mark.gpu

DEV = "cuda"


@pytest.fixture(scope="module")
def ops():
    from mae_b200 import _lib, ops as _ops
    _lib.load()
    return _ops


@pytest.fixture(scope="module")
def O():
    from oracle import mae_oracle
    torch.set_num_threads(max(1, torch.get_num_threads()))
    return mae_oracle


def cu(t):
    return t.to(DEV)


# ---------------------------------------------------------------------------------------------- A1 + A2
def test_reparam_kl_golden(ops, golden):
    for name, c in golden("reparam_kl").items():
        mu, lv = cu(c["mu"]).requires_grad_(), cu(c["logvar"]).requires_grad_()
        z, kl = ops.reparam_kl(mu, lv, cu(c["eps"]))
        ((z * cu(c["gz"])).sum() + (kl * cu(c["gKL"])).sum()).backward()
        r64, r32 = c["ref64"], c["ref32"]
        check_close(z, r64["z"], r32["z"], what=name + " z")
        check_close(kl, r64["KL"], r32["KL"], what=name + " KL")
        check_close(mu.grad, r64["dmu"], r32["dmu"], what=name + " dmu")
        check_close(lv.grad, r64["dlogvar"], r32["dlogvar"], what=name + " dlogvar")


def test_reparam_strided_chunk_views(ops, O):
    torch.manual_seed(1)
    out = torch.randn(37, 2 * 24, device=DEV, requires_grad=True)
    mu, lv = out.chunk(2, 1)                      # row stride 48, like encoder_cores/resnet.py:43
    eps = torch.randn(37, 3, 24, device=DEV)
    z, kl = ops.reparam_kl(mu, lv, eps)
    (z.sum() * 0.3 + kl.sum()).backward()
    o = out.detach().double().cpu().requires_grad_()
    m, l = o.chunk(2, 1)
    zr = O.reparameterize(m, l, eps.double().cpu())
    (zr.sum() * 0.3 + O.analytic_kl(m, l).sum()).backward()
    check_close(z, zr, what="z")
    check_close(out.grad, o.grad, what="grad through chunk views")


# ---------------------------------------------------------------------------------------------------- A3
@pytest.mark.parametrize("path", [1, 2, 3])
def test_mpd_golden(ops, golden, path):
    for name, c in golden("mpd").items():
        mu, lv = cu(c["mu"]).requires_grad_(), cu(c["logvar"]).requires_grad_()
        m_d, s = ops.mpd(mu, lv, bwd_path=path)
        r64, r32 = c["ref64"], c["ref32"]
        assert m_d.shape == r64["m_d"].shape, name
        check_close(m_d, r64["m_d"], r32["m_d"], what="%s m_d" % name)
        check_close(s, r64["S"], r32["S"], what="%s S" % name)
        loss = -c["eta"] * torch.nn.functional.logsigmoid(m_d).sum() + c["gamma"] * s
        loss.backward()
        check_close(mu.grad, r64["dmu"], r32.get("dmu"), what="%s dmu (path %d)" % (name, path))
        check_close(lv.grad, r64["dlogvar"], r32.get("dlogvar"), what="%s dlogvar (path %d)" % (name, path))


def test_mpd_partial_upstream_grads(ops, O):
    """only g_S, only g_m, and a non-uniform g_m"""
    torch.manual_seed(3)
    mu0, lv0 = torch.randn(21, 13), torch.randn(21, 13)
    w = torch.randn(13)
    for mode in ("S", "m", "both"):
        for path in (1, 2, 3):
            mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
            m_d, s = ops.mpd(mu, lv, bwd_path=path)
            f = {"S": lambda m, s_: 2.0 * s_, "m": lambda m, s_: (m * w.to(m)).sum(),
                 "both": lambda m, s_: (m * w.to(m)).sum() - 0.5 * s_}[mode]
            f(m_d, s).backward()
            a, b = mu0.double().requires_grad_(), lv0.double().requires_grad_()
            mr, sr = O.post_kl(a, b)
            f(mr, sr).backward()
            check_close(mu.grad, a.grad, what="dmu %s path %d" % (mode, path))
            check_close(lv.grad, b.grad, what="dlv %s path %d" % (mode, path))


@pytest.mark.parametrize("path", [2, 3])
def test_mpd_tile_backward_multi_chunk(ops, O, path):
    """D > 64 -> several K chunks per row block; B not a multiple of the tile; strided inputs."""
    torch.manual_seed(4)
    out = torch.randn(150, 2 * 70)
    g = torch.randn(70)
    o_gpu = cu(out).requires_grad_()
    mu, lv = o_gpu.chunk(2, 1)
    m_d, s = ops.mpd(mu, lv, bwd_path=path)
    ((m_d * cu(g)).sum() + 0.7 * s).backward()
    o = out.double().requires_grad_()
    a, b = o.chunk(2, 1)
    mr, sr = O.post_kl(a, b)
    ((mr * g.double()).sum() + 0.7 * sr).backward()
    check_close(m_d, mr, what="m_d")
    check_close(s, sr, what="S")
    check_close(o_gpu.grad, o.grad, what="grad (tile path, 3 chunks)")


def test_mpd_large_vs_chunked_oracle(ops, O):
    """B > 2048: 128 x 128 forward tiles (padded batch 3072 here, 2176 = 17 x 128 for B = 2100), with and without the
    stored KL matrices."""
    from mae_b200 import synth
    for b, seed in ((3000, 11), (2100, 15)):
        mu, lv = synth.posterior_params(b, (32,), seed=seed)
        m_ref, s_ref = O.post_kl_chunked(mu, lv, rows_per_chunk=100)
        m_d, s = ops.mpd(cu(mu), cu(lv))
        check_close(m_d, m_ref, what="m_d B=%d" % b)
        check_close(s, s_ref, what="S B=%d" % b)
        m_g, s_g = ops.mpd(cu(mu).requires_grad_(), cu(lv).requires_grad_())     # keeps KL / KL^T (path 3)
        assert torch.equal(m_g, m_d) and torch.equal(s_g, s)


def test_mpd_large_backward_vs_chunked_oracle(ops, O):
    """B = 2100 > 2048: the automatic choice is the stored-KL tile backward (path 3); checked against autograd through
    the row-chunked fp64 oracle, and against the recompute path."""
    from mae_b200 import synth
    b, d = 2100, 12
    mu0, lv0 = synth.posterior_params(b, (d,), seed=16)
    a, c = mu0.double().requires_grad_(), lv0.double().requires_grad_()
    m_ref, s_ref = O.post_kl_chunked(a, c, rows_per_chunk=100)
    (-0.7 * torch.nn.functional.logsigmoid(m_ref).sum() + 1.3 * s_ref).backward()
    g = {}
    for path in (0, 2):
        mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
        m_d, s = ops.mpd(mu, lv, bwd_path=path)
        (-0.7 * torch.nn.functional.logsigmoid(m_d).sum() + 1.3 * s).backward()
        check_close(mu.grad, a.grad, what="dmu B=2100 path %d" % path)
        check_close(lv.grad, c.grad, what="dlv B=2100 path %d" % path)
        g[path] = (mu.grad, lv.grad)
    assert ops._mpd_pick_path(b, d, 0) == 3


@pytest.fixture
def mpd_tc():
    """switch the tensor-core MPD tiles on / off for one test (mae_set_mpd_tensor_cores), restoring the previous setting"""
    from mae_b200 import _lib
    lib = _lib.load()
    prev = lib.mae_set_mpd_tensor_cores(1)
    lib.mae_set_mpd_tensor_cores(prev)

    def set_(on):
        lib.mae_set_mpd_tensor_cores(1 if on else 0)

    yield set_
    lib.mae_set_mpd_tensor_cores(prev)


@pytest.mark.parametrize("regime", ["random", "collapse", "offset"])
def test_mpd_tensor_core_tiles_vs_chunked_oracle(ops, O, mpd_tc, regime):
    """C5 tiles on the tensor cores (3xTF32 tcgen05.mma, csrc/mpd_tc.cuh; padded batch >= 2048, K = 2D a multiple of 128):
    forward (m_d, S) and both backward paths against autograd through the row-chunked fp64 oracle, with the fp32 oracle as
    the measure of what fp32 resolves; beside it the FP32 FFMA tiles on the same inputs.  Regimes: random posteriors,
    posterior collapse (every posterior within 1e-3 of the same one: the reference's A + B - C - 1 cancels) and a common
    offset of 30 sigma (hard part 3)."""
    from mae_b200 import synth
    b, d = 2304, 64
    mu0, lv0 = synth.posterior_params(b, (d,), seed=31)
    if regime == "collapse":
        mu0, lv0 = 0.3 + 1e-3 * mu0, -0.5 + 1e-3 * lv0
    elif regime == "offset":
        mu0 = mu0 + 30.0
    ref = {}
    for key, dt in (("r64", torch.float64), ("r32", torch.float32)):
        a, c = mu0.to(dt).clone().requires_grad_(), lv0.to(dt).clone().requires_grad_()
        m_ref, s_ref = O.post_kl_chunked(a, c, rows_per_chunk=128)
        (-0.7 * torch.nn.functional.logsigmoid(m_ref).sum() + 1.3 * s_ref).backward()
        ref[key] = (m_ref.detach(), s_ref.detach(), a.grad, c.grad)
    worst = {}
    for tc in (1, 0):
        mpd_tc(tc)
        for path in (3, 2):
            mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
            m_d, s = ops.mpd(mu, lv, bwd_path=path)
            (-0.7 * torch.nn.functional.logsigmoid(m_d).sum() + 1.3 * s).backward()
            got = (m_d, s, mu.grad, lv.grad)
            for i, nm in enumerate(("m_d", "S", "dmu", "dlv")):
                e = check_close(got[i], ref["r64"][i], ref["r32"][i], what="%s %s tc=%d path=%d" % (regime, nm, tc, path))
                worst[(tc, nm)] = max(worst.get((tc, nm), 0.0), e)
    print("MPD %s B=%d D=%d, error / tensor scale vs fp64: " % (regime, b, d)
          + ", ".join("%s %s %.1e" % ("tensor" if tc else "FFMA", nm, e) for (tc, nm), e in sorted(worst.items())))


@pytest.mark.parametrize("b,d", [(2500, 64), (3000, 32), (2200, 96), (2100, 256)])
def test_mpd_tensor_core_ragged_sizes(ops, O, mpd_tc, b, d):
    """batches that are not a multiple of the 128-row tile (padded rows and columns are masked in the epilogue and in the
    weights) at every contraction length the tensor-core kernels treat differently: K = 2D = 128 (row operand resident in
    TMEM, contiguous tile ranges, column-half helpers), K = 64 (two chunks per tile, generic issue loop), K = 192 and 512
    (streamed row operand, round-robin tiles, several column chunks in the backward)."""
    from mae_b200 import synth
    mu0, lv0 = synth.posterior_params(b, (d,), seed=37)
    ref = {}
    for key, dt in (("r64", torch.float64), ("r32", torch.float32)):
        a, c = cu(mu0).to(dt).requires_grad_(), cu(lv0).to(dt).requires_grad_()
        m_ref, s_ref = O.post_kl_chunked(a, c, rows_per_chunk=256)
        (-0.7 * torch.nn.functional.logsigmoid(m_ref).sum() + 1.3 * s_ref).backward()
        ref[key] = (m_ref.detach(), s_ref.detach(), a.grad, c.grad)
    mpd_tc(1)
    for path in (3, 2):
        mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
        m_d, s = ops.mpd(mu, lv, bwd_path=path)
        (-0.7 * torch.nn.functional.logsigmoid(m_d).sum() + 1.3 * s).backward()
        for got, i, nm in ((m_d, 0, "m_d"), (s, 1, "S"), (mu.grad, 2, "dmu"), (lv.grad, 3, "dlv")):
            check_close(got, ref["r64"][i], ref["r32"][i], what="B=%d D=%d %s path=%d" % (b, d, nm, path))


def test_mpd_tensor_core_row_blocks(ops, mpd_tc):
    """row blocks (the sharded decomposition) through the tensor-core tiles: partial sums add up to the full batch's, the
    stored KL_ij / KL_ji rows of a block give the same own-row gradients as the full batch"""
    from mae_b200 import synth
    b, d = 2560, 64
    mu0, lv0 = synth.posterior_params(b, (d,), seed=33)
    mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
    mpd_tc(1)
    m_full, s_full = ops.mpd(mu, lv, bwd_path=3)
    (-torch.nn.functional.logsigmoid(m_full).sum() + s_full).backward()
    parts = []
    for r0, n in ((0, 640), (640, 1000), (1640, 920)):
        m_p, ss, kl_all, state = ops.mpd_rows_forward(mu.detach(), lv.detach(), r0, n, full=False)
        assert torch.equal(m_p, m_full.detach()) and state.path == 3
        parts.append(ss.double())
        ops._pool.give(state.ws)
    tot = torch.stack(parts).sum()
    s2 = ops.mpd_finalize(tot.reshape(1), b)
    check_close(s2, s_full.detach().double(), rel=2e-6, what="S from row-block partial sums")
    # own-row gradients of a block, upstream gradients as in the full run
    g_m = -torch.sigmoid(-m_full.detach()).reshape(-1)
    for r0, n in ((640, 1000),):
        m_p, ss, kl_all, state = ops.mpd_rows_forward(mu.detach(), lv.detach(), r0, n, full=False)
        state.s = s_full.detach().reshape(1)
        dmu, dlv = ops.mpd_rows_backward(state, r0, n, g_m.contiguous(), torch.ones(1, device=DEV), None)
        check_close(dmu, mu.grad[r0:r0 + n].double(), rel=1e-5, what="row-block dmu")
        check_close(dlv, lv.grad[r0:r0 + n].double(), rel=1e-5, what="row-block dlv")


def test_mpd_c5_shape_pinned_against_fp64_oracle(ops, O, mpd_tc):
    """One C5 shape (B = 8192, D = 64: the reference itself cannot materialise its [B, B, D] tensor, 34 GB in fp64) pinned
    against the row-chunked fp64 restatement of _postKL (gaussian_encoder.py:162-188), evaluated on the GPU in fp64:
    m_d and S of the tensor-core tiles and of the FP32 FFMA tiles."""
    from mae_b200 import synth
    b, d = 8192, 64
    mu0, lv0 = synth.posterior_params(b, (d,), seed=41)
    with torch.no_grad():
        m_ref, s_ref = O.post_kl_chunked(cu(mu0).double(), cu(lv0).double(), rows_per_chunk=64)
    for tc in (1, 0):
        mpd_tc(tc)
        m_d, s, _ = ops.mpd_forward_only(cu(mu0), cu(lv0))
        e_m = check_close(m_d, m_ref, what="m_d B=8192 tc=%d" % tc)
        e_s = check_close(s, s_ref, what="S B=8192 tc=%d" % tc)
        print("MPD B=8192 D=64 vs fp64 oracle, %s tiles: m_d %.1e, S %.1e" % ("tensor-core" if tc else "FFMA", e_m, e_s))


def test_mpd_row_blocks_add_up_and_deterministic(ops):
    """sum of row-block partial sums == full sum (the multi-GPU decomposition); bitwise repeatability."""
    from mae_b200 import synth
    mu, lv = synth.posterior_params(1000, (64,), seed=12)
    mu, lv = cu(mu), cu(lv)
    m_full, s_full, ss_full = ops.mpd_forward_only(mu, lv)
    parts = []
    for r0, n in ((0, 125), (125, 125), (250, 250), (500, 499), (999, 1)):
        m_p, _, ss = ops.mpd_forward_only(mu, lv, r0, n)
        assert torch.equal(m_p, m_full)
        parts.append(ss)
    tot = torch.stack(parts).sum()
    assert abs(float(tot) - float(ss_full)) <= 1e-9 * abs(float(ss_full))
    s2 = ops.mpd_finalize(tot.reshape(1), 1000)
    check_close(s2, s_full, rel=1e-6, what="S from row blocks")
    m2, s_again, ss2 = ops.mpd_forward_only(mu, lv)
    assert torch.equal(ss2, ss_full) and torch.equal(s_again, s_full)


def test_mpd_full_size_properties(ops):
    """BASELINE C5 size (B=16384, D=64; the [B,B,D] tensor would be 69 GB): permutation and translation
    invariance of (m_d, S), and gradients summing to zero along the batch for mu (translation invariance)."""
    from mae_b200 import synth
    mu, lv = synth.posterior_params(16384, (64,), seed=13)
    mu, lv = cu(mu), cu(lv)
    m1, s1, _ = ops.mpd_forward_only(mu, lv)
    perm = torch.randperm(16384, device=DEV, generator=torch.Generator(DEV).manual_seed(5))
    m2, s2, _ = ops.mpd_forward_only(mu[perm], lv[perm])
    check_close(m2, m1, rel=2e-6, what="m_d permutation")
    check_close(s2, s1, rel=2e-6, what="S permutation")
    m3, s3, _ = ops.mpd_forward_only(mu + 3.0, lv)
    check_close(m3, m1, rel=1e-5, what="m_d translation")
    check_close(s3, s1, rel=1e-5, what="S translation")


def test_mpd_maximum_size_c5(ops):
    """BASELINE C5 maximum (B = 65536, D = 256: the reference's [B,B,D] tensor would be 4.4 TB): the forward's permutation
    invariance; the default backward keeps the two B x B matrices (34 GB of the 180 GB, inside the storage budget) and
    must agree with the recompute backward, which stores nothing; translation property sum_i d/dmu_i = 0 and finiteness."""
    from mae_b200 import synth
    b, d = 65536, 256
    mu, lv = synth.posterior_params(b, (d,), seed=21)
    mu, lv = cu(mu), cu(lv)
    m1, s1, _ = ops.mpd_forward_only(mu, lv)
    perm = torch.randperm(b, device=DEV, generator=torch.Generator(DEV).manual_seed(6))
    m2, s2, _ = ops.mpd_forward_only(mu[perm], lv[perm])
    check_close(m2, m1, rel=2e-6, what="m_d permutation (B=65536)")
    check_close(s2, s1, rel=2e-6, what="S permutation (B=65536)")
    del perm, m2, s2
    assert ops._mpd_pick_path(b, d, 0) == (3 if ops.MPD_KL_BUDGET_BYTES >= (35 << 30) else 2)
    mu.requires_grad_()
    lv.requires_grad_()
    grads = {}
    for path in (0, 2):
        mu.grad = lv.grad = None
        m_d, s = ops.mpd(mu, lv, bwd_path=path)
        assert torch.equal(m_d, m1) and torch.equal(s, s1)
        (-torch.nn.functional.logsigmoid(m_d).sum() + s).backward()
        assert torch.isfinite(mu.grad).all() and torch.isfinite(lv.grad).all()
        col = mu.grad.double().sum(dim=0)
        assert float(col.abs().max()) <= 1e-4 * float(mu.grad.double().abs().sum(dim=0).max())
        grads[path] = (mu.grad.clone(), lv.grad.clone())
        del m_d, s
        torch.cuda.empty_cache()
    check_close(grads[0][0], grads[2][0].double(), rel=2e-5, what="dmu stored-KL vs recompute (B=65536)")
    check_close(grads[0][1], grads[2][1].double(), rel=2e-5, what="dlv stored-KL vs recompute (B=65536)")


def test_mpd_backward_translation_property_tile_path(ops):
    from mae_b200 import synth
    mu, lv = synth.posterior_params(2500, (64,), seed=14)   # > 2048 -> stored-KL tile path automatically
    mu, lv = cu(mu).requires_grad_(), cu(lv).requires_grad_()
    m_d, s = ops.mpd(mu, lv)
    (torch.nn.functional.logsigmoid(m_d).sum() * -1.0 + s).backward()
    col = mu.grad.sum(dim=0)                                  # d/d(common shift) == 0
    assert float(col.abs().max()) <= 1e-4 * float(mu.grad.abs().sum(dim=0).max())
    # and the two backward paths agree on a sub-problem that both can run
    mu_s, lv_s = mu.detach()[:700].clone().requires_grad_(), lv.detach()[:700].clone().requires_grad_()
    g = {}
    for path in (1, 2, 3):
        mu_s.grad = lv_s.grad = None
        m_d, s = ops.mpd(mu_s, lv_s, bwd_path=path)
        (torch.nn.functional.logsigmoid(m_d).sum() * -1.0 + s).backward()
        g[path] = (mu_s.grad.clone(), lv_s.grad.clone())
    check_close(g[2][0], g[1][0].double(), rel=2e-5, what="dmu path2 vs path1")
    check_close(g[2][1], g[1][1].double(), rel=2e-5, what="dlv path2 vs path1")
    check_close(g[3][0], g[1][0].double(), rel=2e-5, what="dmu path3 vs path1")
    check_close(g[3][1], g[1][1].double(), rel=2e-5, what="dlv path3 vs path1")


def test_mpd_rejects_batch_of_one(ops):
    with pytest.raises(RuntimeError, match="at least 2"):
        ops.mpd(torch.zeros(1, 4, device=DEV), torch.zeros(1, 4, device=DEV))


# --------------------------------------------------------------------------------------------------- A4a
def test_bernoulli_golden(ops, golden):
    for name, c in golden("bernoulli").items():
        pr = cu(c["probs"]).requires_grad_()
        err = ops.bernoulli_recon_error(pr, cu(c["x"]))
        (err * cu(c["g"])).sum().backward()
        check_close(err, c["ref64"]["err"], c["ref32"]["err"], what=name + " err")
        check_close(pr.grad, c["ref64"]["dprobs"], c["ref32"]["dprobs"], what=name + " dprobs")


def test_bernoulli_nll_eval_shape(ops, O):
    """[1,4096,784]: the shape calc_nll drives (experiments/binary_image.py:49,271)."""
    from mae_b200 import synth
    d = synth.binary_head(1, 4096, 784, seed=21)
    err = ops.bernoulli_recon_error(cu(d["probs"]), cu(d["x"]))
    ref = O.bernoulli_recon_error(d["probs"].double(), d["x"].double())
    check_close(err, ref, what="bernoulli K=4096")


# --------------------------------------------------------------------------------------------- A4b + A4c
def test_dmol_golden(ops, golden):
    for name, c in golden("dmol").items():
        raw = cu(c["raw"]).requires_grad_()
        err = ops.dmol_recon_error(raw, cu(c["x"]), c["ns"], c["nmix"])
        (err * cu(c["g"])).sum().backward()
        check_close(err, c["ref64"]["err"], c["ref32"]["err"], what=name + " err")
        check_close(raw.grad, c["ref64"]["draw"], c["ref32"]["draw"], what=name + " draw")


def test_dmol_c3_full_size_digest(ops, golden):
    from mae_b200 import synth
    d = golden("dmol_c3_digest")
    inp = synth.color_head(64, 1, 10, 32, 32, seed=int(d["seed"]))
    assert float(inp["raw"].double().sum()) == float(d["input_checksum"])
    raw = cu(inp["raw"]).requires_grad_()
    err = ops.dmol_recon_error(raw, cu(inp["x"]), 1, 10)
    (err / 64.0).sum().backward()
    check_close(err, d["err64"], d["err32"], what="C3 err")
    g = raw.grad
    sc = float(d["draw64_abs_sum"]) / g.numel()          # mean |grad|
    check_close(g.double().sum(dim=(0, 2, 3)), d["draw64_channel_sums"], rel=1e-5, floor=1e-5 * sc * 64 * 1024 ** 0.5,
                what="C3 per-channel gradient sums")
    got = torch.stack([g[:, :, int(h), int(w)] for h, w in d["pixels"]], dim=0)
    check_close(got, d["draw64_pixels"], d["draw32_pixels"], what="C3 gradient pixels")
    # deterministic
    err2 = ops.dmol_recon_error(raw.detach(), cu(inp["x"]), 1, 10)
    assert torch.equal(err2, err.detach())


def test_dmol_iw_shape_vs_oracle(ops, O):
    """one C4 image: K=512 samples of [100,32,32] (209.7 MB) against the CPU oracle in fp64 (a few seconds)."""
    from mae_b200 import synth
    inp = synth.color_head(1, 512, 10, 32, 32, seed=31)
    err = ops.dmol_recon_error(cu(inp["raw"]), cu(inp["x"]), 512, 10)
    ref = O.dmol_recon_error(inp["raw"].double(), inp["x"].double(), 512, 10)
    check_close(err, ref, what="dmol K=512")
    check_bpd(err, ref, 3072, what="dmol K=512 bits/dim")


def test_dmol_shape_errors(ops):
    with pytest.raises(RuntimeError, match="inconsistent"):
        ops.dmol_recon_error(torch.zeros(2, 99, 4, 4, device=DEV), torch.zeros(2, 3, 4, 4, device=DEV), 1, 10)


# ------------------------------------------------------------------------------------------- A5, A6, A7
def test_log_probs_and_iw_golden(ops, golden):
    for name, c in golden("iw").items():
        r64, r32 = c["ref64"], c["ref32"]
        zn = cu(c["z_normal"])
        b, k = zn.shape[:2]
        lp0 = ops.log_prob_prior(zn.reshape(b * k, *zn.shape[2:]))
        lp1 = ops.log_prob_prior(cu(c["z_prior_normal"]), cu(c["logdet_prior"]))
        lq = ops.log_prob_posterior(zn, cu(c["mu"]), cu(c["logvar"]), cu(c["logdet_post"]))
        check_close(lp0, r32["log_prior_noflow"].double(), rel=1e-5, what=name + " log prior")
        check_close(lp1, r64["log_prior_flow"], r32["log_prior_flow"], what=name + " log prior (flow)")
        check_close(lq, r32["log_post"].double(), rel=1e-5, what=name + " log posterior")
        (e, r, kl), iw = ops.iw_nll(cu(c["log_gen"]), cu(c["z_prior_normal"]).view(b, k, -1), cu(c["logdet_prior"]).view(b, k),
                                    zn, cu(c["logdet_post"]), cu(c["mu"]), cu(c["logvar"]))
        # reference fp32 chain as the anchor (its z_normal is the fp32 tensor we feed).  kl = mean_k(log q - log p) is a
        # difference of two log densities of magnitude up to ~1e3 that agree to a few units: each of them carries half an
        # fp32 ulp of ITS magnitude in either implementation, so the difference is resolved to a few ulps of max|log q|,
        # |log p| (not of |kl|) - the floor below is 8 such ulps (2e-4 ... 1e-3 for these fixtures), derived, not tuned
        ulp_big = 2.0 ** -23 * max(float(lp1.abs().max()), float(lq.abs().max()))
        for got, key in ((e, "nll_elbo"), (r, "recon"), (kl, "kl"), (iw, "nll_iw")):
            check_close(got, r32[key].double(), rel=1e-5, floor=8.0 * ulp_big if key == "kl" else 0.0, what="%s %s" % (name, key))


def test_log_prob_gradients(ops, O):
    torch.manual_seed(7)
    b, k, d = 6, 5, 9
    z0, mu0, lv0 = torch.randn(b, k, d), torch.randn(b, d), torch.randn(b, d)
    ld0, g = torch.randn(b, k), torch.randn(b, k)
    z, mu, lv, ld = (cu(t).requires_grad_() for t in (z0, mu0, lv0, ld0))
    (ops.log_prob_posterior(z, mu, lv, ld) * cu(g)).sum().backward()
    zr, mr, lr, ldr = (t.double().requires_grad_() for t in (z0, mu0, lv0, ld0))
    (O.log_prob_posterior(zr, mr, lr, ldr) * g.double()).sum().backward()
    for got, ref, n in ((z, zr, "dz"), (mu, mr, "dmu"), (lv, lr, "dlv"), (ld, ldr, "dlogdet")):
        check_close(got.grad, ref.grad, what="posterior " + n)
    zp = cu(z0.reshape(b * k, d)).requires_grad_()
    ldp = cu(ld0.reshape(-1)).requires_grad_()
    (ops.log_prob_prior(zp, ldp) * cu(g.reshape(-1))).sum().backward()
    zpr = z0.reshape(b * k, d).double().requires_grad_()
    (O.log_prob_prior(zpr, ld0.reshape(-1).double()) * g.reshape(-1).double()).sum().backward()
    check_close(zp.grad, zpr.grad, what="prior dz")
    check_close(ldp.grad, g.reshape(-1), what="prior dlogdet")


def test_iw_eval_color_end_to_end(ops, O):
    """C4 head for 2 images x K=64 on 16x16: DMOL fwd + fused IW kernel vs oracle; bits/dim tolerance."""
    from mae_b200 import synth
    b, k, d = 2, 64, 64
    inp = synth.color_head(b, k, 10, 16, 16, seed=41)
    mu, lv = synth.posterior_params(b, (d,), seed=42, clamp=True)
    z = O.reparameterize(mu, lv, synth.noise(b, k, (d,), seed=43))
    log_gen = -ops.dmol_recon_error(cu(inp["raw"]), cu(inp["x"]), k, 10)
    (e, r, kl), iw = ops.iw_nll(log_gen, cu(z), None, cu(z), None, cu(mu), cu(lv))
    (e_r, r_r, kl_r), iw_r = O.iw_eval_color(inp["raw"].double(), inp["x"].double(), z.double(), mu.double(), lv.double(), 10)
    nx = 3 * 16 * 16
    for got, ref, n in ((e, e_r, "elbo"), (r, r_r, "recon"), (iw, iw_r, "iw")):
        check_close(got, ref, what="C4 " + n)
        check_bpd(got, ref, nx, what="C4 bits/dim " + n)
    check_close(kl, kl_r, rel=1e-5, what="C4 kl")


# ---------------------------------------------------------------------------------------------- A3' + A8
@pytest.mark.parametrize("eta,gamma,free_bits", [(1.0, 1.0, 0.0), (0.0, 0.0, 0.0), (2.0, 0.0, 1e6), (0.0, 0.5, 3.0)])
def test_assemble_loss(ops, O, eta, gamma, free_bits):
    torch.manual_seed(9)
    b, ns, d = 37, 3, 20
    err0, kl0, m0, s0 = torch.rand(b, ns) * 100, torch.rand(b) * 5, torch.randn(d) * 3, torch.rand(()) * 10
    err, kl, m, s = (cu(t).requires_grad_() for t in (err0, kl0, m0, s0))
    loss, stats, recon = ops.assemble_loss(err, kl, m, s, eta, gamma, free_bits)
    (loss * 1.7).backward()
    er, kr, mr, sr = (t.double().requires_grad_() for t in (err0, kl0, m0, s0))
    ref = O.assemble_loss(er, kr, mr, sr, eta, gamma, free_bits)
    (ref[0] * 1.7).backward()
    check_close(loss, ref[0], what="loss")
    check_close(recon, ref[1], what="recon")
    check_close(stats[0], ref[3], what="postKL_mean")
    check_close(stats[1], torch.as_tensor(ref[5]), what="loss_postKL_mean", floor=1e-12)
    check_close(stats[2], torch.as_tensor(ref[6]), what="loss_postKL_std", floor=1e-12)
    zero = lambda t: torch.zeros_like(t) if t.grad is None else t.grad
    check_close(err.grad, zero(er), what="derr")
    check_close(kl.grad, zero(kr), what="dkl", floor=1e-12)
    check_close(zero(m), zero(mr), what="dm_d", floor=1e-12)
    check_close(zero(s), zero(sr), what="dS", floor=1e-12)


# ------------------------------------------------------------------------ fused nodes used by MAE.loss
def test_kl_mpd_fused_node(ops, O):
    torch.manual_seed(11)
    for b, d, side in ((9, 5, False), (64, 64, True), (100, 32, True), (2100, 16, False)):
        mu0, lv0 = torch.randn(b, d), torch.randn(b, d).clamp(-7, 7)
        gk, gm = torch.randn(b), torch.randn(d)
        mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
        kl, m_d, s = ops.kl_mpd(mu, lv, side_stream=side)
        if side:
            ops.join_side(mu.device)
        ((kl * cu(gk)).sum() + (m_d * cu(gm)).sum() + 0.3 * s).backward()
        torch.cuda.synchronize()
        a, c = mu0.double().requires_grad_(), lv0.double().requires_grad_()
        if b <= 512:
            mr, sr = O.post_kl(a, c)
            klr = O.analytic_kl(a, c)
            ((klr * gk.double()).sum() + (mr * gm.double()).sum() + 0.3 * sr).backward()
            check_close(mu.grad, a.grad, what="kl_mpd dmu b=%d" % b)
            check_close(lv.grad, c.grad, what="kl_mpd dlv b=%d" % b)
        else:
            mr, sr = O.post_kl_chunked(mu0, lv0, 100)
            klr = O.analytic_kl(a, c)
        check_close(kl, klr, what="kl b=%d" % b)
        check_close(m_d, mr, what="m_d b=%d" % b)
        check_close(s, sr, what="S b=%d" % b)


def _reg_reference(O, mu0, lv0, eta, gamma, free_bits):
    """fp64 oracle of the encoder-side regulariser of mae.py:132-139 and its gradient."""
    a, c = mu0.double().requires_grad_(), lv0.double().requires_grad_()
    kl = O.analytic_kl(a, c)
    m_d, s = O.post_kl(a, c)
    reg = kl.mean().clamp(min=free_bits)
    if eta > 0:
        reg = reg - eta * torch.nn.functional.logsigmoid(m_d).sum()
    if gamma > 0:
        reg = reg + gamma * s
    reg.backward()
    return kl.detach(), m_d.detach(), s.detach(), a.grad, c.grad, reg.detach()


@pytest.mark.parametrize("b,d", [(2, 1), (5, 3), (8, 64), (9, 7), (16, 16), (17, 5), (21, 13), (64, 64), (65, 33), (100, 32),
                                 (100, 64), (104, 64), (128, 64)])
def test_fused_regulariser_kernel_vs_oracle(ops, O, b, d):
    """mae_kl_mpd_loss_fused (one cluster launch: KL, m_d, S and d reg / d(mu, logvar)) against the fp64 oracle, for every
    cluster size 1..16, ragged last CTA, D below / at the 64-dim limit, and the eta / gamma / free-bits switches."""
    from mae_b200 import synth
    mu0, lv0 = synth.posterior_params(b, (d,), seed=100 + b, clamp=True)
    assert ops.fused_reg_supported(b, d, DEV)
    klm = float(O.analytic_kl(mu0.double(), lv0.double()).mean())
    for eta, gamma, fb in ((1.0, 1.0, 0.0), (0.0, 2.0, 0.0), (0.7, 0.0, 0.0), (1.0, 1.0, 2.0 * klm + 1.0)):
        kl, m_d, s, reg, grads, _ = ops.kl_mpd_reg_fused(cu(mu0), cu(lv0), eta, gamma, fb)
        klr, mr, sr, gmu, glv, regr = _reg_reference(O, mu0, lv0, eta, gamma, fb)
        tag = "B=%d D=%d eta=%g gamma=%g fb=%g" % (b, d, eta, gamma, fb)
        check_close(kl, klr, what="KL " + tag)
        check_close(m_d, mr, what="m_d " + tag)
        check_close(s, sr, what="S " + tag)
        check_close(reg, regr, what="reg " + tag)
        check_close(grads[0], gmu, what="dmu " + tag, floor=1e-12)
        check_close(grads[1], glv, what="dlv " + tag, floor=1e-12)
    # forward only (evaluation): same values, no gradient buffers
    kl2, m2, s2, _, g2, _ = ops.kl_mpd_reg_fused(cu(mu0), cu(lv0), 1.0, 1.0, 0.0, want_grad=False)
    assert g2 is None
    kl1, m1, s1, _, _, _ = ops.kl_mpd_reg_fused(cu(mu0), cu(lv0), 1.0, 1.0, 0.0)
    assert torch.equal(kl1, kl2) and torch.equal(m1, m2) and torch.equal(s1, s2)   # deterministic


@pytest.mark.parametrize("b,d", [(150, 70), (64, 1024), (300, 16), (2100, 8)])
def test_general_regulariser_chain_vs_oracle(ops, O, b, d):
    """ops.kl_mpd_reg outside the cluster kernel's shapes: prepare -> tiles -> device-side upstream gradients -> backward
    kernel (paths 1 and 3), same contract as the one-launch kernel (reg value and d reg / d(mu, logvar))."""
    from mae_b200 import synth
    mu0, lv0 = synth.posterior_params(b, (d,), seed=200 + b, clamp=True)
    assert not ops.fused_reg_supported(b, d, DEV)
    klm = float(O.analytic_kl(mu0.double(), lv0.double()).mean())
    for eta, gamma, fb in ((1.0, 1.0, 0.0), (0.0, 2.0, 0.0), (0.7, 0.0, 2.0 * klm + 1.0)):
        kl, m_d, s, reg, grads, _ = ops.kl_mpd_reg(cu(mu0), cu(lv0), eta, gamma, fb)
        a, c = mu0.double().requires_grad_(), lv0.double().requires_grad_()
        klr = O.analytic_kl(a, c)
        mr, sr = O.post_kl(a, c) if b <= 512 else O.post_kl_chunked(a, c, rows_per_chunk=100)
        regr = klr.mean().clamp(min=fb)
        if eta > 0:
            regr = regr - eta * torch.nn.functional.logsigmoid(mr).sum()
        if gamma > 0:
            regr = regr + gamma * sr
        regr.backward()
        tag = "B=%d D=%d eta=%g gamma=%g fb=%g" % (b, d, eta, gamma, fb)
        check_close(kl, klr, what="KL " + tag)
        check_close(m_d, mr, what="m_d " + tag)
        check_close(s, sr, what="S " + tag)
        check_close(reg, regr, what="reg " + tag)
        check_close(grads[0], a.grad, what="dmu " + tag, floor=1e-12)
        check_close(grads[1], c.grad, what="dlv " + tag, floor=1e-12)
    kl2, m2, s2, reg2, g2, _ = ops.kl_mpd_reg(cu(mu0), cu(lv0), 1.0, 1.0, 0.0, want_grad=False)
    assert g2 is None
    kl1, m1, s1, reg1, _, _ = ops.kl_mpd_reg(cu(mu0), cu(lv0), 1.0, 1.0, 0.0)
    assert torch.equal(m1, m2) and torch.equal(s1, s2) and torch.equal(reg1, reg2)


def test_fused_regulariser_goldens_and_strided_views(ops, O, golden):
    """the reference's own _postKL goldens (incl. the posterior-collapse regime) through the fused kernel, fed with the
    strided chunk views an encoder core returns"""
    n = 0
    for name, c in golden("mpd").items():
        b, d = c["mu"].size(0), c["mu"][0].numel()
        if b > 128 or d > 64:
            continue
        n += 1
        packed = cu(torch.cat([c["mu"].reshape(b, d), c["logvar"].reshape(b, d)], dim=1))
        mu, lv = packed.chunk(2, 1)
        eta, gamma = float(c["eta"]), float(c["gamma"])
        kl, m_d, s, _, grads, _ = ops.kl_mpd_reg_fused(mu, lv, eta, gamma, 0.0, kl_weight=0.0)
        r64, r32 = c["ref64"], c["ref32"]
        check_close(m_d, r64["m_d"].reshape(-1), r32["m_d"].reshape(-1), what="%s m_d" % name)
        check_close(s, r64["S"], r32["S"], what="%s S" % name)
        check_close(grads[0], r64["dmu"].reshape(b, d), r32["dmu"].reshape(b, d) if "dmu" in r32 else None, what="%s dmu" % name)
        check_close(grads[1], r64["dlogvar"].reshape(b, d), r32["dlogvar"].reshape(b, d) if "dlogvar" in r32 else None,
                    what="%s dlogvar" % name)
    assert n >= 2


def test_fused_regulariser_rejects_large_shapes(ops):
    assert not ops.fused_reg_supported(129, 64, DEV) and not ops.fused_reg_supported(64, 65, DEV)
    with pytest.raises(RuntimeError, match="outside the fused path"):
        ops.kl_mpd_reg_fused(torch.zeros(200, 8, device=DEV), torch.zeros(200, 8, device=DEV), 1.0, 1.0, 0.0)


@pytest.mark.parametrize("one_node", [True, False])
@pytest.mark.parametrize("defer", [False, True])
@pytest.mark.parametrize("kind", ["dmol", "bernoulli"])
@pytest.mark.parametrize("eta,gamma,free_bits", [(1.0, 1.0, 0.0), (0.0, 0.0, 0.0), (0.5, 2.0, 1e5)])
def test_fused_loss_heads_vs_oracle(ops, O, kind, eta, gamma, free_bits, defer, one_node, monkeypatch):
    """heads.loss_head_*: one_node = reparam + ONE node for everything else (regulariser kernel on the side stream);
    otherwise reparam + KL/MPD node (side stream) + likelihood/objective node.  Against the oracle's composition of the
    reference formulas, values and all gradients."""
    from mae_b200 import heads, synth
    monkeypatch.setattr(heads, "ONE_NODE", one_node)
    b, ns, d = (12, 2, 10) if kind == "dmol" else (12, 2, 70)     # D = 70: encode_head takes the general kernel chain
    mu0, lv0 = synth.posterior_params(b, (d,), seed=3, clamp=True)
    eps = synth.noise(b, ns, (d,), seed=3)
    mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
    mur, lvr = mu0.double().requires_grad_(), lv0.double().requires_grad_()
    if kind == "dmol":
        inp = synth.color_head(b, ns, 10, 8, 8, seed=4)
        dec = cu(inp["raw"]).requires_grad_()
        decr = inp["raw"].double().requires_grad_()
        out = heads.loss_head_color(mu, lv, cu(eps), dec, cu(inp["x"]), 10, eta, gamma, free_bits, defer_join=defer)
        ref, _ = O.loss_head_color(mur, lvr, eps.double(), lambda z: decr, inp["x"].double(), 10, eta, gamma, free_bits)
    else:
        inp = synth.binary_head(b, ns, 50, seed=4)
        dec = cu(inp["probs"]).requires_grad_()
        decr = inp["probs"].double().requires_grad_()
        out = heads.loss_head_binary(mu, lv, cu(eps), dec, cu(inp["x"]), eta, gamma, free_bits, defer_join=defer)
        ref, _ = O.loss_head_binary(mur, lvr, eps.double(), lambda z: decr, inp["x"].double(), eta, gamma, free_bits)
    loss, recon, kl, stats, z = out
    (loss * 0.9).backward()
    (ref[0] * 0.9).backward()
    torch.cuda.synchronize()
    check_close(loss, ref[0], what="loss")
    check_close(recon, ref[1], what="recon")
    check_close(kl, ref[2], what="KL")
    check_close(stats[0], ref[3], what="postKL_mean")
    check_close(dec.grad, decr.grad, what="d decoder output")
    zero = lambda t, like: torch.zeros_like(like) if t is None else t
    check_close(zero(mu.grad, mu), zero(mur.grad, mur), what="dmu", floor=1e-12)
    check_close(zero(lv.grad, lv), zero(lvr.grad, lvr), what="dlogvar", floor=1e-12)


def test_fused_dmol_forward_backward_full_size(ops, O):
    """32x32, nmix=10: loss_tail takes the one-pass forward+backward kernel (err and the scaled gradient from a single
    read of raw); d loss != 1 exercises the in-place rescale, d loss == 1 the no-op path."""
    from mae_b200 import heads, synth
    b, ns, d = 3, 2, 12
    mu0, lv0 = synth.posterior_params(b, (d,), seed=5, clamp=True)
    eps = synth.noise(b, ns, (d,), seed=5)
    inp = synth.color_head(b, ns, 10, 32, 32, seed=6)
    for scale in (1.0, -2.5):
        mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
        raw = cu(inp["raw"]).requires_grad_()
        loss, recon, kl, stats, z = heads.loss_head_color(mu, lv, cu(eps), raw, cu(inp["x"]), 10, 1.0, 1.0, 0.0)
        (loss * scale).backward()
        mur, lvr, rawr = mu0.double().requires_grad_(), lv0.double().requires_grad_(), inp["raw"].double().requires_grad_()
        ref, _ = O.loss_head_color(mur, lvr, eps.double(), lambda z_: rawr, inp["x"].double(), 10, 1.0, 1.0, 0.0)
        (ref[0] * scale).backward()
        check_close(loss, ref[0], what="loss")
        check_close(recon, ref[1], what="recon")
        check_close(raw.grad, rawr.grad, what="draw scale=%g" % scale)
        check_close(mu.grad, mur.grad, what="dmu")
    # no gradient requested for the decoder output -> plain forward kernel, same value
    loss2 = heads.loss_head_color(cu(mu0), cu(lv0), cu(eps), cu(inp["raw"]), cu(inp["x"]), 10, 1.0, 1.0, 0.0)[0]
    check_close(loss2, ref[0], what="loss (no grad)")


def test_fused_head_under_cuda_graph(ops, O):
    """the training step captured in a CUDA graph (side-stream fork/join included) replays to the same numbers"""
    from mae_b200 import heads, synth
    b, d = 64, 64
    mu0, lv0 = synth.posterior_params(b, (d,), seed=8, clamp=True)
    eps = cu(synth.noise(b, 1, (d,), seed=8))
    inp = synth.color_head(b, 1, 10, 32, 32, seed=9)
    mu, lv = cu(mu0).requires_grad_(), cu(lv0).requires_grad_()
    raw, x = cu(inp["raw"]).requires_grad_(), cu(inp["x"])
    one = torch.ones((), device=DEV)

    def step():
        mu.grad = lv.grad = raw.grad = None
        loss = heads.loss_head_color(mu, lv, eps, raw, x, 10, 1.0, 1.0, 0.0, defer_join=True)[0]
        loss.backward(gradient=one)
        return loss

    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for _ in range(3):
            eager = step()
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    ref_loss, ref_gmu, ref_graw = eager.detach().clone(), mu.grad.clone(), raw.grad.clone()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        cap = step()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    assert torch.equal(cap.detach(), ref_loss)
    assert torch.equal(mu.grad, ref_gmu) and torch.equal(raw.grad, ref_graw)


@pytest.mark.parametrize("one_node", [True, False])
@pytest.mark.parametrize("order", ["ab", "ba", "read_first"])
def test_two_heads_interleaved_with_deferred_join(ops, order, one_node, monkeypatch):
    """Two models on one device whose loss() calls interleave (mae and mae_shadow of experiments/binary_image.py:106-108,304)
    share the per-device side stream and its join bookkeeping (ops._pending_join / _grads_ready).  Every side-stream launch
    of a device is ordered on ONE stream, so a join on the latest event covers all earlier work, and a popped or foreign
    'gradients ready' event falls back to a full stream wait: interleaved deferred heads must reproduce, bit for bit, what
    each head gives alone."""
    from mae_b200 import heads, synth
    monkeypatch.setattr(heads, "ONE_NODE", one_node)

    def inputs(seed, b, d):
        mu0, lv0 = synth.posterior_params(b, (d,), seed=seed, clamp=True)
        inp = synth.color_head(b, 1, 10, 16, 16, seed=seed + 1)
        return [cu(mu0), cu(lv0), cu(synth.noise(b, 1, (d,), seed=seed)), cu(inp["raw"]), cu(inp["x"])]

    def fresh(t):
        return [t[0].clone().requires_grad_(), t[1].clone().requires_grad_(), t[2], t[3].clone().requires_grad_(), t[4]]

    def alone(t):
        a = fresh(t)
        out = heads.loss_head_color(*a, 10, 1.0, 1.0, 0.0, defer_join=False)
        out[0].backward()
        torch.cuda.synchronize()
        return [o.detach().clone() for o in (out[0], out[1], out[2], out[3][0])] + [a[0].grad, a[1].grad, a[3].grad]

    ta, tb = inputs(21, 48, 32), inputs(23, 200, 70)        # cluster-kernel shape and general-chain shape
    ref_a, ref_b = alone(ta), alone(tb)
    a, b = fresh(ta), fresh(tb)
    out_a = heads.loss_head_color(*a, 10, 1.0, 1.0, 0.0, defer_join=True)
    out_b = heads.loss_head_color(*b, 10, 1.0, 1.0, 0.0, defer_join=True)
    if order == "read_first":           # values read before any backward: the caller joins explicitly
        ops.join_side(DEV)
        early = [o.detach().clone() for o in (out_a[0], out_b[0])]
        torch.cuda.synchronize()
        assert torch.equal(early[0], ref_a[0]) and torch.equal(early[1], ref_b[0])
    for which in ("ab" if order == "read_first" else order):
        (out_a if which == "a" else out_b)[0].backward()
    ops.join_side(DEV)
    torch.cuda.synchronize()
    for out, t, ref, nm in ((out_a, a, ref_a, "A"), (out_b, b, ref_b, "B")):
        got = [out[0].detach(), out[1].detach(), out[2].detach(), out[3][0].detach(), t[0].grad, t[1].grad, t[3].grad]
        for g, r, what in zip(got, ref, ("loss", "recon", "kl", "postKL", "dmu", "dlogvar", "draw")):
            assert torch.equal(g, r), "head %s %s differs when interleaved (%s)" % (nm, what, order)


# --------------------------------------------------------------------------------- drop-in: whole MAE
def _build(case, name):
    import stubs
    from mae_b200.modules import MAE, GaussianEncoder
    if name.startswith("binary"):
        flow = stubs.AffineFlow(case["nz"]) if "priorflow" in name else None
        enc = GaussianEncoder(stubs.LinearCore(case["npix"], case["nz"], False), prior_flow=flow)
        dec = stubs.BinaryLinearDecoder(case["nz"], case["npix"])
    else:
        enc = GaussianEncoder(stubs.LinearCore(3 * case["h"] * case["w"], case["nz"], True))
        dec = stubs.ColorLinearDecoder(case["nz"], case["nmix"], case["h"], case["w"])
    mae = MAE(enc, dec)
    mae.load_state_dict(case["state"])
    return mae.to(DEV)


def test_mae_dropin_matches_reference_mae(golden, monkeypatch):
    """Same 7-tuple, parameter gradients and IW-NLL as the reference's MAE driven with identical stand-in conv
    stacks, weights, inputs and noise (tests/golden/make_golden.py::golden_mae)."""
    from mae_b200.modules import GaussianEncoder
    for name, c in golden("mae_stub").items():
        mae = _build(c, name)
        x = cu(c["x"])
        r64, r32 = c["ref64"], c["ref32"]
        monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns, e=c["eps_loss"]: cu(e)))
        out = mae.loss(x, c["ns"], eta=c["eta"], gamma=c["gamma"], free_bits=c["free_bits"])
        assert len(out) == 7
        out[0].backward()
        for i, key in enumerate(["loss", "recon", "KL", "postKL_mean", "postKL_std", "loss_mean", "loss_std"]):
            if not torch.is_tensor(out[i]):
                assert out[i] == 0.0 and float(r64["loss7"][i]) == 0.0, "%s %s" % (name, key)   # mae.py:134-135
                continue
            check_close(out[i], r64["loss7"][i], r32["loss7"][i], what="%s %s" % (name, key))
        for pname, p in mae.named_parameters():
            check_close(p.grad, r64["grads"][pname], r32["grads"][pname], what="%s grad %s" % (name, pname))
        monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns, e=c["eps_nll"]: cu(e)))
        (e, r, kl), iw = mae.nll(x, c["k"])
        nx = c["x"][0].numel()
        for got, key in ((e, "nll_elbo"), (r, "nll_recon"), (iw, "nll_iw")):
            check_close(got, r64[key], r32[key], what="%s %s" % (name, key))
            check_bpd(got, r64[key], nx, what="%s %s bits/dim" % (name, key))
        check_close(kl, r64["nll_kl"], r32["nll_kl"], what="%s nll kl" % name, floor=1e-4)


def test_sampling_api_shapes():
    import stubs
    from mae_b200.modules import MAE, GaussianEncoder
    torch.manual_seed(0)
    mae = MAE(GaussianEncoder(stubs.LinearCore(48, 6, True)), stubs.ColorLinearDecoder(6, 3, 4, 4)).to(DEV)
    x = torch.rand(5, 3, 4, 4, device=DEV) * 2 - 1
    z, (mu, lv, zn, ld) = mae.sample_from_posterior(x, 4)
    assert z.shape == (5, 4, 6) and ld.shape == (5, 4) and mu.shape == (5, 6)
    zp, _ = mae.sample_from_proir(7, device=torch.device(DEV))
    img, _ = mae.decode(zp, random_sample=True)
    assert img.shape == (7, 3, 4, 4) and float(img.abs().max()) <= 1.0
    rec, _ = mae.reconstruct(x, k=3)
    assert rec.shape == (5, 3, 4, 4)
    z2, kl, (m_d, s) = mae.encode(x, 2)
    assert z2.shape == (5, 2, 6) and kl.shape == (5,) and m_d.shape == (6,) and s.dim() == 0
