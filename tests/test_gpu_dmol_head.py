"""SURVEY.md 8f N2 on the GPU: the decoder's last 1x1 convolution fused with the DMOL likelihood (csrc/dmol_head.cu, tcgen05
3xTF32 with TMEM accumulators) against

* the LIVE reference's own last layer + likelihood (tests/golden/dmol_head.pt, made by tests/golden/make_golden_head.py:
  Conv2dWeightNorm -> ColorImageDecoder.reconstruct_error, weight_norm.py:46-86, pixelcnnpp.py:32-34, resnet.py:85,
  color_image_decoder.py:91-125) including the gradients with respect to hidden, weight_v, weight_g and bias;
* the fp64 oracle on seeded inputs at every supported (C, nmix) and with more tiles than SMs (persistent pipeline);
* the unfused kernels through the module interface (reconstruct_error of the shipped decoders in the three head modes)."""
import pytest
import torch

from tolerance import check_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name", ["pixelcnnpp32_c96_m10", "resnet32_c48_m10", "pixelcnnpp64_c64_m5"])
def test_fused_head_matches_reference_last_layer(golden, name):
    from mae_b200 import ops
    from mae_b200.modules.networks import Conv2dWeightNorm
    c = golden("dmol_head")[name]
    last = Conv2dWeightNorm(c["C"], 10 * c["nmix"], 1).to(DEV)
    with torch.no_grad():
        last.conv.weight_v.copy_(c["weight_v"])
        last.conv.weight_g.copy_(c["weight_g"])
        last.conv.bias.copy_(c["bias"])
    hidden = c["hidden"].to(DEV).requires_grad_()
    assert ops.dmol_head_supported(hidden, c["nmix"])
    err = ops.dmol_head_recon_error(hidden, last.conv.weight().flatten(1), last.conv.bias, c["x"].to(DEV), c["ns"], c["nmix"], c["act"])
    (err * c["g"].to(DEV)).sum().backward()
    r64, r32 = c["ref64"], c["ref32"]
    check_close(err, r64["err"], r32["err"], what="%s err" % name)
    for got, key in ((hidden.grad, "d_hidden"), (last.conv.weight_v.grad, "d_weight_v"), (last.conv.weight_g.grad, "d_weight_g"),
                     (last.conv.bias.grad, "d_bias")):
        check_close(got, r64[key], r32[key], what="%s %s" % (name, key))


@pytest.mark.parametrize("b,ns,c,nmix,side,act,grad", [
    (2, 1, 96, 10, 32, True, True), (3, 2, 96, 10, 32, False, True), (2, 2, 48, 10, 32, False, True),
    (1, 2, 64, 5, 64, True, True), (2, 1, 64, 10, 32, True, True),
    (40, 8, 96, 10, 32, True, False),      # 2560 tiles on 148 persistent CTAs: every pipeline phase wraps many times
    (1, 1, 96, 10, 32, False, False),      # 8 tiles: most CTAs idle
])
def test_fused_head_matches_oracle(b, ns, c, nmix, side, act, grad):
    from oracle import mae_oracle as O
    from mae_b200 import ops
    g = torch.Generator().manual_seed(1000 + b * 7 + c)
    n = b * ns
    hid = torch.randn(n, c, side, side, generator=g)
    w = torch.randn(10 * nmix, c, generator=g) * (0.5 / c ** 0.5)
    bias = torch.randn(10 * nmix, generator=g) * 0.1
    x = torch.randint(0, 256, (b, 3, side, side), generator=g).float() / 127.5 - 1.0
    x[:, :, 0, :8] = -1.0
    x[:, :, 1, :8] = 1.0
    gw = torch.randn(b, ns, generator=g)
    ref = {}
    for key, dt in (("r64", torch.float64), ("r32", torch.float32)):
        h_, w_, b_ = (t.to(DEV, dt).requires_grad_(grad) for t in (hid, w, bias))
        a = torch.nn.functional.elu(h_) if act else h_
        raw = torch.einsum("oc,nchw->nohw", w_, a) + b_.view(1, -1, 1, 1)
        e = O.dmol_recon_error(raw, x.to(DEV, dt), ns, nmix)
        if grad:
            (e * gw.to(DEV, dt)).sum().backward()
        ref[key] = (e.detach(), h_.grad, w_.grad, b_.grad)
    h_o, w_o, b_o = (t.to(DEV).requires_grad_(grad) for t in (hid, w, bias))
    err = ops.dmol_head_recon_error(h_o, w_o, b_o, x.to(DEV), ns, nmix, act)
    if grad:
        (err * gw.to(DEV)).sum().backward()
    ours = (err.detach(), h_o.grad, w_o.grad, b_o.grad)
    for i, nm in enumerate(("err", "d hidden", "d weight", "d bias")):
        if ours[i] is not None:
            check_close(ours[i], ref["r64"][i], ref["r32"][i], what=nm)
    # bitwise reproducible (ordered per-sample reduction)
    with torch.no_grad():
        again = ops.dmol_head_recon_error(h_o, w_o, b_o, x.to(DEV), ns, nmix, act)
    assert torch.equal(again, err.detach())


@pytest.mark.parametrize("decoder", ["resnet_color_32x32", "pixelcnn++_color_32x32", "pixelcnn++_color_64x64"])
def test_decoder_head_modes_agree(decoder):
    """reconstruct_error of the shipped colour decoders: raw tensor + DMOL kernels ("off") vs the fused head, forward
    (evaluation, "eval") and with parameter gradients ("always")."""
    from mae_b200 import ops, synth
    from mae_b200.modules import Decoder
    params = {"resnet_color_32x32": dict(z_channels=16, nmix=10),
              "pixelcnn++_color_32x32": dict(z_channels=16, h_channels=16, nmix=10, dropout=0.0, activation="concat_elu"),
              "pixelcnn++_color_64x64": dict(z_channels=32, h_channels=16, nmix=5, dropout=0.0, activation="elu")}[decoder]
    torch.manual_seed(3)
    dec = synth.fill_state(Decoder.by_name(decoder).from_params(dict(params))).to(DEV)
    side = dec.output_size()[1]
    b, ns = 2, 2
    x = (torch.randint(0, 256, (b, 3, side, side)).float() / 127.5 - 1.0).to(DEV)
    z = torch.randn(b, ns, *dec.z_shape(), device=DEV)
    assert dec.head_parts() is not None
    res = {}
    prev = ops.dmol_head_mode()
    try:
        for mode in ("off", "always"):
            ops.set_dmol_head(mode)
            dec.zero_grad()
            n0 = ops.launch_count()
            err = dec.reconstruct_error(x, z)
            err.sum().backward()
            res[mode] = (err.detach().clone(), {n: p.grad.detach().clone() for n, p in dec.named_parameters() if p.grad is not None},
                         ops.launch_count() - n0)
        ops.set_dmol_head("eval")
        with torch.no_grad():
            res["eval"] = dec.reconstruct_error(x, z)
            dec.zero_grad()
        err_train = dec.reconstruct_error(x, z)          # gradients needed: "eval" mode keeps the raw-tensor kernels
        assert type(err_train.grad_fn).__name__ == "_DmolBackward"
    finally:
        ops.set_dmol_head(prev)
    check_close(res["always"][0], res["off"][0].double(), rel=1e-5, what="%s err fused vs unfused" % decoder)
    check_close(res["eval"], res["off"][0].double(), rel=1e-5, what="%s err fused (no_grad) vs unfused" % decoder)
    assert res["always"][2] == 1 and res["off"][2] == 2        # one launch of ours instead of forward + backward kernels
    assert set(res["always"][1]) == set(res["off"][1])
    gscale = max(float(g.abs().max()) for g in res["off"][1].values())
    for n, gref in res["off"][1].items():
        check_close(res["always"][1][n], gref.double(), rel=1e-4, floor=1e-5 * gscale, what="%s grad %s" % (decoder, n))
