"""the PyTorch model shell (conv stacks, masked layers, flows) against the LIVE reference's models.

Fixtures: tests/golden/models.pt (tests/golden/make_golden_models.py) = ``MAE.from_params`` of the reference on its
shipped binary_image.json and on a ResNet colour config, parameters filled from their names (synth.fill_state).  These
tests run on the CPU: they cover construction from the reference's JSON, state_dict key/shape/mask identity (so that
reference checkpoints load strictly), the networks' forward passes, the AF-MADE flow in both directions and the
data-dependent initialisation.  The loss / NLL through the CUDA hot path is tests/test_gpu_models.py."""
import copy

import pytest
import torch

from tolerance import check_close

from mae_b200 import synth
from mae_b200.modules import MAE


@pytest.fixture(scope="module")
def fixtures(golden):
    return golden("models")


def _ours(c):
    return synth.fill_state(MAE.from_params(copy.deepcopy(c["params"])))


ALL = ["binary_image_json", "color_resnet", "color_image32x32_json"]


@pytest.mark.parametrize("name", ALL)
def test_state_dict_identity_with_reference(fixtures, name):
    c = fixtures[name]
    model = MAE.from_params(copy.deepcopy(c["params"]))
    sd = model.state_dict()
    assert {n: tuple(t.shape) for n, t in sd.items()} == c["keys"]          # strict loading of reference checkpoints
    for n, checksum in c["masks"].items():
        t = sd[n]
        got = float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).view(t.shape) % 977).sum())
        assert got == checksum, "mask %s differs from the reference's" % n
    # the reference's own state dict round-trips
    model.load_state_dict({n: torch.zeros(s) for n, s in c["keys"].items()}, strict=True)


@pytest.mark.parametrize("cfg", ["binary_image.json", "color_image32x32.json", "color_image64x64.json"])
def test_shipped_json_configs_build_with_reference_keys(fixtures, cfg):
    """experiments/configs/*.json as written (the colour ones ask for 2 replicas -> DataParallel 'module.' key level)"""
    c = fixtures["literal_configs"][cfg]
    if torch.cuda.is_available() and torch.cuda.device_count() < 2 and "color" in cfg:
        pytest.skip("ngpu=2 configuration needs two devices when CUDA is present")
    model = MAE.from_params(copy.deepcopy(c["params"]))
    assert {n: tuple(t.shape) for n, t in model.state_dict().items()} == c["keys"]


@pytest.mark.parametrize("name", ALL)
def test_encoder_core_and_flow_forward(fixtures, name):
    c = fixtures[name]
    r64, r32 = c["ref64"], c["ref32"]
    model = _ours(c).eval()
    with torch.no_grad():
        mu, logvar = model.encoder.core(c["x"])
        check_close(mu, r64["mu"], r32["mu"], what="%s mu" % name)
        check_close(logvar, r64["logvar"], r32["logvar"], what="%s logvar" % name)
        flow = model.encoder.prior_flow
        if "flow_bwd" in r64:
            zn, ld = flow.backward(mu.view(mu.size(0), *flow.input_size()))
            check_close(zn, r64["flow_bwd"], r32["flow_bwd"], what="flow backward")
            check_close(ld, r64["flow_bwd_logdet"], r32["flow_bwd_logdet"], what="flow backward logdet")
            zf, ldf = flow.forward(zn[:2])
            check_close(zf, r64["flow_fwd"], r32["flow_fwd"], what="flow forward (sequential inverse)")
            check_close(ldf, r64["flow_fwd_logdet"], r32["flow_fwd_logdet"], what="flow forward logdet")
            check_close(zf, mu[:2], rel=1e-4, what="forward(backward(z)) == z")
        else:
            assert flow is None


def _decoder_net(dec, x, z):
    """the decoder network's output for a given z: raw mixture parameters (colour) or Bernoulli probabilities (binary)"""
    zdec = z.view(z.size(0), *dec.z_shape())
    if hasattr(dec, "nmix"):
        return dec(x, zdec)
    if hasattr(dec, "z_transform"):
        return dec.core(torch.cat([x, dec.z_transform(zdec)], dim=1))
    return dec(zdec)


@pytest.mark.parametrize("name", ALL)
def test_decoder_network_output(fixtures, name):
    """ResNet / PixelCNN / PixelCNN++ decoders: the tensor the likelihood kernel reads, against the live reference"""
    c = fixtures[name]
    r64, r32 = c["ref64"], c["ref32"]
    model = _ours(c).eval()
    with torch.no_grad():
        net = _decoder_net(model.decoder, c["x"], r32["mu"])
    assert tuple(net.shape) == tuple(r32["dec_out_shape"])
    check_close(net[:2, ::7, ::3, ::3], r64["dec_out_slice"], r32["dec_out_slice"], rel=2e-5, what="%s decoder output" % name)


def test_color64_networks_forward(fixtures):
    """color_image64x64.json (DenseNet 64x64 core, AF2d-MADE, PixelCNN++ with 4 levels): cores, flow and decoder forward"""
    c = fixtures["color_image64x64_networks"]
    r64, r32 = c["ref64"], c["ref32"]
    model = _ours(c).eval()
    with torch.no_grad():
        mu, logvar = model.encoder.core(c["x"])
        check_close(mu, r64["mu"], r32["mu"], what="mu 64x64")
        check_close(logvar, r64["logvar"], r32["logvar"], what="logvar 64x64")
        zn, ld = model.encoder.prior_flow.backward(r32["mu"])
        check_close(zn, r64["flow_bwd"], r32["flow_bwd"], what="flow backward 64x64")
        check_close(ld, r64["flow_bwd_logdet"], r32["flow_bwd_logdet"], what="flow logdet 64x64")
        net = model.decoder(c["x"], r32["mu"])
    assert tuple(net.shape) == tuple(r32["dec_out_shape"])
    check_close(net[:2, ::7, ::3, ::3], r64["dec_out_slice"], r32["dec_out_slice"], rel=2e-5, what="decoder output 64x64")


@pytest.mark.parametrize("name", ALL)
def test_data_dependent_initialize(fixtures, name):
    """MAE.initialize (weight_norm.py:25-37, masked.py:84-98 through every stack) from identical filled weights ends in
    the same gains and biases."""
    c = fixtures[name]
    model = _ours(c)
    with torch.no_grad():
        model.initialize(c["x_init"], init_scale=1.0)
    sd = model.state_dict()
    for n, ref in c["init"].items():
        check_close(sd[n].reshape(-1)[:32], ref, rel=2e-4, what="%s after initialize" % n)
    for n, ref in c["init_bias"].items():
        check_close(sd[n].reshape(-1)[:32], ref, rel=2e-4, floor=2e-5, what="%s after initialize" % n)


def test_masked_layers_keep_masked_entries_zero():
    """functional masking: masked entries are zero from construction on and get a zero gradient (the reference re-zeroes
    ``weight_v.data`` on every call, masked.py:102)."""
    from mae_b200.modules.networks import MADE, MaskedConv2d
    torch.manual_seed(0)
    made = MADE(6, 2, 18, 'B')
    conv = MaskedConv2d(3, 4, 5, mask_type='A', masked_channels=1, padding=2)
    made(torch.randn(4, 6)).sum().backward()
    conv(torch.randn(2, 3, 7, 7)).sum().backward()
    for layer in [made.input_layer, made.output_layer, made.direct_connect, made.hidden_layers[0], conv]:
        assert float((layer.weight_v * (1 - layer.mask)).abs().max()) == 0.0
        assert float((layer.weight_v.grad * (1 - layer.mask)).abs().max()) == 0.0
    # autoregressive property of MADE order B: output i depends only on inputs > i
    x = torch.randn(1, 6, requires_grad=True)
    jac = torch.autograd.functional.jacobian(lambda t: made(t), x)[0, :, 0, :]
    assert float(jac.tril().abs().max()) == 0.0


def test_registries_hold_the_reference_keys():
    """every registry key the reference's three shipped JSON configurations (and its ResNet variants) use"""
    from mae_b200.modules import Decoder, Encoder, EncoderCore, Flow
    assert Encoder.by_name("gaussian") is not None
    for key in ("resnet_binary_28x28", "resnet_color_32x32", "densenet_color_32x32", "densenet_color_64x64"):
        assert EncoderCore.by_name(key) is not None
    for key in ("resnet_binary_28x28", "resnet_color_32x32", "pixelcnn_binary_28x28", "pixelcnn++_color_32x32",
                "pixelcnn++_color_64x64"):
        assert Decoder.by_name(key) is not None
    assert Flow.by_name("af_made") is not None and Flow.by_name("af2d_made") is not None


def test_pixelcnnpp_is_causal():
    """raw mixture parameters at pixel (i, j) must not depend on x at (i, j) or later in raster order"""
    from mae_b200.modules import PixelCNNPPDecoderColorImage32x32
    torch.manual_seed(0)
    dec = PixelCNNPPDecoderColorImage32x32(4, 4, 2, dropout=0.0, activation='concat_elu').eval()
    z = torch.randn(1, 4, 8, 8)
    x = torch.rand(1, 3, 32, 32) * 2 - 1
    with torch.no_grad():
        base = dec(x, z)
        x2 = x.clone()
        x2[:, :, 13, 17:] += 0.5
        x2[:, :, 14:, :] -= 0.5
        moved = dec(x2, z)
    diff = (moved - base).abs().amax(dim=(0, 1))
    assert float(diff[:13].max()) == 0.0 and float(diff[13, :18].max()) == 0.0     # rows above, and up to (13, 17) itself
    assert float(diff[13, 18:].max()) > 0.0 and float(diff[14:].max()) > 0.0


def test_cached_weights_give_identical_flow_samples():
    """N4: memoised effective weights inside the sequential direction (no-grad) == recomputing them on every call; with
    autograd on the cache is bypassed"""
    from mae_b200.modules import AFMADE, networks
    torch.manual_seed(0)
    flow = AFMADE(6, 2)
    z = torch.randn(3, 6)
    with torch.no_grad():
        y1, ld1 = flow.forward(z)                      # uses cached_weights() internally
        calls = {"n": 0}
        orig = networks._memo

        def no_cache(layer, compute):
            calls["n"] += 1
            return compute()

        networks._memo = no_cache
        try:
            y2, ld2 = flow.forward(z)
        finally:
            networks._memo = orig
    assert calls["n"] > 0 and torch.equal(y1, y2) and torch.equal(ld1, ld2)
    with networks.cached_weights():
        w = flow.blocks[0].fwd.mu.input_layer
        a = w.masked_weight()
        assert w.masked_weight() is not a               # autograd on: never cached
        with torch.no_grad():
            b = w.masked_weight()
            assert w.masked_weight() is b


def test_utils_helpers_match_reference_semantics():
    """norm / logsumexp / logavgexp / deterministic mixture sampling (mae/modules/utils.py) on small inputs"""
    import math
    from mae_b200.modules import logavgexp, logsumexp, norm, sample_from_discretized_mix_logistic
    torch.manual_seed(0)
    p = torch.randn(5, 3, 2, 2)
    assert torch.allclose(norm(p, 0), p.reshape(5, -1).norm(dim=1).view(5, 1, 1, 1))
    assert torch.allclose(norm(p, 3), p.reshape(-1, 2).norm(dim=0).view(1, 1, 1, 2))
    assert torch.allclose(norm(p, None), p.norm())
    x = torch.randn(4, 7)
    assert torch.allclose(logsumexp(x, 1), torch.log(torch.exp(x).sum(1)))
    assert torch.allclose(logavgexp(x, 1), torch.log(torch.exp(x).mean(1)))
    assert torch.allclose(logavgexp(x), torch.log(torch.exp(x).mean()))
    means, ls, co = torch.randn(2, 3, 3, 4, 4), torch.randn(2, 3, 3, 4, 4), torch.tanh(torch.randn(2, 3, 3, 4, 4))
    logits = torch.log_softmax(torch.randn(2, 3, 4, 4), dim=1)
    out = sample_from_discretized_mix_logistic(means, ls, co, logits, False)
    k = logits.argmax(dim=1)
    sel = lambda t: t.gather(1, k[:, None, None].expand(-1, 1, 3, -1, -1)).squeeze(1)
    m, c = sel(means), sel(co)
    x0 = m[:, 0].clamp(-1, 1)
    x1 = (m[:, 1] + c[:, 0] * x0).clamp(-1, 1)
    x2 = (m[:, 2] + c[:, 1] * x0 + c[:, 2] * x1).clamp(-1, 1)
    assert torch.equal(out, torch.stack([x0, x1, x2], 1)) and out.shape == (2, 3, 4, 4)


def test_mae_load_round_trip(tmp_path, fixtures):
    """MAE.load(model_path, device): config.json + model.pt as the reference's training scripts write them"""
    import json
    c = fixtures["binary_image_json"]
    model = _ours(c)
    json.dump(c["params"], open(tmp_path / "config.json", "w"))
    torch.save(model.state_dict(), tmp_path / "model.pt")
    loaded = MAE.load(str(tmp_path), torch.device("cpu"))
    for (n1, t1), (n2, t2) in zip(model.state_dict().items(), loaded.state_dict().items()):
        assert n1 == n2 and torch.equal(t1, t2)
