"""SURVEY.md §8f N1 on the GPU: whole models built by ``MAE.from_params`` from the reference's JSON, on the CUDA hot path,
against the LIVE reference's loss 7-tuple, parameter gradients and importance-weighted NLL (tests/golden/models.pt, made by
tests/golden/make_golden_models.py with name-keyed weights and injected noise).  binary_image_json = the shipped
binary_image.json (ResNet core, AF-MADE prior flow -> Monte-Carlo KL branch, PixelCNN decoder -> Bernoulli kernel);
color_resnet = ResNet core/decoder without flows (analytic-KL branch: encode_head/loss_head nodes at D = 128 fall back to
the general MPD kernels, DMOL kernel on the raw conv output); color_image32x32_json = the shipped color_image32x32.json
(DenseNet core, AF2d-MADE prior flow, PixelCNN++ decoder, D = 1024) with one replica and dropout 0."""
import copy

import pytest
import torch

from tolerance import check_bpd, check_close

pytestmark = pytest.mark.gpu
DEV = "cuda"


def cu(t):
    return t.to(DEV)


@pytest.fixture(scope="module")
def fixtures(golden):
    return golden("models")


@pytest.fixture(autouse=True)
def _no_tf32():
    old = torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("name,head_mode", [
    ("binary_image_json", "eval"), ("color_resnet", "eval"), ("color_image32x32_json", "eval"),
    # N2: "always" also trains through the fused last-convolution + likelihood kernel (loss 7-tuple and every parameter
    # gradient against the live reference); "off" keeps the raw tensor in evaluation as well
    ("color_resnet", "always"), ("color_image32x32_json", "always"), ("color_resnet", "off"), ("color_image32x32_json", "off")])
def test_full_model_loss_grads_and_nll_match_reference(fixtures, name, head_mode, monkeypatch):
    from mae_b200 import ops, synth
    from mae_b200.modules import MAE, GaussianEncoder
    monkeypatch.setattr(ops, "_dmol_head_mode", head_mode)
    c = fixtures[name]
    r64, r32 = c["ref64"], c["ref32"]
    model = synth.fill_state(MAE.from_params(copy.deepcopy(c["params"]))).to(DEV)
    x = cu(c["x"])
    model.train()
    monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns, e=c["eps_loss"]: cu(e)))
    out = model.loss(x, c["ns"], eta=c["eta"], gamma=c["gamma"], free_bits=c["free_bits"])
    assert len(out) == 7
    out[0].backward()
    for i, key in enumerate(["loss", "recon", "KL", "postKL_mean", "postKL_std", "loss_mean", "loss_std"]):
        check_close(out[i], r64["loss7"][i], r32["loss7"][i], rel=2e-5, what="%s %s" % (name, key))
    params = dict(model.named_parameters())
    buffers = dict(model.named_buffers())
    for pname, ref in r64["grads"].items():
        g = params[pname].grad.reshape(-1)[:64]
        ref32 = r32["grads"][pname]
        mask_name = pname[:pname.rfind(".") + 1] + "mask"
        if pname.endswith("weight_v") and mask_name in buffers:
            # masked entries: the reference lets a (meaningless, re-zeroed) gradient through, ours is exactly zero
            m = buffers[mask_name].reshape(-1)[:64].cpu().double()
            assert float((g.cpu().double() * (1 - m)).abs().max()) == 0.0
            ref, ref32 = ref * m, ref32 * m
        check_close(g, ref, ref32, rel=5e-5, what="%s grad %s" % (name, pname))
    model.eval()
    monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns, e=c["eps_nll"]: cu(e)))
    (e, r, kl), iw = model.nll(x, c["k"])
    nx = c["x"][0].numel()
    for got, key in ((e, "nll_elbo"), (r, "nll_recon"), (iw, "nll_iw")):
        check_close(got, r64[key], r32[key], rel=2e-5, what="%s %s" % (name, key))
        check_bpd(got, r64[key], nx, what="%s %s bits/dim" % (name, key))
    check_close(kl, r64["nll_kl"], r32["nll_kl"], rel=2e-5, what="%s nll kl" % name, floor=1e-4)


def test_pixelcnn_decoder_sampling_runs(fixtures):
    """ancestral sampling API of the autoregressive decoder (shapes / value ranges only: 784 network evaluations)"""
    from mae_b200.modules import PixelCNNDecoderBinaryImage28x28
    torch.manual_seed(0)
    dec = PixelCNNDecoderBinaryImage28x28(4, 'small').to(DEV).eval()
    with torch.no_grad():
        img, probs = dec.decode(torch.randn(2, 4, device=DEV), random_sample=True)
    assert img.shape == (2, 1, 28, 28) and probs.shape == (2, 1, 28, 28)
    assert set(img.unique().tolist()) <= {0.0, 1.0} and 0.0 <= float(probs.min()) and float(probs.max()) <= 1.0


def test_batched_calc_nll_equals_per_image_loop(monkeypatch):
    """evaluate.calc_nll (several images x K per call, ragged last batch) == the reference's loop of one image per call
    (experiments/color_image.py:260) when both consume the same noise"""
    import stubs
    from mae_b200 import evaluate
    from mae_b200.modules import MAE, GaussianEncoder
    torch.manual_seed(0)
    model = MAE(GaussianEncoder(stubs.LinearCore(48, 6, True)), stubs.ColorLinearDecoder(6, 3, 4, 4)).to(DEV)
    data = torch.rand(7, 3, 4, 4) * 2 - 1
    k = 16
    noise = torch.randn(7, k, 6, generator=torch.Generator().manual_seed(1))
    cursor = {"i": 0}

    def draw(mu, ns):
        b = mu.size(0)
        e = noise[cursor["i"]:cursor["i"] + b]
        cursor["i"] += b
        return e.to(mu.device)

    monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(draw))
    (e1, r1, k1), iw1, bpd1 = evaluate.calc_nll(model, data, k, images_per_step=1)
    cursor["i"] = 0
    (e3, r3, k3), iw3, bpd3 = evaluate.calc_nll(model, data, k, images_per_step=3)
    for a, b in ((e1, e3), (r1, r3), (k1, k3), (iw1, iw3), (bpd1, bpd3)):
        assert abs(a - b) <= 1e-5 * max(1.0, abs(a)), (a, b)
    assert model.training


def test_calc_nll_cuda_graph_equals_eager(monkeypatch):
    """evaluate.calc_nll(cuda_graph=True): the captured per-batch body (decoder, log densities, fused IW kernel, running
    sums) replayed over the full batches + the eager ragged tail == the eager loop.  Noise is pinned to zero so that both
    see the same samples (a replay draws fresh Philox noise otherwise)."""
    import stubs
    from mae_b200 import evaluate
    from mae_b200.modules import MAE, GaussianEncoder
    torch.manual_seed(0)
    model = MAE(GaussianEncoder(stubs.LinearCore(48, 6, True)), stubs.ColorLinearDecoder(6, 3, 4, 4)).to(DEV)
    data = torch.rand(26, 3, 4, 4) * 2 - 1
    k = 16
    monkeypatch.setattr(GaussianEncoder, "_draw_eps", staticmethod(lambda mu, ns: mu.new_zeros(mu.size(0), ns, *mu.shape[1:])))
    ref = evaluate.calc_nll(model, data, k, images_per_step=4)
    got = evaluate.calc_nll(model, data, k, images_per_step=4, cuda_graph=True)
    flat = lambda r: (r[0][0], r[0][1], r[0][2], r[1], r[2])
    for a, b in zip(flat(got), flat(ref)):
        assert abs(a - b) <= 1e-6 * max(1.0, abs(b)), (a, b)
