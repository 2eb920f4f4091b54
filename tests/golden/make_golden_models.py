"""Golden outputs of the LIVE reference's full models: ``MAE.from_params`` on the shipped
``experiments/configs/binary_image.json`` (ResNet core + AF-MADE prior flow + PixelCNN decoder) and on a ResNet colour
configuration, with parameters filled deterministically from their NAMES (``mae_b200.synth.fill_state``), so that the
fixture holds only inputs and outputs, no weights.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden_models.py
"""
from __future__ import annotations

import copy
import json
import os
import sys
import tempfile
import warnings

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
warnings.filterwarnings("ignore")

from mae.modules import MAE as RefMAE  # noqa: E402

from mae_b200 import synth  # noqa: E402

torch.set_num_threads(8)

COLOR_RESNET = {
    "encoder": {"type": "gaussian", "core": {"type": "resnet_color_32x32", "z_channels": 2}, "ngpu": 1},
    "decoder": {"type": "resnet_color_32x32", "z_channels": 2, "nmix": 10, "ngpu": 1},
}


class _InjectNoise:
    """Feeds fixed noise to the reference's ``new_empty(...).normal_()`` draw (gaussian_encoder.py:61)."""

    def __init__(self, eps):
        self.eps = eps

    def __enter__(self):
        self._orig = torch.Tensor.normal_
        eps = self.eps

        def fake(t, *a, **k):
            assert t.shape == eps.shape, (t.shape, eps.shape)
            return t.copy_(eps.to(t.dtype))

        torch.Tensor.normal_ = fake
        return self

    def __exit__(self, *exc):
        torch.Tensor.normal_ = self._orig


PROBE_GRADS = 6   # gradients of this many parameters (spread over the model) are stored, the first 64 entries each


def run_model(params, x, x_init, ns, k, eta, gamma, free_bits):
    out = {"params": copy.deepcopy(params), "x": x, "x_init": x_init, "ns": ns, "k": k, "eta": eta, "gamma": gamma,
           "free_bits": free_bits}
    model = RefMAE.from_params(copy.deepcopy(params))
    out["keys"] = {n: tuple(t.shape) for n, t in model.state_dict().items()}
    out["masks"] = {n: float((t * torch.arange(1, t.numel() + 1, dtype=torch.float64).view(t.shape) % 977).sum())
                    for n, t in model.state_dict().items() if n.endswith("mask")}
    synth.fill_state(model)
    z_shape = model.encoder.z_shape()
    b = x.size(0)
    torch.manual_seed(7)
    eps_loss = torch.empty(b, ns, *z_shape).normal_()
    eps_nll = torch.empty(b, k, *z_shape).normal_()
    out["eps_loss"], out["eps_nll"] = eps_loss, eps_nll
    names = [n for n, _ in model.named_parameters()]
    probe = [names[i * (len(names) - 1) // (PROBE_GRADS - 1)] for i in range(PROBE_GRADS)]
    for tag in ("ref32", "ref64"):
        m = synth.fill_state(RefMAE.from_params(copy.deepcopy(params)))   # (weight_norm hooks defeat deepcopy)
        m = m.double() if tag == "ref64" else m
        xx = x.double() if tag == "ref64" else x
        m.eval()
        res = {}
        with torch.no_grad():
            mu, logvar = m.encoder.core(xx)
            res["mu"], res["logvar"] = mu.clone(), logvar.clone()
            if m.encoder.prior_flow is not None:
                zn, ld = m.encoder.prior_flow.backward(mu.view(b, *m.encoder.prior_flow.input_size()))
                res["flow_bwd"], res["flow_bwd_logdet"] = zn, ld
                zf, ldf = m.encoder.prior_flow.forward(zn[:2])
                res["flow_fwd"], res["flow_fwd_logdet"] = zf, ldf
        with torch.no_grad():
            # decoder network output for z = mu (what the likelihood kernel reads): raw mixture parameters [B,10*nmix,H,W]
            # (colour) or Bernoulli probabilities (binary); a strided slice keeps the fixture small
            dec = m.decoder
            zdec = mu.view(b, *dec.z_shape())
            if hasattr(dec, "nmix"):
                net = dec(xx, zdec)
            elif hasattr(dec, "z_transform"):      # PixelCNN binary decoder: conditioning maps + image through the core
                net = dec.core(torch.cat([xx, dec.z_transform(zdec)], dim=1))
            else:
                net = dec(zdec)
            res["dec_out_shape"] = tuple(net.shape)
            res["dec_out_slice"] = net[:2, ::7, ::3, ::3].clone()
        m.train()
        with _InjectNoise(eps_loss):
            loss7 = m.loss(xx, ns, eta=eta, gamma=gamma, free_bits=free_bits)
        loss7[0].backward()
        grads = dict(m.named_parameters())
        res["loss7"] = [o.detach().clone() if torch.is_tensor(o) else torch.tensor(o) for o in loss7]
        res["grads"] = {n: grads[n].grad.reshape(-1)[:64].clone() for n in probe}
        res["grad_norms"] = {n: grads[n].grad.norm().clone() for n in probe}
        m.eval()
        with _InjectNoise(eps_nll), torch.no_grad():
            (nll_elbo, recon, kl), nll_iw = m.nll(xx, k)
        res.update(nll_elbo=nll_elbo, nll_recon=recon, nll_kl=kl, nll_iw=nll_iw)
        out[tag] = res
    # data-dependent initialisation from the filled weights (fp32 only): resulting gains of a few layers + output
    m = synth.fill_state(RefMAE.from_params(copy.deepcopy(params)))
    with torch.no_grad():
        m.initialize(x_init, init_scale=1.0)
    sd = m.state_dict()
    gains = [n for n in sd if n.endswith("weight_g")]
    pick = [gains[i * (len(gains) - 1) // 7] for i in range(8)]
    out["init"] = {n: sd[n].reshape(-1)[:32].clone() for n in pick}
    out["init_bias"] = {n.replace("weight_g", "bias"): sd[n.replace("weight_g", "bias")].reshape(-1)[:32].clone() for n in pick
                        if n.replace("weight_g", "bias") in sd}
    return out


def main():
    torch.manual_seed(20260)
    binary = json.load(open("/root/reference/experiments/configs/binary_image.json"))
    xb = (torch.rand(5, 1, 28, 28) < 0.3).float()
    xb_init = (torch.rand(16, 1, 28, 28) < 0.3).float()
    xc = torch.randint(0, 256, (4, 3, 32, 32)).float() / 127.5 - 1.0
    xc_init = torch.randint(0, 256, (8, 3, 32, 32)).float() / 127.5 - 1.0
    # the shipped CIFAR configuration (DenseNet core, AF2d-MADE prior flow, PixelCNN++ decoder); only the replica count
    # (2 GPUs -> 1: DataParallel adds a "module." key level and needs 2 devices) and the dropout rate (0.5 -> 0: the
    # training-mode loss must be deterministic) are changed
    color = json.load(open("/root/reference/experiments/configs/color_image32x32.json"))
    color["encoder"]["ngpu"] = color["decoder"]["ngpu"] = 1
    color["decoder"]["dropout"] = 0.0
    fixtures = {
        "binary_image_json": run_model(binary, xb, xb_init, ns=2, k=4, eta=1.0, gamma=0.5, free_bits=0.0),
        "color_resnet": run_model(COLOR_RESNET, xc, xc_init, ns=2, k=3, eta=0.5, gamma=1.0, free_bits=0.0),
        "color_image32x32_json": run_model(color, xc[:3], xc_init[:4], ns=1, k=2, eta=1.0, gamma=1.0, free_bits=0.0),
    }
    # the shipped 64x64 configuration: networks only (one replica, dropout 0), forward of core / flow / decoder in fp32+fp64
    c64 = json.load(open("/root/reference/experiments/configs/color_image64x64.json"))
    c64["encoder"]["ngpu"] = c64["decoder"]["ngpu"] = 1
    c64["decoder"]["dropout"] = 0.0
    x64 = torch.randint(0, 256, (2, 3, 64, 64)).float() / 127.5 - 1.0
    f64 = {"params": copy.deepcopy(c64), "x": x64}
    for tag in ("ref32", "ref64"):
        m = synth.fill_state(RefMAE.from_params(copy.deepcopy(c64)))
        m = (m.double() if tag == "ref64" else m).eval()
        xx = x64.double() if tag == "ref64" else x64
        with torch.no_grad():
            mu, logvar = m.encoder.core(xx)
            zn, ld = m.encoder.prior_flow.backward(mu)
            net = m.decoder(xx, mu)
        f64[tag] = {"mu": mu, "logvar": logvar, "flow_bwd": zn, "flow_bwd_logdet": ld, "dec_out_shape": tuple(net.shape),
                    "dec_out_slice": net[:2, ::7, ::3, ::3].clone()}
    fixtures["color_image64x64_networks"] = f64
    # state_dict keys / shapes of the three shipped configurations exactly as written (ngpu = 2 for the colour ones ->
    # DataParallel "module." level; pass-through on a CPU-only host)
    literal = {}
    for cfg in ("binary_image.json", "color_image32x32.json", "color_image64x64.json"):
        params = json.load(open("/root/reference/experiments/configs/" + cfg))
        model = RefMAE.from_params(copy.deepcopy(params))
        literal[cfg] = {"params": params, "keys": {n: tuple(t.shape) for n, t in model.state_dict().items()}}
    fixtures["literal_configs"] = literal
    path = os.path.join(HERE, "models.pt")
    torch.save(fixtures, path)
    print("wrote models.pt %.1f KB" % (os.path.getsize(path) / 1024.0))


if __name__ == "__main__":
    main()
