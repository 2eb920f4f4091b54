"""Golden fixture for the fused "last 1x1 convolution + DMOL likelihood" kernel, from the LIVE reference:

    python tests/golden/make_golden_head.py        # build container only (/root/reference is not on the GPU box)

The reference side is its own last layer followed by its own likelihood:  Conv2dWeightNorm(C, 10*nmix, 1)
(mae/modules/networks/weight_norm.py:46-86) applied to ELU(hidden) as in image_decoders/pixelcnnpp.py:32-34 (act) or to
hidden directly as in image_decoders/resnet.py:85, then ColorImageDecoder.reconstruct_error (color_image_decoder.py:91-125).
Stored per case: hidden, the layer's weight_v / weight_g / bias, x, the upstream gradient g, and the reference's err and
gradients in fp32 and (same code on .double() inputs) fp64."""
from __future__ import annotations

import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as MG  # noqa: E402  (sets up the reference import path and the overrides shim)

from mae.modules.networks.weight_norm import Conv2dWeightNorm as RefConv2dWeightNorm  # noqa: E402


class _RefHeadDecoder(MG.RefColorDecoder):
    """The reference's ColorImageDecoder with a conv stack that consists of its last layer only."""

    def __init__(self, c, nmix, act):
        super().__init__(nmix)
        self.act = act
        self.last = RefConv2dWeightNorm(c, 10 * nmix, 1, bias=True)
        self.hidden = None

    def forward(self, x, z):
        h = self.hidden
        return self.last(torch.nn.functional.elu(h) if self.act else h)


def run(case, dtype):
    c, nmix, act, b, ns = case["C"], case["nmix"], case["act"], case["x"].size(0), case["ns"]
    dec = _RefHeadDecoder(c, nmix, act).to(dtype)
    with torch.no_grad():
        dec.last.conv.weight_v.copy_(case["weight_v"].to(dtype))
        dec.last.conv.weight_g.copy_(case["weight_g"].to(dtype))
        dec.last.conv.bias.copy_(case["bias"].to(dtype))
    dec.hidden = case["hidden"].to(dtype).clone().requires_grad_()
    err = dec.reconstruct_error(case["x"].to(dtype), torch.zeros(b, ns, 1, dtype=dtype))
    (err * case["g"].to(dtype)).sum().backward()
    return {"err": err.detach(), "d_hidden": dec.hidden.grad, "d_weight_v": dec.last.conv.weight_v.grad,
            "d_weight_g": dec.last.conv.weight_g.grad, "d_bias": dec.last.conv.bias.grad}


def main():
    out = {}
    specs = {"pixelcnnpp32_c96_m10": dict(C=96, nmix=10, act=True, b=2, ns=2, h=16, w=8),
             "resnet32_c48_m10": dict(C=48, nmix=10, act=False, b=3, ns=1, h=16, w=8),
             "pixelcnnpp64_c64_m5": dict(C=64, nmix=5, act=True, b=2, ns=1, h=16, w=16)}
    for i, (name, sp) in enumerate(specs.items()):
        g = torch.Generator().manual_seed(700 + i)
        n = sp["b"] * sp["ns"]
        v = torch.randn(10 * sp["nmix"], sp["C"], 1, 1, generator=g) * 0.05
        case = dict(C=sp["C"], nmix=sp["nmix"], act=sp["act"], ns=sp["ns"],
                    hidden=torch.randn(n, sp["C"], sp["h"], sp["w"], generator=g) * 1.2,
                    weight_v=v, weight_g=(v.flatten(1).norm(dim=1) * (0.7 + 0.6 * torch.rand(v.size(0), generator=g))).view(-1, 1, 1, 1),
                    bias=torch.randn(10 * sp["nmix"], generator=g) * 0.2,
                    x=torch.randint(0, 256, (sp["b"], 3, sp["h"], sp["w"]), generator=g).float() / 127.5 - 1.0,
                    g=torch.randn(sp["b"], sp["ns"], generator=g))
        case["x"][:, :, 0, :4] = -1.0      # edge bins
        case["x"][:, :, 1, :4] = 1.0
        case["ref32"] = run(case, torch.float32)
        case["ref64"] = run(case, torch.float64)
        out[name] = case
    MG.save("dmol_head", out)


if __name__ == "__main__":
    main()
