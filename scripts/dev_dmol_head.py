"""Development check of the fused 1x1-convolution + DMOL kernel (csrc/dmol_head.cu) against the unfused path
(fp64 convolution -> validated DMOL kernels), plus a timing at the IW-evaluation shape.  Run on a B200 via gpurun."""
import sys
import time

import torch

sys.path.insert(0, ".")
from mae_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False


def case(b, ns, c, nmix, hw_side, act, seed, grad=True):
    g = torch.Generator().manual_seed(seed)
    n = b * ns
    hid = (torch.randn(n, c, hw_side, hw_side, generator=g) * 1.5).to(dev)
    w = (torch.randn(10 * nmix, c, generator=g) * (1.0 / c ** 0.5)).to(dev)
    bias = (torch.randn(10 * nmix, generator=g) * 0.3).to(dev)
    x = (torch.randint(0, 256, (b, 3, hw_side, hw_side), generator=g).float() / 127.5 - 1.0).to(dev)
    x[:, :, 0, :8] = -1.0
    x[:, :, 1, :8] = 1.0
    # reference: fp64 convolution, then the validated DMOL op
    hid_r = hid.clone().requires_grad_(grad)
    w_r = w.clone().requires_grad_(grad)
    b_r = bias.clone().requires_grad_(grad)
    a = torch.nn.functional.elu(hid_r.double()) if act else hid_r.double()
    raw = (torch.einsum("oc,nchw->nohw", w_r.double(), a) + b_r.double().view(1, -1, 1, 1)).float()
    err_ref = ops.dmol_recon_error(raw, x, ns, nmix)
    hid_o = hid.clone().requires_grad_(grad)
    w_o = w.clone().requires_grad_(grad)
    b_o = bias.clone().requires_grad_(grad)
    err = ops.dmol_head_recon_error(hid_o, w_o, b_o, x, ns, nmix, act)
    torch.cuda.synchronize()
    sc = float(err_ref.abs().max())
    e = float((err - err_ref).abs().max()) / sc
    msg = "b=%d ns=%d C=%d nmix=%d hw=%d act=%d: err rel %.2e" % (b, ns, c, nmix, hw_side * hw_side, act, e)
    ok = e < 1e-5
    if grad:
        gw = torch.randn(b, ns, generator=g).to(dev)
        (err_ref * gw).sum().backward()
        (err * gw).sum().backward()
        torch.cuda.synchronize()
        for name, o, r in (("dh", hid_o.grad, hid_r.grad), ("dW", w_o.grad, w_r.grad), ("db", b_o.grad, b_r.grad)):
            ge = float((o - r).abs().max()) / float(r.abs().max())
            msg += " %s %.2e" % (name, ge)
            ok = ok and ge < 2e-5
    print(("OK   " if ok else "FAIL ") + msg, flush=True)
    return ok


def timing():
    b, ns, c, nmix, side = 8, 512, 96, 10, 32
    n = b * ns
    hid = torch.randn(n, c, side, side, device=dev)
    w = torch.randn(100, c, device=dev) / c ** 0.5
    bias = torch.randn(100, device=dev) * 0.3
    x = (torch.randint(0, 256, (b, 3, side, side), device=dev).float() / 127.5 - 1.0)
    conv = torch.nn.Conv2d(c, 100, 1).to(dev)
    with torch.no_grad():
        conv.weight.copy_(w.view(100, c, 1, 1))
        conv.bias.copy_(bias)

        def fused():
            return ops.dmol_head_recon_error(hid, w, bias, x, ns, nmix, True)

        def unfused():
            return ops.dmol_recon_error(conv(torch.nn.functional.elu(hid)), x, ns, nmix)

        for name, fn in (("fused", fused), ("unfused (cuDNN 1x1 conv of ELU(h) + DMOL kernel)", unfused)):
            for _ in range(3):
                out = fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5):
                out = fn()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 5
            px = n * side * side
            print("%s: %.3f ms per %d images x K=%d (%.1f Gpixel/s, %.0f GB/s of hidden)" % (name, ms, b, ns, px / ms / 1e6, px * c * 4 / ms / 1e6),
                  flush=True)
        d = float((fused() - unfused()).abs().max() / unfused().abs().max())
        print("fused vs unfused (fp32 cuDNN) at the evaluation shape: rel %.2e" % d)


if __name__ == "__main__":
    t0 = time.time()
    ok = True
    ok &= case(2, 1, 96, 10, 32, True, 1, grad=False)
    ok &= case(2, 1, 96, 10, 32, True, 1)
    ok &= case(3, 2, 96, 10, 32, False, 2)
    ok &= case(2, 2, 48, 10, 32, False, 3)
    ok &= case(1, 2, 64, 5, 64, True, 4)
    ok &= case(2, 1, 64, 10, 32, True, 5)
    ok &= case(40, 8, 96, 10, 32, True, 6, grad=False)     # more tiles than SMs: persistent loop, all pipeline phases
    print("cases done in %.1f s: %s" % (time.time() - t0, "ALL OK" if ok else "FAILURES"), flush=True)
    if "--time" in sys.argv:
        timing()
    sys.exit(0 if ok else 1)
