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
    """ours vs the oracle (conv + DMOL) in fp64, with the oracle in fp32 as the measure of what fp32 can resolve"""
    from oracle import mae_oracle as O
    sys.path.insert(0, "tests")
    from tolerance import check_close
    g = torch.Generator().manual_seed(seed)
    n = b * ns
    hid = torch.randn(n, c, hw_side, hw_side, generator=g)
    w = torch.randn(10 * nmix, c, generator=g) * (0.5 / c ** 0.5)
    bias = torch.randn(10 * nmix, generator=g) * 0.1
    x = torch.randint(0, 256, (b, 3, hw_side, hw_side), generator=g).float() / 127.5 - 1.0
    x[:, :, 0, :8] = -1.0
    x[:, :, 1, :8] = 1.0
    gw = torch.randn(b, ns, generator=g)
    res = {}
    for name, dt, dv in (("r64", torch.float64, "cuda:0"), ("r32", torch.float32, "cuda:0")):
        h_ = hid.to(dv, dt).requires_grad_(grad)
        w_ = w.to(dv, dt).requires_grad_(grad)
        b_ = bias.to(dv, dt).requires_grad_(grad)
        a = torch.nn.functional.elu(h_) if act else h_
        raw = torch.einsum("oc,nchw->nohw", w_, a) + b_.view(1, -1, 1, 1)
        e = O.dmol_recon_error(raw, x.to(dv, dt), ns, nmix)
        if grad:
            (e * gw.to(dv, dt)).sum().backward()
        res[name] = (e.detach(), h_.grad, w_.grad, b_.grad)
    hid_o = hid.to(dev).requires_grad_(grad)
    w_o = w.to(dev).requires_grad_(grad)
    b_o = bias.to(dev).requires_grad_(grad)
    err = ops.dmol_head_recon_error(hid_o, w_o, b_o, x.to(dev), ns, nmix, act)
    if grad:
        (err * gw.to(dev)).sum().backward()
    torch.cuda.synchronize()
    msg = "b=%d ns=%d C=%d nmix=%d hw=%d act=%d:" % (b, ns, c, nmix, hw_side * hw_side, act)
    ok = True
    ours = (err.detach(), hid_o.grad, w_o.grad, b_o.grad)
    for i, nm in enumerate(("err", "dh", "dW", "db")):
        if ours[i] is None:
            continue
        try:
            e = check_close(ours[i], res["r64"][i], res["r32"][i], what=nm)
            e32 = float((res["r32"][i].double() - res["r64"][i]).abs().max() / res["r64"][i].abs().max())
            msg += " %s %.1e (fp32 oracle %.1e)" % (nm, e, e32)
        except AssertionError as ex:
            ok = False
            msg += " %s FAIL[%s]" % (nm, ex)
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


def role_profile():
    """MAE_HEAD_PROF=1: per-role cycle counters of CTA 0 (loop total, cycles blocked in its two barrier waits)."""
    from mae_b200 import _lib
    lib = _lib.load()
    b, ns, c, nmix, side = 8, 512, 96, 10, 32
    n = b * ns
    hid = torch.randn(n, c, side, side, device=dev)
    w = torch.randn(100, c, device=dev) / c ** 0.5
    bias = torch.randn(100, device=dev) * 0.3
    x = (torch.randint(0, 256, (b, 3, side, side), device=dev).float() / 127.5 - 1.0)
    err = torch.empty(n, device=dev)
    base = lib.mae_dmol_head_workspace_bytes(n, side * side)
    ws = torch.zeros(base + 256, dtype=torch.uint8, device=dev)
    for _ in range(3):
        _lib.check(lib.mae_dmol_head(hid.data_ptr(), w.data_ptr(), bias.data_ptr(), x.data_ptr(), b, ns, nmix, c, side * side, 1, 1.0,
                                     err.data_ptr(), None, ws.data_ptr(), ws.numel(), torch.cuda.current_stream().cuda_stream), "head")
    torch.cuda.synchronize()
    prof = ws[base:base + 256].view(torch.int64).cpu().tolist()
    tiles = (n * side * side // 128 + 147) // 148
    names = {0: "epilogue wg0 (wait tfull, -)", 4: "epilogue wg1", 8: "epilogue wg2", 12: "converter wg0 (wait sfull, wait aempty)",
             16: "converter wg1", 20: "loader (wait sempty, -)", 24: "mma issuer (wait tempty, wait afull)"}
    for k, nm in names.items():
        tot, w0, w1 = prof[k], prof[k + 1], prof[k + 2]
        print("%-44s total %9d cyc (%6.0f / tile)  wait0 %5.1f %%  wait1 %5.1f %%" % (nm, tot, tot / tiles, 100.0 * w0 / max(tot, 1), 100.0 * w1 / max(tot, 1)))


if __name__ == "__main__":
    if "--ncu" in sys.argv:       # three launches at the evaluation shape, nothing else (profiler target)
        b, ns, c = 8, 512, 96
        hid = torch.randn(b * ns, c, 32, 32, device=dev)
        w = torch.randn(100, c, device=dev) / c ** 0.5
        bias = torch.randn(100, device=dev) * 0.3
        x = (torch.randint(0, 256, (b, 3, 32, 32), device=dev).float() / 127.5 - 1.0)
        with torch.no_grad():
            for _ in range(3):
                ops.dmol_head_recon_error(hid, w, bias, x, ns, 10, True)
        torch.cuda.synchronize()
        print("ncu target done")
        sys.exit(0)
    if "--prof" in sys.argv:
        role_profile()
        sys.exit(0)
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
