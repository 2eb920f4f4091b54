"""Development check of the tensor-core MPD tiles (csrc/mpd_tc.cuh) against the FP32 FFMA tiles and the fp64 oracle."""
import sys
import time

import torch

sys.path.insert(0, ".")
from mae_b200 import _lib, ops, synth  # noqa: E402

dev = torch.device("cuda:0")
lib = _lib.load()


def run(b, d, mode, path=3, grad=True):
    lib.mae_set_mpd_tensor_cores(mode)
    mu, lv = synth.posterior_params(b, (d,), seed=11)
    mu, lv = mu.to(dev).requires_grad_(grad), lv.to(dev).requires_grad_(grad)
    m_d, s = ops.mpd(mu, lv, bwd_path=path)
    if grad:
        (-torch.nn.functional.logsigmoid(m_d).sum() + s).backward()
    torch.cuda.synchronize()
    return m_d.detach(), s.detach(), (mu.grad, lv.grad) if grad else None


def timeit(fn, n=5):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


if "--bias" in sys.argv:      # tensor-core vs FFMA gradients as a function of the batch (accumulation length)
    for b, d in ((4096, 64), (16384, 64), (32768, 64), (16384, 256)):
        _, s0, g0 = run(b, d, 0, 3)
        _, s1, g1 = run(b, d, 1, 3)
        eg = [float((a - c).abs().max() / c.abs().max()) for a, c in zip(g1, g0)]
        print("B=%d D=%d: S rel %.1e, dmu rel %.1e, dlv rel %.1e" % (b, d, abs(float(s1) - float(s0)) / float(s0), eg[0], eg[1]), flush=True)
    sys.exit(0)
ok = True
for b, d in ((4096, 64), (2500, 64), (3000, 32), (4096, 256)):
    for path in (3, 2):
        m0, s0, g0 = run(b, d, 0, path)
        m1, s1, g1 = run(b, d, 1, path)
        es = abs(float(s1) - float(s0)) / abs(float(s0))
        eg = max(float((a - c).abs().max() / c.abs().max()) for a, c in zip(g1, g0))
        good = es < 2e-6 and eg < 2e-5 and torch.equal(m0, m1)
        ok &= good
        print("%s B=%d D=%d path=%d: S %.7g vs %.7g (rel %.1e), grads rel %.1e" % ("OK  " if good else "FAIL", b, d, path, float(s1), float(s0), es, eg), flush=True)
if "--oracle" in sys.argv:
    from oracle import mae_oracle as O
    for b, d in ((2500, 64), (2304, 256)):
        mu, lv = synth.posterior_params(b, (d,), seed=11)
        m64, s64 = O.post_kl_chunked(mu.double(), lv.double())
        _, s1, _ = run(b, d, 1, grad=False)
        _, s0, _ = run(b, d, 0, grad=False)
        print("B=%d D=%d oracle fp64 S %.9g: tensor-core rel %.1e, FFMA rel %.1e" % (b, d, float(s64), abs(float(s1) - float(s64)) / float(s64), abs(float(s0) - float(s64)) / float(s64)), flush=True)
for b, d in ((8192, 64), (16384, 64), (16384, 256)):
    mu, lv = synth.posterior_params(b, (d,), seed=5)
    mu, lv = mu.to(dev), lv.to(dev)
    res = []
    for mode in (0, 1):
        lib.mae_set_mpd_tensor_cores(mode)
        res.append(timeit(lambda: ops.mpd_forward_only(mu, lv)))
    mu.requires_grad_()
    lv.requires_grad_()

    def fb():
        mu.grad = lv.grad = None
        m_d, s = ops.mpd(mu, lv)
        (-torch.nn.functional.logsigmoid(m_d).sum() + s).backward()

    for mode in (0, 1):
        lib.mae_set_mpd_tensor_cores(mode)
        res.append(timeit(fb, 3))
    print("B=%d D=%d: forward only (no KL store) FFMA %.3f ms, tensor %.3f ms;  fwd+bwd (stored KL) FFMA %.3f ms, tensor %.3f ms" % (b, d, *res), flush=True)
print("ALL OK" if ok else "FAILURES")
sys.exit(0 if ok else 1)
