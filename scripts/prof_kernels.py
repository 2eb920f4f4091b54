"""Small driver for ncu: launches each hot kernel a few times at the BASELINE shapes.
    python scripts/prof_kernels.py [dmol_c4|dmol_c3|mpd_small|mpd_large|all]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mae_b200 import ops, synth  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "all"
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
reps = 3

if what in ("dmol_c4", "all"):
    raw = torch.randn(1024, 100, 32, 32, device=dev, generator=g)
    x = torch.randint(0, 256, (2, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1
    for _ in range(reps):
        err = ops.dmol_recon_error(raw, x, 512, 10)
    torch.cuda.synchronize()
    print("dmol_c4", float(err.sum()))
    del raw

if what in ("dmol_c3", "all"):
    inp = synth.color_head(64, 1, 10, 32, 32)
    raw = inp["raw"].to(dev).requires_grad_()
    x = inp["x"].to(dev)
    for _ in range(reps):
        raw.grad = None
        err = ops.dmol_recon_error(raw, x, 1, 10)
        err.sum().backward()
    torch.cuda.synchronize()
    print("dmol_c3", float(err.sum()))

if what in ("mpd_small", "all"):
    mu, lv = synth.posterior_params(64, (64,), clamp=True)
    mu, lv = mu.to(dev).requires_grad_(), lv.to(dev).requires_grad_()
    for _ in range(reps):
        mu.grad = lv.grad = None
        m_d, s = ops.mpd(mu, lv)
        (m_d.sum() + s).backward()
    torch.cuda.synchronize()
    print("mpd_small", float(s))

if what in ("mpd_large", "all"):
    mu, lv = synth.posterior_params(8192, (64,), seed=5)
    mu, lv = mu.to(dev).requires_grad_(), lv.to(dev).requires_grad_()
    for _ in range(2):
        mu.grad = lv.grad = None
        m_d, s = ops.mpd(mu, lv)
        (m_d.sum() + s).backward()
    torch.cuda.synchronize()
    print("mpd_large", float(s))

if what in ("reg_fused", "all"):
    for b in (64, 100):
        mu, lv = synth.posterior_params(b, (64,), clamp=True)
        mu, lv = mu.to(dev), lv.to(dev)
        for _ in range(reps):
            kl, m_d, s, reg, grads, _ = ops.kl_mpd_reg_fused(mu, lv, 1.0, 1.0, 0.0)
        torch.cuda.synchronize()
        print("reg_fused B=%d" % b, float(reg))
