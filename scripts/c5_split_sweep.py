"""Tuning aid: time the tensor-core MPD forward+backward at one shape (run once per MAE_MPD_TC_SPLIT value)."""
import sys
import torch
sys.path.insert(0, ".")
from mae_b200 import ops, synth
b, d = int(sys.argv[1]), int(sys.argv[2])
mu, lv = synth.posterior_params(b, (d,), seed=5)
mu, lv = mu.cuda().requires_grad_(), lv.cuda().requires_grad_()


def step():
    mu.grad = lv.grad = None
    m_d, s = ops.mpd(mu, lv)
    (-torch.nn.functional.logsigmoid(m_d).sum() + s).backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
best = 1e9
for _ in range(5):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        step()
    e1.record()
    torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1) / 10)
print("B=%d D=%d: %.3f ms (eager, incl. host launches)" % (b, d, best))
