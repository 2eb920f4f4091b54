"""One MPD forward+backward at a C5 shape (profiler target): python scripts/dev_mpd_once.py B D [tc]"""
import sys
import torch
sys.path.insert(0, ".")
from mae_b200 import _lib, ops, synth
b, d = int(sys.argv[1]), int(sys.argv[2])
_lib.load().mae_set_mpd_tensor_cores(int(sys.argv[3]) if len(sys.argv) > 3 else 1)
mu, lv = synth.posterior_params(b, (d,), seed=5)
mu, lv = mu.cuda().requires_grad_(), lv.cuda().requires_grad_()
for _ in range(2):
    mu.grad = lv.grad = None
    m_d, s = ops.mpd(mu, lv)
    (-torch.nn.functional.logsigmoid(m_d).sum() + s).backward()
torch.cuda.synchronize()
print("done", float(s))
