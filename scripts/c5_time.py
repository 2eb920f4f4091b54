"""C5 timing of the MPD forward / backward kernels alone (B=8192, D=64 by default)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mae_b200 import _lib, synth
lib = _lib.load()
dev = torch.device("cuda")
B, D = int(os.environ.get("B", 8192)), int(os.environ.get("D", 64))
mu, lv = synth.posterior_params(B, (D,), seed=5)
mu, lv = mu.to(dev), lv.to(dev)
P = lambda t: t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
ws = torch.zeros(lib.mae_mpd_workspace_bytes(B, D, 1), dtype=torch.uint8, device=dev)
m_d = torch.empty(D, device=dev); sumsq = torch.empty(1, device=dev, dtype=torch.float64); S = torch.empty((), device=dev)
kl = torch.empty(B, device=dev); dmu = torch.empty(B, D, device=dev); dlv = torch.empty_like(dmu)
g_m = torch.randn(D, device=dev); g_S = torch.ones((), device=dev)
def t(fn, n=10):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
prep = t(lambda: lib.mae_mpd_prepare(P(mu), D, P(lv), D, B, D, P(ws), ws.numel(), P(m_d), P(kl), st))
fwd = t(lambda: lib.mae_mpd_fwd(P(ws), ws.numel(), B, D, 0, B, 1, P(sumsq), P(S), st))
res = {}
for path in (3, 2):
    res[path] = t(lambda: lib.mae_mpd_bwd(P(mu), D, P(lv), D, P(ws), ws.numel(), B, D, 0, B, P(g_m), P(g_S), P(S), P(kl), P(dmu), P(dlv), 0, path, st))
fl = B * B * D
print("B=%d D=%d variant=%s: prepare %.3f ms | fwd %.3f ms (%.1f TF/s of 4B^2D) | bwd path3 %.3f ms (%.1f TF/s of 8B^2D executed) | bwd path2 %.3f ms (%.1f TF/s of 16B^2D executed)" % (
    B, D, os.environ.get("MAE_MPD_BWD_VARIANT", "0"), prep, fwd, 4 * fl / fwd / 1e9, res[3], 8 * fl / res[3] / 1e9, res[2], 16 * fl / res[2] / 1e9))
