"""Times the DMOL forward at the IW-eval shape (N=1024 samples of [100,32,32], ring of 3 slabs > L2) and the C3 shape."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from mae_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
g = torch.Generator(device=dev).manual_seed(1)
def run(N, B, ring, bwd=False):
    raws = [torch.randn(N, 100, 32, 32, device=dev, generator=g) for _ in range(ring)]
    x = torch.randint(0, 256, (B, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1
    err = torch.empty(N, device=dev)
    gg = torch.full((N,), 1.0 / N, device=dev)
    draw = torch.empty_like(raws[0]) if bwd else None
    ws = torch.zeros(lib.mae_dmol_workspace_bytes(N, 1024), dtype=torch.uint8, device=dev)
    st = torch.cuda.current_stream().cuda_stream
    def f(i):
        r = raws[i % ring]
        if bwd:
            lib.mae_dmol_bwd(r.data_ptr(), x.data_ptr(), gg.data_ptr(), 1, 1.0, B, N // B, 10, 1024, draw.data_ptr(), st)
        else:
            lib.mae_dmol_fwd(r.data_ptr(), x.data_ptr(), B, N // B, 10, 1024, err.data_ptr(), ws.data_ptr(), ws.numel(), st)
    for i in range(5): f(i)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    n = 30
    e0.record()
    for i in range(n): f(i)
    e1.record(); torch.cuda.synchronize()
    t = e0.elapsed_time(e1) / n * 1e-3
    bytes_ = 4 * (100 * 1024 * N) * (2 if bwd else 1)
    return t * 1e6, bytes_ / t / 1e9
tag = "SPLIT=%s OCC=%s" % (os.environ.get("MAE_DMOL_SPLIT", "auto"), os.environ.get("MAE_DMOL_S1_OCC", "4"))
t, bw = run(1024, 2, 3)
print("%-22s C4 fwd N=1024: %8.1f us  %7.0f GB/s (%.1f%% of 6541)" % (tag, t, bw, bw / 65.415))
t, bw = run(64, 64, 6)
print("%-22s C3 fwd N=64  : %8.1f us  %7.0f GB/s (%.1f%%)" % (tag, t, bw, bw / 65.415))
t, bw = run(64, 64, 6, bwd=True)
print("%-22s C3 bwd N=64  : %8.1f us  %7.0f GB/s (%.1f%%)" % (tag, t, bw, bw / 65.415))
# fused forward+backward
N = 64
raws = [torch.randn(N, 100, 32, 32, device=dev, generator=g) for _ in range(6)]
x = torch.randint(0, 256, (N, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1
err = torch.empty(N, 64, device=dev); draw = torch.empty_like(raws[0])
st = torch.cuda.current_stream().cuda_stream
def ff(i):
    r = raws[i % 6]
    lib.mae_dmol_fwd_bwd(r.data_ptr(), x.data_ptr(), 1.0 / N, N, 1, 10, 1024, err.data_ptr(), draw.data_ptr(), st)
for i in range(5): ff(i)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for i in range(30): ff(i)
e1.record(); torch.cuda.synchronize()
t = e0.elapsed_time(e1) / 30 * 1e-3
print("%-22s C3 fused N=64: %8.1f us  moved %.0f GB/s (%.1f%%), algorithmic 80.2 MB -> %.1f%%" % (tag, t * 1e6, 53.2e6 / t / 1e9, 53.2e6 / t / 65.415e9, 80.2e6 / t / 65.415e9))
