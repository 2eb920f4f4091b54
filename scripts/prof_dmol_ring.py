"""Profiler target for roofline.traffic: the benched fused DMOL forward+backward kernel (dmol_bwd_kernel<10,1024,5,ERR=true>,
mae_dmol_fwd_bwd at the C3 shape) launched back to back over rings of 8 input and 8 output slabs (> 3x L2), as in bench.py.
Capture a launch in the middle with  ncu --cache-control none  so that the DRAM writes of the evicted d raw lines are counted."""
import sys

import torch

sys.path.insert(0, ".")
from mae_b200 import _lib

lib = _lib.load()
dev = torch.device("cuda:0")
b, nmix, hw = 64, 10, 1024
g = torch.Generator(device=dev).manual_seed(3)
raws = [torch.randn(b, 100, 32, 32, device=dev, generator=g) for _ in range(8)]
draws = [torch.empty(b, 100, 32, 32, device=dev) for _ in range(8)]
x = torch.randint(0, 256, (b, 3, 32, 32), device=dev, generator=g).float() / 127.5 - 1.0
chunks = int(lib.mae_dmol_err_chunks(b, 1, nmix, hw))
errp = torch.empty(b, 1, chunks, device=dev)
s = torch.cuda.current_stream().cuda_stream
for i in range(32):
    _lib.check(lib.mae_dmol_fwd_bwd(raws[i % 8].data_ptr(), x.data_ptr(), 1.0 / b, b, 1, nmix, hw, errp.data_ptr(), draws[i % 8].data_ptr(), s),
               "mae_dmol_fwd_bwd")
torch.cuda.synchronize()
print("done")
