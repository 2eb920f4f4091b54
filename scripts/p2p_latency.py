"""torchrun --nproc-per-node N scripts/p2p_latency.py : in-graph latency of the peer-memory gather vs NCCL all_gather
for the 32 KB/rank panel of the training shapes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from mae_b200.p2p import PeerGather

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x = torch.randn(64, 128, device=dev)
pg = PeerGather(None, x.numel(), dev)
out_nccl = torch.empty(world * 64, 128, device=dev)
REP = 50

def timeit(fn, name):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize(); dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        g.replay()
    e1.record(); torch.cuda.synchronize()
    if rank == 0:
        print("%-28s %7.2f us per call (world %d)" % (name, e0.elapsed_time(e1) / 20 / REP * 1e3, world), flush=True)

res = {}
def f_p2p():
    res["o"] = pg.gather(x)
timeit(f_p2p, "peer-memory gather kernel")
ref = torch.cat([x] * world) if False else None
o = pg.gather(x); torch.cuda.synchronize()
assert torch.equal(o[rank * 64:(rank + 1) * 64], x)
pg.check()
timeit(lambda: dist.all_gather_into_tensor(out_nccl, x), "NCCL all_gather_into_tensor")
sys.stdout.flush()
os._exit(0)
