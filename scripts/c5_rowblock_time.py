"""C5 row-block (sharded) forward+backward timing of one shape under torchrun: python scripts/c5_rowblock_time.py B D"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from mae_b200 import synth
from mae_b200.distributed import ShardedHead
b, d = int(sys.argv[1]), int(sys.argv[2])
head = ShardedHead(check_every=0, overlap=False)
b_loc = b // world
mu_g, lv_g = synth.posterior_params(b, (d,), seed=5)
sl = slice(rank * b_loc, (rank + 1) * b_loc)
mu = mu_g[sl].to(dev).requires_grad_(); lv = lv_g[sl].to(dev).requires_grad_()
def step():
    mu.grad = lv.grad = None
    _, m_d, s_, _ = head.kl_mpd(mu, lv)
    (-torch.nn.functional.logsigmoid(m_d).sum() + s_).backward()
for _ in range(2): step()
ts = []
for rep in range(4):
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); step(); step(); e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / 2], device=dev, dtype=torch.float64); dist.all_reduce(t, op=dist.ReduceOp.MAX); ts.append(float(t))
if rank == 0: print("B=%d D=%d world %d: fwd+bwd ms per step, 4 repeats:" % (b, d, world), ["%.1f" % x for x in ts])
dist.barrier(); os._exit(0)
