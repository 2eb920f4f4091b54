"""Phase timestamps of the one-launch sharded regulariser (build with MAE_DEBUG_TIMING=1).  Run under torchrun."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from mae_b200 import ops, synth, distributed as D
head = D.ShardedHead(check_every=0)
b_loc, d = 64, 64
mu, lv = synth.posterior_params(b_loc, (d,), seed=10 + rank, clamp=True)
packed = torch.cat([mu, lv], 1).to(dev).contiguous()
assert head._rows_cluster_ok(b_loc, d, packed)
scr = head._rows_scratch
names = ["start", "pushed", "rows landed", "phase1", "sync1", "phase2", "sync2", "phase3", "sums landed", "sync3", "end"]
acc = torch.zeros(11, dtype=torch.float64)
n = 0
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
for it in range(300):
    if it % 50 == 0:
        dist.barrier(); torch.cuda.synchronize()
    ops.reg_rows_cluster(packed, head.peer_rows, 1.0, 1.0, 0.0, None, True, scr)
    if it >= 100 and it % 10 == 5:
        torch.cuda.synchronize()
        st = scr[-128:].view(torch.int64)[:11].cpu().double()
        acc += st - st[0]
        n += 1
dist.barrier(); torch.cuda.synchronize()
e0.record()
for it in range(200):
    ops.reg_rows_cluster(packed, head.peer_rows, 1.0, 1.0, 0.0, None, True, scr)
e1.record(); torch.cuda.synchronize()
head.check()
if rank == 0:
    acc /= n
    print("world %d: back-to-back %.2f us/launch; CTA 0 phase stamps (us since start):" % (world, e0.elapsed_time(e1) / 200 * 1e3))
    print("  " + " | ".join("%s %.2f" % (nm, v / 1e3) for nm, v in zip(names, acc.tolist())))
dist.barrier()
dist.destroy_process_group()
