"""Where the multi-GPU C3 step goes: in-graph, dependency-chained latency of every piece of a rank's
regulariser side chain at batch 64 per GPU, global batch 64 x world, and of the main-stream kernels it has to hide behind.
Each piece is captured REP times back to back in one CUDA graph (every copy depends on the previous one) and replayed; the
figure is kernel duration + in-graph dependent-launch gap, i.e. what the piece costs on the critical path.  Run under
torchrun with N ranks (all ranks execute the same sequence: the gather is a cross-rank exchange)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from mae_b200 import _lib, ops, synth, distributed as D  # noqa: E402

lib = _lib.load()
REP = 20


def chain(fn, n=30):
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    dist.barrier()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / n / REP * 1e3], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t)


b_loc, d = 64, 64
bg = b_loc * world
D.ROWS_CLUSTER_MAX_WORLD = 0           # pieces of the gather -> prepare -> tiles -> backward chain first
head = D.ShardedHead(check_every=0, overlap=False)
mu, lv = synth.posterior_params(b_loc, (d,), seed=10 + rank, clamp=True)
packed = torch.cat([mu, lv], 1).to(dev).contiguous()
row0 = rank * b_loc
res = {}
gathered = head._gather(packed)
res["peer gather (push rows over NVLink, flags, copy-out)"] = chain(lambda: head._gather(packed))
mu_all, lv_all = gathered[:, :d], gathered[:, d:]


def fwd():
    m_d, s, kl_all, state = ops.mpd_rows_forward(mu_all, lv_all, row0, b_loc, full=True)
    return m_d, s, kl_all, state


def fwd_only():
    m_d, s, kl_all, state = fwd()
    ops._pool.give(state.ws)


def fwd_bwd():
    m_d, s, kl_all, state = fwd()
    ops.mpd_rows_loss_backward(state, row0, b_loc, 1.0, 1.0, 0.0, 1.0 / bg, m_d, kl_all, True)


def whole():
    g = head._gather(packed)
    m_d, s, kl_all, state = ops.mpd_rows_forward(g[:, :d], g[:, d:], row0, b_loc, full=True)
    ops.mpd_rows_loss_backward(state, row0, b_loc, 1.0, 1.0, 0.0, 1.0 / bg, m_d, kl_all, True)


res["pre-pass + tiles over the %d global rows (2 launches)" % bg] = chain(fwd_only)
t_fb = chain(fwd_bwd)
res["upstream gradients + own-row backward (2 launches)"] = t_fb - res["pre-pass + tiles over the %d global rows (2 launches)" % bg]
res["whole side chain: gather -> pre-pass -> tiles -> backward"] = chain(whole)
# the one-launch exchange + regulariser kernel on the same shape
D.ROWS_CLUSTER_MAX_WORLD = 64
head2 = D.ShardedHead(check_every=0, overlap=False)
if head2._rows_cluster_ok(b_loc, d, packed):
    scr = head2._rows_scratch
    res["one-launch exchange + regulariser cluster kernel"] = chain(lambda: ops.reg_rows_cluster(packed, head2.peer_rows, 1.0, 1.0, 0.0, None, True, scr))
# main-stream kernels of the step (per rank: 64 images)
inp = synth.color_head(b_loc, 1, 10, 32, 32)
raw, x = inp["raw"].to(dev), inp["x"].to(dev)
chunks = int(lib.mae_dmol_err_chunks(b_loc, 1, 10, 1024))
errp = torch.empty(b_loc, 1, chunks, device=dev)
draw = torch.empty_like(raw)
st = lambda: torch.cuda.current_stream().cuda_stream
res["likelihood: fused DMOL forward+backward, 64 images (L2-resident input)"] = chain(
    lambda: lib.mae_dmol_fwd_bwd(raw.data_ptr(), x.data_ptr(), 1.0 / bg, b_loc, 1, 10, 1024, errp.data_ptr(), draw.data_ptr(), st()))
eps = synth.noise(b_loc, 1, (d,)).to(dev)
mu_d, lv_d = mu.to(dev), lv.to(dev)
gz = torch.randn(b_loc, 1, d, device=dev)
gr = torch.randn(2, b_loc, d, device=dev)
one = torch.ones((), device=dev)
dmu, dlv = torch.empty(b_loc, d, device=dev), torch.empty(b_loc, d, device=dev)
res["reparameterisation backward + regulariser gradient (1 launch)"] = chain(
    lambda: lib.mae_reparam_reg_bwd(mu_d.data_ptr(), d, lv_d.data_ptr(), d, eps.data_ptr(), gz.data_ptr(), None, gr[0].data_ptr(),
                                    gr[1].data_ptr(), one.data_ptr(), b_loc, 1, d, dmu.data_ptr(), dlv.data_ptr(), 0, st()))
head.check()
if rank == 0:
    print("world %d, batch %d per GPU, global batch %d: in-graph dependent latency per piece (us, max over ranks)" % (world, b_loc, bg))
    for k, v in res.items():
        print("  %-78s %7.2f" % (k, v))
dist.barrier()
torch.cuda.synchronize()
os._exit(0)
