"""2-GPU NCCL test of the batch-sharded head with the real CUDA engine: each rank's loss / gradients must equal the
single-GPU head evaluated on the concatenated batch.  Skipped on boxes with fewer than 2 GPUs."""
import os
import socket
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, small_limit, out):
    # (the NCCL fallback of the gather is exercised by the gloo tests' all_gather_into_tensor path)
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        from mae_b200 import distributed as D, heads, synth
        D.SMALL_GLOBAL_BATCH = small_limit
        head = D.ShardedHead(check_equal_batches=True)
        b_loc, ns, d = 24, 1, 32
        b = world * b_loc
        mu_g, lv_g = synth.posterior_params(b, (d,), seed=1, clamp=True)
        eps_g = synth.noise(b, ns, (d,), seed=1)
        inp = synth.color_head(b, ns, 10, 32, 32, seed=2)
        sl = slice(rank * b_loc, (rank + 1) * b_loc)
        mu, lv = mu_g[sl].to(dev).requires_grad_(), lv_g[sl].to(dev).requires_grad_()
        raw = inp["raw"][sl].to(dev).requires_grad_()
        loss, recon, kl, stats, z = head.loss_head_color(mu, lv, eps_g[sl].to(dev), raw, inp["x"][sl].to(dev), 10, 1.0, 1.0, 0.0)
        loss.backward()
        g_loss = head.global_loss(loss, stats)
        assert head.peer is not None, "peer-memory gather was not used"
        head.peer.check()
        res = {"loss": g_loss.cpu(), "dmu": mu.grad.cpu(), "dlv": lv.grad.cpu(), "draw": raw.grad.cpu()}
        if rank == 0:   # single-GPU head on the whole batch
            mu_a, lv_a = mu_g.to(dev).requires_grad_(), lv_g.to(dev).requires_grad_()
            raw_a = inp["raw"].to(dev).requires_grad_()
            ref = heads.loss_head_color(mu_a, lv_a, eps_g.to(dev), raw_a, inp["x"].to(dev), 10, 1.0, 1.0, 0.0)
            ref[0].backward()
            res["ref"] = {"loss": ref[0].detach().cpu(), "dmu": mu_a.grad.cpu(), "dlv": lv_a.grad.cpu(), "draw": raw_a.grad.cpu()}
        torch.save(res, out % rank)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("small_limit", [2048, 16])
def test_sharded_head_two_gpus(tmp_path, small_limit):
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from tolerance import check_close
    out = str(tmp_path / "r%d.pt")
    mp.spawn(_worker, args=(2, _free_port(), small_limit, out), nprocs=2, join=True)
    res = [torch.load(out % r) for r in range(2)]
    ref = res[0]["ref"]
    for r in range(2):
        sl = slice(r * 24, (r + 1) * 24)
        check_close(res[r]["loss"], ref["loss"].double(), rel=2e-6, what="rank %d loss" % r)
        check_close(res[r]["dmu"], ref["dmu"][sl].double(), rel=1e-5, what="rank %d dmu" % r)
        check_close(res[r]["dlv"], ref["dlv"][sl].double(), rel=1e-5, what="rank %d dlv" % r)
        check_close(res[r]["draw"], ref["draw"][sl].double(), rel=1e-6, what="rank %d draw" % r)
