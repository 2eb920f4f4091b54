"""Batched, shardable test-NLL evaluation.

The reference's ``calc_nll`` (experiments/binary_image.py:263-287, experiments/color_image.py:252-277) feeds ONE image
per ``MAE.nll`` call: 10 000 calls with K = 512 / 4096 samples each, every call a chain of small launches.  ``calc_nll``
below returns the same tuple ``((nll_elbo, recon, kl), nll_iw, bits_per_dim)`` but pushes ``images_per_step`` images x K
samples through the decoder and the fused IW kernel per call (the kernels are sized for it: the DMOL forward reaches its
bandwidth plateau from ~2 images x 512 samples on), accumulates the four sums on the device (one host read at the end),
and with a ``ShardedHead`` splits the images over ranks with a single all-reduce at the end (§8e)."""
from __future__ import annotations

import math
from typing import Optional

import torch


def calc_nll(model, data: torch.Tensor, k: int, images_per_step: int = 8, sharded=None, device: Optional[torch.device] = None,
             cuda_graph: bool = False):
    """model: ``mae_b200.modules.MAE``; data: [N, *x] tensor (host or device) of test images; k: importance samples.
    -> ((nll_elbo, recon, kl), nll_iw, bits_per_dim) as Python floats, averaged over the N images (all ranks return the
    global averages when ``sharded`` is given).

    ``cuda_graph``: the body of a full batch (``model.nll`` on ``images_per_step`` images and the four running sums) is
    captured once and replayed per batch - evaluation is a long chain of small launches behind the decoder (flows, log
    densities, the IW kernel), and the reference's loop is host-bound; the ragged last batch runs eagerly.  Posterior
    samples come from torch's graph-safe generator, so every replay draws fresh noise."""
    if images_per_step < 1:
        raise ValueError("images_per_step must be >= 1")
    n_total = data.size(0)
    if n_total == 0:
        raise ValueError("calc_nll needs at least one image")
    lo, hi = (0, n_total) if sharded is None else sharded.shard(n_total)
    dev = device if device is not None else next(model.parameters()).device
    nx = data[0].numel()
    was_training = model.training
    model.eval()
    sums = torch.zeros(5, device=dev, dtype=torch.float64)      # nll_elbo, recon, kl, nll_iw, count

    def body(x):
        (elbo, recon, kl), iw = model.nll(x, k)
        sums[0] += elbo.sum()
        sums[1] += recon.sum()
        sums[2] += kl.sum()
        sums[3] += iw.sum()
        sums[4] += x.size(0)

    try:
        with torch.no_grad():
            start = lo
            n_full = (hi - lo) // images_per_step
            if cuda_graph and n_full >= 3 and torch.device(dev).type == "cuda":
                xs = data[start:start + images_per_step].to(dev).clone()
                side = torch.cuda.Stream(device=dev)
                side.wait_stream(torch.cuda.current_stream(dev))
                with torch.cuda.stream(side):
                    body(xs)                                     # warm-up (counts: it IS the first batch)
                torch.cuda.current_stream(dev).wait_stream(side)
                torch.cuda.synchronize(dev)
                start += images_per_step
                g = torch.cuda.CUDAGraph()
                xs.copy_(data[start:start + images_per_step].to(dev, non_blocking=True))
                with torch.cuda.graph(g):
                    body(xs)                                     # captured, not executed
                g.replay()                                       # second batch
                start += images_per_step
                for _ in range(n_full - 2):
                    xs.copy_(data[start:start + images_per_step].to(dev, non_blocking=True), non_blocking=True)
                    g.replay()
                    start += images_per_step
            for st in range(start, hi, images_per_step):
                body(data[st:min(hi, st + images_per_step)].to(dev, non_blocking=True))
    finally:
        model.train(was_training)
    if sharded is not None:
        sums = sharded.reduce_eval(sums)
    s = sums.tolist()                                            # the only host synchronisation
    n = s[4]
    nll_elbo, recon_err, kl_err, nll_iw = (v / n for v in s[:4])
    return (nll_elbo, recon_err, kl_err), nll_iw, nll_iw / (nx * math.log(2.0))
