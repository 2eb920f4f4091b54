"""Tolerances used by the parity tests (BASELINE.json north_star; SURVEY.md §7 hard part 3).

* loss terms and gradients: 1e-5 relative to the reference.  For tensors "relative" is taken against
  the tensor's own scale (max |ref|), the usual norm-wise reading; element-wise relative error is
  meaningless for gradient entries that are themselves the result of cancellation.
* where the fp32 reference itself has lost digits (posterior collapse, cdf differences near the 1e-5
  threshold), the fp64-anchored criterion applies:
      |ours - ref64| <= max(1e-5 * scale, 4 * |ref32 - ref64|)
* eval NLL: 1e-4 bits/dim, nll / (nx * ln 2)  (experiments/color_image.py:275).
"""
import math

import torch

REL = 1e-5
BPD = 1e-4

# every check_close also records the ELEMENT-WISE relative error of the entries that are not themselves cancellation
# residue (|ref| >= 1e-3 of the tensor's scale): max and 99.9th percentile.  conftest.py prints the worst ones at the end of
# the run, so the distance between the norm-wise criterion above and north_star's element-wise wording is on record.
ELEMENTWISE = []


def scale_of(ref: torch.Tensor) -> float:
    ref = ref.detach().double()
    return float(ref.abs().max()) if ref.numel() else 0.0


def check_close(ours, ref64, ref32=None, rel=REL, what="", floor=0.0):
    """Assert |ours-ref64| <= max(rel*scale, 4*|ref32-ref64|max, floor); returns the achieved rel error."""
    ours = torch.as_tensor(ours).detach().double().cpu().reshape(-1)
    r64 = torch.as_tensor(ref64).detach().double().cpu().reshape(-1)
    assert ours.shape == r64.shape, "%s: shape %s vs %s" % (what, tuple(ours.shape), tuple(r64.shape))
    assert torch.isfinite(ours).all(), "%s: non-finite values" % what
    sc = max(scale_of(r64), 1e-30)
    err = float((ours - r64).abs().max()) if ours.numel() else 0.0
    bound = rel * sc
    if ref32 is not None:
        r32 = torch.as_tensor(ref32).detach().double().cpu().reshape(-1)
        bound = max(bound, 4.0 * float((r32 - r64).abs().max()))
    bound = max(bound, floor)
    if ours.numel():
        big = r64.abs() >= 1e-3 * sc
        if bool(big.any()):
            e = ((ours - r64).abs()[big] / r64.abs()[big])
            q = float(torch.quantile(e[:1 << 22], 0.999)) if e.numel() > 1 else float(e.max())
            ELEMENTWISE.append((float(e.max()), q, err / sc, what))
    assert err <= bound, "%s: max abs err %.3e > bound %.3e (scale %.3e, rel %.3e)" % (what, err, bound, sc, err / sc)
    return err / sc


def check_bpd(ours_nll, ref64_nll, nx, what=""):
    ours = torch.as_tensor(ours_nll).detach().double().cpu().reshape(-1)
    r64 = torch.as_tensor(ref64_nll).detach().double().cpu().reshape(-1)
    err = float((ours - r64).abs().max()) / (nx * math.log(2.0))
    assert err <= BPD, "%s: %.3e bits/dim > %.1e" % (what, err, BPD)
    return err
