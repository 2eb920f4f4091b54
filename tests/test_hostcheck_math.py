"""Validates the per-pixel arithmetic shared with the CUDA kernels (mae_b200/csrc/hot_math.cuh) on the CPU:
the header is compiled for the host by g++ (tests/hostcheck/hostcheck.cpp) and compared with the golden
fixtures generated from the live reference.  Catches algebra / branch mistakes without a GPU."""
import ctypes
import os
import subprocess

import pytest
import torch

from tolerance import check_close

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "hostcheck", "hostcheck.cpp")
LIB = os.path.join(HERE, "hostcheck", "libhostcheck.so")


@pytest.fixture(scope="module")
def hc():
    hdr = os.path.join(HERE, "..", "mae_b200", "csrc", "hot_math.cuh")
    if not os.path.exists(LIB) or os.path.getmtime(LIB) < max(os.path.getmtime(SRC), os.path.getmtime(hdr)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-ffp-contract=off", "-x", "c++", SRC, "-o", LIB])
    return ctypes.CDLL(LIB)


def _p(t):
    return ctypes.c_void_p(t.data_ptr())


def test_dmol_pixel_math(hc, golden):
    for name, c in golden("dmol").items():
        raw, x, g = c["raw"].contiguous(), c["x"].contiguous(), c["g"].contiguous()
        b, ns, nmix = x.size(0), c["ns"], c["nmix"]
        hw = x.size(2) * x.size(3)
        err = torch.empty(b, ns)
        draw = torch.empty_like(raw)
        i64 = ctypes.c_int64
        hc.hc_dmol_fwd(_p(raw), _p(x), i64(b), i64(ns), i64(nmix), i64(hw), _p(err))
        hc.hc_dmol_bwd(_p(raw), _p(x), _p(g), i64(b), i64(ns), i64(nmix), i64(hw), _p(draw))
        check_close(err, c["ref64"]["err"], c["ref32"]["err"], what=name + " err")
        check_close(draw, c["ref64"]["draw"], c["ref32"]["draw"], what=name + " draw")
        err2 = torch.empty(b, ns)
        hc.hc_dmol_fwd_pair(_p(raw), _p(x), i64(b), i64(ns), i64(nmix), i64(hw), _p(err2))   # packed two-pixel path
        check_close(err2, c["ref64"]["err"], c["ref32"]["err"], what=name + " err (packed pair path)")


def test_dmol_forward_extremes_vs_oracle(hc):
    """the forward path (one reciprocal and one logarithm per component, product over the three channels) on adversarial
    inputs: log-scales far below the -7 clamp and up to +6, means far outside [-1, 1], every pixel value incl. both edge
    bins, saturated coefficients -> must stay finite and match the fp64 oracle like the goldens do"""
    from oracle import mae_oracle as O
    g = torch.Generator().manual_seed(11)
    b, ns, nmix, h, w = 3, 2, 10, 8, 8
    raw = torch.randn(b * ns, 10 * nmix, h, w, generator=g)
    raw[:, nmix:4 * nmix] *= 3.0                                            # means up to +-9
    raw[:, 4 * nmix:7 * nmix] = torch.empty(b * ns, 3 * nmix, h, w).uniform_(-12.0, 6.0, generator=g)   # raw log-scales
    raw[:, 7 * nmix:] *= 6.0                                                # tanh saturates
    x = (torch.randint(0, 256, (b, 3, h, w), generator=g).float() / 127.5 - 1.0)
    x[0, :, 0, :4] = 1.0                                                    # 255
    x[0, :, 0, 4:] = -1.0                                                   # 0
    err = torch.empty(b, ns)
    i64 = ctypes.c_int64
    hc.hc_dmol_fwd(_p(raw), _p(x), i64(b), i64(ns), i64(nmix), i64(h * w), _p(err))
    assert torch.isfinite(err).all()
    ref64 = O.dmol_recon_error(raw.double(), x.double(), ns, nmix)
    ref32 = O.dmol_recon_error(raw, x, ns, nmix)
    check_close(err, ref64, ref32, what="err on extreme inputs")


def test_bernoulli_math(hc, golden):
    for name, c in golden("bernoulli").items():
        probs, x, g = c["probs"].contiguous(), c["x"].contiguous(), c["g"].contiguous()
        b, ns, p = probs.shape
        err = torch.empty(b, ns)
        d = torch.empty_like(probs)
        i64 = ctypes.c_int64
        hc.hc_bernoulli(_p(probs), _p(x), _p(g), i64(b), i64(ns), i64(p), _p(err), _p(d))
        check_close(err, c["ref64"]["err"], c["ref32"]["err"], what=name + " err")
        check_close(d, c["ref64"]["dprobs"], c["ref32"]["dprobs"], what=name + " dprobs")
