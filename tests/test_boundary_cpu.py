"""CPU-side checks of the drop-in boundary: the C-ABI library builds, loads and exports every symbol the header
declares (no compute calls — there is no GPU here), argument validation returns error codes, the Python op layer
refuses CPU tensors (no fallback), and nothing in the product imports the oracle."""
import ctypes
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from mae_b200 import build, _lib
    build.build_library()
    return _lib.load()


def test_header_symbols_exported(lib):
    from mae_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "mae_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = sorted(set(re.findall(r"\b(mae_[a-z0-9_]+)\s*\(", hdr)))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), "libmae_b200.so does not export %s" % name
    assert sorted(_lib.ALL_SYMBOLS) == declared, "python binding table and header disagree"


def test_abi_version_and_strerror(lib):
    assert lib.mae_abi_version() == 1
    assert b"ok" in lib.mae_strerror(0)
    assert b"NULL" in lib.mae_strerror(-1)
    assert b"workspace" in lib.mae_strerror(-4)


def test_argument_validation_without_gpu(lib):
    # argument errors are detected before any CUDA call, so they are testable on a CPU-only host
    assert lib.mae_reparam_kl_fwd(None, 4, None, 4, None, 1, 1, 4, None, None, None) == -1
    buf = ctypes.create_string_buffer(1 << 16)
    p = ctypes.addressof(buf)
    assert lib.mae_reparam_kl_fwd(p, 2, p, 4, p, 1, 1, 4, p, None, None) == -3       # stride < D
    assert lib.mae_reparam_kl_fwd(p, 4, p, 4, p, 0, 1, 4, p, None, None) == -2       # B = 0
    assert lib.mae_bernoulli_fwd(p, p, 1, 1, 0, p, None) == -2
    assert lib.mae_dmol_fwd(p, p, 1, 1, 0, 16, p, p, 1024, None) == -2
    assert lib.mae_mpd_workspace_bytes(1, 8, 0) == 0                                    # MPD needs B >= 2
    need = lib.mae_mpd_workspace_bytes(100, 64, 0)
    assert lib.mae_mpd_workspace_bytes(100, 64, 1) == need + 2 * 4 * 128 * 128              # KL and KL^T, Bp = 128
    assert need > 0 and need % 256 == 0
    aligned = (p + 255) & ~255
    assert lib.mae_mpd_prepare(aligned, 64, aligned, 64, 100, 64, aligned, need - 1, aligned, None, None) == -4
    assert lib.mae_mpd_prepare(aligned, 64, aligned, 64, 1, 64, aligned, need, aligned, None, None) == -2
    assert lib.mae_mpd_prepare(aligned, 64, aligned, 64, 100, 64, aligned + 4, need, aligned, None, None) == -5
    assert lib.mae_mpd_fwd(aligned, need, 100, 64, 50, 60, 0, aligned, None, None) == -2   # rows past B
    assert lib.mae_dmol_workspace_bytes(64, 1024) >= 64 * 4 * 4


def test_tensor_core_backward_splits_fill_whole_rounds(lib):
    """host logic of the tensor-core MPD backward (csrc/mpd.cu tc_splits): one CTA per SM, so the CTA count
    row tiles x column chunks x splits should fill its last round of 148; the chosen split must be at least as good as
    every smaller one (ties go to the smaller count) and keep >= 4 partner tiles (one accumulation group) per split."""
    f = lib.mae_debug_mpd_tc_splits
    f.restype = ctypes.c_longlong
    f.argtypes = [ctypes.c_longlong] * 3

    def eff(ctas):
        return ctas / (-(-ctas // 148) * 148)

    for b in (2048, 4096, 8192, 16384, 32768, 65536):
        for d in (64, 128, 192, 256):
            for world in (1, 2, 4, 8):
                row_tiles, cols, partners = b // 128 // world, 2 * d // 128, b // 128
                if row_tiles == 0:
                    continue
                s = f(row_tiles, cols, partners)
                assert 1 <= s <= 8 and s <= partners
                assert s == 1 or -(-partners // s) >= 4
                for smaller in range(1, s):
                    assert eff(row_tiles * cols * s) > eff(row_tiles * cols * smaller), (b, d, world, s, smaller)
    # the cases quoted in DESIGN.md 4
    assert f(128, 1, 128) == 8 and f(128, 4, 128) == 2 and f(64, 1, 64) == 2


def test_ops_refuse_cpu_tensors(lib):
    from mae_b200 import ops
    mu, lv, eps = torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 1, 3)
    for fn in (lambda: ops.reparam_kl(mu, lv, eps), lambda: ops.mpd(mu, lv),
               lambda: ops.bernoulli_recon_error(torch.rand(2, 1, 4), torch.rand(2, 4)),
               lambda: ops.dmol_recon_error(torch.rand(2, 10, 2, 2), torch.rand(2, 3, 2, 2), 1, 1),
               lambda: ops.log_prob_prior(torch.rand(2, 3)),
               lambda: ops.assemble_loss(torch.rand(2, 1), torch.rand(2), None, None)):
        with pytest.raises(RuntimeError, match="CUDA tensors only"):
            fn()


def test_missing_library_fails_loudly(monkeypatch):
    from mae_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libmae_b200.so")
    with pytest.raises(_lib.MaeLibraryError, match="no CPU / PyTorch fallback"):
        _lib.load()


def test_product_never_imports_oracle():
    pat = re.compile(r"^\s*(from|import)\s+oracle\b|oracle[./]mae_oracle", re.M)
    for base, _, files in os.walk(os.path.join(ROOT, "mae_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(base, f)).read()
                assert not pat.search(src), "%s references the oracle" % f


def test_rows_helper_keeps_chunk_views():
    from mae_b200.ops import _rows
    out = torch.randn(5, 16)
    mu, lv = out.chunk(2, 1)
    v, stride = _rows(mu)
    assert stride == 16 and v.data_ptr() == mu.data_ptr()
    v, stride = _rows(lv)
    assert stride == 16 and v.data_ptr() == lv.data_ptr()
    out4 = torch.randn(3, 8, 2, 2)
    mu4, _ = out4.chunk(2, 1)
    v, stride = _rows(mu4)
    assert v.shape == (3, 16) and stride == 32 and v.data_ptr() == mu4.data_ptr()
    v, stride = _rows(torch.randn(4, 6).t())
    assert v.is_contiguous() and stride == 4


def test_registries_and_from_params():
    import stubs  # noqa: F401  registers the test cores
    from mae_b200.modules import MAE, Decoder, Encoder, EncoderCore, Flow
    assert Encoder.by_name('gaussian').__name__ == 'GaussianEncoder'
    params = {"encoder": {"type": "gaussian", "core": {"type": "test_linear_core", "nin": 36, "nz": 8, "clamp": False},
                          "prior_flow": {"type": "test_affine_flow", "nz": 8}},
              "decoder": {"type": "test_binary_linear", "nz": 8, "npix": 36}}
    mae = MAE.from_params(params)
    assert mae.z_shape() == (8,)
    keys = set(mae.state_dict().keys())
    assert {"encoder.core.linear.weight", "encoder.prior_flow.a", "decoder.linear.bias"} <= keys
    with pytest.raises(ValueError, match="does not match decoder input shape"):
        MAE(mae.encoder, stubs.BinaryLinearDecoder(7, 36))
    assert EncoderCore.by_name('test_linear_core') is stubs.LinearCore and Flow.by_name('test_affine_flow') is stubs.AffineFlow
    assert Decoder.by_name('test_color_linear') is stubs.ColorLinearDecoder
