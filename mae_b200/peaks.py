"""Measured pipe peaks of the device the process runs on (FP32 FFMA, SFU ex2 / lg2 / rcp): the denominators of the MPD
(C5) and DMOL co-bound rooflines.  Register-only kernels of the library (csrc/microbench.cu), timed with CUDA events."""
from __future__ import annotations

import ctypes

import torch

from . import _lib


def _time(launch, stream, reps):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    launch()
    torch.cuda.synchronize()
    best = float("inf")
    for _ in range(reps):
        e0.record(stream)
        launch()
        e1.record(stream)
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def measure(device=None, reps: int = 5) -> dict:
    """-> {"ffma_tflops", "mufu_ex2_gops", "mufu_lg2_gops", "mufu_rcp_gops", "sm_count", "how"} (best of ``reps`` launches
    of ~2-4 ms each; 'gops' = 1e9 lane-level SFU instructions per second)."""
    lib = _lib.load()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    with torch.cuda.device(dev):
        stream = torch.cuda.current_stream(dev)
        sm = ctypes.c_int(0)
        _lib.check(lib.mae_device_info(ctypes.byref(sm), None, None), "mae_device_info")
        blocks = sm.value * 4                      # 4 x 512 threads = 64 warps per SM
        sink = torch.zeros(blocks * 512, device=dev)
        ops = ctypes.c_double(0.0)
        out = {"sm_count": sm.value}

        def ffma():
            _lib.check(lib.mae_microbench_ffma(blocks, 4096, sink.data_ptr(), ctypes.byref(ops), stream.cuda_stream),
                       "mae_microbench_ffma")

        t = _time(ffma, stream, reps)
        out["ffma_tflops"] = 2.0 * ops.value / t / 1e12
        for op, name in ((0, "ex2"), (1, "lg2"), (2, "rcp")):
            def mufu(op=op):
                _lib.check(lib.mae_microbench_mufu(op, blocks, 2048, sink.data_ptr(), ctypes.byref(ops), stream.cuda_stream),
                           "mae_microbench_mufu")
            t = _time(mufu, stream, reps)
            out["mufu_%s_gops" % name] = ops.value / t / 1e9
        out["how"] = ("csrc/microbench.cu: %d CTAs x 512 threads, 16 independent FFMA chains (8 SFU chains) per thread, "
                      "best of %d launches, CUDA events" % (blocks, reps))
    return out
