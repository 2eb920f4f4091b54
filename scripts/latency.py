"""Warm, dependency-chained latency of each op at the training shapes: every op is captured alone in a CUDA graph
and the graph is replayed back to back on one stream (each replay depends on the previous one), so the figure is
kernel latency + graph-node launch overhead, i.e. what the op costs on the critical path of a captured step."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mae_b200 import ops, synth, _lib  # noqa: E402

dev = torch.device("cuda")
lib = _lib.load()


REP = int(os.environ.get("REP", 40))


def chain(fn, n=20):
    """REP dependent copies of the op inside ONE graph (same stream -> serial kernel nodes); per-op time = replay / REP,
    i.e. kernel duration + the in-graph dependent-launch gap."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(REP):
            fn()
    for _ in range(3):
        g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(n):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n / REP * 1e3


def main():
    B, D = int(os.environ.get("B", 64)), 64
    mu, lv = synth.posterior_params(B, (D,), clamp=True)
    mu, lv = mu.to(dev), lv.to(dev)
    eps = synth.noise(B, 1, (D,)).to(dev)
    inp = synth.color_head(B, 1, 10, 32, 32)
    raw, x = inp["raw"].to(dev), inp["x"].to(dev)
    st = lambda: torch.cuda.current_stream().cuda_stream
    z = torch.empty_like(eps); kl = torch.empty(B, device=dev)
    ws = torch.zeros(lib.mae_mpd_workspace_bytes(B, D, 1), dtype=torch.uint8, device=dev)
    m_d = torch.empty(D, device=dev); sumsq = torch.empty(1, device=dev, dtype=torch.float64); S = torch.empty((), device=dev)
    dmu = torch.empty(B, D, device=dev); dlv = torch.empty(B, D, device=dev)
    g_m = torch.randn(D, device=dev); g_S = torch.ones((), device=dev)
    err = torch.empty(B, 1, device=dev); g = torch.full((B, 1), 1.0 / B, device=dev); draw = torch.empty_like(raw)
    wsd = torch.zeros(lib.mae_dmol_workspace_bytes(B, 1024), dtype=torch.uint8, device=dev)
    sc = torch.empty(6, device=dev); recon = torch.empty(B, device=dev)
    derr = torch.empty(B, 1, device=dev); dkl = torch.empty(B, device=dev); dm = torch.empty(D, device=dev); dS = torch.empty((), device=dev)
    P = lambda t: t.data_ptr()
    res = {}
    res["empty graph node (at::fill)"] = chain(lambda: dS.fill_(1.0))
    res["reparam_kl_fwd"] = chain(lambda: lib.mae_reparam_kl_fwd(P(mu), D, P(lv), D, P(eps), B, 1, D, P(z), P(kl), st()))
    res["reparam_kl_bwd"] = chain(lambda: lib.mae_reparam_kl_bwd(P(mu), D, P(lv), D, P(eps), P(z), P(kl), B, 1, D, P(dmu), P(dlv), 0, st()))
    res["mpd_prepare"] = chain(lambda: lib.mae_mpd_prepare(P(mu), D, P(lv), D, B, D, P(ws), ws.numel(), P(m_d), P(kl), st()))
    res["mpd_fwd(store_kl)"] = chain(lambda: lib.mae_mpd_fwd(P(ws), ws.numel(), B, D, 0, B, 1, P(sumsq), P(S), st()))
    res["mpd_bwd path1"] = chain(lambda: lib.mae_mpd_bwd(P(mu), D, P(lv), D, P(ws), ws.numel(), B, D, 0, B, P(g_m), P(g_S), P(S), P(kl), P(dmu), P(dlv), 0, 1, st()))
    res["mpd_bwd path2"] = chain(lambda: lib.mae_mpd_bwd(P(mu), D, P(lv), D, P(ws), ws.numel(), B, D, 0, B, P(g_m), P(g_S), P(S), P(kl), P(dmu), P(dlv), 0, 2, st()))
    res["dmol_fwd"] = chain(lambda: lib.mae_dmol_fwd(P(raw), P(x), B, 1, 10, 1024, P(err), P(wsd), wsd.numel(), st()))
    res["dmol_bwd"] = chain(lambda: lib.mae_dmol_bwd(P(raw), P(x), P(g), 1, 1.0, B, 1, 10, 1024, P(draw), st()))
    res["assemble_fwd"] = chain(lambda: lib.mae_loss_assemble_fwd(P(err), P(kl), P(m_d), P(S), None, B, 1, D, 0, 1, 1.0, 1.0, 0.0, 0.0, P(sc), P(recon), st()))
    res["assemble_bwd"] = chain(lambda: lib.mae_loss_assemble_bwd(P(g_S), P(sc), P(m_d), B, 1, D, 1.0, 1.0, 0.0, 0.0, P(derr), P(dkl), P(dm), P(dS), st()))
    if lib.mae_kl_mpd_loss_fused_supported(B, D):
        gr = torch.empty(2, B, D, device=dev)
        res["kl_mpd_loss_fused (fwd+bwd)"] = chain(lambda: lib.mae_kl_mpd_loss_fused(
            P(mu), D, P(lv), D, B, D, 1.0, 1.0, 0.0, 1.0 / B, P(m_d), P(S), P(kl), None, P(gr[0]), P(gr[1]), st()))
        res["kl_mpd_loss_fused (fwd only)"] = chain(lambda: lib.mae_kl_mpd_loss_fused(
            P(mu), D, P(lv), D, B, D, 1.0, 1.0, 0.0, 1.0 / B, P(m_d), P(S), P(kl), None, None, None, st()))
    for k, v in res.items():
        print("%-32s %8.2f us" % (k, v))
    if os.environ.get("MAE_DEBUG_TIMING") and lib.mae_kl_mpd_loss_fused_supported(B, D):
        import ctypes
        torch.cuda.synchronize()
        for _ in range(3):
            lib.mae_kl_mpd_loss_fused(P(mu), D, P(lv), D, B, D, 1.0, 1.0, 0.0, 1.0 / B, P(m_d), P(S), P(kl), None, P(gr[0]), P(gr[1]), st())
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * 16)()
        lib.mae_debug_fused_stamps.argtypes = [ctypes.c_void_p]
        lib.mae_debug_fused_stamps(ctypes.addressof(buf))
        t0 = buf[0]
        print("fused: loads+pivot partials %d | elements %d | stats, m_d, KL mean %d | KL rows + block sum %d | "
              "cluster exchange %d | gradient %d   (ns since start, CTA 0)" % tuple(int(buf[i]) - int(t0) for i in range(1, 7)))
    if os.environ.get("MAE_DEBUG_TIMING"):
        # phase timestamps written by the -DMAE_DEBUG_TIMING build (ns, %globaltimer), CTA 0 / last CTA
        torch.cuda.synchronize()
        lib.mae_mpd_prepare(P(mu), D, P(lv), D, B, D, P(ws), ws.numel(), P(m_d), P(kl), st())
        lib.mae_mpd_fwd(P(ws), ws.numel(), B, D, 0, B, 1, P(sumsq), P(S), st())
        torch.cuda.synchronize()
        dbg = ws[32:32 + 24 * 8].view(torch.int64).cpu().tolist()
        t0 = dbg[0]
        print("prepare: start 0 | elems done %d | rows published %d | last-CTA start %d | end %d  (ns)" %
              tuple(d - t0 for d in dbg[1:5]))
        t8 = dbg[8]
        print("fwd: start 0 | K loop done %d | epilogue done %d | finalize done %d  (ns)" % tuple(d - t8 for d in dbg[9:12]))
        print("prepare start -> fwd start: %d ns" % (t8 - t0))


if __name__ == "__main__":
    main()
