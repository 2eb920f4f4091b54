"""One MPD forward+backward at a C5 shape (profiler target): python scripts/dev_mpd_once.py B D [tc]"""
import sys
import torch
sys.path.insert(0, ".")
from mae_b200 import _lib, ops, synth
b, d = int(sys.argv[1]), int(sys.argv[2])
_lib.load().mae_set_mpd_tensor_cores(int(sys.argv[3]) if len(sys.argv) > 3 and sys.argv[3].isdigit() else 1)
mu, lv = synth.posterior_params(b, (d,), seed=5)
mu, lv = mu.cuda().requires_grad_(), lv.cuda().requires_grad_()
fwd_only = "--fwd" in sys.argv           # forward without the KL store (evaluation / recompute-backward callers)
for _ in range(2):
    mu.grad = lv.grad = None
    if fwd_only:
        with torch.no_grad():
            m_d, s = ops.mpd(mu.detach(), lv.detach())
        continue
    m_d, s = ops.mpd(mu, lv)
    (-torch.nn.functional.logsigmoid(m_d).sum() + s).backward()
torch.cuda.synchronize()
print("done", float(s))
import ctypes, os
if os.environ.get("MAE_MPD_PROF"):
    lib = _lib.load()
    out = (ctypes.c_longlong * 64)()
    lib.mae_debug_mpd_prof(out)
    ntile = ((b + 127) // 128) ** 2
    per = -(-ntile // 148)
    names = {0: "epilogue wg0 (wait tfull)", 4: "epilogue wg1", 8: "A producer wg0 (wait aempty)", 12: "A producer wg1", 16: "loader (wait bempty)", 20: "mma issuer (wait tempty, wait afull+bfull)"}
    for k, nm in names.items():
        tot = max(out[k], 1)
        print("%-44s total %9d cyc (%6.0f / tile)  wait0 %5.1f %%  wait1 %5.1f %%" % (nm, out[k], out[k] / per, 100.0 * out[k + 1] / tot, 100.0 * out[k + 2] / tot))
    print("   of the mma issuer's wait1, waiting for B (bulk copies): %.1f %% of its loop" % (100.0 * out[31] / max(out[20], 1)))
    nchunk = ((b + 127) // 128) * 4
    print("backward CTA (0,0,0): chunks of 32 partners =", nchunk, "(divided by the partner splits)")
    for k, nm in {32: "W1 rows (wait wempty, wait gfull)", 36: "W2 rows", 40: "loader (wait pempty)", 44: "mma issuer (wait w1+w2 full, wait pfull, wait gdrained)"}.items():
        tot = max(out[k], 1)
        print("%-58s total %9d cyc  wait0 %5.1f %%  wait1 %5.1f %%  wait2 %5.1f %%" % (nm, out[k], 100.0 * out[k + 1] / tot, 100.0 * out[k + 2] / tot, 100.0 * out[k + 3] / tot))
    print("weight warps of pair 0, cycles per own chunk: W1 loads+first-half split %.0f, second half+TMEM stores+arrive %.0f;  W2 %.0f, %.0f"
          % tuple(out[48 + k] / max(1, nchunk // 8 // 2) for k in range(4)))
    print("   of the second part: second-half split + store issue %.0f, tcgen05.wait::st %.0f (W1);  %.0f, %.0f (W2)"
          % tuple(out[52 + k] / max(1, nchunk // 8 // 2) for k in range(4)))
