"""Guard-band check of every buffer the general regulariser chain writes (B=4, D=128): out-of-bounds writes show up as
changed sentinels."""
import sys, torch
sys.path.insert(0, ".")
from mae_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda")
torch.manual_seed(0)
b, d = int(sys.argv[1]) if len(sys.argv) > 1 else 4, int(sys.argv[2]) if len(sys.argv) > 2 else 128
PAD = 4096
SENT = 12345.678
def guarded(n, dtype=torch.float32):
    buf = torch.full((n + 2 * PAD,), 57 if dtype == torch.uint8 else SENT, device=dev, dtype=dtype)
    return buf, buf[PAD:PAD + n]
def check(name, buf, n):
    lo, hi = buf[:PAD], buf[PAD + n:]
    bad = int((lo != SENT).sum()) + int((hi != SENT).sum())
    if bad:
        il = (lo != SENT).nonzero().flatten().tolist()[:5]; ih = (hi != SENT).nonzero().flatten().tolist()[:5]
        print("OOB WRITE near %s: %d sentinel(s) changed; below idx %s (of %d), above idx %s" % (name, bad, il, PAD, ih))
    return bad
mu_b, mu = guarded(b * d); lv_b, lv = guarded(b * d)
mu.copy_(torch.randn(b * d, device=dev)); lv.copy_(torch.randn(b * d, device=dev) * 0.3)
store = 1
nbytes = lib.mae_mpd_workspace_bytes(b, d, store)
ws_b, ws = guarded(nbytes, torch.uint8)
ws.zero_()
ws_b8 = ws_b  # uint8 sentinel is SENT cast -> compare separately
sent8 = 57
md_b, m_d = guarded(d); s_b, s_ = guarded(1); kl_b, kl = guarded(b); reg_b, reg = guarded(1)
ss_b = torch.full((2 * PAD + 1,), SENT, device=dev, dtype=torch.float64); sumsq = ss_b[PAD:PAD + 1]
g_b, grads = guarded(2 * b * d)
st = torch.cuda.current_stream().cuda_stream
_lib.check(lib.mae_mpd_prepare(mu.data_ptr(), d, lv.data_ptr(), d, b, d, ws.data_ptr(), nbytes, m_d.data_ptr(), kl.data_ptr(), st), "prepare")
_lib.check(lib.mae_mpd_fwd(ws.data_ptr(), nbytes, b, d, 0, b, store, sumsq.data_ptr(), s_.data_ptr(), st), "fwd")
_lib.check(lib.mae_mpd_loss_bwd(mu.data_ptr(), d, lv.data_ptr(), d, ws.data_ptr(), nbytes, b, d, 0, b, 0.5, 1.0, 0.0, 1.0 / b,
                                m_d.data_ptr(), s_.data_ptr(), kl.data_ptr(), reg.data_ptr(), grads.data_ptr(), grads.data_ptr() + 4 * b * d, 1, st), "loss_bwd")
torch.cuda.synchronize()
bad = 0
for name, buf, n in (("mu", mu_b, b * d), ("logvar", lv_b, b * d), ("m_d", md_b, d), ("S", s_b, 1), ("kl", kl_b, b), ("reg", reg_b, 1), ("grads", g_b, 2 * b * d)):
    bad += check(name, buf, n)
lo, hi = ws_b[:PAD], ws_b[PAD + nbytes:]
nb = int((lo != sent8).sum()) + int((hi != sent8).sum())
if nb: print("OOB WRITE near workspace: %d bytes changed (below %s, above %s)" % (nb, (lo != sent8).nonzero().flatten().tolist()[:5], (hi != sent8).nonzero().flatten().tolist()[:5]))
bad += nb
l2, h2 = ss_b[:PAD], ss_b[PAD + 1:]
nb = int((l2 != SENT).sum()) + int((h2 != SENT).sum())
if nb: print("OOB WRITE near sumsq")
bad += nb
print("B=%d D=%d: %s" % (b, d, "no out-of-bounds writes within +-%d elements" % PAD if bad == 0 else "%d guard entries changed" % bad))
