"""CPU check (fp64, torch) of the closed forms the MPD kernels implement (mae_b200/csrc/mpd.cu):
pivoted GEMM decomposition of KL_ij, O(B*D) formula for m_d, and the own-row backward (row side + column
side + rank-1 mean part) in both the tile form and the stored-KL per-pair form.  Compared with autograd
through the oracle.  This validates the algebra only; the CUDA code itself is checked by
tests/test_gpu_parity.py on a B200."""
import torch

from oracle import mae_oracle as O


def kernel_algebra(mu, lv, g_m, g_S, npivot=32):
    B, D = mu.shape
    eps = 1e-12
    P = min(B, npivot)
    mu_bar = mu[:P].mean(0, keepdim=True)            # pivots: mean of the first P rows
    l_bar = lv[:P].mean(0, keepdim=True)
    v_bar = l_bar.exp()
    ivb = 1.0 / v_bar
    mc = mu - mu_bar
    alpha = lv - l_bar
    a = torch.expm1(alpha)                           # v_i / v_bar - 1
    v = v_bar * (1 + a)
    r = 1.0 / (v + eps)
    b = -(a * v_bar + eps) * r                       # v_bar * r_j - 1
    s = mc * mc * ivb
    t = a + (2 * b + b * b) * (1 + a)                # (1+b)^2 (1+a) - 1 = r^2 v v_bar - 1
    X = torch.stack([a + s, mc], dim=2).reshape(B, 2 * D)            # interleaved (2d, 2d+1)
    Y = torch.stack([b, -2.0 * mc * r], dim=2).reshape(B, 2 * D)
    ema = a - alpha                                        # kernel: series for small alpha
    bpa = (v_bar * (alpha * a - ema) + eps * (alpha - 1)) * r   # = b + alpha without cancellation
    assert torch.allclose(bpa, b + alpha, rtol=1e-9, atol=1e-12)
    E = (ema + s).sum(1)
    F = (bpa + mc * mc * r).sum(1)
    KL = 0.5 * (X @ Y.t() + E[:, None] + F[None, :])
    # column statistics (all in the pivot-normalised quantities a, b: no "large minus large") and closed-form m_d
    SA, SB, SAB, SApB = a.sum(0), b.sum(0), (a * b).sum(0), ((v_bar * a * a + eps * (a - 1)) * r).sum(0)   # last = sum(a+b)
    assert torch.allclose(SApB, (a + b).sum(0), rtol=1e-9, atol=1e-12)
    S1, S2, R1, R2 = mc.sum(0), (mc * mc).sum(0), (r * mc).sum(0), (r * mc * mc).sum(0)
    R0 = (B + SB) * ivb.squeeze(0)
    # sum_{i!=j} (v_i r_j - 1) = SA*SB - SAB + (B-1)*sum(a+b);  sum_{i!=j} (mu_i-mu_j)^2 r_j = S2*R0 - 2*S1*R1 + B*R2
    tn2 = SA * SB - SAB + (B - 1) * SApB + (S2 * R0 - 2 * S1 * R1 + B * R2)
    m_d = 0.5 * tn2 / (B * (B - 1))
    m = m_d.sum()
    off = ~torch.eye(B, dtype=torch.bool)
    sumsq = ((KL - m)[off] ** 2).sum()
    S = torch.sqrt(sumsq / (B * (B - 1) - 1))
    # backward, tile form
    beta = g_S / ((B * B - B - 1) * S)
    W1 = beta * (KL - m) * off          # weight of KL_ij, row i
    W2 = beta * (KL.t() - m) * off      # weight of KL_ji seen from row i
    G1 = W1 @ Y
    G2 = W2 @ X
    rs1, rs2 = W1.sum(1, keepdim=True), W2.sum(1, keepdim=True)
    g1r, g1m = G1[:, 0::2], G1[:, 1::2]
    g2x, g2m = G2[:, 0::2], G2[:, 1::2]
    ri = (1 + b) * ivb
    gmu = mc * (g1r + rs1) * ivb + 0.5 * g1m - ri * (g2m - mc * rs2)
    glv = 0.5 * ((1 + a) * g1r + a * rs1) + 0.5 * (-t * rs2 - (1 + t) * (g2x - (2 * mc * g2m - mc * mc * rs2) * ivb))
    aa = g_m / (B * (B - 1))
    dT_dmu = 2 * ((mc * R0 - R1) + ri * (B * mc - S1))
    dT_dlv = ((B - 1) * (a - t) + (1 + a) * (SB - b) - (1 + t) * (SA - a)
              - (1 + t) * ivb * (S2 - 2 * S1 * mc + B * mc * mc))
    gmu = gmu + 0.5 * aa * dT_dmu
    glv = glv + 0.5 * aa * dT_dlv
    # per-pair form used by the stored-KL kernel
    dl = mc[:, None, :] - mc[None, :, :]
    ai, aj, bi, bj, ti = a[:, None], a[None], b[:, None], b[None], t[:, None]
    w1, w2 = W1[:, :, None], W2[:, :, None]
    gmu2 = (dl * (w1 * (1 + bj) + w2 * (1 + bi)) * ivb).sum(1)
    glv2 = 0.5 * (w1 * (ai * bj + ai + bj) + w2 * (-(aj + ti + aj * ti) - dl * dl * (1 + ti) * ivb)).sum(1)
    gmu2 = gmu2 + 0.5 * aa * dT_dmu
    glv2 = glv2 + 0.5 * aa * dT_dlv
    return m_d, S, gmu, glv, gmu2, glv2


def test_mpd_kernel_algebra_matches_oracle_autograd():
    torch.manual_seed(0)
    for B, D, off in [(2, 3, 0.0), (5, 1, 0.0), (9, 4, 20.0), (33, 17, -3.0), (40, 6, 0.5)]:
        mu = (torch.randn(B, D, dtype=torch.float64) + off).requires_grad_()
        lv = torch.randn(B, D, dtype=torch.float64).clamp(-7, 7).requires_grad_()
        g_m = torch.randn(D, dtype=torch.float64)
        g_S = torch.tensor(0.7, dtype=torch.float64)
        m_d_ref, S_ref = O.post_kl(mu, lv)
        ((m_d_ref * g_m).sum() + g_S * S_ref).backward()
        with torch.no_grad():
            m_d, S, gmu, glv, gmu2, glv2 = kernel_algebra(mu.detach(), lv.detach(), g_m, g_S)
        for got, ref, name in [(m_d, m_d_ref, "m_d"), (S, S_ref, "S"), (gmu, mu.grad, "dmu"), (glv, lv.grad, "dlv"),
                               (gmu2, mu.grad, "dmu direct"), (glv2, lv.grad, "dlv direct")]:
            err = (got - ref.detach()).abs().max() / ref.detach().abs().max().clamp_min(1e-300)
            assert err < 1e-9, (B, D, name, float(err))


def cluster_kernel_algebra(mu, lv, eta, gamma, free_bits, npivot=32):
    """fp64 restatement of mpd_fused.cu: direct pair form without the K = 2D panels (both orientations share the squared
    difference), factored gradient terms, upstream gradients derived from the closed-form consumers (mae.py:132-139), the
    analytic-KL part with the free-bits gate, (a, b) -> (r, a+b) reconstruction used when V is rebuilt from (a, b)."""
    B, D = mu.shape
    eps = 1e-12
    P = min(B, npivot)
    mu_bar, l_bar = mu[:P].mean(0, keepdim=True), lv[:P].mean(0, keepdim=True)
    v_bar = l_bar.exp()
    ivb = 1.0 / v_bar
    mc, alpha = mu - mu_bar, lv - l_bar
    a = torch.expm1(alpha)
    v = v_bar * (1 + a)
    r = 1.0 / (v + eps)
    b = -(a * v_bar + eps) * r
    assert torch.allclose((1 + b) * ivb, r, rtol=1e-12)                        # r rebuilt from b
    apb = (v_bar * a * a + eps * (a - 1)) * r
    t = a + (2 * b + b * b) * (1 + a)
    ema = a - alpha
    bpa = (v_bar * (alpha * a - ema) + eps * (alpha - 1)) * r
    E, F = ema.sum(1), bpa.sum(1)                                              # no mc^2 terms: the pair loop carries them
    dl = mc[:, None, :] - mc[None, :, :]
    q2 = dl * dl
    KL = 0.5 * ((a[:, None] * b[None] + q2 * r[None]).sum(2) + E[:, None] + F[None, :])         # KL_ij
    KLT = 0.5 * ((a[None] * b[:, None] + q2 * r[:, None]).sum(2) + E[None, :] + F[:, None])      # KL_ji seen from row i
    assert torch.allclose(KLT, KL.t(), rtol=1e-10, atol=1e-12)
    SA, SB, SAB, SApB = a.sum(0), b.sum(0), (a * b).sum(0), apb.sum(0)
    S1, S2, R1, R2 = mc.sum(0), (mc * mc).sum(0), (r * mc).sum(0), (r * mc * mc).sum(0)
    R0 = (B + SB) * ivb.squeeze(0)
    m_d = 0.5 * (SA * SB - SAB + (B - 1) * SApB + (S2 * R0 - 2 * S1 * R1 + B * R2)) / (B * (B - 1))
    m = m_d.sum()
    off = ~torch.eye(B, dtype=torch.bool)
    S = torch.sqrt(((KL - m)[off] ** 2).sum() / (B * (B - 1) - 1))
    kl = 0.5 * (mu * mu + torch.expm1(lv) - lv).sum(1)
    mean_kl = kl.sum() / B
    reg = torch.clamp(mean_kl, min=free_bits) - eta * torch.nn.functional.logsigmoid(m_d).sum() + gamma * S
    # upstream gradients of the closed-form consumers
    g_m = -eta * torch.sigmoid(-m_d)
    gk = (1.0 / B) if float(mean_kl) >= free_bits else 0.0
    beta = gamma / ((B * B - B - 1) * S)
    w1 = (beta * (KL - m) * off)[:, :, None]
    w2 = (beta * (KLT - m) * off)[:, :, None]
    ai, aj, bj, ri, rj, ti = a[:, None], a[None], b[None], r[:, None], r[None], t[:, None]
    oi = (1 + ti) * ivb
    gmu = (dl * (w1 * rj + w2 * ri)).sum(1)                                                       # r carries 1/v_bar
    glv = 0.5 * (w1 * ((ai + 1) * bj + ai) - w2 * ((ti + 1) * aj + (q2 * oi + ti))).sum(1)        # factored forms
    aa = g_m / (B * (B - 1))
    dT_dmu = 2 * ((mc * R0 - R1) + r * (B * mc - S1))
    dT_dlv = ((B - 1) * (a - t) + (1 + a) * (SB - b) - (1 + t) * (SA - a) - (1 + t) * ivb * (S2 - 2 * S1 * mc + B * mc * mc))
    gmu = gmu + 0.5 * aa * dT_dmu + gk * mu
    glv = glv + 0.5 * aa * dT_dlv + 0.5 * gk * torch.expm1(lv)
    return kl, m_d, S, reg, gmu, glv


def test_cluster_kernel_algebra_matches_oracle_autograd():
    torch.manual_seed(1)
    for B, D, off, eta, gamma, fb in [(2, 3, 0.0, 1.0, 1.0, 0.0), (9, 4, 20.0, 0.5, 2.0, 0.0), (33, 17, -3.0, 1.0, 0.0, 0.0),
                                      (40, 6, 0.5, 0.0, 1.0, 1e6), (100, 8, 0.0, 1.0, 1.0, 0.0)]:
        mu = (torch.randn(B, D, dtype=torch.float64) + off).requires_grad_()
        lv = torch.randn(B, D, dtype=torch.float64).clamp(-7, 7).requires_grad_()
        m_ref, s_ref = O.post_kl(mu, lv)
        kl_ref = O.analytic_kl(mu, lv)
        reg_ref = kl_ref.mean().clamp(min=fb) - eta * torch.nn.functional.logsigmoid(m_ref).sum() + gamma * s_ref
        reg_ref.backward()
        with torch.no_grad():
            kl, m_d, S, reg, gmu, glv = cluster_kernel_algebra(mu.detach(), lv.detach(), eta, gamma, fb)
        for got, ref, name in [(kl, kl_ref, "KL"), (m_d, m_ref, "m_d"), (S, s_ref, "S"), (reg, reg_ref, "reg"),
                               (gmu, mu.grad, "dmu"), (glv, lv.grad, "dlv")]:
            err = (got - ref.detach()).abs().max() / ref.detach().abs().max().clamp_min(1e-300)
            assert err < 1e-9, (B, D, name, float(err))
