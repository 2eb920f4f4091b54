// Per-element arithmetic of the likelihood heads, written once as __host__ __device__ functions so
// that the same source is (a) inlined into the sm_100a kernels and (b) compiled for the host by
// tests/hostcheck to validate the algebra (forward values and analytic gradients) against the
// oracle on a machine without a GPU.  The host build is a TEST of the formulas, not a fallback:
// libmae_b200.so exports no host compute path.
#pragma once

#include <math.h>

#if defined(__CUDACC__)
#define MAE_HD __host__ __device__ __forceinline__
#else
#define MAE_HD inline
#endif

namespace mae {
namespace hm {

constexpr float kEpsF = 1e-12f;
// bin_size = 1/255, lower = 1/255 - 1, upper = 1 - 1/255  (color_image_decoder.py:121-123)
constexpr float kBin = 1.0f / 255.0f;
constexpr float kLower = (float)(1.0 / 255.0 - 1.0);
constexpr float kUpper = (float)(1.0 - 1.0 / 255.0);
constexpr float kLog2Bin = -4.848116181284444f;  // log(2/255)            (utils.py:104)
constexpr float kDeltaMin = 1e-5f;               // cdf_delta threshold   (utils.py:113)
constexpr float kLogScaleMin = -7.0f;            // log_scale.clamp(min=-7) (color_image_decoder.py:52)

// ---- MUFU primitives -------------------------------------------------------------------------------
// On the device these are single SFU instructions (ex2/lg2/rcp.approx.ftz, <= 2^-22 relative error);
// the host versions exist only for tests/hostcheck.
#if defined(__CUDA_ARCH__)
MAE_HD float fast_ex2(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
MAE_HD float fast_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
MAE_HD float fast_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
#else
MAE_HD float fast_ex2(float x) { return exp2f(x); }
MAE_HD float fast_lg2(float x) { return log2f(x); }
MAE_HD float fast_rcp(float x) { return 1.0f / x; }
#endif
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

// log of the probability mass a logistic(mean, exp(ls)) puts on the bin of x, with the reference's
// four-way selection (utils.py:94-118).  cx = x - mean (centred), ls = clamped log-scale.
// When GRAD, also returns d/dcx and d/dls of the selected branch.
//
// Formulation (5 MUFU ops per call on the common path: 3 ex2, 1 rcp, 1 lg2; all selects are branch-free):
//   s = e^-ls, u = s*cx, hi = u + s/255 ("plus_in"), lo = u - s/255 ("min_in")
//   eh = e^-|hi|, el = e^-|lo|, a = 1+eh, b = 1+el, r = 1/(a*b)
//   cdf(hi)-cdf(lo) = num*r with num = el-eh (both > 0), eh-el (both < 0), 1-el*eh (opposite signs):
//     always a difference of SMALL numbers, so no cancellation near cdf = 1 (the reference loses digits there)
//   x == 255:  -softplus(lo)    = log(a*r) - max(lo,0)          x == 0:  hi - softplus(hi) = log(b*r) + min(hi,0)
//   so every non-"approx" case is  log(Q) + lin  with one lg2.
//   "approx" (delta <= 1e-5, rare): u - ls - 2*softplus(u) + log(2/255), taken as a real branch.
template <bool GRAD>
MAE_HD float logistic_bin_logprob(float cx, float ls, float x, float& d_cx, float& d_ls) {
    const float s = fast_ex2(-kLog2e * ls);
    const float sh = s * kBin;
    const float hi = fmaf(s, cx, sh), lo = fmaf(s, cx, -sh);
    const float eh = fast_ex2(-kLog2e * fabsf(hi)), el = fast_ex2(-kLog2e * fabsf(lo));
    const float a = 1.0f + eh, b = 1.0f + el;
    const float r = fast_rcp(a * b);
    // same sign (hi*lo > 0): |el - eh| ; opposite signs: 1 - el*eh
    const float num = (hi * lo > 0.f) ? fabsf(el - eh) : fmaf(-el, eh, 1.0f);
    const float delta = num * r;
    const bool is_up = x > kUpper, is_low = x < kLower;
    const float q = is_up ? a * r : (is_low ? b * r : fmaxf(delta, kEpsF));
    const float lin = is_up ? -fmaxf(lo, 0.f) : (is_low ? fminf(hi, 0.f) : 0.f);
    float L = fmaf(fast_lg2(q), kLn2, lin);
    if (GRAD) {
        // d/dhi and d/dlo of the selected branch
        const float rinv = fast_rcp(num);
        const float dp_mid = eh * b * b * r * rinv;   // sigma'(hi)/delta
        const float dq_mid = -el * a * a * r * rinv;  // -sigma'(lo)/delta
        const float dp = is_up ? 0.f : (is_low ? (hi >= 0.f ? eh : 1.0f) * b * r : dp_mid);   // low: sigma(-hi)
        const float dq = is_up ? -(lo >= 0.f ? 1.0f : el) * a * r : (is_low ? 0.f : dq_mid);  // up: -sigma(lo)
        d_cx = s * (dp + dq);
        d_ls = -(hi * dp + lo * dq);
    }
    if (!is_up && !is_low && !(delta > kDeltaMin)) {  // very wide / very far component: density * bin width
        const float u = s * cx;
        const float eu = fast_ex2(-kLog2e * fabsf(u));
        const float sp = fmaxf(u, 0.f) + kLn2 * fast_lg2(1.0f + eu);  // softplus(u)
        L = u - ls - 2.0f * sp + kLog2Bin;
        if (GRAD) {
            const float du = (u >= 0.f ? (eu - 1.0f) : (1.0f - eu)) * fast_rcp(1.0f + eu);  // 1 - 2*sigma(u)
            d_cx = s * du;
            d_ls = -u * du - 1.0f;
        }
    }
    return L;
}

// One mixture component of one pixel: sum over the three colour channels of the bin log-probability,
// with the linear colour coupling through the TRUE sub-pixels (utils.py:81-85) and the parameter
// post-processing of ColorImageDecoder.execute (tanh of coefficients, clamp of log-scales).
//   mu, rho, kap: raw means / log-scales / coefficients of this component; x: the 3 sub-pixels.
// When GRAD, g_mu/g_rho/g_kap receive d T / d raw for this component.
// tanh(k) = sign(k)(1-e)/(1+e), e = e^-2|k|; the three reciprocals share one rcp; 1-tanh^2 = 4e/(1+e)^2.
template <bool GRAD>
MAE_HD float dmol_component(const float mu[3], const float rho[3], const float kap[3], const float x[3],
                            float g_mu[3], float g_rho[3], float g_kap[3]) {
    const float e0 = fast_ex2(-2.0f * kLog2e * fabsf(kap[0]));
    const float e1 = fast_ex2(-2.0f * kLog2e * fabsf(kap[1]));
    const float e2 = fast_ex2(-2.0f * kLog2e * fabsf(kap[2]));
    const float p0 = 1.0f + e0, p1 = 1.0f + e1, p2 = 1.0f + e2;
    const float p01 = p0 * p1;
    const float r3 = fast_rcp(p01 * p2);
    const float i0 = r3 * p1 * p2, i1 = r3 * p0 * p2, i2 = r3 * p01;  // 1/(1+e_k)
    const float c0 = copysignf((1.0f - e0) * i0, kap[0]);
    const float c1 = copysignf((1.0f - e1) * i1, kap[1]);
    const float c2 = copysignf((1.0f - e2) * i2, kap[2]);
    float mean[3];
    mean[0] = mu[0];
    mean[1] = fmaf(c0, x[0], mu[1]);
    mean[2] = fmaf(c2, x[1], fmaf(c1, x[0], mu[2]));
    float T = 0.f;
    float d_cx[3], d_ls[3];
    if (!GRAD) {
        // forward only: the three channels share ONE reciprocal and ONE logarithm,
        //   sum_c [log(sel_c / den_c) + lin_c] = log(prod sel_c / prod den_c) + sum lin_c,
        // sel_c in (1e-5, 4], den_c = (1+e_hi)(1+e_lo) in [1, 4]: the products stay far inside the fp32 range.  Channels
        // on the rare "approx" branch drop out of the products (sel = den = 1) and contribute through lin.
        float sel_p = 1.0f, den_p = 1.0f;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const float ls = fmaxf(rho[c], kLogScaleMin);
            const float cx = x[c] - mean[c];
            const float s = fast_ex2(-kLog2e * ls);
            const float sh = s * kBin;
            const float hi = fmaf(s, cx, sh), lo = fmaf(s, cx, -sh);
            const float eh = fast_ex2(-kLog2e * fabsf(hi)), el = fast_ex2(-kLog2e * fabsf(lo));
            const float a = 1.0f + eh, b = 1.0f + el;
            const float den = a * b;
            const float num = (hi * lo > 0.f) ? fabsf(el - eh) : fmaf(-el, eh, 1.0f);
            const bool is_up = x[c] > kUpper, is_low = x[c] < kLower;
            float sel = is_up ? a : (is_low ? b : num);
            float dn = den;
            float lin = is_up ? -fmaxf(lo, 0.f) : (is_low ? fminf(hi, 0.f) : 0.f);
            if (!is_up && !is_low && !(num > kDeltaMin * den)) {   // delta = num / den <= 1e-5
                const float u = s * cx;
                const float eu = fast_ex2(-kLog2e * fabsf(u));
                lin = u - ls - 2.0f * (fmaxf(u, 0.f) + kLn2 * fast_lg2(1.0f + eu)) + kLog2Bin;
                sel = 1.0f;
                dn = 1.0f;
            }
            sel_p *= sel;
            den_p *= dn;
            T += lin;
        }
        return fmaf(fast_lg2(sel_p * fast_rcp(den_p)), kLn2, T);
    }
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float ls = fmaxf(rho[c], kLogScaleMin);
        T += logistic_bin_logprob<GRAD>(x[c] - mean[c], ls, x[c], d_cx[c], d_ls[c]);
    }
    if (GRAD) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            g_mu[c] = -d_cx[c];                                     // d/d mean = -d/d cx
            g_rho[c] = rho[c] >= kLogScaleMin ? d_ls[c] : 0.f;      // clamp passes the gradient at equality
        }
        g_kap[0] = x[0] * g_mu[1] * (4.0f * e0 * i0 * i0);
        g_kap[1] = x[0] * g_mu[2] * (4.0f * e1 * i1 * i1);
        g_kap[2] = x[1] * g_mu[2] * (4.0f * e2 * i2 * i2);
    }
    return T;
}

// ---- packed (two pixels at once) forward evaluation ------------------------------------------------------
// Blackwell issues FADD2/FMUL2/FFMA2 (two fp32 operations per lane per instruction, __f{add,mul,fma}2_rn).  The
// forward likelihood is issue-bound once its loads are wide, so the arithmetic of two neighbouring pixels is carried
// in float2 registers; only the SFU ops, |x| and the few compares stay scalar.  Edge bins (x = 0 / 255) are folded
// in with loop-invariant masks instead of selects:  Q = r * (fm*num + fu*(1+eh) + fl*(1+el)),
// lin = fl*min(hi,0) - fu*max(lo,0)  with (fu, fl, fm) = (x is 255, x is 0, neither).
#if !defined(__CUDACC__)
struct float2 {
    float x, y;
};
#endif
MAE_HD float2 mk2(float a, float b) {
    float2 r;
    r.x = a;
    r.y = b;
    return r;
}
MAE_HD float2 splat2(float a) { return mk2(a, a); }
#if defined(__CUDA_ARCH__)
MAE_HD float2 add2(float2 a, float2 b) { return __fadd2_rn(a, b); }
MAE_HD float2 mul2(float2 a, float2 b) { return __fmul2_rn(a, b); }
MAE_HD float2 fma2(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
#else
MAE_HD float2 add2(float2 a, float2 b) { return mk2(a.x + b.x, a.y + b.y); }
MAE_HD float2 mul2(float2 a, float2 b) { return mk2(a.x * b.x, a.y * b.y); }
MAE_HD float2 fma2(float2 a, float2 b, float2 c) { return mk2(fmaf(a.x, b.x, c.x), fmaf(a.y, b.y, c.y)); }
#endif
MAE_HD float2 ex2_negabs2(float2 v, float scale) {  // exp2(-scale*|v|), scale > 0
    return mk2(fast_ex2(-scale * fabsf(v.x)), fast_ex2(-scale * fabsf(v.y)));
}

struct EdgeMask2 {  // loop-invariant per (pixel pair, channel)
    float2 fu, fl, fm;
};
MAE_HD EdgeMask2 edge_mask2(float2 x) {
    EdgeMask2 e;
    e.fu = mk2(x.x > kUpper ? 1.f : 0.f, x.y > kUpper ? 1.f : 0.f);
    e.fl = mk2(x.x < kLower ? 1.f : 0.f, x.y < kLower ? 1.f : 0.f);
    e.fm = fma2(add2(e.fu, e.fl), splat2(-1.0f), splat2(1.0f));
    return e;
}

// rare fix-up: the "approx" branch of one pixel-channel (delta <= 1e-5 and not an edge bin)
MAE_HD float logistic_approx_logprob(float u, float ls) {
    const float eu = fast_ex2(-kLog2e * fabsf(u));
    return u - ls - 2.0f * (fmaxf(u, 0.f) + kLn2 * fast_lg2(1.0f + eu)) + kLog2Bin;
}

// bin log-probability of one channel for two pixels (forward only); cx = x - mean, rho = RAW log-scale
MAE_HD float2 logistic_bin_logprob2(float2 cx, float2 rho, float2 x) {
    const EdgeMask2 em = edge_mask2(x);   // recomputed per call (2 FSET each): cheaper than 36 live mask registers
    const float2 ls = mk2(fmaxf(rho.x, kLogScaleMin), fmaxf(rho.y, kLogScaleMin));
    const float2 t = mul2(ls, splat2(-kLog2e));
    const float2 s = mk2(fast_ex2(t.x), fast_ex2(t.y));
    const float2 sh = mul2(s, splat2(kBin));
    const float2 hi = fma2(s, cx, sh);
    const float2 lo = fma2(sh, splat2(-2.0f), hi);
    const float2 eh = ex2_negabs2(hi, kLog2e), el = ex2_negabs2(lo, kLog2e);
    const float2 a = add2(eh, splat2(1.0f)), b = add2(el, splat2(1.0f));
    const float2 ab = mul2(a, b);
    const float2 r = mk2(fast_rcp(ab.x), fast_rcp(ab.y));
    const float2 d = fma2(eh, splat2(-1.0f), el);                       // el - eh
    const float2 mix = fma2(mul2(el, eh), splat2(-1.0f), splat2(1.0f));  // 1 - el*eh
    const float2 hl = mul2(hi, lo);
    const float2 num = mk2(hl.x > 0.f ? fabsf(d.x) : mix.x, hl.y > 0.f ? fabsf(d.y) : mix.y);
    const float2 sel = fma2(em.fl, b, fma2(em.fu, a, mul2(em.fm, num)));
    float2 q = mul2(sel, r);
    q = mk2(fmaxf(q.x, kEpsF), fmaxf(q.y, kEpsF));
    const float2 lin = fma2(em.fu, mk2(-fmaxf(lo.x, 0.f), -fmaxf(lo.y, 0.f)), mul2(em.fl, mk2(fminf(hi.x, 0.f), fminf(hi.y, 0.f))));
    float2 L = fma2(mk2(fast_lg2(q.x), fast_lg2(q.y)), splat2(kLn2), lin);
    if (em.fm.x > 0.f && !(q.x > kDeltaMin)) L.x = logistic_approx_logprob(hi.x - sh.x, ls.x);
    if (em.fm.y > 0.f && !(q.y > kDeltaMin)) L.y = logistic_approx_logprob(hi.y - sh.y, ls.y);
    return L;
}

// one mixture component for two pixels: T.x, T.y
MAE_HD float2 dmol_component2(const float2 mu[3], const float2 rho[3], const float2 kap[3], const float2 x[3]) {
    const float2 e0 = ex2_negabs2(kap[0], 2.0f * kLog2e), e1 = ex2_negabs2(kap[1], 2.0f * kLog2e),
                 e2 = ex2_negabs2(kap[2], 2.0f * kLog2e);
    const float2 one = splat2(1.0f);
    const float2 p0 = add2(e0, one), p1 = add2(e1, one), p2 = add2(e2, one);
    const float2 p01 = mul2(p0, p1), p12 = mul2(p1, p2);
    const float2 p012 = mul2(p01, p2);
    const float2 r3 = mk2(fast_rcp(p012.x), fast_rcp(p012.y));
    const float2 i0 = mul2(r3, p12), i1 = mul2(mul2(r3, p0), p2), i2 = mul2(r3, p01);
    const float2 m0 = mul2(fma2(e0, splat2(-1.0f), one), i0), m1 = mul2(fma2(e1, splat2(-1.0f), one), i1),
                 m2 = mul2(fma2(e2, splat2(-1.0f), one), i2);
    const float2 c0 = mk2(copysignf(m0.x, kap[0].x), copysignf(m0.y, kap[0].y));
    const float2 c1 = mk2(copysignf(m1.x, kap[1].x), copysignf(m1.y, kap[1].y));
    const float2 c2 = mk2(copysignf(m2.x, kap[2].x), copysignf(m2.y, kap[2].y));
    const float2 mone = splat2(-1.0f);
    // cx = x - mean, mean0 = mu0, mean1 = mu1 + c0 x0, mean2 = mu2 + c1 x0 + c2 x1
    const float2 cx0 = fma2(mu[0], mone, x[0]);
    const float2 cx1 = fma2(fma2(c0, x[0], mu[1]), mone, x[1]);
    const float2 cx2 = fma2(fma2(c2, x[1], fma2(c1, x[0], mu[2])), mone, x[2]);
    float2 T = logistic_bin_logprob2(cx0, rho[0], x[0]);
    T = add2(T, logistic_bin_logprob2(cx1, rho[1], x[1]));
    T = add2(T, logistic_bin_logprob2(cx2, rho[2], x[2]));
    return T;
}

// online log-sum-exp for two pixels
struct Lse2 {
    float2 m, s;
    MAE_HD void init() { m = splat2(-INFINITY); s = splat2(0.f); }
    MAE_HD void push(float2 v) {
        const float2 dv = mk2(v.x - m.x, v.y - m.y);   // (-inf) - (-inf) never occurs: v is finite
        const float2 e = ex2_negabs2(dv, kLog2e);
        const float2 grow = fma2(s, e, splat2(1.0f)), keep = add2(s, e);
        s = mk2(v.x <= m.x ? keep.x : grow.x, v.y <= m.y ? keep.y : grow.y);
        m = mk2(fmaxf(m.x, v.x), fmaxf(m.y, v.y));
    }
    MAE_HD float2 value() const { return fma2(mk2(fast_lg2(s.x), fast_lg2(s.y)), splat2(kLn2), m); }
};

// Bernoulli pixel term and its derivative w.r.t. the probability (binary_image_decoder.py:91).
// "1.0 - p + eps" is evaluated left to right like the reference.
MAE_HD float bernoulli_logprob(float p, float x) {
    return logf(p + kEpsF) * x + logf(1.0f - p + kEpsF) * (1.0f - x);
}
MAE_HD float bernoulli_derr_dp(float p, float x) {  // d(-logprob)/dp
    return (1.0f - x) / (1.0f - p + kEpsF) - x / (p + kEpsF);
}

// Online log-sum-exp accumulator.
struct Lse {
    float m, s;
    MAE_HD void init() { m = -INFINITY; s = 0.f; }
    MAE_HD void push(float v) {   // branch-free, one SFU op: e = exp(-|v-m|); s = v<=m ? s+e : s*e+1
        const float e = fast_ex2(-kLog2e * fabsf(v - m));
        s = (v <= m) ? (s + e) : fmaf(s, e, 1.0f);
        m = fmaxf(m, v);
    }
    MAE_HD void merge(const Lse& o) {
        if (o.m == -INFINITY) return;
        if (m == -INFINITY) { m = o.m; s = o.s; return; }
        if (o.m <= m) s += o.s * expf(o.m - m);
        else { s = s * expf(m - o.m) + o.s; m = o.m; }
    }
    MAE_HD float value() const { return fmaf(fast_lg2(s), kLn2, m); }
};


// ---- whole-pixel routines shared by the kernels and the host check --------------------------------
// `base` points at channel 0 of this pixel; channel c lives at base[c*cs].  Ld::ld is the load policy.
struct LdPlain {
    static MAE_HD float ld(const float* p) { return *p; }
};

template <class Ld>
MAE_HD void dmol_load_component(const float* base, long long cs, int nmix, int m, float& logit, float mu[3], float rho[3],
                                float kap[3]) {
    logit = Ld::ld(base + (long long)m * cs);
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const long long o = (long long)(3 * m + c) * cs;
        mu[c] = Ld::ld(base + (long long)nmix * cs + o);
        rho[c] = Ld::ld(base + (long long)4 * nmix * cs + o);
        kap[c] = Ld::ld(base + (long long)7 * nmix * cs + o);
    }
}

// returns the pixel's contribution to err:  LSE_m(logit_m) - LSE_m(T_m + logit_m);  optionally both LSEs.
template <class Ld>
MAE_HD float dmol_pixel_fwd(const float* base, long long cs, int nmix, const float xs[3], float* lse_t_out,
                            float* lse_l_out) {
    Lse lt, ll;
    lt.init();
    ll.init();
    for (int m = 0; m < nmix; ++m) {
        float logit, mu[3], rho[3], kap[3];
        dmol_load_component<Ld>(base, cs, nmix, m, logit, mu, rho, kap);
        const float T = dmol_component<false>(mu, rho, kap, xs, nullptr, nullptr, nullptr);
        lt.push(T + logit);
        ll.push(logit);
    }
    const float zt = lt.value(), zl = ll.value();
    if (lse_t_out) *lse_t_out = zt;
    if (lse_l_out) *lse_l_out = zl;
    return zl - zt;
}

// writes g * d(pixel err)/d raw for all 10*nmix channels of this pixel through St::st.
struct StPlain {
    static MAE_HD void st(float* p, float v) { *p = v; }
};
template <class Ld, class St>
MAE_HD void dmol_pixel_bwd(const float* base, float* out, long long cs, int nmix, const float xs[3], float g) {
    float zt, zl;
    dmol_pixel_fwd<Ld>(base, cs, nmix, xs, &zt, &zl);
    for (int m = 0; m < nmix; ++m) {
        float logit, mu[3], rho[3], kap[3], gmu[3], grho[3], gkap[3];
        dmol_load_component<Ld>(base, cs, nmix, m, logit, mu, rho, kap);
        const float T = dmol_component<true>(mu, rho, kap, xs, gmu, grho, gkap);
        const float w = fast_ex2(kLog2e * (T + logit - zt));  // responsibility of component m
        const float pi = fast_ex2(kLog2e * (logit - zl));     // softmax(logits)_m
        const float a = -g * w;                // d err / d T_m
        St::st(out + (long long)m * cs, g * (pi - w));
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const long long o = (long long)(3 * m + c) * cs;
            St::st(out + (long long)nmix * cs + o, a * gmu[c]);
            St::st(out + (long long)4 * nmix * cs + o, a * grho[c]);
            St::st(out + (long long)7 * nmix * cs + o, a * gkap[c]);
        }
    }
}

}  // namespace hm
}  // namespace mae
