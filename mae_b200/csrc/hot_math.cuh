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

// exp(-|x|) and 1/(1+exp(-|x|)): the two quantities every sigmoid/softplus below is built from.
struct SigParts {
    float e;  // exp(-|x|)  in (0,1]
    float r;  // 1/(1+e)    in [0.5,1)
};
MAE_HD SigParts sig_parts(float x) {
    SigParts s;
    s.e = expf(-fabsf(x));
    s.r = 1.0f / (1.0f + s.e);
    return s;
}
MAE_HD float sigmoid_from(const SigParts& s, float x) { return x >= 0.f ? s.r : s.e * s.r; }   // sigma(x)
MAE_HD float sigmoid_neg_from(const SigParts& s, float x) { return x >= 0.f ? s.e * s.r : s.r; }  // sigma(-x) = 1-sigma(x)
// softplus(x) = log(1+e^x) = max(x,0) + log1p(e^{-|x|}); F.softplus' threshold-20 shortcut is the same
// number in fp32 (log1p(e^-20) = 2e-9 < ulp(20)/2).
MAE_HD float softplus_from(const SigParts& s, float x) { return fmaxf(x, 0.f) + log1pf(s.e); }

// log of the probability mass a logistic(mean, exp(ls)) puts on the bin of x, with the reference's
// four-way selection (utils.py:94-118).  cx = x - mean (centred), ls = clamped log-scale.
// When GRAD, also returns d/dcx and d/dls of the selected branch.
template <bool GRAD>
MAE_HD float logistic_bin_logprob(float cx, float ls, float x, float& d_cx, float& d_ls) {
    const float s = expf(-ls);
    const float hi = s * (cx + kBin);  // "plus_in"
    const float lo = s * (cx - kBin);  // "min_in"
    float L;
    if (x > kUpper) {  // x == 255: everything above the last bin edge, -softplus(min_in)
        const SigParts a = sig_parts(lo);
        L = -softplus_from(a, lo);
        if (GRAD) {
            const float dq = -sigmoid_from(a, lo);
            d_cx = s * dq;
            d_ls = -lo * dq;
        }
    } else if (x < kLower) {  // x == 0: everything below the first bin edge, plus_in - softplus(plus_in)
        const SigParts a = sig_parts(hi);
        L = hi - softplus_from(a, hi);
        if (GRAD) {
            const float dp = sigmoid_neg_from(a, hi);
            d_cx = s * dp;
            d_ls = -hi * dp;
        }
    } else {
        const SigParts a = sig_parts(hi), b = sig_parts(lo);
        // cdf(hi) - cdf(lo); evaluated through the complements when both arguments are positive so the
        // difference of two numbers near 1 does not cancel (hi > lo always).
        const float delta = lo > 0.f ? (sigmoid_neg_from(b, lo) - sigmoid_neg_from(a, hi))
                                     : (sigmoid_from(a, hi) - sigmoid_from(b, lo));
        if (delta > kDeltaMin) {
            L = logf(fmaxf(delta, kEpsF));
            if (GRAD) {
                const float inv = 1.0f / delta;
                const float dp = a.e * a.r * a.r * inv;    // sigma'(hi)/delta
                const float dq = -b.e * b.r * b.r * inv;   // -sigma'(lo)/delta
                d_cx = s * (dp + dq);
                d_ls = -(hi * dp + lo * dq);
            }
        } else {  // very wide component: density at the bin centre times the bin width
            const float u = s * cx;
            const SigParts c = sig_parts(u);
            L = u - ls - 2.0f * softplus_from(c, u) + kLog2Bin;
            if (GRAD) {
                const float du = 1.0f - 2.0f * sigmoid_from(c, u);
                d_cx = s * du;
                d_ls = -u * du - 1.0f;
            }
        }
    }
    return L;
}

// One mixture component of one pixel: sum over the three colour channels of the bin log-probability,
// with the linear colour coupling through the TRUE sub-pixels (utils.py:81-85) and the parameter
// post-processing of ColorImageDecoder.execute (tanh of coefficients, clamp of log-scales).
//   mu, rho, kap: raw means / log-scales / coefficients of this component; x: the 3 sub-pixels.
// When GRAD, g_mu/g_rho/g_kap receive d T / d raw for this component.
template <bool GRAD>
MAE_HD float dmol_component(const float mu[3], const float rho[3], const float kap[3], const float x[3],
                            float g_mu[3], float g_rho[3], float g_kap[3]) {
    const float c0 = tanhf(kap[0]), c1 = tanhf(kap[1]), c2 = tanhf(kap[2]);
    float mean[3];
    mean[0] = mu[0];
    mean[1] = mu[1] + c0 * x[0];
    mean[2] = mu[2] + c1 * x[0] + c2 * x[1];
    float T = 0.f;
    float d_cx[3], d_ls[3];
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
        g_kap[0] = x[0] * g_mu[1] * (1.0f - c0 * c0);
        g_kap[1] = x[0] * g_mu[2] * (1.0f - c1 * c1);
        g_kap[2] = x[1] * g_mu[2] * (1.0f - c2 * c2);
    }
    return T;
}

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
    MAE_HD void push(float v) {
        if (v <= m) {
            s += expf(v - m);
        } else {
            s = s * expf(m - v) + 1.0f;
            m = v;
        }
    }
    MAE_HD void merge(const Lse& o) {
        if (o.m == -INFINITY) return;
        if (m == -INFINITY) { m = o.m; s = o.s; return; }
        if (o.m <= m) s += o.s * expf(o.m - m);
        else { s = s * expf(m - o.m) + o.s; m = o.m; }
    }
    MAE_HD float value() const { return m + logf(s); }
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
        const float w = expf(T + logit - zt);  // responsibility of component m
        const float pi = expf(logit - zl);     // softmax(logits)_m
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
