// Per-element mathematics of the pivoted MPD decomposition (see the header comment of mpd.cu), shared by the general
// kernels (mpd.cu) and the fused small-batch kernel (mpd_fused.cu).
#pragma once

#include "common.cuh"

namespace mae {

// em1 = e^x - 1 and ema = e^x - 1 - x >= 0, both without cancellation and with ONE exponential at most: Taylor series of
// ema for |x| < 0.5 (truncation < 2e-9 relative; em1 = x + ema), expf(x) - 1 otherwise (no cancellation there).
__device__ __forceinline__ void expm1_pair(float x, float& em1, float& ema) {
    if (fabsf(x) < 0.5f) {
        float p = 1.0f / 362880.0f;
        p = fmaf(p, x, 1.0f / 40320.0f);
        p = fmaf(p, x, 1.0f / 5040.0f);
        p = fmaf(p, x, 1.0f / 720.0f);
        p = fmaf(p, x, 1.0f / 120.0f);
        p = fmaf(p, x, 1.0f / 24.0f);
        p = fmaf(p, x, 1.0f / 6.0f);
        p = fmaf(p, x, 0.5f);
        ema = p * x * x;
        em1 = x + ema;
    } else {
        em1 = expf(x) - 1.0f;
        ema = em1 - x;
    }
}
__device__ __forceinline__ float expm1_minus_x(float x) {
    float em1, ema;
    expm1_pair(x, em1, ema);
    return ema;
}

// per-element quantities of the pivoted decomposition.  fp32 throughout: every quantity below is either a
// product/quotient of well-conditioned terms or evaluated through a cancellation-free form, so fp32 keeps full
// relative precision even when all posteriors coincide (fp64 transcendentals cost ~10 us of pure latency at B = 64).
struct Elem {
    float mc, alpha, a, ema, v, r, b, bpa, apb;
};
__device__ __forceinline__ Elem make_elem(float mu, float lv, float mu_bar, float l_bar, float v_bar) {
    Elem e;
    e.mc = mu - mu_bar;
    e.alpha = lv - l_bar;
    expm1_pair(e.alpha, e.a, e.ema);                         // a = v / v_bar - 1,  ema = a - alpha
    e.v = v_bar * (1.0f + e.a);
    e.r = 1.0f / (e.v + kEps);
    e.b = -(e.a * v_bar + kEps) * e.r;                       // v_bar * r - 1
    e.bpa = (v_bar * (e.alpha * e.a - e.ema) + kEps * (e.alpha - 1.0f)) * e.r;   // b + alpha
    e.apb = (v_bar * e.a * e.a + kEps * (e.a - 1.0f)) * e.r;                     // a + b
    return e;
}

// closed-form gradient of sum_d g_m[d]*m_d w.r.t. (mu_k,d ; lv_k,d): O(1) per element, fp64 products/sums of the
// column statistics and the element's {mc, a, b, t} (no transcendental, no cancellation of large terms):
//   dT/dmu = 2[(mc R0 - R1) + r_k (B mc - S1)],  r_k = (1+b_k)/v_bar, R0 = (B + SB)/v_bar
//   dT/dlv = (B-1)(a_k - t_k) + (1+a_k)(SB - b_k) - (1+t_k)(SA - a_k) - (1+t_k)/v_bar (S2 - 2 S1 mc + B mc^2)
// gkl (nullable scalar for this row): upstream gradient of the analytic KL, adds gkl*mu and 0.5*gkl*(e^lv - 1).
__device__ __forceinline__ void mean_part(const double* __restrict__ cs, const double* __restrict__ pivots,
                                          const float* __restrict__ g_m, int B, int D, int d, const float4& pk, float mu_f,
                                          float lv_f, float& gmu, float& glv, const float* __restrict__ gkl_row = nullptr) {
    if (gkl_row != nullptr) {
        const float gk = gkl_row[0];
        gmu = fmaf(gk, mu_f, gmu);
        glv = fmaf(0.5f * gk, expm1f(lv_f), glv);
    }
    if (g_m == nullptr) return;
    const double nb = (double)B;
    const double w = 0.5 * (double)g_m[d] / (nb * (nb - 1.0));
    const double SA = cs[0 * (int64_t)D + d], SB = cs[1 * (int64_t)D + d], S1 = cs[3 * (int64_t)D + d],
                 S2 = cs[4 * (int64_t)D + d], R1 = cs[5 * (int64_t)D + d];
    const double ivb = pivots[3 * (int64_t)D + d];
    const double mc = pk.x, a = pk.y, b = pk.z, t = pk.w;
    const double r0 = (nb + SB) * ivb, rk = (1.0 + b) * ivb;
    const double dT_dmu = 2.0 * ((mc * r0 - R1) + rk * (nb * mc - S1));
    const double dT_dlv = (nb - 1.0) * (a - t) + (1.0 + a) * (SB - b) - (1.0 + t) * (SA - a) -
                          (1.0 + t) * ivb * (S2 - 2.0 * S1 * mc + nb * mc * mc);
    gmu += (float)(w * dT_dmu);
    glv += (float)(w * dT_dlv);
}

}  // namespace mae
