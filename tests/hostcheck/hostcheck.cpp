// TEST-ONLY host instantiation of mae_b200/csrc/hot_math.cuh (the per-pixel arithmetic the CUDA kernels
// inline).  Lets the CPU test-suite validate the DMOL / Bernoulli formulas and analytic gradients against
// the golden fixtures without a GPU.  Not linked into libmae_b200.so; not reachable from the product.
#include <stdint.h>

#include "../../mae_b200/csrc/hot_math.cuh"

using namespace mae::hm;

extern "C" void hc_dmol_fwd(const float* raw, const float* x, int64_t B, int64_t ns, int64_t nmix, int64_t HW, float* err) {
    for (int64_t n = 0; n < B * ns; ++n) {
        const int64_t b = n / ns;
        float tot = 0.f;
        for (int64_t p = 0; p < HW; ++p) {
            const float xs[3] = {x[(b * 3 + 0) * HW + p], x[(b * 3 + 1) * HW + p], x[(b * 3 + 2) * HW + p]};
            tot += dmol_pixel_fwd<LdPlain>(raw + n * 10 * nmix * HW + p, HW, (int)nmix, xs, nullptr, nullptr);
        }
        err[n] = tot;
    }
}

// forward through the packed two-pixel path (pixels p, p+1; an odd tail pixel is paired with itself)
extern "C" void hc_dmol_fwd_pair(const float* raw, const float* x, int64_t B, int64_t ns, int64_t nmix, int64_t HW, float* err) {
    for (int64_t n = 0; n < B * ns; ++n) {
        const int64_t b = n / ns;
        float tot = 0.f;
        for (int64_t p = 0; p < HW; p += 2) {
            const int64_t p1 = p + 1 < HW ? p + 1 : p;
            float2 xs[3];
            for (int c = 0; c < 3; ++c) xs[c] = mk2(x[(b * 3 + c) * HW + p], x[(b * 3 + c) * HW + p1]);
            const float* base = raw + n * 10 * nmix * HW;
            Lse2 lt, ll;
            lt.init();
            ll.init();
            for (int m = 0; m < nmix; ++m) {
                float2 mu[3], rho[3], kap[3];
                for (int c = 0; c < 3; ++c) {
                    const int64_t o = (3 * m + c) * HW;
                    mu[c] = mk2(base[(nmix)*HW + o + p], base[(nmix)*HW + o + p1]);
                    rho[c] = mk2(base[4 * nmix * HW + o + p], base[4 * nmix * HW + o + p1]);
                    kap[c] = mk2(base[7 * nmix * HW + o + p], base[7 * nmix * HW + o + p1]);
                }
                const float2 lg = mk2(base[m * HW + p], base[m * HW + p1]);
                const float2 T = dmol_component2(mu, rho, kap, xs);
                lt.push(add2(T, lg));
                ll.push(lg);
            }
            const float2 zt = lt.value(), zl = ll.value();
            tot += zl.x - zt.x;
            if (p1 != p) tot += zl.y - zt.y;
        }
        err[n] = tot;
    }
}

extern "C" void hc_dmol_bwd(const float* raw, const float* x, const float* g, int64_t B, int64_t ns, int64_t nmix, int64_t HW,
                            float* draw) {
    for (int64_t n = 0; n < B * ns; ++n) {
        const int64_t b = n / ns;
        for (int64_t p = 0; p < HW; ++p) {
            const float xs[3] = {x[(b * 3 + 0) * HW + p], x[(b * 3 + 1) * HW + p], x[(b * 3 + 2) * HW + p]};
            const int64_t off = n * 10 * nmix * HW + p;
            dmol_pixel_bwd<LdPlain, StPlain>(raw + off, draw + off, HW, (int)nmix, xs, g[n]);
        }
    }
}

extern "C" void hc_bernoulli(const float* probs, const float* x, const float* g, int64_t B, int64_t ns, int64_t P, float* err,
                             float* dprobs) {
    for (int64_t r = 0; r < B * ns; ++r) {
        const int64_t b = r / ns;
        float acc = 0.f;
        for (int64_t p = 0; p < P; ++p) {
            acc += bernoulli_logprob(probs[r * P + p], x[b * P + p]);
            dprobs[r * P + p] = g[r] * bernoulli_derr_dp(probs[r * P + p], x[b * P + p]);
        }
        err[r] = -acc;
    }
}
