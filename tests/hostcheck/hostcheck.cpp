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
