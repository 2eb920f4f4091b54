// A1 + A2: reparameterised posterior sample and analytic KL (gaussian_encoder.py:56-62, :219).
// HBM-bound streaming kernels: z is written once, eps read once; mu/logvar rows are re-read from L1/L2.
#include "common.cuh"

namespace mae {
namespace {

constexpr int kThreads = 256;
constexpr int kSamplesPerCta = 32;  // samples of one datum handled by a CTA (grid.y chunks)

// grid: (B, ceil(ns/kSamplesPerCta)).  Threads sweep the flattened (sample, dim) range of their chunk so
// that consecutive threads touch consecutive addresses of eps/z.
__global__ void __launch_bounds__(kThreads)
reparam_kl_fwd_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                      const float* __restrict__ eps, int ns, int D, float* __restrict__ z, float* __restrict__ kl) {
    __shared__ float scratch[32];
    pdl_launch_dependents();   // the likelihood kernel behind it may be placed while this (tiny) grid drains
    const int b = blockIdx.x;
    const float* mu_b = mu + (int64_t)b * mu_stride;
    const float* lv_b = lv + (int64_t)b * lv_stride;
    const int s0 = blockIdx.y * kSamplesPerCta;
    const int s1 = min(ns, s0 + kSamplesPerCta);
    const int64_t base = ((int64_t)b * ns + s0) * D;
    const int count = (s1 - s0) * D;
    for (int idx = threadIdx.x; idx < count; idx += kThreads) {
        const int d = idx % D;
        const float sd = expf(0.5f * lv_b[d]);
        z[base + idx] = ldg_stream(eps + base + idx) * sd + mu_b[d];
    }
    if (kl != nullptr && blockIdx.y == 0) {
        float acc = 0.f;
        for (int d = threadIdx.x; d < D; d += kThreads) {
            const float m = mu_b[d], l = lv_b[d];
            acc += m * m + expf(l) - l - 1.0f;
        }
        acc = block_sum(acc, scratch);
        if (threadIdx.x == 0) kl[b] = 0.5f * acc;
    }
}

// one thread per (b,d); loops over samples (ns is 1 in training, 5 in validation).
__global__ void __launch_bounds__(kThreads)
reparam_kl_bwd_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                      const float* __restrict__ eps, const float* __restrict__ gz, const float* __restrict__ gkl,
                      int B, int ns, int D, float* __restrict__ dmu, float* __restrict__ dlv, int accumulate,
                      const float* __restrict__ extra_dmu, const float* __restrict__ extra_dlv,
                      const float* __restrict__ extra_scale) {
    pdl_wait();                // gz / extra_* come from the predecessor in the stream
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= (int64_t)B * D) return;
    const int b = (int)(t / D), d = (int)(t % D);
    const float m = mu[(int64_t)b * mu_stride + d], l = lv[(int64_t)b * lv_stride + d];
    float sg = 0.f, sge = 0.f;
    if (gz != nullptr) {
        const int64_t base = (int64_t)b * ns * D + d;
        for (int s = 0; s < ns; ++s) {
            const float g = ldg_stream(gz + base + (int64_t)s * D);
            sg += g;
            sge += g * ldg_stream(eps + base + (int64_t)s * D);
        }
    }
    float gm = sg, gl = 0.5f * expf(0.5f * l) * sge;
    if (gkl != nullptr) {
        const float gk = gkl[b];
        gm += gk * m;
        gl += 0.5f * gk * (expf(l) - 1.0f);
    }
    if (extra_dmu != nullptr) {   // precomputed regulariser gradient (mae_kl_mpd_loss_fused), times its upstream scalar
        const float sc = extra_scale != nullptr ? extra_scale[0] : 1.0f;
        gm = fmaf(sc, extra_dmu[t], gm);
        gl = fmaf(sc, extra_dlv[t], gl);
    }
    if (accumulate) {
        dmu[t] += gm;
        dlv[t] += gl;
    } else {
        dmu[t] = gm;
        dlv[t] = gl;
    }
}

}  // namespace
}  // namespace mae

extern "C" int mae_reparam_kl_fwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                                  const float* eps, int64_t B, int64_t ns, int64_t D, float* z, float* kl,
                                  mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(mu && logvar && eps && z, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && ns > 0 && D > 0 && B < (1ll << 31) && D < (1ll << 24) && ns < (1ll << 24), MAE_ERR_SHAPE);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    const int64_t chunks = ceil_div(ns, kSamplesPerCta);
    MAE_REQUIRE(chunks <= 65535, MAE_ERR_SHAPE);
    dim3 grid((unsigned)B, (unsigned)chunks);
    reparam_kl_fwd_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(mu, mu_stride, logvar, logvar_stride, eps,
                                                                   (int)ns, (int)D, z, kl);
    return launch_status();
}

extern "C" int mae_reparam_reg_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                                   const float* eps, const float* gz, const float* gkl, const float* extra_dmu,
                                   const float* extra_dlogvar, const float* extra_scale, int64_t B, int64_t ns, int64_t D,
                                   float* dmu, float* dlogvar, int accumulate, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(mu && logvar && dmu && dlogvar, MAE_ERR_NULL);
    MAE_REQUIRE(gz == nullptr || eps != nullptr, MAE_ERR_NULL);
    MAE_REQUIRE((extra_dmu == nullptr) == (extra_dlogvar == nullptr), MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && ns > 0 && D > 0 && B < (1ll << 31) && D < (1ll << 24) && ns < (1ll << 24), MAE_ERR_SHAPE);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    const int64_t blocks = ceil_div(B * D, kThreads);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    launch_pdl(reparam_kl_bwd_kernel, dim3((unsigned)blocks), dim3(kThreads), 0, as_stream(stream), mu, mu_stride, logvar,
               logvar_stride, eps, gz, gkl, (int)B, (int)ns, (int)D, dmu, dlogvar, accumulate, extra_dmu, extra_dlogvar,
               extra_scale);
    return launch_status();
}

extern "C" int mae_reparam_kl_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                                  const float* eps, const float* gz, const float* gkl, int64_t B, int64_t ns, int64_t D,
                                  float* dmu, float* dlogvar, int accumulate, mae_stream_t stream) {
    return mae_reparam_reg_bwd(mu, mu_stride, logvar, logvar_stride, eps, gz, gkl, nullptr, nullptr, nullptr, B, ns, D, dmu,
                               dlogvar, accumulate, stream);
}
