// A3' + A8: objective assembly of MAE.loss (mae/modules/mae.py:132-140) as one single-CTA kernel each way.
// At the training shapes (B = 64..100) this replaces ~12 latency-bound ATen launches by one.
#include "common.cuh"
#include "hot_math.cuh"

namespace mae {
namespace {

constexpr int kThreads = 256;

__device__ __forceinline__ float logsigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

__global__ void __launch_bounds__(kThreads)
assemble_fwd_kernel(const float* __restrict__ err, const float* __restrict__ kl, const float* __restrict__ m_d,
                    const float* __restrict__ S, const float* __restrict__ ext_sums, int B, int ns, int D, int kl_count,
                    int err_chunks, float eta, float gamma, float free_bits, float inv_batch, float* __restrict__ scalars,
                    float* __restrict__ recon) {
    __shared__ float scratch4[4][32];
    float s_rec = 0.f, s_kl = 0.f, s_m = 0.f, s_ls = 0.f;
    const float inv_ns = 1.0f / (float)ns;
    for (int b = threadIdx.x; b < B; b += kThreads) {
        float a = 0.f;
        const int cnt = ns * err_chunks;   // err_chunks > 1: per-CTA partial sums of the fused likelihood pass
        for (int s = 0; s < cnt; ++s) a += err[(int64_t)b * cnt + s];
        a *= inv_ns;
        recon[b] = a;
        s_rec += a;
    }
    for (int b = threadIdx.x; b < kl_count; b += kThreads) s_kl += kl[b];   // kl_count > B: the all-gathered KL vector
    if (m_d != nullptr) {
        for (int d = threadIdx.x; d < D; d += kThreads) {
            const float v = m_d[d];
            s_m += v;
            s_ls += logsigmoid(v);
        }
    }
    // the four sums in ONE barrier phase (this kernel sits on the join of every training step: four block_sum calls cost
    // eight barriers and four dependent shuffle trees); fixed order -> deterministic
    {
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        s_rec = warp_sum(s_rec);
        s_kl = warp_sum(s_kl);
        s_m = warp_sum(s_m);
        s_ls = warp_sum(s_ls);
        if (lane == 0) {
            scratch4[0][w] = s_rec;
            scratch4[1][w] = s_kl;
            scratch4[2][w] = s_m;
            scratch4[3][w] = s_ls;
        }
        __syncthreads();
        if (w == 0) {
            constexpr int nw = kThreads / 32;
            float a = lane < nw ? scratch4[0][lane] : 0.f, b = lane < nw ? scratch4[1][lane] : 0.f;
            float c = lane < nw ? scratch4[2][lane] : 0.f, d = lane < nw ? scratch4[3][lane] : 0.f;
            s_rec = warp_sum(a);
            s_kl = warp_sum(b);
            s_m = warp_sum(c);
            s_ls = warp_sum(d);
        }
    }
    if (threadIdx.x == 0) {
        if (ext_sums != nullptr) {
            s_rec += ext_sums[0];
            s_kl += ext_sums[1];
        }
        const float mean_rec = s_rec * inv_batch, mean_kl = s_kl * inv_batch;
        const float l_mean = (eta > 0.f && m_d != nullptr) ? -eta * s_ls : 0.f;
        const float l_std = (gamma > 0.f && S != nullptr) ? gamma * S[0] : 0.f;
        scalars[0] = mean_rec + fmaxf(mean_kl, free_bits) + l_mean + l_std;
        scalars[1] = s_m;
        scalars[2] = l_mean;
        scalars[3] = l_std;
        scalars[4] = mean_kl;
        scalars[5] = mean_rec;
    }
}

__global__ void __launch_bounds__(kThreads)
assemble_bwd_kernel(const float* __restrict__ gloss, const float* __restrict__ scalars, const float* __restrict__ m_d, int B,
                    int ns, int D, float eta, float gamma, float free_bits, float inv_batch, float* __restrict__ derr,
                    float* __restrict__ dkl, float* __restrict__ dm_d, float* __restrict__ dS) {
    const float g = gloss ? gloss[0] : 1.0f;
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (derr != nullptr && t < (int64_t)B * ns) derr[t] = g * inv_batch / (float)ns;
    if (dkl != nullptr && t < B) dkl[t] = scalars[4] >= free_bits ? g * inv_batch : 0.f;  // clamp(min) passes at equality
    if (dm_d != nullptr && m_d != nullptr && t < D) {
        // d/dm [-eta*logsigmoid(m)] = -eta*sigmoid(-m)
        const float v = m_d[t];
        const float e = expf(-fabsf(v));
        const float sig_neg = v >= 0.f ? e / (1.0f + e) : 1.0f / (1.0f + e);
        dm_d[t] = eta > 0.f ? -g * eta * sig_neg : 0.f;
    }
    if (dS != nullptr && t == 0) dS[0] = gamma > 0.f ? g * gamma : 0.f;
}

}  // namespace
}  // namespace mae

extern "C" int mae_loss_assemble_fwd(const float* err, const float* kl, const float* m_d, const float* S,
                                     const float* ext_sums, int64_t B, int64_t ns, int64_t D, int64_t kl_count,
                                     int64_t err_chunks, float eta, float gamma, float free_bits, float inv_batch, float* scalars,
                                     float* recon, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(err && kl && scalars && recon, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && ns > 0 && D >= 0 && B < (1ll << 31) && ns < (1ll << 31) && D < (1ll << 31), MAE_ERR_SHAPE);
    if (inv_batch <= 0.f) inv_batch = 1.0f / (float)B;
    if (kl_count <= 0) kl_count = B;
    if (err_chunks <= 0) err_chunks = 1;
    MAE_REQUIRE(kl_count < (1ll << 31) && ns * err_chunks < (1ll << 31), MAE_ERR_SHAPE);
    assemble_fwd_kernel<<<1, kThreads, 0, as_stream(stream)>>>(err, kl, m_d, S, ext_sums, (int)B, (int)ns, (int)D, (int)kl_count,
                                                              (int)err_chunks, eta, gamma, free_bits, inv_batch, scalars, recon);
    return launch_status();
}

extern "C" int mae_loss_assemble_bwd(const float* gloss, const float* scalars, const float* m_d, int64_t B, int64_t ns,
                                     int64_t D, float eta, float gamma, float free_bits, float inv_batch, float* derr,
                                     float* dkl, float* dm_d, float* dS, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(scalars, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && ns > 0 && D >= 0 && B < (1ll << 31) && ns < (1ll << 31) && D < (1ll << 31), MAE_ERR_SHAPE);
    if (inv_batch <= 0.f) inv_batch = 1.0f / (float)B;
    int64_t n = B * ns;
    if (D > n) n = D;
    const int64_t blocks = ceil_div(n, kThreads);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    assemble_bwd_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(gloss, scalars, m_d, (int)B, (int)ns, (int)D, eta,
                                                                             gamma, free_bits, inv_batch, derr, dkl, dm_d, dS);
    return launch_status();
}
