// A5, A6, A7: Gaussian log densities and the fused importance-weighted NLL estimator
// (gaussian_encoder.py:231-260, 262-301; mae.py:203-215; utils.py:23-38).
// All kernels are single-pass streams over z: one warp per (b,k) sample, lanes strided over the latent
// dimension (coalesced), fixed-order shuffles -> deterministic.
#include "common.cuh"
#include "hot_math.cuh"

namespace mae {
namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;

__global__ void __launch_bounds__(kThreads)
prior_fwd_kernel(const float* __restrict__ z, const float* __restrict__ logdet, int64_t N, int D, float* __restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= N) return;
    const int lane = threadIdx.x & 31;
    const float* zr = z + n * D;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float v = ldg_stream(zr + d);
        acc += v * v;
    }
    acc = warp_sum(acc);
    if (lane == 0) out[n] = -0.5f * (acc + (float)D * kLog2Pi) + (logdet ? logdet[n] : 0.f);
}

__global__ void __launch_bounds__(kThreads)
prior_bwd_kernel(const float* __restrict__ z, const float* __restrict__ g, int64_t total, int D, float* __restrict__ dz) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    dz[t] = -g[t / D] * z[t];
}

__global__ void __launch_bounds__(kThreads)
posterior_fwd_kernel(const float* __restrict__ z, const float* __restrict__ logdet, const float* __restrict__ mu,
                     int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride, int64_t BK, int K, int D,
                     float* __restrict__ out) {
    const int64_t n = (int64_t)blockIdx.x * kWarps + (threadIdx.x >> 5);
    if (n >= BK) return;
    const int lane = threadIdx.x & 31;
    const int64_t b = n / K;
    const float* zr = z + n * D;
    const float* mr = mu + b * mu_stride;
    const float* lr = lv + b * lv_stride;
    float acc = 0.f;
    for (int d = lane; d < D; d += 32) {
        const float l = lr[d];
        const float diff = ldg_stream(zr + d) - mr[d];
        acc += l + diff * diff / (expf(l) + kEps);
    }
    acc = warp_sum(acc);
    if (lane == 0) out[n] = -0.5f * (acc + (float)D * kLog2Pi) + (logdet ? logdet[n] : 0.f);
}

// one thread per (b,d): walks the K samples, writes dz and accumulates dmu / dlogvar.
__global__ void __launch_bounds__(kThreads)
posterior_bwd_kernel(const float* __restrict__ z, const float* __restrict__ g, const float* __restrict__ mu,
                     int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride, int B, int K, int D,
                     float* __restrict__ dz, float* __restrict__ dmu, float* __restrict__ dlv) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= (int64_t)B * D) return;
    const int b = (int)(t / D), d = (int)(t % D);
    const float m = mu[(int64_t)b * mu_stride + d], l = lv[(int64_t)b * lv_stride + d];
    const float v = expf(l);
    const float r = 1.0f / (v + kEps);
    float am = 0.f, al = 0.f;
    for (int k = 0; k < K; ++k) {
        const int64_t o = ((int64_t)b * K + k) * D + d;
        const float gk = g[(int64_t)b * K + k];
        const float diff = z[o] - m;
        const float t1 = gk * diff * r;
        dz[o] = -t1;
        am += t1;
        al += -0.5f * gk * (1.0f - diff * diff * r * r * v);
    }
    dmu[t] = am;
    dlv[t] = al;
}

// grid: (B, chunks).  A CTA handles kSamplesPerCta consecutive samples of one datum; each warp walks its samples
// (two squared norms reduced across lanes per sample) keeping an online log-sum-exp and three running sums.  The
// CTA's merged state goes to the workspace; the last CTA of a datum (ticket) merges the chunk states in order.
constexpr int kSamplesPerCta = 64;

struct IwState {
    float m, s, iw, gen, kl;
};

__global__ void __launch_bounds__(kThreads)
iw_nll_kernel(const float* __restrict__ log_gen, const float* __restrict__ zp, const float* __restrict__ ldp,
              const float* __restrict__ zq, const float* __restrict__ ldq, const float* __restrict__ mu, int64_t mu_stride,
              const float* __restrict__ lv, int64_t lv_stride, int K, int D, int chunks, float* part, unsigned int* tickets,
              float* __restrict__ nll_elbo, float* __restrict__ recon, float* __restrict__ kl, float* __restrict__ nll_iw,
              float gen_scale) {
    __shared__ IwState sh[kWarps];
    __shared__ float scratch[32];
    __shared__ bool is_last;
    pdl_wait();   // log_gen comes from the likelihood kernel right before it in the stream
    const int b = blockIdx.x, chunk = blockIdx.y;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const float* mr = mu + (int64_t)b * mu_stride;
    const float* lr = lv + (int64_t)b * lv_stride;
    float lsum = 0.f;
    for (int d = threadIdx.x; d < D; d += kThreads) lsum += lr[d];
    lsum = block_sum(lsum, scratch);  // sum_d logvar[b,d]
    const float cst = (float)D * kLog2Pi;

    hm::Lse lse;
    lse.init();
    float s_iw = 0.f, s_gen = 0.f, s_kl = 0.f;
    const int k1 = min(K, (chunk + 1) * kSamplesPerCta);
    for (int k = chunk * kSamplesPerCta + w; k < k1; k += kWarps) {
        const int64_t n = (int64_t)b * K + k;
        const float* zpr = zp + n * D;
        const float* zqr = zq + n * D;
        float ap = 0.f, aq = 0.f;
        for (int d = lane; d < D; d += 32) {
            const float a = ldg_stream(zpr + d);
            ap += a * a;
            const float diff = ldg_stream(zqr + d) - mr[d];
            aq += diff * diff / (expf(lr[d]) + kEps);
        }
        ap = warp_sum(ap);
        aq = warp_sum(aq);
        const float lp = -0.5f * (ap + cst) + (ldp ? ldp[n] : 0.f);
        const float lq = -0.5f * (lsum + aq + cst) + (ldq ? ldq[n] : 0.f);
        const float lg = gen_scale * log_gen[n];
        const float liw = lg + lp - lq;
        lse.push(liw);
        s_iw += liw;
        s_gen += lg;
        s_kl += lq - lp;
    }
    if (lane == 0) sh[w] = IwState{lse.m, lse.s, s_iw, s_gen, s_kl};
    __syncthreads();
    if (threadIdx.x == 0) {
        hm::Lse tot;
        tot.init();
        float a = 0.f, c = 0.f, e = 0.f;
        for (int i = 0; i < kWarps; ++i) {
            hm::Lse o;
            o.m = sh[i].m;
            o.s = sh[i].s;
            tot.merge(o);
            a += sh[i].iw;
            c += sh[i].gen;
            e += sh[i].kl;
        }
        float* p = part + ((int64_t)b * chunks + chunk) * 5;
        p[0] = tot.m; p[1] = tot.s; p[2] = a; p[3] = c; p[4] = e;
        is_last = true;
        if (chunks > 1) {
            __threadfence();
            const unsigned int t = atomicAdd(&tickets[b], 1u);
            is_last = (t == (unsigned int)(chunks - 1));
        }
        if (is_last) {
            __threadfence();
            hm::Lse all;
            all.init();
            float sa = 0.f, sc = 0.f, se = 0.f;
            for (int i = 0; i < chunks; ++i) {
                const float* q = part + ((int64_t)b * chunks + i) * 5;
                hm::Lse o;
                o.m = __ldcg(q);
                o.s = __ldcg(q + 1);
                all.merge(o);
                sa += __ldcg(q + 2);
                sc += __ldcg(q + 3);
                se += __ldcg(q + 4);
            }
            const float invk = 1.0f / (float)K;
            nll_elbo[b] = -sa * invk;
            recon[b] = -sc * invk;
            kl[b] = se * invk;
            nll_iw[b] = -all.value() + logf((float)K);
            tickets[b] = 0u;
        }
    }
}

struct IwWs {
    int64_t part_off, ticket_off, total, chunks;
};
IwWs iw_ws(int64_t B, int64_t K) {
    IwWs w;
    w.chunks = ceil_div(K, kSamplesPerCta);
    w.part_off = 0;
    w.ticket_off = round_up(B * w.chunks * 5 * (int64_t)sizeof(float), 256);
    w.total = w.ticket_off + round_up(B * (int64_t)sizeof(unsigned int), 256);
    return w;
}

}  // namespace
}  // namespace mae

extern "C" int mae_log_prob_prior_fwd(const float* z, const float* logdet, int64_t N, int64_t D, float* out,
                                      mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(z && out, MAE_ERR_NULL);
    MAE_REQUIRE(N > 0 && D > 0 && D < (1ll << 30), MAE_ERR_SHAPE);
    const int64_t blocks = ceil_div(N, kWarps);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    prior_fwd_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(z, logdet, N, (int)D, out);
    return launch_status();
}

extern "C" int mae_log_prob_prior_bwd(const float* z, const float* g, int64_t N, int64_t D, float* dz, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(z && g && dz, MAE_ERR_NULL);
    MAE_REQUIRE(N > 0 && D > 0 && D < (1ll << 30), MAE_ERR_SHAPE);
    const int64_t blocks = ceil_div(N * D, kThreads);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    prior_bwd_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(z, g, N * D, (int)D, dz);
    return launch_status();
}

extern "C" int mae_log_prob_posterior_fwd(const float* z, const float* logdet, const float* mu, int64_t mu_stride,
                                          const float* logvar, int64_t logvar_stride, int64_t B, int64_t K, int64_t D,
                                          float* out, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(z && mu && logvar && out, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && K > 0 && D > 0 && D < (1ll << 30) && K < (1ll << 30), MAE_ERR_SHAPE);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    const int64_t blocks = ceil_div(B * K, kWarps);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    posterior_fwd_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(z, logdet, mu, mu_stride, logvar, logvar_stride,
                                                                              B * K, (int)K, (int)D, out);
    return launch_status();
}

extern "C" int mae_log_prob_posterior_bwd(const float* z, const float* g, const float* mu, int64_t mu_stride,
                                          const float* logvar, int64_t logvar_stride, int64_t B, int64_t K, int64_t D,
                                          float* dz, float* dmu, float* dlogvar, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(z && g && mu && logvar && dz && dmu && dlogvar, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && K > 0 && D > 0 && D < (1ll << 30) && K < (1ll << 30) && B < (1ll << 31), MAE_ERR_SHAPE);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    const int64_t blocks = ceil_div(B * D, kThreads);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    posterior_bwd_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(z, g, mu, mu_stride, logvar, logvar_stride, (int)B,
                                                                              (int)K, (int)D, dz, dmu, dlogvar);
    return launch_status();
}

extern "C" int64_t mae_iw_nll_workspace_bytes(int64_t B, int64_t K) {
    if (B <= 0 || K <= 0) return 0;
    return mae::iw_ws(B, K).total;
}

extern "C" int mae_iw_nll(const float* log_gen, const float* z_prior_normal, const float* logdet_prior,
                          const float* z_post_normal, const float* logdet_post, const float* mu, int64_t mu_stride,
                          const float* logvar, int64_t logvar_stride, int64_t B, int64_t K, int64_t D, float gen_scale,
                          float* nll_elbo, float* recon, float* kl, float* nll_iw, void* workspace, int64_t workspace_bytes,
                          mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(log_gen && z_prior_normal && z_post_normal && mu && logvar && nll_elbo && recon && kl && nll_iw && workspace,
                MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && K > 0 && D > 0 && B < (1ll << 31) && D < (1ll << 30) && K < (1ll << 22), MAE_ERR_SHAPE);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    const IwWs w = iw_ws(B, K);
    MAE_REQUIRE(workspace_bytes >= w.total, MAE_ERR_WORKSPACE);
    MAE_REQUIRE(w.chunks <= 65535, MAE_ERR_SHAPE);
    dim3 grid((unsigned)B, (unsigned)w.chunks);
    launch_pdl(iw_nll_kernel, grid, dim3(kThreads), 0, as_stream(stream), log_gen, z_prior_normal, logdet_prior, z_post_normal,
               logdet_post, mu, mu_stride, logvar, logvar_stride, (int)K, (int)D, (int)w.chunks,
               reinterpret_cast<float*>(static_cast<char*>(workspace) + w.part_off),
               reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + w.ticket_off), nll_elbo, recon, kl, nll_iw,
               gen_scale);
    return launch_status();
}
