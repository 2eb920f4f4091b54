// A4a: Bernoulli reconstruction error (binary_image_decoder.py:84-93, pixelcnn.py:136-139).
#include "common.cuh"
#include "hot_math.cuh"

namespace mae {
namespace {

constexpr int kThreads = 256;

// one CTA per (b,s) row of P pixels; fixed-order block reduction -> deterministic.
__global__ void __launch_bounds__(kThreads)
bernoulli_fwd_kernel(const float* __restrict__ probs, const float* __restrict__ x, int ns, int P, float* __restrict__ err) {
    __shared__ float scratch[32];
    const int64_t row = blockIdx.x;
    const int64_t b = row / ns;
    const float* pr = probs + row * P;
    const float* xr = x + b * P;
    float acc = 0.f;
    for (int p = threadIdx.x; p < P; p += kThreads) acc += hm::bernoulli_logprob(ldg_stream(pr + p), xr[p]);
    acc = block_sum(acc, scratch);
    if (threadIdx.x == 0) err[row] = -acc;
}

__global__ void __launch_bounds__(kThreads)
bernoulli_bwd_kernel(const float* __restrict__ probs, const float* __restrict__ x, const float* __restrict__ g,
                     int g_stride, float g_mul, int ns, int P, int64_t total, float* __restrict__ dprobs) {
    const int64_t t = (int64_t)blockIdx.x * kThreads + threadIdx.x;
    if (t >= total) return;
    const int64_t row = t / P;
    const int p = (int)(t % P);
    const int64_t b = row / ns;
    stg_stream(dprobs + t, g[row * g_stride] * g_mul * hm::bernoulli_derr_dp(ldg_stream(probs + t), x[b * P + p]));
}

}  // namespace
}  // namespace mae

extern "C" int mae_bernoulli_fwd(const float* probs, const float* x, int64_t B, int64_t ns, int64_t P, float* err,
                                 mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(probs && x && err, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && ns > 0 && P > 0 && B * ns < (1ll << 31) && P < (1ll << 31), MAE_ERR_SHAPE);
    bernoulli_fwd_kernel<<<(unsigned)(B * ns), kThreads, 0, as_stream(stream)>>>(probs, x, (int)ns, (int)P, err);
    return launch_status();
}

extern "C" int mae_bernoulli_bwd(const float* probs, const float* x, const float* g, int64_t g_stride, float g_mul, int64_t B,
                                 int64_t ns, int64_t P, float* dprobs, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(probs && x && g && dprobs, MAE_ERR_NULL);
    MAE_REQUIRE(g_stride == 0 || g_stride == 1, MAE_ERR_STRIDE);
    MAE_REQUIRE(B > 0 && ns > 0 && P > 0 && B * ns < (1ll << 31) && P < (1ll << 31), MAE_ERR_SHAPE);
    const int64_t total = B * ns * P;
    const int64_t blocks = ceil_div(total, kThreads);
    MAE_REQUIRE(blocks < (1ll << 31), MAE_ERR_SHAPE);
    bernoulli_bwd_kernel<<<(unsigned)blocks, kThreads, 0, as_stream(stream)>>>(probs, x, g, (int)g_stride, g_mul, (int)ns, (int)P, total, dprobs);
    return launch_status();
}
