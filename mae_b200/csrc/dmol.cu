// A4b + A4c: discretized mixture of logistics straight from the raw NCHW decoder output
// (color_image_decoder.py:38-52, 91-125; utils.py:61-123).
//
// Layout: raw [N, 10*nmix, HW].  One thread owns one pixel of one sample and walks the nmix
// components; for a fixed channel the 32 lanes of a warp read 32 consecutive pixels (one 128 B line),
// so every global access is fully coalesced and each byte of `raw` is read exactly once by the
// forward (HBM-bound; SURVEY.md §8d: 4*(10*nmix*HW*N + 3*HW*B + N) bytes).  The backward recomputes
// the forward (second read of `raw` is served by L1/L2) and writes d raw once.
#include "common.cuh"
#include "hot_math.cuh"

namespace mae {
namespace {

constexpr int kThreads = 256;

// Load/store policies for hot_math.cuh's pixel routines (host branch only exists so the shared
// __host__ __device__ templates instantiate cleanly; it is never executed).
struct LdStream {  // read-once: bypass L1 (forward)
    static __host__ __device__ __forceinline__ float ld(const float* p) {
#ifdef __CUDA_ARCH__
        return ldg_stream(p);
#else
        return *p;
#endif
    }
};
struct LdCached {  // read-only path, allocates in L1 so the backward's second pass can hit
    static __host__ __device__ __forceinline__ float ld(const float* p) {
#ifdef __CUDA_ARCH__
        return __ldg(p);
#else
        return *p;
#endif
    }
};
struct StStream {
    static __host__ __device__ __forceinline__ void st(float* p, float v) {
#ifdef __CUDA_ARCH__
        stg_stream(p, v);
#else
        *p = v;
#endif
    }
};

// grid: (N, chunks) with chunks = ceil(HW / kThreads).  err[n] = -sum_pix [ LSE_m(T_m + logit_m) - LSE_m(logit_m) ].
// Cross-CTA reduction of the per-chunk partials is ordered (ticket; the last CTA of a sample adds the
// partials in chunk order), so the result is bit-reproducible.
__global__ void __launch_bounds__(kThreads)
dmol_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ x, int ns, int nmix, int HW, int chunks,
                float* __restrict__ err, float* partial, unsigned int* tickets) {
    __shared__ float scratch[32];
    __shared__ bool is_last;
    const int64_t n = blockIdx.x;
    const int64_t b = n / ns;
    const int pix = blockIdx.y * kThreads + threadIdx.x;
    float val = 0.f;
    if (pix < HW) {
        float xs[3];
#pragma unroll
        for (int c = 0; c < 3; ++c) xs[c] = x[(b * 3 + c) * HW + pix];
        const float* base = raw + n * (int64_t)(10 * nmix) * HW + pix;
        val = hm::dmol_pixel_fwd<LdStream>(base, HW, nmix, xs, nullptr, nullptr);
    }
    val = block_sum(val, scratch);
    if (chunks == 1) {
        if (threadIdx.x == 0) err[n] = val;
        return;
    }
    if (threadIdx.x == 0) {
        partial[n * chunks + blockIdx.y] = val;
        __threadfence();
        const unsigned int t = atomicAdd(&tickets[n], 1u);
        is_last = (t == (unsigned int)(chunks - 1));
    }
    __syncthreads();
    if (is_last && threadIdx.x == 0) {
        __threadfence();
        float tot = 0.f;
        for (int c = 0; c < chunks; ++c) tot += __ldcg(partial + n * chunks + c);
        err[n] = tot;
        tickets[n] = 0u;  // self-reset for the next call
    }
}

// d raw = g[n] * d err[n] / d raw.  Pass 1 recomputes both log-sum-exps, pass 2 recomputes each component
// with its local gradient and scales it by the component's responsibility.
__global__ void __launch_bounds__(kThreads)
dmol_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ x, const float* __restrict__ g, int ns, int nmix,
                int HW, float* __restrict__ draw) {
    const int64_t n = blockIdx.x;
    const int64_t b = n / ns;
    const int pix = blockIdx.y * kThreads + threadIdx.x;
    if (pix >= HW) return;
    float xs[3];
#pragma unroll
    for (int c = 0; c < 3; ++c) xs[c] = x[(b * 3 + c) * HW + pix];
    const int64_t off = n * (int64_t)(10 * nmix) * HW + pix;
    hm::dmol_pixel_bwd<LdCached, StStream>(raw + off, draw + off, HW, nmix, xs, g[n]);
}

struct DmolWs {
    int64_t partial_off, ticket_off, total;
};
DmolWs dmol_ws(int64_t N, int64_t HW) {
    const int64_t chunks = ceil_div(HW, kThreads);
    DmolWs w;
    w.partial_off = 0;
    w.ticket_off = round_up(N * chunks * (int64_t)sizeof(float), 256);
    w.total = w.ticket_off + round_up(N * (int64_t)sizeof(unsigned int), 256);
    return w;
}

int check_dmol(int64_t B, int64_t ns, int64_t nmix, int64_t HW) {
    MAE_REQUIRE(B > 0 && ns > 0 && nmix > 0 && HW > 0, MAE_ERR_SHAPE);
    MAE_REQUIRE(B * ns < (1ll << 31) && nmix <= 1024 && HW < (1ll << 24), MAE_ERR_SHAPE);
    MAE_REQUIRE(ceil_div(HW, kThreads) <= 65535, MAE_ERR_SHAPE);
    return MAE_OK;
}

}  // namespace
}  // namespace mae

extern "C" int64_t mae_dmol_workspace_bytes(int64_t N, int64_t HW) {
    if (N <= 0 || HW <= 0) return 0;
    return mae::dmol_ws(N, HW).total;
}

extern "C" int mae_dmol_fwd(const float* raw, const float* x, int64_t B, int64_t ns, int64_t nmix, int64_t HW, float* err,
                            void* workspace, int64_t workspace_bytes, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(raw && x && err, MAE_ERR_NULL);
    int rc = check_dmol(B, ns, nmix, HW);
    if (rc != MAE_OK) return rc;
    const int64_t N = B * ns;
    const int64_t chunks = ceil_div(HW, kThreads);
    const DmolWs w = dmol_ws(N, HW);
    float* partial = nullptr;
    unsigned int* tickets = nullptr;
    if (chunks > 1) {
        MAE_REQUIRE(workspace != nullptr, MAE_ERR_NULL);
        MAE_REQUIRE(workspace_bytes >= w.total, MAE_ERR_WORKSPACE);
        partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.partial_off);
        tickets = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + w.ticket_off);
    }
    dim3 grid((unsigned)N, (unsigned)chunks);
    dmol_fwd_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(raw, x, (int)ns, (int)nmix, (int)HW, (int)chunks, err, partial,
                                                             tickets);
    return launch_status();
}

extern "C" int mae_dmol_bwd(const float* raw, const float* x, const float* g, int64_t B, int64_t ns, int64_t nmix, int64_t HW,
                            float* draw, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(raw && x && g && draw, MAE_ERR_NULL);
    int rc = check_dmol(B, ns, nmix, HW);
    if (rc != MAE_OK) return rc;
    const int64_t N = B * ns;
    dim3 grid((unsigned)N, (unsigned)ceil_div(HW, kThreads));
    dmol_bwd_kernel<<<grid, kThreads, 0, as_stream(stream)>>>(raw, x, g, (int)ns, (int)nmix, (int)HW, draw);
    return launch_status();
}
