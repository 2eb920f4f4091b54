// Pipe-peak micro-benchmarks for the roofline denominators bench.py reports ("FP32-FFMA and MUFU peaks
// are not in MEASURED_PEAKS.json -> micro-benchmark both on the box").  Each kernel is a register-only loop of
// independent dependent-chains on ONE pipe, timed by the caller with CUDA events:
//   ffma : 16 independent fma.rn.f32 chains per thread            -> 2 flop per instruction
//   mufu : 8 independent chains of one SFU instruction per thread -> ex2.approx / lg2.approx / rcp.approx
// Results are written only under a condition the compiler cannot evaluate, so nothing is optimised away and nothing
// is stored on the common path.
#include "common.cuh"

namespace mae {
namespace {

constexpr int kMbThreads = 512;

__global__ void __launch_bounds__(kMbThreads) ffma_peak_kernel(int iters, float x, float y, float* __restrict__ sink) {
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = (float)(threadIdx.x + i) * 1e-3f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], x, y);
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += a[i];
    if (s == 123.456f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int OP>
__device__ __forceinline__ float mufu_op(float v) {
    float r;
    if (OP == 0) asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    else if (OP == 1) asm volatile("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    else asm volatile("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
    return r;
}

// ex2: v -> 2^(-v) stays in (0.5, 1); lg2: v -> lg2(v) + 3 stays near 4.6; rcp: v -> 1/v + 0.5 converges to 1.28.  The negation / add
// rides on the FMA pipe (4x the SFU issue rate), so the SFU stays the limiter.
template <int OP>
__global__ void __launch_bounds__(kMbThreads) mufu_peak_kernel(int iters, float seed, float* __restrict__ sink) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = seed + (float)((threadIdx.x + i) & 7) * 0.03125f;
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                if (OP == 0) a[i] = mufu_op<0>(-a[i]);
                else if (OP == 1) a[i] = mufu_op<1>(a[i]) + 3.0f;
                else a[i] = mufu_op<2>(a[i]) + 0.5f;
            }
    }
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += a[i];
    if (s == 123.456f) sink[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

}  // namespace
}  // namespace mae

// One launch of the FFMA loop: blocks x 512 threads, each executing iters * 64 FFMAs.  *ops_out (host, nullable) receives
// the number of FFMA instructions (lane-level) of the launch: flop = 2 * ops.
extern "C" int mae_microbench_ffma(int64_t blocks, int64_t iters, float* sink, double* ops_out, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(sink, MAE_ERR_NULL);
    MAE_REQUIRE(blocks > 0 && blocks < (1 << 20) && iters > 0 && iters < (1ll << 30), MAE_ERR_SHAPE);
    ffma_peak_kernel<<<(unsigned)blocks, kMbThreads, 0, as_stream(stream)>>>((int)iters, 0.999f, 1e-3f, sink);
    if (ops_out) *ops_out = (double)blocks * kMbThreads * (double)iters * 64.0;
    return launch_status();
}

// op: 0 ex2.approx, 1 lg2.approx, 2 rcp.approx.  Each thread executes iters * 32 SFU instructions.
extern "C" int mae_microbench_mufu(int op, int64_t blocks, int64_t iters, float* sink, double* ops_out, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(sink, MAE_ERR_NULL);
    MAE_REQUIRE(op >= 0 && op <= 2, MAE_ERR_UNSUPPORTED);
    MAE_REQUIRE(blocks > 0 && blocks < (1 << 20) && iters > 0 && iters < (1ll << 30), MAE_ERR_SHAPE);
    cudaStream_t s = as_stream(stream);
    if (op == 0) mufu_peak_kernel<0><<<(unsigned)blocks, kMbThreads, 0, s>>>((int)iters, 0.6f, sink);
    else if (op == 1) mufu_peak_kernel<1><<<(unsigned)blocks, kMbThreads, 0, s>>>((int)iters, 4.6f, sink);
    else mufu_peak_kernel<2><<<(unsigned)blocks, kMbThreads, 0, s>>>((int)iters, 1.7f, sink);
    if (ops_out) *ops_out = (double)blocks * kMbThreads * (double)iters * 32.0;
    return launch_status();
}
