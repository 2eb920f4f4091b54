// Micro-benchmark: how fast does ONE thread retire tcgen05.mma.kind::tf32 instructions, as a function of the instruction
// shape (N = 128 / 256), the source of A (TMEM or shared memory) and the shared-memory layout of B (un-swizzled canonical
// 8 x 16 B core matrices, as the kernels of this repo use, or SWIZZLE_128B)?  Timing only: operands are constant ones, the
// result is not checked.  Guide floor: max(M,128) * N / 256 cycles per instruction (64 at N = 128, 128 at N = 256).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o scripts/mma_rate scripts/mma_rate.cu && scripts/mma_rate
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "../mae_b200/csrc/tc5.cuh"

namespace tc5 = mae::tc5;
namespace {

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr) {      // K-major, SWIZZLE_128B: 8 rows x 128 B atoms, SBO = 1024
    return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)(1024 >> 4) << 32) | (1ull << 46) | (2ull << 61);
}

// compile-time variant: N256, A in shared memory (SS) or TMEM (TS), SWIZZLE_128B or un-swizzled B, TWO accumulators in turn.
// The eight descriptors of two stages x four k-steps are formed before the timed loop, which is unrolled by 8: nothing but
// the MMA instructions (and their predicate) is on the issuing thread's path.
template <int N256, int A_SMEM, int SW128, int TWO_ACC, int TRAFFIC>
__global__ void __launch_bounds__(256, 1) rate_kernel(long long* out, int iters, const float* src) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint64_t bar;
    __shared__ uint32_t slot;
    __shared__ volatile int stop;
    __shared__ uint64_t cbar[4];
    float* f = reinterpret_cast<float*>(smem);
    for (int i = threadIdx.x; i < (96 * 1024) / 4; i += blockDim.x) f[i] = 1.0f;
    if (threadIdx.x == 0) { tc5::mbar_init(&bar, 1); tc5::fence_barrier_init(); stop = 0; for (int q = 0; q < 4; ++q) tc5::mbar_init(&cbar[q], 1); }
    if (threadIdx.x < 32) tc5::tmem_alloc(&slot, 512);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem = slot;
    if (threadIdx.x < 128) {   // A operand in TMEM: ones in columns 384..447 of every lane
        float v[16];
        for (int i = 0; i < 16; ++i) v[i] = 1.0f;
        const uint32_t lane_base = tmem + ((uint32_t)(threadIdx.x & ~31u) << 16);
        for (int c = 0; c < 64; c += 16) tc5::tmem_st16(lane_base + 384 + c, v);
        tc5::tmem_st_wait();
    }
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    if (threadIdx.x == 0) {
        constexpr int N = N256 ? 256 : 128;
        constexpr uint32_t idesc = tc5::idesc_tf32(128, N);
        const uint32_t sb = tc5::smem_u32(smem);              // B: 64 KB region
        const uint32_t sa = sb + 64 * 1024;                   // A (SS mode): 32 KB region
        constexpr uint32_t lbo = (uint32_t)N * 16;
        uint64_t bd[8], ad[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) {
            const int j = q & 3, st = q >> 2;
            bd[q] = SW128 ? desc_sw128(sb + st * 32768 + j * 32) : tc5::smem_desc(sb + st * 32768 + 2 * j * lbo, lbo, 128);
            ad[q] = SW128 ? desc_sw128(sa + st * 16384 + j * 32) : tc5::smem_desc(sa + st * 16384 + 2 * j * 2048, 2048, 128);
        }
        const long long t0 = clock64();
        for (int it = 0; it < iters; it += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                const uint32_t d = tmem + (TWO_ACC ? (uint32_t)((q & 1) * N) : 0u);
                if (A_SMEM) tc5::mma_tf32(d, ad[q], bd[q], idesc, 1u);
                else tc5::mma_tf32_ts(d, tmem + 384 + 8 * (q & 3), bd[q], idesc, 1u);
            }
        }
        const long long t1 = clock64();
        tc5::mma_commit(&bar);
        tc5::mbar_wait(&bar, 0);
        const long long t2 = clock64();
        if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
        stop = 1;
    } else if (TRAFFIC == 3) {
        // a loader keeping four 16 KB bulk copies (global -> shared) in flight into a region the MMAs do not read
        if (threadIdx.x == 128) {
            long long n = 0;
            for (int q = 0; q < 4; ++q) {
                tc5::mbar_expect_tx(&cbar[q], 16384);
                tc5::bulk_g2s(smem + 96 * 1024 + q * 16384, src + (size_t)((blockIdx.x * 4 + q) % 64) * 4096, 16384, &cbar[q]);
            }
            while (!stop) {
                const int q = (int)(n & 3);
                tc5::mbar_wait(&cbar[q], (uint32_t)(n >> 2) & 1u);
                tc5::mbar_expect_tx(&cbar[q], 16384);
                tc5::bulk_g2s(smem + 96 * 1024 + q * 16384, src + (size_t)((blockIdx.x * 4 + n) % 64) * 4096, 16384, &cbar[q]);
                ++n;
            }
            for (int q = 0; q < 4; ++q) { const long long m = n + q; tc5::mbar_wait(&cbar[m & 3], (uint32_t)(m >> 2) & 1u); }
            if (blockIdx.x == 0) out[2] = n;
        }
    } else if (threadIdx.x >= 128 && TRAFFIC != 0) {
        // the other roles of a real kernel: an epilogue reading an accumulator (columns 256..383) or producers writing operand
        // stages (columns 448..511), lanes of this warp's TMEM quarter, until the issuer is done
        const uint32_t lane_base = tmem + ((uint32_t)((threadIdx.x & 127) & ~31u) << 16);
        float v[16];
        for (int i = 0; i < 16; ++i) v[i] = 1.0f;
        float acc = 0.f;
        long long n = 0;
        while (!stop) {
            if (TRAFFIC == 1) {
#pragma unroll
                for (int c = 0; c < 128; c += 16) { tc5::tmem_ld16(lane_base + 256 + c, v); tc5::tmem_ld_wait(); acc += v[0] + v[15]; }
            } else {
#pragma unroll
                for (int c = 0; c < 64; c += 16) { tc5::tmem_st16(lane_base + 448 + c, v); }
                tc5::tmem_st_wait();
            }
            ++n;
        }
        if (acc == 123.456f) out[3] = n;
        if (blockIdx.x == 0 && threadIdx.x == 128) out[2] = n;
    }
    __syncthreads();
    if (threadIdx.x < 32) tc5::tmem_dealloc(tmem, 512);
}

template <int N256, int A_SMEM, int SW128, int TWO_ACC, int TRAFFIC = 0>
void run(const char* name, int grid, int iters, long long* out) {
    auto k = rate_kernel<N256, A_SMEM, SW128, TWO_ACC, TRAFFIC>;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 164 * 1024);
    static float* src = nullptr;
    if (src == nullptr) { cudaMalloc(&src, 64 * 16384); cudaMemset(src, 0, 64 * 16384); }
    long long h[3] = {0, 0, 0};
    for (int rep = 0; rep < 2; ++rep) {
        k<<<grid, 256, 164 * 1024>>>(out, iters, src);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, out, 24, cudaMemcpyDeviceToHost);
    printf("  %s issue %6.1f cyc/MMA, complete %6.1f cyc/MMA (floor %d)", name, (double)h[0] / iters, (double)h[1] / iters, N256 ? 128 : 64);
    if (TRAFFIC == 1) printf("   [4 warps tcgen05.ld: %.1f B/clk]", (double)h[2] * 128 * 128 * 4 / (double)h[1]);
    if (TRAFFIC == 2) printf("   [4 warps tcgen05.st: %.1f B/clk]", (double)h[2] * 64 * 128 * 4 / (double)h[1]);
    if (TRAFFIC == 3) printf("   [bulk copies into smem: %.1f B/clk]", (double)h[2] * 16384 / (double)h[1]);
    printf("\n");
}

}  // namespace

int main() {
    long long* out;
    cudaMalloc(&out, 64);
    cudaMemset(out, 0, 64);
    const int iters = 8192;
    for (int grid : {1, 148}) {
        printf("grid = %d CTA(s), %d instructions M=128 x N x K=8 (tf32) from one thread\n", grid, iters);
        run<0, 0, 0, 0>("TS  N=128 no-swizzle (this repo's layout)", grid, iters, out);
        run<0, 0, 0, 1>("TS  N=128 no-swizzle, two accumulators   ", grid, iters, out);
        run<0, 0, 1, 0>("TS  N=128 SWIZZLE_128B                   ", grid, iters, out);
        run<1, 0, 0, 0>("TS  N=256 no-swizzle                     ", grid, iters, out);
        run<1, 0, 1, 0>("TS  N=256 SWIZZLE_128B                   ", grid, iters, out);
        run<0, 1, 0, 0>("SS  N=128 no-swizzle                     ", grid, iters, out);
        run<0, 1, 1, 0>("SS  N=128 SWIZZLE_128B                   ", grid, iters, out);
        run<1, 1, 0, 0>("SS  N=256 no-swizzle                     ", grid, iters, out);
        run<1, 1, 1, 0>("SS  N=256 SWIZZLE_128B                   ", grid, iters, out);
        run<0, 0, 0, 0, 1>("TS  N=128 + epilogue warps reading TMEM  ", grid, iters, out);
        run<0, 1, 0, 0, 1>("SS  N=128 + epilogue warps reading TMEM  ", grid, iters, out);
        run<1, 0, 0, 0, 1>("TS  N=256 + epilogue warps reading TMEM  ", grid, iters, out);
        run<0, 0, 0, 0, 2>("TS  N=128 + producer warps writing TMEM  ", grid, iters, out);
        run<1, 0, 0, 0, 2>("TS  N=256 + producer warps writing TMEM  ", grid, iters, out);
        run<0, 0, 0, 0, 3>("TS  N=128 + bulk copies into smem        ", grid, iters, out);
        run<0, 1, 0, 0, 3>("SS  N=128 + bulk copies into smem        ", grid, iters, out);
        run<1, 0, 0, 0, 3>("TS  N=256 + bulk copies into smem        ", grid, iters, out);
    }
    return 0;
}
