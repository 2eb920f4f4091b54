// N2: the colour decoder's final 1x1 convolution fused with the discretized-mixture-of-logistics likelihood.
//
//   reference:  raw = Conv2dWeightNorm(hidden -> 10*nmix, 1x1)(ELU(h))       image_decoders/pixelcnnpp.py:32-34
//               raw = Conv2dWeightNorm(48 -> 10*nmix, 1x1)(h)                image_decoders/resnet.py:85
//               err = ColorImageDecoder.reconstruct_error(raw, x)            color_image_decoder.py:91-125, utils.py:61-123
//
// The reference writes raw [N, 10*nmix, H, W] to HBM (400 B / pixel at nmix = 10) and reads it back in the likelihood; here
// the raw parameters of a pixel never leave the SM.  Per tile of 128 pixels of one sample:
//
//     producers (2 warpgroups)   h[n, c, p0 + t] (coalesced 128 B lines per channel) -> ELU -> 3xTF32 split -> shared memory
//                                in the K-major canonical UMMA layout, 32 channels per pipeline stage (mbarrier ring)
//     MMA issuer (1 thread)      D[128 pixels x 112] (+)= A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T   tcgen05.mma kind::tf32,
//                                M = 128, N = 112, K = 8 per instruction; W (hi / lo split, rows permuted so that the 10
//                                parameters of a mixture component are adjacent) stays resident in shared memory;
//                                accumulators in TMEM, one 128-column buffer per epilogue warpgroup
//     epilogue (3 warpgroups)    thread = pixel = TMEM lane: tcgen05.ld of a component's 10 columns, + bias, DMOL evaluation
//                                (hot_math.cuh, the same arithmetic as dmol.cu), online log-sum-exp, tile sum -> per-sample
//                                ordered reduction.  GRAD: a second sweep over the TMEM-resident parameters emits
//                                d err / d raw (the only HBM write), for the library GEMMs of the convolution's backward.
//
// HBM traffic per pixel: 4*C bytes of h (+ 12 of x) instead of 4*C + 2 * 400 (conv read + raw write + raw read).
// 3xTF32 keeps the contraction at fp32-level accuracy (1e-5 parity tolerance); the tensor pipe is ~2000 cycles per tile at
// C = 96, hidden behind the epilogue's SFU / issue-bound evaluation.
#include "common.cuh"
#include "hot_math.cuh"
#include "tc5.cuh"

namespace mae {
namespace {

template <int NMIX, int C>
struct HeadGeo {
    static constexpr int kNOut = 10 * NMIX;                         // raw channels
    static constexpr int kNMma = (kNOut + 15) / 16 * 16;            // MMA N (multiple of 16 for M = 128)
    static constexpr int kAccCols = kNMma <= 64 ? 64 : 128;         // TMEM columns per accumulator buffer
    static constexpr int kKC = (C % 32 == 0) ? 32 : 16;             // channels per pipeline stage
    static constexpr int kChunks = C / kKC;
    static constexpr int kProdWG = 2, kEpiWG = 3;
    static constexpr int kThreads = (kProdWG + kEpiWG) * 128 + 32;  // + the MMA / TMEM-allocator warp
    static constexpr int kTmemCols = kEpiWG * kAccCols <= 256 ? 256 : 512;
    static constexpr int kLboA = 16 * 128;                          // 128 rows = 16 core matrices along M
    static constexpr int kLboW = (kNMma / 8) * 128;
    static constexpr int kSbo = 128;
    static constexpr int kWBytes = kNMma * C * 4;                   // one of W_hi / W_lo
    static constexpr int kStageBytes = 2 * 128 * kKC * 4;           // A_hi | A_lo
    static constexpr int kStages = kKC == 32 ? 4 : 6;
    static constexpr int kBiasOff = 2 * kWBytes + kStages * kStageBytes;
    static constexpr int kBarOff = kBiasOff + 128 * 4;
    static constexpr int kNumBars = 2 * kStages + 2 * kEpiWG;
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kRedOff = kSlotOff + 16;
    static constexpr int kSmemBytes = kRedOff + kEpiWG * 2 * 4 * 4 + 16;
    static_assert(C % 16 == 0 && C >= 16 && C <= 128, "hidden channels: a multiple of 16, at most 128");
    static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");
};

// TMEM / W-row column `col` (component-major: 10 * m + {logit, mu0..2, rho0..2, kap0..2}) -> raw channel of the reference's
// layout [logits(nmix) | mu(nmix,3) | log_scale(nmix,3) | coeffs(nmix,3)]   (color_image_decoder.py:44-50); -1: padding
template <int NMIX>
__device__ __forceinline__ int raw_channel(int col) {
    if (col >= 10 * NMIX) return -1;
    const int m = col / 10, p = col % 10;
    if (p == 0) return m;
    const int grp = (p - 1) / 3, c = (p - 1) % 3;    // grp 0: mu, 1: log-scale, 2: coefficient
    return (1 + 3 * grp) * NMIX + 3 * m + c;
}

template <int NMIX, int C, bool GRAD>
__global__ void __launch_bounds__(HeadGeo<NMIX, C>::kThreads, 1)
dmol_head_kernel(const float* __restrict__ hid, const float* __restrict__ wgt, const float* __restrict__ bias,
                 const float* __restrict__ x, int ns, int HW, int tps, int ntiles, int act, float g_mul,
                 float* __restrict__ err, float* __restrict__ draw, float* partial, unsigned int* tickets) {
    using G = HeadGeo<NMIX, C>;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + G::kWBytes;
    uint8_t* stages = smem + 2 * G::kWBytes;
    float* bias_s = reinterpret_cast<float*>(smem + G::kBiasOff);
    uint64_t* full = reinterpret_cast<uint64_t*>(smem + G::kBarOff);
    uint64_t* empty = full + G::kStages;
    uint64_t* tfull = empty + G::kStages;
    uint64_t* tempty = tfull + G::kEpiWG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G::kSlotOff);
    float* red = reinterpret_cast<float*>(smem + G::kRedOff);   // [kEpiWG][2][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr int kEpiWarps = 4 * G::kEpiWG, kProdWarps = 4 * G::kProdWG;

    // ---------------------------------------------------------------- one-time setup
    // W (row-permuted, 3xTF32 split) into the canonical layout; consecutive threads -> consecutive rows -> contiguous 16 B
    for (int idx = threadIdx.x; idx < G::kNMma * (C / 4); idx += G::kThreads) {
        const int n = idx % G::kNMma, kg = idx / G::kNMma;
        const int src = raw_channel<NMIX>(n);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) v = __ldg(reinterpret_cast<const float4*>(wgt + (int64_t)src * C + 4 * kg));
        float4 h, l;
        tc5::split_tf32(v.x, h.x, l.x);
        tc5::split_tf32(v.y, h.y, l.y);
        tc5::split_tf32(v.z, h.z, l.z);
        tc5::split_tf32(v.w, h.w, l.w);
        const int off = kg * G::kLboW + n * 16;
        *reinterpret_cast<float4*>(w_hi + off) = h;
        *reinterpret_cast<float4*>(w_lo + off) = l;
    }
    if (threadIdx.x < 128) {
        const int src = raw_channel<NMIX>(threadIdx.x);
        bias_s[threadIdx.x] = (src >= 0 && bias != nullptr) ? __ldg(bias + src) : 0.f;
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < G::kStages; ++s) {
            tc5::mbar_init(&full[s], 4);     // the four warps of the producing warpgroup
            tc5::mbar_init(&empty[s], 1);    // tcgen05.commit
        }
        for (int a = 0; a < G::kEpiWG; ++a) {
            tc5::mbar_init(&tfull[a], 1);    // tcgen05.commit
            tc5::mbar_init(&tempty[a], 4);   // the four warps of the epilogue warpgroup
        }
        tc5::fence_barrier_init();
    }
    if (warp == kEpiWarps + kProdWarps) tc5::tmem_alloc(tmem_slot, G::kTmemCols);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < kEpiWarps) {
        // ------------------------------------------------------------ epilogue: thread = pixel = TMEM lane
        const int wg = warp >> 2, q = warp & 3;
        const int t = (q << 5) | lane;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(wg * G::kAccCols);
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            if (it % G::kEpiWG != wg) continue;
            const int use = it / G::kEpiWG;
            const int n = tile / tps, p0 = (tile - n * tps) * 128;
            const int b = n / ns;
            float xs[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) xs[c] = __ldg(x + ((int64_t)b * 3 + c) * HW + p0 + t);
            tc5::mbar_wait(&tfull[wg], use & 1);
            tc5::fence_after_sync();

            hm::Lse lt, ll;
            lt.init();
            ll.init();
#pragma unroll 1
            for (int m = 0; m < NMIX; ++m) {
                float v[10];
                tc5::tmem_ld8(tacc + 10 * m, v);
                tc5::tmem_ld2(tacc + 10 * m + 8, v + 8);
                tc5::tmem_ld_wait();
#pragma unroll
                for (int i = 0; i < 10; ++i) v[i] += bias_s[10 * m + i];
                const float T = hm::dmol_component<false>(v + 1, v + 4, v + 7, xs, nullptr, nullptr, nullptr);
                lt.push(T + v[0]);
                ll.push(v[0]);
            }
            const float zt = lt.value(), zl = ll.value();
            if (GRAD) {
                // second sweep over the TMEM-resident parameters: responsibilities and local gradients -> d err / d raw
                float* ob = draw + (int64_t)n * G::kNOut * HW + p0 + t;
#pragma unroll 1
                for (int m = 0; m < NMIX; ++m) {
                    float v[10];
                    tc5::tmem_ld8(tacc + 10 * m, v);
                    tc5::tmem_ld2(tacc + 10 * m + 8, v + 8);
                    tc5::tmem_ld_wait();
#pragma unroll
                    for (int i = 0; i < 10; ++i) v[i] += bias_s[10 * m + i];
                    float gmu[3], grho[3], gkap[3];
                    const float T = hm::dmol_component<true>(v + 1, v + 4, v + 7, xs, gmu, grho, gkap);
                    const float wgt_m = hm::fast_ex2(hm::kLog2e * (T + v[0] - zt));   // responsibility of the component
                    const float pri = hm::fast_ex2(hm::kLog2e * (v[0] - zl));         // softmax(logits)
                    const float a = -g_mul * wgt_m;
                    stg_stream(ob + (int64_t)m * HW, g_mul * (pri - wgt_m));
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        float* o = ob + (int64_t)(3 * m + c) * HW;
                        stg_stream(o + (int64_t)NMIX * HW, a * gmu[c]);
                        stg_stream(o + (int64_t)4 * NMIX * HW, a * grho[c]);
                        stg_stream(o + (int64_t)7 * NMIX * HW, a * gkap[c]);
                    }
                }
            }
            // the accumulator buffer may be overwritten by the MMA of this warpgroup's next tile
            tc5::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(&tempty[wg]);

            // tile sum -> ordered per-sample reduction (bit-reproducible: partials are added in tile order)
            float val = warp_sum(zl - zt);
            float* r = red + (wg * 2 + (use & 1)) * 4;
            if (lane == 0) r[q] = val;
            tc5::named_sync(1 + wg, 128);
            if (t == 0) {
                const float tot = (r[0] + r[1]) + (r[2] + r[3]);
                if (tps == 1) {
                    err[n] = tot;
                } else {
                    partial[tile] = tot;
                    __threadfence();
                    const unsigned int tk = atomicAdd(&tickets[n], 1u);
                    if (tk == (unsigned int)(tps - 1)) {
                        __threadfence();
                        float s = 0.f;
                        for (int c = 0; c < tps; ++c) s += __ldcg(partial + (int64_t)n * tps + c);
                        err[n] = s;
                        tickets[n] = 0u;   // self-reset for the next call
                    }
                }
            }
        }
    } else if (warp < kEpiWarps + kProdWarps) {
        // ------------------------------------------------------------ producers: h -> ELU -> hi / lo -> canonical smem
        const int pw = (warp - kEpiWarps) >> 2;
        const int t = (((warp - kEpiWarps) & 3) << 5) | lane;
        int qn = 0;   // running chunk counter of this CTA (same sequence in the MMA issuer)
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            const int n = tile / tps, p0 = (tile - n * tps) * 128;
            const float* src = hid + (int64_t)n * C * HW + p0 + t;
            for (int kc = 0; kc < G::kChunks; ++kc, ++qn) {
                if (qn % G::kProdWG != pw) continue;
                float v[G::kKC];
#pragma unroll
                for (int j = 0; j < G::kKC; ++j) v[j] = ldg_stream(src + (int64_t)(kc * G::kKC + j) * HW);
                const int stage = qn % G::kStages;
                const uint32_t ph = (uint32_t)(qn / G::kStages) & 1u;
                tc5::mbar_wait(&empty[stage], ph ^ 1u);
                tc5::fence_after_sync();
                uint8_t* a_hi = stages + stage * G::kStageBytes + t * 16;
                uint8_t* a_lo = a_hi + G::kStageBytes / 2;
#pragma unroll
                for (int kg = 0; kg < G::kKC / 4; ++kg) {
                    float h[4], l[4];
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        float u = v[4 * kg + i];
                        if (act) u = u > 0.f ? u : hm::fast_ex2(hm::kLog2e * u) - 1.0f;   // ELU(alpha = 1)
                        tc5::split_tf32(u, h[i], l[i]);
                    }
                    *reinterpret_cast<float4*>(a_hi + kg * G::kLboA) = make_float4(h[0], h[1], h[2], h[3]);
                    *reinterpret_cast<float4*>(a_lo + kg * G::kLboA) = make_float4(l[0], l[1], l[2], l[3]);
                }
                tc5::fence_proxy_async();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(&full[stage]);
            }
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        constexpr uint32_t idesc = tc5::idesc_tf32(128, G::kNMma);
        const uint32_t w_hi_a = tc5::smem_u32(w_hi), w_lo_a = tc5::smem_u32(w_lo), st_a = tc5::smem_u32(stages);
        int qn = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int a = it % G::kEpiWG, use = it / G::kEpiWG;
            tc5::mbar_wait(&tempty[a], ((uint32_t)use & 1u) ^ 1u);
            tc5::fence_after_sync();
            const uint32_t d = tmem_base + (uint32_t)(a * G::kAccCols);
            for (int kc = 0; kc < G::kChunks; ++kc, ++qn) {
                const int stage = qn % G::kStages;
                const uint32_t ph = (uint32_t)(qn / G::kStages) & 1u;
                tc5::mbar_wait(&full[stage], ph);
                tc5::fence_after_sync();
                const uint32_t a_hi = st_a + stage * G::kStageBytes, a_lo = a_hi + G::kStageBytes / 2;
#pragma unroll
                for (int j = 0; j < G::kKC / 8; ++j) {
                    const int ks = kc * (G::kKC / 8) + j;   // global k-step
                    const uint64_t ah = tc5::smem_desc(a_hi + 2 * j * G::kLboA, G::kLboA, G::kSbo);
                    const uint64_t al = tc5::smem_desc(a_lo + 2 * j * G::kLboA, G::kLboA, G::kSbo);
                    const uint64_t bh = tc5::smem_desc(w_hi_a + 2 * ks * G::kLboW, G::kLboW, G::kSbo);
                    const uint64_t bl = tc5::smem_desc(w_lo_a + 2 * ks * G::kLboW, G::kLboW, G::kSbo);
                    tc5::mma_tf32(d, al, bh, idesc, (kc | j) != 0);
                    tc5::mma_tf32(d, ah, bl, idesc, 1u);
                    tc5::mma_tf32(d, ah, bh, idesc, 1u);
                }
                tc5::mma_commit(&empty[stage]);
            }
            tc5::mma_commit(&tfull[a]);
        }
    }

    // ---------------------------------------------------------------- teardown
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == kEpiWarps + kProdWarps) {
        tc5::fence_after_sync();
        tc5::tmem_dealloc(tmem_base, G::kTmemCols);
    }
}

struct HeadWs {
    int64_t partial_off, ticket_off, total;
};
HeadWs head_ws(int64_t N, int64_t HW) {
    HeadWs w;
    w.partial_off = 0;
    w.ticket_off = round_up(N * ceil_div(HW, 128) * (int64_t)sizeof(float), 256);
    w.total = w.ticket_off + round_up(N * (int64_t)sizeof(unsigned int), 256);
    return w;
}

template <int NMIX, int C>
int launch_head(const float* hid, const float* wgt, const float* bias, const float* x, int64_t N, int ns, int64_t HW, int act,
                float g_mul, float* err, float* draw, float* partial, unsigned int* tickets, cudaStream_t s) {
    using G = HeadGeo<NMIX, C>;
    const int tps = (int)(HW / 128);
    const int64_t ntiles = N * tps;
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    auto go = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        kernel<<<grid, G::kThreads, G::kSmemBytes, s>>>(hid, wgt, bias, x, ns, (int)HW, tps, (int)ntiles, act, g_mul, err, draw,
                                                       partial, tickets);
        return launch_status();
    };
    return draw != nullptr ? go(dmol_head_kernel<NMIX, C, true>) : go(dmol_head_kernel<NMIX, C, false>);
}

}  // namespace
}  // namespace mae

extern "C" int mae_dmol_head_supported(int64_t C, int64_t nmix, int64_t HW) {
    const bool shape = (nmix == 10 && (C == 96 || C == 48 || C == 64)) || (nmix == 5 && C == 64);
    return (shape && HW > 0 && HW % 128 == 0) ? 1 : 0;
}

extern "C" int64_t mae_dmol_head_workspace_bytes(int64_t N, int64_t HW) {
    if (N <= 0 || HW <= 0) return 0;
    return mae::head_ws(N, HW).total;
}

extern "C" int mae_dmol_head(const float* hidden, const float* weight, const float* bias, const float* x, int64_t B, int64_t ns,
                             int64_t nmix, int64_t C, int64_t HW, int act, float g_mul, float* err, float* draw, void* workspace,
                             int64_t workspace_bytes, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(hidden && weight && x && err && workspace, MAE_ERR_NULL);
    MAE_REQUIRE(B > 0 && ns > 0 && B * ns * (HW / 128) < (1ll << 31), MAE_ERR_SHAPE);
    MAE_REQUIRE(mae_dmol_head_supported(C, nmix, HW), MAE_ERR_UNSUPPORTED);
    MAE_REQUIRE((reinterpret_cast<uintptr_t>(weight) & 15) == 0, MAE_ERR_ALIGN);
    const int64_t N = B * ns;
    const HeadWs w = head_ws(N, HW);
    MAE_REQUIRE(workspace_bytes >= w.total, MAE_ERR_WORKSPACE);
    float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.partial_off);
    unsigned int* tickets = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + w.ticket_off);
    cudaStream_t s = as_stream(stream);
#define MAE_HEAD_CASE(NM, CC) \
    if (nmix == NM && C == CC) return launch_head<NM, CC>(hidden, weight, bias, x, N, (int)ns, HW, act, g_mul, err, draw, partial, tickets, s)
    MAE_HEAD_CASE(10, 96);
    MAE_HEAD_CASE(10, 48);
    MAE_HEAD_CASE(10, 64);
    MAE_HEAD_CASE(5, 64);
#undef MAE_HEAD_CASE
    return MAE_ERR_UNSUPPORTED;
}
