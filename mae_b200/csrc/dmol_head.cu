// N2: the colour decoder's final 1x1 convolution fused with the discretized-mixture-of-logistics likelihood.
//
//   reference:  raw = Conv2dWeightNorm(hidden -> 10*nmix, 1x1)(ELU(h))       image_decoders/pixelcnnpp.py:32-34
//               raw = Conv2dWeightNorm(48 -> 10*nmix, 1x1)(h)                image_decoders/resnet.py:85
//               err = ColorImageDecoder.reconstruct_error(raw, x)            color_image_decoder.py:91-125, utils.py:61-123
//
// The reference writes raw [N, 10*nmix, H, W] to HBM (400 B / pixel at nmix = 10) and reads it back in the likelihood; here
// the raw parameters of a pixel never leave the SM.  Per tile of 128 pixels of one sample:
//
//     loader (1 thread)          h[n, c0 .. c0 + 31, p0 .. p0 + 127]: ONE TMA box (cp.async.bulk.tensor.3d over a tensor map of
//                                [N, C, HW]) per chunk into an fp32 staging ring [channel][pixel], mbarrier transaction
//                                counts: 64 KB in flight per SM independent of the other warps' progress.  (Per-channel
//                                512-byte cp.async.bulk copies were measured first: ~78 cycles of issue each, 1.8 TB/s.)
//     converters (2 warpgroups)  staging -> ELU -> 3xTF32 split -> tcgen05.st into a TMEM operand stage (thread = pixel = lane,
//                                one tf32 per column).  A in shared memory (SS mode) was measured first: with K = 8 per
//                                instruction the tensor core re-reads 7.5 KB of operands per MMA and the kernel was
//                                shared-memory-bandwidth-bound (477 KB per tile); with A in TMEM only W is read from smem.
//     MMA issuer (1 thread)      D[128 pixels x 112] = 1 b^T + A_lo W_hi^T + A_hi W_lo^T + A_hi W_hi^T   tcgen05.mma
//                                kind::tf32, M = 128, N = 112, K = 8 per instruction; W (hi / lo split, rows permuted so that
//                                the 10 parameters of a mixture component are adjacent, the bias as one more k-step against a
//                                constant block of ones) stays resident in shared memory; accumulators in TMEM, one
//                                128-column buffer per epilogue warpgroup
//     epilogue (3 warpgroups)    thread = pixel = TMEM lane: tcgen05.ld of a component's 10 columns, DMOL evaluation
//                                (hot_math.cuh, the same arithmetic as dmol.cu), online log-sum-exp, tile sum -> per-sample
//                                ordered reduction.  GRAD: a second sweep over the TMEM-resident parameters emits
//                                d err / d raw (the only HBM write), for the library GEMMs of the convolution's backward.
//
// HBM traffic per pixel: 4*C bytes of h (+ 12 of x) instead of 4*C + 2 * 400 (conv read + raw write + raw read).
// 3xTF32 keeps the contraction at fp32-level accuracy (1e-5 parity tolerance); the tensor pipe is ~2000 cycles per tile at
// C = 96, hidden behind the epilogue's SFU / issue-bound evaluation.
#include <cuda.h>   // CUtensorMap (types only: the encoder is fetched through cudaGetDriverEntryPoint, no libcuda link)

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
    static constexpr int kKW = C + 8;                               // K extent of W: the extra k-step carries the bias
    static constexpr int kConvWG = 2, kEpiWG = 3;
    static constexpr int kEpiWarps = 4 * kEpiWG, kConvWarps = 4 * kConvWG;
    static constexpr int kLoaderWarp = kEpiWarps + kConvWarps, kMmaWarp = kLoaderWarp + 1;
    static constexpr int kThreads = (kMmaWarp + 1) * 32;
    static constexpr int kLboA = 16 * 128;                          // 128 rows = 16 core matrices along M
    static constexpr int kLboW = (kNMma / 8) * 128;
    static constexpr int kSbo = 128;
    static constexpr int kWBytes = kNMma * kKW * 4;                 // one of W_hi / W_lo
    static constexpr int kOnesBytes = 128 * 8 * 4;                  // A block of the bias k-step: column 0 = 1
    static constexpr int kAStages = kConvWG;                        // one TMEM operand stage per converter warpgroup
    static constexpr int kAStageCols = 2 * kKC;                     // A_hi | A_lo, one tf32 per column, row = TMEM lane
    static constexpr int kACol0 = kEpiWG * kAccCols;                // operand stages sit behind the accumulators
    static constexpr int kTmemUsed = kACol0 + kAStages * kAStageCols;
    static constexpr int kTmemCols = kTmemUsed <= 256 ? 256 : 512;
    static constexpr int kSStageBytes = kKC * 128 * 4;              // fp32 staging slot [channel][pixel]
    static constexpr int kSStages = 6;                              // ring filled by TMA
    static constexpr int kOnesOff = 2 * kWBytes;
    static constexpr int kSOff = kOnesOff + kOnesBytes;
    static constexpr int kBarOff = kSOff + kSStages * kSStageBytes;
    static constexpr int kNumBars = 2 * kSStages + 2 * kAStages + 2 * kEpiWG;
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kRedOff = kSlotOff + 16;
    static constexpr int kSmemBytes = kRedOff + kEpiWG * 2 * 4 * 4 + 16;
    static_assert(kTmemUsed <= 512, "TMEM columns");
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
dmol_head_kernel(const __grid_constant__ CUtensorMap hid_map, const float* __restrict__ wgt, const float* __restrict__ bias,
                 const float* __restrict__ x, int ns, int HW, int tps, int ntiles, int act, float g_mul,
                 float* __restrict__ err, float* __restrict__ draw, float* partial, unsigned int* tickets, long long* prof) {
    using G = HeadGeo<NMIX, C>;
    // prof != nullptr (MAE_HEAD_PROF=1, profiling runs only): CTA 0 reports, per role, the cycles spent in its loop and in
    // each of its barrier waits -> which role is the pipeline's bottleneck
    long long pw0 = 0, pw1 = 0;
    auto wait = [&](uint64_t* bar, uint32_t parity, long long& acc) {
        if (prof != nullptr) {
            const long long t0 = clock64();
            tc5::mbar_wait(bar, parity);
            acc += clock64() - t0;
        } else {
            tc5::mbar_wait(bar, parity);
        }
    };
    const long long t_start = prof != nullptr ? clock64() : 0;
    auto report = [&](int slot) {
        if (prof != nullptr && blockIdx.x == 0) {
            prof[slot] = clock64() - t_start;
            prof[slot + 1] = pw0;
            prof[slot + 2] = pw1;
        }
    };
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + G::kWBytes;
    uint8_t* ones = smem + G::kOnesOff;
    uint8_t* sstages = smem + G::kSOff;
    uint64_t* sfull = reinterpret_cast<uint64_t*>(smem + G::kBarOff);
    uint64_t* sempty = sfull + G::kSStages;
    uint64_t* afull = sempty + G::kSStages;
    uint64_t* aempty = afull + G::kAStages;
    uint64_t* tfull = aempty + G::kAStages;
    uint64_t* tempty = tfull + G::kEpiWG;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + G::kSlotOff);
    float* red = reinterpret_cast<float*>(smem + G::kRedOff);   // [kEpiWG][2][4]

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    // ---------------------------------------------------------------- one-time setup
    // W (row-permuted, 3xTF32 split) into the canonical layout; consecutive threads -> consecutive rows -> contiguous 16 B.
    // The k-group behind the C channels holds the bias in its first column (multiplied by the constant "ones" A block).
    for (int idx = threadIdx.x; idx < G::kNMma * (G::kKW / 4); idx += G::kThreads) {
        const int n = idx % G::kNMma, kg = idx / G::kNMma;
        const int src = raw_channel<NMIX>(n);
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (src >= 0) {
            if (kg < C / 4) v = __ldg(reinterpret_cast<const float4*>(wgt + (int64_t)src * C + 4 * kg));
            else if (kg == C / 4 && bias != nullptr) v.x = __ldg(bias + src);
        }
        float4 h, l;
        tc5::split_tf32(v.x, h.x, l.x);
        tc5::split_tf32(v.y, h.y, l.y);
        tc5::split_tf32(v.z, h.z, l.z);
        tc5::split_tf32(v.w, h.w, l.w);
        const int off = kg * G::kLboW + n * 16;
        *reinterpret_cast<float4*>(w_hi + off) = h;
        *reinterpret_cast<float4*>(w_lo + off) = l;
    }
    for (int idx = threadIdx.x; idx < 256; idx += G::kThreads)      // 128 rows x 2 k-groups of 16 bytes
        *reinterpret_cast<float4*>(ones + idx * 16) = make_float4(idx < 128 ? 1.f : 0.f, 0.f, 0.f, 0.f);
    if (threadIdx.x == 0) {
        for (int s = 0; s < G::kSStages; ++s) {
            tc5::mbar_init(&sfull[s], 1);    // the loader's arrive.expect_tx; completed by the bulk copies' bytes
            tc5::mbar_init(&sempty[s], 4);   // the four warps of the converting warpgroup
        }
        for (int s = 0; s < G::kAStages; ++s) {
            tc5::mbar_init(&afull[s], 4);    // the four warps of the converting warpgroup
            tc5::mbar_init(&aempty[s], 1);   // tcgen05.commit
        }
        for (int a = 0; a < G::kEpiWG; ++a) {
            tc5::mbar_init(&tfull[a], 1);    // tcgen05.commit
            tc5::mbar_init(&tempty[a], 4);   // the four warps of the epilogue warpgroup
        }
        tc5::fence_barrier_init();
    }
    if (warp == G::kMmaWarp) tc5::tmem_alloc(tmem_slot, G::kTmemCols);
    tc5::fence_proxy_async();
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < G::kEpiWarps) {
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
            wait(&tfull[wg], use & 1, pw0);
            tc5::fence_after_sync();

            hm::Lse lt, ll;
            lt.init();
            ll.init();
            // two components per iteration: their evaluations are independent instruction streams the scheduler can interleave
            // (one pixel per thread and 3 warps per scheduler leave the SFU / FMA latencies exposed otherwise)
#pragma unroll 1
            for (int m = 0; m + 1 < NMIX; m += 2) {
                float v[20];
                tc5::tmem_ld16(tacc + 10 * m, v);
                tc5::tmem_ld4(tacc + 10 * m + 16, v + 16);
                tc5::tmem_ld_wait();
                const float T0 = hm::dmol_component<false>(v + 1, v + 4, v + 7, xs, nullptr, nullptr, nullptr);
                const float T1 = hm::dmol_component<false>(v + 11, v + 14, v + 17, xs, nullptr, nullptr, nullptr);
                lt.push(T0 + v[0]);
                ll.push(v[0]);
                lt.push(T1 + v[10]);
                ll.push(v[10]);
            }
            if (NMIX & 1) {
                float v[10];
                tc5::tmem_ld8(tacc + 10 * (NMIX - 1), v);
                tc5::tmem_ld2(tacc + 10 * (NMIX - 1) + 8, v + 8);
                tc5::tmem_ld_wait();
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
        if (t == 0) report(4 * wg);
    } else if (warp < G::kLoaderWarp) {
        // ------------------------------------------------------------ converters: staging -> ELU -> hi / lo -> canonical smem
        const int pw = (warp - G::kEpiWarps) >> 2;
        const int t = ((warp & 3) << 5) | lane;
        // this thread's row (TMEM lane) of its warpgroup's operand stage: A_hi in columns [0, kKC), A_lo in [kKC, 2 kKC)
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(G::kACol0 + pw * G::kAStageCols);
        int qn = 0;   // running chunk counter of this CTA (the same sequence in the loader and in the MMA issuer)
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
            for (int kc = 0; kc < G::kChunks; ++kc, ++qn) {
                if (qn % G::kConvWG != pw) continue;
                const int slot = qn % G::kSStages;
                wait(&sfull[slot], (uint32_t)(qn / G::kSStages) & 1u, pw0);
                const float* st = reinterpret_cast<const float*>(sstages + slot * G::kSStageBytes) + t;
                float h[G::kKC], l[G::kKC];
#pragma unroll
                for (int j = 0; j < G::kKC; ++j) h[j] = st[j * 128];
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(&sempty[slot]);     // the slot may be refilled
#pragma unroll
                for (int j = 0; j < G::kKC; ++j) {
                    float u = h[j];
                    if (act) u = u > 0.f ? u : hm::fast_ex2(hm::kLog2e * u) - 1.0f;   // ELU(alpha = 1)
                    tc5::split_tf32_fast(u, h[j], l[j]);      // both parts rounded to tf32 (a truncated lo would bias the sum)
                }
                wait(&aempty[pw], ((uint32_t)(qn / G::kConvWG) & 1u) ^ 1u, pw1);   // the MMAs that read this stage are done
                tc5::fence_after_sync();
#pragma unroll
                for (int j = 0; j < G::kKC; j += 16) {
                    tc5::tmem_st16(ta + j, h + j);
                    tc5::tmem_st16(ta + G::kKC + j, l + j);
                }
                tc5::tmem_st_wait();
                tc5::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(&afull[pw]);
            }
        }
        if (t == 0) report(12 + 4 * pw);
    } else if (warp == G::kLoaderWarp) {
        // ------------------------------------------------------------ loader: one TMA box [kKC channels x 128 pixels] per chunk
        if (lane == 0) {
            tc5::prefetch_tmap(&hid_map);
            int qn = 0;
            for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
                const int n = tile / tps, p0 = (tile - n * tps) * 128;
                for (int kc = 0; kc < G::kChunks; ++kc, ++qn) {
                    const int slot = qn % G::kSStages;
                    wait(&sempty[slot], ((uint32_t)(qn / G::kSStages) & 1u) ^ 1u, pw0);
                    tc5::mbar_expect_tx(&sfull[slot], G::kSStageBytes);
                    tc5::tma_load_3d(sstages + slot * G::kSStageBytes, &hid_map, p0, kc * G::kKC, n, &sfull[slot]);
                }
            }
        }
        if (lane == 0) report(20);
    } else if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer (one thread)
        constexpr uint32_t idesc = tc5::idesc_tf32(128, G::kNMma);
        const uint32_t w_hi_a = tc5::smem_u32(w_hi), w_lo_a = tc5::smem_u32(w_lo);
        const uint64_t one_d = tc5::smem_desc(tc5::smem_u32(ones), G::kLboA, G::kSbo);
        const uint64_t bias_h = tc5::smem_desc(w_hi_a + 2 * (C / 8) * G::kLboW, G::kLboW, G::kSbo);
        const uint64_t bias_l = tc5::smem_desc(w_lo_a + 2 * (C / 8) * G::kLboW, G::kLboW, G::kSbo);
        int qn = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int a = it % G::kEpiWG, use = it / G::kEpiWG;
            wait(&tempty[a], ((uint32_t)use & 1u) ^ 1u, pw0);
            tc5::fence_after_sync();
            const uint32_t d = tmem_base + (uint32_t)(a * G::kAccCols);
            tc5::mma_tf32(d, one_d, bias_l, idesc, 0u);     // D = 1 * bias: initialises the accumulator (A from smem)
            tc5::mma_tf32(d, one_d, bias_h, idesc, 1u);
            for (int kc = 0; kc < G::kChunks; ++kc, ++qn) {
                const int stage = qn % G::kAStages;
                wait(&afull[stage], (uint32_t)(qn / G::kAStages) & 1u, pw1);
                tc5::fence_after_sync();
                const uint32_t a_hi = tmem_base + (uint32_t)(G::kACol0 + stage * G::kAStageCols), a_lo = a_hi + G::kKC;
#pragma unroll
                for (int j = 0; j < G::kKC / 8; ++j) {
                    const int ks = kc * (G::kKC / 8) + j;   // global k-step
                    const uint64_t bh = tc5::smem_desc(w_hi_a + 2 * ks * G::kLboW, G::kLboW, G::kSbo);
                    const uint64_t bl = tc5::smem_desc(w_lo_a + 2 * ks * G::kLboW, G::kLboW, G::kSbo);
                    tc5::mma_tf32_ts(d, a_lo + 8 * j, bh, idesc, 1u);
                    tc5::mma_tf32_ts(d, a_hi + 8 * j, bl, idesc, 1u);
                    tc5::mma_tf32_ts(d, a_hi + 8 * j, bh, idesc, 1u);
                }
                tc5::mma_commit(&aempty[stage]);
            }
            tc5::mma_commit(&tfull[a]);
        }
        report(24);
    }

    // ---------------------------------------------------------------- teardown
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == G::kMmaWarp) {
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

typedef CUresult (*TmapEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
TmapEncodeFn tmap_encoder() {
    static TmapEncodeFn fn = []() -> TmapEncodeFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<TmapEncodeFn>(p);
    }();
    return fn;
}

template <int NMIX, int C>
int launch_head(const float* hid, const float* wgt, const float* bias, const float* x, int64_t N, int ns, int64_t HW, int act,
                float g_mul, float* err, float* draw, float* partial, unsigned int* tickets, long long* prof, cudaStream_t s) {
    using G = HeadGeo<NMIX, C>;
    const int tps = (int)(HW / 128);
    const int64_t ntiles = N * tps;
    const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
    // tensor map of hidden [N][C][HW] fp32 (innermost first); one box = 128 pixels x kKC channels of one sample
    TmapEncodeFn enc = tmap_encoder();
    if (enc == nullptr) return MAE_ERR_UNSUPPORTED;
    CUtensorMap map;
    const cuuint64_t gdim[3] = {(cuuint64_t)HW, (cuuint64_t)C, (cuuint64_t)N};
    const cuuint64_t gstr[2] = {(cuuint64_t)HW * 4, (cuuint64_t)HW * 4 * C};
    const cuuint32_t box[3] = {128, (cuuint32_t)G::kKC, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(hid), gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
            CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return MAE_ERR_UNSUPPORTED;
    auto go = [&](auto kernel) -> int {
        cudaError_t e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, G::kSmemBytes);
        if (e != cudaSuccess) return (int)e;
        kernel<<<grid, G::kThreads, G::kSmemBytes, s>>>(map, wgt, bias, x, ns, (int)HW, tps, (int)ntiles, act, g_mul, err, draw,
                                                       partial, tickets, prof);
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
    MAE_REQUIRE((reinterpret_cast<uintptr_t>(weight) & 15) == 0 && (reinterpret_cast<uintptr_t>(hidden) & 15) == 0, MAE_ERR_ALIGN);
    const int64_t N = B * ns;
    const HeadWs w = head_ws(N, HW);
    MAE_REQUIRE(workspace_bytes >= w.total, MAE_ERR_WORKSPACE);
    float* partial = reinterpret_cast<float*>(static_cast<char*>(workspace) + w.partial_off);
    unsigned int* tickets = reinterpret_cast<unsigned int*>(static_cast<char*>(workspace) + w.ticket_off);
    cudaStream_t s = as_stream(stream);
    // profiling runs: MAE_HEAD_PROF=1 and a workspace with 256 spare bytes behind the tickets -> per-role cycle counters
    static const bool want_prof = getenv("MAE_HEAD_PROF") != nullptr;
    long long* prof = (want_prof && workspace_bytes >= w.total + 256) ? reinterpret_cast<long long*>(static_cast<char*>(workspace) + w.total)
                                                                      : nullptr;
#define MAE_HEAD_CASE(NM, CC) \
    if (nmix == NM && C == CC) return launch_head<NM, CC>(hidden, weight, bias, x, N, (int)ns, HW, act, g_mul, err, draw, partial, tickets, prof, s)
    MAE_HEAD_CASE(10, 96);
    MAE_HEAD_CASE(10, 48);
    MAE_HEAD_CASE(10, 64);
    MAE_HEAD_CASE(5, 64);
#undef MAE_HEAD_CASE
    return MAE_ERR_UNSUPPORTED;
}
