// Tensor-core tiles of the MPD regulariser for large batches (C5: B >= 2048): the K = 2D contraction of the forward
//   2 KL_ij = x~_i . y~_j + E_i + F_j
// and the partner (K = B) contractions of the stored-KL backward
//   G1_i = sum_j (KL_ij - m) y~_j ,  G2_i = sum_j (KL_ji - m) x~_j
// as 3xTF32 tcgen05.mma (fp32-level accuracy: every operand is split into tf32(v) and v - tf32(v), three products, fp32
// accumulation in TMEM; tc5.cuh).  Included by mpd.cu inside namespace mae::{anonymous} (shares Hdr / Layout).
//
// Both kernels are warp-specialised, one persistent CTA per SM:
//   A side     thread = tile row = TMEM lane: loads its 128 contiguous bytes of the chunk from global (L2-resident panels or
//              the stored KL rows), splits, tcgen05.st into a TMEM operand stage -> the tensor core reads only B from smem
//              (with K = 8 per tf32 instruction an SS-mode MMA re-reads 8 KB of operands per 64 cycles: the shared-memory
//              bandwidth, measured in dmol_head.cu)
//   B side     loads + splits the column panel chunk into the K-major canonical shared-memory layout
//   MMA        one thread: 3 tcgen05.mma (128 x 128 x 8) per k-step, tcgen05.commit -> mbarriers
//   epilogue   thread = row: tcgen05.ld of the accumulator, KL / weights arithmetic, stores
#pragma once

// (tc5.cuh is included by mpd.cu at file scope)

constexpr int kTcT = 128;            // tile edge (rows = TMEM lanes, columns = MMA N)
constexpr int kTcKC = 32;            // contraction elements per pipeline stage
constexpr int kTcLbo = 16 * 128;     // canonical layout: 128 rows x 16 B per k-group
constexpr int kTcStageB = 2 * kTcT * kTcKC * 4;   // B_hi | B_lo of one stage
constexpr int kTcStages = 2;

struct FwdTcSmem {
    static constexpr int kBOff = 0;
    static constexpr int kFOff = kTcStages * kTcStageB;            // f_j of the current tile, one copy per epilogue warpgroup
    static constexpr int kBarOff = kFOff + 2 * kTcT * 4;
    static constexpr int kNumBars = 3 * kTcStages + 4;             // afull, bfull, sempty per stage; tfull, tempty per accumulator
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kRedOff = kSlotOff + 16;
    static constexpr int kBytes = kRedOff + 8 * 8 + 16;
};
constexpr int kFwdTcThreads = (8 + 4 + 4 + 1) * 32;   // 2 epilogue warpgroups, A producer, B producer, MMA warp

// TMEM columns: accumulators at 0 and 128, A operand stages (hi | lo, kTcKC columns each) behind them
constexpr int kFwdTcACol = 256;
constexpr int kFwdTcTmemCols = 512;

__global__ void __launch_bounds__(kFwdTcThreads, 1)
mpd_fwd_tc_kernel(const float* __restrict__ Xp, const float* __restrict__ Yp, const float* __restrict__ evec,
                  const float* __restrict__ fvec, Hdr* hdr, double* part, int B, int Kp, int row0, int row1, int tr0, int ntr,
                  int ntc, float* __restrict__ KL, float* __restrict__ KLT, int Bp, double* __restrict__ sumsq_out,
                  float* __restrict__ S_out, int tc0, int kl_base) {
    using SM = FwdTcSmem;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* bst = smem + SM::kBOff;
    float* fs = reinterpret_cast<float*>(smem + SM::kFOff);
    uint64_t* afull = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
    uint64_t* bfull = afull + kTcStages;
    uint64_t* sempty = bfull + kTcStages;
    uint64_t* tfull = sempty + kTcStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::kSlotOff);
    double* red = reinterpret_cast<double*>(smem + SM::kRedOff);
    __shared__ bool is_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = ntr * ntc;
    const int nchunk = Kp / kTcKC;
    pdl_wait();                // panels, E / F and m come from mpd_prepare_kernel
    pdl_launch_dependents();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc5::mbar_init(&afull[s], 4);
            tc5::mbar_init(&bfull[s], 4);
            tc5::mbar_init(&sempty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc5::mbar_init(&tfull[a], 1);
            tc5::mbar_init(&tempty[a], 4);
        }
        tc5::fence_barrier_init();
    }
    if (warp == 16) tc5::tmem_alloc(tmem_slot, kFwdTcTmemCols);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ------------------------------------------------------------ epilogue: thread = row of the tile
        const int wg = warp >> 2, q = warp & 3;
        const int t = (q << 5) | lane;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(wg * kTcT);
        float* fw = fs + wg * kTcT;
        const float m = (float)hdr->m;
        double total = 0.0;
        int it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            if ((it & 1) != wg) continue;
            const int use = it >> 1;
            const int ti = tr0 + tile / ntc, tj = tc0 + tile % ntc;
            const int i = ti * kTcT + t;
            const float ei = evec[i];
            tc5::named_sync(1 + wg, 128);                 // the previous tile's readers of fw are done
            fw[t] = fvec[tj * kTcT + t];
            tc5::named_sync(1 + wg, 128);
            const bool row_ok = (i >= row0) && (i < row1);
            const bool plain = (ti != tj) && (tj * kTcT + kTcT <= B);      // no diagonal, no padded columns
            tc5::mbar_wait(&tfull[wg], use & 1);
            tc5::fence_after_sync();
            float psum = 0.f;
            float* klrow = KL != nullptr ? KL + (int64_t)(i - kl_base) * Bp + tj * kTcT : nullptr;
            float* kltcol = KLT != nullptr ? KLT + (int64_t)(tj * kTcT - kl_base) * Bp + ti * kTcT + t : nullptr;
#pragma unroll 1
            for (int c0 = 0; c0 < kTcT; c0 += 16) {
                float v[16];
                tc5::tmem_ld16(tacc + c0, v);
                tc5::tmem_ld_wait();
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float kl = 0.5f * (v[c] + ei + fw[c0 + c]);
                    v[c] = kl;
                    const float dev = kl - m;
                    const int j = tj * kTcT + c0 + c;
                    const bool ok = row_ok && (plain || (j < B && j != i));
                    psum = ok ? fmaf(dev, dev, psum) : psum;
                }
                if (klrow != nullptr) {
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4)
                        *reinterpret_cast<float4*>(klrow + c0 + 4 * g4) = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
                }
                if (kltcol != nullptr) {
#pragma unroll
                    for (int c = 0; c < 16; ++c) kltcol[(int64_t)(c0 + c) * Bp] = v[c];      // lanes = consecutive i: coalesced
                }
            }
            tc5::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(&tempty[wg]);
            total += (double)psum;
        }
        // CTA total over the 256 epilogue threads, fixed order
        total = warp_sum(total);
        if (lane == 0) red[warp] = total;
        tc5::named_sync(3, 256);
        if (sumsq_out != nullptr && threadIdx.x == 0) {
            double s = 0.0;
            for (int w = 0; w < 8; ++w) s += red[w];
            if (gridDim.x == 1) {
                is_last = true;
                part[0] = s;
            } else {
                part[blockIdx.x] = s;
                __threadfence();
                const unsigned int tk = atomicAdd(&hdr->ticket_fwd, 1u);
                is_last = (tk == gridDim.x - 1);
            }
        }
        tc5::named_sync(3, 256);
        if (sumsq_out != nullptr && is_last && warp == 0) {
            __threadfence();
            double s = 0.0;
            for (int k = lane; k < (int)gridDim.x; k += 32) s += __ldcg(part + k);
            s = warp_sum(s);
            if (lane == 0) {
                hdr->sumsq = s;
                sumsq_out[0] = s;
                if (S_out != nullptr) {
                    const double nb = (double)B;
                    S_out[0] = (float)sqrt(s / (nb * (nb - 1.0) - 1.0));
                }
                hdr->ticket_fwd = 0u;
            }
        }
    } else if (warp < 12) {
        // ------------------------------------------------------------ A producer: x~ rows -> hi / lo -> TMEM operand stage
        const int t = ((warp & 3) << 5) | lane;
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)kFwdTcACol;
        const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int total = my_tiles * nchunk;
        // chunk g of this CTA's sequence -> the 128 bytes of this thread's row; the loads of chunk g + 1 are in flight while
        // chunk g is split and stored (the panels are L2-resident: latency, not bandwidth)
        auto src_of = [&](int g) {
            const int tile = blockIdx.x + (g / nchunk) * gridDim.x;
            return Xp + (int64_t)((tr0 + tile / ntc) * kTcT + t) * Kp + (g % nchunk) * kTcKC;
        };
        float4 nxt[kTcKC / 4];
        if (total > 0) {
            const float* p = src_of(0);
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) nxt[g4] = __ldg(reinterpret_cast<const float4*>(p + 4 * g4));
        }
        for (int qn = 0; qn < total; ++qn) {
            float h[kTcKC], l[kTcKC];
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) {
                h[4 * g4] = nxt[g4].x; h[4 * g4 + 1] = nxt[g4].y; h[4 * g4 + 2] = nxt[g4].z; h[4 * g4 + 3] = nxt[g4].w;
            }
            if (qn + 1 < total) {
                const float* p = src_of(qn + 1);
#pragma unroll
                for (int g4 = 0; g4 < kTcKC / 4; ++g4) nxt[g4] = __ldg(reinterpret_cast<const float4*>(p + 4 * g4));
            }
#pragma unroll
            for (int j = 0; j < kTcKC; ++j) tc5::split_tf32(h[j], h[j], l[j]);     // both parts ROUNDED: a truncated lo biases long contractions
            const int s = qn % kTcStages;
            tc5::mbar_wait(&sempty[s], ((uint32_t)(qn / kTcStages) & 1u) ^ 1u);
            tc5::fence_after_sync();
            const uint32_t tas = ta + (uint32_t)(s * 2 * kTcKC);
#pragma unroll
            for (int j = 0; j < kTcKC; j += 16) {
                tc5::tmem_st16(tas + j, h + j);
                tc5::tmem_st16(tas + kTcKC + j, l + j);
            }
            tc5::tmem_st_wait();
            tc5::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(&afull[s]);
        }
    } else if (warp < 16) {
        // ------------------------------------------------------------ B producer: y~ rows -> hi / lo -> canonical smem
        const int t = ((warp & 3) << 5) | lane;
        const int my_tiles = blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
        const int total = my_tiles * nchunk;
        auto src_of = [&](int g) {
            const int tile = blockIdx.x + (g / nchunk) * gridDim.x;
            return Yp + (int64_t)((tc0 + tile % ntc) * kTcT + t) * Kp + (g % nchunk) * kTcKC;
        };
        float4 nxt[kTcKC / 4];
        if (total > 0) {
            const float* p = src_of(0);
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) nxt[g4] = __ldg(reinterpret_cast<const float4*>(p + 4 * g4));
        }
        for (int qn = 0; qn < total; ++qn) {
            float4 v[kTcKC / 4];
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) v[g4] = nxt[g4];
            if (qn + 1 < total) {
                const float* p = src_of(qn + 1);
#pragma unroll
                for (int g4 = 0; g4 < kTcKC / 4; ++g4) nxt[g4] = __ldg(reinterpret_cast<const float4*>(p + 4 * g4));
            }
            const int s = qn % kTcStages;
            tc5::mbar_wait(&sempty[s], ((uint32_t)(qn / kTcStages) & 1u) ^ 1u);
            tc5::fence_after_sync();
            uint8_t* b_hi = bst + s * kTcStageB + t * 16;
            uint8_t* b_lo = b_hi + kTcStageB / 2;
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) {
                float4 h, l;
                tc5::split_tf32(v[g4].x, h.x, l.x);
                tc5::split_tf32(v[g4].y, h.y, l.y);
                tc5::split_tf32(v[g4].z, h.z, l.z);
                tc5::split_tf32(v[g4].w, h.w, l.w);
                *reinterpret_cast<float4*>(b_hi + g4 * kTcLbo) = h;
                *reinterpret_cast<float4*>(b_lo + g4 * kTcLbo) = l;
            }
            tc5::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(&bfull[s]);
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc5::idesc_tf32(kTcT, kTcT);
        const uint32_t b_a = tc5::smem_u32(bst);
        int qn = 0, it = 0;
        for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x, ++it) {
            const int a = it & 1, use = it >> 1;
            tc5::mbar_wait(&tempty[a], ((uint32_t)use & 1u) ^ 1u);
            tc5::fence_after_sync();
            const uint32_t d = tmem_base + (uint32_t)(a * kTcT);
            for (int kc = 0; kc < nchunk; ++kc, ++qn) {
                const int s = qn % kTcStages;
                const uint32_t ph = (uint32_t)(qn / kTcStages) & 1u;
                tc5::mbar_wait(&afull[s], ph);
                tc5::mbar_wait(&bfull[s], ph);
                tc5::fence_after_sync();
                const uint32_t a_hi = tmem_base + (uint32_t)(kFwdTcACol + s * 2 * kTcKC), a_lo = a_hi + kTcKC;
                const uint32_t b_hi = b_a + s * kTcStageB, b_lo = b_hi + kTcStageB / 2;
#pragma unroll
                for (int j = 0; j < kTcKC / 8; ++j) {
                    const uint64_t bh = tc5::smem_desc(b_hi + 2 * j * kTcLbo, kTcLbo, 128);
                    const uint64_t bl = tc5::smem_desc(b_lo + 2 * j * kTcLbo, kTcLbo, 128);
                    tc5::mma_tf32_ts(d, a_lo + 8 * j, bh, idesc, (kc | j) != 0);
                    tc5::mma_tf32_ts(d, a_hi + 8 * j, bl, idesc, 1u);
                    tc5::mma_tf32_ts(d, a_hi + 8 * j, bh, idesc, 1u);
                }
                tc5::mma_commit(&sempty[s]);
            }
            tc5::mma_commit(&tfull[a]);
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 16) {
        tc5::fence_after_sync();
        tc5::tmem_dealloc(tmem_base, kFwdTcTmemCols);
    }
}

// ------------------------------------------------------------------------------------------------ backward, stored KL
// G1[i][k] = sum_j (KL_ij - m) y~_j[k],  G2[i][k] = sum_j (KL_ji - m) x~_j[k]  over a range of partner tiles, UNSCALED (the
// weight scale beta = g_S / ((B^2 - B - 1) S) is a global factor; mpd_bwd_tc_finish_kernel applies it together with the
// closed-form per-element terms).  CTA = (row tile of 128, 128 gradient columns, partner split); TMEM: G1 at columns 0,
// G2 at 128, weight stages W1 (hi | lo) at 256 + 64 s, W2 at 384 + 64 s.  The weights of a row are formed by the thread that
// owns the row (= TMEM lane) straight from the stored KL rows (128 contiguous bytes per chunk, streamed once per column
// chunk), so A never touches shared memory; B = the partner panels transposed on the fly into [k][j] (K-major) tiles.
constexpr int kBwdTcLbo = 16 * 128 + 16;          // +16 B: the transposing 4-byte stores of 32 partners hit 32 banks
constexpr int kBwdTcMat = (kTcKC / 4) * kBwdTcLbo;   // one of Y_hi, Y_lo, X_hi, X_lo of a stage
constexpr int kBwdTcStageB = 4 * kBwdTcMat;
struct BwdTcSmem {
    static constexpr int kBarOff = kTcStages * kBwdTcStageB;
    static constexpr int kNumBars = 5 * kTcStages + 2;    // w1full, w2full, yfull, xfull, sempty per stage; gfull, gdrained
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kBytes = kSlotOff + 32;
};
constexpr int kBwdTcThreads = (4 + 4 + 4 + 4 + 1) * 32;   // W1 rows, W2 rows, Y panel, X panel, MMA warp

__global__ void __launch_bounds__(kBwdTcThreads, 1)
mpd_bwd_tc_kernel(const float* __restrict__ Xp, const float* __restrict__ Yp, const Hdr* __restrict__ hdr, int B, int Kp,
                  int row0, int row1, int tr0, int ntc, int tiles_per_split, const float* __restrict__ KL,
                  const float* __restrict__ KLT, int Bp, float* __restrict__ gpart, float* __restrict__ rspart, int rows_pad,
                  int gbase, int group_chunks) {
    // group_chunks: chunks of 32 partners per TMEM accumulation.  The tensor core's accumulator does not round to nearest, so
    // the error of a running sum grows with its length; after every group the owners of the rows add the accumulator to the
    // CTA's fp32 partial sums in global memory (round to nearest) and the next group starts from zero.
    // KL / KLT: pre-offset so that global row i is at [i * Bp] (row blocks keep only their rows).  gpart
    // [split][2][rows_pad][Kp], rspart [split][2][rows_pad], local row = i - gbase.
    using SM = BwdTcSmem;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* w1full = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
    uint64_t* w2full = w1full + kTcStages;
    uint64_t* yfull = w2full + kTcStages;
    uint64_t* xfull = yfull + kTcStages;
    uint64_t* sempty = xfull + kTcStages;
    uint64_t* gfull = sempty + kTcStages;
    uint64_t* gdrained = gfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::kSlotOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ti = tr0 + blockIdx.x;
    const int kbase = blockIdx.y * kTcT;
    const int split = blockIdx.z;
    const int tj0 = split * tiles_per_split;
    const int tj1 = min(ntc, tj0 + tiles_per_split);
    const int total = (tj1 - tj0) * (kTcT / kTcKC);      // chunks of 32 partners
    pdl_wait();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kTcStages; ++s) {
            tc5::mbar_init(&w1full[s], 4);
            tc5::mbar_init(&w2full[s], 4);
            tc5::mbar_init(&yfull[s], 4);
            tc5::mbar_init(&xfull[s], 4);
            tc5::mbar_init(&sempty[s], 1);
        }
        tc5::mbar_init(gfull, 1);
        tc5::mbar_init(gdrained, 8);
        tc5::fence_barrier_init();
    }
    if (warp == 16) tc5::tmem_alloc(tmem_slot, 512);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 8) {
        // ------------------------------------------------------------ weights: thread = row i, which = 0: KL_ij, 1: KL_ji
        const int which = warp >> 2;
        const int t = ((warp & 3) << 5) | lane;
        const int i = ti * kTcT + t;
        const bool row_ok = (i >= row0) && (i < row1);
        const float m = (float)hdr->m;
        const float* krow = (which == 0 ? KL : KLT) + (int64_t)i * Bp;
        uint64_t* wfull = which == 0 ? w1full : w2full;
        const uint32_t tw = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 + which * 128);
        float rs = 0.f;
        const int lrow = i - gbase;
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(which * kTcT);
        float* dst = gpart + (((int64_t)split * 2 + which) * rows_pad + lrow) * Kp + kbase;
        float4 nxt[kTcKC / 4];
        if (total > 0) {
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) nxt[g4] = __ldcs(reinterpret_cast<const float4*>(krow + tj0 * kTcT + 4 * g4));
        }
        for (int qn = 0; qn < total; ++qn) {
            float h[kTcKC], l[kTcKC];
#pragma unroll
            for (int g4 = 0; g4 < kTcKC / 4; ++g4) {
                h[4 * g4] = nxt[g4].x; h[4 * g4 + 1] = nxt[g4].y; h[4 * g4 + 2] = nxt[g4].z; h[4 * g4 + 3] = nxt[g4].w;
            }
            const int j0 = tj0 * kTcT + qn * kTcKC;
            if (qn + 1 < total) {
#pragma unroll
                for (int g4 = 0; g4 < kTcKC / 4; ++g4) nxt[g4] = __ldcs(reinterpret_cast<const float4*>(krow + j0 + kTcKC + 4 * g4));
            }
            const bool plain = row_ok && (j0 + kTcKC <= B) && (i < j0 || i >= j0 + kTcKC);
#pragma unroll
            for (int j = 0; j < kTcKC; ++j) {
                const bool ok = plain || (row_ok && (j0 + j) < B && (j0 + j) != i);
                const float w = ok ? h[j] - m : 0.f;
                rs += w;
                tc5::split_tf32(w, h[j], l[j]);
            }
            const int s = qn % kTcStages;
            tc5::mbar_wait(&sempty[s], ((uint32_t)(qn / kTcStages) & 1u) ^ 1u);
            tc5::fence_after_sync();
            const uint32_t tws = tw + (uint32_t)(s * 2 * kTcKC);
#pragma unroll
            for (int j = 0; j < kTcKC; j += 16) {
                tc5::tmem_st16(tws + j, h + j);
                tc5::tmem_st16(tws + kTcKC + j, l + j);
            }
            tc5::tmem_st_wait();
            tc5::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(&wfull[s]);
            if ((qn + 1) % group_chunks == 0 || qn + 1 == total) {
                // end of an accumulation group: this warpgroup's accumulator (G1 or G2) -> fp32 partial sums in global memory
                const int g = qn / group_chunks;
                tc5::mbar_wait(gfull, (uint32_t)g & 1u);
                tc5::fence_after_sync();
#pragma unroll 1
                for (int c0 = 0; c0 < kTcT; c0 += 16) {
                    float v[16];
                    tc5::tmem_ld16(tacc + c0, v);
                    tc5::tmem_ld_wait();
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4) {
                        float4* p = reinterpret_cast<float4*>(dst + c0 + 4 * g4);
                        float4 o = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
                        if (g > 0) {
                            const float4 b = *p;
                            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
                        }
                        *p = o;
                    }
                }
                tc5::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(gdrained);
            }
        }
        if (blockIdx.y == 0) rspart[((int64_t)split * 2 + which) * rows_pad + lrow] = rs;
        if (total == 0) {
            for (int c0 = 0; c0 < kTcT; c0 += 4) *reinterpret_cast<float4*>(dst + c0) = make_float4(0.f, 0.f, 0.f, 0.f);
        }
    } else if (warp < 16) {
        // ------------------------------------------------------------ panels: lane = partner j of the chunk, transposed store
        const int which = (warp - 8) >> 2;                  // 0: Y panel (pairs with W1), 1: X panel (pairs with W2)
        const int w4 = warp & 3;
        const float* P = which == 0 ? Yp : Xp;
        uint64_t* pfull = which == 0 ? yfull : xfull;
        float4 nxt[8];
        auto load = [&](int qn, float4* dstv) {
            const float* src = P + (int64_t)(tj0 * kTcT + qn * kTcKC + lane) * Kp + kbase + w4 * 32;
#pragma unroll
            for (int it = 0; it < 8; ++it) dstv[it] = __ldg(reinterpret_cast<const float4*>(src + 4 * it));
        };
        if (total > 0) load(0, nxt);
        for (int qn = 0; qn < total; ++qn) {
            float4 v[8];
#pragma unroll
            for (int it = 0; it < 8; ++it) v[it] = nxt[it];
            if (qn + 1 < total) load(qn + 1, nxt);
            const int s = qn % kTcStages;
            tc5::mbar_wait(&sempty[s], ((uint32_t)(qn / kTcStages) & 1u) ^ 1u);
            tc5::fence_after_sync();
            // element (n = k, K = j): (j / 4) * LBO + n * 16 + (j % 4) * 4
            uint8_t* hi = smem + s * kBwdTcStageB + which * 2 * kBwdTcMat + (lane >> 2) * kBwdTcLbo + (lane & 3) * 4;
            uint8_t* lo = hi + kBwdTcMat;
#pragma unroll
            for (int it = 0; it < 8; ++it) {
                const int n = w4 * 32 + 4 * it;
                float h4[4], l4[4];
                tc5::split_tf32(v[it].x, h4[0], l4[0]);
                tc5::split_tf32(v[it].y, h4[1], l4[1]);
                tc5::split_tf32(v[it].z, h4[2], l4[2]);
                tc5::split_tf32(v[it].w, h4[3], l4[3]);
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                    *reinterpret_cast<float*>(hi + (n + e) * 16) = h4[e];
                    *reinterpret_cast<float*>(lo + (n + e) * 16) = l4[e];
                }
            }
            tc5::fence_proxy_async();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(&pfull[s]);
        }
    } else if (lane == 0) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc5::idesc_tf32(kTcT, kTcT);
        const uint32_t st_a = tc5::smem_u32(smem);
        const uint32_t dG1 = tmem_base, dG2 = tmem_base + kTcT;
        for (int qn = 0; qn < total; ++qn) {
            const int s = qn % kTcStages;
            const uint32_t ph = (uint32_t)(qn / kTcStages) & 1u;
            const int gq = qn % group_chunks;
            if (gq == 0 && qn > 0) {       // the previous group's sums have been read out of the accumulators
                tc5::mbar_wait(gdrained, (uint32_t)(qn / group_chunks - 1) & 1u);
                tc5::fence_after_sync();
            }
            tc5::mbar_wait(&w1full[s], ph);
            tc5::mbar_wait(&w2full[s], ph);
            tc5::mbar_wait(&yfull[s], ph);
            tc5::mbar_wait(&xfull[s], ph);
            tc5::fence_after_sync();
            const uint32_t w1h = tmem_base + (uint32_t)(256 + s * 2 * kTcKC), w1l = w1h + kTcKC;
            const uint32_t w2h = tmem_base + (uint32_t)(384 + s * 2 * kTcKC), w2l = w2h + kTcKC;
            const uint32_t yh = st_a + s * kBwdTcStageB, yl = yh + kBwdTcMat, xh = yl + kBwdTcMat, xl = xh + kBwdTcMat;
#pragma unroll
            for (int j = 0; j < kTcKC / 8; ++j) {
                const uint32_t acc = (gq | j) != 0;
                const uint64_t dyh = tc5::smem_desc(yh + 2 * j * kBwdTcLbo, kBwdTcLbo, 128);
                const uint64_t dyl = tc5::smem_desc(yl + 2 * j * kBwdTcLbo, kBwdTcLbo, 128);
                const uint64_t dxh = tc5::smem_desc(xh + 2 * j * kBwdTcLbo, kBwdTcLbo, 128);
                const uint64_t dxl = tc5::smem_desc(xl + 2 * j * kBwdTcLbo, kBwdTcLbo, 128);
                tc5::mma_tf32_ts(dG1, w1l + 8 * j, dyh, idesc, acc);
                tc5::mma_tf32_ts(dG1, w1h + 8 * j, dyl, idesc, 1u);
                tc5::mma_tf32_ts(dG1, w1h + 8 * j, dyh, idesc, 1u);
                tc5::mma_tf32_ts(dG2, w2l + 8 * j, dxh, idesc, acc);
                tc5::mma_tf32_ts(dG2, w2h + 8 * j, dxl, idesc, 1u);
                tc5::mma_tf32_ts(dG2, w2h + 8 * j, dxh, idesc, 1u);
            }
            tc5::mma_commit(&sempty[s]);
            if (gq + 1 == group_chunks || qn + 1 == total) tc5::mma_commit(gfull);
        }
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 16) {
        tc5::fence_after_sync();
        tc5::tmem_dealloc(tmem_base, 512);
    }
}

// own-row gradient from the partner sums: thread = (row, latent dim); the splits are added in a fixed order
__global__ void __launch_bounds__(256)
mpd_bwd_tc_finish_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                         const float4* __restrict__ Vp, const double* __restrict__ colstats, const double* __restrict__ pivots,
                         const float* __restrict__ g_m, const float* __restrict__ g_S, const float* __restrict__ S,
                         const float* __restrict__ gkl, int B, int D, int Kp, int row0, int nrows, int nsplit,
                         const float* __restrict__ gpart, const float* __restrict__ rspart, int rows_pad, int gbase,
                         float* __restrict__ dmu, float* __restrict__ dlv, int accumulate) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= (int64_t)nrows * D) return;
    const int i = row0 + (int)(idx / D), d = (int)(idx % D);
    const int lrow = i - gbase;
    const float beta = weight_scale(g_S, S, B);
    float g1r = 0.f, g1m = 0.f, g2x = 0.f, g2m = 0.f, rs1 = 0.f, rs2 = 0.f;
    if (beta != 0.f) {
        for (int sp = 0; sp < nsplit; ++sp) {
            const float2 a = *reinterpret_cast<const float2*>(gpart + (((int64_t)sp * 2 + 0) * rows_pad + lrow) * Kp + 2 * d);
            const float2 b = *reinterpret_cast<const float2*>(gpart + (((int64_t)sp * 2 + 1) * rows_pad + lrow) * Kp + 2 * d);
            g1r += a.x; g1m += a.y; g2x += b.x; g2m += b.y;
            rs1 += rspart[((int64_t)sp * 2 + 0) * rows_pad + lrow];
            rs2 += rspart[((int64_t)sp * 2 + 1) * rows_pad + lrow];
        }
        g1r *= beta; g1m *= beta; g2x *= beta; g2m *= beta; rs1 *= beta; rs2 *= beta;
    }
    const float4 pi = Vp[(int64_t)i * D + d];   // {mc, a, b, t}
    const float ivb = (float)pivots[3 * (int64_t)D + d];
    const float mc = pi.x, a = pi.y, ri = (1.0f + pi.z) * ivb, t = pi.w;
    float gmu = mc * (g1r + rs1) * ivb + 0.5f * g1m - ri * (g2m - mc * rs2);
    float glv = 0.5f * ((1.0f + a) * g1r + a * rs1) + 0.5f * (-t * rs2 - (1.0f + t) * (g2x - (2.0f * mc * g2m - mc * mc * rs2) * ivb));
    mean_part(colstats, pivots, g_m, B, D, d, pi, mu[(int64_t)i * mu_stride + d], lv[(int64_t)i * lv_stride + d], gmu, glv,
              gkl ? gkl + i : nullptr);
    const int64_t o = (int64_t)(i - row0) * D + d;
    if (accumulate) {
        dmu[o] += gmu;
        dlv[o] += glv;
    } else {
        dmu[o] = gmu;
        dlv[o] = glv;
    }
}
