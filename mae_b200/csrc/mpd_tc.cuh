// Tensor-core tiles of the MPD regulariser for large batches (C5: B >= 2048): the K = 2D contraction of the forward
//   2 KL_ij = x~_i . y~_j + E_i + F_j
// and the partner (K = B) contractions of the stored-KL backward
//   G1_i = sum_j (KL_ij - m) y~_j ,  G2_i = sum_j (KL_ji - m) x~_j
// as 3xTF32 tcgen05.mma (fp32-level accuracy: every operand is split into tf32(v) and v - tf32(v), three products, fp32
// accumulation in TMEM; tc5.cuh).  Included by mpd.cu inside namespace mae::{anonymous} (shares Hdr / Layout).
//
// Both kernels are warp-specialised, one persistent CTA per SM:
//   A side     thread = tile row = TMEM lane: loads its part of the chunk from global (the pre-split tiled panel, or the stored
//              KL rows which it turns into weights and splits), tcgen05.st into a TMEM operand stage -> the tensor core
//              reads only B from smem (with K = 8 per tf32 instruction an SS-mode MMA re-reads 8 KB of operands per 64
//              cycles: the shared-memory bandwidth, measured in dmol_head.cu)
//   B side     one thread: cp.async.bulk of the chunk's 16 KB halves from the tiled panels (written by
//              mpd_tile_panels_kernel already split and already in the K-major canonical shared-memory layout)
//   MMA        one thread: 3 tcgen05.mma (128 x 128 x 8) per k-step, tcgen05.commit -> mbarriers
//   epilogue   thread = row: tcgen05.ld of the accumulator, KL / weights arithmetic, stores
#pragma once

// (tc5.cuh is included by mpd.cu at file scope)

constexpr int kTcT = 128;            // tile edge (rows = TMEM lanes, columns = MMA N)
constexpr int kTcKC = 32;            // contraction elements per pipeline stage
constexpr int kTcLbo = 16 * 128;     // canonical layout: 128 rows x 16 B per k-group
constexpr int kTcHalf = kTcT * kTcKC * 4;         // one of hi / lo of an operand chunk: 16 KB, contiguous in the tiled panels
constexpr int kTcStageB = 2 * kTcHalf;            // B_hi | B_lo of one stage

// Tiled, pre-split copies of the panels (workspace region L.tp, 8 matrices of Bp x Kp floats; written once per forward by
// mpd_tile_panels_kernel, read by both tensor-core kernels).  "Row" tilings 0..3 = X_hi, X_lo, Y_hi, Y_lo hold element
// (row r, k) at  ((r / 128 * (Kp / 4) + k / 4) * 128 + r % 128) * 4 + k % 4 : a chunk of 32 k of a 128-row tile is 16 KB of
// CONTIGUOUS memory that already is the K-major canonical UMMA layout (one cp.async.bulk per operand half, or coalesced
// 16-byte loads with thread = row).  "Partner" tilings 4..7 = Y_hi, Y_lo, X_hi, X_lo hold element (n = k, K = partner j) of
// the backward's B operands at  ((((j / 128 * (Kp / 128) + k / 128) * 32 + j % 128 / 4) * 128 + k % 128) * 4 + j % 4 : a chunk
// of 32 partners x 128 gradient columns is again 16 KB contiguous in canonical layout.
__global__ void __launch_bounds__(256)
mpd_tile_panels_kernel(const float* __restrict__ Xp, const float* __restrict__ Yp, int Bp, int Kp, float* __restrict__ tp,
                       int64_t mat, int with_partner) {
    const int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int kq = Kp / 4;
    if (idx >= (int64_t)Bp * kq) return;
    const int r = (int)(idx % Bp), g = (int)(idx / Bp);      // consecutive threads: consecutive rows -> coalesced stores
    const float4 x = *reinterpret_cast<const float4*>(Xp + (int64_t)r * Kp + 4 * g);
    const float4 y = *reinterpret_cast<const float4*>(Yp + (int64_t)r * Kp + 4 * g);
    float4 xh, xl, yh, yl;
    tc5::split_tf32(x.x, xh.x, xl.x); tc5::split_tf32(x.y, xh.y, xl.y); tc5::split_tf32(x.z, xh.z, xl.z); tc5::split_tf32(x.w, xh.w, xl.w);
    tc5::split_tf32(y.x, yh.x, yl.x); tc5::split_tf32(y.y, yh.y, yl.y); tc5::split_tf32(y.z, yh.z, yl.z); tc5::split_tf32(y.w, yh.w, yl.w);
    const int64_t o = (((int64_t)(r / 128) * kq + g) * 128 + r % 128) * 4;
    *reinterpret_cast<float4*>(tp + 0 * mat + o) = xh;
    *reinterpret_cast<float4*>(tp + 1 * mat + o) = xl;
    *reinterpret_cast<float4*>(tp + 2 * mat + o) = yh;
    *reinterpret_cast<float4*>(tp + 3 * mat + o) = yl;
    if (with_partner) {
        const int nc = (4 * g) / 128, kk = (4 * g) % 128;
        const int64_t ob = ((((int64_t)(r / 128) * (Kp / 128) + nc) * 32 + (r % 128) / 4) * 128 + kk) * 4 + r % 4;
        const float yhv[4] = {yh.x, yh.y, yh.z, yh.w}, ylv[4] = {yl.x, yl.y, yl.z, yl.w};
        const float xhv[4] = {xh.x, xh.y, xh.z, xh.w}, xlv[4] = {xl.x, xl.y, xl.z, xl.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            tp[4 * mat + ob + 4 * e] = yhv[e];
            tp[5 * mat + ob + 4 * e] = ylv[e];
            tp[6 * mat + ob + 4 * e] = xhv[e];
            tp[7 * mat + ob + 4 * e] = xlv[e];
        }
    }
}

__device__ long long g_mpd_prof[64];     // MAE_MPD_PROF=1: per-role cycle counters (development aid)

constexpr int kFwdAStages = 4;       // TMEM operand stages (A = x~ rows): 256 + 4 * 64 = all 512 columns
constexpr int kFwdBStages = 4;       // shared-memory stages (B = y~ rows, filled by the bulk-copy engine); six stages with run-time
                                     // stage indices in the issuer measured slower (0.475 vs 0.443 ms at B = 16384 / D = 64)
constexpr int kStagePitch = 36;      // floats per row of a 32 x 32 staging tile (16-byte aligned, bank-conflict-free)
struct FwdTcSmem {
    static constexpr int kBOff = 0;
    static constexpr int kFOff = kFwdBStages * kTcStageB;          // f_j of the current tile, one copy per epilogue / helper warpgroup
    static constexpr int kBarOff = kFOff + 4 * kTcT * 4;
    static constexpr int kNumBars = 2 * kFwdAStages + 2 * kFwdBStages + 4;
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kRedOff = kSlotOff + 16;
    static constexpr int kStageOff = (kRedOff + 16 * 8 + 16 + 15) / 16 * 16;     // per-warp KL staging tiles (16 warps)
    static constexpr int kBytes = kStageOff + 16 * 32 * kStagePitch * 4;
};
constexpr int kFwdTcThreads = (8 + 8 + 1 + 1) * 32;   // 2 epilogue warpgroups, 2 A warpgroups, MMA warp, loader warp

// TMEM columns: accumulators at 0 and 128, A operand stages (hi | lo, kTcKC columns each) behind them
constexpr int kFwdTcACol = 256;
constexpr int kFwdTcTmemCols = 512;

__global__ void __launch_bounds__(kFwdTcThreads, 1)
mpd_fwd_tc_kernel(const float* __restrict__ tp, int64_t mat, const float* __restrict__ evec,
                  const float* __restrict__ fvec, Hdr* hdr, double* part, int B, int Kp, int row0, int row1, int tr0, int ntr,
                  int ntc, float* __restrict__ KL, float* __restrict__ KLT, int Bp, double* __restrict__ sumsq_out,
                  float* __restrict__ S_out, int tc0, int kl_base, long long* prof) {
    using SM = FwdTcSmem;
    // prof != nullptr (MAE_MPD_PROF=1, profiling runs): CTA 0 reports per role the cycles of its loop and of its barrier waits
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
    uint8_t* bst = smem + SM::kBOff;
    float* fs = reinterpret_cast<float*>(smem + SM::kFOff);
    uint64_t* afull = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
    uint64_t* aempty = afull + kFwdAStages;
    uint64_t* bfull = aempty + kFwdAStages;
    uint64_t* bempty = bfull + kFwdBStages;
    uint64_t* tfull = bempty + kFwdBStages;
    uint64_t* tempty = tfull + 2;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::kSlotOff);
    double* red = reinterpret_cast<double*>(smem + SM::kRedOff);
    __shared__ bool is_last;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int ntiles = ntr * ntc;
    const int nchunk = Kp / kTcKC;
    const int kq = Kp / 4;
    // a CTA takes a CONTIGUOUS range of tiles in row-major order, so consecutive tiles share the row tile.  With K = 2D = 128
    // the four A stages in TMEM hold the whole pre-split row operand: it is stored once per run of equal row tiles and every
    // tile of the run reads it in place ("resident").  Why it matters: L2 delivers ~43 B/clk to an SM (~6300 B/clk chip-wide),
    // a 128 x 128 tile at the MMA floor (48 MMAs x 64 cycles) would need 83 B/clk when both operands are re-read per tile
    // (profiles/r2_mma_rate.txt); with A resident only B streams: 43 B/clk.
    // Otherwise (K > 128) the tiles are dealt round-robin: the CTAs then work on neighbouring tiles of one row at a time and
    // share its A operand in L2 (contiguous ranges measured 30 % slower there).
    const bool resident = (nchunk == kFwdAStages);
    // the idle A warpgroups help with the epilogue only when KL rows are stored (the evaluation-only forward is MMA-bound with
    // four epilogue warps per tile; eight measured 4 % slower there)
    const bool helpers = resident && KL != nullptr;
    const int per_cta = (ntiles + (int)gridDim.x - 1) / (int)gridDim.x;
    const int t_begin = resident ? min(ntiles, (int)blockIdx.x * per_cta) : (int)blockIdx.x;
    const int t_step = resident ? 1 : (int)gridDim.x;
    const int my_tiles = resident ? min(ntiles, t_begin + per_cta) - t_begin
                                  : ((int)blockIdx.x < ntiles ? (ntiles - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0);
    const int total = my_tiles * nchunk;         // chunks of this CTA, in (tile, k) order
    pdl_wait();                // tiled panels, E / F and m come from the kernels before this one
    pdl_launch_dependents();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kFwdAStages; ++s) {
            tc5::mbar_init(&afull[s], 4);
            tc5::mbar_init(&aempty[s], 1);
        }
        for (int s = 0; s < kFwdBStages; ++s) {
            tc5::mbar_init(&bfull[s], 1);
            tc5::mbar_init(&bempty[s], 1);
        }
        for (int a = 0; a < 2; ++a) {
            tc5::mbar_init(&tfull[a], 1);
            tc5::mbar_init(&tempty[a], helpers ? 8 : 4);      // one arrival per epilogue (and helper) warp
        }
        tc5::fence_barrier_init();
    }
    if (warp == 16) tc5::tmem_alloc(tmem_slot, kFwdTcTmemCols);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    // ---------------------------------------------------------------- epilogue of one tile: thread = row of the tile, columns
    // [c_lo, c_hi) of accumulator `acc`.  Run by the epilogue warpgroup that owns the accumulator and, in resident mode, by the
    // otherwise idle A warpgroup of the same parity on the other half of the columns (8 warps per tile instead of 4).
    const float m_piv = (float)hdr->m;
    auto epilogue_tile = [&](int it, int acc, int q, int c_lo, int c_hi, float* fw, int bar_id, double& tot, long long& pwacc) {
        const int t = (q << 5) | lane;
        const int tile = t_begin + it * t_step;
        const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * kTcT);
        const int use = it >> 1;
        const int ti = tr0 + tile / ntc, tj = tc0 + tile % ntc;
        const int i = ti * kTcT + t;
        const float ei = evec[i];
        const float m = m_piv;
        tc5::named_sync(bar_id, 128);                 // the previous tile's readers of fw are done
        fw[t] = fvec[tj * kTcT + t];
        tc5::named_sync(bar_id, 128);
        const bool row_ok = (i >= row0) && (i < row1);
        const bool plain = (ti != tj) && (tj * kTcT + kTcT <= B);      // no diagonal, no padded columns
        wait(&tfull[acc], use & 1, pwacc);
        tc5::fence_after_sync();
        float psum = 0.f;
        float* klrow = KL != nullptr ? KL + (int64_t)(i - kl_base) * Bp + tj * kTcT : nullptr;
        float* kltcol = KLT != nullptr ? KLT + (int64_t)(tj * kTcT - kl_base) * Bp + ti * kTcT + t : nullptr;
        float* stage = reinterpret_cast<float*>(smem + SM::kStageOff) + warp * (32 * kStagePitch);
        float v[16], w[16];
        tc5::tmem_ld16(tacc + c_lo, v);
        tc5::tmem_ld_wait();
#pragma unroll 1
        for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
            if (c0 + 16 < c_hi) tc5::tmem_ld16(tacc + c0 + 16, w);     // in flight while this batch is evaluated
            if (plain && row_ok) {       // interior tile: no masks; f_j by 16-byte broadcast loads
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4) {
                    const float4 f4 = *reinterpret_cast<const float4*>(fw + c0 + 4 * g4);
                    const float ff[4] = {f4.x, f4.y, f4.z, f4.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        const float kl = 0.5f * (v[4 * g4 + e] + (ei + ff[e]));
                        v[4 * g4 + e] = kl;
                        const float dev = kl - m;
                        psum = fmaf(dev, dev, psum);
                    }
                }
            } else {
#pragma unroll
                for (int c = 0; c < 16; ++c) {
                    const float kl = 0.5f * (v[c] + (ei + fw[c0 + c]));
                    v[c] = kl;
                    const float dev = kl - m;
                    const int j = tj * kTcT + c0 + c;
                    const bool ok = row_ok && j < B && j != i;
                    psum = ok ? fmaf(dev, dev, psum) : psum;
                }
            }
            if (klrow != nullptr) {
                // KL rows leave through a per-warp staging tile (32 rows x 32 columns, pitch 36 floats: conflict-free for the
                // 16-byte row writes and for the 16-byte reads below), so that one store instruction writes four complete
                // 128-byte lines.  Thread = row storing 16 bytes of its own row touched 32 lines per instruction: 4096 L1
                // wavefronts per tile, on the data path the MMA reads its B operand through (issuer blocked at 131 cycles / MMA).
                float* strow = stage + lane * kStagePitch + (c0 & 16);
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4)
                    *reinterpret_cast<float4*>(strow + 4 * g4) = make_float4(v[4 * g4], v[4 * g4 + 1], v[4 * g4 + 2], v[4 * g4 + 3]);
                if (c0 & 16) {
                    __syncwarp();
                    const int cb = c0 - 16;                    // first column of the staged 32
                    float* gdst = KL + (int64_t)(ti * kTcT + q * 32 - kl_base) * Bp + tj * kTcT + cb + (lane & 7) * 4;
#pragma unroll
                    for (int r4 = 0; r4 < 8; ++r4) {
                        const int r = r4 * 4 + (lane >> 3);
                        const float4 x4 = *reinterpret_cast<const float4*>(stage + r * kStagePitch + (lane & 7) * 4);
                        *reinterpret_cast<float4*>(gdst + (int64_t)r * Bp) = x4;
                    }
                    __syncwarp();
                }
            }
            if (kltcol != nullptr) {
#pragma unroll
                for (int c = 0; c < 16; ++c) kltcol[(int64_t)(c0 + c) * Bp] = v[c];      // lanes = consecutive i: coalesced
            }
            tc5::tmem_ld_wait();
#pragma unroll
            for (int c = 0; c < 16; ++c) v[c] = w[c];
        }
        tc5::fence_before_sync();
        __syncwarp();
        if (lane == 0) tc5::mbar_arrive(&tempty[acc]);
        tot += (double)psum;
    };
    // CTA total of the squared deviations over the warps that ran epilogues, fixed order; the last CTA adds the CTA totals
    const int nred = helpers ? 16 : 8;
    auto reduce_total = [&](double tot) {
        tot = warp_sum(tot);
        if (lane == 0) red[warp] = tot;
        tc5::named_sync(3, nred * 32);
        if (sumsq_out != nullptr && threadIdx.x == 0) {
            double s = 0.0;
            for (int w8 = 0; w8 < nred; ++w8) s += red[w8];
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
        tc5::named_sync(3, nred * 32);
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
    };

    if (warp < 8) {
        // ------------------------------------------------------------ epilogue warpgroups: accumulator wg, tiles of parity wg
        const int wg = warp >> 2, q = warp & 3;
        const int c_hi = helpers ? kTcT / 2 : kTcT;
        double tot = 0.0;
        for (int it = wg; it < my_tiles; it += 2) epilogue_tile(it, wg, q, 0, c_hi, fs + wg * kTcT, 1 + wg, tot, pw0);
        if (((q << 5) | lane) == 0) report(4 * wg);
        reduce_total(tot);
    } else if (warp < 16) {
        // ------------------------------------------------------------ A: pre-split x~ rows -> TMEM operand stage; the two
        // warpgroups take alternate chunks, so one has its 16 loads in flight while the other stores
        const int pw = (warp - 8) >> 2;
        const int t = ((warp & 3) << 5) | lane;
        const uint32_t ta = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)kFwdTcACol;
        const float* xh = tp, * xl = tp + mat;
        auto produce = [&](int ti, int kc, int s, uint32_t empty_parity) {
            const int64_t o = (((int64_t)ti * kq + kc * 8) * 128 + t) * 4;
            float h[kTcKC], l[kTcKC];
#pragma unroll
            for (int g = 0; g < 8; ++g) {
                const float4 a = __ldg(reinterpret_cast<const float4*>(xh + o + g * 512));
                const float4 b = __ldg(reinterpret_cast<const float4*>(xl + o + g * 512));
                h[4 * g] = a.x; h[4 * g + 1] = a.y; h[4 * g + 2] = a.z; h[4 * g + 3] = a.w;
                l[4 * g] = b.x; l[4 * g + 1] = b.y; l[4 * g + 2] = b.z; l[4 * g + 3] = b.w;
            }
            wait(&aempty[s], empty_parity, pw0);
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
        };
        if (resident) {
            // per tile, in order: refill the four stages when a new row tile starts (once the previous run's MMAs are done),
            // then help the epilogue warpgroup of this parity with the upper half of the columns
            int run = 0, prev_ti = -1;
            double tot = 0.0;
            for (int it = 0; it < my_tiles; ++it) {
                const int ti = tr0 + (t_begin + it) / ntc;
                if (ti != prev_ti) {
                    prev_ti = ti;
                    for (int kc = pw; kc < nchunk; kc += 2) produce(ti, kc, kc, ((uint32_t)run & 1u) ^ 1u);
                    ++run;
                }
                if (helpers && (it & 1) == pw) epilogue_tile(it, pw, warp & 3, kTcT / 2, kTcT, fs + (2 + pw) * kTcT, 4 + pw, tot, pw1);
            }
            if (t == 0) report(8 + 4 * pw);
            if (helpers) reduce_total(tot);
        } else {
            for (int qn = pw; qn < total; qn += 2) {
                const int tile = t_begin + (qn / nchunk) * t_step;
                produce(tr0 + tile / ntc, qn % nchunk, qn % kFwdAStages, ((uint32_t)(qn / kFwdAStages) & 1u) ^ 1u);
            }
            if (t == 0) report(8 + 4 * pw);
        }
    } else if (warp == 17) {
        // ------------------------------------------------------------ loader: B = y~ chunk, two contiguous 16 KB bulk copies
        if (tc5::elect_one()) {
            const float* yh = tp + 2 * mat, * yl = tp + 3 * mat;
            for (int qn = 0; qn < total; ++qn) {
                const int tile = t_begin + (qn / nchunk) * t_step;
                const int tj = tc0 + tile % ntc, kc = qn % nchunk;
                const int64_t o = (((int64_t)tj * kq + kc * 8) * 128) * 4;
                const int s = qn % kFwdBStages;
                wait(&bempty[s], ((uint32_t)(qn / kFwdBStages) & 1u) ^ 1u, pw0);
                tc5::mbar_expect_tx(&bfull[s], kTcStageB);
                tc5::bulk_g2s(bst + s * kTcStageB, yh + o, kTcHalf, &bfull[s]);
                tc5::bulk_g2s(bst + s * kTcStageB + kTcHalf, yl + o, kTcHalf, &bfull[s]);
            }
            report(16);
        }
    } else if (tc5::elect_one()) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc5::idesc_tf32(kTcT, kTcT);
        const uint32_t b_a = tc5::smem_u32(bst);
        int qn = 0, run = -1, prev_ti = -1;
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = t_begin + it * t_step;
            const int a = it & 1, use = it >> 1;
            const int ti = tr0 + tile / ntc;
            const bool new_row = resident && ti != prev_ti;                    // wait for the row operand once per run ...
            const bool run_ends = resident && (it + 1 == my_tiles || tr0 + (tile + 1) / ntc != ti);    // ... release it after the last tile
            if (new_row) { ++run; prev_ti = ti; }
            wait(&tempty[a], ((uint32_t)use & 1u) ^ 1u, pw0);
            tc5::fence_after_sync();
            const uint32_t d = tmem_base + (uint32_t)(a * kTcT);
            // one chunk: wait for both operands, 4 k-steps x 3 products, release the stages.  With K = 2D a multiple of 128
            // the chunk count per tile is a multiple of the stage count, so the stage index is the unroll index and every
            // descriptor is loop-invariant (the issuing thread's uniform-datapath arithmetic was on the critical path:
            // 75 % busy at 2.2x the tensor pipe's 64 cycles per MMA)
            auto chunk = [&](int kc, int sa, int sb) {
                if (!resident) wait(&afull[sa], (uint32_t)(qn / kFwdAStages) & 1u, pw1);
                else if (new_row) wait(&afull[sa], (uint32_t)run & 1u, pw1);
                { long long tmpw = 0; wait(&bfull[sb], (uint32_t)(qn / kFwdBStages) & 1u, tmpw); pw1 += tmpw; if (prof) prof[31] += tmpw; }
                tc5::fence_after_sync();
                const uint32_t a_hi = tmem_base + (uint32_t)(kFwdTcACol + sa * 2 * kTcKC), a_lo = a_hi + kTcKC;
                const uint32_t b_hi = b_a + sb * kTcStageB, b_lo = b_hi + kTcHalf;
#pragma unroll
                for (int j = 0; j < kTcKC / 8; ++j) {
                    const uint64_t bh = tc5::smem_desc(b_hi + 2 * j * kTcLbo, kTcLbo, 128);
                    const uint64_t bl = tc5::smem_desc(b_lo + 2 * j * kTcLbo, kTcLbo, 128);
                    tc5::mma_tf32_ts(d, a_lo + 8 * j, bh, idesc, (kc | j) != 0);
                    tc5::mma_tf32_ts(d, a_hi + 8 * j, bl, idesc, 1u);
                    tc5::mma_tf32_ts(d, a_hi + 8 * j, bh, idesc, 1u);
                }
                if (!resident || run_ends) tc5::mma_commit(&aempty[sa]);
                tc5::mma_commit(&bempty[sb]);
                ++qn;
            };
            static_assert(kFwdAStages == 4 && kFwdBStages == 4, "the unrolled issue loop assumes four stages of each operand");
            if ((nchunk & 3) == 0) {
                for (int kc = 0; kc < nchunk; kc += 4) {
                    chunk(kc, 0, 0);
                    chunk(kc + 1, 1, 1);
                    chunk(kc + 2, 2, 2);
                    chunk(kc + 3, 3, 3);
                }
            } else {
                for (int kc = 0; kc < nchunk; ++kc) chunk(kc, qn % kFwdAStages, qn % kFwdBStages);
            }
            tc5::mma_commit(&tfull[a]);
        }
        report(20);
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
constexpr int kBwdWStages = 2;                     // TMEM weight stages per W (hi | lo, 64 columns each)
constexpr int kBwdPStages = 2;                     // shared-memory panel stages: Y_hi | Y_lo | X_hi | X_lo, 16 KB each
constexpr int kBwdTcStageB = 4 * kTcHalf;
struct BwdTcSmem {
    static constexpr int kBarOff = kBwdPStages * kBwdTcStageB;
    static constexpr int kNumBars = 3 * kBwdWStages + 2 * kBwdPStages + 2;   // w1full, w2full, wempty; pfull, pempty; gfull, gdrained
    static constexpr int kSlotOff = kBarOff + kNumBars * 8;
    static constexpr int kBytes = kSlotOff + 32;
};
constexpr int kBwdTcThreads = (16 + 1 + 1) * 32;   // two pairs of (W1 rows, W2 rows) warpgroups, MMA warp, loader warp

__global__ void __launch_bounds__(kBwdTcThreads, 1)
mpd_bwd_tc_kernel(const float* __restrict__ tp, int64_t mat, const Hdr* __restrict__ hdr, int B, int Kp,
                  int row0, int row1, int tr0, int ntc, int tiles_per_split, const float* __restrict__ KL,
                  const float* __restrict__ KLT, int Bp, float* __restrict__ gpart, float* __restrict__ rspart, int rows_pad,
                  int gbase, int group_chunks, long long* prof) {
    long long pw0 = 0, pw1 = 0, pw2 = 0;
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
        if (prof != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            prof[slot] = clock64() - t_start;
            prof[slot + 1] = pw0;
            prof[slot + 2] = pw1;
            prof[slot + 3] = pw2;
        }
    };
    // gpart [split][2][Kp][rows_pad] (column-major per matrix), rspart [split][2][2 pairs][rows_pad].
    // KL / KLT: pre-offset so that global row i is at [i * Bp] (row blocks keep only their rows).  KLT == nullptr (full
    // batch): KL_ji of row i is read from COLUMN i of KL - for a fixed partner j the 32 lanes of a warp (= 32 consecutive
    // rows i) read one 128-byte line, so the transposed matrix is neither written by the forward nor needed here.  gpart
    // [split][2][rows_pad][Kp], rspart [split][2][rows_pad], local row = i - gbase.
    // group_chunks: chunks of 32 partners per TMEM accumulation.  The tensor core's accumulator does not round to nearest, so
    // the error of a running sum grows with its length; after every group the owners of the rows add the accumulator to the
    // CTA's fp32 partial sums in global memory (round to nearest) and the next group starts from zero.
    using SM = BwdTcSmem;
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* w1full = reinterpret_cast<uint64_t*>(smem + SM::kBarOff);
    uint64_t* w2full = w1full + kBwdWStages;
    uint64_t* wempty = w2full + kBwdWStages;
    uint64_t* pfull = wempty + kBwdWStages;
    uint64_t* pempty = pfull + kBwdPStages;
    uint64_t* gfull = pempty + kBwdPStages;
    uint64_t* gdrained = gfull + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM::kSlotOff);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // grid = (column chunks, row tiles, splits): the CTAs that share the KL rows of one row tile (one per 128 gradient columns)
    // are launched next to each other and run in step, so the rows come from HBM once and from L2 for the other chunks
    const int ti = tr0 + blockIdx.y;
    const int cchunk = blockIdx.x;
    const int kbase = cchunk * kTcT;
    const int split = blockIdx.z;
    const int tj0 = split * tiles_per_split;
    const int tj1 = min(ntc, tj0 + tiles_per_split);
    const int total = (tj1 - tj0) * (kTcT / kTcKC);      // chunks of 32 partners
    pdl_wait();

    if (threadIdx.x == 0) {
        for (int s = 0; s < kBwdWStages; ++s) {
            tc5::mbar_init(&w1full[s], 4);
            tc5::mbar_init(&w2full[s], 4);
            tc5::mbar_init(&wempty[s], 1);
        }
        for (int s = 0; s < kBwdPStages; ++s) {
            tc5::mbar_init(&pfull[s], 1);
            tc5::mbar_init(&pempty[s], 1);
        }
        tc5::mbar_init(gfull, 1);
        tc5::mbar_init(gdrained, 16);
        tc5::fence_barrier_init();
    }
    if (warp == 16) tc5::tmem_alloc(tmem_slot, 512);
    tc5::fence_before_sync();
    __syncthreads();
    tc5::fence_after_sync();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < 16) {
        // ------------------------------------------------------------ weights: thread = row i; which = 0: KL_ij, 1: KL_ji;
        // two PAIRS of (W1, W2) warpgroups take alternate chunks (pair p owns TMEM weight stage p): while one pair waits for
        // its 32 loads from HBM the other forms, splits and stores its weights - latency is hidden by the second pair instead
        // of a register prefetch (a single pair was busy 95 % and the MMA issuer waited 29-50 % for it)
        const int pair = warp >> 3, which = (warp >> 2) & 1;
        const int t = ((warp & 3) << 5) | lane;
        const int i = ti * kTcT + t;
        const bool row_ok = (i >= row0) && (i < row1);
        const float m = (float)hdr->m;
        const bool by_col = which == 1 && KLT == nullptr;
        const float* krow = by_col ? KL + i : (which == 0 ? KL : KLT) + (int64_t)i * Bp;
        uint64_t* wfull = which == 0 ? w1full : w2full;
        const uint32_t tws = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(256 + which * 128 + pair * 2 * kTcKC);
        float rs = 0.f;
        const int lrow = i - gbase;
        // this pair drains columns [64 pair, 64 pair + 64) of its accumulator
        const uint32_t tacc = tmem_base + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)(which * kTcT + pair * 64);
        float* dst = gpart + (((int64_t)split * 2 + which) * Kp + kbase + pair * 64) * rows_pad + lrow;   // [k][row], see the drain
        const int ngroups = (total + group_chunks - 1) / group_chunks;
        int use = 0;       // chunks this pair has produced (parity of its stage's empty barrier)
        long long ph_load = 0, ph_store = 0, ph_split2 = 0, ph_stwait = 0, tp1 = 0;      // profiling runs: where a weight warp's time goes
        for (int g = 0; g < ngroups; ++g) {
            const int q_end = min(total, (g + 1) * group_chunks);
            int q0 = g * group_chunks;
            q0 += ((q0 & 1) != pair) ? 1 : 0;
            for (int qn = q0; qn < q_end; qn += 2, ++use) {
                const int j0 = tj0 * kTcT + qn * kTcKC;
                const long long tp0 = prof != nullptr ? clock64() : 0;
                float v[kTcKC];
                if (by_col) {        // KL_ji from column i of KL: for a fixed j the lanes (consecutive i) read one 128-byte line
#pragma unroll
                    for (int j = 0; j < kTcKC; ++j) v[j] = __ldg(krow + (int64_t)(j0 + j) * Bp);
                } else {
#pragma unroll
                    for (int g4 = 0; g4 < kTcKC / 4; ++g4) {
                        const float4 q4 = __ldg(reinterpret_cast<const float4*>(krow + j0 + 4 * g4));
                        v[4 * g4] = q4.x; v[4 * g4 + 1] = q4.y; v[4 * g4 + 2] = q4.z; v[4 * g4 + 3] = q4.w;
                    }
                }
                const bool plain = row_ok && (j0 + kTcKC <= B) && (i < j0 || i >= j0 + kTcKC);
                // the first half of the chunk is split BEFORE the wait for the free stage: less work between the MMA's release
                // of the stage and this pair's 'full' arrival (that interval is on the issuer's critical path)
                // interior chunks (no diagonal, no padded partners or rows: warp-uniform test) take a branch-free loop of 7
                // instructions per element; per-element mask branches left ~20 instructions per element with no overlap
                // between elements (1 450 cycles for 16 elements, the critical resource of this kernel)
                const bool all_plain = __all_sync(0xffffffffu, plain);
#pragma unroll
                for (int h0 = 0; h0 < kTcKC; h0 += 16) {
                    float h[16], l[16];
                    if (all_plain) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const float w = v[h0 + j] - m;
                            rs += w;
                            tc5::split_tf32_fast(w, h[j], l[j]);
                        }
                    } else {
#pragma unroll
                        for (int j = 0; j < 16; ++j) {
                            const bool ok = row_ok && (j0 + h0 + j) < B && (j0 + h0 + j) != i;
                            const float w = ok ? v[h0 + j] - m : 0.f;
                            rs += w;
                            tc5::split_tf32_fast(w, h[j], l[j]);
                        }
                    }
                    if (h0 == 0) {
                        if (prof != nullptr) ph_load += clock64() - tp0;      // loads landed + first half split
                        wait(&wempty[pair], ((uint32_t)use & 1u) ^ 1u, pw0);
                        tc5::fence_after_sync();
                        if (prof != nullptr) tp1 = clock64();
                    }
                    tc5::tmem_st16(tws + h0, h);
                    tc5::tmem_st16(tws + kTcKC + h0, l);
                }
                const long long tp2 = prof != nullptr ? clock64() : 0;
                tc5::tmem_st_wait();
                const long long tp3 = prof != nullptr ? clock64() : 0;
                tc5::fence_before_sync();
                __syncwarp();
                if (lane == 0) tc5::mbar_arrive(&wfull[pair]);
                if (prof != nullptr) { ph_store += clock64() - tp1; ph_split2 += tp2 - tp1; ph_stwait += tp3 - tp2; }             // second half split + TMEM stores + arrive
            }
            // end of an accumulation group: this warpgroup's half of its accumulator (G1 or G2) -> fp32 partial sums.
            // The drain keeps these warps from loading ahead, so the KL lines of this pair's first chunk of the next group are
            // pulled into L2 now: after the drain its loads cost an L2 hit instead of an HBM round trip while the issuer waits.
            {
                int qp = (g + 1) * group_chunks;
                qp += ((qp & 1) != pair) ? 1 : 0;
                if (qp < total) {
                    const int jp = tj0 * kTcT + qp * kTcKC;
                    const float* pf = by_col ? krow + (int64_t)(jp + lane) * Bp - lane      // line of partner jp + lane, this warp's rows
                                             : krow + jp;                                   // this row's 128 bytes
                    asm volatile("prefetch.global.L2 [%0];" ::"l"(pf));
                }
            }
            wait(gfull, (uint32_t)g & 1u, pw1);
            tc5::fence_after_sync();
            const long long td0 = prof != nullptr ? clock64() : 0;
            // gpart is COLUMN-major per (split, which): [k][row].  The owner of row i holds one accumulator row (TMEM lane), so
            // for a fixed column the 32 lanes of a warp touch 32 consecutive floats = one 128-byte line: every load and store
            // of the drain is fully coalesced.  (Row-major partial sums - 16-byte accesses 512 bytes apart, 32 half-used
            // sectors per request - made the drains 0.28 ms of a 1.25 ms kernel; serialised loads were not the cause.)
            if (g > 0) {
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float old[16];
#pragma unroll
                    for (int c = 0; c < 16; ++c) old[c] = __ldcg(dst + (int64_t)(c0 + c) * rows_pad);
                    float a16[16];
                    tc5::tmem_ld16(tacc + c0, a16);
                    tc5::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) dst[(int64_t)(c0 + c) * rows_pad] = a16[c] + old[c];
                }
            } else {
#pragma unroll 1
                for (int c0 = 0; c0 < 64; c0 += 16) {
                    float a16[16];
                    tc5::tmem_ld16(tacc + c0, a16);
                    tc5::tmem_ld_wait();
#pragma unroll
                    for (int c = 0; c < 16; ++c) dst[(int64_t)(c0 + c) * rows_pad] = a16[c];
                }
            }
            tc5::fence_before_sync();
            __syncwarp();
            if (lane == 0) tc5::mbar_arrive(gdrained);
            if (prof != nullptr) pw2 += clock64() - td0;
        }
        if (t == 0 && pair == 0) report(4 * which);
        if (prof != nullptr && t == 0 && pair == 0 && blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0) {
            prof[16 + 2 * which] = ph_load;
            prof[17 + 2 * which] = ph_store;
            prof[20 + 2 * which] = ph_split2;
            prof[21 + 2 * which] = ph_stwait;
        }
        if (cchunk == 0) rspart[(((int64_t)split * 2 + which) * 2 + pair) * rows_pad + lrow] = rs;
        if (total == 0) {
            for (int c = 0; c < 64; ++c) dst[(int64_t)c * rows_pad] = 0.f;
        }
    } else if (warp == 17) {
        // ------------------------------------------------------------ loader: the four 16 KB operand halves of a chunk
        if (tc5::elect_one()) {
            const int ncn = Kp / 128;
            for (int qn = 0; qn < total; ++qn) {
                const int tj = tj0 + qn / (kTcT / kTcKC), jc = qn % (kTcT / kTcKC);
                const int64_t o = ((((int64_t)tj * ncn + cchunk) * 32 + jc * 8) * 128) * 4;
                const int s = qn % kBwdPStages;
                wait(&pempty[s], ((uint32_t)(qn / kBwdPStages) & 1u) ^ 1u, pw0);
                tc5::mbar_expect_tx(&pfull[s], kBwdTcStageB);
                uint8_t* st = smem + s * kBwdTcStageB;
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) tc5::bulk_g2s(st + q4 * kTcHalf, tp + (4 + q4) * mat + o, kTcHalf, &pfull[s]);
            }
            report(8);
        }
    } else if (warp == 16 && tc5::elect_one()) {
        // ------------------------------------------------------------ MMA issuer
        constexpr uint32_t idesc = tc5::idesc_tf32(kTcT, kTcT);
        const uint32_t st_a = tc5::smem_u32(smem);
        const uint32_t dG1 = tmem_base, dG2 = tmem_base + kTcT;
        for (int qn = 0; qn < total; ++qn) {
            const int sw = qn % kBwdWStages, sp = qn % kBwdPStages;
            const int gq = qn % group_chunks;
            if (gq == 0 && qn > 0) {       // the previous group's sums have been read out of the accumulators
                wait(gdrained, (uint32_t)(qn / group_chunks - 1) & 1u, pw2);
                tc5::fence_after_sync();
            }
            wait(&w1full[sw], (uint32_t)(qn / kBwdWStages) & 1u, pw0);
            wait(&w2full[sw], (uint32_t)(qn / kBwdWStages) & 1u, pw0);
            wait(&pfull[sp], (uint32_t)(qn / kBwdPStages) & 1u, pw1);
            tc5::fence_after_sync();
            const uint32_t w1h = tmem_base + (uint32_t)(256 + sw * 2 * kTcKC), w1l = w1h + kTcKC;
            const uint32_t w2h = tmem_base + (uint32_t)(384 + sw * 2 * kTcKC), w2l = w2h + kTcKC;
            const uint32_t yh = st_a + sp * kBwdTcStageB, yl = yh + kTcHalf, xh = yl + kTcHalf, xl = xh + kTcHalf;
#pragma unroll
            for (int j = 0; j < kTcKC / 8; ++j) {
                const uint32_t acc = (gq | j) != 0;
                const uint64_t dyh = tc5::smem_desc(yh + 2 * j * kTcLbo, kTcLbo, 128);
                const uint64_t dyl = tc5::smem_desc(yl + 2 * j * kTcLbo, kTcLbo, 128);
                const uint64_t dxh = tc5::smem_desc(xh + 2 * j * kTcLbo, kTcLbo, 128);
                const uint64_t dxl = tc5::smem_desc(xl + 2 * j * kTcLbo, kTcLbo, 128);
                tc5::mma_tf32_ts(dG1, w1l + 8 * j, dyh, idesc, acc);
                tc5::mma_tf32_ts(dG1, w1h + 8 * j, dyl, idesc, 1u);
                tc5::mma_tf32_ts(dG1, w1h + 8 * j, dyh, idesc, 1u);
                tc5::mma_tf32_ts(dG2, w2l + 8 * j, dxh, idesc, acc);
                tc5::mma_tf32_ts(dG2, w2h + 8 * j, dxl, idesc, 1u);
                tc5::mma_tf32_ts(dG2, w2h + 8 * j, dxh, idesc, 1u);
            }
            tc5::mma_commit(&wempty[sw]);
            tc5::mma_commit(&pempty[sp]);
            if (gq + 1 == group_chunks || qn + 1 == total) tc5::mma_commit(gfull);
        }
        report(12);
    }
    tc5::fence_before_sync();
    __syncthreads();
    if (warp == 16) {
        tc5::fence_after_sync();
        tc5::tmem_dealloc(tmem_base, 512);
    }
}

// own-row gradient from the partner sums; the splits are added in a fixed order.  Block = 32 rows x 8 latent dims.  The partial
// sums are column-major ([k][row]), everything else ((mc, a, b, t), mu, logvar, the outputs) is row-major ([row][dim]): phase 1
// reads the partial sums with the ROWS along the lanes, a shared-memory transpose hands the six sums of an element to the
// thread that has the DIMS along the lanes, and phase 2 reads / writes whole sectors.  (One mapping for both phases left
// 4-byte accesses 4 D bytes apart on one side: 32 sectors per request, 0.062 ms for 0.02 ms worth of bytes at B = 16384.)
constexpr int kFinRows = 32, kFinDims = 8, kFinPitch = 36;
__global__ void __launch_bounds__(kFinRows * kFinDims)
mpd_bwd_tc_finish_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                         const float4* __restrict__ Vp, const double* __restrict__ colstats, const double* __restrict__ pivots,
                         const float* __restrict__ g_m, const float* __restrict__ g_S, const float* __restrict__ S,
                         const float* __restrict__ gkl, int B, int D, int Kp, int row0, int nrows, int nsplit,
                         const float* __restrict__ gpart, const float* __restrict__ rspart, int rows_pad, int gbase,
                         float* __restrict__ dmu, float* __restrict__ dlv, int accumulate) {
    __shared__ float sm[6][kFinDims][kFinPitch];
    const int r_blk = blockIdx.x * kFinRows, d_blk = blockIdx.y * kFinDims;
    const float beta = weight_scale(g_S, S, B);
    {   // phase 1: lanes = rows
        const int r = threadIdx.x % kFinRows, dd = threadIdx.x / kFinRows;
        const int li = r_blk + r, d = d_blk + dd;
        float g1r = 0.f, g1m = 0.f, g2x = 0.f, g2m = 0.f, rs1 = 0.f, rs2 = 0.f;
        if (beta != 0.f && li < nrows && d < D) {
            const int lrow = row0 + li - gbase;
            for (int sp = 0; sp < nsplit; ++sp) {
                const float* p1 = gpart + (((int64_t)sp * 2 + 0) * Kp + 2 * d) * rows_pad + lrow;
                const float* p2 = gpart + (((int64_t)sp * 2 + 1) * Kp + 2 * d) * rows_pad + lrow;
                g1r += p1[0]; g1m += p1[rows_pad]; g2x += p2[0]; g2m += p2[rows_pad];
                rs1 += rspart[(((int64_t)sp * 2 + 0) * 2 + 0) * rows_pad + lrow] + rspart[(((int64_t)sp * 2 + 0) * 2 + 1) * rows_pad + lrow];
                rs2 += rspart[(((int64_t)sp * 2 + 1) * 2 + 0) * rows_pad + lrow] + rspart[(((int64_t)sp * 2 + 1) * 2 + 1) * rows_pad + lrow];
            }
            g1r *= beta; g1m *= beta; g2x *= beta; g2m *= beta; rs1 *= beta; rs2 *= beta;
        }
        sm[0][dd][r] = g1r; sm[1][dd][r] = g1m; sm[2][dd][r] = g2x; sm[3][dd][r] = g2m; sm[4][dd][r] = rs1; sm[5][dd][r] = rs2;
    }
    __syncthreads();
    // phase 2: lanes = dims
    const int dd = threadIdx.x % kFinDims, r = threadIdx.x / kFinDims;
    const int li = r_blk + r, d = d_blk + dd;
    if (li >= nrows || d >= D) return;
    const int i = row0 + li;
    const float g1r = sm[0][dd][r], g1m = sm[1][dd][r], g2x = sm[2][dd][r], g2m = sm[3][dd][r], rs1 = sm[4][dd][r], rs2 = sm[5][dd][r];
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
