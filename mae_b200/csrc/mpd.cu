// A3: mutual posterior divergence regulariser (gaussian_encoder.py:162-188) without the [B,B,D] tensor.
//
// Pivoted GEMM decomposition (K = 2D contraction, FP32 FFMA, register-tiled).  With per-dimension pivots
// mu_bar, l_bar (mean of the first <=32 rows; KL is invariant to a common shift of mu and a common scale of v),
// v_bar = exp(l_bar), alpha = lv - l_bar, a = expm1(alpha) = v/v_bar - 1, b = v_bar*r - 1, r = 1/(v + 1e-12),
// mc = mu - mu_bar, s = mc^2/v_bar:
//   2*KL_ij = x~_i . y~_j + E_i + F_j
//     x~_i = [a_i + s_i, mc_i]        (interleaved per latent dim: k = 2d, 2d+1)
//     y~_j = [b_j, -2 mc_j r_j]
//     E_i  = sum_d (a_i - alpha_i + s_i),   F_j = sum_d (b_j + alpha_j + mc_j^2 r_j)
//   Every term is small when the posteriors are alike, so the tiles keep full relative precision in the
//   posterior-collapse regime where the reference's own (A + B - C - 1) cancels (hard part 3).
//   The O(B*D) pre-pass evaluates the per-element quantities in fp64.
//
// Kernels
//   mpd_prepare_kernel   O(B*D): pivots, panels X~, Y~ (row-major [Bp][Kp]), V (float4 {mc,a,b,t}), E, F,
//                        per-dimension column sums in fp64, m_d and m = sum_d m_d (closed form, SURVEY.md §8a "A3-alg").
//   mpd_fwd_kernel       64x64 tiles of KL over rows [row0,row1) x all columns; accumulates sum (KL_ij - m)^2
//                        (m is known beforehand, so no E[x^2]-E[x]^2 cancellation); optional KL / KL^T store.
//   mpd_bwd_small_kernel B <= 2048: per-(i,d) gradient from the stored KL and KL^T (L2-resident).
//   mpd_bwd_tile_kernel  large B: flash-attention-style recompute; per row block, for every column block
//                        recompute KL_ij and KL_ji tiles, form the weights W, and contract W with the
//                        column panels (two more FFMA GEMMs); the row owner receives its complete gradient
//                        (as the i and as the j argument), so there are no atomics and no reduce-scatter.
// All cross-CTA reductions are ticketed and summed in a fixed order -> bit-reproducible results.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"
#include "mpd_math.cuh"
#include "tc5.cuh"

namespace mae {
namespace {

constexpr int kT = 64;         // tile edge (rows and columns per CTA tile)
constexpr int kKC = 16;        // K chunk staged in shared memory per step
constexpr int kNT = 256;       // threads per CTA: 16 x 16 threads, 4 x 4 micro-tile each
constexpr int kLd = kT + 4;    // padded leading dimension (keeps float4 alignment)
// K columns of gradient owned by one CTA of the tile backward: template parameter OC = 128 (64 latent dims), or 64 when the
// grid would otherwise leave SMs idle (B = 8192, D = 64: 128 row tiles x 1 chunk on 148 SMs -> 256 CTAs, 2 per SM)
constexpr int kMaxFwdCtas = 592;  // 4 x 148; fixed so partial-sum order does not depend on the device
constexpr int kBwdBatch = 16;      // panel rows fetched per batch by the stored-KL backward
constexpr int kStats = 8;         // per-dimension column statistics: SA SB SAB S1 S2 R1 R2 S(a+b)
constexpr int kPivotRows = 32;    // pivots = mean of the first min(B,32) rows

struct Hdr {  // first 256 bytes of the workspace
    unsigned int ticket_prep;
    unsigned int ticket_fwd;
    unsigned int pad[2];
    double m;      // sum_d m_d
    double sumsq;  // last forward result
    unsigned long long dbg[24];  // phase timestamps (ns), written only when built with -DMAE_DEBUG_TIMING
};

#ifdef MAE_DEBUG_TIMING
__device__ __forceinline__ unsigned long long now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define MAE_STAMP(hdr, slot)                                  \
    do {                                                      \
        __syncthreads();                                      \
        if (threadIdx.x == 0 && blockIdx.x == 0 && blockIdx.y == 0) (hdr)->dbg[slot] = now_ns(); \
    } while (0)
#define MAE_STAMP_ANY(hdr, slot)                              \
    do {                                                      \
        if (threadIdx.x == 0) (hdr)->dbg[slot] = now_ns();    \
    } while (0)
#else
#define MAE_STAMP(hdr, slot) do {} while (0)
#define MAE_STAMP_ANY(hdr, slot) do {} while (0)
#endif

struct Layout {
    int64_t Bp, Kp, nblk;
    int64_t prep_rows, prep_grid;  // rows per CTA / CTAs of mpd_prepare_kernel
    int64_t hdr, colstats, pivots, c, l, xp, yp, vp, part_prep, part_fwd, lossg, kl, klt, total, total_nokl;
    int64_t tp, tp_mat;   // tensor-core path: 8 tiled, pre-split panel copies of tp_mat floats each (0: not eligible)
};

constexpr int64_t kBigTileRows = 2048;  // padded batch from which the forward uses 128 x 128 tiles
constexpr int64_t kSmallTileRows = 512;  // padded batch up to which it uses 32 x 32 tiles (latency regime)

Layout make_layout(int64_t B, int64_t D) {
    Layout L;
    L.Bp = round_up(B, kT);
    if (L.Bp >= kBigTileRows) L.Bp = round_up(B, 128);  // the 128 x 128 forward tiles need whole 128-row blocks
    L.Kp = round_up(2 * D, kKC);
    L.nblk = L.Bp / kT;
    // prepare: 8 rows per CTA for small batches (latency), at most ~1184 CTAs for large ones, <= 512 rows per CTA
    {
        int64_t rows = round_up(ceil_div(L.Bp, 1184), 8);
        if (rows > 512) rows = 512;
        L.prep_rows = rows;
        L.prep_grid = ceil_div(L.Bp, rows);
    }
    int64_t off = 0;
    auto take = [&](int64_t bytes) {
        int64_t o = off;
        off += round_up(bytes, 256);
        return o;
    };
    L.hdr = take(256);
    L.colstats = take(8 * D * 8);
    L.pivots = take(4 * D * 8);  // double mu_bar[D], l_bar[D], v_bar[D], 1/v_bar[D]
    L.c = take(L.Bp * 4);
    L.l = take(L.Bp * 4);
    L.xp = take(L.Bp * L.Kp * 4);
    L.yp = take(L.Bp * L.Kp * 4);
    L.vp = take(L.Bp * D * 16);
    L.part_prep = take(L.prep_grid * kStats * D * 8);
    L.part_fwd = take(kMaxFwdCtas * 8);
    L.lossg = take((D + B + 8) * 4);   // mae_mpd_loss_bwd: g_m[D] | g_S | pad[7] | gkl[B]
    L.tp_mat = (L.Bp >= kBigTileRows && L.Kp % 32 == 0 && L.Kp <= 512) ? L.Bp * L.Kp : 0;   // see mpd_tc.cuh
    L.tp = take(8 * L.tp_mat * 4);
    L.total_nokl = off;           // everything the forward needs when the KL tiles are not kept
    L.kl = take(L.Bp * L.Bp * 4);  // optional tail: KL[Bp][Bp] and its transpose (stored-KL backward paths)
    L.klt = take(L.Bp * L.Bp * 4);
    L.total = off;
    return L;
}

template <typename T>
T* at(void* ws, int64_t off) {
    return reinterpret_cast<T*>(static_cast<char*>(ws) + off);
}

// ---------------------------------------------------------------------------------------------- prepare
// grid: G CTAs; CTA `blk` owns rows [blk*R, blk*R+R) (R a multiple of 8, so one warp owns whole rows).  Work is
// organised in chunks of 32 latent dims: lane = dim inside the chunk, warp = row (stride 8).  Every element is
// evaluated once and feeds the panels, the row sums E_i/F_i/KL_i and the per-dimension column partials (fp64 adds).
__global__ void __launch_bounds__(kNT)
mpd_prepare_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride, int B,
                   int D, int Kp, int R, int Bp, float* __restrict__ Xp, float* __restrict__ Yp, float4* __restrict__ Vp,
                   float* __restrict__ evec, float* __restrict__ fvec, double* part, double* __restrict__ colstats,
                   double* pivots, Hdr* hdr, float* __restrict__ m_d_out, float* __restrict__ kl_out,
                   double half_inv_pairs /* 0.5 / (B (B-1)) */) {
    extern __shared__ __align__(16) unsigned char prep_smem[];
    double* row_e = reinterpret_cast<double*>(prep_smem);  // [R]
    double* row_f = row_e + R;                             // [R]
    double* row_k = row_f + R;                             // [R] analytic KL(q_i || N(0,I)) accumulators
    __shared__ double col_part[kNT / 32][kStats][64];
    __shared__ double scratch[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int blk = blockIdx.x;
    const int i0 = blk * R;
    constexpr int kWarps = kNT / 32;
    pdl_wait();                // mu / logvar may come from the gather kernel right before it in the stream
    pdl_launch_dependents();   // the tile kernel may be placed while this grid (and its last-CTA reduction) drains

    MAE_STAMP(hdr, 0);
    for (int r = tid; r < R; r += kNT) {
        row_e[r] = 0.0;
        row_f[r] = 0.0;
        row_k[r] = 0.0;
    }
    const int P = B < kPivotRows ? B : kPivotRows;
    // chunks of 64 latent dims: every lane owns dims d0 = ch*64 + lane and d1 = d0 + 32 and carries both through
    // the same instruction stream, so the two dependent chains (loads -> pivots -> transcendental math) overlap
    const int nchunk = (D + 63) / 64;
    for (int ch = 0; ch < nchunk; ++ch) {
        int dd[2];
        bool dok[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            dd[h] = ch * 64 + h * 32 + lane;
            dok[h] = dd[h] < D;
        }
        // first row of this warp: issued together with the pivot loads (one L2 round trip instead of two)
        float mu_first[2] = {0.f, 0.f}, lv_first[2] = {0.f, 0.f};
        {
            const int i = i0 + w;
#pragma unroll
            for (int h = 0; h < 2; ++h)
                if (dok[h] && w < R && i < B) {
                    mu_first[h] = mu[(int64_t)i * mu_stride + dd[h]];
                    lv_first[h] = lv[(int64_t)i * lv_stride + dd[h]];
                }
        }
        // pivots: mean of the first P rows (every CTA computes the same values in the same order)
        float mu_bar[2], l_bar[2], v_bar[2], ivb[2];
        {
            float pm[2][kPivotRows], pl[2][kPivotRows];
#pragma unroll
            for (int h = 0; h < 2; ++h)
#pragma unroll
                for (int i = 0; i < kPivotRows; ++i) {   // fixed trip count + predication: all loads in flight together
                    pm[h][i] = (dok[h] && i < P) ? mu[(int64_t)i * mu_stride + dd[h]] : 0.f;
                    pl[h][i] = (dok[h] && i < P) ? lv[(int64_t)i * lv_stride + dd[h]] : 0.f;
                }
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                float sm = 0.f, sl = 0.f;
#pragma unroll
                for (int i = 0; i < kPivotRows; ++i) {
                    sm += pm[h][i];
                    sl += pl[h][i];
                }
                mu_bar[h] = sm / (float)P;
                l_bar[h] = sl / (float)P;
                v_bar[h] = expf(l_bar[h]);
                ivb[h] = 1.0f / v_bar[h];
                if (dok[h] && blk == 0 && w == 0) {
                    pivots[dd[h]] = mu_bar[h];
                    pivots[(int64_t)D + dd[h]] = l_bar[h];
                    pivots[2 * (int64_t)D + dd[h]] = v_bar[h];
                    pivots[3 * (int64_t)D + dd[h]] = 1.0 / (double)v_bar[h];
                }
            }
        }
        double acc[2][kStats];
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < kStats; ++q) acc[h][q] = 0.0;
        for (int r = w; r < R; r += kWarps) {
            const int i = i0 + r;
            if (i >= Bp) break;
            float es = 0.f, fs = 0.f, ks = 0.f;
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                if (!dok[h]) continue;
                const int d = dd[h];
                float2 xv = make_float2(0.f, 0.f), yv = xv;
                if (i < B) {
                    const float mu_raw = (r == w) ? mu_first[h] : mu[(int64_t)i * mu_stride + d];
                    const float lv_raw = (r == w) ? lv_first[h] : lv[(int64_t)i * lv_stride + d];
                    const Elem e = make_elem(mu_raw, lv_raw, mu_bar[h], l_bar[h], v_bar[h]);
                    const float sq = e.mc * e.mc;
                    const float s = sq * ivb[h];
                    const float t = e.a + (2.0f * e.b + e.b * e.b) * (1.0f + e.a);  // r^2 v v_bar - 1
                    xv = make_float2(e.a + s, e.mc);
                    yv = make_float2(e.b, -2.0f * e.mc * e.r);
                    Vp[(int64_t)i * D + d] = make_float4(e.mc, e.a, e.b, t);
                    es += e.ema + s;            // all terms >= 0: well-conditioned fp32 sums
                    fs += e.bpa + sq * e.r;
                    // analytic KL term mu^2 + (e^lv - lv - 1), gaussian_encoder.py:219, second part cancellation-free
                    ks += fmaf(mu_raw, mu_raw, expm1_minus_x(lv_raw));
                    acc[h][0] += (double)e.a;
                    acc[h][1] += (double)e.b;
                    acc[h][2] += (double)e.a * (double)e.b;
                    acc[h][3] += (double)e.mc;
                    acc[h][4] += (double)sq;
                    acc[h][5] += (double)e.r * (double)e.mc;
                    acc[h][6] += (double)e.r * (double)sq;
                    acc[h][7] += (double)e.apb;
                }
                *reinterpret_cast<float2*>(Xp + (int64_t)i * Kp + 2 * d) = xv;
                *reinterpret_cast<float2*>(Yp + (int64_t)i * Kp + 2 * d) = yv;
            }
            es = warp_sum(es);
            fs = warp_sum(fs);
            if (kl_out != nullptr) ks = warp_sum(ks);
            if (lane == 0) {
                row_e[r] += (double)es;
                row_f[r] += (double)fs;
                row_k[r] += (double)ks;
            }
        }
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
            for (int q = 0; q < kStats; ++q) col_part[w][q][h * 32 + lane] = acc[h][q];
        __syncthreads();
        for (int o = tid; o < kStats * 64; o += kNT) {   // 8 statistics x 64 dims of this chunk
            const int q = o >> 6, l = o & 63;
            if (ch * 64 + l < D) {
                double t = 0.0;
#pragma unroll
                for (int ww = 0; ww < kWarps; ++ww) t += col_part[ww][q][l];
                part[((int64_t)blk * kStats + q) * D + ch * 64 + l] = t;
            }
        }
        __syncthreads();
    }
    MAE_STAMP(hdr, 1);
    // zero the K padding of the panels and publish the row sums
    for (int r = w; r < R; r += kWarps) {
        const int i = i0 + r;
        if (i >= Bp) break;
        for (int k = 2 * D + lane; k < Kp; k += 32) {
            Xp[(int64_t)i * Kp + k] = 0.f;
            Yp[(int64_t)i * Kp + k] = 0.f;
        }
        if (lane == 0) {
            evec[i] = (float)row_e[r];
            fvec[i] = (float)row_f[r];
            if (kl_out != nullptr && i < B) kl_out[i] = (float)(0.5 * row_k[r]);
        }
    }

    MAE_STAMP(hdr, 2);
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const unsigned int t = atomicAdd(&hdr->ticket_prep, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    MAE_STAMP_ANY(hdr, 3);

    // last CTA: fixed-order sum of the per-CTA partials, one (statistic, dim) item per thread (the CTA count grows with
    // B: a loop over the CTAs by only D threads made this step 8-10 us at B = 512), then the closed-form m_d (fp64)
    const double nb = (double)B;
    const int G = (int)gridDim.x;
    double (*sh)[64] = col_part[0];       // [kStats][64]: the sums of one chunk of 64 dims (the column partials are done)
    double msum = 0.0;
    for (int ch = 0; ch < nchunk; ++ch) {
        for (int o = tid; o < kStats * 64; o += kNT) {
            const int q = o >> 6, l = o & 63, d = ch * 64 + l;
            if (d < D) {
                const double* p = part + (int64_t)q * D + d;      // partials are [bk][q][d]
                const int64_t step = (int64_t)kStats * D;
                double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
                int bk = 0;
                for (; bk + 3 < G; bk += 4) {
                    s0 += __ldcg(p + (bk + 0) * step);
                    s1 += __ldcg(p + (bk + 1) * step);
                    s2 += __ldcg(p + (bk + 2) * step);
                    s3 += __ldcg(p + (bk + 3) * step);
                }
                for (; bk < G; ++bk) s0 += __ldcg(p + bk * step);
                const double t = (s0 + s1) + (s2 + s3);
                sh[q][l] = t;
                colstats[q * (int64_t)D + d] = t;
            }
        }
        __syncthreads();
        const int d = ch * 64 + tid;
        if (tid < 64 && d < D) {
            double s[kStats];
#pragma unroll
            for (int q = 0; q < kStats; ++q) s[q] = sh[q][tid];
            // sum_{i!=j} (v_i r_j - 1)        = SA*SB - SAB + (B-1)*S(a+b)
            // sum_{i!=j} (mu_i-mu_j)^2 r_j    = S2*R0 - 2*S1*R1 + B*R2,   R0 = (B + SB)/v_bar
            const double r0 = (nb + s[1]) * pivots[3 * (int64_t)D + d];
            const double tn = s[0] * s[1] - s[2] + (nb - 1.0) * s[7] + (s[4] * r0 - 2.0 * s[3] * s[5] + nb * s[6]);
            const double md = half_inv_pairs * tn;
            m_d_out[d] = (float)md;
            msum += md;
        }
        __syncthreads();
    }
    msum = block_sum(msum, scratch);
    if (tid == 0) {
        hdr->m = msum;
        hdr->ticket_prep = 0u;
    }
    MAE_STAMP_ANY(hdr, 4);
}

// ------------------------------------------------------------------------------------------ tile helpers
// Stage a 64-row x 16-k slab of a row-major panel into shared memory, transposed to [k][row].
__device__ __forceinline__ void stage_slab(const float* __restrict__ panel, int64_t row_base, int Kp, int k0,
                                           float (*dst)[kLd], int tid) {
    const int r = tid >> 2, kq = (tid & 3) * 4;
    const float4 a = *reinterpret_cast<const float4*>(panel + (row_base + r) * Kp + k0 + kq);
    dst[kq + 0][r] = a.x;
    dst[kq + 1][r] = a.y;
    dst[kq + 2][r] = a.z;
    dst[kq + 3][r] = a.w;
}

__device__ __forceinline__ void fma_tile(const float (*A)[kLd], const float (*Bm)[kLd], int ty, int tx, float acc[4][4]) {
#pragma unroll
    for (int k = 0; k < kKC; ++k) {
        const float4 a = *reinterpret_cast<const float4*>(&A[k][ty * 4]);
        const float4 b = *reinterpret_cast<const float4*>(&Bm[k][tx * 4]);
        const float av[4] = {a.x, a.y, a.z, a.w};
        const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
}

// ---------------------------------------------------------------------------------------------- forward
// TM x TM micro-tile per thread, 16 x 16 threads -> (16*TM)^2 tile: TM = 4 (64 x 64) for throughput, TM = 2 (32 x 32)
// for small batches where a handful of 64-tiles would leave the chip empty and serialise ~2k FFMA per thread.
template <int TM>
__global__ void __launch_bounds__(kNT)
mpd_fwd_kernel(const float* __restrict__ Xp, const float* __restrict__ Yp, const float* __restrict__ evec,
               const float* __restrict__ fvec, Hdr* hdr, double* part, int B, int Kp, int row0, int row1, int tr0,
               int ntr, int ntc, float* __restrict__ KL, float* __restrict__ KLT, int Bp, double* __restrict__ sumsq_out,
               float* __restrict__ S_out, int tc0, int kl_base) {
    // tc0: first column tile (0 except for the transposed pass of a row block); kl_base: global row stored in row 0 of the
    // KL / KLT buffers (row blocks keep only their rows); sumsq_out == nullptr: store pass only, no reduction
    constexpr int T = 16 * TM;       // tile edge
    constexpr int LD = T + 4;        // padded leading dimension
    constexpr int NLOAD = T * 4;     // float4 loads per operand slab (T rows x 16 k)
    constexpr int LPT = (2 * NLOAD) / kNT;   // slab loads per thread per K step: 1 (TM=2), 2 (TM=4), 4 (TM=8)
    __shared__ __align__(16) float As[2][kKC][LD];
    __shared__ __align__(16) float Bs[2][kKC][LD];
    __shared__ double scratch[32];
    __shared__ bool is_last;
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    pdl_wait();                // panels, E / F and m come from mpd_prepare_kernel
    pdl_launch_dependents();   // (after the wait: successors are placed one kernel ahead, not a whole chain)
    MAE_STAMP(hdr, 8);
    const float m = (float)hdr->m;
    double total = 0.0;
    const int ntiles = ntr * ntc;
    // micro-tile coordinates: contiguous TM x TM for TM <= 4; for TM = 8 two groups of 4, 64 apart, so that the
    // 16-byte shared-memory reads of a quarter warp fall into distinct banks
    auto rc = [](int t, int i) { return TM == 8 ? ((i < 4) ? t * 4 + i : 64 + t * 4 + (i - 4)) : t * TM + i; };
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
        const int ti = tr0 + tile / ntc, tj = tc0 + tile % ntc;
        float acc[TM][TM];
#pragma unroll
        for (int i = 0; i < TM; ++i)
#pragma unroll
            for (int j = 0; j < TM; ++j) acc[i][j] = 0.f;
        // K loop: the next slab is fetched into registers while the current one is multiplied
        {
            const float* src[LPT];
            float (*dstp[LPT])[kKC][LD];
            int rr[LPT], kk[LPT];
#pragma unroll
            for (int q = 0; q < LPT; ++q) {
                const int item = tid + q * kNT;          // [0, NLOAD): X slab, [NLOAD, 2 NLOAD): Y slab
                const bool isx = item < NLOAD;
                const int idx = isx ? item : item - NLOAD;
                rr[q] = idx >> 2;
                kk[q] = (idx & 3) * 4;
                src[q] = (isx ? Xp + ((int64_t)ti * T + rr[q]) * Kp : Yp + ((int64_t)tj * T + rr[q]) * Kp) + kk[q];
                dstp[q] = isx ? As : Bs;
            }
            float4 reg[LPT];
#pragma unroll
            for (int q = 0; q < LPT; ++q) reg[q] = *reinterpret_cast<const float4*>(src[q]);
            int buf = 0;
            for (int k0 = 0; k0 < Kp; k0 += kKC) {
#pragma unroll
                for (int q = 0; q < LPT; ++q) {
                    dstp[q][buf][kk[q] + 0][rr[q]] = reg[q].x;
                    dstp[q][buf][kk[q] + 1][rr[q]] = reg[q].y;
                    dstp[q][buf][kk[q] + 2][rr[q]] = reg[q].z;
                    dstp[q][buf][kk[q] + 3][rr[q]] = reg[q].w;
                }
                __syncthreads();
                if (k0 + kKC < Kp) {
#pragma unroll
                    for (int q = 0; q < LPT; ++q) reg[q] = *reinterpret_cast<const float4*>(src[q] + k0 + kKC);
                }
#pragma unroll
                for (int k = 0; k < kKC; ++k) {
                    float av[TM], bv[TM];
#pragma unroll
                    for (int i = 0; i < TM; ++i) {
                        av[i] = As[buf][k][rc(ty, i)];
                        bv[i] = Bs[buf][k][rc(tx, i)];
                    }
#pragma unroll
                    for (int i = 0; i < TM; ++i)
#pragma unroll
                        for (int j = 0; j < TM; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
                }
                buf ^= 1;
            }
            __syncthreads();  // all reads of the last slab done before the next tile overwrites the buffers
        }
        MAE_STAMP(hdr, 9);
        float psum = 0.f;
        float fj[TM];
#pragma unroll
        for (int jj = 0; jj < TM; ++jj) fj[jj] = fvec[tj * T + rc(tx, jj)];
        float klv[TM][TM];
#pragma unroll
        for (int ii = 0; ii < TM; ++ii) {
            const int i = ti * T + rc(ty, ii);
            const float ei = evec[i];
            const bool row_ok = (i >= row0) && (i < row1);
#pragma unroll
            for (int jj = 0; jj < TM; ++jj) {
                const int j = tj * T + rc(tx, jj);
                const float kl = 0.5f * (acc[ii][jj] + ei + fj[jj]);
                klv[ii][jj] = kl;
                if (row_ok && j < B && j != i) {
                    const float dev = kl - m;
                    psum = fmaf(dev, dev, psum);
                }
            }
            if (KL != nullptr) {
                // a thread's columns come in groups of 4 consecutive ones (rc): 16-byte stores, a quarter warp covers 256
                // contiguous bytes of the row (scalar stores wrote every 32-byte sector four times over)
                if constexpr (TM >= 4) {
#pragma unroll
                    for (int gq = 0; gq < TM / 4; ++gq)
                        *reinterpret_cast<float4*>(KL + (int64_t)(i - kl_base) * Bp + tj * T + rc(tx, gq * 4)) =
                            make_float4(klv[ii][gq * 4], klv[ii][gq * 4 + 1], klv[ii][gq * 4 + 2], klv[ii][gq * 4 + 3]);
                } else {
#pragma unroll
                    for (int jj = 0; jj < TM; ++jj) KL[(int64_t)(i - kl_base) * Bp + tj * T + rc(tx, jj)] = klv[ii][jj];
                }
            }
        }
        if (KLT != nullptr) {
#pragma unroll
            for (int jj = 0; jj < TM; ++jj) {
                float* dst = KLT + (int64_t)(tj * T + rc(tx, jj) - kl_base) * Bp + ti * T;
                if constexpr (TM >= 4) {
#pragma unroll
                    for (int gq = 0; gq < TM / 4; ++gq)
                        *reinterpret_cast<float4*>(dst + rc(ty, gq * 4)) =
                            make_float4(klv[gq * 4][jj], klv[gq * 4 + 1][jj], klv[gq * 4 + 2][jj], klv[gq * 4 + 3][jj]);
                } else {
#pragma unroll
                    for (int ii = 0; ii < TM; ++ii) dst[rc(ty, ii)] = klv[ii][jj];
                }
            }
        }
        total += (double)psum;
    }
    MAE_STAMP(hdr, 10);
    if (sumsq_out == nullptr) return;   // store-only pass (KL_ji of a row block)
    total = block_sum(total, scratch);
    if (gridDim.x == 1) {  // single CTA: nothing to combine
        is_last = true;
        if (tid == 0) part[0] = total;
    } else if (tid == 0) {
        part[blockIdx.x] = total;
        __threadfence();
        const unsigned int t = atomicAdd(&hdr->ticket_fwd, 1u);
        is_last = (t == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && tid < 32) {
        __threadfence();
        // warp 0: lane-strided fixed-order sums of the per-CTA partials, then a fixed xor tree
        double s = 0.0;
        if (gridDim.x == 1) s = total;
        else {
            for (int i = tid; i < (int)gridDim.x; i += 32) s += __ldcg(part + i);
            s = warp_sum(s);
        }
        if (tid == 0) {
            hdr->sumsq = s;
            sumsq_out[0] = s;
            if (S_out != nullptr) {
                const double nb = (double)B;
                S_out[0] = (float)sqrt(s / (nb * (nb - 1.0) - 1.0));
            }
            hdr->ticket_fwd = 0u;
            MAE_STAMP_ANY(hdr, 11);
        }
    }
}

__global__ void mpd_finalize_kernel(const double* __restrict__ sumsq, double nb, float* __restrict__ S) {
    S[0] = (float)sqrt(sumsq[0] / (nb * (nb - 1.0) - 1.0));
}

// upstream gradients of the regulariser's closed-form consumers (mae.py:132-139), evaluated on the device so that the
// MPD backward can be enqueued in the FORWARD pass behind the tile kernel (no round trip through the objective):
//   reg = clamp(kl_weight * sum_b KL_b, min = free_bits) - eta * sum_d logsigmoid(m_d) + gamma * S
__global__ void __launch_bounds__(kNT)
mpd_loss_grads_kernel(const float* __restrict__ m_d, const float* __restrict__ S, const float* __restrict__ kl_all, int B, int D,
                      float eta, float gamma, float free_bits, float kl_weight, float* __restrict__ g_m, float* __restrict__ g_S,
                      float* __restrict__ gkl, float* __restrict__ reg) {
    __shared__ double scratch[32];
    const int tid = threadIdx.x;
    pdl_wait();                // S comes from the tile kernel right before it
    pdl_launch_dependents();
    double ks = 0.0, ls = 0.0;
    if (kl_all != nullptr)
        for (int i = tid; i < B; i += kNT) ks += (double)kl_all[i];
    for (int d = tid; d < D; d += kNT) {
        const float v = m_d[d];
        const float e = expf(-fabsf(v));
        const float sig_neg = v >= 0.f ? e / (1.0f + e) : 1.0f / (1.0f + e);
        g_m[d] = -eta * sig_neg;
        ls += (double)(fminf(v, 0.f) - log1pf(e));
    }
    ks = block_sum(ks, scratch);
    ls = block_sum(ls, scratch);
    const float mean_kl = (float)(ks * (double)kl_weight);
    const bool kl_on = kl_all != nullptr && kl_weight > 0.f;
    const float gk = (kl_on && mean_kl >= free_bits) ? kl_weight : 0.f;   // clamp(min) passes at equality
    if (gkl != nullptr)
        for (int i = tid; i < B; i += kNT) gkl[i] = gk;
    if (tid == 0) {
        g_S[0] = gamma > 0.f ? gamma : 0.f;
        if (reg != nullptr) {
            float r = kl_on ? fmaxf(mean_kl, free_bits) : 0.f;
            if (eta > 0.f) r -= eta * (float)ls;
            if (gamma > 0.f) r += gamma * S[0];
            reg[0] = r;
        }
    }
}

// --------------------------------------------------------------------------------- backward: common parts
// weight scale of (KL - m): g_S / ((B^2 - B - 1) * S)     (SURVEY.md §8a "A3-bwd")
__device__ __forceinline__ float weight_scale(const float* g_S, const float* S, int B) {
    if (g_S == nullptr) return 0.f;
    const double s = (double)S[0];
    if (!(s > 0.0)) return 0.f;
    const double nb = (double)B;
    return (float)((double)g_S[0] / ((nb * nb - nb - 1.0) * s));
}

// ------------------------------------------------------------------------- backward, stored-KL (small B)
// one warp per (row i, 32 latent dims); walks all partners j.  V = {mc, a, b, t} (see the header comment):
//   d KL_ij/d mu_i = dl*r_j,  d KL_ji/d mu_i = dl*r_i,                        r = (1+b)/v_bar
//   2 d KL_ij/d lv_i = v_i r_j - 1 = a_i b_j + a_i + b_j
//   2 d KL_ji/d lv_i = 1 - (v_j + dl^2) r_i^2 v_i = -(a_j + t_i + a_j t_i) - dl^2 (1+t_i)/v_bar
// PS warps share one (row, 32-dim) item, each walking every PS-th block of 32 partners; their partial sums meet in
// shared memory (fixed order).  PS = 1 for small batches, larger when a row has hundreds of partners (sharded runs
// keep 64 own rows but the partner count grows with the world size).
template <int PS>
__global__ void __launch_bounds__(kNT)
mpd_bwd_small_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                     const float4* __restrict__ Vp, const float* __restrict__ KL, const float* __restrict__ KLT, int Bp,
                     const double* __restrict__ colstats, const double* __restrict__ pivots, const Hdr* __restrict__ hdr,
                     const float* __restrict__ g_m, const float* __restrict__ g_S, const float* __restrict__ S,
                     const float* __restrict__ gkl, int B, int D, int row0, int nrows, float* __restrict__ dmu,
                     float* __restrict__ dlv, int accumulate) {
    constexpr int kItems = (kNT / 32) / PS;   // (row, dim-chunk) items per CTA
    __shared__ float red[2][kNT / 32][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    pdl_wait();                // g_m / g_S / gkl may come from mpd_loss_grads_kernel right before it
    const int item_local = w / PS, ps = w % PS;
    const int nd32 = (D + 31) / 32;
    const int64_t item = (int64_t)blockIdx.x * kItems + item_local;
    const bool item_ok = item < (int64_t)nrows * nd32;
    const int i = row0 + (int)((item_ok ? item : 0) / nd32);
    const int d = (int)((item_ok ? item : 0) % nd32) * 32 + lane;
    const bool active = item_ok && d < D;
    const float m = (float)hdr->m;
    const float beta = weight_scale(g_S, S, B);
    const float4 zero4 = make_float4(0.f, 0.f, 0.f, 0.f);
    const float4 pi = active ? Vp[(int64_t)i * D + d] : zero4;  // {mc, a, b, t} of row i
    float gmu = 0.f, glv = 0.f;
    if (beta != 0.f && item_ok) {
        const float ivb = active ? (float)pivots[3 * (int64_t)D + d] : 0.f;
        const float* kl_row = KL + (int64_t)i * Bp;    // KL_ij
        const float* klt_row = KLT + (int64_t)i * Bp;  // KL_ji
        const float opb_i = 1.0f + pi.z, opt_i = 1.0f + pi.w, opa_i = 1.0f + pi.y, oi = opt_i * ivb;
        const float4* vcol = Vp + d;
        for (int jb = ps * 32; jb < B; jb += 32 * PS) {
            // the 32 lanes fetch the weights of 32 partners at once (one coalesced line each), then broadcast
            const int jl = jb + lane;
            const bool ok = jl < B && jl != i;
            const float w1l = ok ? beta * (kl_row[jl] - m) : 0.f;
            const float w2l = ok ? beta * (klt_row[jl] - m) : 0.f;
#pragma unroll
            for (int j0 = 0; j0 < 32; j0 += kBwdBatch) {
                // explicit batch: kBwdBatch independent 16-byte panel loads in flight before any of them is consumed
                float4 pj[kBwdBatch];
#pragma unroll
                for (int u = 0; u < kBwdBatch; ++u)
                    pj[u] = (active && jb + j0 + u < B) ? __ldg(vcol + (int64_t)(jb + j0 + u) * D) : zero4;
#pragma unroll
                for (int u = 0; u < kBwdBatch; ++u) {
                    const float w1 = __shfl_sync(0xffffffffu, w1l, j0 + u);
                    const float w2 = __shfl_sync(0xffffffffu, w2l, j0 + u);
                    const float dl = pi.x - pj[u].x;
                    gmu = fmaf(dl, fmaf(w1, 1.0f + pj[u].z, w2 * opb_i), gmu);
                    glv = fmaf(w1, fmaf(opa_i, pj[u].z, pi.y), glv);                                  // (a_i+1) b_j + a_i
                    glv = fmaf(-w2, fmaf(opt_i, pj[u].y, fmaf(dl * dl, oi, pi.w)), glv);               // (t_i+1) a_j + t_i + dl^2 o_i
                }
            }
        }
        gmu *= ivb;
        glv *= 0.5f;
    }
    if (PS > 1) {
        red[0][w][lane] = gmu;
        red[1][w][lane] = glv;
        __syncthreads();
        if (ps != 0) return;
        gmu = 0.f;
        glv = 0.f;
#pragma unroll
        for (int q = 0; q < PS; ++q) {
            gmu += red[0][item_local * PS + q][lane];
            glv += red[1][item_local * PS + q][lane];
        }
    }
    if (!active) return;
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

// ----------------------------------------------------------------------- backward, recompute tiles (large B)
template <int OC>
struct BwdSmem {
    float W1T[kT][kLd], W2T[kT][kLd];  // [j][i]
    float PY[kT][OC + 4], PX[kT][OC + 4];  // column panels of block J, this CTA's K chunk
    float XI[kKC][kLd], YI[kKC][kLd], XJ[kKC][kLd], YJ[kKC][kLd];  // recompute path only (last: the stored path omits them)
};
template <int OC>
constexpr size_t bwd_stored_smem() { return offsetof(BwdSmem<OC>, XI); }

// grid: (row tiles, K chunks of kOC columns).  CTA owns rows [i0,i0+64) and gradient dims of its chunk.
// STORED: the KL_ij / KL_ji tiles are read from the matrices the forward kept (one streaming pass over 2 B^2 floats)
// instead of being recomputed (two K = 2D contractions per tile pair), which halves the backward's FFMA work.
template <bool STORED, int OC, int MINB>
__global__ void __launch_bounds__(kNT, MINB)
mpd_bwd_tile_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                    const float* __restrict__ Xp, const float* __restrict__ Yp, const float* __restrict__ evec,
                    const float* __restrict__ fvec, const float4* __restrict__ Vp, const double* __restrict__ colstats,
                    const double* __restrict__ pivots, const Hdr* __restrict__ hdr, const float* __restrict__ g_m, const float* __restrict__ g_S,
                    const float* __restrict__ S, const float* __restrict__ gkl, int B, int D, int Kp, int row0, int row1,
                    int tr0, int ntc, float* __restrict__ dmu, float* __restrict__ dlv, int accumulate,
                    const float* __restrict__ KL, const float* __restrict__ KLT, int Bp) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int NG = OC / 16;   // gradient columns per thread: groups of 4 at tx*4 (+64 for OC = 128)
    BwdSmem<OC>& sm = *reinterpret_cast<BwdSmem<OC>*>(smem_raw);
    const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
    const int ti = tr0 + blockIdx.x;
    const int i0 = ti * kT;
    const int kbase = blockIdx.y * OC;
    const float m = (float)hdr->m;
    const float beta = weight_scale(g_S, S, B);

    float G1[4][NG], G2[4][NG], rs1[4], rs2[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        rs1[i] = rs2[i] = 0.f;
#pragma unroll
        for (int j = 0; j < NG; ++j) G1[i][j] = G2[i][j] = 0.f;
    }
    float ei[4], fi[4];
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        ei[ii] = evec[i0 + ty * 4 + ii];
        fi[ii] = fvec[i0 + ty * 4 + ii];
    }

    if (beta != 0.f) {
        for (int tj = 0; tj < ntc; ++tj) {
            const int j0 = tj * kT;
            float a1[4][4], a2[4][4];
            if (STORED) {
                // KL_ij = KL[i][j], KL_ji = KLT[i][j]: both row-contiguous in j -> 16-byte loads
#pragma unroll
                for (int ii = 0; ii < 4; ++ii) {
                    const int64_t o = (int64_t)(i0 + ty * 4 + ii) * Bp + j0 + tx * 4;
                    const float4 k1 = __ldcs(reinterpret_cast<const float4*>(KL + o));
                    const float4 k2 = __ldcs(reinterpret_cast<const float4*>(KLT + o));
                    a1[ii][0] = k1.x; a1[ii][1] = k1.y; a1[ii][2] = k1.z; a1[ii][3] = k1.w;
                    a2[ii][0] = k2.x; a2[ii][1] = k2.y; a2[ii][2] = k2.z; a2[ii][3] = k2.w;
                }
            } else {
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int j = 0; j < 4; ++j) a1[i][j] = a2[i][j] = 0.f;
                for (int k0 = 0; k0 < Kp; k0 += kKC) {
                    stage_slab(Xp, (int64_t)i0, Kp, k0, sm.XI, tid);
                    stage_slab(Yp, (int64_t)i0, Kp, k0, sm.YI, tid);
                    stage_slab(Xp, (int64_t)j0, Kp, k0, sm.XJ, tid);
                    stage_slab(Yp, (int64_t)j0, Kp, k0, sm.YJ, tid);
                    __syncthreads();
                    fma_tile(sm.XI, sm.YJ, ty, tx, a1);  // x~_i . y~_j  -> KL_ij
                    fma_tile(sm.YI, sm.XJ, ty, tx, a2);  // y~_i . x~_j  -> KL_ji
                    __syncthreads();
                }
            }
            // weights; W1T/W2T stored [j][i]
            const int jb = j0 + tx * 4;
            const float4 e4 = *reinterpret_cast<const float4*>(evec + jb);
            const float4 f4 = *reinterpret_cast<const float4*>(fvec + jb);
            const float ej[4] = {e4.x, e4.y, e4.z, e4.w};
            const float fj[4] = {f4.x, f4.y, f4.z, f4.w};
            float w1[4][4], w2[4][4];
#pragma unroll
            for (int ii = 0; ii < 4; ++ii) {
                const int i = i0 + ty * 4 + ii;
                const bool row_ok = (i >= row0) && (i < row1);
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                    const int j = jb + jj;
                    const bool ok = row_ok && j < B && j != i;
                    const float kij = STORED ? a1[ii][jj] : 0.5f * (a1[ii][jj] + ei[ii] + fj[jj]);
                    const float kji = STORED ? a2[ii][jj] : 0.5f * (a2[ii][jj] + ej[jj] + fi[ii]);
                    w1[ii][jj] = ok ? beta * (kij - m) : 0.f;
                    w2[ii][jj] = ok ? beta * (kji - m) : 0.f;
                    rs1[ii] += w1[ii][jj];
                    rs2[ii] += w2[ii][jj];
                }
            }
#pragma unroll
            for (int jj = 0; jj < 4; ++jj) {
                *reinterpret_cast<float4*>(&sm.W1T[tx * 4 + jj][ty * 4]) = make_float4(w1[0][jj], w1[1][jj], w1[2][jj], w1[3][jj]);
                *reinterpret_cast<float4*>(&sm.W2T[tx * 4 + jj][ty * 4]) = make_float4(w2[0][jj], w2[1][jj], w2[2][jj], w2[3][jj]);
            }
            // column panels of block J restricted to this CTA's K chunk: 64 rows x 128 columns each
            for (int q = tid; q < kT * (OC / 4); q += kNT) {
                const int r = q / (OC / 4), c = (q % (OC / 4)) * 4;
                float4 y = make_float4(0.f, 0.f, 0.f, 0.f), x = y;
                if (kbase + c < Kp) {
                    y = *reinterpret_cast<const float4*>(Yp + (int64_t)(j0 + r) * Kp + kbase + c);
                    x = *reinterpret_cast<const float4*>(Xp + (int64_t)(j0 + r) * Kp + kbase + c);
                }
                *reinterpret_cast<float4*>(&sm.PY[r][c]) = y;
                *reinterpret_cast<float4*>(&sm.PX[r][c]) = x;
            }
            __syncthreads();
            // G1[i][k] += sum_j W1[i][j] * y~_j[k];  G2[i][k] += sum_j W2[i][j] * x~_j[k]
#pragma unroll 4
            for (int j = 0; j < kT; ++j) {
                const float4 wa = *reinterpret_cast<const float4*>(&sm.W1T[j][ty * 4]);
                const float4 wb = *reinterpret_cast<const float4*>(&sm.W2T[j][ty * 4]);
                float yv[NG], xv[NG];
#pragma unroll
                for (int gq = 0; gq < NG / 4; ++gq) {
                    const float4 yq = *reinterpret_cast<const float4*>(&sm.PY[j][gq * 64 + tx * 4]);
                    const float4 xq = *reinterpret_cast<const float4*>(&sm.PX[j][gq * 64 + tx * 4]);
                    yv[gq * 4 + 0] = yq.x; yv[gq * 4 + 1] = yq.y; yv[gq * 4 + 2] = yq.z; yv[gq * 4 + 3] = yq.w;
                    xv[gq * 4 + 0] = xq.x; xv[gq * 4 + 1] = xq.y; xv[gq * 4 + 2] = xq.z; xv[gq * 4 + 3] = xq.w;
                }
                const float wav[4] = {wa.x, wa.y, wa.z, wa.w};
                const float wbv[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                for (int ii = 0; ii < 4; ++ii)
#pragma unroll
                    for (int c = 0; c < NG; ++c) {
                        G1[ii][c] = fmaf(wav[ii], yv[c], G1[ii][c]);
                        G2[ii][c] = fmaf(wbv[ii], xv[c], G2[ii][c]);
                    }
            }
            __syncthreads();
        }
        // row sums of W: reduce the per-thread partials over the 16 threads (tx) that share a row group
#pragma unroll
        for (int ii = 0; ii < 4; ++ii) {
#pragma unroll
            for (int o = 8; o > 0; o >>= 1) {
                rs1[ii] += __shfl_xor_sync(0xffffffffu, rs1[ii], o);
                rs2[ii] += __shfl_xor_sync(0xffffffffu, rs2[ii], o);
            }
        }
    }

    // epilogue: this thread holds, for 4 rows, K columns {tx*4..tx*4+3} and {64+tx*4..}: 4 latent dims per row
#pragma unroll
    for (int ii = 0; ii < 4; ++ii) {
        const int i = i0 + ty * 4 + ii;
        if (i < row0 || i >= row1) continue;
#pragma unroll
        for (int h = 0; h < NG / 2; ++h) {
            const int cb = (h >> 1) * 4 + (h & 1) * 2;           // register column of the pair (2d, 2d+1)
            const int kcol = kbase + (h >> 1) * 64 + tx * 4 + (h & 1) * 2;
            const int d = kcol >> 1;
            if (d >= D) continue;
            const float4 pi = Vp[(int64_t)i * D + d];  // {mc, a, b, t}
            const float ivb = (float)pivots[3 * (int64_t)D + d];
            const float g1r = G1[ii][cb], g1m = G1[ii][cb + 1], g2x = G2[ii][cb], g2m = G2[ii][cb + 1];
            const float mc = pi.x, a = pi.y, ri = (1.0f + pi.z) * ivb, t = pi.w;
            // row side: mu_i sum_j W r_j - sum_j W mu_j r_j ; column side: -r_i (sum_j W' mu_j - mu_i sum_j W')
            float gmu = mc * (g1r + rs1[ii]) * ivb + 0.5f * g1m - ri * (g2m - mc * rs2[ii]);
            float glv = 0.5f * ((1.0f + a) * g1r + a * rs1[ii]) +
                        0.5f * (-t * rs2[ii] - (1.0f + t) * (g2x - (2.0f * mc * g2m - mc * mc * rs2[ii]) * ivb));
            mean_part(colstats, pivots, g_m, B, D, d, pi, mu[(int64_t)i * mu_stride + d], lv[(int64_t)i * lv_stride + d], gmu,
                      glv, gkl ? gkl + i : nullptr);
            const int64_t o = (int64_t)(i - row0) * D + d;
            if (accumulate) {
                dmu[o] += gmu;
                dlv[o] += glv;
            } else {
                dmu[o] = gmu;
                dlv[o] = glv;
            }
        }
    }
}

#include "mpd_tc.cuh"

int g_mpd_tc = []() { const char* e = getenv("MAE_MPD_TC"); return (e && atoi(e) == 0) ? 0 : 1; }();

// Stored-KL buffers of a row block [row0, row0 + nrows): only the rows [base, base + rows_pad) of KL (= KL_ij, own i) and
// of KL^T (= KL_ji, own i) are kept, base a multiple of the largest tile edge.  The full batch keeps Bp rows (base 0).
struct KlGeom {
    int64_t base, rows_pad, kl_off, klt_off, total;
    int64_t tc_split, gp_off, rs_off;   // tensor-core backward: partner splits, partial G1 / G2 sums, partial row sums
};
// partner splits of the tensor-core backward.  The kernel runs one CTA per SM (131 KB of shared memory, 576 threads), so its
// time is whole rounds of 148 CTAs: pick the split count whose CTA count (row tiles x column chunks x splits) fills its last
// round best (B = 16384 / D = 64: 3 splits = 2.6 rounds -> 8 splits = 6.9 rounds; D = 256: 1 split = 3.5 rounds -> 2 splits =
// 6.9), the smaller count on a tie.  A split keeps at least one accumulation group (4 partner tiles).
int64_t tc_splits(int64_t row_tiles, int64_t col_chunks, int64_t partner_tiles) {
    static const int forced = []() { const char* e = getenv("MAE_MPD_TC_SPLIT"); return e ? atoi(e) : 0; }();   // tuning aid
    if (forced > 0) return forced > partner_tiles ? partner_tiles : forced;
    const int64_t base = row_tiles * col_chunks;
    int64_t best = 1;
    double best_eff = 0.0;
    for (int64_t s = 1; s <= 8 && s <= partner_tiles; ++s) {
        if (s > 1 && ceil_div(partner_tiles, s) < 4) break;
        const int64_t ctas = base * s, rounds = ceil_div(ctas, 148);
        const double eff = (double)ctas / (double)(rounds * 148);
        if (eff > best_eff + 0.02) {
            best_eff = eff;
            best = s;
        }
    }
    return best;
}
// chunks of 32 partners per TMEM accumulation of the tensor-core backward (default 16 = 512 partners = 192 accumulating
// MMAs): the tensor core's accumulator does not round to nearest, so the error of a running sum grows with its length
// (measured against the FFMA path at B = 32768: 7e-5 of the gradient scale with one accumulation over all partners, 9e-6
// with 2048 partners per accumulation, 2e-6 with 512, independent of B); the groups are added in fp32 by the row owners
int tc_group_chunks() {
    static const int v = []() { const char* e = getenv("MAE_MPD_TC_GROUP"); const int g = e ? atoi(e) : 16; return g < 1 ? 1 : g; }();
    return v;
}
KlGeom kl_geom(const Layout& L, int64_t B, int64_t row0, int64_t nrows) {
    KlGeom g;
    if (row0 == 0 && nrows == B) {
        g.base = 0;
        g.rows_pad = L.Bp;
    } else {
        g.base = (row0 / 128) * 128;
        int64_t end = round_up(row0 + nrows, 128);
        if (end > L.Bp) end = L.Bp;
        g.rows_pad = end - g.base;
    }
    g.kl_off = L.kl;
    g.klt_off = L.kl + round_up(g.rows_pad * L.Bp * 4, 256);
    g.total = g.klt_off + round_up(g.rows_pad * L.Bp * 4, 256);
    g.tc_split = 0;
    g.gp_off = g.rs_off = g.total;
    if (L.tp_mat > 0 && L.Kp % 128 == 0) {
        g.tc_split = tc_splits(g.rows_pad / 128, L.Kp / 128, L.Bp / 128);
        g.gp_off = g.total;
        g.rs_off = g.gp_off + round_up(2 * g.tc_split * g.rows_pad * L.Kp * 4, 256);
        g.total = g.rs_off + round_up(4 * g.tc_split * g.rows_pad * 4, 256);
    }
    return g;
}

int check_common(int64_t B, int64_t D, const void* ws, int64_t ws_bytes, Layout& L) {
    MAE_REQUIRE(ws != nullptr, MAE_ERR_NULL);
    MAE_REQUIRE(B >= 2 && D >= 1 && B <= (1ll << 22) && D <= (1ll << 20), MAE_ERR_SHAPE);
    MAE_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, MAE_ERR_ALIGN);
    L = make_layout(B, D);
    MAE_REQUIRE(ws_bytes >= L.total_nokl, MAE_ERR_WORKSPACE);
    return MAE_OK;
}

}  // namespace
}  // namespace mae

// development aid (not part of include/mae_b200.h): per-role cycle counters of the last profiled tensor-core forward
extern "C" int mae_debug_mpd_prof(long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, mae::g_mpd_prof, sizeof(long long) * 64);
}

// development aid (not part of include/mae_b200.h): the partner-split count the tensor-core backward would use (host logic
// only, callable without a GPU; tests/test_boundary_cpu.py checks the whole-round property)
extern "C" long long mae_debug_mpd_tc_splits(long long row_tiles, long long col_chunks, long long partner_tiles) {
    return mae::tc_splits(row_tiles, col_chunks, partner_tiles);
}

extern "C" int mae_set_mpd_tensor_cores(int enabled) {
    const int before = mae::g_mpd_tc;
    mae::g_mpd_tc = enabled ? 1 : 0;
    return before;
}

extern "C" int64_t mae_mpd_workspace_bytes(int64_t B, int64_t D, int store_kl) {
    if (B < 2 || D < 1) return 0;
    const mae::Layout L = mae::make_layout(B, D);
    return store_kl ? mae::kl_geom(L, B, 0, B).total : L.total_nokl;
}

extern "C" int64_t mae_mpd_workspace_bytes_rows(int64_t B, int64_t D, int64_t row0, int64_t nrows) {
    if (B < 2 || D < 1 || row0 < 0 || nrows < 1 || row0 + nrows > B) return 0;
    const mae::Layout L = mae::make_layout(B, D);
    return mae::kl_geom(L, B, row0, nrows).total;
}

extern "C" int mae_mpd_prepare(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride, int64_t B,
                               int64_t D, void* workspace, int64_t workspace_bytes, float* m_d, float* kl,
                               mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(mu && logvar && m_d, MAE_ERR_NULL);
    Layout L;
    int rc = check_common(B, D, workspace, workspace_bytes, L);
    if (rc != MAE_OK) return rc;
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    launch_pdl(mpd_prepare_kernel, dim3((unsigned)L.prep_grid), dim3(kNT), (size_t)(3 * L.prep_rows * sizeof(double)), as_stream(stream),
        mu, mu_stride, logvar, logvar_stride, (int)B, (int)D, (int)L.Kp, (int)L.prep_rows, (int)L.Bp,
        at<float>(workspace, L.xp), at<float>(workspace, L.yp),
        at<float4>(workspace, L.vp), at<float>(workspace, L.c), at<float>(workspace, L.l), at<double>(workspace, L.part_prep),
        at<double>(workspace, L.colstats), at<double>(workspace, L.pivots), at<Hdr>(workspace, L.hdr), m_d, kl,
        0.5 / ((double)B * ((double)B - 1.0)));
    if (g_mpd_tc && L.tp_mat > 0) {
        // tensor-core tiles: pre-split copies of the panels in the K-major canonical tile layout (mpd_tc.cuh)
        const int64_t n = L.Bp * (L.Kp / 4);
        mpd_tile_panels_kernel<<<(unsigned)ceil_div(n, 256), 256, 0, as_stream(stream)>>>(
            at<float>(workspace, L.xp), at<float>(workspace, L.yp), (int)L.Bp, (int)L.Kp, at<float>(workspace, L.tp), L.tp_mat,
            L.Kp % 128 == 0 ? 1 : 0);
    }
    return launch_status();
}

extern "C" int mae_mpd_fwd(void* workspace, int64_t workspace_bytes, int64_t B, int64_t D, int64_t row0, int64_t nrows,
                           int store_kl, double* sumsq, float* S, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(sumsq, MAE_ERR_NULL);
    Layout L;
    int rc = check_common(B, D, workspace, workspace_bytes, L);
    if (rc != MAE_OK) return rc;
    MAE_REQUIRE(row0 >= 0 && nrows >= 1 && row0 + nrows <= B, MAE_ERR_SHAPE);
    const bool full = row0 == 0 && nrows == B;
    const KlGeom kg = kl_geom(L, B, row0, nrows);
    if (store_kl) MAE_REQUIRE(workspace_bytes >= kg.total, MAE_ERR_WORKSPACE);
    float* klp = store_kl ? at<float>(workspace, kg.kl_off) : nullptr;
    float* kltp = store_kl ? at<float>(workspace, kg.klt_off) : nullptr;
    auto launch = [&](auto tm_tag) {
        constexpr int TM = decltype(tm_tag)::value;
        constexpr int T = 16 * TM;
        const int tr0 = (int)(row0 / T);
        const int ntr = (int)(ceil_div(row0 + nrows, T) - tr0);
        const int ntc = (int)(L.Bp / T);
        const int64_t ntiles = (int64_t)ntr * ntc;
        const int grid = (int)(ntiles < kMaxFwdCtas ? ntiles : kMaxFwdCtas);
        // pass 1: tiles (own rows) x (all columns): sum of squares, KL_ij rows; the full batch also gets KL^T from the same tiles
        launch_pdl(mpd_fwd_kernel<TM>, dim3((unsigned)grid), dim3(kNT), 0, as_stream(stream),
            at<float>(workspace, L.xp), at<float>(workspace, L.yp), at<float>(workspace, L.c), at<float>(workspace, L.l),
            at<Hdr>(workspace, L.hdr), at<double>(workspace, L.part_fwd), (int)B, (int)L.Kp, (int)row0, (int)(row0 + nrows), tr0,
            ntr, ntc, klp, full ? kltp : nullptr, (int)L.Bp, sumsq, S, 0, (int)kg.base);
        if (store_kl && !full) {
            // pass 2 (row blocks): tiles (all rows) x (own columns), stored transposed = KL_ji for the own i
            const int tcb = (int)(kg.base / T), ntcb = (int)(kg.rows_pad / T), ntra = (int)(L.Bp / T);
            const int64_t nt2 = (int64_t)ntra * ntcb;
            const int grid2 = (int)(nt2 < kMaxFwdCtas ? nt2 : kMaxFwdCtas);
            mpd_fwd_kernel<TM><<<grid2, kNT, 0, as_stream(stream)>>>(
                at<float>(workspace, L.xp), at<float>(workspace, L.yp), at<float>(workspace, L.c), at<float>(workspace, L.l),
                at<Hdr>(workspace, L.hdr), at<double>(workspace, L.part_fwd), (int)B, (int)L.Kp, 0, 0, 0, ntra, ntcb, nullptr, kltp,
                (int)L.Bp, nullptr, nullptr, tcb, (int)kg.base);
        }
    };
    if (g_mpd_tc && L.tp_mat > 0) {   // padded batch >= 2048, K = 2D <= 512 (<= 192 accumulating MMAs)
        // tensor-core tiles (3xTF32 tcgen05.mma): one persistent CTA per SM
        static const cudaError_t attr = cudaFuncSetAttribute(mpd_fwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                             FwdTcSmem::kBytes);
        if (attr != cudaSuccess) return (int)attr;
        const int T = kTcT;
        const int tr0 = (int)(row0 / T);
        const int ntr = (int)(ceil_div(row0 + nrows, T) - tr0);
        const int ntc = (int)(L.Bp / T);
        const int64_t ntiles = (int64_t)ntr * ntc;
        const int grid = (int)(ntiles < sm_count() ? ntiles : sm_count());
        // profiling runs (MAE_MPD_PROF=1): per-role cycle counters of CTA 0 in the (unused) second half of the per-CTA partials
        static const bool want_prof = getenv("MAE_MPD_PROF") != nullptr;
        long long* prof = nullptr;
        if (want_prof) {
            cudaGetSymbolAddress(reinterpret_cast<void**>(&prof), g_mpd_prof);
            cudaMemsetAsync(prof, 0, sizeof(long long) * 32, as_stream(stream));
        }
        launch_pdl(mpd_fwd_tc_kernel, dim3((unsigned)grid), dim3(kFwdTcThreads), (size_t)FwdTcSmem::kBytes, as_stream(stream),
            at<float>(workspace, L.tp), L.tp_mat, at<float>(workspace, L.c), at<float>(workspace, L.l),
            at<Hdr>(workspace, L.hdr), at<double>(workspace, L.part_fwd), (int)B, (int)L.Kp, (int)row0, (int)(row0 + nrows), tr0,
            ntr, ntc, klp, (full && kg.tc_split == 0) ? kltp : nullptr, (int)L.Bp, sumsq, S, 0, (int)kg.base, prof);
        // (full batch with the tensor-core backward: no KL^T, mpd_bwd_tc_kernel reads KL_ji from the columns of KL; K = 2D not a
        //  multiple of 128: the FFMA stored-KL backward follows and wants both matrices)
        if (store_kl && !full) {
            const int tcb = (int)(kg.base / T), ntcb = (int)(kg.rows_pad / T), ntra = (int)(L.Bp / T);
            const int64_t nt2 = (int64_t)ntra * ntcb;
            const int grid2 = (int)(nt2 < sm_count() ? nt2 : sm_count());
            mpd_fwd_tc_kernel<<<grid2, kFwdTcThreads, FwdTcSmem::kBytes, as_stream(stream)>>>(
                at<float>(workspace, L.tp), L.tp_mat, at<float>(workspace, L.c), at<float>(workspace, L.l),
                at<Hdr>(workspace, L.hdr), at<double>(workspace, L.part_fwd), (int)B, (int)L.Kp, 0, 0, 0, ntra, ntcb, nullptr, kltp,
                (int)L.Bp, nullptr, nullptr, tcb, (int)kg.base, nullptr);
        }
        return launch_status();
    }
    const int tm_env = []() { const char* e = getenv("MAE_MPD_FWD_TM"); return e ? atoi(e) : 0; }();   // tuning only
    const int64_t small_limit = tm_env == 2 ? 1024 : (tm_env == 4 ? 256 : kSmallTileRows);
    if (L.Bp <= small_limit) launch(std::integral_constant<int, 2>{});   // small batches: 32 x 32 tiles, more CTAs, less latency
    else if (L.Bp < kBigTileRows) launch(std::integral_constant<int, 4>{});   // 64 x 64
    else launch(std::integral_constant<int, 8>{});                    // 128 x 128, 8 x 8 micro-tiles: FFMA-bound regime
    return launch_status();
}

extern "C" int mae_mpd_finalize(const double* sumsq, int64_t B, float* S, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(sumsq && S, MAE_ERR_NULL);
    MAE_REQUIRE(B >= 2, MAE_ERR_SHAPE);
    mpd_finalize_kernel<<<1, 1, 0, as_stream(stream)>>>(sumsq, (double)B, S);
    return launch_status();
}

extern "C" int mae_mpd_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride, void* workspace,
                           int64_t workspace_bytes, int64_t B, int64_t D, int64_t row0, int64_t nrows, const float* g_m,
                           const float* g_S, const float* S, const float* gkl, float* dmu, float* dlogvar, int accumulate,
                           int path, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(mu && logvar && dmu && dlogvar, MAE_ERR_NULL);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    MAE_REQUIRE(g_S == nullptr || S != nullptr, MAE_ERR_NULL);
    Layout L;
    int rc = check_common(B, D, workspace, workspace_bytes, L);
    if (rc != MAE_OK) return rc;
    MAE_REQUIRE(row0 >= 0 && nrows >= 1 && row0 + nrows <= B, MAE_ERR_SHAPE);
    MAE_REQUIRE(path >= 0 && path <= 3, MAE_ERR_SHAPE);
    if (path == 0) path = (B <= MAE_MPD_SMALL_B) ? 1 : 2;
    const KlGeom kg = kl_geom(L, B, path == 1 ? 0 : row0, path == 1 ? B : nrows);   // path 1 always holds the full matrices
    if (path == 1 || path == 3) MAE_REQUIRE(workspace_bytes >= kg.total, MAE_ERR_WORKSPACE);
    if (path == 1) {
        MAE_REQUIRE(B <= MAE_MPD_SMALL_B, MAE_ERR_UNSUPPORTED);
        const int64_t items = nrows * ceil_div(D, 32);
        auto launch = [&](auto ps_tag) {
            constexpr int PS = decltype(ps_tag)::value;
            const int64_t blocks = ceil_div(items, (kNT / 32) / PS);
            launch_pdl(mpd_bwd_small_kernel<PS>, dim3((unsigned)blocks), dim3(kNT), 0, as_stream(stream),
                mu, mu_stride, logvar, logvar_stride, at<float4>(workspace, L.vp), at<float>(workspace, L.kl),
                at<float>(workspace, L.klt), (int)L.Bp, at<double>(workspace, L.colstats), at<double>(workspace, L.pivots),
                at<Hdr>(workspace, L.hdr), g_m, g_S, S, gkl, (int)B, (int)D, (int)row0, (int)nrows, dmu, dlogvar, accumulate);
        };
        MAE_REQUIRE(ceil_div(items, 1) < (1ll << 31), MAE_ERR_SHAPE);
        // partners per row = B: keep every warp's walk at <= ~4 blocks of 32 partners while the grid is small
        if (B > 256 && items <= 4096) launch(std::integral_constant<int, 8>{});
        else if (B > 96 && items <= 8192) launch(std::integral_constant<int, 4>{});
        else launch(std::integral_constant<int, 1>{});
        return launch_status();
    }
    if (path == 3 && g_mpd_tc && kg.tc_split > 0) {
        // tensor-core partner contractions (3xTF32 tcgen05.mma) + the closed-form finish
        static const cudaError_t attr = cudaFuncSetAttribute(mpd_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                             BwdTcSmem::kBytes);
        if (attr != cudaSuccess) return (int)attr;
        const bool full_rows = row0 == 0 && nrows == B;     // then the forward kept KL only: KL_ji is read from its columns
        const int tr0t = (int)(row0 / kTcT);
        const int ntrt = (int)(ceil_div(row0 + nrows, kTcT) - tr0t);
        const int ntc = (int)(L.Bp / kTcT);
        const int nsplit = (int)kg.tc_split;
        const int tps = (int)ceil_div(ntc, nsplit);
        float* gp = at<float>(workspace, kg.gp_off);
        float* rsp = at<float>(workspace, kg.rs_off);
        dim3 grid((unsigned)(L.Kp / kTcT), (unsigned)ntrt, (unsigned)nsplit);
        static const bool want_bprof = getenv("MAE_MPD_PROF") != nullptr;
        long long* bprof = nullptr;
        if (want_bprof) {
            cudaGetSymbolAddress(reinterpret_cast<void**>(&bprof), g_mpd_prof);
            bprof += 32;
        }
        launch_pdl(mpd_bwd_tc_kernel, grid, dim3(kBwdTcThreads), (size_t)BwdTcSmem::kBytes, as_stream(stream),
            at<float>(workspace, L.tp), L.tp_mat, at<Hdr>(workspace, L.hdr), (int)B, (int)L.Kp, (int)row0,
            (int)(row0 + nrows), tr0t, ntc, tps, at<float>(workspace, kg.kl_off) - kg.base * L.Bp,
            full_rows ? nullptr : at<float>(workspace, kg.klt_off) - kg.base * L.Bp, (int)L.Bp, gp, rsp, (int)kg.rows_pad, (int)kg.base,
            tc_group_chunks(), bprof);
        const dim3 fgrid((unsigned)ceil_div(nrows, kFinRows), (unsigned)ceil_div(D, kFinDims));
        mpd_bwd_tc_finish_kernel<<<fgrid, kFinRows * kFinDims, 0, as_stream(stream)>>>(
            mu, mu_stride, logvar, logvar_stride, at<float4>(workspace, L.vp), at<double>(workspace, L.colstats),
            at<double>(workspace, L.pivots), g_m, g_S, S, gkl, (int)B, (int)D, (int)L.Kp, (int)row0, (int)nrows, nsplit, gp, rsp,
            (int)kg.rows_pad, (int)kg.base, dmu, dlogvar, accumulate);
        return launch_status();
    }
    const int tr0 = (int)(row0 / kT);
    const int ntr = (int)(ceil_div(row0 + nrows, kT) - tr0);
    auto launch = [&](auto stored_tag, auto oc_tag, auto minb_tag) -> int {
        constexpr bool STORED = decltype(stored_tag)::value;
        constexpr int OC = decltype(oc_tag)::value;
        constexpr int MINB = decltype(minb_tag)::value;
        const int nchunks = (int)ceil_div(L.Kp, OC);
        if (nchunks > 65535) return MAE_ERR_SHAPE;
        dim3 grid((unsigned)ntr, (unsigned)nchunks);
        const size_t smem = STORED ? bwd_stored_smem<OC>() : sizeof(BwdSmem<OC>);
        cudaError_t e = cudaFuncSetAttribute(mpd_bwd_tile_kernel<STORED, OC, MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return (int)e;
        mpd_bwd_tile_kernel<STORED, OC, MINB><<<grid, kNT, smem, as_stream(stream)>>>(
            mu, mu_stride, logvar, logvar_stride, at<float>(workspace, L.xp), at<float>(workspace, L.yp),
            at<float>(workspace, L.c), at<float>(workspace, L.l), at<float4>(workspace, L.vp),
            at<double>(workspace, L.colstats), at<double>(workspace, L.pivots), at<Hdr>(workspace, L.hdr), g_m, g_S, S, gkl,
            (int)B, (int)D, (int)L.Kp, (int)row0, (int)(row0 + nrows), tr0, (int)L.nblk, dmu, dlogvar, accumulate,
            STORED ? at<float>(workspace, kg.kl_off) - kg.base * L.Bp : nullptr,
            STORED ? at<float>(workspace, kg.klt_off) - kg.base * L.Bp : nullptr, (int)L.Bp);
        return MAE_OK;
    };
    // 64-column chunks double the CTA count when 128-column chunks would not even fill the SMs once (stored path: the
    // KL tiles are then streamed twice, which costs far less than idle SMs); the recompute path keeps 128
    const bool narrow = path == 3 && (int64_t)ntr * ceil_div(L.Kp, 128) < 148 && L.Kp > 64;
    using I1 = std::integral_constant<int, 1>;
    using I2 = std::integral_constant<int, 2>;
    using OC128 = std::integral_constant<int, 128>;
    const int variant = []() { const char* e = getenv("MAE_MPD_BWD_VARIANT"); return e ? atoi(e) : 0; }();   // tuning only
    if (path == 3) {
        // grids that do not fill the SMs once (B = 8192, D = 64: 128 CTAs): one CTA per SM with the full register budget
        // (160 registers, deeper load/FFMA overlap) measured 1.12 ms vs 1.21 ms for 64-column chunks x 2 CTAs/SM and
        // 1.40 ms for 128 registers; larger grids keep 2 CTAs per SM
        if (narrow && variant == 2) rc = launch(std::true_type{}, std::integral_constant<int, 64>{}, I2{});
        else if (narrow) rc = launch(std::true_type{}, OC128{}, I1{});
        else rc = launch(std::true_type{}, OC128{}, I2{});
    } else {
        rc = launch(std::false_type{}, OC128{}, I1{});
    }
    if (rc != MAE_OK) return rc;
    return launch_status();
}

extern "C" int mae_mpd_loss_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride, void* workspace,
                                int64_t workspace_bytes, int64_t B, int64_t D, int64_t row0, int64_t nrows, float eta, float gamma,
                                float free_bits, float kl_weight, const float* m_d, const float* S, const float* kl_all,
                                float* reg, float* dmu, float* dlogvar, int path, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(m_d && S, MAE_ERR_NULL);
    Layout L;
    int rc = check_common(B, D, workspace, workspace_bytes, L);
    if (rc != MAE_OK) return rc;
    float* g_m = at<float>(workspace, L.lossg);
    float* g_S = g_m + D;
    float* gkl = (kl_all != nullptr && kl_weight > 0.f) ? g_m + D + 8 : nullptr;
    launch_pdl(mpd_loss_grads_kernel, dim3(1), dim3(kNT), 0, as_stream(stream), m_d, S, kl_all, (int)B, (int)D, eta, gamma, free_bits,
               kl_weight, g_m, g_S, gkl, reg);
    rc = launch_status();
    if (rc != MAE_OK || dmu == nullptr) return rc;   // dmu NULL: only reg is wanted (evaluation)
    return mae_mpd_bwd(mu, mu_stride, logvar, logvar_stride, workspace, workspace_bytes, B, D, row0, nrows,
                       eta > 0.f ? g_m : nullptr, gamma > 0.f ? g_S : nullptr, S, gkl, dmu, dlogvar, 0, path, stream);
}
