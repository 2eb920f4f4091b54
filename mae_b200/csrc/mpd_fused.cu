// A2 + A3 + A3' forward AND backward in ONE launch for training-sized batches (B <= 128, D <= 64: the reference's
// MNIST / OMNIGLOT / CIFAR heads):
//     reg(mu, logvar) = clamp(kl_weight * sum_b KL_b, min = free_bits) - eta * sum_d logsigmoid(m_d) + gamma * S
// (mae.py:132-139 with gaussian_encoder.py:162-188, :219).  The kernel returns KL[B], m_d[D], S and d reg / d(mu, logvar),
// so the regulariser leaves the critical path of a training step: nothing of it remains in the backward pass except an
// optional rescale by d loss.  The general three-kernel path (mpd.cu) spends ~21 us of dependent launches here.
//
// One thread-block CLUSTER of up to 16 CTAs; CTA c owns R = ceil(B/16) rows.  Every CTA evaluates the O(B*D) pre-pass for
// ALL rows redundantly (4-8k elements: cheaper than exchanging them - distributed shared memory moves only ~20 B/clk per
// SM, measured: an all-to-all of the per-element results cost 9 us), keeps V = {mc, a, b, r} of every (row, dim) in its
// shared memory, and then only works on its own rows:
//   phase 1  pivots, per-element quantities, column statistics (fp64), row sums E_i, F_i, analytic KL_i, closed-form m_d
//   phase 2  KL_ij and KL_ji for own i, all j:  2 KL_ij = sum_d [a_i b_j + (mc_i - mc_j)^2 r_j] + E_i + F_j
//            (pivoted form of mpd.cu without the K = 2D panels: both orientations share the squared difference)
//   exchange the CTAs' partial sums of (KL_ij - m)^2 meet through distributed shared memory (fixed order) -> S
//   phase 3  gradient of own rows: complete (as the i and as the j argument), same formulas as mpd_bwd_small_kernel
// Deterministic: fixed reduction orders everywhere, no atomics.
#include <cooperative_groups.h>

#include "common.cuh"
#include "mpd_math.cuh"

namespace cg = cooperative_groups;

namespace mae {
namespace {

constexpr int kFT = 512;           // threads per CTA
constexpr int kFD = 64;            // latent dims supported (one thread column per dim)
constexpr int kFG = kFT / kFD;     // 8 row groups in the pre-pass
constexpr int kFStats = 8;         // SA SB SAB S1 S2 R1 R2 S(a+b), as in mpd.cu
constexpr int kFPivotRows = 32;
constexpr int kFMaxB = 128;
constexpr int kFMaxCluster = 16;    // CTAs per cluster (non-portable size, opt-in attribute)

#ifdef MAE_DEBUG_TIMING
__device__ unsigned long long g_fused_dbg[16];
__device__ __forceinline__ unsigned long long fused_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define FUSED_STAMP(slot)                                                          \
    do {                                                                           \
        __syncthreads();                                                           \
        if (threadIdx.x == 0 && blockIdx.x == 0) g_fused_dbg[slot] = fused_now_ns(); \
    } while (0)
#else
#define FUSED_STAMP(slot) do {} while (0)
#endif

// Block-wide sums of three values at once (fixed order); every thread returns the totals.
__device__ __forceinline__ void block_sum3(double& a, double& b, double& c, double (*scratch)[kFT / 32]) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    a = warp_sum(a);
    b = warp_sum(b);
    c = warp_sum(c);
    __syncthreads();
    if (lane == 0) {
        scratch[0][w] = a;
        scratch[1][w] = b;
        scratch[2][w] = c;
    }
    __syncthreads();
    constexpr int NW = kFT / 32;
    a = warp_sum(lane < NW ? scratch[0][lane] : 0.0);
    b = warp_sum(lane < NW ? scratch[1][lane] : 0.0);
    c = warp_sum(lane < NW ? scratch[2][lane] : 0.0);
}

// Lane sums of N per-lane values at once: butterfly that halves the number of values a lane carries at every step
// (N-1 shuffles) followed by plain xor steps over the remaining lane bits, instead of 5 N shuffles.  Afterwards every lane
// holds the complete sum of value number  (lane >> (5 - log2 N)) & (N - 1).
template <int N>
__device__ __forceinline__ float multi_warp_sum(float (&x)[N], int lane) {
    int off = 16;
#pragma unroll
    for (int n = N; n > 1; n >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int k = 0; k < n / 2; ++k) {
            const float send = upper ? x[k] : x[k + n / 2];
            const float keep = upper ? x[k + n / 2] : x[k];
            x[k] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    float v = x[0];
    for (; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    return v;
}

__device__ __forceinline__ float logsigmoid_f(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

template <int NJ>
struct FusedSmem {
    double cs[kFStats][kFD];
    double piv[4][kFD];
    double scratch[3][kFT / 32];
    double xs[kFMaxCluster];   // every CTA's sum of squares, pushed by its owner
    float pivpart[2][kFG][kFD];
    float rowsum[3][2][NJ];   // E, F, analytic KL: one partial per 32-dim half
    float E[NJ], F[NJ], K[NJ];
    float w1[NJ / 16][NJ], w2[NJ / 16][NJ];   // (KL_ij - m), (KL_ji - m) of the own rows, 0 where excluded
    float gm[kFD];
    float gk;
};

// Shared-memory carve-up behind FusedSmem (all offsets multiples of 16 bytes):
//   V    float4 [D][Bs]             {mc, a, b, r} of every (dim, row)
//   U    union: part double [8][8][64] (column statistics per row group) | red float2 [8][RM][NJ] | gpart float [8][RM][2][64]
template <int NJ>
__host__ __device__ constexpr size_t fused_off_v() { return ((sizeof(FusedSmem<NJ>) + 15) / 16) * 16; }
// union area: part 32 KB, red 8 * RM * NJ * 8 B = 16 KB (NJ = 64) / 64 KB (NJ = 128), gpart 16 / 32 KB
static inline size_t fused_union_bytes(int nj) { return nj == 64 ? 32 * 1024 : 64 * 1024; }

template <int NJ>
__global__ void __launch_bounds__(kFT, 1)
kl_mpd_loss_fused_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                         int B, int D, int Bs, int R, float eta, float gamma, float free_bits, float kl_weight,
                         float* __restrict__ m_d_out, float* __restrict__ S_out, float* __restrict__ kl_out,
                         float* __restrict__ reg_out, float* __restrict__ dmu, float* __restrict__ dlv) {
    constexpr int RM = NJ / 16;   // own rows per CTA at most
    extern __shared__ __align__(16) unsigned char fused_smem[];
    FusedSmem<NJ>& sm = *reinterpret_cast<FusedSmem<NJ>*>(fused_smem);
    float4* V = reinterpret_cast<float4*>(fused_smem + fused_off_v<NJ>());
    unsigned char* U = reinterpret_cast<unsigned char*>(V + (size_t)D * Bs);
    double* part = reinterpret_cast<double*>(U);
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int d = tid & (kFD - 1), g = tid >> 6, half = w & 1;
    const bool dok = d < D;
    const int i0 = crank * R;
    const double nb = (double)B;

    FUSED_STAMP(0);
    // ------------------------------------------------------------------ phase 1: every row, every dim, redundantly
    // (exchanging the per-element results instead costs more: distributed shared memory moves ~20 B/clk per SM)
    constexpr int RPT = NJ / kFG;   // rows per thread
    float mur[RPT], lvr[RPT];
#pragma unroll
    for (int q = 0; q < RPT; ++q) {
        const int i = g + q * kFG;
        const bool ok = dok && i < B;
        mur[q] = ok ? mu[(int64_t)i * mu_stride + d] : 0.f;
        lvr[q] = ok ? lv[(int64_t)i * lv_stride + d] : 0.f;
    }
    const int P = B < kFPivotRows ? B : kFPivotRows;
    {
        float pm = 0.f, pl = 0.f;
#pragma unroll
        for (int q = 0; q < RPT; ++q)
            if (g + q * kFG < P) {
                pm += mur[q];
                pl += lvr[q];
            }
        sm.pivpart[0][g][d] = pm;
        sm.pivpart[1][g][d] = pl;
    }
    __syncthreads();
    FUSED_STAMP(1);
    float mu_bar = 0.f, l_bar = 0.f;
#pragma unroll
    for (int gg = 0; gg < kFG; ++gg) {
        mu_bar += sm.pivpart[0][gg][d];
        l_bar += sm.pivpart[1][gg][d];
    }
    mu_bar /= (float)P;
    l_bar /= (float)P;
    const float v_bar = expf(l_bar);
    const float ivb = 1.0f / v_bar;
    if (g == 0) {
        sm.piv[0][d] = mu_bar;
        sm.piv[1][d] = l_bar;
        sm.piv[2][d] = v_bar;
        sm.piv[3][d] = 1.0 / (double)v_bar;
    }
    {
        // per-thread partial column statistics over <= 16 rows in fp32 (every term is small and well-conditioned in the
        // pivoted variables), widened to fp64 for the sums across threads and CTAs
        float acc[kFStats];
#pragma unroll
        for (int q = 0; q < kFStats; ++q) acc[q] = 0.f;
        float es[RPT], fs[RPT], ks[RPT];
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int i = g + q * kFG;
            es[q] = fs[q] = ks[q] = 0.f;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dok && i < B) {
                const Elem e = make_elem(mur[q], lvr[q], mu_bar, l_bar, v_bar);
                const float sq = e.mc * e.mc;
                v4 = make_float4(e.mc, e.a, e.b, e.r);
                es[q] = e.ema;   // a - alpha >= 0
                fs[q] = e.bpa;   // b + alpha
                ks[q] = fmaf(mur[q], mur[q], expm1_minus_x(lvr[q]));   // gaussian_encoder.py:219, cancellation-free
                acc[0] += e.a;
                acc[1] += e.b;
                acc[2] = fmaf(e.a, e.b, acc[2]);
                acc[3] += e.mc;
                acc[4] += sq;
                acc[5] = fmaf(e.r, e.mc, acc[5]);
                acc[6] = fmaf(e.r, sq, acc[6]);
                acc[7] += e.apb;
            }
            if (dok && i < Bs) V[d * Bs + i] = v4;   // rows in [B, Bs) are zero padding
        }
        {
            constexpr int SH = RPT == 8 ? 2 : 1;   // lanes sharing one row's total after the butterfly: 4 or 2
            const float e_ = multi_warp_sum<RPT>(es, lane);
            const float f_ = multi_warp_sum<RPT>(fs, lane);
            const float k_ = multi_warp_sum<RPT>(ks, lane);
            if ((lane & ((1 << SH) - 1)) == 0) {
                const int i = g + ((lane >> SH) & (RPT - 1)) * kFG;
                sm.rowsum[0][half][i] = e_;
                sm.rowsum[1][half][i] = f_;
                sm.rowsum[2][half][i] = k_;
            }
        }
#pragma unroll
        for (int q = 0; q < kFStats; ++q) part[(g * kFStats + q) * kFD + d] = (double)acc[q];
    }
    __syncthreads();
    FUSED_STAMP(2);
    {
        const int q = g;   // statistic (kFT / kFD == kFStats)
        double t = 0.0;
#pragma unroll
        for (int gg = 0; gg < kFG; ++gg) t += part[(gg * kFStats + q) * kFD + d];
        sm.cs[q][d] = t;
        for (int i = tid; i < NJ; i += kFT) {
            sm.E[i] = sm.rowsum[0][0][i] + sm.rowsum[0][1][i];
            sm.F[i] = sm.rowsum[1][0][i] + sm.rowsum[1][1][i];
            sm.K[i] = 0.5f * (sm.rowsum[2][0][i] + sm.rowsum[2][1][i]);
        }
    }
    __syncthreads();
    // closed-form m_d (SURVEY.md §8a "A3-alg" in the pivoted variables; same expression as mpd_prepare_kernel)
    double mdv = 0.0, lsig = 0.0;
    if (tid < D) {
        const double SA = sm.cs[0][tid], SB = sm.cs[1][tid], SAB = sm.cs[2][tid], S1 = sm.cs[3][tid], S2 = sm.cs[4][tid],
                     R1 = sm.cs[5][tid], R2 = sm.cs[6][tid], SAPB = sm.cs[7][tid];
        const double r0 = (nb + SB) * sm.piv[3][tid];
        const double tn = SA * SB - SAB + (nb - 1.0) * SAPB + (S2 * r0 - 2.0 * S1 * R1 + nb * R2);
        mdv = 0.5 / (nb * (nb - 1.0)) * tn;
        const float v = (float)mdv;
        const float e = expf(-fabsf(v));
        const float sig_neg = v >= 0.f ? e / (1.0f + e) : 1.0f / (1.0f + e);
        sm.gm[tid] = -eta * sig_neg;   // d/dm [-eta * logsigmoid(m)]
        lsig = (double)logsigmoid_f(v);
        if (crank == 0) m_d_out[tid] = v;
    }
    double klsum = (tid < B) ? (double)sm.K[tid] : 0.0;
    block_sum3(mdv, klsum, lsig, sm.scratch);
    const float m = (float)mdv;
    const float mean_kl = (float)(klsum * (double)kl_weight);
    if (tid == 0) sm.gk = (kl_weight > 0.f && mean_kl >= free_bits) ? kl_weight : 0.f;
    if (kl_out != nullptr && crank == 0 && tid < B) kl_out[tid] = sm.K[tid];

    FUSED_STAMP(3);
    // ------------------------------------------------------------------ phase 2: KL_ij, KL_ji of the own rows
    {
        // thread = (partner column j0 [+64], group of 8 latent dims): with two partner columns per thread (B > 64) every
        // broadcast load of an own row's V serves two pairs and the two accumulation chains interleave
        constexpr int NDG = kFT / 64;     // 8 dim groups
        constexpr int DPG = kFD / NDG;    // 8 dims per group
        constexpr int JPT = NJ / 64;      // partner columns per thread
        float2* red = reinterpret_cast<float2*>(U);   // [NDG][RM][NJ]
        const int j0 = tid & 63, dg = tid >> 6;
        float a1[JPT][RM], a2[JPT][RM];
#pragma unroll
        for (int u = 0; u < JPT; ++u)
#pragma unroll
            for (int r = 0; r < RM; ++r) a1[u][r] = a2[u][r] = 0.f;
#pragma unroll 2
        for (int dd = 0; dd < DPG; ++dd) {
            const int dc = dg * DPG + dd;
            if (dc >= D) break;
            const float4* vrow = V + dc * Bs;
            float4 vj[JPT];
#pragma unroll
            for (int u = 0; u < JPT; ++u) {
                const int j = j0 + 64 * u;
                vj[u] = j < B ? vrow[j] : make_float4(0.f, 0.f, 0.f, 0.f);
            }
#pragma unroll
            for (int r = 0; r < RM; ++r) {
                if (r < R) {
                    const float4 vi = vrow[i0 + r];   // warp-wide broadcast
#pragma unroll
                    for (int u = 0; u < JPT; ++u) {
                        const float dl = vi.x - vj[u].x;
                        const float q2 = dl * dl;
                        a1[u][r] = fmaf(vi.y, vj[u].z, fmaf(q2, vj[u].w, a1[u][r]));
                        a2[u][r] = fmaf(vj[u].y, vi.z, fmaf(q2, vi.w, a2[u][r]));
                    }
                }
            }
        }
#pragma unroll
        for (int u = 0; u < JPT; ++u)
#pragma unroll
            for (int r = 0; r < RM; ++r) red[(dg * RM + r) * NJ + j0 + 64 * u] = make_float2(a1[u][r], a2[u][r]);
        __syncthreads();
        float psum = 0.f;
        for (int item = tid; item < RM * NJ; item += kFT) {
            const int r = item / NJ, jj = item % NJ;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < NDG; ++q) {
                const float2 t = red[(q * RM + r) * NJ + jj];
                s1 += t.x;
                s2 += t.y;
            }
            const int i = i0 + r;
            const bool ok = r < R && i < B && jj < B && jj != i;
            float kij = 0.f, kji = 0.f;
            if (ok) {
                kij = 0.5f * (s1 + sm.E[i] + sm.F[jj]) - m;
                kji = 0.5f * (s2 + sm.E[jj] + sm.F[i]) - m;
                psum = fmaf(kij, kij, psum);
            }
            sm.w1[r][jj] = kij;
            sm.w2[r][jj] = kji;
        }
        double mine = (double)psum, z0 = 0.0, z1 = 0.0;
        block_sum3(mine, z0, z1, sm.scratch);
        if (tid < C) *cluster.map_shared_rank(&sm.xs[crank], tid) = mine;   // thread c tells CTA c
    }
    FUSED_STAMP(4);
    // ------------------------------------------------------------------ exchange: S needs every CTA's rows
    cluster.sync();   // last remote access of the kernel: afterwards every CTA only touches its own shared memory
    double total = 0.0;
    for (int c = 0; c < C; ++c) total += sm.xs[c];
    const float S = (float)sqrt(total / (nb * (nb - 1.0) - 1.0));
    if (crank == 0 && tid == 0) {
        S_out[0] = S;
        if (reg_out != nullptr) {
            float reg = kl_weight > 0.f ? fmaxf(mean_kl, free_bits) : 0.f;
            if (eta > 0.f) reg -= eta * (float)lsig;
            if (gamma > 0.f) reg += gamma * S;
            reg_out[0] = reg;
        }
    }
    FUSED_STAMP(5);
    if (dmu == nullptr) return;

    // ------------------------------------------------------------------ phase 3: gradient of the own rows
    float beta = 0.f;   // d(gamma * S) / d KL_ij = beta * (KL_ij - m)     (SURVEY.md §8a "A3-bwd")
    if (gamma > 0.f && S > 0.f) beta = (float)((double)gamma / ((nb * nb - nb - 1.0) * (double)S));
    float* gpart = reinterpret_cast<float*>(U);   // [8 partner groups][RM][2][64]
    const float4* vrow = V + d * Bs;
    if (beta != 0.f) {
        // thread = (latent dim, partner group): all own rows against a 1/8 slice of the partners, so that every partner's
        // V is read once per dim and serves RM rows
        constexpr int JP = NJ / kFG;   // partners per group (multiple of 4)
        const int jb0 = g * JP;
        const int Bq = (B + 3) & ~3;
        float4 pi[RM];
        float ti[RM], oi[RM], gmu[RM], glv[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            pi[r] = (dok && r < R) ? vrow[i0 + r] : make_float4(0.f, 0.f, 0.f, 0.f);
            ti[r] = pi[r].y + (2.0f * pi[r].z + pi[r].z * pi[r].z) * (1.0f + pi[r].y);   // r_i^2 v_i v_bar - 1
            oi[r] = (1.0f + ti[r]) * ivb;
            gmu[r] = glv[r] = 0.f;
        }
        if (dok) {
            for (int jb = jb0; jb < jb0 + JP && jb < Bq; jb += 4) {
                float4 pj[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) pj[u] = vrow[jb + u];
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    if (r < R) {
                        const float4 wa = *reinterpret_cast<const float4*>(&sm.w1[r][jb]);
                        const float4 wb = *reinterpret_cast<const float4*>(&sm.w2[r][jb]);
                        const float w1v[4] = {wa.x, wa.y, wa.z, wa.w};
                        const float w2v[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            // a_i b_j + a_i + b_j = (a_i + 1) b_j + a_i ;  a_j t_i + a_j + t_i + dl^2 o_i = (t_i + 1) a_j + (t_i + dl^2 o_i)
                            const float dl = pi[r].x - pj[u].x;
                            gmu[r] = fmaf(dl, fmaf(w1v[u], pj[u].w, w2v[u] * pi[r].w), gmu[r]);
                            glv[r] = fmaf(w1v[u], fmaf(pi[r].y + 1.0f, pj[u].z, pi[r].y), glv[r]);
                            glv[r] = fmaf(-w2v[u], fmaf(ti[r] + 1.0f, pj[u].y, fmaf(dl * dl, oi[r], ti[r])), glv[r]);
                        }
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) {
            gpart[((g * RM + r) * 2 + 0) * kFD + d] = gmu[r];
            gpart[((g * RM + r) * 2 + 1) * kFD + d] = glv[r];
        }
        __syncthreads();
    }
    // thread = (latent dim, own row): fixed-order sum of the partner groups, closed-form mean part, analytic-KL part
    for (int r = g; r < R; r += kFG) {
        const int i = i0 + r;
        if (!dok || i >= B) continue;
        float gmu = 0.f, glv = 0.f;
        if (beta != 0.f) {
#pragma unroll
            for (int q = 0; q < kFG; ++q) {
                gmu += gpart[((q * RM + r) * 2 + 0) * kFD + d];
                glv += gpart[((q * RM + r) * 2 + 1) * kFD + d];
            }
            gmu *= beta;
            glv *= 0.5f * beta;
        }
        const float4 p = vrow[i];
        const float t_i = p.y + (2.0f * p.z + p.z * p.z) * (1.0f + p.y);
        mean_part(&sm.cs[0][0], &sm.piv[0][0], eta > 0.f ? sm.gm : nullptr, B, kFD, d, make_float4(p.x, p.y, p.z, t_i),
                  mu[(int64_t)i * mu_stride + d], lv[(int64_t)i * lv_stride + d], gmu, glv, sm.gk != 0.f ? &sm.gk : nullptr);
        dmu[(int64_t)i * D + d] = gmu;
        dlv[(int64_t)i * D + d] = glv;
    }
#ifdef MAE_DEBUG_TIMING
    __syncthreads();
    if (tid == 0 && blockIdx.x == 0) g_fused_dbg[6] = fused_now_ns();
#endif
}

struct FusedGeom {
    int nj, cluster, bs, rows;
    size_t smem;
};

bool fused_geom(int64_t B, int64_t D, FusedGeom& g) {
    if (B < 2 || B > kFMaxB || D < 1 || D > kFD) return false;
    g.nj = B <= 64 ? 64 : 128;
    g.rows = (int)ceil_div(B, kFMaxCluster);          // own rows per CTA: <= NJ/16
    g.cluster = (int)ceil_div(B, g.rows);
    int64_t bs = round_up(B, 8);
    if (bs < (int64_t)g.cluster * g.rows) bs = (int64_t)g.cluster * g.rows;
    g.bs = (int)(bs | 1);   // odd row pitch (in float4): conflict-free for lane = row and for lane = dim
    const size_t off_v = g.nj == 64 ? fused_off_v<64>() : fused_off_v<128>();
    g.smem = off_v + (size_t)D * g.bs * sizeof(float4) + fused_union_bytes(g.nj);
    return g.smem <= 227 * 1024;
}

template <int NJ>
int fused_config(const FusedGeom& g, cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, cudaStream_t stream) {
    auto kern = kl_mpd_loss_fused_kernel<NJ>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g.smem);
    if (e != cudaSuccess) return (int)e;
    if (g.cluster > 8) {
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
    }
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3((unsigned)g.cluster);
    cfg.blockDim = dim3(kFT);
    cfg.dynamicSmemBytes = g.smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)g.cluster;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return MAE_OK;
}

}  // namespace
}  // namespace mae

#ifdef MAE_DEBUG_TIMING
extern "C" int mae_debug_fused_stamps(unsigned long long* out16) {
    return (int)cudaMemcpyFromSymbol(out16, mae::g_fused_dbg, 16 * sizeof(unsigned long long));
}
#endif

extern "C" int mae_kl_mpd_loss_fused_supported(int64_t B, int64_t D) {
    using namespace mae;
    FusedGeom g;
    if (!fused_geom(B, D, g)) return 0;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int n = 0;
    cudaError_t e;
    if (g.nj == 64) {
        if (fused_config<64>(g, cfg, attr, nullptr) != MAE_OK) return (void)cudaGetLastError(), 0;
        e = cudaOccupancyMaxActiveClusters(&n, kl_mpd_loss_fused_kernel<64>, &cfg);
    } else {
        if (fused_config<128>(g, cfg, attr, nullptr) != MAE_OK) return (void)cudaGetLastError(), 0;
        e = cudaOccupancyMaxActiveClusters(&n, kl_mpd_loss_fused_kernel<128>, &cfg);
    }
    if (e != cudaSuccess) {
        (void)cudaGetLastError();
        return 0;
    }
    return n >= 1 ? 1 : 0;
}

extern "C" int mae_kl_mpd_loss_fused(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride, int64_t B,
                                     int64_t D, float eta, float gamma, float free_bits, float kl_weight, float* m_d, float* S,
                                     float* kl, float* reg, float* dmu, float* dlogvar, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(mu && logvar && m_d && S, MAE_ERR_NULL);
    MAE_REQUIRE((dmu == nullptr) == (dlogvar == nullptr), MAE_ERR_NULL);
    MAE_REQUIRE(B >= 2 && D >= 1, MAE_ERR_SHAPE);
    MAE_REQUIRE(mu_stride >= D && logvar_stride >= D, MAE_ERR_STRIDE);
    FusedGeom g;
    MAE_REQUIRE(fused_geom(B, D, g), MAE_ERR_UNSUPPORTED);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int bi = (int)B, di = (int)D;
    cudaError_t e;
    if (g.nj == 64) {
        int rc = fused_config<64>(g, cfg, attr, as_stream(stream));
        if (rc != MAE_OK) return rc;
        e = cudaLaunchKernelEx(&cfg, kl_mpd_loss_fused_kernel<64>, mu, mu_stride, logvar, logvar_stride, bi, di, g.bs, g.rows, eta,
                               gamma, free_bits, kl_weight, m_d, S, kl, reg, dmu, dlogvar);
    } else {
        int rc = fused_config<128>(g, cfg, attr, as_stream(stream));
        if (rc != MAE_OK) return rc;
        e = cudaLaunchKernelEx(&cfg, kl_mpd_loss_fused_kernel<128>, mu, mu_stride, logvar, logvar_stride, bi, di, g.bs, g.rows, eta,
                               gamma, free_bits, kl_weight, m_d, S, kl, reg, dmu, dlogvar);
    }
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}
