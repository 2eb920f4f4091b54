// A2 + A3 + A3' forward AND backward in ONE launch for training-sized batches (B <= 128, D <= 64: the reference's
// MNIST / OMNIGLOT / CIFAR heads):
//     reg(mu, logvar) = clamp(kl_weight * sum_b KL_b, min = free_bits) - eta * sum_d logsigmoid(m_d) + gamma * S
// (mae.py:132-139 with gaussian_encoder.py:162-188, :219).  The kernel returns KL[B], m_d[D], S and d reg / d(mu, logvar),
// so the regulariser leaves the critical path of a training step: nothing of it remains in the backward pass except an
// optional rescale by d loss.  The general three-kernel path (mpd.cu) spends ~21 us of dependent launches here.
//
// One thread-block CLUSTER of ceil(B/8) CTAs; CTA c owns rows [8c, 8c+8).  Every CTA evaluates the O(B*D) pre-pass for
// ALL rows redundantly (4-8k elements: cheaper than exchanging them), keeps V = {mc, a, b, r} of every (row, dim) in its
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
constexpr int kFR = 8;             // own rows per CTA
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

__device__ __forceinline__ float logsigmoid_f(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

template <int NJ>
struct FusedSmem {
    double cs[kFStats][kFD];
    double piv[4][kFD];
    double scratch[3][kFT / 32];
    double xs[kFMaxCluster];   // every CTA's sum of squares, pushed by its owner
    float pivpart[2][kFG][kFD];
    float rowsum[3][2][NJ];   // E, F, analytic KL: one partial per 32-dim half, pushed by the row's owner
    float E[NJ], F[NJ], K[NJ];
    float w1[NJ / 16][NJ], w2[NJ / 16][NJ];   // (KL_ij - m), (KL_ji - m) of the own rows, 0 where excluded
    float gm[kFD];
    float gk;
};

// Shared-memory carve-up behind FusedSmem (all offsets multiples of 16 bytes):
//   V    float4 [D][Bs]             {mc, a, b, r} of every (dim, row); rows pushed by their owners
//   loc  double [RM][8][64]         this CTA's per-row column statistics before they are summed and pushed
//   U    union: part double [C][8][64] (column statistics of every CTA) | red float2 [512/NJ][RM][NJ] | gpart float [8][RM][2][64]
template <int NJ>
__host__ __device__ constexpr size_t fused_off_v() { return ((sizeof(FusedSmem<NJ>) + 15) / 16) * 16; }
__host__ __device__ inline size_t fused_union_bytes(int nj, int cluster) {
    const size_t part = (size_t)cluster * kFStats * kFD * sizeof(double);
    const size_t other = (size_t)(nj / 16) * 8 * 2 * kFD * sizeof(float) * 2;   // red and gpart: 16 KB (NJ=64) / 32 KB
    return part > other ? part : other;
}

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
    double* loc = reinterpret_cast<double*>(V + (size_t)D * Bs);
    unsigned char* U = reinterpret_cast<unsigned char*>(loc + RM * kFStats * kFD);
    double* part = reinterpret_cast<double*>(U);
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), C = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int d = tid & (kFD - 1), g = tid >> 6, half = w & 1;
    const bool dok = d < D;
    const int i0 = crank * R;
    const double nb = (double)B;

    FUSED_STAMP(0);
    // ------------------------------------------------------------------ phase A: own rows only, results pushed to all
    // what no peer writes: the padding rows of V and the row sums of rows >= B
    for (int q = tid; q < 3 * 2 * NJ; q += kFT) (&sm.rowsum[0][0][0])[q] = 0.f;
    for (int q = tid; q < D * (Bs - B); q += kFT) V[(q / (Bs - B)) * Bs + B + q % (Bs - B)] = make_float4(0.f, 0.f, 0.f, 0.f);
    cluster.barrier_arrive();   // "I have started and initialised": peers may write into this CTA after the matching wait

    const int i_own = i0 + g;
    const bool own_row = g < R && i_own < B;   // warp-uniform
    const bool own_ok = own_row && dok;
    const int P = B < kFPivotRows ? B : kFPivotRows;
    float mu_own = 0.f, lv_own = 0.f;
    {
        constexpr int PQ = kFPivotRows / kFG;   // pivot rows per thread
        float pm[PQ], pl[PQ];
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
            const int i = g + q * kFG;
            const bool ok = dok && i < P;
            pm[q] = ok ? mu[(int64_t)i * mu_stride + d] : 0.f;
            pl[q] = ok ? lv[(int64_t)i * lv_stride + d] : 0.f;
        }
        if (own_ok) {
            mu_own = mu[(int64_t)i_own * mu_stride + d];
            lv_own = lv[(int64_t)i_own * lv_stride + d];
        }
        float sm_ = 0.f, sl_ = 0.f;
#pragma unroll
        for (int q = 0; q < PQ; ++q) {
            sm_ += pm[q];
            sl_ += pl[q];
        }
        sm.pivpart[0][g][d] = sm_;
        sm.pivpart[1][g][d] = sl_;
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
        double acc[kFStats];
#pragma unroll
        for (int q = 0; q < kFStats; ++q) acc[q] = 0.0;
        float es = 0.f, fs = 0.f, ks = 0.f;
        float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
        if (own_ok) {
            const Elem e = make_elem(mu_own, lv_own, mu_bar, l_bar, v_bar);
            const float sq = e.mc * e.mc;
            v4 = make_float4(e.mc, e.a, e.b, e.r);
            es = e.ema;   // a - alpha >= 0
            fs = e.bpa;   // b + alpha
            ks = fmaf(mu_own, mu_own, expm1_minus_x(lv_own, expm1f(lv_own)));   // gaussian_encoder.py:219
            acc[0] = (double)e.a;
            acc[1] = (double)e.b;
            acc[2] = (double)e.a * (double)e.b;
            acc[3] = (double)e.mc;
            acc[4] = (double)sq;
            acc[5] = (double)e.r * (double)e.mc;
            acc[6] = (double)e.r * (double)sq;
            acc[7] = (double)e.apb;
        }
        if (g < RM) {
#pragma unroll
            for (int q = 0; q < kFStats; ++q) loc[(g * kFStats + q) * kFD + d] = acc[q];
        }
        es = warp_sum(es);
        fs = warp_sum(fs);
        ks = warp_sum(ks);
        cluster.barrier_wait();   // every CTA of the cluster runs and has initialised its shared memory
        if (own_ok) {
            for (int c = 0; c < C; ++c) cluster.map_shared_rank(V, c)[d * Bs + i_own] = v4;
        }
        if (own_row && lane < C) {   // lane c tells CTA c
            float* rs = cluster.map_shared_rank(&sm.rowsum[0][0][0], lane);
            rs[(0 * 2 + half) * NJ + i_own] = es;
            rs[(1 * 2 + half) * NJ + i_own] = fs;
            rs[(2 * 2 + half) * NJ + i_own] = ks;
        }
        __syncthreads();
        {
            const int q = g;   // statistic (kFT / kFD == kFStats)
            double t = 0.0;
#pragma unroll
            for (int gg = 0; gg < RM; ++gg) t += loc[(gg * kFStats + q) * kFD + d];
            for (int c = 0; c < C; ++c) cluster.map_shared_rank(part, c)[(crank * kFStats + q) * kFD + d] = t;
        }
    }
    cluster.sync();
    FUSED_STAMP(2);
    // ------------------------------------------------------------------ phase B: column statistics, m_d, scalars
    {
        const int q = g;
        double t = 0.0;
        for (int c = 0; c < C; ++c) t += part[(c * kFStats + q) * kFD + d];
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
        constexpr int NDG = kFT / NJ;     // dim groups
        constexpr int DPG = kFD / NDG;    // dims per group
        float2* red = reinterpret_cast<float2*>(U);   // [NDG][RM][NJ]
        const int j = tid % NJ, dg = tid / NJ;
        float a1[RM], a2[RM];
#pragma unroll
        for (int r = 0; r < RM; ++r) a1[r] = a2[r] = 0.f;
        if (j < B) {
#pragma unroll 2
            for (int dd = 0; dd < DPG; ++dd) {
                const int dc = dg * DPG + dd;
                if (dc >= D) break;
                const float4* vrow = V + dc * Bs;
                const float4 vj = vrow[j];
#pragma unroll
                for (int r = 0; r < RM; ++r) {
                    if (r < R) {
                        const float4 vi = vrow[i0 + r];   // warp-wide broadcast
                        const float dl = vi.x - vj.x;
                        const float q2 = dl * dl;
                        a1[r] = fmaf(vi.y, vj.z, fmaf(q2, vj.w, a1[r]));
                        a2[r] = fmaf(vj.y, vi.z, fmaf(q2, vi.w, a2[r]));
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RM; ++r) red[(dg * RM + r) * NJ + j] = make_float2(a1[r], a2[r]);
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
                            const float dl = pi[r].x - pj[u].x;
                            gmu[r] = fmaf(dl, fmaf(w1v[u], pj[u].w, w2v[u] * pi[r].w), gmu[r]);
                            glv[r] += w1v[u] * (fmaf(pi[r].y, pj[u].z, pi[r].y) + pj[u].z) -
                                      w2v[u] * (fmaf(pj[u].y, ti[r], pj[u].y) + ti[r] + dl * dl * oi[r]);
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
    g.smem = off_v + (size_t)D * g.bs * sizeof(float4) + (size_t)(g.nj / 16) * kFStats * kFD * sizeof(double) +
             fused_union_bytes(g.nj, g.cluster);
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
