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

template <int NJ>
struct FusedSmem {
    double cs[kFStats][kFD];
    double piv[4][kFD];
    double scratch[32];
    double xchg;       // this CTA's sum of squares, read by the cluster peers
    double total;      // cluster-wide sum
    union {
        double part[kFG][kFStats][kFD];       // pre-pass: column partials per row group
        float2 red[kFT / NJ][kFR][NJ];        // phase 2: {KL_ij, KL_ji} partials per dim group
    } u;
    float pivpart[2][kFG][kFD];
    float rowsum[3][2][NJ];   // E, F, analytic KL: one partial per 32-dim half
    float E[NJ], F[NJ], K[NJ];
    float w1[kFR][NJ], w2[kFR][NJ];   // (KL_ij - m), (KL_ji - m) of the own rows, 0 where excluded
    float gm[kFD];
    float gk;
};

template <int NJ>
__global__ void __launch_bounds__(kFT, 1)
kl_mpd_loss_fused_kernel(const float* __restrict__ mu, int64_t mu_stride, const float* __restrict__ lv, int64_t lv_stride,
                         int B, int D, int Bs, float eta, float gamma, float free_bits, float kl_weight,
                         float* __restrict__ m_d_out, float* __restrict__ S_out, float* __restrict__ kl_out,
                         float* __restrict__ dmu, float* __restrict__ dlv) {
    extern __shared__ __align__(16) unsigned char fused_smem[];
    FusedSmem<NJ>& sm = *reinterpret_cast<FusedSmem<NJ>*>(fused_smem);
    float4* V = reinterpret_cast<float4*>(fused_smem + ((sizeof(FusedSmem<NJ>) + 15) / 16) * 16);   // [D][Bs]
    cg::cluster_group cluster = cg::this_cluster();
    const int crank = (int)cluster.block_rank(), csize = (int)cluster.num_blocks();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int d = tid & (kFD - 1), g = tid >> 6, half = w & 1;
    const bool dok = d < D;
    const int i0 = crank * kFR;
    const double nb = (double)B;

    // ------------------------------------------------------------------ phase 1: every row, every dim, redundantly
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
#pragma unroll
        for (int q = 0; q < RPT; ++q) {
            const int i = g + q * kFG;
            float es = 0.f, fs = 0.f, ks = 0.f;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dok && i < B) {
                const Elem e = make_elem(mur[q], lvr[q], mu_bar, l_bar, v_bar);
                const float sq = e.mc * e.mc;
                v4 = make_float4(e.mc, e.a, e.b, e.r);
                es = e.ema;   // a - alpha >= 0
                fs = e.bpa;   // b + alpha
                ks = fmaf(mur[q], mur[q], expm1_minus_x(lvr[q], expm1f(lvr[q])));   // gaussian_encoder.py:219
                acc[0] += (double)e.a;
                acc[1] += (double)e.b;
                acc[2] += (double)e.a * (double)e.b;
                acc[3] += (double)e.mc;
                acc[4] += (double)sq;
                acc[5] += (double)e.r * (double)e.mc;
                acc[6] += (double)e.r * (double)sq;
                acc[7] += (double)e.apb;
            }
            if (dok && i < Bs) V[d * Bs + i] = v4;   // rows in [B, Bs) are zero padding
            es = warp_sum(es);
            fs = warp_sum(fs);
            ks = warp_sum(ks);
            if (lane == 0) {
                sm.rowsum[0][half][i] = es;
                sm.rowsum[1][half][i] = fs;
                sm.rowsum[2][half][i] = ks;
            }
        }
#pragma unroll
        for (int q = 0; q < kFStats; ++q) sm.u.part[g][q][d] = acc[q];
    }
    __syncthreads();
    {
        const int q = tid >> 6;   // statistic (kFT / kFD == kFStats)
        double t = 0.0;
#pragma unroll
        for (int gg = 0; gg < kFG; ++gg) t += sm.u.part[gg][q][d];
        sm.cs[q][d] = t;
        for (int i = tid; i < NJ; i += kFT) {
            sm.E[i] = sm.rowsum[0][0][i] + sm.rowsum[0][1][i];
            sm.F[i] = sm.rowsum[1][0][i] + sm.rowsum[1][1][i];
            sm.K[i] = 0.5f * (sm.rowsum[2][0][i] + sm.rowsum[2][1][i]);
        }
    }
    __syncthreads();
    // closed-form m_d (SURVEY.md §8a "A3-alg" in the pivoted variables; same expression as mpd_prepare_kernel)
    double mdv = 0.0;
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
        if (crank == 0) m_d_out[tid] = v;
    }
    const float m = (float)block_sum(mdv, sm.scratch);
    const double klsum = block_sum((tid < B) ? (double)sm.K[tid] : 0.0, sm.scratch);
    if (tid == 0) sm.gk = (kl_weight > 0.f && (float)(klsum * (double)kl_weight) >= free_bits) ? kl_weight : 0.f;
    if (kl_out != nullptr && crank == 0 && tid < B) kl_out[tid] = sm.K[tid];

    // ------------------------------------------------------------------ phase 2: KL_ij, KL_ji of the own rows
    {
        constexpr int NDG = kFT / NJ;     // dim groups
        constexpr int DPG = kFD / NDG;    // dims per group
        const int j = tid % NJ, dg = tid / NJ;
        float a1[kFR], a2[kFR];
#pragma unroll
        for (int r = 0; r < kFR; ++r) a1[r] = a2[r] = 0.f;
        if (j < B) {
#pragma unroll 2
            for (int dd = 0; dd < DPG; ++dd) {
                const int dc = dg * DPG + dd;
                if (dc >= D) break;
                const float4* vrow = V + dc * Bs;
                const float4 vj = vrow[j];
#pragma unroll
                for (int r = 0; r < kFR; ++r) {
                    const float4 vi = vrow[i0 + r];   // warp-wide broadcast
                    const float dl = vi.x - vj.x;
                    const float q2 = dl * dl;
                    a1[r] = fmaf(vi.y, vj.z, fmaf(q2, vj.w, a1[r]));
                    a2[r] = fmaf(vj.y, vi.z, fmaf(q2, vi.w, a2[r]));
                }
            }
        }
#pragma unroll
        for (int r = 0; r < kFR; ++r) sm.u.red[dg][r][j] = make_float2(a1[r], a2[r]);
        __syncthreads();
        float psum = 0.f;
        for (int item = tid; item < kFR * NJ; item += kFT) {
            const int r = item / NJ, jj = item % NJ;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int q = 0; q < NDG; ++q) {
                const float2 t = sm.u.red[q][r][jj];
                s1 += t.x;
                s2 += t.y;
            }
            const int i = i0 + r;
            const bool ok = i < B && jj < B && jj != i;
            const float kij = 0.5f * (s1 + sm.E[i] + sm.F[jj]);
            const float kji = 0.5f * (s2 + sm.E[jj] + sm.F[i]);
            sm.w1[r][jj] = ok ? kij - m : 0.f;
            sm.w2[r][jj] = ok ? kji - m : 0.f;
            if (ok) psum = fmaf(kij - m, kij - m, psum);
        }
        const double mine = block_sum((double)psum, sm.scratch);
        if (tid == 0) sm.xchg = mine;
    }
    // ------------------------------------------------------------------ exchange: S needs every CTA's rows
    cluster.sync();
    if (w == 0) {
        double t = 0.0;
        if (lane < csize) t = *cluster.map_shared_rank(&sm.xchg, lane);
        t = warp_sum(t);   // xor tree: same order in every CTA
        if (lane == 0) sm.total = t;
    }
    cluster.sync();   // also keeps every CTA's shared memory alive until all peers have read it
    const float S = (float)sqrt(sm.total / (nb * (nb - 1.0) - 1.0));
    if (crank == 0 && tid == 0) S_out[0] = S;
    if (dmu == nullptr) return;

    // ------------------------------------------------------------------ phase 3: gradient of the own rows
    const int i = i0 + g;
    if (!dok || i >= B) return;
    float beta = 0.f;   // d(gamma * S) / d KL_ij = beta * (KL_ij - m)     (SURVEY.md §8a "A3-bwd")
    if (gamma > 0.f && S > 0.f) beta = (float)((double)gamma / ((nb * nb - nb - 1.0) * (double)S));
    const float4* vrow = V + d * Bs;
    const float4 pi = vrow[i];
    const float t_i = pi.y + (2.0f * pi.z + pi.z * pi.z) * (1.0f + pi.y);   // r_i^2 v_i v_bar - 1
    float gmu = 0.f, glv = 0.f;
    if (beta != 0.f) {
        const float opt_i = (1.0f + t_i) * ivb;
        for (int jb = 0; jb < B; jb += 4) {
            const float4 wa = *reinterpret_cast<const float4*>(&sm.w1[g][jb]);
            const float4 wb = *reinterpret_cast<const float4*>(&sm.w2[g][jb]);
            const float w1v[4] = {wa.x, wa.y, wa.z, wa.w};
            const float w2v[4] = {wb.x, wb.y, wb.z, wb.w};
            float4 pj[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) pj[u] = vrow[jb + u];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float dl = pi.x - pj[u].x;
                gmu = fmaf(dl, fmaf(w1v[u], pj[u].w, w2v[u] * pi.w), gmu);
                glv += w1v[u] * (fmaf(pi.y, pj[u].z, pi.y) + pj[u].z) -
                       w2v[u] * (fmaf(pj[u].y, t_i, pj[u].y) + t_i + dl * dl * opt_i);
            }
        }
        gmu *= beta;
        glv *= 0.5f * beta;
    }
    const float gk = sm.gk;
    mean_part(&sm.cs[0][0], &sm.piv[0][0], eta > 0.f ? sm.gm : nullptr, B, kFD, d, make_float4(pi.x, pi.y, pi.z, t_i),
              mu[(int64_t)i * mu_stride + d], lv[(int64_t)i * lv_stride + d], gmu, glv, gk != 0.f ? &sm.gk : nullptr);
    dmu[(int64_t)i * D + d] = gmu;
    dlv[(int64_t)i * D + d] = glv;
}

struct FusedGeom {
    int nj, cluster, bs;
    size_t smem;
};

bool fused_geom(int64_t B, int64_t D, FusedGeom& g) {
    if (B < 2 || B > kFMaxB || D < 1 || D > kFD) return false;
    g.nj = B <= 64 ? 64 : 128;
    g.cluster = (int)ceil_div(B, kFR);
    g.bs = (int)(round_up(B, 8) | 1);   // odd row pitch (in float4): conflict-free for lane = row and for lane = dim
    const size_t fixed = g.nj == 64 ? sizeof(FusedSmem<64>) : sizeof(FusedSmem<128>);
    g.smem = ((fixed + 15) / 16) * 16 + (size_t)D * g.bs * sizeof(float4);
    return true;
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
                                     float* kl, float* dmu, float* dlogvar, mae_stream_t stream) {
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
        e = cudaLaunchKernelEx(&cfg, kl_mpd_loss_fused_kernel<64>, mu, mu_stride, logvar, logvar_stride, bi, di, g.bs, eta, gamma,
                               free_bits, kl_weight, m_d, S, kl, dmu, dlogvar);
    } else {
        int rc = fused_config<128>(g, cfg, attr, as_stream(stream));
        if (rc != MAE_OK) return rc;
        e = cudaLaunchKernelEx(&cfg, kl_mpd_loss_fused_kernel<128>, mu, mu_stride, logvar, logvar_stride, bi, di, g.bs, eta, gamma,
                               free_bits, kl_weight, m_d, S, kl, dmu, dlogvar);
    }
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}
