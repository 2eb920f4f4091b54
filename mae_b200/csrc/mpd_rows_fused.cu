// Batch-sharded training step, encoder side, in ONE launch per rank: the NVLink exchange of the packed (mu | logvar) rows
// AND the regulariser of MAE.loss over the GLOBAL batch, forward and backward for this rank's rows
//     reg = clamp(kl_weight * sum_b KL_b, min = free_bits) - eta * sum_d logsigmoid(m_d) + gamma * S        (mae.py:132-139)
// (gaussian_encoder.py:162-188, :219; replaces, for the loss head, the DataParallel gather of gaussian_encoder.py:41-42).
// Before, a sharded step spent a dependent chain of five launches on it (push-gather kernel with copy-out -> O(B*D) pre-pass
// -> B x B tiles, evaluated redundantly on every rank -> device-side upstream gradients -> row backward: 45-60 us at global
// batches 256-512 against a 29 us single-GPU step).  Here:
//
//   one thread-block cluster of 16 CTAs x 512 threads per rank; global batch B = world * b_loc <= 512, b_loc <= 64, D <= 64
//   CTA c owns the PARTNER SLICE  j in [c*SL, (c+1)*SL),  SL = ceil(B / 16) <= 32,  and works on (own rows i) x (slice j):
//     phase 0  push the local rows into every peer's landing buffer (16-byte stores over NVLink, CTA p serves peer p), raise
//              the peer's flag, then wait ONLY for the ranks whose rows this CTA reads; the landed rows are consumed in
//              place (no copy-out)
//     phase 1  pivots (first <= 32 global rows), per-element quantities of the own rows and of the slice, row sums E / F,
//              analytic KL, column statistics of the slice (every global row belongs to exactly one slice)
//              -> cluster exchange of the statistics through a global scratch buffer            [cluster.sync 1]
//     phase 2  closed-form m_d, m;  KL_ij and KL_ji for own i, slice j (pivoted pair form of mpd_fused.cu)
//              -> partial sum of (KL_ij - m)^2 -> CTA 0 -> the ranks exchange ONE double each     [cluster.sync 2 + flags]
//     phase 3  partial gradient of every own row over the slice -> global scratch                [cluster.sync 3]
//     phase 4  CTA c finishes own rows 4c..4c+3: fixed-order sum over the 16 slices, closed-form mean part, analytic-KL part
//   Every rank computes only ITS rows of the B x B pair matrix (the old path evaluated all of it on every rank).
//
// Peer protocol (same as p2p_gather.cu): epochs live on the device, landing parities alternate, a rank can start epoch
// e+1 only after every peer has started epoch e.  Spins are bounded (2 s): on a timeout the sticky error flag of the
// buffer is raised and every output of this and all later launches is NaN.  Deterministic: fixed reduction orders.
#include <cooperative_groups.h>

#include "common.cuh"
#include "mpd_math.cuh"

namespace cg = cooperative_groups;

namespace mae {
namespace {

constexpr int kRT = 512;          // threads per CTA
constexpr int kRD = 64;           // latent dims (one thread column per dim)
constexpr int kRG = kRT / kRD;    // 8 row groups
constexpr int kRC = 16;           // CTAs per cluster
constexpr int kROwn = 64;         // own rows per rank at most
constexpr int kRSl = 32;          // partner rows per CTA at most (B <= 512)
constexpr int kRMaxB = kRC * kRSl;
constexpr int kRStats = 8;
constexpr int kRPivotRows = 32;
constexpr int kRMaxWorld = 8;
constexpr int kPo = kROwn + 1;    // float4 pitch of the own-row panel (odd: conflict-free for lane = row and lane = dim)
constexpr int kPp = kRSl + 1;
constexpr unsigned long long kRTimeoutNs = 2000000000ull;

struct RowsHdr {                  // 1024 bytes at the start of the communication buffer
    unsigned int error;
    unsigned int pad[63];
    unsigned int epoch[kRC];      // one launch counter per CTA (all advance in lockstep)
    unsigned int pad2[64 - kRC];
    unsigned int flags[2][kRMaxWorld];    // rows landed:   value = epoch
    unsigned int flags2[2][kRMaxWorld];   // sum landed:    value = epoch
    double sums[2][kRMaxWorld];           // every rank's sum of (KL_ij - m)^2 over its rows
};
static_assert(sizeof(RowsHdr) <= 1024, "header layout");
__host__ __device__ inline int64_t rows_landing_off() { return 1024; }

__device__ __forceinline__ unsigned long long rows_now_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// bounded spin until *flag == want; returns false on timeout or when the buffer is already in the error state
__device__ __forceinline__ bool rows_wait(const unsigned int* flag, unsigned int want, volatile unsigned int* error) {
    if (*error != 0u) return false;
    const unsigned long long t0 = rows_now_ns();
    while (ld_acquire_sys_u32(flag) != want) {
        if (rows_now_ns() - t0 > kRTimeoutNs) return false;
    }
    return true;
}

__device__ __forceinline__ float rows_logsigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

struct RowsSmem {
    double cs[kRStats][kRD];
    double piv[4][kRD];
    double part[kRG][kRStats][kRD];      // column statistics per row group (phase 1), reused as scratch afterwards
    double scratch[3][kRT / 32];
    double xs[kRC];                      // CTA 0: every CTA's partial sum of squares
    double kls[kRC];                     // every CTA: every CTA's partial sum of the analytic KL
    float4 Vo[kRD][kPo];                 // {mc, a, b, r} of the own rows
    float4 Vp[kRD][kPp];                 // ... of the partner slice
    float pivpart[2][kRG][kRD];
    float rso[3][2][kROwn], rsp[3][2][kRSl];   // E / F / KL row sums, one partial per 32-dim half
    float Eo[kROwn], Fo[kROwn], Ep[kRSl], Fp[kRSl], Kp[kRSl];
    float w1[kROwn][kPp], w2[kROwn][kPp];      // KL_ij - m, KL_ji - m (0 where excluded)
    float gm[kRD];
    float gk;
    int bad;
};

__device__ __forceinline__ void rows_block_sum3(double& a, double& b, double& c, double (*scratch)[kRT / 32]) {
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
    constexpr int NW = kRT / 32;
    a = warp_sum(lane < NW ? scratch[0][lane] : 0.0);
    b = warp_sum(lane < NW ? scratch[1][lane] : 0.0);
    c = warp_sum(lane < NW ? scratch[2][lane] : 0.0);
}

// local: this rank's packed rows [b_loc][2D] (mu | logvar).  peers: device table of the ranks' communication buffers as
// mapped in this process (entry rank = own buffer).  gstats [16][8][64] doubles and gpart [16][64][2][64] floats: scratch.
__global__ void __launch_bounds__(kRT, 1)
kl_mpd_rows_cluster_kernel(const float* __restrict__ local, char* const* __restrict__ peers, int rank, int world, int64_t slot_floats,
                           int b_loc, int D, float eta, float gamma, float free_bits, float kl_weight,
                           float* __restrict__ m_d_out, float* __restrict__ S_out, float* __restrict__ kl_all_out,
                           float* __restrict__ reg_out, float* __restrict__ dmu, float* __restrict__ dlv,
                           double* __restrict__ gstats, float* __restrict__ gpart) {
    extern __shared__ __align__(16) unsigned char rows_smem[];
    RowsSmem& sm = *reinterpret_cast<RowsSmem*>(rows_smem);
    cg::cluster_group cluster = cg::this_cluster();
    const int c = (int)cluster.block_rank();
    const int tid = threadIdx.x, lane = tid & 31, w = tid >> 5;
    const int d = tid & (kRD - 1), g = tid >> 6, half = w & 1;
    const bool dok = d < D;
    const int B = world * b_loc;
    const int SL = (B + kRC - 1) / kRC;
    const int j0 = c * SL;                                   // first global row of the slice
    const int nsl = min(SL, max(B - j0, 0));                 // rows in the slice (0 for trailing CTAs of small batches)
    const int row0 = rank * b_loc;
    const int W = 2 * D;                                     // floats per packed row
    const double nb = (double)B;

    char* own = peers[rank];
    RowsHdr* hdr = reinterpret_cast<RowsHdr*>(own);
    volatile unsigned int* err_flag = &hdr->error;
    const unsigned int e = hdr->epoch[c] + 1;
    const int parity = e & 1;
    const float* landing = reinterpret_cast<const float*>(own + rows_landing_off()) + (int64_t)parity * world * slot_floats;

    // ------------------------------------------------------------------ phase 0: push, flag, wait
    if (tid == 0) sm.bad = 0;
    if (c < world && c != rank) {
        char* dst_base = peers[c];
        float* dst = reinterpret_cast<float*>(dst_base + rows_landing_off()) + ((int64_t)parity * world + rank) * slot_floats;
        const int n4 = (b_loc * W) / 4;
        const float4* src4 = reinterpret_cast<const float4*>(local);
        float4* dst4 = reinterpret_cast<float4*>(dst);
        for (int i = tid; i < n4; i += kRT) dst4[i] = src4[i];
        __syncthreads();
        if (tid == 0) {
            __threadfence_system();
            st_release_sys_u32(&reinterpret_cast<RowsHdr*>(dst_base)->flags[parity][rank], e);
        }
    }
    __syncthreads();
    if (tid == 0) {
        // ranks whose rows this CTA reads: the owner(s) of the slice and rank 0's first rows (pivots)
        bool ok = true;
        const int p_hi = min(B, kRPivotRows) - 1;
        for (int q = 0; q < world && ok; ++q) {
            if (q == rank) continue;
            const int q_lo = q * b_loc, q_hi = q_lo + b_loc - 1;
            const bool need = (nsl > 0 && q_hi >= j0 && q_lo <= j0 + nsl - 1) || (q_lo <= p_hi);
            if (need) ok = rows_wait(&hdr->flags[parity][q], e, err_flag);
        }
        if (!ok) {
            hdr->error = 1u;
            sm.bad = 1;
        }
        hdr->epoch[c] = e;
    }
    __syncthreads();
    auto row_ptr = [&](int gi) -> const float* {            // packed row of global index gi (landed rows are read in place)
        const int q = gi / b_loc, li = gi - q * b_loc;
        return q == rank ? local + (int64_t)li * W : landing + (int64_t)q * slot_floats + (int64_t)li * W;
    };
    auto ld_row = [&](int gi, int col) -> float {
        const int q = gi / b_loc;
        const float* p = row_ptr(gi) + col;
        return q == rank ? __ldg(p) : __ldcg(p);             // peers' NVLink writes: L2 is the coherence point
    };

    // ------------------------------------------------------------------ phase 1: pivots, elements, statistics
    const int P = B < kRPivotRows ? B : kRPivotRows;
    {
        float pm = 0.f, pl = 0.f;
        for (int i = g; i < P; i += kRG)
            if (dok) {
                pm += ld_row(i, d);
                pl += ld_row(i, D + d);
            }
        sm.pivpart[0][g][d] = pm;
        sm.pivpart[1][g][d] = pl;
    }
    // raw values of the rows this thread evaluates: own rows g, g+8, ... and slice rows g, g+8, ...
    constexpr int RO = kROwn / kRG;   // 8
    constexpr int RP = kRSl / kRG;    // 4
    float mu_o[RO], lv_o[RO], mu_p[RP], lv_p[RP];
#pragma unroll
    for (int q = 0; q < RO; ++q) {
        const int i = g + q * kRG;
        const bool ok = dok && i < b_loc;
        mu_o[q] = ok ? __ldg(local + (int64_t)i * W + d) : 0.f;
        lv_o[q] = ok ? __ldg(local + (int64_t)i * W + D + d) : 0.f;
    }
#pragma unroll
    for (int q = 0; q < RP; ++q) {
        const int jl = g + q * kRG;
        const bool ok = dok && jl < nsl;
        mu_p[q] = ok ? ld_row(j0 + jl, d) : 0.f;
        lv_p[q] = ok ? ld_row(j0 + jl, D + d) : 0.f;
    }
    __syncthreads();
    float mu_bar = 0.f, l_bar = 0.f;
#pragma unroll
    for (int gg = 0; gg < kRG; ++gg) {
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
        // own rows: V, E, F (no statistics: they are counted by the CTA whose slice holds them)
#pragma unroll
        for (int q = 0; q < RO; ++q) {
            const int i = g + q * kRG;
            float es = 0.f, fs = 0.f;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dok && i < b_loc) {
                const Elem el = make_elem(mu_o[q], lv_o[q], mu_bar, l_bar, v_bar);
                v4 = make_float4(el.mc, el.a, el.b, el.r);
                es = el.ema;
                fs = el.bpa;
            }
            sm.Vo[d][i] = v4;
            es = warp_sum(es);
            fs = warp_sum(fs);
            if (lane == 0) {
                sm.rso[0][half][i] = es;
                sm.rso[1][half][i] = fs;
            }
        }
        float acc[kRStats];
#pragma unroll
        for (int q = 0; q < kRStats; ++q) acc[q] = 0.f;
#pragma unroll
        for (int q = 0; q < RP; ++q) {
            const int jl = g + q * kRG;
            float es = 0.f, fs = 0.f, ks = 0.f;
            float4 v4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (dok && jl < nsl) {
                const Elem el = make_elem(mu_p[q], lv_p[q], mu_bar, l_bar, v_bar);
                const float sq = el.mc * el.mc;
                v4 = make_float4(el.mc, el.a, el.b, el.r);
                es = el.ema;
                fs = el.bpa;
                ks = fmaf(mu_p[q], mu_p[q], expm1_minus_x(lv_p[q]));   // gaussian_encoder.py:219, cancellation-free
                acc[0] += el.a;
                acc[1] += el.b;
                acc[2] = fmaf(el.a, el.b, acc[2]);
                acc[3] += el.mc;
                acc[4] += sq;
                acc[5] = fmaf(el.r, el.mc, acc[5]);
                acc[6] = fmaf(el.r, sq, acc[6]);
                acc[7] += el.apb;
            }
            sm.Vp[d][jl] = v4;
            es = warp_sum(es);
            fs = warp_sum(fs);
            ks = warp_sum(ks);
            if (lane == 0) {
                sm.rsp[0][half][jl] = es;
                sm.rsp[1][half][jl] = fs;
                sm.rsp[2][half][jl] = ks;
            }
        }
#pragma unroll
        for (int q = 0; q < kRStats; ++q) sm.part[g][q][d] = (double)acc[q];
    }
    __syncthreads();
    {
        const int q = g;   // statistic (kRT / kRD == kRStats)
        double t = 0.0;
#pragma unroll
        for (int gg = 0; gg < kRG; ++gg) t += sm.part[gg][q][d];
        gstats[((int64_t)c * kRStats + q) * kRD + d] = t;
        for (int i = tid; i < kROwn; i += kRT) {
            sm.Eo[i] = sm.rso[0][0][i] + sm.rso[0][1][i];
            sm.Fo[i] = sm.rso[1][0][i] + sm.rso[1][1][i];
        }
        for (int jl = tid; jl < kRSl; jl += kRT) {
            sm.Ep[jl] = sm.rsp[0][0][jl] + sm.rsp[0][1][jl];
            sm.Fp[jl] = sm.rsp[1][0][jl] + sm.rsp[1][1][jl];
            const float k = 0.5f * (sm.rsp[2][0][jl] + sm.rsp[2][1][jl]);
            sm.Kp[jl] = k;
            if (jl < nsl && kl_all_out != nullptr) kl_all_out[j0 + jl] = k;   // overwritten with NaN below after a timeout
        }
    }
    __syncthreads();
    {
        // partial sum of the analytic KL over the slice -> every CTA (DSMEM), fixed order
        double ksum = (tid < nsl) ? (double)sm.Kp[tid] : 0.0, z0 = 0.0, z1 = 0.0;
        rows_block_sum3(ksum, z0, z1, sm.scratch);
        if (tid < kRC) *cluster.map_shared_rank(&sm.kls[c], tid) = ksum;
    }
    __threadfence();
    cluster.sync();   // (1) statistics of every slice are in gstats, KL sums in kls

    // ------------------------------------------------------------------ phase 2: m_d, m, KL of (own rows) x (slice)
    {
        const int q = g;
        double t = 0.0;
        for (int cc = 0; cc < kRC; ++cc) t += __ldcg(gstats + ((int64_t)cc * kRStats + q) * kRD + d);
        sm.cs[q][d] = t;
    }
    __syncthreads();
    double mdv = 0.0, lsig = 0.0, z2 = 0.0;
    if (tid < D) {
        const double SA = sm.cs[0][tid], SB = sm.cs[1][tid], SAB = sm.cs[2][tid], S1 = sm.cs[3][tid], S2 = sm.cs[4][tid],
                     R1 = sm.cs[5][tid], R2 = sm.cs[6][tid], SAPB = sm.cs[7][tid];
        const double r0 = (nb + SB) * sm.piv[3][tid];
        const double tn = SA * SB - SAB + (nb - 1.0) * SAPB + (S2 * r0 - 2.0 * S1 * R1 + nb * R2);
        mdv = 0.5 / (nb * (nb - 1.0)) * tn;
        const float v = (float)mdv;
        const float ex = expf(-fabsf(v));
        const float sig_neg = v >= 0.f ? ex / (1.0f + ex) : 1.0f / (1.0f + ex);
        sm.gm[tid] = -eta * sig_neg;   // d/dm [-eta * logsigmoid(m)]
        lsig = (double)rows_logsigmoid(v);
        if (c == 0) m_d_out[tid] = v;
    }
    rows_block_sum3(mdv, lsig, z2, sm.scratch);
    const float m = (float)mdv;
    double klsum = 0.0;
    for (int cc = 0; cc < kRC; ++cc) klsum += sm.kls[cc];
    const float mean_kl = (float)(klsum * (double)kl_weight);
    if (tid == 0) sm.gk = (kl_weight > 0.f && mean_kl >= free_bits) ? kl_weight : 0.f;
    {
        // thread = (slice row jl = lane, 4 own rows 4w..4w+3): both orientations share the squared difference
        constexpr int RW = kROwn / (kRT / 32);   // 4 own rows per warp
        const int jl = lane;
        float a1[RW], a2[RW];
#pragma unroll
        for (int r = 0; r < RW; ++r) a1[r] = a2[r] = 0.f;
        if (nsl > 0) {
#pragma unroll 4
            for (int dc = 0; dc < D; ++dc) {
                const float4 vj = sm.Vp[dc][jl];
#pragma unroll
                for (int r = 0; r < RW; ++r) {
                    const float4 vi = sm.Vo[dc][w * RW + r];   // warp-wide broadcast
                    const float dl = vi.x - vj.x;
                    const float q2 = dl * dl;
                    a1[r] = fmaf(vi.y, vj.z, fmaf(q2, vj.w, a1[r]));
                    a2[r] = fmaf(vj.y, vi.z, fmaf(q2, vi.w, a2[r]));
                }
            }
        }
        float psum = 0.f;
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const int i = w * RW + r;
            const bool ok = i < b_loc && jl < nsl && (j0 + jl) != (row0 + i);
            float kij = 0.f, kji = 0.f;
            if (ok) {
                kij = 0.5f * (a1[r] + sm.Eo[i] + sm.Fp[jl]) - m;
                kji = 0.5f * (a2[r] + sm.Ep[jl] + sm.Fo[i]) - m;
                psum = fmaf(kij, kij, psum);
            }
            sm.w1[i][jl] = kij;
            sm.w2[i][jl] = kji;
        }
        double mine = (double)psum, z0 = 0.0, z1 = 0.0;
        rows_block_sum3(mine, z0, z1, sm.scratch);
        if (tid == 0) *cluster.map_shared_rank(&sm.xs[c], 0) = mine;   // -> CTA 0
    }
    cluster.sync();   // (2) CTA 0 holds every slice's partial sum of squares
    if (c == 0 && tid == 0) {
        double mine = 0.0;
        for (int cc = 0; cc < kRC; ++cc) mine += sm.xs[cc];
        // this rank's sum -> every rank (own buffer included), then the flag
        for (int q = 0; q < world; ++q) {
            RowsHdr* h = reinterpret_cast<RowsHdr*>(peers[q]);
            h->sums[parity][rank] = mine;
        }
        __threadfence_system();
        for (int q = 0; q < world; ++q) st_release_sys_u32(&reinterpret_cast<RowsHdr*>(peers[q])->flags2[parity][rank], e);
    }
    if (tid == 0) {
        bool ok = true;
        for (int q = 0; q < world && ok; ++q) ok = rows_wait(&hdr->flags2[parity][q], e, err_flag);
        if (!ok) {
            hdr->error = 1u;
            sm.bad = 1;
        }
    }
    __syncthreads();
    double total = 0.0;
    for (int q = 0; q < world; ++q) total += __ldcg(&hdr->sums[parity][q]);   // rank order: identical on every rank
    const float S = (float)sqrt(total / (nb * (nb - 1.0) - 1.0));
    float reg = 0.f;
    if (c == 0 && tid == 0) {
        reg = kl_weight > 0.f ? fmaxf(mean_kl, free_bits) : 0.f;
        if (eta > 0.f) reg -= eta * (float)lsig;
        if (gamma > 0.f) reg += gamma * S;
    }

    // ------------------------------------------------------------------ phase 3: partial gradient over the slice
    float beta = 0.f;   // d(gamma * S) / d KL_ij = beta * (KL_ij - m)     (SURVEY.md §8a "A3-bwd")
    if (gamma > 0.f && S > 0.f) beta = (float)((double)gamma / ((nb * nb - nb - 1.0) * (double)S));
    if (dmu != nullptr) {
        // thread = (latent dim d, row group g): own rows 8g..8g+7 against every row of the slice
        constexpr int RT = kROwn / kRG;   // 8
        float4 pi[RT];
        float ti[RT], oi[RT], gmu[RT], glv[RT];
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            pi[r] = sm.Vo[d][g * RT + r];
            ti[r] = pi[r].y + (2.0f * pi[r].z + pi[r].z * pi[r].z) * (1.0f + pi[r].y);   // r_i^2 v_i v_bar - 1
            oi[r] = (1.0f + ti[r]) * ivb;
            gmu[r] = glv[r] = 0.f;
        }
        if (beta != 0.f && dok) {
            for (int jl = 0; jl < nsl; ++jl) {
                const float4 pj = sm.Vp[d][jl];
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float w1v = sm.w1[g * RT + r][jl], w2v = sm.w2[g * RT + r][jl];   // warp-wide broadcasts
                    const float dl = pi[r].x - pj.x;
                    gmu[r] = fmaf(dl, fmaf(w1v, pj.w, w2v * pi[r].w), gmu[r]);
                    glv[r] = fmaf(w1v, fmaf(pi[r].y + 1.0f, pj.z, pi[r].y), glv[r]);
                    glv[r] = fmaf(-w2v, fmaf(ti[r] + 1.0f, pj.y, fmaf(dl * dl, oi[r], ti[r])), glv[r]);
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RT; ++r) {
            const int i = g * RT + r;
            gpart[(((int64_t)c * kROwn + i) * 2 + 0) * kRD + d] = gmu[r];
            gpart[(((int64_t)c * kROwn + i) * 2 + 1) * kRD + d] = glv[r];
        }
        __threadfence();
    }
    cluster.sync();   // (3) every slice's partial gradients are in gpart; every error flag raised so far is visible
    const bool poisoned = (*err_flag != 0u);
    const float nanv = __int_as_float(0x7fc00000);
    if (c == 0 && tid == 0) {
        S_out[0] = poisoned ? nanv : S;
        if (reg_out != nullptr) reg_out[0] = poisoned ? nanv : reg;
    }
    if (poisoned) {
        if (c == 0 && tid < D) m_d_out[tid] = nanv;
        if (kl_all_out != nullptr)
            for (int jl = tid; jl < nsl; jl += kRT) kl_all_out[j0 + jl] = nanv;
    }
    if (dmu == nullptr) return;

    // ------------------------------------------------------------------ phase 4: CTA c finishes own rows 4c .. 4c+3
    {
        constexpr int RF = kROwn / kRC;   // 4
        const int r = g;                   // groups 0..3 carry a row each
        const int i = c * RF + r;
        if (r < RF && dok && i < b_loc) {
            float gmu = 0.f, glv = 0.f;
            if (beta != 0.f) {
                for (int cc = 0; cc < kRC; ++cc) {
                    gmu += __ldcg(gpart + (((int64_t)cc * kROwn + i) * 2 + 0) * kRD + d);
                    glv += __ldcg(gpart + (((int64_t)cc * kROwn + i) * 2 + 1) * kRD + d);
                }
                gmu *= beta;
                glv *= 0.5f * beta;
            }
            const float4 p = sm.Vo[d][i];
            const float t_i = p.y + (2.0f * p.z + p.z * p.z) * (1.0f + p.y);
            mean_part(&sm.cs[0][0], &sm.piv[0][0], eta > 0.f ? sm.gm : nullptr, B, kRD, d, make_float4(p.x, p.y, p.z, t_i),
                      __ldg(local + (int64_t)i * W + d), __ldg(local + (int64_t)i * W + D + d), gmu, glv,
                      sm.gk != 0.f ? &sm.gk : nullptr);
            dmu[(int64_t)i * D + d] = poisoned ? nanv : gmu;
            dlv[(int64_t)i * D + d] = poisoned ? nanv : glv;
        }
    }
}

int rows_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, cudaStream_t stream) {
    auto kern = kl_mpd_rows_cluster_kernel;
    static bool done = false;
    if (!done) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(RowsSmem));
        if (e != cudaSuccess) return (int)e;
        e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        if (e != cudaSuccess) return (int)e;
        done = true;
    }
    cfg = cudaLaunchConfig_t{};
    cfg.gridDim = dim3(kRC);
    cfg.blockDim = dim3(kRT);
    cfg.dynamicSmemBytes = sizeof(RowsSmem);
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = kRC;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return MAE_OK;
}

bool rows_shape_ok(int64_t world, int64_t b_loc, int64_t D) {
    return world >= 2 && world <= kRMaxWorld && b_loc >= 1 && b_loc <= kROwn && D >= 1 && D <= kRD && world * b_loc <= kRMaxB &&
           world * b_loc >= 2 && (b_loc * 2 * D) % 4 == 0;
}

}  // namespace
}  // namespace mae

extern "C" int64_t mae_rows_comm_bytes(int64_t world, int64_t slot_floats) {
    if (world < 1 || world > mae::kRMaxWorld || slot_floats < 1) return 0;
    return mae::rows_landing_off() + 2 * world * mae::round_up(slot_floats, 4) * 4;
}

extern "C" int64_t mae_rows_scratch_bytes(void) {
    return (int64_t)mae::kRC * mae::kRStats * mae::kRD * 8 + (int64_t)mae::kRC * mae::kROwn * 2 * mae::kRD * 4;
}

extern "C" int mae_kl_mpd_rows_fused_supported(int64_t world, int64_t b_loc, int64_t D) {
    using namespace mae;
    if (!rows_shape_ok(world, b_loc, D)) return 0;
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    if (rows_config(cfg, attr, nullptr) != MAE_OK) return (void)cudaGetLastError(), 0;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kl_mpd_rows_cluster_kernel, &cfg) != cudaSuccess) return (void)cudaGetLastError(), 0;
    return n >= 1 ? 1 : 0;
}

extern "C" int mae_kl_mpd_rows_fused(const float* local_rows, void* const* peer_buffers, int64_t rank, int64_t world,
                                     int64_t slot_floats, int64_t b_loc, int64_t D, float eta, float gamma, float free_bits,
                                     float kl_weight, float* m_d, float* S, float* kl_all, float* reg, float* dmu, float* dlogvar,
                                     void* scratch, int64_t scratch_bytes, mae_stream_t stream) {
    using namespace mae;
    MAE_REQUIRE(local_rows && peer_buffers && m_d && S && scratch, MAE_ERR_NULL);
    MAE_REQUIRE((dmu == nullptr) == (dlogvar == nullptr), MAE_ERR_NULL);
    MAE_REQUIRE(rank >= 0 && rank < world, MAE_ERR_SHAPE);
    MAE_REQUIRE(rows_shape_ok(world, b_loc, D), MAE_ERR_UNSUPPORTED);
    MAE_REQUIRE(slot_floats % 4 == 0 && slot_floats >= b_loc * 2 * D, MAE_ERR_SHAPE);
    MAE_REQUIRE((reinterpret_cast<uintptr_t>(local_rows) & 15) == 0 && (reinterpret_cast<uintptr_t>(scratch) & 15) == 0, MAE_ERR_ALIGN);
    MAE_REQUIRE(scratch_bytes >= mae_rows_scratch_bytes(), MAE_ERR_WORKSPACE);
    cudaLaunchConfig_t cfg;
    cudaLaunchAttribute attr[1];
    int rc = rows_config(cfg, attr, as_stream(stream));
    if (rc != MAE_OK) return rc;
    double* gstats = static_cast<double*>(scratch);
    float* gpart = reinterpret_cast<float*>(static_cast<char*>(scratch) + (int64_t)kRC * kRStats * kRD * 8);
    cudaError_t e = cudaLaunchKernelEx(&cfg, kl_mpd_rows_cluster_kernel, local_rows, reinterpret_cast<char* const*>(peer_buffers),
                                       (int)rank, (int)world, slot_floats, (int)b_loc, (int)D, eta, gamma, free_bits, kl_weight,
                                       m_d, S, kl_all, reg, dmu, dlogvar, gstats, gpart);
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}
