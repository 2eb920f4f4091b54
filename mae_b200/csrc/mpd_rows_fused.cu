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
//     phase 1  pivots (first <= 32 global rows), per-element quantities {mc, a, b, r}, row sums E / F, analytic KL and column
//              statistics of the SLICE only: every global row is evaluated once, by the CTA whose slice holds it, and
//              published through a global scratch buffer                                         [cluster.sync 1]
//     phase 2  fetch the own rows' elements, closed-form m_d, m;  KL_ij and KL_ji for own i, slice j (pivoted pair form of
//              mpd_fused.cu) -> partial sum of (KL_ij - m)^2 -> CTA 0 -> the ranks exchange ONE double each [cluster.sync 2]
//     phase 3  partial gradient sums of every own row over the slice (expanded form, linear in the weights, so it runs
//              while the ranks' sums are in flight) -> global scratch; then wait for the sums -> S   [cluster.sync 3]
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

#ifdef MAE_DEBUG_TIMING
// phase timestamps of CTA 0 (ns, %globaltimer) behind the scratch area: profiling builds only
#define ROWS_STAMP(slot)                                                                     \
    do {                                                                                     \
        __syncthreads();                                                                     \
        if (threadIdx.x == 0 && c == 0) dbg[slot] = rows_now_ns();                           \
    } while (0)
#else
#define ROWS_STAMP(slot) do {} while (0)
#endif

__device__ __forceinline__ float rows_logsigmoid(float x) { return fminf(x, 0.f) - log1pf(expf(-fabsf(x))); }

struct RowsSmem {
    double cs[kRStats][kRD];
    double piv[4][kRD];
    double part[kRG][kRStats][kRD];      // column statistics per row group (phase 1)
    double scratch[3][kRT / 32];
    double xs[kRC];                      // CTA 0: every CTA's partial sum of squares
    double kls[kRC];                     // every CTA: every CTA's partial sum of the analytic KL
    float4 Vo[kRD][kPo];                 // {mc, a, b, r} of the own rows
    float4 Vp[kRD][kPp];                 // ... of the partner slice
    float pivpart[2][kRG][kRD];
    float rsp[3][2][kRSl];               // E / F / KL row sums of the slice, one partial per 32-dim half
    float Eo[kROwn], Fo[kROwn], Ep[kRSl], Fp[kRSl], Kp[kRSl];
    float w1[kROwn][kRSl], w2[kROwn][kRSl];    // KL_ij - m, KL_ji - m (0 where excluded); rows are 128 B: float4 reads over j
    float gm[kRD];
    float gk;
    const float* prow[kRPivotRows];      // packed pivot rows (own memory or landing buffer)
    const float* srow[kRSl];             // packed rows of the slice
    int bad;
};

// global scratch (device memory owned by the caller): everything the CTAs of a cluster exchange
struct RowsScratch {
    double gstats[kRC][kRStats][kRD];    // column statistics of every slice
    float4 Vg[kRMaxB][kRD];              // {mc, a, b, r} of every global row, written by the CTA whose slice holds it
    float Eg[kRMaxB], Fg[kRMaxB];
    float gA[kRC][kROwn][6][kRD];        // partial gradient sums of the own rows over every slice
    float grs[kRC][kROwn][2];            // partial row sums of w1 / w2
    unsigned long long dbg[16];
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

// KL_ij / KL_ji of (own rows) x (slice): thread = (slice row jl = tid % NSLP, own rows RW * (tid / NSLP) ...): every thread
// of the CTA carries RW pairs whatever the world size (NSLP = slice rows rounded up to 8 / 16 / 32)
template <int NSLP>
__device__ __forceinline__ float rows_pairs(RowsSmem& sm, int D, int b_loc, int nsl, int j0, int row0, float m) {
    constexpr int RW = NSLP / 8;                  // own rows per thread: 512 threads cover 64 rows x NSLP slice rows
    const int jl = threadIdx.x % NSLP, i0 = (threadIdx.x / NSLP) * RW;
    float a1[RW], a2[RW];
#pragma unroll
    for (int r = 0; r < RW; ++r) a1[r] = a2[r] = 0.f;
#pragma unroll 4
    for (int dc = 0; dc < D; ++dc) {
        const float4 vj = sm.Vp[dc][jl];
#pragma unroll
        for (int r = 0; r < RW; ++r) {
            const float4 vi = sm.Vo[dc][i0 + r];
            const float dl = vi.x - vj.x;
            const float q2 = dl * dl;
            a1[r] = fmaf(vi.y, vj.z, fmaf(q2, vj.w, a1[r]));
            a2[r] = fmaf(vj.y, vi.z, fmaf(q2, vi.w, a2[r]));
        }
    }
    float psum = 0.f;
#pragma unroll
    for (int r = 0; r < RW; ++r) {
        const int i = i0 + r;
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
    return psum;
}

// local: this rank's packed rows [b_loc][2D] (mu | logvar).  peers: device table of the ranks' communication buffers as
// mapped in this process (entry rank = own buffer).
__global__ void __launch_bounds__(kRT, 1)
kl_mpd_rows_cluster_kernel(const float* __restrict__ local, char* const* __restrict__ peers, int rank, int world, int64_t slot_floats,
                           int b_loc, int D, float eta, float gamma, float free_bits, float kl_weight,
                           float* __restrict__ m_d_out, float* __restrict__ S_out, float* __restrict__ kl_all_out,
                           float* __restrict__ reg_out, float* __restrict__ dmu, float* __restrict__ dlv, RowsScratch* __restrict__ gs) {
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

#ifdef MAE_DEBUG_TIMING
    unsigned long long* dbg = gs->dbg;
#endif
    ROWS_STAMP(0);
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
    // packed-row pointers of the pivot rows and of the slice (landed rows are read in place, no copy-out)
    const int P = B < kRPivotRows ? B : kRPivotRows;
    if (tid < kRPivotRows + kRSl) {
        const int gi = tid < kRPivotRows ? tid : j0 + (tid - kRPivotRows);
        const bool valid = tid < kRPivotRows ? (tid < P) : (tid - kRPivotRows < nsl);
        const float* p = local;
        if (valid) {
            const int q = gi / b_loc, li = gi - q * b_loc;
            p = q == rank ? local + (int64_t)li * W : landing + (int64_t)q * slot_floats + (int64_t)li * W;
        }
        if (tid < kRPivotRows) sm.prow[tid] = p;
        else sm.srow[tid - kRPivotRows] = p;
    }
    __syncthreads();
    ROWS_STAMP(1);   // rows pushed
    if (tid == 0) {
        // ranks whose rows this CTA reads: the owner(s) of the slice and of the pivot rows
        bool ok = true;
        const int p_hi = P - 1;
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
    ROWS_STAMP(2);   // peers' rows landed

    // ------------------------------------------------------------------ phase 1: pivots, elements of the SLICE, statistics
    // (peers' NVLink writes: L2 is the coherence point, hence .cg loads)
    constexpr int RP = kRSl / kRG;    // 4 slice rows per thread
    float mu_p[RP], lv_p[RP];
#pragma unroll
    for (int q = 0; q < RP; ++q) {
        const int jl = g + q * kRG;
        const bool ok = dok && jl < nsl;
        mu_p[q] = ok ? __ldcg(sm.srow[jl] + d) : 0.f;
        lv_p[q] = ok ? __ldcg(sm.srow[jl] + D + d) : 0.f;
    }
    {
        float pm = 0.f, pl = 0.f;
#pragma unroll
        for (int q = 0; q < kRPivotRows / kRG; ++q) {
            const int i = g + q * kRG;
            if (dok && i < P) {
                pm += __ldcg(sm.prow[i] + d);
                pl += __ldcg(sm.prow[i] + D + d);
            }
        }
        sm.pivpart[0][g][d] = pm;
        sm.pivpart[1][g][d] = pl;
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
            if (jl < nsl) gs->Vg[j0 + jl][d] = v4;      // every global row is evaluated ONCE, by the CTA of its slice
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
        gs->gstats[c][q][d] = t;
        if (tid < kRSl) {
            const int jl = tid;
            const float ev = sm.rsp[0][0][jl] + sm.rsp[0][1][jl], fv = sm.rsp[1][0][jl] + sm.rsp[1][1][jl];
            const float k = 0.5f * (sm.rsp[2][0][jl] + sm.rsp[2][1][jl]);
            sm.Ep[jl] = ev;
            sm.Fp[jl] = fv;
            sm.Kp[jl] = k;
            if (jl < nsl) {
                gs->Eg[j0 + jl] = ev;
                gs->Fg[j0 + jl] = fv;
                if (kl_all_out != nullptr) kl_all_out[j0 + jl] = k;   // overwritten with NaN below after a timeout
            }
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
    ROWS_STAMP(3);   // phase 1 done
    cluster.sync();   // (1) every slice's elements, row sums and statistics are in the global scratch, KL sums in kls
    ROWS_STAMP(4);

    // ------------------------------------------------------------------ phase 2: own rows, m_d, m, KL of (own rows) x (slice)
    {
        constexpr int RO = kROwn / kRG;   // 8 own rows per thread
        float4 vo[RO];
#pragma unroll
        for (int q = 0; q < RO; ++q) {
            const int i = g + q * kRG;
            vo[q] = (dok && i < b_loc) ? __ldcg(&gs->Vg[row0 + i][d]) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        double t = 0.0;
        {
            double tq[kRC];
#pragma unroll
            for (int cc = 0; cc < kRC; ++cc) tq[cc] = __ldcg(&gs->gstats[cc][g][d]);
#pragma unroll
            for (int cc = 0; cc < kRC; ++cc) t += tq[cc];
        }
        sm.cs[g][d] = t;
#pragma unroll
        for (int q = 0; q < RO; ++q) sm.Vo[d][g + q * kRG] = vo[q];
        if (tid < kROwn) {
            sm.Eo[tid] = tid < b_loc ? __ldcg(&gs->Eg[row0 + tid]) : 0.f;
            sm.Fo[tid] = tid < b_loc ? __ldcg(&gs->Fg[row0 + tid]) : 0.f;
        }
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
        float psum;
        if (nsl <= 8) psum = rows_pairs<8>(sm, D, b_loc, nsl, j0, row0, m);
        else if (nsl <= 16) psum = rows_pairs<16>(sm, D, b_loc, nsl, j0, row0, m);
        else psum = rows_pairs<32>(sm, D, b_loc, nsl, j0, row0, m);
        // rows_pairs<NSLP> writes every column < NSLP >= the float4 groups phase 3 reads
        double mine = (double)psum, z0 = 0.0, z1 = 0.0;
        rows_block_sum3(mine, z0, z1, sm.scratch);
        if (tid == 0) *cluster.map_shared_rank(&sm.xs[c], 0) = mine;   // -> CTA 0
    }
    ROWS_STAMP(5);   // phase 2 done
    cluster.sync();   // (2) CTA 0 holds every slice's partial sum of squares
    ROWS_STAMP(6);
    if (c == 0 && tid == 0) {
        double mine = 0.0;
        for (int cc = 0; cc < kRC; ++cc) mine += sm.xs[cc];
        // this rank's sum -> every rank (own buffer included), then the flag; the answer is awaited AFTER phase 3
        for (int q = 0; q < world; ++q) reinterpret_cast<RowsHdr*>(peers[q])->sums[parity][rank] = mine;
        __threadfence_system();
        for (int q = 0; q < world; ++q) st_release_sys_u32(&reinterpret_cast<RowsHdr*>(peers[q])->flags2[parity][rank], e);
    }

    // ------------------------------------------------------------------ phase 3: partial gradient sums over the slice
    // Linear in the weights w1 = KL_ij - m, w2 = KL_ji - m, so it does not need S (beta multiplies at the very end) and
    // overlaps the cross-rank exchange of the sums.  Expanded (GEMM) form: with p = {mc, a, b, r} of partner j,
    //   A1 = sum w1 r_j   A2 = sum w1 mc_j r_j   A3 = sum w1 b_j        A4 = sum w2 mc_j   A5 = sum w2 a_j   A6 = sum w2 mc_j^2
    // 6 FFMA per (own row, partner, dim) instead of the 10 operations of the difference form.
    if (dmu != nullptr) {
        if (tid < kROwn) {   // partial row sums of the weights
            float r1 = 0.f, r2 = 0.f;
            for (int jl = 0; jl < nsl; ++jl) {
                r1 += sm.w1[tid][jl];
                r2 += sm.w2[tid][jl];
            }
            gs->grs[c][tid][0] = r1;
            gs->grs[c][tid][1] = r2;
        }
        // thread = (latent dim d, row group g): own rows 8g..8g+7 against every row of the slice, 4 partners at a time
        constexpr int RT = kROwn / kRG;   // 8
        float A[RT][6];
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) A[r][q] = 0.f;
        if (gamma > 0.f && dok) {
            for (int jb = 0; jb < nsl; jb += 4) {
                float pr[4], pmr[4], pb[4], pm[4], pa[4], pm2[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const float4 pj = sm.Vp[d][jb + u];      // rows beyond the slice are zero
                    pr[u] = pj.w;
                    pmr[u] = pj.x * pj.w;
                    pb[u] = pj.z;
                    pm[u] = pj.x;
                    pa[u] = pj.y;
                    pm2[u] = pj.x * pj.x;
                }
#pragma unroll
                for (int r = 0; r < RT; ++r) {
                    const float4 wa = *reinterpret_cast<const float4*>(&sm.w1[g * RT + r][jb]);   // warp-wide broadcasts
                    const float4 wb = *reinterpret_cast<const float4*>(&sm.w2[g * RT + r][jb]);
                    const float w1v[4] = {wa.x, wa.y, wa.z, wa.w};
                    const float w2v[4] = {wb.x, wb.y, wb.z, wb.w};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        A[r][0] = fmaf(w1v[u], pr[u], A[r][0]);
                        A[r][1] = fmaf(w1v[u], pmr[u], A[r][1]);
                        A[r][2] = fmaf(w1v[u], pb[u], A[r][2]);
                        A[r][3] = fmaf(w2v[u], pm[u], A[r][3]);
                        A[r][4] = fmaf(w2v[u], pa[u], A[r][4]);
                        A[r][5] = fmaf(w2v[u], pm2[u], A[r][5]);
                    }
                }
            }
        }
#pragma unroll
        for (int r = 0; r < RT; ++r)
#pragma unroll
            for (int q = 0; q < 6; ++q) gs->gA[c][g * RT + r][q][d] = A[r][q];
        __threadfence();
    }
    ROWS_STAMP(7);   // phase 3 done
    if (tid == 0) {
        bool ok = true;
        for (int q = 0; q < world && ok; ++q) ok = rows_wait(&hdr->flags2[parity][q], e, err_flag);
        if (!ok) {
            hdr->error = 1u;
            sm.bad = 1;
        }
    }
    __syncthreads();
    ROWS_STAMP(8);   // every rank's sum landed
    double total = 0.0;
    for (int q = 0; q < world; ++q) total += __ldcg(&hdr->sums[parity][q]);   // rank order: identical on every rank
    const float S = (float)sqrt(total / (nb * (nb - 1.0) - 1.0));
    float reg = 0.f;
    if (c == 0 && tid == 0) {
        reg = kl_weight > 0.f ? fmaxf(mean_kl, free_bits) : 0.f;
        if (eta > 0.f) reg -= eta * (float)lsig;
        if (gamma > 0.f) reg += gamma * S;
    }
    float beta = 0.f;   // d(gamma * S) / d KL_ij = beta * (KL_ij - m)     (SURVEY.md §8a "A3-bwd")
    if (gamma > 0.f && S > 0.f) beta = (float)((double)gamma / ((nb * nb - nb - 1.0) * (double)S));
    cluster.sync();   // (3) every slice's partial sums are in the scratch; every error flag raised so far is visible
    ROWS_STAMP(9);
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
            const float4 p = sm.Vo[d][i];   // {mc, a, b, r}
            const float t_i = p.y + (2.0f * p.z + p.z * p.z) * (1.0f + p.y);   // r_i^2 v_i v_bar - 1
            if (beta != 0.f) {
                float A[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f}, rs1 = 0.f, rs2 = 0.f;
                for (int cc = 0; cc < kRC; ++cc) {
#pragma unroll
                    for (int q = 0; q < 6; ++q) A[q] += __ldcg(&gs->gA[cc][i][q][d]);
                    rs1 += __ldcg(&gs->grs[cc][i][0]);
                    rs2 += __ldcg(&gs->grs[cc][i][1]);
                }
                const float mc = p.x, o_i = (1.0f + t_i) * ivb;
                // sum_j dl (w1 r_j + w2 r_i),  dl = mc_i - mc_j
                gmu = fmaf(mc, A[0], -A[1]) + p.w * fmaf(mc, rs2, -A[3]);
                // sum_j w1 ((a_i + 1) b_j + a_i) - w2 ((t_i + 1) a_j + t_i + dl^2 o_i)
                glv = fmaf(p.y + 1.0f, A[2], p.y * rs1) - fmaf(t_i + 1.0f, A[4], t_i * rs2) -
                      o_i * (fmaf(mc * mc, rs2, A[5]) - 2.0f * mc * A[3]);
                gmu *= beta;
                glv *= 0.5f * beta;
            }
            mean_part(&sm.cs[0][0], &sm.piv[0][0], eta > 0.f ? sm.gm : nullptr, B, kRD, d, make_float4(p.x, p.y, p.z, t_i),
                      __ldg(local + (int64_t)i * W + d), __ldg(local + (int64_t)i * W + D + d), gmu, glv,
                      sm.gk != 0.f ? &sm.gk : nullptr);
            dmu[(int64_t)i * D + d] = poisoned ? nanv : gmu;
            dlv[(int64_t)i * D + d] = poisoned ? nanv : glv;
        }
    }
    ROWS_STAMP(10);
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
    return (int64_t)sizeof(mae::RowsScratch);
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
    cudaError_t e = cudaLaunchKernelEx(&cfg, kl_mpd_rows_cluster_kernel, local_rows, reinterpret_cast<char* const*>(peer_buffers),
                                       (int)rank, (int)world, slot_floats, (int)b_loc, (int)D, eta, gamma, free_bits, kl_weight,
                                       m_d, S, kl_all, reg, dmu, dlogvar, static_cast<RowsScratch*>(scratch));
    if (e != cudaSuccess) return (int)e;
    return launch_status();
}
