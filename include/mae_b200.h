/*
 * mae_b200.h — C ABI of libmae_b200.so: the B200 (sm_100a) implementation of the loss-and-evaluation
 * hot path of XuezheMax/mae.
 *
 * The reference has no FFI / plugin API of its own (it is pure Python): its seam is
 * the Python class interface in mae/modules.  Every entry point below therefore cites the Python
 * function (reference file:line) whose arithmetic it replaces; the Python host layer in
 * mae_b200/ops.py binds these symbols with ctypes and mae_b200/modules mirrors the class interface.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to fp32 data unless stated otherwise; sizes are int64_t;
 *     "stride" arguments are ROW strides in elements (the reference's mu is a chunk(2,1) view with
 *     row stride 2*D, encoder_cores/resnet.py:43); the innermost dimension is always contiguous.
 *   - no allocation and no synchronisation inside: the caller owns every buffer, including
 *     workspaces and outputs; all work is enqueued on `stream` (a cudaStream_t), so every call can be
 *     captured into a CUDA graph.  The only process-wide state is the launch-attribute switch
 *     mae_set_pdl() and the peer-memory communication buffers (mae_comm_*), which the library allocates
 *     because they must be cudaMalloc'ed to be shareable through CUDA IPC.
 *   - workspaces must be zero-filled once when allocated (they carry self-resetting tickets) and
 *     must not be shared by calls that may run concurrently.
 *   - return value: 0 on success, a negative MAE_ERR_* code for an argument error, or a positive
 *     cudaError_t from the launch.  mae_strerror() maps either to text.
 *   - "nullable" pointers may be NULL; the documented default is used instead.
 */
#ifndef MAE_B200_H
#define MAE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* mae_stream_t; /* cudaStream_t */

#define MAE_ABI_VERSION 1

enum {
    MAE_OK = 0,
    MAE_ERR_NULL = -1,      /* a required pointer is NULL */
    MAE_ERR_SHAPE = -2,     /* a size is <= 0, too large, or inconsistent (e.g. MPD with B < 2) */
    MAE_ERR_STRIDE = -3,    /* a row stride is smaller than the row length */
    MAE_ERR_WORKSPACE = -4, /* workspace too small for the given sizes */
    MAE_ERR_ALIGN = -5,     /* a pointer is not 16-byte aligned where the kernel needs it */
    MAE_ERR_UNSUPPORTED = -6
};

int mae_abi_version(void);
/* Programmatic dependent launch between the kernels of a dependent chain (each kernel still waits for its predecessor's
 * completion before reading its output, so results are unaffected).  On by default; mae_set_pdl(0) switches it off for the
 * process (the batch-sharded head does: at 8 GPUs early-placed successors cost 8 us per step while ranks wait for each
 * other's rows); MAE_NO_PDL=1 in the environment has the same effect.  Returns the previous setting. */
int mae_set_pdl(int enabled);
/* Tensor-core tiles of the MPD regulariser (3xTF32 tcgen05.mma, csrc/mpd_tc.cuh) for padded batches >= 2048: on by default
 * (MAE_MPD_TC=0 in the environment starts with them off); off = the FP32 FFMA tile kernels.  Returns the previous setting. */
int mae_set_mpd_tensor_cores(int enabled);
const char* mae_strerror(int code);
/* SM count / compute capability of the current device; used by the host layer to refuse non-sm_100 parts. */
int mae_device_info(int* sm_count, int* cc_major, int* cc_minor);

/* ------------------------------------------------------------------------------------------------
 * A1 + A2  reparameterised sample and analytic KL
 *   replaces GaussianEncoder.reparameterize  (mae/modules/encoders/gaussian_encoder.py:56-62)
 *        and the analytic KL of encode()     (mae/modules/encoders/gaussian_encoder.py:219)
 *   z[b,s,d]  = eps[b,s,d] * exp(0.5*logvar[b,d]) + mu[b,d]
 *   kl[b]     = 0.5 * sum_d (mu^2 + exp(logvar) - logvar - 1)            (kl nullable: skipped)
 *   eps is drawn by the caller (torch Philox) so RNG parity with the reference is exact.
 * bwd: gz [B,ns,D] nullable, gkl [B] nullable ->
 *   dmu[b,d] (+)= sum_s gz + gkl[b]*mu ;  dlogvar[b,d] (+)= 0.5*exp(0.5*logvar)*sum_s(gz*eps) + 0.5*gkl[b]*(exp(logvar)-1)
 *   accumulate != 0 adds into dmu/dlogvar instead of overwriting (dmu/dlogvar are dense [B,D]).
 * ---------------------------------------------------------------------------------------------- */
int mae_reparam_kl_fwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                       const float* eps, int64_t B, int64_t ns, int64_t D,
                       float* z, float* kl, mae_stream_t stream);
int mae_reparam_kl_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                       const float* eps, const float* gz, const float* gkl,
                       int64_t B, int64_t ns, int64_t D,
                       float* dmu, float* dlogvar, int accumulate, mae_stream_t stream);
/* Same, plus the precomputed gradient of the encoder-side regulariser (extra_dmu / extra_dlogvar [B,D] dense, written by
 * mae_kl_mpd_loss_fused) times its upstream device scalar extra_scale[0] (NULL = 1): the whole gradient of a training
 * step w.r.t. the encoder outputs in one launch (no separate rescale, no accumulation kernels). */
int mae_reparam_reg_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                        const float* eps, const float* gz, const float* gkl,
                        const float* extra_dmu, const float* extra_dlogvar, const float* extra_scale,
                        int64_t B, int64_t ns, int64_t D,
                        float* dmu, float* dlogvar, int accumulate, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A3  mutual posterior divergence (MPD) regulariser
 *   replaces GaussianEncoder._postKL (mae/modules/encoders/gaussian_encoder.py:162-188)
 *
 *   KL_ij = 0.5 * sum_d [ v_i r_j + (mu_i-mu_j)^2 r_j - (lv_i - lv_j) - 1 ],  v = exp(lv), r = 1/(v+1e-12)
 *   m_d   = sum_{i!=j} KL_ij,d / (B(B-1))                      (per latent dimension, [D])
 *   S     = sqrt( sum_{i!=j} (KL_ij - m)^2 / (B(B-1) - 1) ),   m = sum_d m_d
 *   (this is what the Eye*mean fill, unbiased .std() and the cc/dd factors at :181-187 reduce to.)
 *
 * The [B,B,D] tensor is never formed: tiles of 2*KL_ij = x~_i . y~_j + E_i + F_j (K = 2D contraction in FP32 FFMA) are
 * produced in shared memory / registers from pivoted per-element quantities (pivots mu_bar, l_bar per latent dim;
 * x~ = [a + mc^2/v_bar, mc], y~ = [b, -2 mc r], a = expm1(lv - l_bar), b = v_bar r - 1, mc = mu - mu_bar), which keep full
 * relative precision when the posteriors coincide (posterior collapse), where the textbook form cancels.
 *
 * mae_mpd_prepare : O(B*D) pre-pass over ALL B posteriors (after an all-gather in multi-GPU runs):
 *                   packed panels, per-dimension column statistics, m_d (written to m_d[D]) and m.  If kl is
 *                   non-NULL it also writes the analytic KL(q_b || N(0,I)) [B] (gaussian_encoder.py:219), which
 *                   shares every operand with the pre-pass (one launch less on the training path).
 * mae_mpd_fwd     : tiles for rows [row0,row0+nrows) x all B columns.  Writes sumsq[0] (double) =
 *                   sum over those rows and all j != i of (KL_ij - m)^2, and, if S is non-NULL,
 *                   S[0] = sqrt(sumsq/(B(B-1)-1)) (meaningful when the row range covers all rows).
 *                   store_kl != 0 additionally keeps the KL tiles (and their transpose) in the workspace for the
 *                   stored-KL backward paths (requires the full row range and a workspace sized with store_kl).
 * mae_mpd_finalize: S[0] = sqrt(sumsq[0]/(B(B-1)-1)) after sumsq has been all-reduced across ranks.
 * mae_mpd_bwd     : gradient of  sum_d g_m[d]*m_d + g_S*S  w.r.t. mu and logvar of rows
 *                   [row0,row0+nrows): dmu/dlogvar are dense [nrows,D].  Each row's contribution both
 *                   as the "i" and as the "j" argument of KL_ij is included, so no cross-rank
 *                   reduce-scatter of column gradients is needed.  g_S and S are device scalars.
 *                   gkl (nullable, [B] indexed by global row) is the upstream gradient of the analytic KL written
 *                   by mae_mpd_prepare; its contribution gkl*mu / 0.5*gkl*(exp(logvar)-1) is added in the epilogue.
 *                   path: 0 = choose automatically, 1 = stored-KL per-row kernel (B <= 2048), 2 = recompute tiles
 *                   (nothing stored), 3 = stored-KL tile GEMMs (any B whose KL tiles were kept by mae_mpd_fwd:
 *                   8*B^2*D flop instead of 16*B^2*D, plus one streaming read of the two B x B matrices).
 * ---------------------------------------------------------------------------------------------- */
#define MAE_MPD_SMALL_B 2048
/* store_kl != 0: room for KL[Bp][Bp] and its transpose (8*Bp^2 bytes, Bp = B rounded up to 64) at the end of the
 * workspace, needed by the stored-KL backward paths 1 and 3; B200's 180 GB make this affordable far beyond the point
 * where recomputing the tiles (path 2, 16*B^2*D flop) is the only option (B = 65536 -> 34 GB). */
int64_t mae_mpd_workspace_bytes(int64_t B, int64_t D, int store_kl);
/* workspace for a ROW BLOCK [row0, row0 + nrows) that keeps KL_ij and KL_ji of its rows (store_kl of mae_mpd_fwd with a
 * row block; sharded runs): everything of mae_mpd_workspace_bytes(B, D, 0) plus 2 * rows_pad * Bp floats. */
int64_t mae_mpd_workspace_bytes_rows(int64_t B, int64_t D, int64_t row0, int64_t nrows);
int mae_mpd_prepare(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                    int64_t B, int64_t D, void* workspace, int64_t workspace_bytes,
                    float* m_d, float* kl, mae_stream_t stream);
int mae_mpd_fwd(void* workspace, int64_t workspace_bytes, int64_t B, int64_t D,
                int64_t row0, int64_t nrows, int store_kl,
                double* sumsq, float* S, mae_stream_t stream);
int mae_mpd_finalize(const double* sumsq, int64_t B, float* S, mae_stream_t stream);
int mae_mpd_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                void* workspace, int64_t workspace_bytes, int64_t B, int64_t D,
                int64_t row0, int64_t nrows,
                const float* g_m, const float* g_S, const float* S, const float* gkl,
                float* dmu, float* dlogvar, int accumulate, int path, mae_stream_t stream);

/* The regulariser's consumers are closed forms (mae.py:132-139), so its gradient does not have to wait for the objective:
 *     reg = clamp(kl_weight * sum_b kl_all[b], min = free_bits) - eta * sum_d logsigmoid(m_d[d]) + gamma * S
 * mae_mpd_loss_bwd = d reg / d(mu, logvar) of rows [row0,row0+nrows) (dense [nrows,D]) and reg[0] (nullable), enqueued in
 * the FORWARD pass right behind mae_mpd_prepare / mae_mpd_fwd (two launches: the upstream gradients g_m, g_S, gkl are
 * derived on the device into the workspace, then the mae_mpd_bwd kernel of the chosen path runs).  m_d / S / kl_all are the
 * outputs of prepare / fwd (after mae_mpd_finalize in row-sharded runs); kl_all NULL or kl_weight = 0 leaves the KL term
 * out; dmu / dlogvar NULL: only reg is written.  Works for any B, D (the general-shape companion of
 * mae_kl_mpd_loss_fused below). */
int mae_mpd_loss_bwd(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                     void* workspace, int64_t workspace_bytes, int64_t B, int64_t D,
                     int64_t row0, int64_t nrows, float eta, float gamma, float free_bits, float kl_weight,
                     const float* m_d, const float* S, const float* kl_all, float* reg,
                     float* dmu, float* dlogvar, int path, mae_stream_t stream);

/* Training-sized batches (B <= 128, D <= 64: MNIST / OMNIGLOT / CIFAR heads): the whole encoder-side regulariser
 *     reg = clamp(kl_weight * sum_b KL_b, min = free_bits) - eta * sum_d logsigmoid(m_d) + gamma * S
 * (mae.py:132-139; KL_b gaussian_encoder.py:219; m_d, S gaussian_encoder.py:162-188) forward AND backward in ONE launch of
 * one thread-block cluster, no workspace.  Writes m_d[D], S[1], kl[B] (nullable), reg[1] (nullable) and, when dmu /
 * dlogvar are non-NULL, d reg / d mu and d reg / d logvar [B,D] (dense), which mae_reparam_reg_bwd folds into the
 * encoder-output gradient together with their upstream scalar.
 * kl_weight = 1/B for the reference's KL.mean(); 0 leaves the KL term out (flows: KL is a Monte-Carlo estimate).
 * mae_kl_mpd_loss_fused_supported: 1 if this (B, D) can take the path on the current device, else 0 (use the
 * prepare / fwd / bwd entry points above). */
int mae_kl_mpd_loss_fused_supported(int64_t B, int64_t D);
int mae_kl_mpd_loss_fused(const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                          int64_t B, int64_t D, float eta, float gamma, float free_bits, float kl_weight,
                          float* m_d, float* S, float* kl, float* reg, float* dmu, float* dlogvar, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A4a  Bernoulli reconstruction error
 *   replaces BinaryImageDecoder.reconstruct_error (.../image_decoders/binary_image_decoder.py:84-93)
 *        and its duplicate in PixelCNNDecoderBinaryImage28x28 (.../image_decoders/pixelcnn.py:136-139)
 *   err[b,s] = - sum_p [ log(probs[b,s,p] + 1e-12) * x[b,p] + log(1 - probs[b,s,p] + 1e-12) * (1 - x[b,p]) ]
 *   probs are post-sigmoid probabilities [B,ns,P]; x is [B,P].
 * bwd: dprobs[b,s,p] = G[b,s] * ( (1-x)/(1-probs+1e-12) - x/(probs+1e-12) ),  G[n] = g[n*g_stride]*g_mul
 *      (g_stride = 1: per-sample upstream gradient; g_stride = 0: g is ONE device scalar, e.g. d loss, and
 *       g_mul = 1/(B*ns) — lets the likelihood backward start without waiting for the objective's backward)
 * ---------------------------------------------------------------------------------------------- */
int mae_bernoulli_fwd(const float* probs, const float* x, int64_t B, int64_t ns, int64_t P,
                      float* err, mae_stream_t stream);
int mae_bernoulli_bwd(const float* probs, const float* x, const float* g, int64_t g_stride, float g_mul,
                      int64_t B, int64_t ns, int64_t P, float* dprobs, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A4b + A4c  discretized mixture of logistics, from the RAW decoder output
 *   replaces ColorImageDecoder.execute            (.../image_decoders/color_image_decoder.py:38-52)
 *            ColorImageDecoder.reconstruct_error  (.../image_decoders/color_image_decoder.py:91-125)
 *            discretized_mix_logistic_loss        (mae/modules/utils.py:61-123)
 *   raw [B*ns, 10*nmix, HW] (NCHW: logits | means | log-scales | coefficients, mixture-major and
 *   colour-minor inside each group), x [B,3,HW] in [-1,1]; err [B,ns].
 *   log-softmax of the logits, clamp(min=-7) of the log-scales and tanh of the coefficients happen
 *   inside the kernel.  bin 1/255, lower 1/255-1, upper 1-1/255 as in the reference.
 *   workspace: mae_dmol_workspace_bytes(B*ns, HW) bytes (per-CTA partial sums + tickets).
 * bwd: draw [B*ns,10*nmix,HW] = G[b,s] * d err[b,s] / d raw   (clamp gate passes at equality),
 *      G[n] = g[n*g_stride]*g_mul as for the Bernoulli backward.
 * ---------------------------------------------------------------------------------------------- */
int64_t mae_dmol_workspace_bytes(int64_t N, int64_t HW);
int mae_dmol_fwd(const float* raw, const float* x, int64_t B, int64_t ns, int64_t nmix, int64_t HW,
                 float* err, void* workspace, int64_t workspace_bytes, mae_stream_t stream);
int mae_dmol_bwd(const float* raw, const float* x, const float* g, int64_t g_stride, float g_mul,
                 int64_t B, int64_t ns, int64_t nmix, int64_t HW, float* draw, mae_stream_t stream);
/* Fused forward+backward for the training step (the objective is a mean over samples, so d loss/d err is the constant
 * g_mul = 1/(B*ns) times d loss, known up to the scalar d loss): ONE pass reads raw, writes err[b,s] and
 * draw = g_mul * d err/d raw  (53 MB instead of 27 + 53 MB at the C3 shape).  The autograd backward then only calls
 * mae_scale_inplace(draw, d loss), which returns without touching memory when d loss == 1 (loss.backward()).
 * Only nmix = 10 and H*W in {1024, 4096} (MAE_ERR_UNSUPPORTED otherwise: use fwd + bwd).
 * err_partial is [B*ns, mae_dmol_err_chunks(...)]: per-CTA partial sums, reduced by mae_loss_assemble_fwd(err_chunks)
 * (no tickets / cross-CTA dependency in the heavy kernel). */
int64_t mae_dmol_err_chunks(int64_t B, int64_t ns, int64_t nmix, int64_t HW);
int mae_dmol_fwd_bwd(const float* raw, const float* x, float g_mul, int64_t B, int64_t ns, int64_t nmix, int64_t HW,
                     float* err_partial, float* draw, mae_stream_t stream);
/* data[i] *= scale[0] (device scalar); a no-op kernel when scale[0] == 1. */
int mae_scale_inplace(float* data, int64_t n, const float* scale, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * N2  the decoder's final 1x1 convolution fused with the DMOL likelihood (tcgen05 3xTF32, accumulators in TMEM)
 *   replaces the LAST layer of the colour decoders' conv stacks together with A4b + A4c:
 *     nn.Sequential(..., nn.ELU(), Conv2dWeightNorm(hidden, 10*nmix, 1))   image_decoders/pixelcnnpp.py:32-34  (act = 1)
 *     Conv2dWeightNorm(48, 10*nmix, 1)                                      image_decoders/resnet.py:85         (act = 0)
 *     ColorImageDecoder.reconstruct_error                                   color_image_decoder.py:91-125
 *   hidden [B*ns, C, HW] fp32 NCHW: the input of that last layer (before its ELU when act = 1);
 *   weight [10*nmix, C] row-major: the EFFECTIVE weights (weight normalisation g * v / |v| applied by the caller,
 *   weight_norm.py) in the reference's channel order; bias [10*nmix] (nullable); x [B, 3, HW] in [-1, 1].
 *   err [B*ns] = -log p(x | z) summed over pixels.  draw (nullable) [B*ns, 10*nmix, HW] = g_mul * d err / d raw for the
 *   convolution's backward (plain library GEMMs on the caller's side); the raw tensor itself is never written.
 *   supported(C, nmix, HW): nmix = 10 with C in {48, 64, 96}, nmix = 5 with C = 64; HW a multiple of 128.
 * ---------------------------------------------------------------------------------------------- */
int mae_dmol_head_supported(int64_t C, int64_t nmix, int64_t HW);
int64_t mae_dmol_head_workspace_bytes(int64_t N, int64_t HW);
int mae_dmol_head(const float* hidden, const float* weight, const float* bias, const float* x, int64_t B, int64_t ns,
                  int64_t nmix, int64_t C, int64_t HW, int act, float g_mul, float* err, float* draw,
                  void* workspace, int64_t workspace_bytes, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A5 / A6  Gaussian log densities
 *   replaces GaussianEncoder.log_probability_prior     (gaussian_encoder.py:231-260, lines 256-258)
 *            GaussianEncoder.log_probability_posterior (gaussian_encoder.py:262-301, lines 297-299)
 *   prior:     out[n]   = -0.5 * sum_d (z[n,d]^2 + log 2pi) + logdet[n]                (logdet nullable = 0)
 *   posterior: out[b,k] = -0.5 * sum_d (lv + (z-mu)^2/(exp(lv)+1e-12) + log 2pi) + logdet[b,k]
 *   the flow transforms that produce z_normal/logdet stay in PyTorch.
 * bwd (for the Monte-Carlo KL branch of encode(), gaussian_encoder.py:220-225):
 *   prior:     dz = -g[n]*z
 *   posterior: dz = -g*(z-mu)*r ; dmu = sum_k g*(z-mu)*r ; dlogvar = sum_k -0.5*g*(1 - (z-mu)^2 r^2 v)
 *   (d logdet = g is a pass-through done by the host layer).
 * ---------------------------------------------------------------------------------------------- */
int mae_log_prob_prior_fwd(const float* z_normal, const float* logdet, int64_t N, int64_t D,
                           float* out, mae_stream_t stream);
int mae_log_prob_prior_bwd(const float* z_normal, const float* g, int64_t N, int64_t D,
                           float* dz, mae_stream_t stream);
int mae_log_prob_posterior_fwd(const float* z_normal, const float* logdet,
                               const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                               int64_t B, int64_t K, int64_t D, float* out, mae_stream_t stream);
int mae_log_prob_posterior_bwd(const float* z_normal, const float* g,
                               const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
                               int64_t B, int64_t K, int64_t D,
                               float* dz, float* dmu, float* dlogvar, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A5 + A6 + A7  fused importance-weighted NLL estimator
 *   replaces the tail of MAE.nll (mae/modules/mae.py:203-215) and logsumexp (mae/modules/utils.py:23-38)
 *   log_gen [B,K] (= -reconstruct_error), z_prior_normal [B,K,D] + logdet_prior [B,K] (nullable = 0),
 *   z_post_normal [B,K,D] + logdet_post [B,K] (nullable = 0), mu/logvar [B,D]  ->
 *   nll_elbo[b] = -mean_k log_iw ; recon[b] = -mean_k log_gen ; kl[b] = mean_k (log q - log p) ;
 *   nll_iw[b]   = -(max_k log_iw + log sum_k exp(log_iw - max)) + log K ,  log_iw = log_gen + log p - log q
 *   gen_scale multiplies log_gen on the fly: -1 feeds the reconstruction ERROR written by mae_dmol_fwd / mae_bernoulli_fwd
 *   directly (log p(x|z) = -error, decoder.py log_probability), saving a negation pass.
 *   workspace: mae_iw_nll_workspace_bytes(B, K) bytes (per-chunk partial states + tickets), zero-filled once.
 */
int64_t mae_iw_nll_workspace_bytes(int64_t B, int64_t K);
int mae_iw_nll(const float* log_gen, const float* z_prior_normal, const float* logdet_prior,
               const float* z_post_normal, const float* logdet_post,
               const float* mu, int64_t mu_stride, const float* logvar, int64_t logvar_stride,
               int64_t B, int64_t K, int64_t D, float gen_scale,
               float* nll_elbo, float* recon, float* kl, float* nll_iw,
               void* workspace, int64_t workspace_bytes, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * A3' + A8  objective assembly
 *   replaces MAE.loss lines mae/modules/mae.py:132-140
 *   err [B,ns], kl [B], m_d [D] (nullable), S [1] (nullable) ->
 *   recon[b] = mean_s err[b,s]
 *   scalars[0] = loss = mean_b recon + max(mean_b kl, free_bits) + loss_mean + loss_std
 *   scalars[1] = postKL_mean = sum_d m_d         scalars[2] = loss_mean = -eta * sum_d logsigmoid(m_d)  (0 if eta<=0)
 *   scalars[3] = loss_std = gamma * S (0 if gamma<=0)      scalars[4] = mean_b kl       scalars[5] = mean_b recon
 *   inv_batch: 1/B_global (>0) lets multi-GPU callers pass local tensors; 0 -> 1/B.
 *   ext_sums (nullable, device float[2]): {sum of recon, sum of kl} contributed by the OTHER ranks of a
 *   batch-sharded run; they are added to the local sums before the means (and the free-bits clamp) are taken.
 *   err_chunks (0 -> 1): err is [B, ns, err_chunks] per-CTA partial sums (mae_dmol_fwd_bwd) that are added here.
 *   kl_count (0 -> B): number of entries of `kl` summed for the KL mean; a batch-sharded caller passes the
 *   all-gathered KL vector [B_global] here (err stays local [B,ns]) instead of a separate reduction.
 * bwd, given gloss (device scalar, nullable = 1):
 *   derr[b,s] = gloss*inv_batch/ns ; dkl[b] = gloss*inv_batch*[mean kl >= free_bits] ;
 *   dm_d[d] = -gloss*eta*sigmoid(-m_d) ; dS = gloss*gamma
 *   kl_mean is read from scalars[4] written by the forward.
 * ---------------------------------------------------------------------------------------------- */
int mae_loss_assemble_fwd(const float* err, const float* kl, const float* m_d, const float* S, const float* ext_sums,
                          int64_t B, int64_t ns, int64_t D, int64_t kl_count, int64_t err_chunks,
                          float eta, float gamma, float free_bits, float inv_batch,
                          float* scalars, float* recon, mae_stream_t stream);
int mae_loss_assemble_bwd(const float* gloss, const float* scalars, const float* m_d,
                          int64_t B, int64_t ns, int64_t D, float eta, float gamma, float free_bits, float inv_batch,
                          float* derr, float* dkl, float* dm_d, float* dS, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU: all-gather of the packed (mu | logvar) rows over NVLink peer memory (replaces the
 * DataParallel gather of gaussian_encoder.py:41-42 for the loss head).  One kernel per rank pushes its rows into
 * every peer's landing buffer, flags, waits for the peers' flags and copies the landed slots to `out`
 * [world, n_floats].  Buffers come from mae_comm_alloc (cudaMalloc, zeroed) and are opened by the peers through
 * CUDA IPC handles (64 bytes each, exchanged by the host layer over torch.distributed).
 *   peer_buffers: DEVICE array of `world` pointers, entry r = rank r's buffer as mapped in this process.
 *   Spins are bounded (2 s): on timeout the error flag (mae_comm_error) is set instead of hanging the GPU.
 * ---------------------------------------------------------------------------------------------- */
int64_t mae_comm_buffer_bytes(int64_t world, int64_t slot_floats);
int mae_comm_alloc(int64_t bytes, void** ptr);
int mae_comm_free(void* ptr);
int mae_comm_get_handle(void* ptr, void* handle64);
int mae_comm_open_handle(const void* handle64, void** ptr);
int mae_comm_close_handle(void* ptr);
int mae_comm_error(const void* own_buffer, int* error_out);
int mae_allgather_rows(const float* local, int64_t n_floats, void* const* peer_buffers, int64_t rank, int64_t world,
                       int64_t slot_floats, float* out, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Multi-GPU, training batches: the exchange of the packed rows AND the regulariser over the global batch (forward, and
 * backward for this rank's rows) in ONE cluster launch per rank (csrc/mpd_rows_fused.cu).  Supported shapes:
 * 2 <= world <= 8, b_loc <= 64, D <= 64, world * b_loc <= 512 (mae_kl_mpd_rows_fused_supported).
 *   local_rows   this rank's packed rows [b_loc][2 D] = (mu | logvar), 16-byte aligned
 *   peer_buffers DEVICE array of `world` pointers to communication buffers of mae_rows_comm_bytes(world, slot_floats)
 *                bytes (mae_comm_alloc, opened through IPC like the gather's); slot_floats >= b_loc * 2 D, multiple of 4
 *   outputs      m_d [D], S [1], kl_all [world * b_loc] (analytic KL of every global row, nullable), reg [1] (nullable),
 *                dmu / dlogvar [b_loc][D] = d reg / d(mu, logvar) of the own rows (both NULL: forward only)
 *   scratch      mae_rows_scratch_bytes() bytes of device memory owned by the caller (no initialisation needed)
 * Same timeout behaviour as mae_allgather_rows: the sticky error flag (mae_comm_error) and NaN outputs.
 * ---------------------------------------------------------------------------------------------- */
int64_t mae_rows_comm_bytes(int64_t world, int64_t slot_floats);
int64_t mae_rows_scratch_bytes(void);
int mae_kl_mpd_rows_fused_supported(int64_t world, int64_t b_loc, int64_t D);
int mae_kl_mpd_rows_fused(const float* local_rows, void* const* peer_buffers, int64_t rank, int64_t world,
                          int64_t slot_floats, int64_t b_loc, int64_t D, float eta, float gamma, float free_bits,
                          float kl_weight, float* m_d, float* S, float* kl_all, float* reg, float* dmu, float* dlogvar,
                          void* scratch, int64_t scratch_bytes, mae_stream_t stream);

/* ------------------------------------------------------------------------------------------------
 * Pipe-peak micro-benchmarks (measurement support, SURVEY.md §7 / §8d: the FP32-FFMA and SFU peaks that the MPD and
 * DMOL rooflines are quoted against are measured on the box, not assumed).  Register-only loops; the caller times
 * the launch with CUDA events.  *ops_out (host pointer, nullable) receives the lane-level instruction count of the
 * launch (FFMA: flop = 2 * ops).  mufu op: 0 ex2.approx, 1 lg2.approx, 2 rcp.approx.
 * ---------------------------------------------------------------------------------------------- */
int mae_microbench_ffma(int64_t blocks, int64_t iters, float* sink, double* ops_out, mae_stream_t stream);
int mae_microbench_mufu(int op, int64_t blocks, int64_t iters, float* sink, double* ops_out, mae_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* MAE_B200_H */
