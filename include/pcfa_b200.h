/*
 * pcfa_b200.h — C ABI of libpcfa_b200.so, the B200 (sm_100a) implementation of the
 * cost-volume / warping / objective operators on PCFA's optimisation hot path.
 *
 * Conventions (all entry points):
 *   - plain C, raw DEVICE pointers, int sizes, `stream` is a cudaStream_t passed as void*;
 *   - tensors are dense, contiguous, fp32, NCHW unless stated; index arithmetic is int32/int64;
 *   - asynchronous and stream-ordered: no allocation, no host synchronisation, re-entrant;
 *   - return 0 on success, a positive cudaError_t value if a CUDA call/launch failed, or a
 *     negative PCFA_E_* code for argument errors.  Nothing throws.
 *   - outputs documented "(overwritten)" need no initialisation; "(accumulated)" are += targets.
 *
 * Every function cites the reference interface it replaces (paths relative to the
 * cv-stuttgart/PCFA checkout).  The reference binds these through pybind11 modules taking
 * at::Tensor; the stub a maintainer would add on the reference side is in INTEGRATION.md.
 */
#ifndef PCFA_B200_H
#define PCFA_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCFA_ABI_VERSION 1

#define PCFA_OK            0
#define PCFA_E_BADARG     -1   /* null pointer / non-positive size / unsupported parameter */
#define PCFA_E_TOOLARGE   -2   /* a dimension exceeds what the int32 index maps support     */
#define PCFA_E_NODEVICE   -3   /* no sm_100 device / driver entry point unavailable         */
#define PCFA_E_WORKSPACE  -4   /* workspace pointer null or too small                       */

typedef void* pcfa_stream_t;

/* ---- library info -------------------------------------------------------------------- */
int         pcfa_abi_version(void);
/* Human-readable text for a status returned by any entry point (static storage). */
const char* pcfa_status_string(int status);
/* Number of kernels launched by this library since load (bench.py's gpu_launches). */
int64_t     pcfa_launch_count(void);

/* ======================================================================================
 * (a1) All-pairs correlation pyramid — replaces CorrBlock.__init__ + CorrBlock.corr
 *      models/raft/corr.py:13-27,52-60 ; models/gma/corr.py:16-30,55-63
 *
 *   level0[b*N+i, y, x] = (1/sqrt(C)) * sum_c fmap1[b,c,i] * fmap2[b,c,y*W+x],  N = H*W
 *   level(l+1) = avg_pool2d(level l, 2, stride 2)   (floor sizes: H_{l+1} = H_l/2)
 *
 * The pyramid lives in ONE flat fp32 buffer; level l starts at offsets[l] floats and is a dense
 * [B*N, H_l, W_l] array.  pcfa_corr_pyramid_layout fills offsets/hs/ws (host arrays of
 * num_levels entries, offsets has num_levels+1: the last one is the total float count).
 * ====================================================================================== */
int pcfa_corr_pyramid_layout(int B, int H, int W, int num_levels,
                             int64_t* offsets, int* hs, int* ws);

/* Bytes of scratch needed by pcfa_corr_pyramid_forward / _backward for these sizes. */
int64_t pcfa_corr_pyramid_workspace_bytes(int B, int C, int H, int W, int num_levels);

/* impl: 0 = auto (tcgen05 path when the shape qualifies, otherwise SIMT), 1 = force SIMT fp32,
 *       2 = force tcgen05.  Forward: bf16 hi/mid split (3 products), fp32 accumulate in TMEM, needs
 *       C % 64 == 0 and C <= 256.  Backward: TF32 gradient pyramid x TF32 hi/lo split features, needs
 *       C % 16 == 0, C <= 256 and every level's H_l*W_l a multiple of 4 (TMA stride rule). */
int pcfa_corr_pyramid_forward(const float* fmap1, const float* fmap2,
                              float* pyramid /* (overwritten) */,
                              void* workspace, int64_t workspace_bytes,
                              int B, int C, int H, int W, int num_levels, int impl,
                              pcfa_stream_t stream);

/* Backward of the build (autograd of matmul + avg_pool2d in the reference, corr.py:25-27,58-60).
 * grad_pyramid has the pyramid's flat layout and holds dL/d(level l) for every level.
 *   grad_fmap1[b,c,i] = (1/sqrt C) sum_l sum_j g_l[b*N+i, j] * pool_l(fmap2)[b,c,j]
 *   grad_fmap2        = (1/sqrt C) sum_l unpool_l( g_l^T * fmap1 )
 * grad_fmap1 / grad_fmap2 are overwritten. */
int pcfa_corr_pyramid_backward(const float* grad_pyramid,
                               const float* fmap1, const float* fmap2,
                               float* grad_fmap1, float* grad_fmap2,
                               void* workspace, int64_t workspace_bytes,
                               int B, int C, int H, int W, int num_levels, int impl,
                               pcfa_stream_t stream);

/* ======================================================================================
 * (a2/a3) Multi-level bilinear lookup — replaces CorrBlock.__call__ → bilinear_sampler →
 *      F.grid_sample(align_corners=True, zeros)   models/raft/corr.py:29-50,
 *      models/raft/utils/utils.py:57-71 (same in models/gma/corr.py:32-53)
 *
 *   coords [B,2,H,W] (channel 0 = x, 1 = y, level-0 pixel units)
 *   out    [B, num_levels*(2r+1)^2, H, W],  channel = l*(2r+1)^2 + a*(2r+1) + b,
 *          sample position (x/2^l + a - r, y/2^l + b - r)   — a moves x (x-major window).
 * Backward scatters into grad_pyramid (accumulated; flat pyramid layout); coords get no
 * gradient (detached in the reference, models/raft/raft.py:123).
 * ====================================================================================== */
int pcfa_corr_lookup_forward(const float* pyramid, const float* coords,
                             float* out /* (overwritten) */,
                             int B, int H, int W, int num_levels, int radius,
                             pcfa_stream_t stream);

/* channels-last variants: out / grad_out are [B][H][W][num_levels*(2r+1)^2] in memory (torch.channels_last) */
int pcfa_corr_lookup_forward_cl(const float* pyramid, const float* coords, float* out, int B, int H, int W,
                                int num_levels, int radius, pcfa_stream_t stream);
int pcfa_corr_lookup_backward_cl(const float* grad_out, const float* coords, float* grad_pyramid, int B, int H, int W,
                                 int num_levels, int radius, pcfa_stream_t stream);
int pcfa_corr_lookup_backward(const float* grad_out, const float* coords,
                              float* grad_pyramid /* (accumulated) */,
                              int B, int H, int W, int num_levels, int radius,
                              pcfa_stream_t stream);

/* Sparse backward (no reference counterpart: the reference's autograd materialises and sums one dense 261 MB gradient per
 * lookup, models/raft/corr.py:29-50).  The lookup windows of all iterations cover only part of the gradient pyramid (29-52 %
 * of its 32-query x 32-cell blocks at 55x128, profiles/g_sparsity_r2.txt).  `occupancy` is a bitmap of those blocks
 * (uint32 words [B][ceil(HW/32)][words], pcfa_corr_occupancy_bytes):
 *   pcfa_corr_occupancy_mark          ONE launch per backward pass: sets the bits of every block the lookups at
 *                                     coords_list[0..n_lookups) (a HOST array of device pointers, each [B,2,H,W] as passed to
 *                                     pcfa_corr_lookup_backward[_cl]) may have written; bits are OR-ed in
 *   pcfa_corr_pyramid_backward_occ    = pcfa_corr_pyramid_backward, reading only marked blocks (tensor-core path; the other
 *                                     paths ignore the bitmap)
 * Contract: the caller zero-fills `occupancy` (e.g. together with grad_pyramid) and marks EVERY lookup that accumulated into
 * grad_pyramid; an unmarked non-zero block is silently dropped. */
int64_t pcfa_corr_occupancy_bytes(int B, int H, int W, int num_levels);
int pcfa_corr_occupancy_mark(const float* const* coords_list, int n_lookups, uint32_t* occupancy /* (bits set) */,
                             int B, int H, int W, int num_levels, int radius, pcfa_stream_t stream);
int pcfa_corr_pyramid_backward_occ(const float* grad_pyramid, const uint32_t* occupancy, const float* fmap1, const float* fmap2,
                                   float* grad_fmap1, float* grad_fmap2, void* workspace, int64_t workspace_bytes,
                                   int B, int C, int H, int W, int num_levels, int impl, pcfa_stream_t stream);

/* ======================================================================================
 * (a5) Spatial correlation sampler — replaces spatial_correlation_sampler_backend.forward /
 *      .backward  (Correlation_Module/correlation_sampler.cpp:58-117, CPU semantics of
 *      Correlation_Module/correlation.cpp:75-178).  PWCNet call: models/PWCNet/PWCNet.py:45-58.
 *
 *   out[n,ph,pw,h,w] = sum_c sum_{i<kH} sum_{j<kW} in1[n,c,i1,j1] * in2[n,c,i2,j2]
 *   i1 = -padH + h*dH + i*dilH, i2 = i1 + (ph - (patchH-1)/2)*dilPatchH  (same for j / W);
 *   terms with any index out of range are skipped.  out is [B,patchH,patchW,oH,oW] with
 *   oH = (iH + 2*padH - ((kH-1)*dilH+1))/dH + 1.   `scale` multiplies the result (1.0 for the
 *   sampler itself; PWCNet's "/C" may be folded in by the caller).
 * ====================================================================================== */
typedef struct pcfa_scs_params {
    int kH, kW, patchH, patchW, padH, padW, dilH, dilW, dilPatchH, dilPatchW, dH, dW;
} pcfa_scs_params;

int pcfa_scs_output_size(int iH, int iW, const pcfa_scs_params* p, int* oH, int* oW);

int pcfa_scs_forward(const float* in1, const float* in2, float* out /* (overwritten) */,
                     int B, int C, int iH, int iW, const pcfa_scs_params* p, float scale,
                     pcfa_stream_t stream);

int pcfa_scs_backward(const float* in1, const float* in2, const float* grad_out,
                      float* grad_in1, float* grad_in2 /* (both overwritten) */,
                      int B, int C, int iH, int iW, const pcfa_scs_params* p, float scale,
                      pcfa_stream_t stream);

/* ======================================================================================
 * (a7) FlowNet2 correlation — replaces correlation_cuda.forward / .backward
 *      models/FlowNet/correlation_package/correlation_cuda.cc:10-87,89-171 and the kernels in
 *      correlation_cuda_kernel.cu:46-334.  The padded NHWC copies rbot1/rbot2 of the reference are
 *      not materialised (zero padding is a predicate).  corr_multiply is accepted and ignored, as
 *      in the reference.  Output [B, D*D, oH, oW], D = 2*(max_disp/stride2)+1, already divided by
 *      kernel_size^2 * C.  Backward follows the reference's truncating integer division
 *      (correlation_cuda_kernel.cu:172-176,290-294).
 * ====================================================================================== */
int pcfa_fn2corr_output_size(int H, int W, int pad_size, int kernel_size, int max_disp,
                             int stride1, int stride2, int* outC, int* oH, int* oW);

int pcfa_fn2corr_forward(const float* in1, const float* in2, float* out /* (overwritten) */,
                         int B, int C, int H, int W,
                         int pad_size, int kernel_size, int max_disp, int stride1, int stride2,
                         pcfa_stream_t stream);

int pcfa_fn2corr_backward(const float* in1, const float* in2, const float* grad_out,
                          float* grad_in1, float* grad_in2 /* (both overwritten) */,
                          int B, int C, int H, int W,
                          int pad_size, int kernel_size, int max_disp, int stride1, int stride2,
                          pcfa_stream_t stream);

/* ======================================================================================
 * (a8) Resample2d — replaces resample2d_cuda.forward / .backward
 *      models/FlowNet/resample2d_package/resample2d_cuda.cc:6-31, resample2d_kernel.cu:15-198.
 *   img [B,C,H,W], flow [B,2,oH,oW] → out [B,C,oH,oW]; border-clamped bilinear (or nearest when
 *   bilinear == 0).  Only kernel_size == 1 is supported (the only value the reference uses,
 *   models/FlowNet/FlowNet2.py:39-80).  Backward reproduces the reference's quirks: image weights
 *   use xf - trunc(xf) (kernel.cu:105-106), flow gradient uses floor (kernel.cu:163-193).
 * ====================================================================================== */
int pcfa_resample2d_forward(const float* img, const float* flow, float* out /* (overwritten) */,
                            int B, int C, int H, int W, int oH, int oW,
                            int kernel_size, int bilinear, pcfa_stream_t stream);

int pcfa_resample2d_backward(const float* img, const float* flow, const float* grad_out,
                             float* grad_img  /* (accumulated; caller zero-fills like resample2d.py:38) */,
                             float* grad_flow /* (overwritten) */,
                             int B, int C, int H, int W, int oH, int oW,
                             int kernel_size, int bilinear, pcfa_stream_t stream);

/* ======================================================================================
 * (a9) ChannelNorm — replaces channelnorm_cuda.forward / .backward
 *      models/FlowNet/channelnorm_package/channelnorm_cuda.cc:6-30, channelnorm_kernel.cu:18-96.
 *   out[b,0,y,x] = sqrt(sum_c x^2);  gin = gout * x / (out + 1e-9).  norm_deg is ignored (L2).
 * ====================================================================================== */
int pcfa_channelnorm_forward(const float* x, float* out, int B, int C, int H, int W,
                             int norm_deg, pcfa_stream_t stream);
int pcfa_channelnorm_backward(const float* x, const float* out, const float* grad_out,
                              float* grad_x /* (overwritten) */, int B, int C, int H, int W,
                              int norm_deg, pcfa_stream_t stream);

/* ======================================================================================
 * (a6) PWCNet backward warp — replaces PWCDCNet.warp (models/PWCNet/PWCNet.py:166-206):
 *   two F.grid_sample calls (bilinear, zeros, align_corners=False on an align_corners=True-style
 *   normalisation) and the >= 1e-4 validity mask, fused.
 *   x [B,C,H,W], flow [B,2,H,W] → out [B,C,H,W].  Backward gives d/dx (accumulated, caller
 *   zero-fills) and d/dflow (overwritten); the mask carries no gradient.
 * ====================================================================================== */
int pcfa_pwc_warp_forward(const float* x, const float* flow, float* out,
                          int B, int C, int H, int W, pcfa_stream_t stream);
int pcfa_pwc_warp_backward(const float* x, const float* flow, const float* grad_out,
                           float* grad_x, float* grad_flow,
                           int B, int C, int H, int W, pcfa_stream_t stream);

/* ======================================================================================
 * (a10) Box constraint / input transform — replaces ScaledInputModel.forward's pre-processing
 *      (helper_functions/own_models.py:62-85) fused with extract_deltas / extract_deltas_joint
 *      (attack_PCFA.py:20-37).
 *
 *   mode PCFA_BOX_COV      : var = w (per image);   x = clamp(0.5*(tanh(w)+(1-eps))/(1-eps),0,1)
 *                            delta = 0.5*(tanh(w)+(1-eps))/(1-eps) - image
 *   mode PCFA_BOX_CLIP     : var = image+delta;     x = clamp(var,0,1);  delta = x - image
 *   mode PCFA_BOX_JOINT    : var = delta (shared);  x = clamp(image+delta,0,1);
 *                            delta' = clamp(clamp(delta+other_max,0,1)-other_max+other_min,0,1)-other_min
 *                            (penalty sees delta'; `aux_max`/`aux_min` = max/min(image1,image2))
 *   mode PCFA_BOX_UNIVERSAL: var = delta [1,C,H,W] broadcast over the batch; x = clamp(image+delta,0,1);
 *                            penalty sees the raw delta.
 *   net_in = scale * x  (scale = 255 for RAFT/GMA/FlowNet2, 1 for PWCNet/SpyNet).
 *   sumsq_partials[PCFA_BOX_PARTIALS] receives deterministic per-block partial sums of delta^2
 *   (sum them in index order; pcfa_objective_loss does).
 * ====================================================================================== */
#define PCFA_BOX_COV        0
#define PCFA_BOX_CLIP       1
#define PCFA_BOX_JOINT      2
#define PCFA_BOX_UNIVERSAL  3
#define PCFA_BOX_PARTIALS   1024

int pcfa_box_forward(const float* var, const float* image,
                     const float* aux_max, const float* aux_min,   /* JOINT only, else NULL */
                     float* net_in /* (overwritten) [B,C,H,W] */,
                     float* delta_out /* optional (may be NULL) */,
                     float* sumsq_partials /* (overwritten) [PCFA_BOX_PARTIALS] */,
                     int mode, int B, int64_t chw, float eps_box, float scale,
                     pcfa_stream_t stream);

/* dL/dvar (overwritten; for UNIVERSAL summed over the batch) from dL/dnet_in and the penalty:
 *   penalty_coef = mu * 2 / numel_total if the penalty is active else 0, read from
 *   loss_terms[2] (device; written by pcfa_objective_loss) so no host sync is needed. */
int pcfa_box_backward(const float* var, const float* image,
                      const float* aux_max, const float* aux_min,
                      const float* grad_net_in /* may be NULL: penalty only */,
                      const float* loss_terms,
                      float* grad_var /* overwritten, or accumulated when accumulate != 0 (second
                                         image of a shared perturbation) */,
                      int accumulate,
                      int mode, int B, int64_t chw, float eps_box, float scale,
                      pcfa_stream_t stream);


/* d delta/d var applied to an arbitrary gradient (autograd of extract_deltas / extract_deltas_joint,
 * attack_PCFA.py:20-37, when the deltas feed something other than the fused loss).  UNIVERSAL: identity. */
int pcfa_box_delta_backward(const float* var, const float* aux_max, const float* aux_min,
                            const float* grad_delta, float* grad_var, int accumulate,
                            int mode, int64_t numel, float eps_box, pcfa_stream_t stream);

/* Deterministic partial sums of x^2 (PCFA_BOX_PARTIALS floats) — two_norm_avg_delta_squared's
 * reduction (helper_functions/losses.py:110-126) for deltas that did not come from pcfa_box_forward. */
int pcfa_sumsq_partials(const float* x, int64_t numel, float* sumsq_partials, pcfa_stream_t stream);

/* ======================================================================================
 * (a11/a12) Loss + penalty — replaces losses.loss_delta_constraint (helper_functions/losses.py:
 *      200-230) with avg_epe :3-30 / avg_mse :32-44 / f_cosim :76-88 (including its operator-
 *      precedence quirk), relu_penalty :177-197, and InputPadder.unpad (ownutilities.py:51-62)
 *      without the .cpu() round trip of postprocess_flow (ownutilities.py:297).
 *
 *   flow [B,2,Hp,Wp] is the PADDED network output; target [B,2,H,W] is unpadded; the crop is
 *   rows [pad_top, pad_top+H), cols [pad_left, pad_left+W).
 *   loss_terms (device, 4 floats, overwritten):
 *     [0] = loss = sim + mu*max(0, sumsq/numel - bound^2), [1] = sim, [2] = penalty_coef
 *     (mu*2/numel if active else 0), [3] = sumsq/numel.
 *   grad_flow [B,2,Hp,Wp] (overwritten; zero in the padding) = d sim / d flow.
 *   sumsq_partials: one or two arrays of PCFA_BOX_PARTIALS floats (second may be NULL);
 *   numel_total = numel(delta1)+numel(delta2) as in losses.py:122-126.
 * ====================================================================================== */
#define PCFA_LOSS_AEE   0
#define PCFA_LOSS_MSE   1
#define PCFA_LOSS_COSIM 2

int pcfa_objective_loss(const float* flow, const float* target,
                        const float* sumsq_partials1, const float* sumsq_partials2,
                        float sumsq_weight1, float sumsq_weight2,
                        float* loss_terms, float* grad_flow /* may be NULL */,
                        void* workspace /* >= pcfa_objective_workspace_bytes() */,
                        int loss_type, int B, int H, int W, int Hp, int Wp,
                        int pad_top, int pad_left,
                        double numel_total, float delta_bound, float mu,
                        pcfa_stream_t stream);
int64_t pcfa_objective_workspace_bytes(void);

/* --------------------------------------------------------------------------- encoder glue
 * Instance normalisation with optional fused ReLU on NCHW fp32 planes: y = relu((x - mean) / sqrt(var + eps)), biased
 * variance over H*W per (b, c) — nn.InstanceNorm2d(affine=False) + nn.ReLU as RAFT's feature encoder applies them
 * (models/raft/extractor.py:13-55,118-150; F.instance_norm -> ATen batch_norm on a [1, B*C, H, W] view).
 * `stats` receives (mean, rstd) per plane: [B*C][2] floats, and is an input of the backward call together with the
 * forward INPUT x (the ReLU mask is recomputed from it).  `workspace`: pcfa_instnorm_workspace_bytes().
 * channels_last != 0: all tensors are [B][H][W][C] in memory (torch.channels_last); needs C % 4 == 0, C <= 1024. */
int64_t pcfa_instnorm_workspace_bytes(int B, int C, int H, int W);
int pcfa_instnorm_forward(const float* x, float* y, float* stats, void* workspace, int B, int C, int H, int W, float eps,
                          int relu, int channels_last, pcfa_stream_t stream);
int pcfa_instnorm_backward(const float* x, const float* grad_y, const float* stats, float* grad_x, void* workspace,
                           int B, int C, int H, int W, int relu, int channels_last, pcfa_stream_t stream);
/* Channels-last with x stored in fp16 (GMA's fp16-autocast encoder): y and grad_y fp32, grad_x fp16, fp32 arithmetic — what
 * autocast makes of F.instance_norm(conv_out), without the conversion copies around it. */
int pcfa_instnorm_forward_h(const void* x_half, float* y, float* stats, void* workspace, int B, int C, int H, int W, float eps,
                            int relu, pcfa_stream_t stream);
int pcfa_instnorm_backward_h(const void* x_half, const float* grad_y, const float* stats, void* grad_x_half, void* workspace, int B,
                             int C, int H, int W, int relu, pcfa_stream_t stream);

/* Element-wise halves of the convolutional GRU (models/raft/update.py:16-60): z = sigmoid, r = sigmoid, rh = r*h from the
 * concatenated pre-activations zr = [B][2C][H*W] (z first; n = C*H*W elements per sample), and the state update
 * q = tanh(q_pre), h_new = (1-z)*h + z*q.  All tensors NCHW-contiguous fp32.  grad_z / grad_rh may be NULL (= zero). */
int pcfa_gru_gates_forward(const float* zr, const float* h, float* z, float* r, float* rh, int B, int64_t n,
                           int channels_last_C /* 0: NCHW; C: all tensors channels-last with C (2C for zr) channels */,
                           pcfa_stream_t stream);
int pcfa_gru_gates_backward(const float* z, const float* r, const float* h, const float* grad_z, const float* grad_rh,
                            float* grad_zr, float* grad_h, int B, int64_t n, int channels_last_C, pcfa_stream_t stream);
int pcfa_gru_blend_forward(const float* z, const float* q_pre, const float* h, float* q, float* h_new, int64_t numel,
                           pcfa_stream_t stream);
int pcfa_gru_blend_backward(const float* z, const float* q, const float* h, const float* grad_h_new, float* grad_z,
                            float* grad_q_pre, float* grad_h, int64_t numel, pcfa_stream_t stream);

/* Channels-last GRU halves with a hoisted pre-activation addend and concatenated outputs (NHWC update block):
 *   gates_x : z = sigmoid(zr[:, :C] + addend[:, :C]), r = sigmoid(zr[:, C:] + addend[:, C:]), rhm = [r*h | m]
 *   blend_x : q = tanh(q_pre + addend), h_new = (1-z)*h + z*q, hm = [h_new | m] (hm may be NULL)
 * All tensors [npix][channels] (torch.channels_last), C and Cm multiples of 4, 16-byte aligned.  In the backward calls
 * grad_rhm / grad_hm are the gradients of the concatenated outputs (their first C channels are consumed here, the tail
 * is the caller's gradient of m); NULL gradient pointers mean zero. */
int pcfa_gru_gates_x_forward(const float* zr, const float* addend, const float* h, const float* m, float* z, float* r, float* rhm,
                             int C, int Cm, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_gates_x_backward(const float* z, const float* r, const float* h, const float* grad_z, const float* grad_rhm,
                              float* grad_zr, float* grad_h, int C, int Cm, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_blend_x_forward(const float* z, const float* q_pre, const float* addend, const float* h, const float* m, float* q,
                             float* h_new, float* hm, int C, int Cm, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_blend_x_backward(const float* z, const float* q, const float* h, const float* grad_h_new, const float* grad_hm,
                              float* grad_z, float* grad_q_pre, float* grad_h, int C, int Cm, int64_t npix, pcfa_stream_t stream);

/* Whole-step variants for one autograd node per SepConvGRU step (pcfa_b200/gru_ops.py::gru_step_x):
 *   *_backward_acc : as above, with a second dense addend for grad_h_new and an accumulator for the gradient of the hoisted
 *                    addend (acc_mode 0 none, 1 acc = grad, 2 acc += grad; summed over the GRU iterations in place);
 *   pcfa_gru_step_combine : grad_h = gh_a + gh_b + cat0[:, :C];  grad_m = sum_k cat_k[:, C:]  over the four [npix][C+Cm]
 *                    gradients of the step's concatenated convolution inputs — every gradient is written once. */
int pcfa_gru_gates_x_backward_acc(const float* z, const float* r, const float* h, const float* grad_z, const float* grad_rhm,
                                  float* grad_zr, float* grad_h, float* acc, int acc_mode, int C, int Cm, int64_t npix,
                                  pcfa_stream_t stream);
int pcfa_gru_blend_x_backward_acc(const float* z, const float* q, const float* h, const float* grad_h_new_a, const float* grad_h_new_b,
                                  const float* grad_hm, float* grad_z, float* grad_q_pre, float* grad_h, float* acc, int acc_mode,
                                  int C, int Cm, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_step_combine(const float* gh_a, const float* gh_b, const float* cat0, const float* cat1, const float* cat2,
                          const float* cat3, float* grad_h, float* grad_m, int C, int Cm, int64_t npix, pcfa_stream_t stream);

/* fp16-storage variants of the whole-step GRU kernels (GMA under fp16 autocast; csrc/gru_half.cu): same argument meaning as the
 * fp32 entry points above, every tensor __half channels-last, fp32 arithmetic per element, 8-byte aligned pointers. */
int pcfa_cat2_channels_last_h(const void* a, const void* b, void* out, int Ca, int Cb, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_gates_x_forward_h(const void* zr, const void* addend, const void* h, const void* m, void* z, void* r, void* rhm, int C, int Cm,
                               int64_t npix, pcfa_stream_t stream);
int pcfa_gru_blend_x_forward_h(const void* z, const void* q_pre, const void* addend, const void* h, const void* m, void* q, void* h_new,
                               void* hm, int C, int Cm, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_gates_x_backward_acc_h(const void* z, const void* r, const void* h, const void* grad_z, const void* grad_rhm, void* grad_zr,
                                    void* grad_h, void* acc, int acc_mode, int C, int Cm, int64_t npix, pcfa_stream_t stream);
int pcfa_gru_blend_x_backward_acc_h(const void* z, const void* q, const void* h, const void* grad_h_new_a, const void* grad_h_new_b,
                                    const void* grad_hm, void* grad_z, void* grad_q_pre, void* grad_h, void* acc, int acc_mode, int C, int Cm,
                                    int64_t npix, pcfa_stream_t stream);
int pcfa_gru_step_combine_h(const void* gh_a, const void* gh_b, const void* cat0, const void* cat1, const void* cat2, const void* cat3,
                            void* grad_h, void* grad_m, int C, int Cm, int64_t npix, pcfa_stream_t stream);

/* Channel concatenation of up to four channels-last tensors [npix][C_k] -> [npix][sum C_k] (torch.cat(dim=1) of
 * torch.channels_last tensors, which ATen runs on a slow path). */
int pcfa_cat_channels_last(const float* const* inputs, const int* channels, int n_inputs, float* out, int64_t npix,
                           pcfa_stream_t stream);
/* Same for arbitrary channel counts, with zero channels appended up to out_channels (a multiple of 4, >= the sum): the
 * consumer is a convolution with zero-padded input-channel weights, so cuDNN needs no channel-padding launches. */
int pcfa_cat_channels_last_pad(const float* const* inputs, const int* channels, int n_inputs, float* out, int64_t npix,
                               int out_channels, pcfa_stream_t stream);

/* --------------------------------------------------------------------------- convex up-sampling
 * RAFT.upsample_flow (models/raft/raft.py:72-83; GMA: models/gma/network.py:59-70): softmax over the 9 taps of the
 * mask head's 576 channels, convex combination of the 3x3 neighbourhood of 8*flow, pixel-shuffle to [N,2,8H,8W].
 *   flow        [N,2,H,W]    fp32 NCHW
 *   mask_cl     [N,H,W,576]  fp32, the mask head's output in torch.channels_last memory; channel = k*64 + i*8 + j
 *   mask_scale  multiplies the mask before the softmax (0.25 in models/raft/update.py:135) and its gradient
 *   up          [N,2,8H,8W]  fp32 NCHW
 * backward: grad_mask_cl has mask_cl's layout; workspace >= pcfa_convex_upsample_workspace_bytes(N,H,W).  No atomics. */
int64_t pcfa_convex_upsample_workspace_bytes(int N, int H, int W);
int pcfa_convex_upsample_forward(const float* flow, const float* mask_cl, float* up, int N, int H, int W, float mask_scale,
                                 pcfa_stream_t stream);
int pcfa_convex_upsample_backward(const float* flow, const float* mask_cl, const float* grad_up, float* grad_flow,
                                  float* grad_mask_cl, void* workspace, int64_t workspace_bytes, int N, int H, int W,
                                  float mask_scale, pcfa_stream_t stream);

/* --------------------------------------------------------------------------- GMA attention softmax
 * attn = softmax(sim, dim=-1) of models/gma/gma.py:73-74 for the shipped fp16-autocast configuration: fp16 rows in,
 * fp16 rows out, fp32 arithmetic (what the reference's fp32 softmax followed by the aggregation GEMM's fp16 cast yields).
 *   sim, attn, grad_*: [rows][cols] __half, 16-byte aligned, cols % 8 == 0 and cols <= 16384 (else PCFA_E_BADARG).
 * backward: grad_sim = attn * (grad_attn - sum_j attn_j * grad_attn_j). */
int pcfa_softmax_rows_f16_forward(const void* sim, void* attn, int64_t rows, int cols, pcfa_stream_t stream);
int pcfa_softmax_rows_f16_backward(const void* attn, const void* grad_attn, void* grad_sim, int64_t rows, int cols,
                                   pcfa_stream_t stream);

/* --------------------------------------------------------------------------- convolution epilogue (row f-4, glue)
 * y = act?(x + bias[c]) in place on the output of a bias-free cuDNN convolution (act = max(v, slope*v): ReLU for slope 0,
 * LeakyReLU otherwise; 0 <= slope < 1), and the activation mask of the backward (y > 0 ? g : slope*g)
 * (torch runs the bias as a broadcasting ATen add and the ReLU as a clamp: two launches per convolution).
 * dtype 0 = fp32, 1 = fp16.  Layout by `inner`: 1 = channels-last ([..., C] innermost; C % 4 (8 for fp16) == 0),
 * H*W = NCHW (inner % 4 (8) == 0).  n % 4 (8) == 0; pointers 16-byte aligned; otherwise PCFA_E_BADARG (caller falls back). */
int pcfa_bias_act_forward(void* x, const void* bias, int64_t n, int C, int64_t inner, int relu, float slope, int dtype,
                          pcfa_stream_t stream);
int pcfa_relu_mask_backward(const void* y, const void* grad_y, void* grad_x, int64_t n, float slope, int dtype, pcfa_stream_t stream);
/* grad_y given as rows `ld` elements apart (a channel slice of a wider channels-last tensor, i.e. the gradient of a
 * concatenation): saves the .contiguous() copy.  y, grad_x: dense [rows][C]; C % 4 (8) == 0, ld % 4 (8) == 0. */
int pcfa_relu_mask_backward_rows(const void* y, const void* grad_y, void* grad_x, int64_t rows, int C, int64_t ld, float slope,
                                 int dtype, pcfa_stream_t stream);
/* dst[r][c] += src[r*ld + c] (fp32): a dense [rows][C] gradient plus a channel slice of a wider channels-last gradient — the
 * skip branch of x = cat(conv(x), x) (models/PWCNet/PWCNet.py:253-257 and the four levels below it).  C % 4 == 0, ld % 4 == 0. */
int pcfa_add_rows_inplace(float* dst, const float* src, int64_t rows, int C, int64_t ld, pcfa_stream_t stream);
/* out = a + src-slice, out-of-place: where a tensor feeds a convolution AND a later concatenation (the encoder skips of
 * models/FlowNet/FlowNetS.py:63-88, FlowNetC.py:106-121, FlowNetSD.py:69-99, FlowNetFusion.py:50-65). */
int pcfa_add_rows(float* out, const float* a, const float* src, int64_t rows, int C, int64_t ld, pcfa_stream_t stream);
/* grad_x = y > 0 ? g1 + g2 : 0 (fp32, any dense layout shared by the four tensors): the ReLU mask of a residual tail applied
 * to the sum of the gradients of its two consumers (next block's convolution and skip branch) in one pass. */
int pcfa_relu_mask2_backward(const float* y, const float* g1, const float* g2, float* grad_x, int64_t n, pcfa_stream_t stream);
/* out = relu(a + b), element-wise, any dense layout shared by the three tensors: the tail of the encoders' residual blocks
 * (models/raft/extractor.py:56,116) in one pass.  dtype 0 = fp32 (n % 4 == 0), 1 = fp16 (n % 8 == 0). */
int pcfa_add_relu_forward(const void* a, const void* b, void* out, int64_t n, int dtype, pcfa_stream_t stream);
/* Coordinate bookkeeping of one RAFT/GMA refinement iteration in one launch (models/raft/raft.py:123-131):
 * new_coords1 = coords1 + delta[..., 0:2], flow_cl = new_coords1 - coords0.  coords*: [B,2,H,W]; delta: channels-last with
 * delta_ld (even) channels per pixel; flow_cl: [B,H,W,flow_ld] = torch.channels_last memory of [B,flow_ld,H,W], channels 2.. are
 * zero-filled (flow_ld = 8 lets the 7x7 convolution on the flow run as a tensor-core implicit GEMM). */
int pcfa_flow_step(const float* coords1, const float* coords0, const void* delta, int delta_ld, int delta_dtype /* 0 fp32, 1 fp16 */,
                   float* new_coords1, float* flow_cl, int flow_ld, int B, int H, int W, pcfa_stream_t stream);


/* --------------------------------------------------------------------------- on-device L-BFGS
 * The vector algebra of torch.optim.LBFGS.step (torch/optim/lbfgs.py; the reference's optimiser, attack_PCFA.py:97,114)
 * without its ~4*history ATen launches per iteration.  History: ring buffers S, Y of [history_capacity][n] floats.
 *   pcfa_lbfgs_store_pair : y = grad - grad_prev -> y_slot, s = t*d -> s_slot, grad_prev <- grad;
 *                           scalars_out = { <y,s>, <y,y> } (device floats)
 *   pcfa_lbfgs_direction  : two-loop recursion over the `num_old` pairs starting at ring index `start` (oldest first),
 *                           ro[slot] = 1/<y,s>, *h_diag = <y,s>/<y,y> of the newest pair (device floats);
 *                           d <- -H*grad;  scalars_out = { <grad,d>, max|d| }.  One cooperative launch.
 * workspace: pcfa_lbfgs_workspace_bytes(). */
int64_t pcfa_lbfgs_workspace_bytes(void);
/* One host round trip per iteration (the f-1 criterion): the history bookkeeping and torch's pre-update break test live on
 * the device.
 *   pcfa_lbfgs_update_history : <y,s>, <y,y> of the candidate pair; if <y,s> > 1e-10 it replaces the oldest / next free
 *                               ring slot, ro[slot] = 1/<y,s>, *h_diag = <y,s>/<y,y>; ring_out = new {start, num_old}
 *                               (ring_in != ring_out, both int[2] on the device); always grad_prev <- grad.
 *   pcfa_lbfgs_direction_step : two-loop recursion over ring = {start, num_old}, then param += t*d unless
 *                               <grad,d> > -tol_change; scalars_out = { <grad,d>, max|d| }. */
int pcfa_lbfgs_update_history(const float* grad, float* grad_prev, const float* d, float t, float* S, float* Y, float* ro,
                              float* h_diag, const int* ring_in, int* ring_out, float* scalars_out, void* workspace, int64_t n,
                              int history_capacity, pcfa_stream_t stream);
int pcfa_lbfgs_direction_step(const float* S, const float* Y, const float* ro, const float* grad, const float* h_diag, float* d,
                              const int* ring, float* param, float t, float tol_change, float* scalars_out, void* workspace,
                              int64_t n, int history_capacity, pcfa_stream_t stream);
/* The same direction in the compact (Byrd-Nocedal-Schnabel) representation: three launches without a grid barrier instead
 * of one cooperative launch with 2*num_old of them (csrc/lbfgs_compact.cu).  pair_scalars = the {<y,s>, <y,y>} that
 * pcfa_lbfgs_update_history wrote; `state` (pcfa_lbfgs_compact_workspace_bytes(history_capacity), zero-initialised, 8-byte
 * aligned) carries S^T Y, Y^T Y and the previous S^T grad, Y^T grad between calls. */
int64_t pcfa_lbfgs_compact_workspace_bytes(int history_capacity);
int pcfa_lbfgs_direction_compact(const float* S, const float* Y, const float* grad, const float* h_diag, float* d, const int* ring,
                                 const float* pair_scalars, float* param, float t, float tol_change, float* scalars_out,
                                 void* state, int64_t n, int history_capacity, pcfa_stream_t stream);
int pcfa_lbfgs_store_pair(const float* grad, float* grad_prev, const float* d, float t, float* s_slot, float* y_slot,
                          float* scalars_out, void* workspace, int64_t n, pcfa_stream_t stream);
int pcfa_lbfgs_direction(const float* S, const float* Y, const float* ro, const float* grad, const float* h_diag, float* d,
                         float* scalars_out, void* workspace, int64_t n, int history_capacity, int start, int num_old,
                         pcfa_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* PCFA_B200_H */
