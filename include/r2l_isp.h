/*
 * r2l_isp.h -- C ABI of the B200-native differentiable ISP (raw Bayer -> RGB, forward + backward).
 *
 * This is the drop-in boundary for the hot path of aiaudit-org/raw2logit's processing/pipeline_torch.py.
 * The reference has no FFI (it is pure Python over stock torch ops); each entry point below names the reference
 * interface it replaces (file:line relative to the reference tree).  The host side that binds these symbols is
 * raw2logit_b200/_lib.py (ctypes); the torch custom ops and the autograd.Function sit above it
 * (raw2logit_b200/ops.py); INTEGRATION.md shows the stub a reference maintainer would add.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer on the current CUDA device unless noted;
 *   - all buffers are owned by the caller (outputs, workspace); the library keeps no state between calls;
 *   - `stream` is a cudaStream_t passed as void*; calls only enqueue work, they never synchronise;
 *   - re-entrant from several host threads (autograd's backward thread calls r2l_isp_backward);
 *   - return value: R2L_OK (0) or a negative R2L_ERR_* code; nothing throws across the boundary.
 *
 * Layouts (all row-major, contiguous):
 *   raw      (B, H, W)     float32, or uint16 with value = u / raw_denominator (dataset.py:87: img/(2**bits-1))
 *   out      (B, 3, H, W)  float32                                  (pipeline_torch.py:175-225 return value)
 *   grad_out (B, 3, H, W)  float32, grad_raw (B, H, W) float32
 *   RGGB phase of a site: par(y,x) = 2*(y&1) + (x&1) -> R, G1, G2, B  (pipeline_torch.py:256-259)
 */
#ifndef R2L_ISP_H
#define R2L_ISP_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define R2L_ABI_VERSION 6

enum {
    R2L_OK = 0,
    R2L_ERR_BAD_SHAPE = -1,     /* B < 0, or H < 3 or W < 3 (reference: reflect pad 2 raises, pipeline_torch.py:165,202) */
    R2L_ERR_BAD_DTYPE = -2,     /* raw_dtype not in {R2L_F32, R2L_U16} */
    R2L_ERR_NULL_POINTER = -3,  /* a required pointer is NULL */
    R2L_ERR_MISALIGNED = -4,    /* out/grad pointers not 4-byte aligned, raw not element aligned */
    R2L_ERR_WORKSPACE = -5,     /* workspace too small (see r2l_isp_workspace_bytes) */
    R2L_ERR_CUDA = -6,          /* a CUDA runtime call failed; r2l_isp_last_cuda_error() has the code */
    R2L_ERR_BAD_ARGUMENT = -7   /* inconsistent flags / unsupported mode */
};

enum { R2L_F32 = 0, R2L_U16 = 1 };

/* The 132 trainable scalars + the two fixed colour-space matrices of ParametrizedProcessing
 * (pipeline_torch.py:154-171).  Device pointers to float32, each tensor contiguous in its torch layout. */
typedef struct r2l_isp_params {
    const float* black_level;        /* (4,)        :154 */
    const float* white_balance;      /* (1,3)       :155 */
    const float* colour_correction;  /* (3,3) [k][c] :156 */
    const float* gamma_correct;      /* (1,)        :158 */
    const float* debayer_weight;     /* (3,3,3,3) [k][c][i][j]  Debayer :228-237 */
    const float* sharpen_weight;     /* (1,1,3,3)   :162-163 */
    const float* gauss_weight;       /* (1,1,5,5)   :165-166 */
    const float* rgb2yuv;            /* (3,3) buffer M_RGB_2_YUV :170 */
    const float* yuv2rgb;            /* (3,3) buffer M_YUV_2_RGB :171 */
} r2l_isp_params;

/* Offsets of each parameter group inside the flat 132-float gradient vector written by r2l_isp_backward. */
enum {
    R2L_G_BLACK_LEVEL = 0,    /* 4  */
    R2L_G_WHITE_BALANCE = 4,  /* 3  */
    R2L_G_COLOUR = 7,         /* 9  */
    R2L_G_GAMMA = 16,         /* 1  */
    R2L_G_DEBAYER = 17,       /* 81 */
    R2L_G_SHARPEN = 98,       /* 9  */
    R2L_G_GAUSS = 107,        /* 25 */
    R2L_NUM_PARAM_GRADS = 132
};

/* Optional fused tail of the forward (both may be NULL):
 *   additive  (3,H,W) float32 broadcast over B       -- additive_layer, pipeline_torch.py:129-131, 212-214
 *   affine    6 floats {scale[3], shift[3]}: out = o*scale[c] + shift[c]
 *             -- BatchNorm2d(3, affine=False) in eval mode folded to scale/shift, pipeline_torch.py:168, 216-217 */
typedef struct r2l_isp_tail {
    const float* additive;
    const float* affine;
} r2l_isp_tail;

int r2l_isp_abi_version(void);
const char* r2l_isp_error_string(int code);
int r2l_isp_last_cuda_error(void);

/* Fused forward: replaces ParametrizedProcessing.forward (pipeline_torch.py:175-225) minus stage tracking:
 * raw2rgb :183 -> Debayer :187 -> WB :190 -> CCM :191 -> RGB2YUV :194 -> sharpen(Y) :195 -> Gaussian(Y) :202
 * -> YUV2RGB :203 -> clip :206 -> gamma :209 [-> additive :213] [-> eval-BN :217].  tail may be NULL. */
/* saved_luma (NULL, or r2l_isp_saved_luma_floats(B,H,W) floats, 32-byte aligned: written / read with 256-bit accesses): when given, the kernel also
 * keeps the two luma planes it computes on the way -- Y0 (after RGB->YUV, :194) and Y1 (after the sharpening filter,
 * :195) -- laid out [2][ceil(B/2)][H][W][2] (images 2p, 2p+1 interleaved per site; an odd last image is paired with
 * itself).  They are what torch autograd would keep for the two convolutions' weight gradients; handing them to
 * r2l_isp_backward lets it skip every recompute.  Only the vectorised kernel writes them: ask
 * r2l_isp_luma_supported() first (R2L_ERR_BAD_ARGUMENT otherwise). */
int r2l_isp_forward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                    const r2l_isp_params* params, const r2l_isp_tail* tail, float* out, float* saved_luma,
                    void* stream);
size_t r2l_isp_saved_luma_floats(int B, int H, int W);
/* 1 when a forward call with this raw / out / additive (pointer alignment, shape: W % 4 == 0, 16-byte aligned rows)
 * takes the kernel that can write saved_luma, else 0 (then pass saved_luma = NULL to forward and backward). */
int r2l_isp_luma_supported(const void* raw, int raw_dtype, int B, int H, int W, const float* out,
                           const float* additive);

/* Workspace every call below accepts (a fixed upper bound, independent of the shape). */
size_t r2l_isp_workspace_bytes(int B, int H, int W);

/* Fused forward + train-mode BatchNorm2d(3, affine=False) tail (pipeline_torch.py:168, 216-217; train.py:196
 * always enables it): forward kernel that also reduces per-channel sum / sum of squares, a finish kernel that
 * forms batch mean / biased variance, writes saved_affine = {1/sqrt(var+eps)[3], -mean/sqrt(var+eps)[3]} and
 * updates running_mean / running_var in place like torch (momentum, unbiased variance; either may be NULL), and
 * an in-place normalisation of out.  additive may be NULL.  num_batches_tracked (may be NULL): one int64 on the
 * device, incremented by one (nn.BatchNorm2d's counter; in the kernel, so a captured launch counts its replays). */
int r2l_isp_forward_bn_train(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                             const r2l_isp_params* params, const float* additive, float* out,
                             float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                             float eps, float* saved_affine, float* saved_luma, void* workspace, size_t workspace_bytes,
                             void* stream);

/* Backward of that tail, part 1: reduces sum(grad_out), sum(grad_out * out) per channel and writes the 15-float
 * grad_tail {gs[3], c1[3], c2[3], ysc[3], ysh[3]} that r2l_isp_backward consumes.
 * With the full workspace (r2l_isp_workspace_bytes) the tail is DEFERRED: the per-CTA sums stay in the workspace, c1 / c2
 * of grad_tail carry a tag, and the next r2l_isp_backward / r2l_isp_backward_dp call on the same stream WITH THE SAME
 * WORKSPACE finishes them inside its kernel (no separate finish launch); treat grad_tail as opaque between the two
 * calls.  With a smaller workspace (>= 3 * 296 * 2 floats) grad_tail is complete when this call's work is done. */
int r2l_isp_bn_backward_prepare(const float* grad_out, const float* out, const float* saved_affine, int B, int H,
                                int W, float* grad_tail, void* workspace, size_t workspace_bytes, void* stream);

/* Fused backward: replaces the autograd graph of the same chain (79 nodes).
 * grad_out is dL/d(output).  grad_tail (NULL or 15 floats {gs, c1, c2, ysc, ysh} x 3 channels) describes the
 * BatchNorm tail the forward applied: dL/d(o) = gs*(grad_out - c1 - c2*yhat), yhat = (o + additive)*ysc + ysh,
 * formed inside the kernel (eval mode: gs = 1/sqrt(running_var+eps), c1 = c2 = 0).  additive (NULL or (3,H,W))
 * is only read for yhat.  grad_raw may be NULL (the training case: raw does not require grad).  grad_params
 * receives R2L_NUM_PARAM_GRADS floats laid out per the R2L_G_* offsets.
 * out (NULL or the (B,3,H,W) output the forward produced for the same raw / params / tail -- what torch autograd
 * keeps alive anyway): when given, the kernel derives the clip mask and the gamma derivative from it instead of
 * recomputing the Gaussian and the colour tail (about 20 % fewer instructions for 12 B/px more reads; the chip has
 * the bandwidth to spare).  With a tail, grad_tail's ysc/ysh and additive must describe the
 * affine map that produced out.  NULL = full recompute from raw.
 * saved_luma (NULL or the planes the forward saved for the same call; only read together with out): the kernel then
 * recomputes nothing -- no raw window, no Y0 / Y1 stencils -- and reads the three centre values its statistics need
 * (Y1, Y0, raw) straight from memory; the 132 gradients are finished by the last CTA of the same launch. */
int r2l_isp_backward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                     const r2l_isp_params* params, const float* grad_out, const float* grad_tail,
                     const float* additive, const float* out, const float* saved_luma, float* grad_raw,
                     float* grad_params, void* workspace, size_t workspace_bytes, void* stream);

/* Data-parallel variant (new build: the reference is single-GPU, train.py:362): the same backward, and the
 * 132 parameter gradients leave the call ALREADY SUMMED OVER ALL RANKS.  The last CTA of the launch -- the one that
 * turns the per-CTA sums into the gradients -- writes them, each as an 8-byte {value, epoch} word, into its slot of every
 * rank's exchange buffer over NVLink peer memory, polls the other ranks' slots in its own buffer until they carry the
 * epoch and adds them up in rank order (bit-identical result on every rank): a one-shot all-reduce of 528 bytes inside
 * the kernel, no fence, no flag, no collective launch.  grad_raw stays local (images are independent).
 *   peers     DEVICE array of `world` pointers; peers[r] = rank r's exchange buffer as mapped into THIS process
 *             (CUDA IPC / symmetric memory, e.g. torch.distributed._symmetric_memory: buffer_ptrs_dev), each
 *             r2l_isp_exchange_bytes(world) bytes, zero-filled once before the first call
 *   epoch     1, 2, 3, ... -- the same value on every rank, incremented on every call that uses the buffer; or
 *             R2L_EPOCH_DEVICE: the kernel keeps the count itself, in a word at the end of this rank's exchange buffer
 *             (every rank makes the same calls, so the counts agree) -- the launch then carries no per-call value and
 *             can be captured in a CUDA graph and replayed.  One buffer must be used in one of the two ways only.
 *   scale     factor applied to the sum (1/world for an average, 1 for a sum)
 * Needs the path that finishes the gradients in the launch (out and saved_luma given, shape served by the vectorised
 * kernel); otherwise R2L_ERR_BAD_ARGUMENT and nothing is launched (use r2l_isp_backward + a collective).  Every rank
 * must make the call, or the others wait for ever, as with any collective. */
#define R2L_EPOCH_DEVICE 0xFFFFFFFFu
typedef struct {
    int world, rank;
    float* const* peers;
    unsigned epoch;
    float scale;
} r2l_isp_allreduce;
size_t r2l_isp_exchange_bytes(int world);
int r2l_isp_backward_dp(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                        const r2l_isp_params* params, const float* grad_out, const float* grad_tail,
                        const float* additive, const float* out, const float* saved_luma, float* grad_raw,
                        float* grad_params, void* workspace, size_t workspace_bytes, const r2l_isp_allreduce* dp,
                        void* stream);

/* out[c][i] = scale[c] * sum_b x[b][c][i]  (scale may be NULL): gradient of the broadcast additive_layer
 * (pipeline_torch.py:212-214), x = grad_out (B,C,HW), out (C,HW).  Deterministic (fixed summation order). */
int r2l_isp_batch_sum(const float* x, const float* scale, int B, int C, int HW, float* out, void* stream);

/* CFA split, replaces raw2rgb (pipeline_torch.py:240-283) / RawToRGB.forward (:65-80).
 * reduce_size=1: out (B,C,H/2,W/2) packed (C=3 averages the greens), needs even H,W (the reference raises
 * otherwise); reduce_size=0: out (B,C,H,W) zero-filled mosaic.  black_level: NULL or 4 floats (device). */
int r2l_isp_mosaic(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                   const float* black_level, int reduce_size, int out_channels, float* out, void* stream);
/* Its adjoint with respect to raw: grad_raw (B,H,W) from grad_out in the layout above. */
int r2l_isp_mosaic_backward(const float* grad_out, int B, int H, int W, int reduce_size, int out_channels,
                            float* grad_raw, void* stream);

/* ---- SSIM regulariser of adversarial training ------------------------------------------------------
 * Replaces utils/ssim.py:19-39 (`_ssim`: five grouped 11x11 Gaussian-window convolutions with zero padding 5 on
 * img1, img2, img1^2, img2^2, img1*img2, the SSIM map and its mean), called as SSIM(window_size=11) by train.py:261-262
 * through AuxLoss (utils/base.py:346-358).  img1 / img2: (B, C, H, W) float32, contiguous.  window_size must be 11
 * (sigma 1.5, the reference's only use); R2L_ERR_BAD_ARGUMENT otherwise.
 * forward: partial receives r2l_isp_ssim_partial_count() doubles, the sums of the SSIM map over the 32x32 tiles laid
 *   out [B][C][tiles]: their total / (B*C*H*W) is `ssim_map.mean()` (:36), the per-image totals / (C*H*W) the
 *   size_average=False result (:38).
 * backward: grad1 / grad2 (either may be NULL) receive d(result)/d(img1) / d(img2) for an upstream gradient given as
 *   scale[b] = grad_result[b] / (number of averaged elements), B floats on the device.  Nothing is kept between the
 *   two calls: the backward recomputes the window moments. */
size_t r2l_isp_ssim_partial_count(int B, int C, int H, int W);
int r2l_isp_ssim_forward(const float* img1, const float* img2, int B, int C, int H, int W, int window_size,
                         double* partial, void* stream);
int r2l_isp_ssim_backward(const float* img1, const float* img2, const float* scale, int B, int C, int H, int W,
                          int window_size, float* grad1, float* grad2, void* stream);

/* ---- augmentation + hand-off to the task model in one pass -----------------------------------------
 * Replaces the chain  augmentation_weak (utils/augmentation.py:70-74: RandomHorizontalFlip, RandomVerticalFlip,
 * RandomRotate90 :8-11; applied at model.py:79-82)  ->  .contiguous(memory_format=channels_last)  ->  .to(bfloat16):
 * dst[b][c][y][x] = src[b][c][a0 + a1 y + a2 x][b0 + b1 y + b2 x]  for a dihedral map6 = {a0, a1, a2, b0, b1, b2} (the
 * composition of the drawn flips / quarter turns, formed by the host), both tensors addressed through element strides
 * {batch, channel, row, column} (NCHW, channels_last, ...) and dtype codes 0 = float32, 2 = bfloat16 (round to nearest
 * even, like Tensor.to).  The adjoint is the same call with the inverse map and the gradient as source.  C <= 4. */
int r2l_isp_dihedral_copy(const void* src, int src_dtype, const long long* src_strides, void* dst, int dst_dtype,
                          const long long* dst_strides, int B, int C, int H_dst, int W_dst, int H_src, int W_src,
                          const int* map6, void* stream);

/* ---- numpy-compatible static pipeline --------------------------------------------------------------
 * Replaces processing/pipeline_numpy.py:70-141 `processing(img, black_level, white_balance, colour_matrix,
 * debayer='bilinear', sharpening=..., denoising=..., gamma=2.2)` -- the per-image chain of --processing_mode static
 * (RawProcessingPipeline.__call__, :56-68; 16 DataLoader workers, train.py:316-320) -- for a whole batch in one kernel,
 * with that chain's boundary rules (scipy half-sample reflection for the bilinear demosaic and the Gaussian filter, zero
 * fill for the sharpening filter) and its clip to [0, 1].  Forward only.
 *   black_level[4], white_balance[3], colour_matrix[9]: HOST arrays (camera constants, not trainable here)
 *   sharpening_filter: 1 = the 3x3 sharpening filter of :178-191, 0 = none
 *   gaussian_denoising: 1 = scipy.ndimage.gaussian_filter(Y, gaussian_sigma) (:203-209; radius int(4 sigma + 0.5) <= 2), 0 = none
 *   out (B, 3, H, W) float32. */
int r2l_isp_numpy_forward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                          const float* black_level, const float* white_balance, const float* colour_matrix,
                          int sharpening_filter, int gaussian_denoising, float gaussian_sigma, float gamma,
                          float* out, void* stream);

/* ---- staged mode: one kernel per stage of ParametrizedProcessing.forward with track_stages=True ----------------------
 * Replaces the per-stage torch ops of pipeline_torch.py:183-214 whose outputs the reference keeps in `self.stages`
 * (and whose .grad model.track_images reads, model.py:229-254).  Every linear stage is a 3 -> 3 channel K x K
 * correlation of the previous stage's tensor (colour stage :187-191 = CCM.diag(wb).Debayer, K = 3, reflect-1 pad;
 * sharpening :194-198 = M_yuv2rgb.[sharpen(Y)|U|V].M_rgb2yuv, K = 3, zero pad; Gaussian :199-203, K = 5, reflect-2 pad);
 * the host forms the combined weight [3][3][K][K] (raw2logit_b200/staged.py).
 *   r2l_isp_stage_conv:           y[b][co] = sum_ci corr(x_pad[b][ci], weight[co][ci]);  x, y (B, 3, H, W) float32,
 *                                 K in {3, 5}, pad_mode 0 = zeros, 1 = reflect (R2L_ERR_BAD_SHAPE when H or W <= K/2)
 *   r2l_isp_stage_conv_backward:  grad_x (nullable) and grad_weight (nullable, 9 K K floats; needs x and a workspace of
 *                                 r2l_isp_stage_workspace_bytes(K) bytes, 8-byte aligned); deterministic two-stage sums
 *   r2l_isp_stage_clip(_backward):  y = clamp(x, lo, hi) (:206); grad_x = grad_y where lo <= x <= hi (torch.clip)
 *   r2l_isp_stage_gamma(_backward): y = exp((1 / gamma) * log(x)) (:209), gamma = one float on the device;
 *                                 grad_x (nullable) = grad_y y / (gamma x), grad_gamma[0] = -sum(grad_y y log x) / gamma^2;
 *                                 workspace as above (any K). */
size_t r2l_isp_stage_workspace_bytes(int K);
int r2l_isp_stage_conv(const float* x, const float* weight, int B, int H, int W, int K, int pad_mode, float* y, void* stream);
int r2l_isp_stage_conv_backward(const float* x, const float* weight, const float* grad_y, int B, int H, int W, int K,
                                int pad_mode, float* grad_x, float* grad_weight, void* workspace, size_t workspace_bytes,
                                void* stream);
int r2l_isp_stage_clip(const float* x, long long n, float lo, float hi, float* y, void* stream);
int r2l_isp_stage_clip_backward(const float* x, const float* grad_y, long long n, float lo, float hi, float* grad_x,
                                void* stream);
int r2l_isp_stage_gamma(const float* x, const float* gamma, long long n, float* y, void* stream);
int r2l_isp_stage_gamma_backward(const float* x, const float* y, const float* grad_y, const float* gamma, long long n,
                                 float* grad_x, float* grad_gamma, void* workspace, size_t workspace_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* R2L_ISP_H */
