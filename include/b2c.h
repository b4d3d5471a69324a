/* b2c.h - C ABI of libb2c.so: the B200 (sm_100a) kernels underneath the reference's
 * ImageCompressor Python API (Balle-2017 codec hot path).
 *
 * The reference (Yuval-H/iclr_17_compression) is pure Python over PyTorch eager and has
 * NO native interface of its own; each entry point below therefore cites the reference
 * Python call site (file:line, relative to the reference root) whose arithmetic it
 * replaces.  INTEGRATION.md shows the ctypes binding a maintainer adds on the reference
 * side.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes; no torch / C++ types.
 *   - every pointer is DEVICE memory owned by the caller unless the name says host;
 *     tensors are dense fp32 NCHW exactly as the reference holds them.
 *   - the library never allocates or frees caller-visible memory and never synchronises
 *     the device; work is enqueued on `stream` (a cudaStream_t passed as void*).
 *   - returns 0 on success, a negative B2C_E* code otherwise; b2c_last_error() returns a
 *     thread-local message.  No C++ exception crosses the boundary.
 *   - re-entrant: no global mutable state apart from lazily created, mutex-guarded,
 *     per-device caches (function attributes, tensor maps).
 *   - there is no CPU fallback: a non-sm_100 device is an error (B2C_EARCH).
 */
#ifndef B2C_H_
#define B2C_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define B2C_OK 0
#define B2C_EINVAL (-1)  /* bad shape / null pointer / unsupported geometry */
#define B2C_EARCH (-2)   /* device is not sm_100 */
#define B2C_ECUDA (-3)   /* CUDA runtime / launch failure */
#define B2C_EWORKSPACE (-4) /* workspace too small */

/* number of per-channel parameter vectors of the factorized entropy model:
 * f1.{h,b,a}, f2.{h,b,a}, f3.{h,b,a}, f4.{h,b}  (models/bitEstimator.py:13-18,33-36) */
#define B2C_NUM_BITPARAMS 11

int b2c_version(void);
const char* b2c_last_error(void);
/* 0 if the CURRENT device is compute capability 10.x, B2C_EARCH otherwise. */
int b2c_check_device(void);
/* counts kernels launched by this library in this process (bench.py's gpu_launches). */
uint64_t b2c_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Convolutions (exact-fp32 path).
 * --------------------------------------------------------------------------------- */
/* nn.Conv2d forward: models/analysis_17.py:14,18,22 + :33-35.
 * x[B,Cin,H,W], w[Cout,Cin,k,k] (OIHW), bias[Cout] or NULL -> y[B,Cout,Ho,Wo],
 * Ho = (H+2p-k)/s+1.  Also serves as dgrad of conv_transpose2d (same weight memory). */
int b2c_conv2d_fwd(const float* x, const float* w, const float* bias, float* y,
                   int B, int Cin, int H, int W, int Cout, int k, int stride, int pad,
                   void* stream);

/* nn.ConvTranspose2d forward: models/synthesis_17.py:15,19,23 + :28-30.
 * x[B,Cin,H,W], w[Cin,Cout,k,k] (IOHW), bias[Cout] or NULL -> y[B,Cout,Ho,Wo],
 * Ho = (H-1)s-2p+k+out_pad.  Computed in s*s sub-pixel phases (no zero insertion).
 * Also serves as dgrad of conv2d (same weight memory, Cin:=conv Cout). */
int b2c_conv_transpose2d_fwd(const float* x, const float* w, const float* bias, float* y,
                             int B, int Cin, int H, int W, int Cout, int k, int stride,
                             int pad, int out_pad, void* stream);

/* weight gradient shared by both layer types (autograd of the above; train.py:105).
 *   dW[Cs][Cb][k][k] (+)= sum_{n,ys,xs} small[n,Cs,ys,xs] * big[n,Cb,s*ys-p+ky,s*xs-p+kx]
 * conv2d:            small = grad_out [B,Cout,Ho,Wo], big = x  [B,Cin,H,W]  -> dW is OIHW
 * conv_transpose2d:  small = x [B,Cin,H,W],          big = grad_out [B,Cout,Ho,Wo] -> IOHW
 * dW must be zero-filled by the caller (split-K partial sums are added atomically). */
int b2c_conv_wgrad(const float* small_t, const float* big_t, float* dW,
                   int B, int Cs, int hs, int ws, int Cb, int hb, int wb,
                   int k, int stride, int pad, void* stream);

/* out[c] = sum_{n,pix} t[n,c,pix]  (bias gradients). */
int b2c_channel_sum(const float* t, float* out, int B, int C, int HW, void* stream);

/* ---------------------------------------------------------------------------------
 * GDN / IGDN: models/GDN.py:64-94 (forward), :10-24 (LowerBound), autograd for bwd.
 * beta_p[C], gamma_p[C,C] are the raw (reparametrised) parameters as in state_dict.
 * --------------------------------------------------------------------------------- */
int b2c_gdn_fwd(const float* x, const float* beta_p, const float* gamma_p, float* y,
                int B, int C, int HW, int inverse, void* stream);
/* gx[B,C,HW], dbeta_p[C], dgamma_p[C,C] are assigned;
 * workspace: B*C*HW floats (dL/dnorm) + C + C*C floats. */
size_t b2c_gdn_bwd_workspace_bytes(int B, int C, int HW);
int b2c_gdn_bwd(const float* x, const float* beta_p, const float* gamma_p, const float* gy,
                float* gx, float* dbeta_p, float* dgamma_p,
                int B, int C, int HW, int inverse, void* workspace, size_t workspace_bytes,
                void* stream);

/* chain rule through eff = max(p, bound)^2 - pedestal with LowerBound's gradient rule (models/GDN.py:18-24,73-79):
 * (dL/dbeta_eff, dL/dgamma_eff) -> (dL/dbeta_p, dL/dgamma_p); all fp32, gamma [C][C]. */
int b2c_gdn_reparam_bwd(const float* beta_p, const float* gamma_p, const float* dbeta_eff, const float* dgamma_eff,
                        float* dbeta_p, float* dgamma_p, int C, void* stream);

/* ---------------------------------------------------------------------------------
 * Quantizer + factorized entropy model: model.py:53-56 (quantizer), model.py:71-78 +
 * models/bitEstimator.py:20-25,38-42 (bits).
 * y[B,C,HW]; noise[B,C,HW] (training: z = y + noise) or NULL (eval: z = rint(y), half to
 * even).  host_params: HOST array of B2C_NUM_BITPARAMS device pointers, each [C], order
 * f1.h f1.b f1.a f2.h f2.b f2.a f3.h f3.b f3.a f4.h f4.b.
 * z[B,C,HW] out; plane_bits[B*C] out (bits of each (image, channel) plane);
 * image_bits[B] out (deterministic fixed-order sum over channels).
 * --------------------------------------------------------------------------------- */
int b2c_quant_entropy_fwd(const float* y, const float* noise, const float* const* host_params,
                          float* z, float* plane_bits, float* image_bits,
                          int B, int C, int HW, void* stream);
/* Inference form of the same pass (eval quantizer only): hardware tanh / exp2 / log2 approximations in the
 * per-element chain (bits change by ~1e-6 relative), per-image bits only, and the rounded latent is also written
 * as NHWC bf16 (z_nhwc_bf16, may be NULL) for the tensor-core synthesis transform.  workspace: one float per
 * (image, 32-pixel chunk). */
size_t b2c_quant_entropy_infer_workspace_bytes(int B, int HW);
/* Table-driven form: at inference z = round(y) is an integer, so an element's bits depend only on (channel, z).
 * b2c_entropy_lut_build fills lut[C][2*half_range + 1] (fp32, entry z + half_range) with the same per-element formula
 * - once per weight version; b2c_quant_entropy_infer_lut then reads y, writes z (fp32 NCHW, and NHWC bf16 when
 * z_nhwc_bf16 != NULL) and sums table entries (|z| > half_range falls back to the formula): an HBM-bound pass.
 * Requires C % 8 == 0, C <= 256 and HW % 4 == 0; same workspace as b2c_quant_entropy_infer. */
int b2c_entropy_lut_build(const float* const* host_params, float* lut, int C, int half_range, void* stream);
/* range_flag (both forms; may be NULL): a device-accessible int (device memory, or mapped pinned host memory so the
 * host can poll it without synchronising) that the kernel sets to 1 when some |z| > 256 - beyond that the bf16 copy
 * z_nhwc_bf16 is no longer exact and the tensor-core synthesis transform must not consume it.  Sticky: never cleared
 * by the library. */
int b2c_quant_entropy_infer_lut(const float* y, const float* const* host_params, const float* lut, int half_range,
                                float* z, void* z_nhwc_bf16, float* image_bits, int* range_flag, int B, int C, int HW,
                                void* workspace, size_t workspace_bytes, void* stream);
int b2c_quant_entropy_infer(const float* y, const float* const* host_params, float* z, void* z_nhwc_bf16,
                            float* image_bits, int* range_flag, int B, int C, int HW, void* workspace,
                            size_t workspace_bytes, void* stream);
/* backward of total_bits = sum(image_bits): g_bits is a DEVICE scalar (dL/dtotal_bits).
 * dz[B,C,HW] out (assigned); host_dparams: HOST array of 11 device pointers [C], must be
 * zero-filled by the caller (atomically accumulated). */
int b2c_entropy_bwd(const float* z, const float* const* host_params, const float* g_bits,
                    float* dz, float* const* host_dparams, int B, int C, int HW, void* stream);

/* One Bitparm layer as a stand-alone call: models/bitEstimator.py:20-25 (non-final: t = x*softplus(h)+b,
 * out = t + tanh(t)*tanh(a); final (a == NULL): out = sigmoid(t)).  Four of them chained are BitEstimator.forward
 * (bitEstimator.py:38-42), the CDF a caller evaluates itself (model.py:71-72).  x, out [B,C,HW]; h, b, a [C].
 * Exact fp32 formula.  Backward: dx assigned, dh/db/da[C] assigned (deterministic fixed-order sums);
 * workspace: 3*B*C floats. */
int b2c_bitparm_fwd(const float* x, const float* h, const float* b, const float* a, float* out,
                    int B, int C, int HW, void* stream);
size_t b2c_bitparm_bwd_workspace_bytes(int B, int C);
int b2c_bitparm_bwd(const float* x, const float* h, const float* b, const float* a, const float* g,
                    float* dx, float* dh, float* db, float* da, int B, int C, int HW,
                    void* workspace, size_t workspace_bytes, void* stream);

/* ---------------------------------------------------------------------------------
 * Distortion: model.py:59 (clamp) and :61 (mse on the UNclamped reconstruction),
 * train.py:172 (mse on the clamped one).
 * sse[2*B]: per image {sum (recon-x)^2, sum (clamp(recon)-x)^2}; deterministic.
 * --------------------------------------------------------------------------------- */
int b2c_clamp_sse_fwd(const float* recon, const float* x, float* clipped, float* sse,
                      int B, int CHW, void* workspace, size_t workspace_bytes, void* stream);
size_t b2c_clamp_sse_workspace_bytes(int B, int CHW);
/* out[n*width + j] = sum_i partial[(n*per_image + i)*width + j], fixed order (deterministic); width <= 4. */
int b2c_partial_reduce(const float* partial, float* out, int B, int per_image, int width, void* stream);

/* End of an inference forward in ONE launch (model.py:61,73,78): sse[B][2] = fixed-order sums of the last layer's
 * per-tile partial sums sse_partial[B][tiles][2]; totals[4] = {sum of image_bits, sum of sse[:,0], that / numel (mse on
 * the unclamped reconstruction), bits / pixels (bpp)}.  With bits_partial[B][bits_slots] (the fused quantizer / entropy
 * epilogue of b2c_tc_layer.lut) image_bits is an OUTPUT, reduced here first; otherwise it is read.  `ticket`: one
 * zero-initialised device word the library uses to find the last block; the kernel leaves it at zero. */
int b2c_infer_stats(const float* sse_partial, int B, int tiles, float* image_bits, const float* bits_partial,
                    int bits_slots, double numel, double pixels, float* sse, float* totals, unsigned* ticket,
                    void* stream);
/* grad_recon = g_clipped * 1[0<=recon<=1] + g_sse[0] * 2 (recon - x);
 * g_clipped may be NULL; g_sse is a DEVICE scalar (dL/d total unclamped SSE). */
int b2c_clamp_sse_bwd(const float* recon, const float* x, const float* g_clipped,
                      const float* g_sse, float* g_recon, size_t n, void* stream);

/* =================================================================================
 * Tensor-core inference path (tcgen05 implicit GEMM, TMA-staged NHWC bf16 tiles, TMEM
 * accumulators) for out_channel_N = 128.  Activations between layers are NHWC bf16; the
 * analysis transform carries every tensor as a (hi, lo) bf16 pair and issues three MMAs
 * per product (hi*hi + hi*lo + lo*hi) so that rounded latents match the fp32 reference;
 * the synthesis transform runs in plain bf16 (split3 = 0).  GDN / IGDN (models/GDN.py:64-94)
 * is fused into the producing convolution's epilogue as a second tcgen05 MMA.
 * ================================================================================= */
#define B2C_TC_CONV1 0     /* 9x9 s4 p4, 3->128, input = fp32 NCHW image   (models/analysis_17.py:14,33)  */
#define B2C_TC_CONV5S2 1   /* 5x5 s2 p2, 128->128                          (models/analysis_17.py:18,22)  */
#define B2C_TC_DECONV5S2 2 /* 5x5 s2 p2 op1 transposed, 128->128, 4 phases (models/synthesis_17.py:15,19) */
#define B2C_TC_DECONV3 3   /* 9x9 s4 p4 op3 transposed, 128->3, as a 3x3 conv to 16 phases x 3 channels
                              (models/synthesis_17.py:23) */

#define B2C_TC_DECONV3X 4  /* the same layer (128 / 192 / 256 channels) with the horizontal input shifts folded into the GEMM's
                              N dimension (N = 112, 24 MMAs per tile instead of 72 of N = 48) and summed by warp shuffles
                              in the epilogue; tiles advance by 14 columns.  Own packed-weight layout: pack with this kind. */

/* `channels` = out_channel_N in {128, 192, 256} (model.py:39; 192 / 256 are the wider-latent variants of
 * BASELINE config 5; the training kinds - gdn 3/4, the norm plane, b2c_tc_wgrad* - exist for 128 only). */
/* bf16 elements in ONE packed weight plane (hi or lo) of a layer kind. */
size_t b2c_tc_packed_weight_elems(int kind, int channels);
/* fp32 OIHW (conv) / IOHW (deconv) weights as in state_dict -> tap-major bf16 panels; w_lo may be NULL. */
int b2c_tc_pack_weights(int kind, int channels, const float* w, void* w_hi, void* w_lo, void* stream);
/* GDN parameters (raw, reparametrised) -> gamma_eff as bf16 hi/lo [C][C] and beta_eff fp32 [C]. */
int b2c_tc_pack_gdn(int channels, const float* beta_p, const float* gamma_p, void* gamma_hi, void* gamma_lo,
                    float* beta_eff, void* stream);

/* gamma_eff transposed (bf16 [C][C]) for the GDN backward epilogue, and beta_eff */
int b2c_tc_pack_gdn_t(int channels, const float* beta_p, const float* gamma_p, void* gamma_t, float* beta_eff,
                      void* stream);

/* Several packing jobs in ONE launch (a training step re-packs 19 operands after every optimizer step; as separate
 * launches they sat on the critical path in front of the second layer).  kind = a layer kind (src = weights, dst_hi /
 * dst_lo as b2c_tc_pack_weights) or B2C_TC_PACK_GDN / _GDN_T (src = beta_p, src2 = gamma_p, dst_hi / dst_lo = gamma,
 * beta_eff as b2c_tc_pack_gdn / b2c_tc_pack_gdn_t).  Up to 16 jobs per call; 128 channels only. */
#define B2C_TC_PACK_GDN 16
#define B2C_TC_PACK_GDN_T 17
typedef struct b2c_tc_pack_job {
  int kind, channels;
  const float* src;
  const float* src2;
  void* dst_hi;
  void* dst_lo;
  float* beta_eff;
} b2c_tc_pack_job;
int b2c_tc_pack_many(const b2c_tc_pack_job* jobs, int n_jobs, void* stream);

typedef struct b2c_tc_layer {
  int kind;         /* B2C_TC_*                                                                   */
  int split3;       /* 1: (hi,lo) operands, 3 MMAs per product; 0: plain bf16                      */
  int gdn;          /* 0: none, 1: GDN, 2: IGDN fused after bias; 3 / 4: backward of GDN / IGDN applied to
                       the result (the launch is then the dgrad conv of the layer that consumed the GDN
                       output): out_hi = dL/dx (pre-GDN), out_lo = dL/dnorm, gamma_hi = gamma_eff^T        */
  int out_f32_nchw; /* 1: fp32 NCHW output in out_hi (required for DECONV3), 0: NHWC bf16 hi[/lo]  */
  int B, Hin, Win;  /* input grid                                                                  */
  const void* in_hi; /* NHWC bf16 [B,Hin,Win,128]  (CONV1: fp32 NCHW image [B,3,Hin,Win])          */
  const void* in_lo; /* NHWC bf16 low part when split3 (unused for CONV1)                          */
  const void* w_hi;  /* packed weights (b2c_tc_pack_weights)                                       */
  const void* w_lo;
  const float* bias; /* [Cout] or NULL                                                             */
  const void* gamma_hi; /* b2c_tc_pack_gdn outputs, when gdn != 0                                  */
  const void* gamma_lo;
  const float* beta_eff;
  void* out_hi;
  void* out_lo;
  /* DECONV3 / DECONV3X only, optional fused distortion (model.py:59,61; train.py:172): with ref_image = x (fp32 NCHW)
   * out_hi receives clamp(recon, 0, 1) and sse_partial[B][tiles][2] the per-tile sums of (recon - x)^2 and
   * (clamp(recon) - x)^2; reduce them with b2c_partial_reduce(sse_partial, sse, B, tiles, 2). */
  const float* ref_image;
  float* sse_partial;
  /* gdn 3 / 4: the forward GDN output y and norm n = beta + gamma x^2 saved by the forward launch
   * (a bf16 forward launch with gdn 1 / 2 writes n to out_lo), NHWC bf16 shaped like this launch's output */
  const void* saved_y;
  const void* saved_n;
  int channels;     /* 0 or 128, 192, 256 (see above)                                               */
  float* grad_out;  /* DECONV3 with ref_image (training): also write grad_scale * (recon - x), fp32 NCHW - the
                       gradient of (grad_scale / 2) * sum (recon - x)^2 w.r.t. the unclamped reconstruction
                       (model.py:61 through train.py:102), so no separate distortion forward / backward pass runs */
  float grad_scale;
  int ksplit;       /* 0 / 1: off.  2..25 (CONV5S2, out_f32_nchw, gdn == 0 only): the 25 taps are split over
                       ksplit CTAs per tile which ADD their partial sums into out_hi - the caller zero-fills it.
                       For small batches where one CTA per tile leaves most SMs idle (training, batch 8).   */
  float* dbias_out; /* DECONV3X with grad_out (training), optional: dbias_out[o] += sum over (n, y, x) of grad_out -
                       the gradient of Decoder.deconv3.bias (3 floats, accumulated with atomic adds: zero-fill or
                       accumulate on purpose), instead of a separate pass over grad_out                          */
  /* Last analysis layer at inference (CONV5S2, split3, gdn 0, out_f32_nchw, 128 channels: the persistent kernel), optional:
   * with `lut` (b2c_entropy_lut_build, half range lut_half_range) the epilogue is the quantizer and the entropy model too
   * (model.py:55-56,71-73): out_hi receives z = rint(conv3 output) instead of the pre-quantization value, z_nhwc_out its
   * bf16 NHWC copy (what the synthesis transform reads), bits_partial[B][8 * tiles] the bits of every (tile, epilogue warp)
   * - reduce with b2c_partial_reduce(bits_partial, image_bits, B, 8 * tiles, 1); bit_params = HOST array of the 11 Bitparm
   * device vectors (for |z| beyond the table); range_flag as in b2c_quant_entropy_infer (may be NULL). */
  const float* lut;
  int lut_half_range;
  const float* const* bit_params;
  void* z_nhwc_out;
  float* bits_partial;
  int* range_flag;
} b2c_tc_layer;

/* number of 8x16-pixel tiles (CTAs per image and phase) a layer kind launches for an input grid = the `tiles`
 * dimension of sse_partial (DECONV3X: tiles advance by 14 columns and every one of a tile's 16 epilogue warps owns a
 * partial-sum slot, so this is 16 x the tile count) */
int b2c_tc_tiles_per_image(int kind, int Hin, int Win);

int b2c_tc_layer_run(const b2c_tc_layer* layer, void* stream);

/* Tensor-core weight gradient (bf16 training path):
 *   dW[m][n][ky][kx] += sum_{img,ys,xs} small[img,ys,xs,m] * big[img, stride*ys - pad + ky, stride*xs - pad + kx, n]
 * small [B,hs,ws,128], big [B,hb,wb,128]: NHWC bf16; k in {1,5}, stride in {1,2}.
 *   conv2d weight grad (OIHW): small = grad_out, big = input;  conv_transpose2d (IOHW): small = input, big = grad_out;
 *   GDN gamma grad: k = 1, small = dL/dnorm, big = x^2.                               (autograd of train.py:105)
 * With a workspace of b2c_tc_wgrad_workspace_bytes() bytes (need not be zeroed): per-CTA partial tiles + a
 * fixed-order second kernel (deterministic, no atomics) that ADDS the result to dW in the state_dict layout
 * [128][128][k][k].  With workspace = NULL (legacy): dW is a zero-filled TAP-MAJOR [k][k][128][128] buffer that the
 * CTAs accumulate into with red.global.add. */
size_t b2c_tc_wgrad_workspace_bytes(int B, int hs, int ws, int k);
int b2c_tc_wgrad(const void* small_t, const void* big_t, float* dW, int B, int hs, int ws, int hb, int wb, int k,
                 int stride, int pad, void* workspace, size_t workspace_bytes, void* stream);
/* The red.global form of b2c_tc_wgrad accumulates tap-major, tmp[tap][m][n]; this writes (accumulate = 0) or adds
 * (1) it into the parameter's own layout out[m][n][tap] (OIHW / IOHW with ky*k + kx = tap). */
int b2c_tc_wgrad_untap(const float* tmp, float* out, int ntaps, int accumulate, void* stream);

/* 9x9 / stride-4 / pad-4 weight gradient on tensor cores: dW[m][c][ky][kx] += sum small[img,ys,xs,m] *
 * image[img,c,4ys-4+ky,4xs-4+kx]; small NHWC bf16 [B,h,w,128], image fp32 NCHW [B,3,4h,4w], dW fp32 [128][3][9][9]
 * zero-filled by the caller.  conv1 (OIHW): small = grad_out, image = x; deconv3 (IOHW): small = input, image = grad_out.
 * With a workspace of b2c_tc_wgrad9_workspace_bytes() bytes (need not be zeroed) the per-CTA sums are written as
 * tiles and reduced in a fixed order by a second kernel (deterministic, no atomics); with workspace = NULL they are
 * accumulated with red.global.add. */
size_t b2c_tc_wgrad9_workspace_bytes(int B, int h, int w);
int b2c_tc_wgrad9(const void* small_t, const float* image, float* dW, int B, int h, int w, void* workspace,
                  size_t workspace_bytes, void* stream);

/* helpers of the bf16 training path: x^2 of a GDN input from the saved output and norm (GDN: y^2 n, IGDN: y^2/n;
 * count bf16 elements, multiple of 8), and out[c] += sum_rows t[row][c] for a bf16 [P][128] tensor (out zero-filled). */
int b2c_tc_xsq(const void* y, const void* n, void* out, size_t count, int inverse, void* stream);
int b2c_tc_nhwc_colsum(const void* t, float* out, size_t P, void* stream);

/* layout / precision converters between the reference's fp32 NCHW and the NHWC bf16 (hi[,lo]) tiles */
int b2c_tc_nchw_to_nhwc(const float* in, void* out_hi, void* out_lo, int B, int C, int HW, void* stream);
int b2c_tc_nhwc_to_nchw(const void* in_hi, const void* in_lo, float* out, int B, int C, int HW, void* stream);

/* ---- fused training-step tail (next row N2): gradient clamp + Adam for all parameter tensors in one launch ----------
 * Replaces train.py:106-112 (`p.grad.data.clamp_(-5, 5)` per tensor, then `optim.Adam(...).step()`: betas (0.9,
 * 0.999), eps 1e-8, no weight decay, train.py:233).  params[i] / numel[i]: the n_tensors (<= 32) fp32 parameter tensors
 * (host arrays of device pointers / element counts); grads_flat, exp_avg, exp_avg_sq: flat fp32 device buffers holding
 * the gradients and both moments concatenated in the same order; step_dev: device int32 holding the 1-based step number
 * of THIS update; lr_dev: device float learning rate (both read on the device, so a CUDA graph can replay the launch
 * while the host advances them).  grad_clip <= 0 disables the clamp.  beta1 / beta2 / eps are doubles because torch
 * forms 1 - beta in double before rounding the scalar to fp32. */
int b2c_adam_clamp_step(float* const* params, const size_t* numel, int n_tensors, const float* grads_flat,
                        float* exp_avg, float* exp_avg_sq, const int* step_dev, const float* lr_dev, double beta1,
                        double beta2, double eps, float grad_clip, void* stream);

/* ---- ImageCompressor_new encoder tail (next row N1): sigmoid + straight-through binarizer --------------------------
 * models/analysis_17_new.py:35-37 + models/binarizer.py:19-27.  y: conv3 output, fp32 NCHW [B,C,HW]; sig = sigmoid(y),
 * code = (sig > 0.5) as 0.0 / 1.0 (both fp32 NCHW); code_nhwc_bf16 (optional, may be NULL): the code as NHWC bf16 for
 * the tensor-core synthesis transform.  Backward (straight-through): g_y = (g_code + g_sig) * sig * (1 - sig); either
 * incoming gradient may be NULL. */
int b2c_sigmoid_binarize_fwd(const float* y, float* sig, float* code, void* code_nhwc_bf16, int B, int C, int HW,
                             void* stream);
int b2c_sigmoid_binarize_bwd(const float* sig, const float* g_code, const float* g_sig, float* g_y, size_t count,
                             void* stream);

/* ---- SSIM / MS-SSIM on the GPU (next row N3; models/ms_ssim_torch.py, used by train.py:178 on CPU copies) -----------
 * b2c_ssim_level: X, Y fp32 NCHW [B,C,H,W]; win: device array of win_size (odd, <= 33) 1-D Gaussian taps
 * (ms_ssim_torch.py:6-18); valid padding; sums[b*2 + 0] = sum over (C, H-win+1, W-win+1) of the ssim map,
 * sums[b*2 + 1] = the same for the cs map (ms_ssim_torch.py:59-71); C1 = (0.01 data_range)^2, C2 = (0.03 data_range)^2.
 * b2c_avgpool2: F.avg_pool2d(x, 2, padding=(pad_h, pad_w)) over `planes` HxW planes (ms_ssim_torch.py:164-166).
 * b2c_msssim_combine: out[b] = prod_{l<levels-1} (cs_l[b])^w_l * (ssim_last[b])^((levels-1) w_last) with
 * sums[levels][B][2] and the per-level element counts (device float[levels]).  The (levels-1) exponent is the
 * reference's: ms_ssim_torch.py:169-170 broadcasts the ssim factor inside the product over the first levels-1 rows. */
size_t b2c_ssim_workspace_bytes(int B, int C, int H, int W, int win_size);
int b2c_ssim_level(const float* X, const float* Y, const float* win, int win_size, float C1, float C2, float* sums, int B,
                   int C, int H, int W, void* workspace, size_t workspace_bytes, void* stream);
int b2c_avgpool2(const float* in, float* out, int planes, int H, int W, int pad_h, int pad_w, void* stream);
int b2c_msssim_combine(const float* sums, const float* counts, const float* weights, float* out, int levels, int B,
                       void* stream);

/* ---- 8-bit input (host pipelines): out[i] = (float)in[i] / 255, the conversion torchvision's ToTensor performs on the
 * host in the reference's data loader (datasets.py:24-34), bit-identical (IEEE division); count % 4 == 0.  Lets a
 * caller ship uint8 images over PCIe (4x fewer bytes than the fp32 tensors of train.py:96,170). */
int b2c_u8_to_f32(const void* in_u8, float* out, size_t count, void* stream);

/* ---- gradient all-reduce over NVLink / NVSwitch peer memory (data-parallel training, replaces the reduce of
 * nn.DataParallel, train.py:228) -------------------------------------------------------------------------------------
 * In-place mean/sum of a flat fp32 buffer that lives in SYMMETRIC memory (the same allocation mapped into every rank,
 * e.g. torch.distributed._symmetric_memory): this rank reduces elements [rank * count / world, (rank+1) * count / world)
 * over all ranks, multiplies by `scale` (1 / world for the mean) and writes the result to every rank's copy.
 * multicast_ptr != NULL: NVLS form (multimem.ld_reduce / multimem.st on the multicast address; the switch sums);
 * otherwise peer_ptrs[0..world) (host array of the ranks' buffer addresses as mapped in THIS process) are read and
 * written directly, sums in rank order (deterministic).  count % (4 * world) == 0.  The caller brackets the launch with
 * a cross-rank barrier on the same stream (all inputs written before, all results visible after). */
int b2c_allreduce_symm(void* const* peer_ptrs, void* multicast_ptr, int rank, int world, size_t count, float scale,
                       void* stream);
/* The same over elements [begin, begin + count) only (both multiples of 4; the float4s are split over the ranks as
 * evenly as possible) and with at most max_blocks CTAs (<= 0: no cap): one gradient bucket reduced beside the compute
 * kernels of the rest of the backward pass (train_tc.GraphedTrainStep captures two such launches in its step graph). */
int b2c_allreduce_symm_range(void* const* peer_ptrs, void* multicast_ptr, int rank, int world, size_t begin,
                             size_t count, float scale, int max_blocks, void* stream);
/* Barrier -> reduce -> barrier in ONE launch.  The symmetric buffer carries b2c_allreduce_symm_flag_words() uint32
 * words per channel behind its data, zero-filled once (flag_offset: first word of the channel used by this launch, in
 * elements from the buffer base, >= begin + count; one channel per gradient bucket).  peer_ptrs is always required
 * (flags travel through peer memory), multicast_ptr selects the NVLS data path.  The launch number lives in the flag
 * block, not in an argument: the launch can be captured in a CUDA graph and replayed.  Every rank must launch the
 * same sequence per channel.  max_blocks <= 0: up to 148 CTAs. */
int b2c_allreduce_symm_flag_words(void);
int b2c_allreduce_symm_fused(void* const* peer_ptrs, void* multicast_ptr, int rank, int world, size_t begin,
                             size_t count, size_t flag_offset, float scale, int max_blocks, void* stream);

/* ---- real bitstream for the integer latents (the reference only estimates the size, model.py:71-74, and
 * approximates files with gzip(np.packbits(latent == 0)), NewTests/testModelRecon.py:80) ----------------------------
 * Warp-interleaved range-ANS coder over the BitEstimator's per-channel probabilities (format: csrc/b2c_rans.cu).
 * Tables: freq_cum [C][A] uint32 = f | (start << 16), slot_sym [C][2^prec] uint8, A = 2 * half_range + 2 (last symbol
 * = escape + 16 raw bits), 1 <= half_range <= 126, 10 <= prec <= 15.  z: fp32 NCHW [B][C][HW] holding integers
 * (|z| <= 32767).  One warp codes a block of block_symbols (multiple of 32) consecutive latents of one image; blocks
 * per image = b2c_rans_num_blocks.  Encode writes block_bytes[B * blocks], offsets[B * blocks + 1] (bytes, exclusive
 * scan) and the concatenated block streams to out; *status (device int, caller zeroes it) gets bit 0 = value out of
 * range, bit 1 = out_capacity too small (nothing past the capacity is written), bit 2 (decode) = corrupt stream.
 * Decode reads the same offsets + bytes and writes z.  Nothing synchronises; all buffers are caller-owned. */
int b2c_rans_build_tables(const float* const* host_params, int C, int half_range, int prec, uint32_t* freq_cum,
                          uint8_t* slot_sym, void* stream);
int b2c_rans_num_blocks(int C, int HW, int block_symbols);
size_t b2c_rans_scratch_bytes(int B, int C, int HW, int block_symbols);
int b2c_rans_encode(const float* z, const uint32_t* freq_cum, int half_range, int prec, int B, int C, int HW,
                    int block_symbols, void* scratch, size_t scratch_bytes, uint32_t* block_bytes, uint64_t* offsets,
                    uint8_t* out, size_t out_capacity, int* status, void* stream);
int b2c_rans_decode(const uint8_t* in, const uint64_t* offsets, const uint32_t* freq_cum, const uint8_t* slot_sym,
                    int half_range, int prec, int B, int C, int HW, int block_symbols, float* z, int* status,
                    void* stream);

/* ---- training input transform on the device ---------------------------------------------------------
 * Replaces, per batch, the reference's per-image CPU Compose of datasets.py:24-34: RandomResizedCrop(out_size) ->
 * RandomHorizontalFlip -> RandomVerticalFlip -> ToTensor, i.e. torchvision resized_crop (PIL crop + antialiased BILINEAR
 * resize, Pillow src/libImaging/Resample.c, 22-bit fixed point, horizontal pass first), flips, uint8 / 255.  The random
 * draws stay on the host.  src: decoded RGB images, uint8 HWC, packed in one device buffer (16-byte aligned, each image
 * at a byte offset that is a multiple of 16).  desc: device int64 [B][8] = {byte offset, H, W, top, left, crop_h,
 * crop_w, flags (bit 0 horizontal flip, bit 1 vertical flip)}.  out: fp32 [B][3][out_size][out_size], bit-identical to
 * the reference transform.  out_size: multiple of 32 in [32, 512].  max_crop_h / max_crop_w: upper bounds of the crop
 * sizes in desc (they size the shared memory; B2C_EINVAL if they need more than 227 KB).  *status (device int, caller
 * zeroes it) gets bit 0 when a descriptor is outside its image / the buffer / the maxima (that image is not written). */
int b2c_resized_crop_flip(const uint8_t* src, size_t src_bytes, const int64_t* desc, int B, int out_size, int max_crop_h,
                          int max_crop_w, float* out, int* status, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* B2C_H_ */
