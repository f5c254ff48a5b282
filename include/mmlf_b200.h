/* mmlf_b200 -- C ABI of the B200 (sm_100a) hot path of titus-leistner/mmlf.
 *
 * The reference is pure Python/PyTorch and has no FFI of its own; each entry point below
 * names the reference call site(s) it replaces (paths relative to /root/reference).
 * INTEGRATION.md shows the ctypes binding a maintainer would add on the reference side.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer unless the parameter is documented as host memory;
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued asynchronously on it;
 *   - return value 0 = success, anything else = failure, text via mmlf_last_error();
 *   - nothing here allocates device memory: the caller owns every buffer.
 *
 * "Slot" layout: activations of a batch of B images of H x W pixels are
 * stored channel-last in bf16 as rows of a 2-D array [n_slots][ld] with
 *     Hp = H + 1, Wp = W + 1, n_slots = B * Hp * Wp, slot(b, sy, sx) = (b * Hp + sy) * Wp + sx.
 * An H x W tensor lives at slots (y + 1, x + 1) with slot row 0 / slot column 0 holding zeros;
 * an (H+1) x (W+1) tensor (output of the padded first conv of a block) fills the whole grid.
 * With this layout each tap of a 2x2 convolution is a constant row offset, so the implicit
 * GEMM reads its A operand with plain 2-D TMA boxes.
 */
#ifndef MMLF_B200_H
#define MMLF_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MMLF_ABI_VERSION 5

/* 16-bit storage formats.  Forward activations and weights are fp16 (11 significant bits; the reference's own GPU
 * path multiplies in TF32, 11 bits), gradients are bf16 (fp32 exponent range); accumulation is always fp32. */
#define MMLF_BF16 0
#define MMLF_FP16 1

const char* mmlf_last_error(void);
int mmlf_abi_version(void);
/* 0 if the current device is sm_100 and the kernels can run, else an error. */
int mmlf_check_device(void);

/* ------------------------------------------------------------------ light-field input (mmlf/data/hci4d.py) */

/* View-index extraction + u8 -> f32 (hci4d.py:142-193, HCI4D.load_scene).
 * views: [n*n][H][W][3] uint8 in sorted-file (row-major grid) order.
 * h, v, i, d: [n][3][H][W] float32 stacks (centre row, centre column, rising diagonal reversed,
 * falling diagonal); center: [3][H][W] = v[n/2] (may be NULL).  Bit-exact: x / 255. */
int mmlf_lf_extract_u8(const uint8_t* views, int n, int H, int W, float* h, float* v, float* i, float* d,
                       float* center, void* stream);

/* Shift.__call__ (hci4d.py:907-990) on the four stacks, out of place, one launch.
 * src_x / dst_x: [batch][n][3][H][W] float32, src != dst.  Two-tap circular lerp per view; diagonal stacks get the
 * W lerp, an fp32 rounding, then the H lerp (opposite sign for the i stack).  Bit-exact. */
int mmlf_lf_shift(const float* src_h, const float* src_v, const float* src_i, const float* src_d, float* dst_h,
                  float* dst_v, float* dst_i, float* dst_d, int batch, int n, int H, int W, double disp,
                  void* stream);

/* create_mask_texture (hci4d.py:38-69): mask[b][y][x] = (mean over 3 colours x wsize^2 zero-padded neighbours of
 * |center[b][c][y+dy][x+dx] - center[b][c][y][x]| >= threshold) and (y, x) at least wsize / 2 away from the border.
 * center: (B, 3, H, W) f32; mask: (B, H, W) int32; mae (optional): the mean-L1 map, (B, H, W) f32. */
int mmlf_texture_mask(const float* center, int B, int H, int W, int wsize, double threshold, int32_t* mask, float* mae,
                      void* stream);

/* Host helper: the per-view taps of Shift (hci4d.py:934-938): weights rounded to f32, integer shifts.
 * All four arrays are HOST memory of length n. */
int mmlf_shift_taps(double disp, int n, float* w0, float* w1, int* s0, int* s1);

/* Fold views into channels and convert to the bf16 slot layout (feed_forward.py:226-232).
 * views: [B][C][H][W] float32 (C = n*3); out: [n_slots][ld] bf16, channels >= C and halo slots zeroed. */
int mmlf_pack_views(const float* views, int B, int C, int H, int W, void* out, int ld, int dtype, void* stream);

/* Split-precision packing: hi = fp16(x) to columns [0, c_pad), lo = fp16(x - hi) to columns [c_pad, 2 c_pad) of
 * out [n_slots][ld] (ld >= 2 * c_pad, c_pad a multiple of 8 and >= C). */
int mmlf_pack_views_split(const float* views, int B, int C, int H, int W, void* out, int ld, int c_pad, void* stream);

/* Fused Shift + pack for the ESE sweep (ensamble.py:63-70 + feed_forward.py:226-232): the shifted fp32 value
 * is rounded to bf16 and written straight into the slot layout.  stack: 0 = h, 1 = v, 2 = i, 3 = d. */
int mmlf_shift_pack(const float* src, int stack, int B, int n, int H, int W, double disp, void* out, int ld,
                    int dtype, void* stream);
/* All stacks of a forward pass in ONE launch (grid.z = stack), each optionally written twice from the same read: `out`
 * in `dtype` and `out2` (NULL = none) in `dtype2` -- the bf16 twin of the fp16 activations that the weight-gradient GEMM
 * reads.  views / out / out2: HOST arrays of n_stacks device pointers, stacks[i] = 0 h, 1 v, 2 i, 3 d (only used when
 * do_shift != 0: the ESE Shift of mmlf_shift_pack with disparity `disp`).  Each source (B, n, 3, H, W) f32. */
int mmlf_pack_stacks(const float* const* views, const int* stacks, int n_stacks, int B, int n, int H, int W,
                     void* const* out, void* const* out2, int ld, int dtype, int dtype2, int do_shift, double disp,
                     void* stream);

/* ------------------------------------------------------------------ convolution (mmlf/model/feed_forward.py:122-137) */

/* Weight packing: canonical nn.Conv2d weight (cout, cin, 2, 2) f32 -> bf16 K-major GEMM B operand
 * [n_pad][4 * kc * 64] (kc = ceil(cin_pad / 64)); column (tap * kc * 64 + c), tap = dy * 2 + dx.
 *   spatial: 0 = as is (v / d streams), 1 = transposed taps (h stream, replaces the permutes of
 *            feed_forward.py:236-241), 2 = flipped + transposed (i stream, feed_forward.py:248-256)
 *   dgrad  : 0 = forward operand; 1 = data-gradient operand (taps rotated by 180 degrees, cin/cout swapped:
 *            rows = cin, columns = cout)
 *   in_groups/group_real/group_pad: the input channels form `in_groups` groups of `group_real` real channels
 *            stored with pitch `group_pad` (the 4 x 70 -> 4 x 80 concatenated feature buffer); 1/cin/cin_pad else. */
int mmlf_pack_conv_weight(const float* w, int cout, int cin, int spatial, int dgrad, int in_groups, int group_real,
                          int group_pad, void* out, int n_pad, int cin_pad, int dtype, void* stream);

/* Split-precision operand for mmlf_conv_args.split_in: fp16 [n_pad][4 * 3 * kc * 64]; per tap the K blocks are
 * [w_hi | w_lo | w_hi] (w' = w * weight_scale, w_hi = fp16(w'), w_lo = fp16(w' - w_hi)), matching the activation
 * blocks [hi | hi | lo].  weight_scale is a power of two (64 in the engine) that lifts the residuals of typical conv
 * weights (|w| ~ 0.03) out of the fp16 subnormal range; the caller divides it out through `scale` in the epilogue. */
int mmlf_pack_conv_weight_split(const float* w, int cout, int cin, int spatial, int in_groups, int group_real,
                                int group_pad, void* out, int n_pad, int cin_pad, float weight_scale, void* stream);

/* All weight packings of a training / inference step in ONE launch (the per-layer calls above cost ~150 launches of
 * 5-7 us per step, 2.5 % of a 64-patch step).  jobs: DEVICE array; a job = one mmlf_pack_conv_weight[_split] call,
 * plus the zero-padded fp32 copy of the bias when bias / bias_pad are set. */
typedef struct mmlf_pack_job {
  const float* w;          /* (cout, cin, 2, 2) f32                                            */
  void* out;               /* packed operand [n_pad][4 * terms * ceil(cin_pad / 64) * 64]      */
  const float* bias;       /* (cout) f32 or NULL                                               */
  float* bias_pad;         /* (n_pad) f32 or NULL: bias_pad[i] = i < cout ? bias[i] : 0        */
  int cout, cin, spatial, dgrad, in_groups, group_real, group_pad, n_pad, cin_pad, dtype, split;
  float weight_scale;      /* split only (else 1)                                              */
} mmlf_pack_job;
int mmlf_pack_conv_weights_batch(const mmlf_pack_job* jobs, int n_jobs, int64_t max_elems, void* stream);

/* Inverse for gradients: dw_packed [n_pad][4][cin_pad] f32 (same tap/column convention, forward orientation)
 * -> canonical (cout, cin, 2, 2) f32; accumulate != 0 adds (two streams share one module). */
int mmlf_unpack_conv_wgrad(const float* dw_packed, int n_pad, int cin_pad, int cout, int cin, int spatial,
                           int in_groups, int group_real, int group_pad, float* dw, int accumulate, void* stream);

typedef struct mmlf_conv_args {
  const void* in;        /* 16-bit [n_slots][ld_in]; first cin_pad channels are the operand       */
  int ld_in;             /* row pitch in elements (multiple of 8)                                 */
  int cin_pad;           /* multiple of 16                                                         */
  const void* wpack;     /* from mmlf_pack_conv_weight, [n_pad][4 * kc * 64], same format as `in`  */
  int n_pad;             /* multiple of 16, <= 320                                                 */
  int B, H, W;           /* image geometry (slot grid is (H+1) x (W+1))                            */
  int type;              /* 0: padded conv (nn.Conv2d(.., 2, padding=1), feed_forward.py:123): output on the
                               whole slot grid, taps at rows {0, 1, Wp, Wp+1};
                            1: valid conv (padding=0, feed_forward.py:125): output at slots (y+1, x+1), taps at
                               rows {-Wp-1, -Wp, -1, 0}; halo slots are written as zero                 */
  const float* bias;     /* [n_pad] or NULL                                                        */
  const float* scale;    /* [n_pad] or NULL: y = (acc + bias) * scale + shift (eval-mode BN fold)   */
  const float* shift;    /* [n_pad] or NULL                                                        */
  int relu;              /* apply max(.,0) (nn.ReLU, feed_forward.py:124,135)                       */
  const uint32_t* gate_bits; /* [n_slots][ld_bits] or NULL: out[slot][c] is zeroed unless bit (c & 31) of word c / 32
                            is set -- ReLU backward with the sign bits saved by the forward pass (out_mode 0)   */
  uint32_t* relu_bits;   /* [n_slots][ld_bits] or NULL: receives the bits (out[slot][c] > 0) (out_mode 0)   */
  int ld_bits;           /* words per slot of gate_bits / relu_bits, >= ceil(n_pad / 32), <= 10      */
  void* out;             /* see out_mode                                                           */
  int ld_out;
  int out_mode;          /* 0: 16-bit [n_slots][ld_out]; 1: f32 [n_slots][ld_out];
                            2: f32 planar (B, n_real, Ho, Wo), Ho x Wo = (H+1)x(W+1) for type 0, H x W for type 1 */
  int n_real;            /* channels written in out_mode 2 (<= n_pad); out_mode 0/1 write n_pad channels */
  void* out2;            /* NULL or a second copy of the out_mode-0 output in out2_dtype, [n_slots][ld_out2]:
                            the forward pass of training keeps fp16 for the next conv and bf16 for the weight
                            gradient (tcgen05 kind::f16 needs both operands in one format)              */
  int ld_out2;
  double* col_sums;      /* NULL or [2][n_pad]: += per-channel sum and sum of squares of the stored (rounded)
                            output over all slots (halo slots are zero): BatchNorm batch statistics
                            (feed_forward.py:134) and bias gradients come out of the conv epilogue (out_mode 0) */
  int ab_dtype;          /* storage format of `in` and `wpack`: MMLF_BF16 or MMLF_FP16                */
  int out_dtype;         /* storage format of `out` in out_mode 0                                     */
  int out2_dtype;        /* storage format of `out2`                                                  */
  /* Split-precision inference ("3 x fp16", fp32-class products on the fp16 tensor cores): a value x is stored as
   * hi = fp16(x) and lo = fp16(x - hi) and x * w is evaluated as hi*w_hi + hi*w_lo + lo*w_hi in the fp32 accumulator.
   * split_in  != 0: `in` holds hi in columns [0, cin_pad) and lo in columns [split_in, split_in + cin_pad); `wpack`
   *                 comes from mmlf_pack_conv_weight_split (three K blocks per tap); ab_dtype must be MMLF_FP16.
   * split_out != 0: out_mode 0 writes hi to `out` and lo to `out + split_out` elements (same pitch); out2 must be NULL. */
  int split_in;
  int split_out;
  /* BatchNorm-backward statistics fused into a data-gradient launch (out_mode 0, col_sums required): the output of this
   * launch is g = d loss / d y for y = relu(z * bn_scale + bn_shift), the output of a training-mode BatchNorm whose
   * pre-activation z (16-bit [n_slots][ld_z], format bn_z_dtype) the forward pass kept.  With m = (z * bn_scale +
   * bn_shift > 0), col_sums then receives  [0][c] += sum_slots g*m  and  [1][c] += sum_slots g*m*(z - bn_mean)  over
   * the stored (rounded) g -- exactly what mmlf_bn_bwd_reduce computes up to the factor invstd (mmlf_bn_bwd_apply with
   * train = 2 applies it), without reading g and z again. */
  const void* bn_z;      /* NULL = plain column sums                                                 */
  int ld_z;
  int bn_z_dtype;
  const float* bn_scale; /* [n_pad] from mmlf_bn_finalize                                            */
  const float* bn_shift;
  const float* bn_mean;  /* [n_pad] save_mean                                                        */
} mmlf_conv_args;

/* 2x2 convolution as an implicit GEMM on tcgen05 (TMA-fed, TMEM accumulators, fused epilogue).  Forward of
 * nn.Conv2d at feed_forward.py:123/125 and, with dgrad-packed weights, its data gradient. */
int mmlf_conv2x2(const mmlf_conv_args* args, void* stream);

/* Same contract on CUDA cores (fp32 FMA over the bf16 operands); slow; used by the tests to cross-check the
 * tensor-core kernel on the device. */
int mmlf_conv2x2_simt(const mmlf_conv_args* args, void* stream);

/* Weight gradient of the same convolution (autograd of feed_forward.py:123/125):
 *   dw[n][tap][c] = sum_slots dout[slot][n] * act[slot + tap_offset(type)][c]
 * dout: bf16 [n_slots][ld_dout] (n_pad channels), act: bf16 [n_slots][ld_act] (cin_pad channels).
 * workspace: f32, at least mmlf_conv2x2_wgrad_workspace(...) bytes.  dw: f32 [n_pad][4][cin_pad].
 * act_dtype may differ from dout_dtype (fp16 activations x bf16 gradients): tcgen05.mma kind::f16 faults on mixed f16 x bf16
 * operands (measured on B200), so the kernel converts its activation boxes to dout_dtype in shared memory after the TMA
 * load -- bit-identical to mmlf_convert16 followed by a same-format launch, without the extra copy in HBM. */
int64_t mmlf_conv2x2_wgrad_workspace(int n_pad, int cin_pad);
int mmlf_conv2x2_wgrad(const void* dout, int ld_dout, int n_pad, const void* act, int ld_act, int cin_pad, int B,
                       int H, int W, int type, int act_dtype, int dout_dtype, float* workspace, float* dw,
                       void* stream);
/* The same weight gradient written straight into the canonical (cout, cin, 2, 2) f32 tensor of the parameter (+= when
 * accumulate): the K-split reduction undoes the per-stream tap mapping (spatial) and the channel-group padding of
 * mmlf_pack_conv_weight itself, i.e. mmlf_conv2x2_wgrad + mmlf_unpack_conv_wgrad in one reduction launch. */
int mmlf_conv2x2_wgrad_canonical(const void* dout, int ld_dout, int n_pad, const void* act, int ld_act, int cin_pad,
                                 int B, int H, int W, int type, int act_dtype, int dout_dtype, float* workspace, int cout,
                                 int cin, int spatial, int in_groups, int group_real, int group_pad, float* dw,
                                 int accumulate, void* stream);

/* Format conversion of a 16-bit slot array (C channels per slot, multiple of 8). */
int mmlf_convert16(const void* src, int ld_src, int src_dtype, void* dst, int ld_dst, int dst_dtype, int C,
                   int64_t n_slots, void* stream);

/* Column sums of a 16-bit slot array: out[c] (+)= sum_slots x[slot][c]  (bias gradients). */
int mmlf_colsum16(const void* x, int ld, int C, int64_t n_slots, int dtype, float* out, int accumulate, void* stream);

/* ------------------------------------------------------------------ BatchNorm (feed_forward.py:134) */

/* Per-channel sum and sum of squares over the valid pixels of z (bf16 slots, H x W tensor at (y+1, x+1)).
 * sums: double[2][C], must be zeroed by the caller. */
int mmlf_bn_stats(const void* z, int ld, int C, int B, int H, int W, int act_dtype, double* sums, void* stream);

/* Training-mode finalisation: batch mean / biased var -> scale = gamma * invstd, shift = beta - mean * scale;
 * running_mean / running_var (unbiased) momentum update and num_batches_tracked += 1, as nn.BatchNorm2d does.
 * save_mean / save_invstd (f32 [C]) are kept for the backward pass. */
int mmlf_bn_finalize(const double* sums, int C_real, int C, int64_t count, const float* gamma, const float* beta,
                     float* running_mean, float* running_var, int64_t* num_batches_tracked, float momentum,
                     float eps, float* scale, float* shift, float* save_mean, float* save_invstd, void* stream);

/* Eval-mode fold: scale = gamma / sqrt(running_var + eps), shift = (conv_bias - running_mean) * scale + beta;
 * the conv epilogue then needs no separate bias. */
int mmlf_bn_fold_eval(int C_real, int C, const float* gamma, const float* beta, const float* running_mean,
                      const float* running_var, const float* conv_bias, float eps, float* scale, float* shift,
                      void* stream);

/* y = relu(z * scale + shift) on valid slots, zero on halo slots; 16-bit in, 16-bit out (ReLU at feed_forward.py:135).
 * y2 (optional, NULL to skip): second copy of y in y2_dtype -- training keeps fp16 for the next convolution and bf16
 * for its weight gradient. */
int mmlf_bn_apply_relu(const void* z, int ld_z, const float* scale, const float* shift, int C, int B, int H, int W,
                       int act_dtype, void* y, int ld_y, void* y2, int ld_y2, int y2_dtype, void* stream);

/* Backward of BN(+ReLU) in training mode.  The ReLU mask is recomputed from z exactly as the forward pass computed
 * y = relu(fma(z, scale, shift)) (scale / shift from mmlf_bn_finalize), so y itself is not read.
 * Pass 1: with g = dy * (z * scale + shift > 0) and xhat = (z - mean) * invstd,
 * sums[0][c] = sum g, sums[1][c] = sum g * xhat (double[2][C], zeroed by the caller). */
int mmlf_bn_bwd_reduce(const void* dy, int ld_dy, const void* z, int ld_z, const float* scale, const float* shift,
                       const float* save_mean, const float* save_invstd, int C, int B, int H, int W, int grad_dtype,
                       int act_dtype, double* sums, void* stream);
/* Pass 2 (train = 2: sums[1] holds sum g*(z - mean) from a fused conv epilogue and is multiplied by invstd first):
 * dz = gamma * invstd * (g - sum_g / count - xhat * sum_gx / count) (16-bit, halo zero);
 * dgamma (+)= sum_gx, dbeta (+)= sum_g (f32 [C_real]; `accumulate` != 0 adds to what is there -- the shared in-nets are
 * called twice per forward).  gamma: the module's weight, f32 [C_real] (NOT padded).  fsums: f32 scratch [3 * C].
 * train = 0 gives the eval-mode gradient dz = g * gamma * invstd.  dz_colsum (optional, f32 [C]): += sum over slots of
 * dz as stored = the bias gradient of the convolution in front of the BatchNorm (feed_forward.py:125). */
int mmlf_bn_bwd_apply(const void* dy, int ld_dy, const void* z, int ld_z, const float* scale, const float* shift,
                      const float* gamma, const float* save_mean, const float* save_invstd, const double* sums,
                      int64_t count, int train, int C_real, int C, int B, int H, int W, int grad_dtype, int act_dtype,
                      void* dz, int ld_dz, float* dgamma, float* dbeta, int accumulate, float* fsums, float* dz_colsum,
                      void* stream);

/* ReLU backward alone (blocks without BatchNorm): dz = dy * (y > 0), bf16 slots. */
int mmlf_relu_bwd(const void* dy, int ld_dy, const void* y, int ld_y, int C, int64_t n_slots, int grad_dtype,
                  int act_dtype, void* dz, int ld_dz, void* stream);

/* ------------------------------------------------------------------ heads (feed_forward.py:270-302) */

/* Last conv of the BASE / UPR head: nn.Conv2d(OC, OC, 2, padding=0) with OC in {1, 2} in fp32 on CUDA cores.
 * mid: f32 [n_slots][ld_mid] (relu'd output of the first head conv, whole slot grid); w2: (OC, OC, 2, 2),
 * b2: (OC); out: (B, OC, H, W) f32.  mean = out[:, 0], logvar = out[:, 1] (feed_forward.py:270, 293). */
int mmlf_head_small(const float* mid, int ld_mid, int OC, const float* w2, const float* b2, int B, int H, int W,
                    float* out, void* stream);
/* Its backward: gout (B, OC, H, W) -> gmid bf16 [n_slots][ld_gmid] (first OC channels, rest zero, gated by
 * mid > 0), dw2 (OC, OC, 2, 2) and db2 (OC) accumulated into zeroed f32 buffers. */
int mmlf_head_small_bwd(const float* gout, const float* mid, int ld_mid, int OC, const float* w2, int B, int H,
                        int W, void* gmid, int ld_gmid, float* dw2, float* db2, void* stream);

/* UPR posterior (feed_forward.py:292-302 + laplacian :9-12): post[b][j] = exp(-|x_j - mean| / e^logvar) / (2 e^logvar). */
int mmlf_upr_posterior(const float* mean, const float* logvar, const float* bins, int steps, int64_t B, int64_t HW,
                       float* posterior, void* stream);

/* DPP head (feed_forward.py:276-290): one_hot = (max_c s == s); posterior = exp(s) / sum exp(s) (unstabilised);
 * mean = sum_c bins_t[c] * one_hot[c]; logvar = log sum_c (bins_n[c] - mean)^2 * posterior[c].
 * bins_t = torch.linspace table, bins_n = np.linspace table (they differ by 1 ulp in places).
 * one_hot / posterior may be NULL to skip materialising them. */
int mmlf_dpp_head(const float* scores, const float* bins_t, const float* bins_n, int steps, int64_t B, int64_t HW,
                  float* one_hot, float* posterior, float* mean, float* logvar, void* stream);

/* DPP targets (utils/dl.py:109-157). */
int mmlf_reg_to_class(const float* gt, const float* bins_t, int steps, double half_step, int64_t B, int64_t HW,
                      float* out, void* stream);
int mmlf_mpi_to_weights(const float* mpi, int K, const float* bins_t, int steps, double half_step, int64_t B,
                        int64_t HW, float* out, void* stream);

/* ------------------------------------------------------------------ losses (mmlf/model/loss.py), value + gradient fused */

/* Normaliser pre-pass.  sums (double[8], zeroed by caller):
 *   [0] sum(mask)                                              (loss.py:71 etc.)
 *   [1] sum(mask_padding), if given                            (loss.py:274)
 *   [2] sum_px sum_k w_k, if mpi given                         (loss.py:356)
 *   [3] count(sum_k w_k < 0.01), if mpi given                  (loss.py:359)
 *   [4] number of pixels B * HW                                (loss.py:275,361: flatten().shape[0])
 * mpi: (B, K, 5, H, W) f32 or NULL.  In a multi-rank job the caller all-reduces sums before the main pass. */
int mmlf_loss_prepass(const int32_t* mask, const int32_t* mask_padding, const float* mpi, int K, int64_t B,
                      int64_t HW, double* sums, void* stream);

/* kind: 0 MaskedL1Loss (loss.py:46-77)            1 MultiMaskedL1Loss (:88-103)
 *       2 ImprovedUncertaintyL1Loss (:262-294)    3 ImprovedMultiUncertaintyL1Loss (:344-372)
 *       4 MaskedMSELoss (:114-122, value only)    5 MaskedBadPix (:177-187, value only, threshold in `param`)
 * target: gt (B, H, W) for kinds 0/2/4/5, mpi (B, K, 5, H, W) for kinds 1/3.  sums: the (all-reduced) pre-pass sums.
 * loss_sum (double[1], zeroed): receives the un-normalised masked sum; value = loss_sum / max(sums[0], 1 if 0).
 * g_mean / g_logvar (B, H, W) f32 or NULL: d value / d mean, d value / d logvar.
 * pred_stride: element stride between the images of mean / logvar / g_mean / g_logvar (0 = HW, i.e. dense): with
 * 2 * HW, mean = out and logvar = out + HW address the planes of the network output (B, 2, H, W) in place
 * (feed_forward.py:270,293: `output[:, 0]`, `output[:, 1]`), and the gradients land in a (B, 2, H, W) tensor likewise. */
int mmlf_loss_regression(int kind, const float* mean, const float* logvar, const float* target, int K,
                         const int32_t* mask, const int32_t* mask_padding, const double* sums, double param,
                         int64_t B, int64_t HW, double* loss_sum, float* g_mean, float* g_logvar, int64_t pred_stride,
                         void* stream);

/* MaskedCrossEntropy (loss.py:145-160): l = logsumexp(relu(s)) - sum_c relu(s_c) t_c, masked mean.
 * target (B, S, H, W) f32 or, if NULL, built on the fly from gt with reg_to_class (utils/dl.py:109-131).
 * g_scores (B, S, H, W) f32 or NULL. */
int mmlf_loss_cross_entropy(const float* scores, const float* target, const float* gt, const float* bins_t,
                            double half_step, int steps, const int32_t* mask, const double* sums, int64_t B,
                            int64_t HW, double* loss_sum, float* g_scores, void* stream);

/* ------------------------------------------------------------------ ESE (mmlf/model/ensamble.py:78-101) */
/* means / logvars: (K, B, H, W) f32; disp: np.linspace(min, max, K) as f32 [K].
 * mean / logvar (B, H, W): member with minimal logvar (first on ties); posterior (B, K, H, W): Laplace mixture. */
int mmlf_ese_reduce(const float* means, const float* logvars, const float* disp, int K, int64_t B, int64_t HW,
                    float* mean, float* logvar, float* posterior, void* stream);

/* ------------------------------------------------------------------ optimiser (mmlf/train/cli.py:113-118,258) */
/* torch.optim.Adam update (betas, eps as given; no weight decay, no amsgrad) over one flat f32 buffer.
 * step = 1-based step count after the increment. */
int mmlf_adam_step(float* p, const float* g, float* m, float* v, int64_t n, double lr, double beta1, double beta2,
                   double eps, int64_t step, void* stream);
/* The same update with its step-dependent scalars in DEVICE memory, so that the launch can sit in a captured CUDA graph
 * and be replayed with a new learning rate / step count (train/cli.py:233-241 changes lr every iteration):
 * hyper: double[4] = { lr (written by the host before each replay), step count BEFORE this update (advanced here),
 * scratch, scratch }.  Two launches: a one-thread kernel advances the step and derives lr / bias_correction1 and
 * sqrt(bias_correction2) in float64 exactly like the host version, then the element-wise update. */
int mmlf_adam_step_dev(float* p, const float* g, float* m, float* v, int64_t n, double* hyper, double beta1,
                       double beta2, double eps, void* stream);

/* ------------------------------------------------------------------ small fixed-cost helpers of a training step */
/* cudaMemsetAsync(ptr, 0, bytes) on `stream` (accumulator scratch of the column-sum / statistics kernels). */
int mmlf_zero(void* ptr, int64_t bytes, void* stream);

/* One launch for all the short per-parameter vector updates at the end of a backward pass (conv bias gradients from the
 * float64 / padded-pitch column sums of the kernels that produced the gradients, accumulation for shared modules):
 *     dst[i] = (accumulate ? dst[i] : 0) + (float) src[i] [+ (float) src2[i]],  i < n.
 * jobs: DEVICE array of n_jobs descriptors (built once by the host; every pointer in it is a device pointer).  Two jobs
 * of one launch must not share a dst (they run concurrently): a parameter with two contributions -- the shared in-nets
 * are called twice per forward -- names both in ONE job. */
typedef struct mmlf_vec_job {
  const void* src;    /* f32 or f64 */
  const void* src2;   /* optional second contribution, or NULL */
  float* dst;
  int32_t n;
  int32_t src_f64;    /* bit 0: src is double, bit 1: src2 is double */
  int32_t accumulate;
  int32_t pad_;
} mmlf_vec_job;
int mmlf_vec_jobs(const mmlf_vec_job* jobs, int n_jobs, void* stream);

/* loss value of the masked losses without a host sync and without framework glue: out[0] = (float)(loss_sum[0] /
 * (sums[0] == 0 ? 1 : sums[0]))  (loss.py:73-77: no division when the mask is empty). */
int mmlf_loss_finish(const double* loss_sum, const double* sums, float* out, void* stream);

/* ------------------------------------------------------------------ generic float32 layers
 * Odd --model_ksize (feed_forward.py:86-92) and the --model_unet out-net (unet.py:8-132) run on CUDA-core float32
 * kernels over dense channel-last tensors: B x H x W pixels, C channels, rows of pitch ld >= C ([B*H*W][ld]). */

/* nn.Conv2d / nn.ConvTranspose2d weight -> GEMM operand [K][N] (float32):
 *   mode 0: conv forward, w (cout, cin, k, k) -> [(dy*k + dx)*cin + ci][co]
 *   mode 1: conv data gradient -> [((k-1-dy)*k + (k-1-dx))*cout + co][ci]
 *   mode 2: ConvTranspose2d(cin, cout, 2, stride 2) forward, w (cin, cout, 2, 2) -> [ci][(dy*2 + dx)*cout + co]
 *   mode 3: its data gradient -> [(dy*2 + dx)*cout + co][ci]
 * spatial (modes 0 / 1): the stream plumbing of feed_forward.py:236-256 folded into the taps -- 0 none, 1 transposed
 * (h stream), 2 transposed + flipped (i stream): effective tap (dy, dx) reads w[..][dx][k-1-dy]. */
int mmlf_g_pack_weight(const float* w, int cout, int cin, int k, int spatial, int mode, float* out, void* stream);
/* y[b][oy][ox][n] = act(bias[n] + sum_{dy,dx,ci} x[b][oy+dy-pad][ox+dx-pad][ci] * wg[(dy*k+dx)*cin+ci][n]), zero padding,
 * output (H + 2 pad - k + 1) x (W + 2 pad - k + 1); bias may be NULL; relu != 0 clamps at zero. */
int mmlf_g_conv(const float* x, int ld_x, const float* wg, const float* bias, int B, int H, int W, int cin, int cout,
                int k, int pad, int relu, float* y, int ld_y, void* stream);
/* Weight gradient, ADDED into the canonical parameter gradient dw (zeroed by the caller once per backward pass):
 * (cout, cin, k, k) with the `spatial` tap mapping undone, or -- transposed != 0, k = 1, cout = 4 * cout_t, dy in
 * space-to-depth form -- the (cin, cout_t, 2, 2) gradient of a ConvTranspose2d. */
int mmlf_g_conv_wgrad(const float* x, int ld_x, const float* dy, int ld_dy, int B, int H, int W, int cin, int cout, int k,
                      int pad, int spatial, int transposed, float* dw, void* stream);
/* out[c] += sum_rows x[row][c] (bias gradients). */
int mmlf_g_colsum(const float* x, int ld, int C, int64_t n_rows, float* out, void* stream);
/* BatchNorm statistics: sums[c] += sum x, sums[C + c] += sum x^2 (double[2 * C], zeroed by the caller); finish with
 * mmlf_bn_finalize(sums, C, C, ...). */
int mmlf_g_bn_stats(const float* x, int ld, int C, int64_t n_rows, double* sums, void* stream);
/* y = x * scale[c] + shift[c], optionally rectified (BatchNorm apply; scale / shift from mmlf_bn_finalize / _fold_eval). */
int mmlf_g_affine(const float* x, int ld_x, const float* scale, const float* shift, int C, int64_t n_rows, int relu,
                  float* y, int ld_y, void* stream);
/* BatchNorm backward.  x: the layer's input, gate (optional): the layer's OUTPUT after a ReLU (g = dy * (gate > 0)).
 * sums: double[2 * C] scratch, zeroed by the caller.  dx = gamma * invstd * (g - mean(g) - xhat * mean(g * xhat)) in
 * training mode, g * gamma * invstd for an eval-mode BatchNorm (train = 0).  dgamma / dbeta are ADDED to. */
int mmlf_g_bn_bwd(const float* dy, int ld_dy, const float* x, int ld_x, const float* gate, int ld_gate,
                  const float* gamma, const float* mean, const float* invstd, double* sums, int64_t count, int train,
                  int C, int64_t n_rows, float* dx, int ld_dx, float* dgamma, float* dbeta, void* stream);
/* out = dy * (y > 0). */
int mmlf_g_relu_bwd(const float* dy, int ld_dy, const float* y, int ld_y, int C, int64_t n_rows, float* out, int ld_out,
                    void* stream);
/* F.max_pool2d(x, 2) (unet.py:71) on dense (B, H, W, C): y (B, H/2, W/2, C), idx = arg-max position in the window
 * (first maximum, row-major); the backward pass writes every element of dx (B, H, W, C). */
int mmlf_g_maxpool2(const float* x, int B, int H, int W, int C, float* y, uint8_t* idx, void* stream);
int mmlf_g_maxpool2_bwd(const float* dy, const uint8_t* idx, int B, int H, int W, int C, float* dx, void* stream);
/* dst[b][yd0+y][xd0+x][cd0+c] (+)= src[b][ys0+y][xs0+x][cs0+c] for a h x w x C window (centre crop + concat of
 * unet.py:118-131 and their gradients). */
int mmlf_g_copy_window(const float* src, int Hs, int Ws, int ld_s, int cs0, int ys0, int xs0, float* dst, int Hd, int Wd,
                       int ld_d, int cd0, int yd0, int xd0, int B, int h, int w, int C, int accumulate, void* stream);
/* Depth-to-space of a transposed convolution evaluated as a 1x1 convolution to 4 * C channels: y4 (B, H, W, 4*C dense,
 * tap-major) -> out (B, 2H, 2W, C at channel offset c0 of pitch ld); inverse != 0 gathers the other way. */
int mmlf_g_depth_to_space(float* y4, float* out, int ld, int c0, int B, int H, int W, int C, int inverse, void* stream);
/* (B, C, H, W) <-> (B, H, W, C at pitch ld): to_nhwc != 0 reads nchw and writes nhwc, else the reverse. */
int mmlf_g_layout(float* nchw, float* nhwc, int ld, int B, int C, int H, int W, int to_nhwc, void* stream);

/* ------------------------------------------------------------------ validation metrics */
/* laplace_to_discrete / lmm_to_discrete (validate/cli.py:91-118): means / logvars (K, B, HW) f32 (K = 1: one Laplacian),
 * out (B, n_bins, HW) f64 = mean over the members of the Laplace-CDF differences over n_bins + 1 edges from
 * x_min - step/2 to x_max + step/2.  float64 arithmetic; var = float32 exp of the float32 logvar, as in the reference. */
int mmlf_lmm_to_discrete(const float* means, const float* logvars, int K, int64_t B, int64_t HW, int n_bins, double x_min,
                         double x_max, double* out, void* stream);
/* kl_divergence (validate/cli.py:174-187): dist / dist_gt (B, S, HW) f64 are epsilon-shifted and normalised IN PLACE
 * like the reference does; value (optional, f64 [B * HW]) = sum_c gt log(gt / dist) per pixel; sums (optional, f64 [2],
 * zeroed by the caller) += (sum value * mask, sum mask), mask (f64 [B * HW]) or NULL for all ones. */
int mmlf_kl_divergence(double* dist, double* dist_gt, int S, int64_t B, int64_t HW, const double* mask, double* value,
                       double* sums, void* stream);
/* nll_discrete (validate/cli.py:51-70): weights /= sum, posterior /= sum * 7 after the epsilon shift (in place),
 * value = sum_c weights * -log(posterior). */
int mmlf_nll_discrete(double* weights, double* posterior, int S, int64_t B, int64_t HW, const double* mask, double* value,
                      double* sums, void* stream);

/* ------------------------------------------------------------------ training augmentation chain */
/* RandomDownSampling -> RandomShift -> RandomCrop(ps + 16) -> CenterCrop(ps) -> RandomRotate -> RedistColor -> Brightness
 * -> Contrast (train/cli.py:78-87; hci4d.py:483-530, 894-1028, 533-664, 1031-1087, 667-785) as gather kernels over
 * scenes resident in HBM, with the per-sample random parameters drawn by the HOST (python `random`, same draw order as
 * the reference) and passed explicitly. */
typedef struct mmlf_aug_sample {
  int scene;             /* index into the scene arrays                                                  */
  int f;                 /* down-sampling factor (x[::f, ::f])                                           */
  int cy, cx;            /* top-left of the final ps x ps patch in the down-sampled image (crop y + 8, x + 8) */
  int r;                 /* number of 90-degree rotations, 0..3                                          */
  int src[4];            /* output stack s (h, v, i, d) reads input stack src[s] ...                      */
  int flip[4];           /* ... with the view order reversed if flip[s]   (Rotate90 swaps h<->v, i<->d)  */
  float w0[16], w1[16];  /* Shift taps of view k in the down-sampled image (mmlf_shift_taps)              */
  int s0[16], s1[16];
  double mat[9];         /* RedistColor matrix, row major                                                */
  float bright;          /* Brightness factor (float32 of the python float)                              */
  float contrast;        /* Contrast factor                                                              */
  float one_minus_contrast; /* float32(1.0 - alpha), alpha the python float                              */
  float disp_f;          /* float32(disp): gt -= disp                                                    */
  double disp;           /* mpi[:, 4] -= disp (float64 array in the reference)                           */
} mmlf_aug_sample;

/* Host helper: fills src / flip / the Shift taps of a sample from (r, disp, n). */
int mmlf_augment_fill(mmlf_aug_sample* s, int r, double disp, int n);

/* Scenes: stacks (S, 4, n, 3, H, W) f32 (h, v, i, d), center (S, 3, H, W) f32, gt (S, H, W) f32, mpi (S, K, 5, H, W) f32,
 * mask (S, H, W) int32.  samples: DEVICE array of B mmlf_aug_sample.
 * Outputs: views (4, B, n, 3, ps, ps) f32 and center (B, 3, ps, ps) f32 up to and including Brightness; view_sums
 * (double[B], zeroed by the caller) receives the sum of the h stack of each sample (for Contrast's mean);
 * gt (B, ps, ps) f32, mpi (B, K, 5, ps, ps) f32 (float64 arithmetic, rounded once), mask (B, ps, ps) int32 (cropped
 * but, like in the reference, NOT rotated). */
int mmlf_augment_patches(const float* stacks, const float* center, const float* gt, const float* mpi,
                         const int32_t* mask, int S, int n, int K, int H, int W, const mmlf_aug_sample* samples, int B,
                         int ps, float* out_views, float* out_center, double* view_sums, float* out_gt, float* out_mpi,
                         int32_t* out_mask, void* stream);
/* Contrast (hci4d.py:739-751): x * alpha + mean * (1 - alpha) in place on views and center; mean[b] = float32(
 * view_sums[b] / (n * 3 * ps * ps)) unless mean_override (f32 [B], e.g. numpy's own pairwise float32 mean) is given. */
int mmlf_augment_contrast(float* views, float* center, const mmlf_aug_sample* samples, const double* view_sums,
                          const float* mean_override, int B, int n, int ps, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MMLF_B200_H */
