/* vaegan_b200.h -- C ABI of libvaegan_b200.so (hand-written sm_100a kernels for the VAE-GAN train step).
 *
 * The reference (Andrey1408/vae-gan-mark) has no native/FFI boundary of its own: its hot path is
 * Python nn.Modules calling ATen/cuDNN.  This header is the seam this project *creates* underneath
 * that nn.Module surface ("lower surface").  Each entry point names the reference
 * call site whose ATen dispatch it replaces.
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless it says "host";
 *   - activations are NHWC bf16 ("pixel rows"): element (n,h,w,c) at ((n*H+h)*W+w)*ld + c_off + c, where
 *     ld >= C lets a tensor live in a channel slice of a wider buffer (fused concat);
 *   - parameters and their gradients are fp32 in PyTorch layouts (OIHW conv, IOHW conv-transpose);
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *   - return 0 on success, negative on error (never throws/aborts); vg_last_error() gives the message
 *     of the calling thread's last failure;
 *   - no hidden allocations: scratch is caller-provided.
 */
#ifndef VAEGAN_B200_H_
#define VAEGAN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VG_API_VERSION 1
#define VG_MAX_TAPS 16          /* kernel positions of one convolution */
#define VG_MAX_FPROP_TAPS 96    /* fprop K-loop entries: taps x split-precision operand combinations */

const char* vg_last_error(void);
int vg_version(void);
int vg_device_info(int* sm_count, int* cc_major, int* cc_minor);   /* host pointers */
unsigned long long vg_launch_count(void);
/* Cap the grid of the persistent tensor-core kernels at n SMs (0 = all) so that a concurrent stream of small,
 * latency-bound kernels (the recurrent text encoder) finds free SMs; host-side setting, read at launch time. */
int vg_set_conv_sm_limit(int n);   /* kernels this library has launched so far in this process */

/* ---------------------------------------------------------------------------------------------
 * Tensor-core implicit GEMMs (tcgen05 / TMEM / TMA)
 *
 * A "tap" addresses the activation operand for one kernel position: {c_base, dw, sh, dh}.
 * The NHWC tensor x[N][H][W][ld] is viewed through stride s (1 or 2) as (s*ld, W/s, s, H/s, N);
 * for output pixel (n, oh, ow) the tap reads channels [c_base + c, ...) at view coordinates
 * (ow + dw, sh, oh + dh, n).  For s = 1: c_base = channel offset, (dh,dw) = (r - pad, q - pad), sh = 0.
 * For s = 2 and input row 2*oh + r - pad = 2*(oh + dh) + sh (same for columns, the column parity
 * being folded into c_base = parity*ld + channel offset).  Out-of-range pixels read as zero.
 * ------------------------------------------------------------------------------------------- */

/* out[pixel][n] = sum_{tap,c} x[pixel@tap][c] * w[n][tap*cin + c]   (+bias, activation)
 * Replaces Conv2d / ConvTranspose2d forward and input-gradient dispatches to cuDNN
 * (vae-gan.py:52-60,76-81,153-157; vae-gan-v2.py:123-127,168-176,199-241; vae-gan-unet.py:148-154,194-221). */
typedef struct VgConvFprop {
  const void* x;            /* bf16 activations */
  int x_n, x_h, x_w, x_ld;  /* physical dims of x and its pixel stride (elements) */
  int x_stride;             /* view stride s: 1 or 2 */
  int m_n, m_h, m_w;        /* output pixel grid (GEMM M = m_n*m_h*m_w) */
  int cin;                  /* channels per tap, multiple of 64 */
  int num_taps;             /* 1..VG_MAX_FPROP_TAPS */
  int taps[VG_MAX_FPROP_TAPS][4]; /* {c_base, dw, sh, dh} */
  int use_wk;               /* 0: tap i reads weight columns [i*cin, (i+1)*cin); 1: columns [wk[i], wk[i]+cin) */
  int wk[VG_MAX_FPROP_TAPS];
  const void* w;            /* bf16 [n_gemm][w_ld], K = tap-major then channel */
  int w_ld;
  int n_gemm;               /* GEMM N */
  void* out;                /* destination */
  int out_kind;             /* 0: bf16 store, 1: fp32 store, 2: fp32 atomic add (split-K) */
  int out_h, out_w, out_ld, out_coff;   /* destination spatial dims, pixel stride, channel offset */
  int su_h, su_w;           /* pixel-shuffle factors: dest pixel = (oh*su_h + dh', ow*su_w + dw') */
  int sub_h0, sub_w0;       /* fixed sub-pixel offset added to (dh', dw') */
  int cout_per_sub;         /* GEMM column n -> sub = n / cout_per_sub (dh' = sub / su_w, dw' = sub % su_w),
                               dest channel = n % cout_per_sub */
  const float* bias;        /* optional fp32 [cout_per_sub] */
  int act;                  /* 0 none, 1 ReLU, 2 LeakyReLU(0.2) */
  int ksplit;               /* 0 = auto; >1 requires out_kind 2 (caller zeroes out) */
  int force_bn;             /* 0 = auto; 64/128/256 forces the N tile (testing) */
  int b_mn_major;           /* 0: w is [n_gemm][K] (K contiguous).  1: w is [w_rows = K per tap][w_ld] with the N index
                               contiguous: GEMM column n of tap t sits at w[k][wk[t] + n] (use_wk required).  This is a
                               conv's FORWARD operand [Cout][taps*Cin] read as B of its own data gradient */
  int w_rows;               /* b_mn_major: rows of w (K extent of one tap; rows beyond read as zero) */
  int num_groups;           /* 0/1: one problem.  2..4: independent problems sharing x, w, M, N and the destination,
                               differing in their taps (listed group after group in taps[] / wk[]) and in the sub-pixel
                               offset they write -- the output-parity classes of a stride-2 data gradient in ONE launch.
                               sub_h0 / sub_w0 are ignored; no split-K */
  int group_ntaps[4];       /* taps of each group; must add up to num_taps */
  int group_sub[4][2];      /* (sub_h0, sub_w0) of each group */
  int halo_mode;            /* 0 = auto (3x3 stride-1 tap sets with n_gemm <= 64 and >= 64K pixels load each 18 x 10
                               activation halo once and address the nine taps as shifted UMMA descriptors; the weights
                               stay resident in shared memory when they fit); -1 = never; 1 = force (testing) */
  float* stats;             /* optional fp32 [2][cout_per_sub], zeroed by the call: per-destination-channel sum and sum of
                               squares of the values as stored (after bias / activation, rounded to the output type) over
                               all valid pixels -- the batch statistics of the BatchNorm2d that follows the convolution
                               (vae-gan-v2.py:172-177, vae-gan.py:52-55,77-80), produced by the epilogue instead of a
                               separate pass over the tensor.  Needs out_kind 0/1, no split-K, an aligned destination */
} VgConvFprop;
int vg_conv_fprop(const VgConvFprop* desc /*host*/, void* stream);
/* CTA-pair (tcgen05 cta_group::2, clusters of two CTAs, M = 256 MMAs, half the weight tile per CTA) variant of the
 * forward kernel for wide layers (256-column N tiles, K-major weights) on (default) / off */
int vg_set_fprop_cta_pairs(int on);

/* dw[co][tap*cin + ci] = sum_pixels g[pixel][co] * x[pixel@tap][ci]      (fp32 result)
 * Replaces Conv2d / ConvTranspose2d weight-gradient dispatches to cuDNN (same call sites, backward). */
typedef struct VgConvWgrad {
  const void* g;            /* bf16 output-side gradient, NHWC over the (m_n, m_h, m_w) grid */
  int g_ld, g_coff;
  int cout;                 /* rows of dw */
  const void* x;            /* bf16 input-side activations */
  int x_n, x_h, x_w, x_ld, x_stride;
  int m_n, m_h, m_w;
  int cin, num_taps;
  int taps[VG_MAX_TAPS][4];
  int num_combos;           /* 0/1: plain; >1: split-precision operand pairs accumulated into the same dw */
  int combo_g[8], combo_x[8]; /* channel offsets added to the g / x channel coordinates for each pair */
  float* dw;                /* fp32 [cout][dw_ld]; overwritten (zeroed internally when split) */
  int dw_ld;
  int ksplit;               /* 0 = auto */
  int force_bn;             /* 0 = auto; 64/128/192/256 */
  void* workspace;          /* optional scratch for a DETERMINISTIC split reduction: each split stores its partial tile
                               here and a second kernel adds them in a fixed order (bit-identical results from run to run);
                               NULL: the splits are combined with fp32 atomics in whatever order they finish */
  long long workspace_bytes; /* >= vg_conv_wgrad_workspace(desc) */
} VgConvWgrad;
int vg_conv_wgrad(const VgConvWgrad* desc /*host*/, void* stream);
/* CTA-pair (tcgen05 cta_group::2, clusters of two CTAs) variant of the weight-gradient kernel on (default) / off */
int vg_set_cta_pairs(int wgrad_on);
long long vg_conv_wgrad_workspace(const VgConvWgrad* desc /*host*/);   /* bytes; 0 if the launch does not split; < 0 bad desc */


/* ---------------------------------------------------------------------------------------------
 * Normalisation + activation (+ fused 2x2 max-pool).  groups = 1: BatchNorm2d (training) statistics over all
 * rows; groups = N: InstanceNorm2d statistics per sample.  act: 0 none, 1 ReLU, 2 LeakyReLU(0.2).
 * Replaces nn.BatchNorm2d + nn.ReLU (+ nn.MaxPool2d) of the generator (vae-gan.py:52-55,76-81;
 * vae-gan-v2.py:157-177,236-242) and nn.InstanceNorm2d + nn.LeakyReLU of D (vae-gan.py:154-156).
 * ------------------------------------------------------------------------------------------- */
/* sums[g][0][c] = sum x, sums[g][1][c] = sum x^2 (fp32, zeroed internally) */
int vg_norm_stats(const void* x, int x_ld, int x_coff, int groups, long long rows_per_group, int c, float* sums,
                  int dtype /*0 bf16, 1 fp32 activations*/, void* stream);
/* Batch statistics of a [n][h][w][c] tensor whose h (>= 3) rows stand for virt_h rows: row 0 and row h-1 count once,
 * interior rows (virt_h-2)/(h-2) times.  Used for the FiLM parameter maps (vae-gan-v2.py:138-145), whose rows are all
 * equal except the first and the last, so that 3 rows carry the whole map; finalize with rows = n*virt_h*w. */
int vg_norm_stats_rows(const void* x, int x_ld, int x_coff, int n, int h, int w, int c, int virt_h, float* sums,
                       int dtype, void* stream);
/* mean_rstd[g][0][c] = mean, [g][1][c] = 1/sqrt(var+eps); when groups == 1 and running_mean != NULL also updates the
 * running statistics (momentum, unbiased variance) and increments *num_batches_tracked (int64, nullable). */
int vg_norm_finalize(const float* sums, int groups, long long rows_per_group, int c, float eps, float* mean_rstd,
                     float momentum, float* running_mean, float* running_var, long long* num_batches_tracked,
                     void* stream);
typedef struct VgNormApply {
  const void* x; int x_ld, x_coff;      /* raw conv output, bf16 NHWC */
  int n, h, w, c;
  const float* mean_rstd; int per_sample;
  const float* gamma; const float* beta; /* nullable */
  int act;
  void* y; int y_ld, y_coff;            /* full-resolution output (may be a channel slice of a concat buffer) */
  void* pool; int p_ld, p_coff;         /* optional 2x2 max-pooled output (NULL = none) */
  int dtype;                            /* activation storage: 0 bf16, 1 fp32 (high-accuracy mode) */
} VgNormApply;
int vg_norm_apply(const VgNormApply* desc /*host*/, void* stream);
typedef struct VgNormBackward {
  const void* x; int x_ld, x_coff;      /* raw conv output saved by the forward */
  const void* dy; int dy_ld, dy_coff;   /* grad wrt y (nullable) */
  const void* dpool; int dp_ld, dp_coff;/* grad wrt the pooled output (nullable) */
  int n, h, w, c;
  const float* mean_rstd; int per_sample;
  const float* gamma; const float* beta;
  int act;
  float* sums;                          /* scratch fp32 [groups][2][c] */
  void* dx; int dx_ld, dx_coff;         /* grad wrt x, bf16 */
  float* dgamma; float* dbeta;          /* fp32 [c], nullable */
  int accumulate;                       /* add into dgamma/dbeta instead of overwriting */
  int dtype;                            /* activation storage: 0 bf16, 1 fp32 */
  int virt_h;                           /* 0, or: the h (>= 3) rows stand for virt_h rows whose interior rows are all
                                           equal (see vg_norm_stats_rows); dy then holds per-class SUMS of gradients */
} VgNormBackward;
int vg_norm_backward(const VgNormBackward* desc /*host*/, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Data movement / elementwise
 * ------------------------------------------------------------------------------------------- */
/* out[sum k_d*out_strides[d]] (=|+=) scale * in[sum k_d*in_strides[d]] over a 5-D index space (dims, host arrays;
 * dtype codes 0 = fp32, 1 = bf16; scale is an optional DEVICE scalar, inverted when scale_inverse).  Used for the
 * fp32 OIHW <-> bf16 GEMM-layout weight/gradient re-layouts, NCHW fp32 <-> NHWC bf16 at the module boundary
 * (torch.cat of image+mask vae-gan.py:139, z/text concat vae-gan-v2.py:249-251) and W/sigma of spectral norm. */
int vg_strided_copy(const void* in, int in_dtype, void* out, int out_dtype, const long long* dims,
                    const long long* in_strides, const long long* out_strides, const float* scale, int scale_inverse,
                    int accumulate, void* stream);

/* Test hook: walks the tiled-copy plan of vg_strided_copy on fp32 HOST buffers (no GPU needed) so the tiling logic
 * can be checked on CPU.  Returns the tile count; 0 means the plan falls back to the generic kernel. */
long long vg_debug_copy_plan_host(const float* in, float* out, const long long* dims, const long long* in_strides,
                                  const long long* out_strides);
/* MaxPool2d(2, 2) on NHWC activations [n][h][w][c] (pixel stride x_ld), y / dy dense [n][h/2][w/2][c], dx dense
 * [n][h][w][c] (fully written; the gradient goes to the first maximum of each window).  The VGG16 feature extractor of
 * the perceptual loss (vae-gan.py:300-311) is the user; the U-Net's pools are fused into the normalisation kernels. */
int vg_maxpool2x2_fwd(const void* x, int x_ld, void* y, int n, int h, int w, int c, int dtype, void* stream);
int vg_maxpool2x2_bwd(const void* x, int x_ld, const void* dy, void* dx, int n, int h, int w, int c, int dtype, void* stream);

/* dx = dy * act'(y) for an activation that was fused into a conv epilogue (act 1 ReLU, 2 LeakyReLU(0.2)); bf16 rows */
int vg_act_bwd(const void* y, int y_ld, const void* dy, int dy_ld, void* dx, int dx_ld, long long rows, int c, int act,
               int dtype, void* stream);
/* y = act(y) in place (where the activation cannot be fused into the producing epilogue: split-K launches) */
int vg_act_fwd(void* y, int y_ld, long long rows, int c, int act, int dtype, void* stream);
/* out[c] (=|+=) sum_r in[r*ld + c], fp32 (bias gradients of small matrices) */
int vg_colsum_f32(const float* in, long long rows, int cols, int ld, float* out, int accumulate, void* stream);
/* FiLM (vae-gan-v2.py:146-149): y = gb[:, :c] * x + gb[:, c:]; gb and y dense bf16 [rows][2c] / [rows][c] */
int vg_film_fwd(const void* gb, const void* x, int x_ld, int x_coff, void* y, long long rows, int c, int dtype,
                void* stream);
int vg_film_bwd(const void* gb, const void* x, int x_ld, int x_coff, const void* dy, void* dgb, void* dx, int dx_ld,
                int dx_coff, long long rows, int c, int dtype, void* stream);
/* FiLM whose parameter map gb has gh = 3 rows per image (first row | any interior row | last row): the maps of
 * vae-gan-v2.py:138-145 are row-constant away from the zero-padded border.  y = gamma(class of row) * x + beta(...).
 * The backward writes dx and, into dgb [n][3][w][2c], the gradient SUMMED over the rows of each class. */
int vg_film_rows_fwd(const void* gb, int gh, const void* x, int x_ld, int x_coff, void* y, int n, int h, int w, int c,
                     int dtype, void* stream);
int vg_film_rows_bwd(const void* gb, int gh, const void* x, int x_ld, int x_coff, const void* dy, void* dgb, void* dx,
                     int dx_ld, int dx_coff, int n, int h, int w, int c, int dtype, void* stream);
/* F.interpolate(bilinear, align_corners=False) of a (1 x w0) map to (h x w) (vae-gan-v2.py:138-140) */
int vg_upsample_w_fwd(const void* t, int t_ld, int t_coff, int n, int w0, int c, void* y, int h, int w, int dtype,
                      void* stream);
int vg_upsample_w_bwd(const void* dy, int n, int h, int w, int c, int w0, float* dt /*fp32 [n][w0][c]*/, int dtype,
                      void* stream);
/* Bilinear resize along H only (align_corners=False); with the W-only pair above it composes F.interpolate(bilinear)
 * of a multi-row map (the 4-row text map of vae-gan-oldv.py:165-176,286-291).  t / dt: [n][h0][w][c], y / dy:
 * [n][h][w][c], all dense; dt is fp32 and fully written. */
int vg_upsample_h_fwd(const void* t, int n, int h0, int w, int c, void* y, int h, int dtype, void* stream);
int vg_upsample_h_bwd(const void* dy, int n, int h, int w, int c, int h0, float* dt, int dtype, void* stream);
/* Per-channel gate y = x * scale[c] (GatedSkipConnection, vae-gan-oldv.py:226-231; scale = sigmoid(alpha), fp32 [c]),
 * written into a channel slice (y_ld, y_coff).  Backward in one pass: dx = dy * scale, dscale[0..c) = sum dy * x
 * (dscale is fp32 [2*c]; the second half is scratch). */
int vg_channel_scale_fwd(const void* x, int x_ld, const float* scale, void* y, int y_ld, int y_coff, long long rows, int c,
                         int dtype, void* stream);
int vg_channel_scale_bwd(const void* x, int x_ld, const void* dy, int dy_ld, int dy_coff, const float* scale, void* dx,
                         int dx_ld, long long rows, int c, float* dscale, int dtype, void* stream);
/* `dtype` on the activation-touching entry points selects the activation storage: 0 = bf16 (default), 1 = fp32 (the
 * high-accuracy mode, in which the tensor core is fed three bf16 planes per fp32 operand, see vg_split3). */
/* fp32 [rows][c] (row stride ld_in) -> bf16 [rows][3*cp]: planes hi | mid | lo with hi+mid+lo == x to ~2^-24 */
int vg_split3(const float* in, int ld_in, long long rows, int c, int cp, void* out, void* stream);
/* im2col / col2im for few-channel images (first conv of the encoder vae-gan-v2.py:154 and of D vae-gan.py:153) */
int vg_im2col(const void* src, int n, int h, int w, int ld, int c, int kh, int kw, int stride, int pad, void* col,
              int kpad, int dtype, void* stream);
int vg_col2im(const void* dcol, int kpad, int n, int h, int w, int c, int kh, int kw, int stride, int pad,
              float* dsrc_nchw, int dtype, void* stream);
/* direct stride-1 convs with <= 4 output channels; w fp32 [cout][kh][kw][cin]; out/dy fp32 NHWC [n][oh][ow][cout]
 * (final_image_conv vae-gan-v2.py:232, decode.15 vae-gan.py:81, D's patch head vae-gan.py:157) */
int vg_conv_smalln_fwd(const void* x, int x_ld, int x_coff, int n, int h, int w, int cin, const float* wt,
                       const float* bias, int cout, int kh, int kw, int pad, float* out, int dtype, void* stream);
int vg_conv_smalln_dgrad(const float* dy, int n, int h, int w, int cin, const float* wt, int cout, int kh, int kw,
                         int pad, void* dx, int dx_ld, int dx_coff, int dtype, void* stream);
int vg_conv_smalln_wgrad(const float* dy, const void* x, int x_ld, int x_coff, int n, int h, int w, int cin, int cout,
                         int kh, int kw, int pad, float* dw, float* dbias, int dtype, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Recurrent text encoder: the time recurrence of one bidirectional GRU layer (torch.nn.GRU semantics, gate order
 * r, z, n; replaces the cuDNN RNN the reference reaches through nn.GRU, vae-gan-v2.py:84-89,105).  One launch
 * walks all `steps` time steps of both directions; hidden must be 256.  All tensors fp32, device, contiguous:
 *   xproj [batch][steps][2][3*hidden]  x W_ih^T + b_ih of direction 0 | direction 1 (a time-parallel GEMM, caller's)
 *   w_hh  [2][3*hidden][hidden], b_hh [2][3*hidden]
 *   out   [batch][steps][2*hidden]     h_t of direction d in columns [d*hidden, (d+1)*hidden)  (h_0 = 0)
 *   gates [2][batch][steps][4][hidden] r, z, n and (W_hn h + b_hn), saved for the backward (may be NULL)
 * backward: dout = gradient of `out`; writes dgx [batch][steps][2][3*hidden] (gradient w.r.t. xproj) and
 *   dgh [2][batch][steps][3*hidden] (gradient w.r.t. h W_hh^T + b_hh); weight gradients are GEMMs of these. */
int vg_gru_seq_fwd(const float* xproj, const float* w_hh, const float* b_hh, float* out, float* gates, int batch,
                   int steps, int hidden, void* stream);
int vg_gru_seq_bwd(const float* dout, const float* out, const float* gates, const float* w_hh, float* dgx, float* dgh,
                   int batch, int steps, int hidden, void* stream);
/* How many 8-CTA clusters of the forward (backward != 0: backward) kernel the device keeps resident at once, for
 * batch_group = 8 or 16 rows per cluster; the launchers use the smallest group that runs in a single wave. */
int vg_gru_max_active_clusters(int backward, int batch_group);

/* Front end of CharacterTokenEncoder on the device (vae-gan-v2.py:65-114; the reference tokenises with a Python loop and a
 * dict lookup per character on the host, :89-100, and goes through ATen's embedding / adaptive_avg_pool1d kernels):
 *   vg_tokenize       idx[i] = lut[codepoints[i]] (0 = padding for code points >= lut_size); codepoints are the UTF-32
 *                     code units of the strings, zero padded to max_len (uint32 [n]); idx int64 [n]
 *   vg_embedding_fwd  out[t][:] = weight[idx[t]][:]            (fp32 [vocab][dim] -> fp32 [n][dim])
 *   vg_embedding_bwd  dw[r][:] = sum_{t: idx[t] == r} g[t][:], row padding_idx = 0; fixed summation order (deterministic),
 *                     every row of dw written
 *   vg_seqpool_fwd    adaptive average pooling of a sequence [b][l][c] (row stride in_ld) along l into [b][w][c] (row
 *                     stride out_ld) -- i.e. the NHWC text map [b][1][w][c]; bins as torch.nn.AdaptiveAvgPool1d
 *   vg_seqpool_bwd    its adjoint: dy [b][w][c] -> dseq [b][l][c] (fully written)
 * dtype codes: 0 = bf16, 1 = fp32. */
int vg_tokenize(const uint32_t* codepoints, long long n, const int* lut, int lut_size, long long* idx, void* stream);
int vg_embedding_fwd(const long long* idx, long long n, const float* weight, int vocab, int dim, float* out, void* stream);
int vg_embedding_bwd(const long long* idx, long long n, const float* g, int vocab, int dim, int padding_idx, float* dw,
                     void* stream);
int vg_seqpool_fwd(const void* seq, int in_dtype, int in_ld, int b, int l, int c, int w, void* out, int out_dtype,
                   int out_ld, void* stream);
int vg_seqpool_bwd(const void* dy, int dy_dtype, int dy_ld, int b, int l, int c, int w, void* dseq, int out_dtype,
                   int out_ld, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Losses, reparameterisation, spectral norm, optimiser (fp32)
 * ------------------------------------------------------------------------------------------- */
/* heads fp32 [b][2z] (mu | logvar before bias); writes mu, logvar, z = mu + eps*exp(logvar/2) and
 * *kl_out = mean_b(-0.5 mean_c(1 + lv - mu^2 - e^lv))   (vae-gan.py:133-136, :420) */
int vg_reparam_kl_fwd(const float* heads, const float* bias_mu, const float* bias_lv, const float* eps, int b, int z,
                      float* mu, float* logvar, float* zout, float* kl_out, void* stream);
/* dheads[b][2z] from dz, optional external dmu/dlogvar and the scalar dkl (device); optional bf16 copy (row stride bf_ld) */
int vg_reparam_kl_bwd(const float* mu, const float* logvar, const float* eps, const float* dz, const float* dmu_ext,
                      const float* dlv_ext, const float* dkl, int b, int z, float* dheads, void* dheads_bf16, int bf_ld,
                      void* stream);
int vg_sigmoid_fwd(const float* pre_nhwc, int n, int c, int hw, float* y_nchw, void* stream);      /* vae-gan.py:82 */
int vg_sigmoid_bwd(const float* y_nchw, const float* dy_nchw, int n, int c, int hw, float* dpre_nhwc, void* stream);
int vg_l1_fwd(const float* a, const float* b, long long n, float* out, void* stream);              /* vae-gan.py:419 */
int vg_l1_bwd(const float* a, const float* b, long long n, const float* gout, float* da, int accumulate, void* stream);
/* hinge_loss (vae-gan.py:313-320): mode 1 real, 0 fake, 2 generator */
int vg_hinge_fwd(const float* p, long long n, int mode, float* out, void* stream);
int vg_hinge_bwd(const float* p, long long n, int mode, const float* gout, float* dp, void* stream);
/* spectral_norm (vae-gan.py:153-156): one power iteration in training mode (u, v updated in place), sigma = u.Wv;
 * scratch fp32 [rows + cols + 2] */
int vg_spectral_sigma(const float* w, int rows, int cols, float* u, float* v, int training, float eps, float* sigma,
                      float* scratch, void* stream);
/* dw_orig (=|+=) g/sigma - (<g, w_orig>/sigma^2) u v^T ; scratch fp32 [1] */
int vg_spectral_bwd(const float* g, const float* w_orig, const float* u, const float* v, const float* sigma, int rows,
                    int cols, float* dw, int accumulate, float* scratch, void* stream);
/* clip_grad_norm_ + Adam over flat buffers (vae-gan.py:424, :541-542): *out (+)= sum g^2; the Adam pass scales g by
 * min(1, max_norm/(sqrt(*gnorm_sq)+1e-6)) when gnorm_sq != NULL and max_norm > 0 */
int vg_sumsq(const float* g, long long n, float* out, int zero_first, void* stream);
/* multi-tensor forms: `table` is a DEVICE array of `count` entries (g == NULL entries are skipped); `state` is a
 * device float[4] {step, 1-beta1^step, sqrt(1-beta2^step), lr} advanced by vg_adam_prepare, so a captured CUDA graph of
 * the step replays with the right bias corrections; vg_multi_adam with lr < 0 reads the learning rate from state[3], so
 * a scheduler (ReduceLROnPlateau of vae-gan-lr-sh.py:751-760, vae-gan-v2.py:944-953) can change it between replays */
/* shadow (nullable): bf16 [n] in the memory order of p, rewritten by vg_multi_adam with the updated values -- for conv
 * weights kept in [Cout][kh][kw][Cin] order it IS the tensor-core operand of the next step's forward, so the per-step
 * fp32 -> bf16 operand conversion disappears */
typedef struct VgAdamTensor { float* p; float* g; float* m; float* v; long long n; void* shadow; } VgAdamTensor;
int vg_adam_prepare(float* state, float beta1, float beta2, void* stream);
/* deterministic (fixed-order) reduction; scratch: device float[scratch_len], scratch_len >= 1 (use >= 4 x #SMs) */
int vg_multi_sumsq(const VgAdamTensor* table, int count, float* out, float* scratch, int scratch_len, void* stream);
int vg_multi_adam(const VgAdamTensor* table, int count, float lr, float beta1, float beta2, float eps,
                  const float* state, const float* gnorm_sq, float max_norm, int write_back_grad, void* stream);
int vg_adam_step(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2, float eps,
                 int step, const float* gnorm_sq, float max_norm, int write_back_grad, void* stream);

/* ---- data path: patch extraction (perspective_crop + T.ToTensor, vae-gan.py:163-188,275-281; the same function in every
 * script).  Restates cv2.getPerspectiveTransform + cv2.warpPerspective(INTER_LINEAR, BORDER_REPLICATE) for 8-bit images
 * byte for byte (OpenCV's fixed-point arithmetic).  vg_perspective_crop_matrix runs on the HOST: bbox = 4 (x, y) float
 * pairs, result = the inverse (destination -> source) map, 9 doubles.  vg_warp_perspective_u8: src is a DEVICE uint8
 * image [src_h][src_w][channels] with src_row_bytes per row, minv the host matrix; writes the uint8 patch
 * [out_h][out_w][channels] and / or the float32 tensor [channels][out_h][out_w] = value / 255 (either may be NULL). */
int vg_perspective_crop_matrix(const float* bbox, int out_w, int out_h, double* minv);
int vg_warp_perspective_u8(const unsigned char* src, int src_h, int src_w, int channels, long long src_row_bytes,
                           const double* minv, int out_h, int out_w, unsigned char* dst_u8, float* dst_chw, void* stream);
/* general form: inverse (destination -> source) map of cv2.getPerspectiveTransform(src_quad -> dst_quad), HOST */
int vg_perspective_matrix(const float* src_quad, const float* dst_quad, double* minv);
/* perspective_unwarp (vae-gan.py:190-200): the generated patch pasted back into the page.  vg_perspective_unwarp_matrix
 * (HOST): patch rectangle [0, patch_w - 1] x [0, patch_h - 1] -> bbox.  vg_warp_perspective_u8_transparent: the warp with
 * borderMode = BORDER_TRANSPARENT into the caller's DEVICE canvas [out_h][out_w][channels] uint8 (the reference zero-fills
 * it first): only destination pixels whose source coordinate falls inside the patch are written (cv2 4.13 semantics). */
int vg_perspective_unwarp_matrix(const float* bbox, int patch_w, int patch_h, double* minv);
int vg_warp_perspective_u8_transparent(const unsigned char* src, int src_h, int src_w, int channels, long long src_row_bytes,
                                       const double* minv, int out_h, int out_w, unsigned char* canvas_u8, void* stream);
/* a whole batch of warps (the patches of one training batch: vae-gan.py:268-283 runs perspective_crop three times per
 * sample) in ONE launch, one grid row per job.  The caller owns the table: `jobs_host` (validated here) and an identical
 * copy `jobs_device` in device memory, made on `stream` before the call. */
typedef struct VgWarpJob {
  const unsigned char* src; int src_h, src_w, channels; long long src_row_bytes;
  double minv[9];
  int out_h, out_w;
  unsigned char* dst_u8; float* dst_chw;      /* either may be NULL (not both) */
  int transparent;                            /* 0: BORDER_REPLICATE (crop), 1: BORDER_TRANSPARENT into dst_u8 (unwarp) */
} VgWarpJob;
int vg_warp_perspective_u8_batch(const VgWarpJob* jobs_host, const VgWarpJob* jobs_device, int count, void* stream);
/* test hooks: the kernel's per-pixel code on HOST buffers (not a fallback; nothing in the package calls them) */
int vg_debug_warp_perspective_host(const unsigned char* src, int src_h, int src_w, int channels, long long src_row_bytes,
                                   const double* minv, int out_h, int out_w, unsigned char* dst_u8, float* dst_chw);
int vg_debug_warp_perspective_transparent_host(const unsigned char* src, int src_h, int src_w, int channels,
                                               long long src_row_bytes, const double* minv, int out_h, int out_w,
                                               unsigned char* canvas_u8);

#ifdef __cplusplus
}
#endif
#endif /* VAEGAN_B200_H_ */
