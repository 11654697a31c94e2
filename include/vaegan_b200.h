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
#define VG_MAX_TAPS 16

const char* vg_last_error(void);
int vg_version(void);
int vg_device_info(int* sm_count, int* cc_major, int* cc_minor);   /* host pointers */

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
  int num_taps;             /* 1..VG_MAX_TAPS */
  int taps[VG_MAX_TAPS][4]; /* {c_base, dw, sh, dh} */
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
} VgConvFprop;
int vg_conv_fprop(const VgConvFprop* desc /*host*/, void* stream);

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
  float* dw;                /* fp32 [cout][dw_ld]; overwritten (zeroed internally when split) */
  int dw_ld;
  int ksplit;               /* 0 = auto */
  int force_bn;             /* 0 = auto; 64/128/192/256 */
} VgConvWgrad;
int vg_conv_wgrad(const VgConvWgrad* desc /*host*/, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VAEGAN_B200_H_ */
