// Patch extraction of the data path: perspective_crop + T.ToTensor of the reference (vae-gan.py:163-188, 275-281),
// i.e. cv2.getPerspectiveTransform + cv2.warpPerspective(INTER_LINEAR, BORDER_REPLICATE) on 8-bit images.
// Also the inverse direction, perspective_unwarp (vae-gan.py:190-200): the same warp with borderMode=BORDER_TRANSPARENT into
// a caller-provided canvas, and a batched launch over a device table of jobs (one grid row per patch) for a whole training
// batch of crops.
// Integer / byte work: the arithmetic below restates OpenCV's (imgwarp.cpp: fixed-point coordinates with 5 fractional
// bits, 15-bit bilinear weights, blocks of 64 destination columns) so that the bytes are identical; oracle/warp.py is
// the numpy restatement pinned against cv2, tests/test_warp_*.py hold this file to it.
#include <cmath>
#include <cstdint>

#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

struct WarpMat { double v[9]; };

constexpr int kWarpBlockW = 64;     // WarpPerspectiveInvoker: BLOCK_SZ * BLOCK_SZ / min(BLOCK_SZ / 2, height) columns per block

#define VG_HD __host__ __device__ __forceinline__
// Explicitly rounded double operations: intrinsics on the device (no fused multiply-add contraction), separately
// stored results on the host (the host twin below lets the CPU test suite check the per-pixel code without a GPU).
VG_HD double dmul(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dmul_rn(a, b);
#else
  volatile double r = a * b; return r;
#endif
}
VG_HD double dadd(double a, double b) {
#ifdef __CUDA_ARCH__
  return __dadd_rn(a, b);
#else
  volatile double r = a + b; return r;
#endif
}
VG_HD double ddiv(double a, double b) {
#ifdef __CUDA_ARCH__
  return __ddiv_rn(a, b);
#else
  volatile double r = a / b; return r;
#endif
}
VG_HD int d2i_rn(double a) {          // cvRound: to nearest, ties to even; `a` is already clamped to the int range
#ifdef __CUDA_ARCH__
  return __double2int_rn(a);
#else
  return static_cast<int>(std::nearbyint(a));
#endif
}
VG_HD float to_unit(int v) {          // T.ToTensor(): float32(v) / 255, IEEE division
#ifdef __CUDA_ARCH__
  return __fdiv_rn(static_cast<float>(v), 255.f);
#else
  volatile float r = static_cast<float>(v) / 255.f; return r;
#endif
}

// One destination pixel.  Source coordinate in 1/32 pixel units in OpenCV's operation order: X0 = M0*bx + M1*y + M2 per
// 64-column block, then X = cvRound((X0 + M0*x1) * (32 / (W0 + M6*x1))) clamped to int; bilinear blend with 15-bit weights.
VG_HD void warp_pixel(const unsigned char* __restrict__ src, int sh, int sw, int ch, long long row_bytes, const WarpMat& m,
                      int oh, int ow, int x, int y, unsigned char* __restrict__ dst_u8, float* __restrict__ dst_chw,
                      int transparent = 0) {
  const int bx = (x / kWarpBlockW) * kWarpBlockW;
  const double dbx = static_cast<double>(bx), dy = static_cast<double>(y), dx1 = static_cast<double>(x - bx);
  const double X0 = dadd(dadd(dmul(m.v[0], dbx), dmul(m.v[1], dy)), m.v[2]);
  const double Y0 = dadd(dadd(dmul(m.v[3], dbx), dmul(m.v[4], dy)), m.v[5]);
  const double W0 = dadd(dadd(dmul(m.v[6], dbx), dmul(m.v[7], dy)), m.v[8]);
  double W = dadd(W0, dmul(m.v[6], dx1));
  W = (W != 0.0) ? ddiv(32.0, W) : 0.0;
  const double fX = fmax(-2147483648.0, fmin(2147483647.0, dmul(dadd(X0, dmul(m.v[0], dx1)), W)));
  const double fY = fmax(-2147483648.0, fmin(2147483647.0, dmul(dadd(Y0, dmul(m.v[3], dx1)), W)));
  const int X = d2i_rn(fX), Y = d2i_rn(fY);
  int sx = X >> 5, sy = Y >> 5;                                                     // arithmetic shifts
  sx = sx < -32768 ? -32768 : (sx > 32767 ? 32767 : sx);                            // stored as short
  sy = sy < -32768 ? -32768 : (sy > 32767 ? 32767 : sy);
  // BORDER_TRANSPARENT (cv2 4.13, pinned in tests/test_warp_oracle.py): the destination pixel is written iff the integer
  // part of its source coordinate lies inside the source; its value is the BORDER_REPLICATE one
  if (transparent && (sx < 0 || sx > sw - 1 || sy < 0 || sy > sh - 1)) return;
  const int ax = X & 31, ay = Y & 31;
  const int x0 = sx < 0 ? 0 : (sx > sw - 1 ? sw - 1 : sx), x1 = sx + 1 < 0 ? 0 : (sx + 1 > sw - 1 ? sw - 1 : sx + 1);   // BORDER_REPLICATE
  const int y0 = sy < 0 ? 0 : (sy > sh - 1 ? sh - 1 : sy), y1 = sy + 1 < 0 ? 0 : (sy + 1 > sh - 1 ? sh - 1 : sy + 1);
  // 15-bit weights (32-ay)(32-ax)*32 ...: exact, they sum to 1 << 15
  const int w00 = (32 - ay) * (32 - ax), w01 = (32 - ay) * ax, w10 = ay * (32 - ax), w11 = ay * ax;
  const unsigned char* r0 = src + static_cast<long long>(y0) * row_bytes;
  const unsigned char* r1 = src + static_cast<long long>(y1) * row_bytes;
  for (int c = 0; c < ch; ++c) {
    const int acc = r0[x0 * ch + c] * w00 + r0[x1 * ch + c] * w01 + r1[x0 * ch + c] * w10 + r1[x1 * ch + c] * w11;
    const int v = (acc * 32 + (1 << 14)) >> 15;
    if (dst_u8 != nullptr) dst_u8[(static_cast<long long>(y) * ow + x) * ch + c] = static_cast<unsigned char>(v);
    if (dst_chw != nullptr) dst_chw[(static_cast<long long>(c) * oh + y) * ow + x] = to_unit(v);
  }
}

__global__ void warp_perspective_u8_kernel(const unsigned char* __restrict__ src, int sh, int sw, int ch, long long row_bytes,
                                           const WarpMat m, int oh, int ow, unsigned char* __restrict__ dst_u8,
                                           float* __restrict__ dst_chw, int transparent) {
  const int total = oh * ow;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / ow;
    warp_pixel(src, sh, sw, ch, row_bytes, m, oh, ow, idx - y * ow, y, dst_u8, dst_chw, transparent);
  }
}

// one grid row (blockIdx.y) per job of a device-resident table: a whole batch of patches in ONE launch
__global__ void warp_perspective_u8_batch_kernel(const VgWarpJob* __restrict__ jobs) {
  const VgWarpJob& j = jobs[blockIdx.y];
  WarpMat m;
#pragma unroll
  for (int i = 0; i < 9; ++i) m.v[i] = j.minv[i];
  const int total = j.out_h * j.out_w;
  for (int idx = blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += gridDim.x * blockDim.x) {
    const int y = idx / j.out_w;
    warp_pixel(j.src, j.src_h, j.src_w, j.channels, j.src_row_bytes, m, j.out_h, j.out_w, idx - y * j.out_w, y, j.dst_u8, j.dst_chw,
               j.transparent);
  }
}

}  // namespace vg

using namespace vg;

// cv::getPerspectiveTransform(bbox -> output rectangle) followed by cv::invert, on the host, operation for operation:
// the products of the system matrix are float32 products (Point2f members), the 8x8 system is solved by Gaussian
// elimination with partial pivoting (hal::LU64f), the 3x3 inverse is cofactors times the reciprocal determinant.
// Inverse (destination -> source) map of cv::getPerspectiveTransform(src_quad -> dst_quad): 4 (x, y) float pairs each.
extern "C" int vg_perspective_matrix(const float* src_quad, const float* dst_quad, double* minv) {
  VG_CHECK(src_quad != nullptr && dst_quad != nullptr && minv != nullptr, -1, "vg_perspective_matrix: bad arguments");
  double a[8][8], b[8];
  for (int i = 0; i < 8; ++i) for (int j = 0; j < 8; ++j) a[i][j] = 0.0;
  for (int i = 0; i < 4; ++i) {
    const float sx = src_quad[2 * i], sy = src_quad[2 * i + 1], dx = dst_quad[2 * i], dy = dst_quad[2 * i + 1];
    a[i][0] = a[i + 4][3] = sx;
    a[i][1] = a[i + 4][4] = sy;
    a[i][2] = a[i + 4][5] = 1.0;
    volatile float p0 = -sx * dx, p1 = -sy * dx, p2 = -sx * dy, p3 = -sy * dy;      // float32 products
    a[i][6] = p0; a[i][7] = p1; a[i + 4][6] = p2; a[i + 4][7] = p3;
    b[i] = dx; b[i + 4] = dy;
  }
  const int n = 8;
  for (int i = 0; i < n; ++i) {
    int k = i;
    for (int j = i + 1; j < n; ++j)
      if (std::fabs(a[j][i]) > std::fabs(a[k][i])) k = j;
    VG_CHECK(std::fabs(a[k][i]) >= 2.220446049250313e-16 * 100, -4, "vg_perspective_matrix: degenerate quadrilateral");
    if (k != i) {
      for (int j = i; j < n; ++j) { const double t = a[i][j]; a[i][j] = a[k][j]; a[k][j] = t; }
      const double t = b[i]; b[i] = b[k]; b[k] = t;
    }
    const double d = -1.0 / a[i][i];
    for (int j = i + 1; j < n; ++j) {
      const double alpha = a[j][i] * d;
      for (int c = i + 1; c < n; ++c) { volatile double pr = alpha * a[i][c]; a[j][c] = a[j][c] + pr; }
      volatile double pb = alpha * b[i];
      b[j] = b[j] + pb;
    }
  }
  for (int i = n - 1; i >= 0; --i) {
    double s = b[i];
    for (int c = i + 1; c < n; ++c) { volatile double pr = a[i][c] * b[c]; s = s - pr; }
    b[i] = s / a[i][i];
  }
  const double M[3][3] = {{b[0], b[1], b[2]}, {b[3], b[4], b[5]}, {b[6], b[7], 1.0}};
#define S(i, j) M[i][j]
  // volatile temporaries keep every product a separately rounded double (no contraction into fused multiply-adds)
  auto mul = [](double x, double y) { volatile double r = x * y; return static_cast<double>(r); };
  const double det = mul(S(0, 0), mul(S(1, 1), S(2, 2)) - mul(S(1, 2), S(2, 1))) -
                     mul(S(0, 1), mul(S(1, 0), S(2, 2)) - mul(S(1, 2), S(2, 0))) +
                     mul(S(0, 2), mul(S(1, 0), S(2, 1)) - mul(S(1, 1), S(2, 0)));
  VG_CHECK(det != 0.0, -4, "vg_perspective_matrix: singular transform");
  const double d = 1.0 / det;
  minv[0] = mul(mul(S(1, 1), S(2, 2)) - mul(S(1, 2), S(2, 1)), d);
  minv[1] = mul(mul(S(0, 2), S(2, 1)) - mul(S(0, 1), S(2, 2)), d);
  minv[2] = mul(mul(S(0, 1), S(1, 2)) - mul(S(0, 2), S(1, 1)), d);
  minv[3] = mul(mul(S(1, 2), S(2, 0)) - mul(S(1, 0), S(2, 2)), d);
  minv[4] = mul(mul(S(0, 0), S(2, 2)) - mul(S(0, 2), S(2, 0)), d);
  minv[5] = mul(mul(S(0, 2), S(1, 0)) - mul(S(0, 0), S(1, 2)), d);
  minv[6] = mul(mul(S(1, 0), S(2, 1)) - mul(S(1, 1), S(2, 0)), d);
  minv[7] = mul(mul(S(0, 1), S(2, 0)) - mul(S(0, 0), S(2, 1)), d);
  minv[8] = mul(mul(S(0, 0), S(1, 1)) - mul(S(0, 1), S(1, 0)), d);
#undef S
  return 0;
}

// perspective_crop (vae-gan.py:176-178): bbox -> the output rectangle [0, W-1] x [0, H-1]
extern "C" int vg_perspective_crop_matrix(const float* bbox, int out_w, int out_h, double* minv) {
  VG_CHECK(bbox != nullptr && minv != nullptr && out_w >= 1 && out_h >= 1, -1, "vg_perspective_crop_matrix: bad arguments");
  const float rect[8] = {0.f, 0.f, static_cast<float>(out_w - 1), 0.f, static_cast<float>(out_w - 1), static_cast<float>(out_h - 1),
                         0.f, static_cast<float>(out_h - 1)};
  return vg_perspective_matrix(bbox, rect, minv);
}
// perspective_unwarp (vae-gan.py:193-196): the patch rectangle [0, w-1] x [0, h-1] -> bbox on the canvas
extern "C" int vg_perspective_unwarp_matrix(const float* bbox, int patch_w, int patch_h, double* minv) {
  VG_CHECK(bbox != nullptr && minv != nullptr && patch_w >= 1 && patch_h >= 1, -1, "vg_perspective_unwarp_matrix: bad arguments");
  const float rect[8] = {0.f, 0.f, static_cast<float>(patch_w - 1), 0.f, static_cast<float>(patch_w - 1),
                         static_cast<float>(patch_h - 1), 0.f, static_cast<float>(patch_h - 1)};
  return vg_perspective_matrix(rect, bbox, minv);
}

static int warp_check(const char* who, const unsigned char* src, int src_h, int src_w, int channels, long long src_row_bytes,
                      const double* minv, int out_h, int out_w) {
  VG_CHECK(src != nullptr && minv != nullptr, -1, "%s: null pointer", who);
  VG_CHECK(src_h >= 1 && src_w >= 1 && src_h <= 32767 && src_w <= 32767 && channels >= 1 && channels <= 4 && out_h >= 1 &&
               out_w >= 1 && src_row_bytes >= static_cast<long long>(src_w) * channels,
           -1, "%s: sizes (source up to 32767 x 32767, 1..4 channels)", who);
  VG_CHECK(static_cast<long long>(out_h) * out_w < (1LL << 31), -1, "%s: output too large", who);
  return 0;
}

static int warp_launch(const unsigned char* src, int src_h, int src_w, int channels, long long src_row_bytes, const double* minv,
                       int out_h, int out_w, unsigned char* dst_u8, float* dst_chw, int transparent, void* stream_) {
  WarpMat m;
  for (int i = 0; i < 9; ++i) m.v[i] = minv[i];
  const int total = out_h * out_w;
  const int grid = std::max(1, std::min((total + 255) / 256, num_sms() * 8));
  warp_perspective_u8_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream_)>>>(src, src_h, src_w, channels, src_row_bytes, m,
                                                                                  out_h, out_w, dst_u8, dst_chw, transparent);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_warp_perspective_u8(const unsigned char* src, int src_h, int src_w, int channels, long long src_row_bytes,
                                      const double* minv, int out_h, int out_w, unsigned char* dst_u8, float* dst_chw,
                                      void* stream_) {
  VG_CHECK(dst_u8 != nullptr || dst_chw != nullptr, -1, "vg_warp_perspective_u8: null pointer");
  if (int rc = warp_check("vg_warp_perspective_u8", src, src_h, src_w, channels, src_row_bytes, minv, out_h, out_w)) return rc;
  return warp_launch(src, src_h, src_w, channels, src_row_bytes, minv, out_h, out_w, dst_u8, dst_chw, 0, stream_);
}

// perspective_unwarp's warp: BORDER_TRANSPARENT into the caller's canvas (out_h x out_w x channels uint8, modified in place)
extern "C" int vg_warp_perspective_u8_transparent(const unsigned char* src, int src_h, int src_w, int channels,
                                                  long long src_row_bytes, const double* minv, int out_h, int out_w,
                                                  unsigned char* canvas_u8, void* stream_) {
  VG_CHECK(canvas_u8 != nullptr, -1, "vg_warp_perspective_u8_transparent: null pointer");
  if (int rc = warp_check("vg_warp_perspective_u8_transparent", src, src_h, src_w, channels, src_row_bytes, minv, out_h, out_w))
    return rc;
  return warp_launch(src, src_h, src_w, channels, src_row_bytes, minv, out_h, out_w, canvas_u8, nullptr, 1, stream_);
}

// A batch of warps in one launch.  `jobs_host` is validated here; `jobs_device` is the same table in device memory (the
// caller copies it on `stream` before this call -- the library never allocates or copies on its own).
extern "C" int vg_warp_perspective_u8_batch(const VgWarpJob* jobs_host, const VgWarpJob* jobs_device, int count, void* stream_) {
  VG_CHECK(jobs_host != nullptr && jobs_device != nullptr && count >= 1 && count <= 65535, -1,
           "vg_warp_perspective_u8_batch: bad table (1..65535 jobs)");
  int max_total = 1;
  for (int i = 0; i < count; ++i) {
    const VgWarpJob& j = jobs_host[i];
    VG_CHECK(j.dst_u8 != nullptr || j.dst_chw != nullptr, -1, "vg_warp_perspective_u8_batch: job %d has no destination", i);
    VG_CHECK(!j.transparent || j.dst_chw == nullptr, -1, "vg_warp_perspective_u8_batch: job %d: transparent jobs write a uint8 canvas", i);
    if (int rc = warp_check("vg_warp_perspective_u8_batch", j.src, j.src_h, j.src_w, j.channels, j.src_row_bytes, j.minv, j.out_h,
                            j.out_w))
      return rc;
    max_total = std::max(max_total, j.out_h * j.out_w);
  }
  const int gx = std::max(1, std::min((max_total + 255) / 256, std::max(1, num_sms() * 8 / count)));
  warp_perspective_u8_batch_kernel<<<dim3(gx, count), 256, 0, static_cast<cudaStream_t>(stream_)>>>(jobs_device);
  VG_LAUNCH_OK();
  return 0;
}

// Host twin of the kernel (same per-pixel code, host buffers): lets the CPU test suite hold the arithmetic to the oracle
// without a GPU, like vg_debug_copy_plan_host.  Not a fallback -- nothing in the package calls it.
extern "C" int vg_debug_warp_perspective_host(const unsigned char* src, int src_h, int src_w, int channels,
                                              long long src_row_bytes, const double* minv, int out_h, int out_w,
                                              unsigned char* dst_u8, float* dst_chw) {
  VG_CHECK(src != nullptr && minv != nullptr && channels >= 1 && channels <= 4, -1, "vg_debug_warp_perspective_host: arguments");
  WarpMat m;
  for (int i = 0; i < 9; ++i) m.v[i] = minv[i];
  for (int y = 0; y < out_h; ++y)
    for (int x = 0; x < out_w; ++x) warp_pixel(src, src_h, src_w, channels, src_row_bytes, m, out_h, out_w, x, y, dst_u8, dst_chw);
  return 0;
}
extern "C" int vg_debug_warp_perspective_transparent_host(const unsigned char* src, int src_h, int src_w, int channels,
                                                          long long src_row_bytes, const double* minv, int out_h, int out_w,
                                                          unsigned char* canvas_u8) {
  VG_CHECK(src != nullptr && minv != nullptr && canvas_u8 != nullptr && channels >= 1 && channels <= 4, -1,
           "vg_debug_warp_perspective_transparent_host: arguments");
  WarpMat m;
  for (int i = 0; i < 9; ++i) m.v[i] = minv[i];
  for (int y = 0; y < out_h; ++y)
    for (int x = 0; x < out_w; ++x) warp_pixel(src, src_h, src_w, channels, src_row_bytes, m, out_h, out_w, x, y, canvas_u8, nullptr, 1);
  return 0;
}
