// HBM-bound data-movement and elementwise kernels around the tensor-core convolutions:
//   * generic 5-D strided copy with dtype conversion / device-scalar scaling / accumulation (weight re-layouts
//     fp32 OIHW -> bf16 GEMM layouts, gradient re-layouts back, NCHW fp32 <-> NHWC bf16 at the module boundary,
//     z/text concat of vae-gan-v2.py:249-251, spectral-norm division by sigma);
//   * FiLM modulation gamma*x+beta and its backward (vae-gan-v2.py:146-149);
//   * bilinear upsampling of the (1 x W/16) text map and its backward (vae-gan-v2.py:138-140);
//   * im2col / col2im for the 3- and 4-channel image-side layers (first conv of the encoder and of D);
//   * direct convolutions with <= 4 output channels (final_image_conv vae-gan-v2.py:232, decode.15
//     vae-gan.py:81, D's patch head vae-gan.py:157) -- HBM-bound GEMV-like work, kept off the tensor pipe.
#include <algorithm>
#include <stdlib.h>
#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

static int ew_grid(long long items, int per_block = 256) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 16;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

// ---------------------------------------------------------------------------------------------
// strided copy
// ---------------------------------------------------------------------------------------------
struct CopyParams {
  long long dims[5], is[5], os[5];
};
template <typename TI, typename TO>
__global__ void strided_copy_kernel(const TI* __restrict__ in, TO* __restrict__ out, CopyParams p,
                                    const float* __restrict__ scale, int scale_inverse, int accumulate) {
  const long long total = p.dims[0] * p.dims[1] * p.dims[2] * p.dims[3] * p.dims[4];
  float sc = 1.f;
  if (scale != nullptr) sc = scale_inverse ? 1.f / *scale : *scale;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    long long r = i, io = 0, oo = 0;
#pragma unroll
    for (int d = 4; d >= 0; --d) {
      const long long k = r % p.dims[d];
      r /= p.dims[d];
      io += k * p.is[d];
      oo += k * p.os[d];
    }
    float v = static_cast<float>(in[io]) * sc;
    if (accumulate) v += static_cast<float>(out[oo]);
    out[oo] = static_cast<TO>(v);
  }
}

// Tiled variant: every block moves one <=4096-element tile through shared memory, reading it in the order of
// the input strides and writing it in the order of the output strides, so both sides of a permutation
// (OIHW -> [Cout][r][s][Cin], OIHW -> [r][s][Cin][Cout], [Cout][r][s][Cin] -> OIHW, NCHW <-> NHWC) move in
// >=128-byte runs.  Slots 0..4 are the (merged) dims sorted by input stride ("ld") and by output stride ("st").
constexpr int kCopyTileMax = 4096;
constexpr int kCopySmem = 6144;
struct CopySide {
  int tile[5], dim[5], ntile[5], div[5], ss[5];
  unsigned magic[5];          // ceil(2^32 / tile) for tile >= 2 (exact quotient for numerators < 2^16)
  long long gs[5];
  int gsi[5];                 // gs as int: offsets INSIDE a tile stay below 2^31 (checked by the planner)
  int nt;                     // slots [0, nt) have tile > 1 (the only ones the per-element decode walks)
};
struct TiledCopyParams {
  CopySide ld, st;
  int tile_elems;
};
__host__ __device__ __forceinline__ unsigned copy_mulhi(unsigned x, unsigned y) {
#ifdef __CUDA_ARCH__
  return __umulhi(x, y);
#else
  return static_cast<unsigned>((static_cast<unsigned long long>(x) * y) >> 32);
#endif
}
__host__ __device__ __forceinline__ bool copy_decode(const CopySide& s, const int (&ext)[5], int l, int& goff, int& soff) {
  unsigned r = static_cast<unsigned>(l);
  bool ok = true;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    if (k < s.nt) {
      const unsigned t = static_cast<unsigned>(s.tile[k]);
      const unsigned q = copy_mulhi(r, s.magic[k]);
      const int c = static_cast<int>(r - q * t);
      r = q;
      ok = ok && c < ext[k];
      goff += c * s.gsi[k];
      soff += c * s.ss[k];
    }
  }
  return ok;
}
__host__ __device__ __forceinline__ long long copy_origin(const CopySide& s, unsigned bid, int (&ext)[5]) {
  long long g = 0;
#pragma unroll
  for (int k = 0; k < 5; ++k) {
    const int t = static_cast<int>((bid / static_cast<unsigned>(s.div[k])) % static_cast<unsigned>(s.ntile[k]));
    const int o = t * s.tile[k];
    ext[k] = s.tile[k] < s.dim[k] - o ? s.tile[k] : s.dim[k] - o;
    g += static_cast<long long>(o) * s.gs[k];
  }
  return g;
}
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) tiled_copy_kernel(const TI* __restrict__ in, TO* __restrict__ out,
                                                         const __grid_constant__ TiledCopyParams p,
                                                         const float* __restrict__ scale, int scale_inverse,
                                                         int accumulate) {
  __shared__ float sm[kCopySmem];
  float sc = 1.f;
  if (scale != nullptr) sc = scale_inverse ? 1.f / *scale : *scale;
  int ext[5];
  const long long ibase = copy_origin(p.ld, blockIdx.x, ext);
  // all (up to 16) loads of a thread are issued before the first use: with one 4-byte load in flight per thread the
  // kernel is bound by latency x bytes-in-flight (~0.6 TB/s), not by HBM
  constexpr int kPerThread = kCopyTileMax / 256;
  float v[kPerThread];
  int so[kPerThread];
#pragma unroll
  for (int u = 0; u < kPerThread; ++u) {
    const int l = threadIdx.x + u * 256;
    int goff = 0, soff = 0;
    const bool ok = l < p.tile_elems && copy_decode(p.ld, ext, l, goff, soff);
    so[u] = ok ? soff : -1;
    v[u] = ok ? static_cast<float>(in[ibase + goff]) : 0.f;
  }
#pragma unroll
  for (int u = 0; u < kPerThread; ++u)
    if (so[u] >= 0) sm[so[u]] = v[u];
  __syncthreads();
  const long long obase = copy_origin(p.st, blockIdx.x, ext);
#pragma unroll 4
  for (int l = threadIdx.x; l < p.tile_elems; l += 256) {
    int goff = 0, soff = 0;
    if (copy_decode(p.st, ext, l, goff, soff)) {
      float v = sm[soff] * sc;
      if (accumulate) v += static_cast<float>(out[obase + goff]);
      out[obase + goff] = static_cast<TO>(v);
    }
  }
}

// Same-order copy (both sides dense): 8 elements per thread per trip, 16/32-byte accesses
template <typename TI, typename TO>
__global__ void contig_copy_kernel(const TI* __restrict__ in, TO* __restrict__ out, long long n8,
                                   const float* __restrict__ scale, int scale_inverse, int accumulate) {
  float sc = 1.f;
  if (scale != nullptr) sc = scale_inverse ? 1.f / *scale : *scale;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n8;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float f[8];
    load8(in + i * 8, f);
    if (accumulate) {
      float o[8];
      load8(out + i * 8, o);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] = fmaf(f[k], sc, o[k]);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] *= sc;
    }
    store8(out + i * 8, f);
  }
}
static bool same_order_dense(const long long* dims, const long long* is, const long long* os, long long* total) {
  long long run = 1;
  for (int i = 4; i >= 0; --i) {
    if (dims[i] == 1) continue;
    if (is[i] != run || os[i] != run) return false;
    run *= dims[i];
  }
  *total = run;
  return true;
}

// Plans the tiling; returns the number of tiles, or 0 when the generic kernel should be used.
static long long plan_tiled_copy(const long long* dims, const long long* is, const long long* os, TiledCopyParams& p) {
  long long d[5], a[5], b[5];
  int n = 0;
  for (int i = 0; i < 5; ++i)
    if (dims[i] > 1) { d[n] = dims[i]; a[n] = is[i]; b[n] = os[i]; ++n; }
  for (int i = 0; i + 1 < n;) {   // merge dims that are contiguous on both sides
    if (a[i] == a[i + 1] * d[i + 1] && b[i] == b[i + 1] * d[i + 1]) {
      d[i] *= d[i + 1]; a[i] = a[i + 1]; b[i] = b[i + 1];
      for (int j = i + 1; j + 1 < n; ++j) { d[j] = d[j + 1]; a[j] = a[j + 1]; b[j] = b[j + 1]; }
      --n;
    } else {
      ++i;
    }
  }
  if (n == 0) { n = 1; d[0] = 1; a[0] = 1; b[0] = 1; }
  for (int i = 0; i < n; ++i)
    if (d[i] >= (1LL << 31) || a[i] < 0 || b[i] <= 0) return 0;
  for (int i = n; i < 5; ++i) { d[i] = 1; a[i] = 0; b[i] = 0; }
  int lo[5] = {0, 1, 2, 3, 4}, so[5] = {0, 1, 2, 3, 4};
  auto key = [](long long stride, long long dim) { return (dim == 1 || stride == 0) ? (1LL << 62) : stride; };
  std::stable_sort(lo, lo + 5, [&](int x, int y) { return key(a[x], d[x]) < key(a[y], d[y]); });
  std::stable_sort(so, so + 5, [&](int x, int y) { return key(b[x], d[x]) < key(b[y], d[y]); });
  long long t[5] = {1, 1, 1, 1, 1};
  auto cdiv = [](long long x, long long y) { return (x + y - 1) / y; };
  long long prod = 1;
  for (int second = 1; second >= 0; --second) {   // second = also extend each side into its next contiguous dim
    for (int i = 0; i < 5; ++i) t[i] = 1;
    const int s0 = so[0];
    t[s0] = std::min<long long>(d[s0], 64);
    if (second && t[s0] < 64 && n > 1) {
      const int s1 = so[1];
      if (b[s1] == b[s0] * d[s0]) t[s1] = std::min<long long>(d[s1], cdiv(64, t[s0]));
    }
    const int l0 = lo[0];
    t[l0] = std::max<long long>(t[l0], std::min<long long>(d[l0], 32));
    if (second && t[l0] == d[l0] && t[l0] < 32 && n > 1) {
      const int l1 = lo[1];
      if (a[l1] == a[l0] * d[l0]) t[l1] = std::max<long long>(t[l1], std::min<long long>(d[l1], cdiv(32, t[l0])));
    }
    prod = 1;
    for (int i = 0; i < 5; ++i) prod *= t[i];
    if (prod <= kCopyTileMax) break;
  }
  for (int k = 0; k < 5 && prod < 2048; ++k) {   // small tiles: grow along the output-fast dims
    const int i = so[k];
    const long long nt = std::min<long long>(d[i], t[i] * (2048 / prod));
    prod = prod / t[i] * nt;
    t[i] = nt;
  }
  if (prod > kCopyTileMax) return 0;
  // shared-memory strides follow the output order; odd strides keep the transposing phase off a single bank
  int ss[5];
  long long run = 1;
  for (int k = 0; k < 5; ++k) {
    const int i = so[k];
    ss[i] = static_cast<int>(run);
    run *= t[i];
    if (k < 4 && run > 1 && run % 2 == 0 && t[so[k + 1]] > 1) run += 1;
  }
  if (run > kCopySmem) return 0;
  long long ntile[5], div[5], tiles = 1;
  for (int i = 0; i < 5; ++i) { ntile[i] = cdiv(d[i], t[i]); div[i] = tiles; tiles *= ntile[i]; }
  if (tiles >= (1LL << 31)) return 0;
  bool fits = true;
  auto fill = [&](CopySide& s, const int* order_in, const long long* gstride) {
    int order[5], m = 0;
    for (int k = 0; k < 5; ++k)
      if (t[order_in[k]] > 1) order[m++] = order_in[k];      // tiled dims first (stride order kept) ...
    s.nt = m;
    for (int k = 0; k < 5; ++k)
      if (t[order_in[k]] <= 1) order[m++] = order_in[k];     // ... then the ones the per-element decode skips
    long long span = 0;
    for (int k = 0; k < 5; ++k) {
      const int i = order[k];
      span += (t[i] - 1) * gstride[i];
      s.gsi[k] = static_cast<int>(gstride[i]);
      s.tile[k] = static_cast<int>(t[i]); s.dim[k] = static_cast<int>(d[i]); s.ntile[k] = static_cast<int>(ntile[i]);
      s.div[k] = static_cast<int>(div[i]); s.ss[k] = ss[i]; s.gs[k] = gstride[i];
      s.magic[k] = t[i] >= 2 ? static_cast<unsigned>(((1ULL << 32) + t[i] - 1) / t[i]) : 0u;
    }
    if (span >= (1LL << 31)) fits = false;
  };
  fill(p.ld, lo, a);
  fill(p.st, so, b);
  if (!fits) return 0;
  p.tile_elems = static_cast<int>(prod);
  return tiles;
}

// ---------------------------------------------------------------------------------------------
// FiLM
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void film_fwd_kernel(const T* __restrict__ gb, const T* __restrict__ x, int x_ld,
                                int x_coff, T* __restrict__ y, long long rows, int c) {
  const int cv = c / 8;
  const long long total = rows * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cv;
    const int ch = static_cast<int>(i % cv) * 8;
    float g[8], b[8], xv[8], o[8];
    load8(gb + r * 2 * c + ch, g);
    load8(gb + r * 2 * c + c + ch, b);
    load8(x + r * x_ld + x_coff + ch, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(g[k], xv[k], b[k]);
    store8(y + r * c + ch, o);
  }
}
template <typename T>
__global__ void film_bwd_kernel(const T* __restrict__ gb, const T* __restrict__ x, int x_ld,
                                int x_coff, const T* __restrict__ dy, T* __restrict__ dgb,
                                T* __restrict__ dx, int dx_ld, int dx_coff, long long rows, int c) {
  const int cv = c / 8;
  const long long total = rows * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cv;
    const int ch = static_cast<int>(i % cv) * 8;
    float g[8], xv[8], d[8], o1[8], o2[8];
    load8(gb + r * 2 * c + ch, g);
    load8(x + r * x_ld + x_coff + ch, xv);
    load8(dy + r * c + ch, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) { o1[k] = d[k] * xv[k]; o2[k] = d[k] * g[k]; }
    store8(dgb + r * 2 * c + ch, o1);      // d gamma
    store8(dgb + r * 2 * c + c + ch, d);   // d beta
    store8(dx + r * dx_ld + dx_coff + ch, o2);
  }
}

// ---------------------------------------------------------------------------------------------
// FiLM with row-class parameter maps: gb has only gh (= 3) rows per image -- first row, any interior row, last row --
// because the maps the reference builds (vae-gan-v2.py:138-145) are row-constant except at the zero-padded border.
// ---------------------------------------------------------------------------------------------
VG_DEVICE int film_row_class(int yy, int h, int gh) { return yy == 0 ? 0 : (yy == h - 1 ? gh - 1 : 1); }

template <typename T>
__global__ void film_rows_fwd_kernel(const T* __restrict__ gb, int gh, const T* __restrict__ x, int x_ld, int x_coff,
                                     T* __restrict__ y, int n, int h, int w, int c) {
  const int cv = c / 8;
  const long long total = static_cast<long long>(n) * h * w * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cv;                       // pixel (b, yy, xx)
    const int ch = static_cast<int>(i - r * cv) * 8;
    const int xx = static_cast<int>(r % w);
    const long long by = r / w;
    const int yy = static_cast<int>(by % h);
    const long long b = by / h;
    const T* g = gb + ((b * gh + film_row_class(yy, h, gh)) * w + xx) * 2 * c + ch;
    float ga[8], be[8], xv[8], o[8];
    load8(g, ga);
    load8(g + c, be);
    load8(x + r * x_ld + x_coff + ch, xv);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = fmaf(ga[k], xv[k], be[k]);
    store8(y + r * c + ch, o);
  }
}
// one thread = 8 channels of one image column: walks the h rows, writes dx and the per-class SUMS of d gamma / d beta
template <typename T>
__global__ void film_rows_bwd_kernel(const T* __restrict__ gb, int gh, const T* __restrict__ x, int x_ld, int x_coff,
                                     const T* __restrict__ dy, T* __restrict__ dgb, T* __restrict__ dx, int dx_ld,
                                     int dx_coff, int n, int h, int w, int c) {
  const int cv = c / 8;
  const long long total = static_cast<long long>(n) * w * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long bx = i / cv;
    const int ch = static_cast<int>(i - bx * cv) * 8;
    const int xx = static_cast<int>(bx % w);
    const long long b = bx / w;
    float gam[3][8];
#pragma unroll
    for (int k = 0; k < 3; ++k) load8(gb + ((b * gh + (k == 2 ? gh - 1 : k)) * w + xx) * 2 * c + ch, gam[k]);
    float ag[8], ab[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) ag[k] = ab[k] = 0.f;
#pragma unroll 4
    for (int yy = 0; yy < h; ++yy) {
      const long long r = (b * h + yy) * w + xx;
      float xv[8], d[8], o[8];
      load8(x + r * x_ld + x_coff + ch, xv);
      load8(dy + r * c + ch, d);
      const int k = yy == 0 ? 0 : (yy == h - 1 ? 2 : 1);
#pragma unroll
      for (int e = 0; e < 8; ++e) o[e] = d[e] * gam[k][e];
      store8(dx + r * dx_ld + dx_coff + ch, o);
      if (k == 1) {
#pragma unroll
        for (int e = 0; e < 8; ++e) { ag[e] = fmaf(d[e], xv[e], ag[e]); ab[e] += d[e]; }
      } else {
        T* o2 = dgb + ((b * gh + (k == 2 ? gh - 1 : 0)) * w + xx) * 2 * c + ch;
#pragma unroll
        for (int e = 0; e < 8; ++e) o[e] = d[e] * xv[e];
        store8(o2, o);
        store8(o2 + c, d);
      }
    }
    T* o1 = dgb + ((b * gh + 1) * w + xx) * 2 * c + ch;
    store8(o1, ag);
    store8(o1 + c, ab);
  }
}

// ---------------------------------------------------------------------------------------------
// bilinear upsample of a (1 x w0) map to (h x w), align_corners = False; rows of the result are identical
// ---------------------------------------------------------------------------------------------
VG_DEVICE void lerp_src(int j, int w0, int w, int& j0, int& j1, float& lam) {
  float src = (static_cast<float>(j) + 0.5f) * (static_cast<float>(w0) / static_cast<float>(w)) - 0.5f;
  if (src < 0.f) src = 0.f;
  j0 = static_cast<int>(src);
  if (j0 > w0 - 1) j0 = w0 - 1;
  j1 = j0 + (j0 < w0 - 1 ? 1 : 0);
  lam = src - static_cast<float>(j0);
}
// every output row is the same interpolated line: one thread interpolates 8 channels of one column once and
// stores them to rows_per_thread rows (write-only traffic, 16-byte stores, 8 lanes = 128 contiguous bytes)
template <typename T>
__global__ void upsample_fwd_kernel(const T* __restrict__ t, int t_ld, int t_coff, int n, int w0, int c,
                                    T* __restrict__ y, int h, int w, int rows_per_thread) {
  const unsigned cv = static_cast<unsigned>(c / 8);
  const unsigned line = static_cast<unsigned>(w) * cv;                 // vectors per output row
  const unsigned ichunks = static_cast<unsigned>((h + rows_per_thread - 1) / rows_per_thread);
  const long long total = static_cast<long long>(n) * ichunks * line;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned bi = static_cast<unsigned>(i / line);              // b * ichunks + chunk
    const unsigned v = static_cast<unsigned>(i - static_cast<long long>(bi) * line);
    const int j = static_cast<int>(v / cv);
    const int ch = static_cast<int>(v - static_cast<unsigned>(j) * cv) * 8;
    const int b = static_cast<int>(bi / ichunks);
    const int r0 = static_cast<int>(bi - static_cast<unsigned>(b) * ichunks) * rows_per_thread;
    int j0, j1;
    float lam;
    lerp_src(j, w0, w, j0, j1, lam);
    float a[8], bb[8], o[8];
    load8(t + (static_cast<long long>(b) * w0 + j0) * t_ld + t_coff + ch, a);
    load8(t + (static_cast<long long>(b) * w0 + j1) * t_ld + t_coff + ch, bb);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (1.f - lam) * a[k] + lam * bb[k];
    T* dst = y + ((static_cast<long long>(b) * h + r0) * w + j) * c + ch;
    const int r1 = min(h, r0 + rows_per_thread);
    for (int r = r0; r < r1; ++r, dst += static_cast<long long>(w) * c) store8(dst, o);
  }
}
// dt[b][js][c] (fp32, accumulated) = sum_i sum_j weight(j -> js) dy[b][i][j][c].
// One thread = 8 channels x a run of `jlen` destination columns x `rows_per_thread` rows: every dy element is read
// exactly once (each column feeds the two sources it was interpolated from); a run is short enough (<= half the
// upsampling ratio) to touch at most three consecutive sources, whose partial sums live in registers.
template <typename T>
__global__ void __launch_bounds__(256, 3) upsample_bwd_kernel(const T* __restrict__ dy, int n, int h, int w, int c, int w0,
                                    float* __restrict__ dt, int rows_per_thread, int jlen) {
  const int cv = c / 8;
  const int ichunks = (h + rows_per_thread - 1) / rows_per_thread;
  const int jruns = (w + jlen - 1) / jlen;
  const long long total = static_cast<long long>(n) * jruns * cv * ichunks;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % cv) * 8;
    long long r = idx / cv;
    const int jr = static_cast<int>(r % jruns);
    r /= jruns;
    const int ic = static_cast<int>(r % ichunks);
    const int b = static_cast<int>(r / ichunks);
    const int j_begin = jr * jlen, j_end = min(w, j_begin + jlen);
    const int i_begin = ic * rows_per_thread, i_end = min(h, i_begin + rows_per_thread);
    int base, tmp;
    float lam0;
    lerp_src(j_begin, w0, w, base, tmp, lam0);
    float acc[3][8];
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3)
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[s3][k] = 0.f;
    for (int j = j_begin; j < j_end; ++j) {
      int j0, j1;
      float lam;
      lerp_src(j, w0, w, j0, j1, lam);
      float colsum[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) colsum[k] = 0.f;
      // all rows of the run are loaded before any is consumed: 8 independent 16-byte loads in flight per thread
      // (one dependent load per row left the kernel latency-bound at 0.46 of the copy bandwidth)
      const T* col = dy + ((static_cast<long long>(b) * h + i_begin) * w + j) * c + ch;
      const long long rstride = static_cast<long long>(w) * c;
      for (int i0 = i_begin; i0 < i_end; i0 += 8) {
        Raw8<T> v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u)
          if (i0 + u < i_end) v[u] = Raw8<T>::load(col + (i0 - i_begin + u) * rstride);
#pragma unroll
        for (int u = 0; u < 8; ++u) {
          if (i0 + u < i_end) {
            float d[8];
            v[u].unpack(d);
#pragma unroll
            for (int k = 0; k < 8; ++k) colsum[k] += d[k];
          }
        }
      }
      const int s0 = j0 - base, s1 = j1 - base;          // 0..2 by construction of jlen
#pragma unroll
      for (int s3 = 0; s3 < 3; ++s3) {
        const float wgt = (s3 == s0 ? 1.f - lam : 0.f) + (s3 == s1 ? lam : 0.f);
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[s3][k] = fmaf(wgt, colsum[k], acc[s3][k]);
      }
    }
#pragma unroll
    for (int s3 = 0; s3 < 3; ++s3) {
      if (base + s3 >= w0) break;
      float* o = dt + (static_cast<long long>(b) * w0 + base + s3) * c + ch;
      bool any = false;
#pragma unroll
      for (int k = 0; k < 8; ++k) any = any || acc[s3][k] != 0.f;
      if (any) {
#pragma unroll
        for (int k = 0; k < 8; ++k) atomicAdd(o + k, acc[s3][k]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// Bilinear resize along H only (align_corners=False): y[b][i][x][:] = (1-mu) t[b][i0][x][:] + mu t[b][i1][x][:].
// With the W-only kernels above this composes F.interpolate(..., mode='bilinear') of a multi-row map (the 4-row text
// map of vae-gan-oldv.py:165-176, 286-291) -- bilinear interpolation is separable.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void upsample_h_fwd_kernel(const T* __restrict__ t, int n, int h0, int w, int c, T* __restrict__ y, int h) {
  const int cv = c / 8;
  const long long total = static_cast<long long>(n) * h * w * cv;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % cv) * 8;
    long long r = idx / cv;
    const int x = static_cast<int>(r % w);
    r /= w;
    const int i = static_cast<int>(r % h);
    const long long b = r / h;
    int i0, i1;
    float mu;
    lerp_src(i, h0, h, i0, i1, mu);
    float a[8], bb[8], o[8];
    load8(t + ((b * h0 + i0) * w + x) * c + ch, a);
    load8(t + ((b * h0 + i1) * w + x) * c + ch, bb);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = (1.f - mu) * a[k] + mu * bb[k];
    store8(y + ((b * h + i) * w + x) * c + ch, o);
  }
}
// dt[b][r][x][:] = sum_i weight(i -> r) dy[b][i][x][:]   (one thread per source element; fp32 output)
template <typename T>
__global__ void upsample_h_bwd_kernel(const T* __restrict__ dy, int n, int h, int w, int c, int h0, float* __restrict__ dt) {
  const int cv = c / 8;
  const long long total = static_cast<long long>(n) * h0 * w * cv;
  const float ratio = static_cast<float>(h) / static_cast<float>(h0);
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(idx % cv) * 8;
    long long r = idx / cv;
    const int x = static_cast<int>(r % w);
    r /= w;
    const int rs = static_cast<int>(r % h0);
    const long long b = r / h0;
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    int ilo = static_cast<int>((static_cast<float>(rs) - 1.f) * ratio) - 2;
    int ihi = static_cast<int>((static_cast<float>(rs) + 2.f) * ratio) + 2;
    if (ilo < 0) ilo = 0;
    if (ihi > h - 1) ihi = h - 1;
    for (int i = ilo; i <= ihi; ++i) {
      int i0, i1;
      float mu;
      lerp_src(i, h0, h, i0, i1, mu);
      float wgt = 0.f;
      if (i0 == rs) wgt += 1.f - mu;
      if (i1 == rs) wgt += mu;
      if (wgt == 0.f) continue;
      float d[8];
      load8(dy + ((b * h + i) * w + x) * c + ch, d);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] = fmaf(wgt, d[k], acc[k]);
    }
    float* o = dt + ((b * h0 + rs) * w + x) * c + ch;
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = acc[k];
  }
}

// ---------------------------------------------------------------------------------------------
// im2col / col2im for few-channel NHWC bf16 images: col[m][(r*kw+q)*c + ch], zero padded to kpad columns
// ---------------------------------------------------------------------------------------------
constexpr int kIm2colMaxK = 1024;
template <typename T>
__global__ void im2col_kernel(const T* __restrict__ src, int n, int h, int w, int ld, int c, int kh,
                              int kw, int stride, int pad, int oh, int ow, T* __restrict__ col, int kpad) {
  // one thread = 8 consecutive columns of one output pixel row (one 16-byte store); the column -> (r, q, ch)
  // decode is tabulated once per block so the inner loop carries no integer division
  __shared__ int tab[kIm2colMaxK];
  const int kvalid = kh * kw * c;
  for (int k = threadIdx.x; k < kpad; k += blockDim.x) {
    int e = -1;
    if (k < kvalid) {
      const int ch = k % c, tap = k / c;
      e = ((tap / kw) << 16) | ((tap % kw) << 8) | ch;
    }
    tab[k] = e;
  }
  __syncthreads();
  const unsigned kv = static_cast<unsigned>(kpad / 8);
  const long long total = static_cast<long long>(n) * oh * ow * kv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const unsigned m = static_cast<unsigned>(i / kv);
    const int k0 = static_cast<int>(i - static_cast<long long>(m) * kv) * 8;
    const unsigned row = m / static_cast<unsigned>(ow);
    const int ox = static_cast<int>(m - row * ow);
    const int b = static_cast<int>(row / static_cast<unsigned>(oh));
    const int oy = static_cast<int>(row - static_cast<unsigned>(b) * oh);
    const T* img = src + static_cast<long long>(b) * h * w * ld;
    const int y0 = oy * stride - pad, x0 = ox * stride - pad;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int e = tab[k0 + j];
      float v = 0.f;
      if (e >= 0) {
        const int iy = y0 + (e >> 16), ix = x0 + ((e >> 8) & 255);
        if (iy >= 0 && iy < h && ix >= 0 && ix < w) v = static_cast<float>(img[(iy * w + ix) * ld + (e & 255)]);
      }
      f[j] = v;
    }
    store8(col + static_cast<long long>(m) * kpad + k0, f);
  }
}
// dsrc[b][c][iy][ix] (fp32 NCHW, overwritten) = sum over taps of dcol
template <typename T>
__global__ void col2im_kernel(const T* __restrict__ dcol, int kpad, int n, int h, int w, int c, int kh,
                              int kw, int stride, int pad, int oh, int ow, float* __restrict__ dsrc) {
  const long long total = static_cast<long long>(n) * c * h * w;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ix = static_cast<int>(i % w), iy = static_cast<int>((i / w) % h);
    const int ch = static_cast<int>((i / (static_cast<long long>(w) * h)) % c);
    const int b = static_cast<int>(i / (static_cast<long long>(w) * h * c));
    float acc = 0.f;
    for (int r = 0; r < kh; ++r) {
      const int ty = iy + pad - r;
      if (ty < 0 || ty % stride != 0) continue;
      const int oy = ty / stride;
      if (oy >= oh) continue;
      for (int q = 0; q < kw; ++q) {
        const int tx = ix + pad - q;
        if (tx < 0 || tx % stride != 0) continue;
        const int ox = tx / stride;
        if (ox >= ow) continue;
        acc += static_cast<float>(dcol[((static_cast<long long>(b) * oh + oy) * ow + ox) * kpad + (r * kw + q) * c + ch]);
      }
    }
    dsrc[i] = acc;
  }
}

// ---------------------------------------------------------------------------------------------
// direct stride-1 convolutions with few (<= 4) output channels; weights fp32 [cout][kh][kw][cin]
// ---------------------------------------------------------------------------------------------
constexpr int kMaxSmallN = 4;
// forward: a group of G lanes per output pixel, lanes stride over (tap, channel-vector)
template <typename T>
__global__ void smalln_fwd_kernel(const T* __restrict__ x, int x_ld, int x_coff, int n, int h, int w,
                                  int cin, const float* __restrict__ wt, const float* __restrict__ bias, int cout,
                                  int kh, int kw, int pad, int oh, int ow, float* __restrict__ out, int lanes_per_px) {
  const int G = lanes_per_px;
  const long long pixels = static_cast<long long>(n) * oh * ow;
  const int cv = cin / 8;
  const int kvecs = kh * kw * cv;
  const int lane_in_g = threadIdx.x % G;
  if (kh == 1 && kw == 1 && pad == 0 && kvecs == G) {
    // pointwise conv with one channel vector per lane (final_image_conv, vae-gan-v2.py:232): the lane's weights stay
    // in registers, the pixel index needs no decoding, each warp reads 32/G whole pixel rows (contiguous bytes)
    float wreg[kMaxSmallN][8];
    for (int o = 0; o < kMaxSmallN; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) wreg[o][k] = o < cout ? __ldg(wt + static_cast<long long>(o) * cin + lane_in_g * 8 + k) : 0.f;
    float bv[kMaxSmallN];
    for (int o = 0; o < kMaxSmallN; ++o) bv[o] = (bias != nullptr && o < cout) ? bias[o] : 0.f;
    // four pixels per trip: all four 16-byte loads are in flight before the first use
    const long long gstride = static_cast<long long>(gridDim.x) * blockDim.x / G;
    const long long px0 = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
    const long long trips = (pixels + 4 * gstride - 1) / (4 * gstride);       // identical for every thread (shuffles)
    for (long long trip = 0; trip < trips; ++trip) {
      Raw8<T> raw[4];
      long long pxs[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        pxs[u] = px0 + (trip * 4 + u) * gstride;
        if (pxs[u] < pixels) raw[u] = Raw8<T>::load(x + pxs[u] * x_ld + x_coff + lane_in_g * 8);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        float acc[kMaxSmallN] = {0.f, 0.f, 0.f, 0.f};
        if (pxs[u] < pixels) {
          float f[8];
          raw[u].unpack(f);
#pragma unroll
          for (int o = 0; o < kMaxSmallN; ++o)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[o] = fmaf(f[k], wreg[o][k], acc[o]);
        }
        if (G == 8) {
          // transposing reduction of the 4 partial outputs over the 8 lanes of the pixel: 4 shuffles instead of 12;
          // lane l ends up with the full sum of output (l >> 1) & 3
          const bool hi4 = lane_in_g & 4, hi2 = lane_in_g & 2;
          const float r0 = __shfl_xor_sync(0xffffffffu, hi4 ? acc[0] : acc[2], 4);
          const float r1 = __shfl_xor_sync(0xffffffffu, hi4 ? acc[1] : acc[3], 4);
          const float a0 = (hi4 ? acc[2] : acc[0]) + r0, a1 = (hi4 ? acc[3] : acc[1]) + r1;
          float b = (hi2 ? a1 : a0) + __shfl_xor_sync(0xffffffffu, hi2 ? a0 : a1, 2);
          b += __shfl_xor_sync(0xffffffffu, b, 1);
          const int o = (lane_in_g >> 1) & 3;
          if ((lane_in_g & 1) == 0 && pxs[u] < pixels && o < cout) out[pxs[u] * cout + o] = b + bv[o];
        } else {
#pragma unroll
          for (int o = 0; o < kMaxSmallN; ++o) {
            float v = acc[o];
            for (int sft = G / 2; sft > 0; sft >>= 1) v += __shfl_xor_sync(0xffffffffu, v, sft);
            if (lane_in_g == 0 && pxs[u] < pixels && o < cout) out[pxs[u] * cout + o] = v + bv[o];
          }
        }
      }
    }
    return;
  }
  for (long long px = (static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x) / G;
       px < ((pixels + (32 / G) - 1) / (32 / G)) * (32 / G);      // keep whole warps in the loop for the shuffles
       px += static_cast<long long>(gridDim.x) * blockDim.x / G) {
    float acc[kMaxSmallN] = {0.f, 0.f, 0.f, 0.f};
    if (px < pixels) {
      const int ox = static_cast<int>(px % ow), oy = static_cast<int>((px / ow) % oh);
      const int b = static_cast<int>(px / (static_cast<long long>(ow) * oh));
      for (int kv = lane_in_g; kv < kvecs; kv += G) {
        const int tap = kv / cv, ch = (kv % cv) * 8;
        const int iy = oy + tap / kw - pad, ix = ox + tap % kw - pad;
        if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
        float f[8];
        load8(x + ((static_cast<long long>(b) * h + iy) * w + ix) * x_ld + x_coff + ch, f);
        for (int o = 0; o < cout; ++o) {
          const float* wp = wt + (static_cast<long long>(o) * kh * kw + tap) * cin + ch;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[o] = fmaf(f[k], __ldg(wp + k), acc[o]);
        }
      }
    }
    for (int o = 0; o < cout; ++o) {
      float v = acc[o];
      for (int s = G / 2; s > 0; s >>= 1) v += __shfl_xor_sync(0xffffffffu, v, s);
      if (lane_in_g == 0 && px < pixels) out[px * cout + o] = v + (bias ? bias[o] : 0.f);
    }
  }
}
// data gradient: thread per (input pixel, channel vector)
template <typename T>
__global__ void smalln_dgrad_kernel(const float* __restrict__ dy, int n, int oh, int ow, int cout,
                                    const float* __restrict__ wt, int kh, int kw, int pad, int h, int w, int cin,
                                    T* __restrict__ dx, int dx_ld, int dx_coff) {
  const int cv = cin / 8;
  const long long total = static_cast<long long>(n) * h * w * cv;
  const long long nthreads = static_cast<long long>(gridDim.x) * blockDim.x;
  if (kh == 1 && kw == 1 && pad == 0 && nthreads % cv == 0) {
    // pointwise: the thread's channel vector never changes -> weights in registers, no index decoding
    const long long i0 = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    const int ch = static_cast<int>(i0 % cv) * 8;
    float wreg[kMaxSmallN][8];
    for (int o = 0; o < kMaxSmallN; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) wreg[o][k] = o < cout ? __ldg(wt + static_cast<long long>(o) * cin + ch + k) : 0.f;
    for (long long i = i0; i < total; i += 4 * nthreads) {
      float gv[4][kMaxSmallN];
      long long pix[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long iu = i + u * nthreads;
        pix[u] = iu < total ? iu / cv : -1;
#pragma unroll
        for (int o = 0; o < kMaxSmallN; ++o) gv[u][o] = (pix[u] >= 0 && o < cout) ? __ldg(dy + pix[u] * cout + o) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (pix[u] < 0) continue;
        float acc[8];
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll
        for (int o = 0; o < kMaxSmallN; ++o)
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(gv[u][o], wreg[o][k], acc[k]);
        store8(dx + pix[u] * dx_ld + dx_coff + ch, acc);
      }
    }
    return;
  }
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int ch = static_cast<int>(i % cv) * 8;
    const long long pix = i / cv;
    const int ix = static_cast<int>(pix % w), iy = static_cast<int>((pix / w) % h);
    const int b = static_cast<int>(pix / (static_cast<long long>(w) * h));
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] = 0.f;
    for (int r = 0; r < kh; ++r) {
      const int oy = iy + pad - r;
      if (oy < 0 || oy >= oh) continue;
      for (int q = 0; q < kw; ++q) {
        const int ox = ix + pad - q;
        if (ox < 0 || ox >= ow) continue;
        const float* g = dy + ((static_cast<long long>(b) * oh + oy) * ow + ox) * cout;
        for (int o = 0; o < cout; ++o) {
          const float gv = g[o];
          const float* wp = wt + (static_cast<long long>(o) * kh * kw + r * kw + q) * cin + ch;
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[k] = fmaf(gv, __ldg(wp + k), acc[k]);
        }
      }
    }
    store8(dx + pix * dx_ld + dx_coff + ch, acc);
  }
}
// pointwise (1x1, pad 0) weight gradient: thread = (pixel lane, channel vector); four pixels per trip with all loads in
// flight first; registers -> shared memory across the pixel lanes of the block -> one fp32 atomic per (block, element)
template <typename T>
__global__ void __launch_bounds__(256) pw_wgrad_kernel(const float* __restrict__ dy, const T* __restrict__ x, int x_ld,
                                                       int x_coff, long long pixels, int cin, int cout,
                                                       float* __restrict__ dw, float* __restrict__ dbias) {
  const int cv = cin / 8;                 // <= 32 (checked by the host)
  const int ppar = 256 / cv;
  const int tid = threadIdx.x;
  const int pl = tid / cv, cvi = tid % cv;
  const int ch = cvi * 8;
  float acc[kMaxSmallN][8], bacc[kMaxSmallN];
#pragma unroll
  for (int o = 0; o < kMaxSmallN; ++o) {
    bacc[o] = 0.f;
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[o][k] = 0.f;
  }
  const long long stride = static_cast<long long>(gridDim.x) * ppar;
  if (pl < ppar) {
    for (long long px = static_cast<long long>(blockIdx.x) * ppar + pl; px < pixels; px += 4 * stride) {
      Raw8<T> raw[4];
      float g[4][kMaxSmallN];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const long long pu = px + u * stride;
        const bool ok = pu < pixels;
        if (ok) raw[u] = Raw8<T>::load(x + pu * x_ld + x_coff + ch);
#pragma unroll
        for (int o = 0; o < kMaxSmallN; ++o) g[u][o] = (ok && o < cout) ? __ldg(dy + pu * cout + o) : 0.f;
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        if (px + u * stride >= pixels) continue;
        float f[8];
        raw[u].unpack(f);
#pragma unroll
        for (int o = 0; o < kMaxSmallN; ++o) {
          bacc[o] += g[u][o];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[o][k] = fmaf(g[u][o], f[k], acc[o][k]);
        }
      }
    }
  }
  __shared__ float red[256][8 * kMaxSmallN + 1];
#pragma unroll
  for (int o = 0; o < kMaxSmallN; ++o)
#pragma unroll
    for (int k = 0; k < 8; ++k) red[tid][o * 8 + k] = acc[o][k];
  __syncthreads();
  // thread t reduces element (t % 32) of channel vector (t / 32) over the pixel lanes
  for (int item = tid; item < cv * 32; item += 256) {
    const int v = item >> 5, e = item & 31, o = e >> 3, k = e & 7;
    if (o < cout) {
      float a = 0.f;
      for (int p2 = 0; p2 < ppar; ++p2) a += red[p2 * cv + v][e];
      atomicAdd(dw + static_cast<long long>(o) * cin + v * 8 + k, a);
    }
  }
  if (dbias != nullptr) {
    __syncthreads();
    if (cvi == 0 && pl < ppar)
#pragma unroll
      for (int o = 0; o < kMaxSmallN; ++o) red[pl][o] = bacc[o];
    __syncthreads();
    if (tid < cout) {
      float a = 0.f;
      for (int p2 = 0; p2 < ppar; ++p2) a += red[p2][tid];
      atomicAdd(dbias + tid, a);
    }
  }
}

// weight gradient: block = (tap, pixel chunk); thread = (channel vector, pixel lane); fp32 atomics per block
template <typename T>
__global__ void __launch_bounds__(256) smalln_wgrad_kernel(const float* __restrict__ dy, const T* __restrict__ x,
                                                           int x_ld, int x_coff, int n, int h, int w, int cin, int cout,
                                                           int kh, int kw, int pad, int oh, int ow,
                                                           float* __restrict__ dw, float* __restrict__ dbias) {
  const int cv = cin / 8;
  const int cvl = cv < 256 ? cv : 256;
  const int ppar = 256 / cvl;
  const int tap = blockIdx.y;
  const int r = tap / kw, q = tap % kw;
  const int tid = threadIdx.x;
  const int pl = tid / cvl, cvi = tid % cvl;
  const long long pixels = static_cast<long long>(n) * oh * ow;
  __shared__ float red[256][8 * kMaxSmallN + 1];
  for (int cv0 = 0; cv0 < cv; cv0 += cvl) {
    const int cvec = cv0 + cvi;
    const int ch = cvec * 8;
    float acc[kMaxSmallN][8];
    float bacc[kMaxSmallN];
    for (int o = 0; o < kMaxSmallN; ++o) {
      bacc[o] = 0.f;
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[o][k] = 0.f;
    }
    if (pl < ppar && cvec < cv) {
      for (long long px = static_cast<long long>(blockIdx.x) * ppar + pl; px < pixels;
           px += static_cast<long long>(gridDim.x) * ppar) {
        const float* g = dy + px * cout;
        if (tap == 0 && cvec == 0)
          for (int o = 0; o < cout; ++o) bacc[o] += g[o];
        long long xpix = px;                        // pointwise conv: input pixel == output pixel
        if (kh * kw > 1 || pad != 0) {
          const int ox = static_cast<int>(px % ow), oy = static_cast<int>((px / ow) % oh);
          const int b = static_cast<int>(px / (static_cast<long long>(ow) * oh));
          const int iy = oy + r - pad, ix = ox + q - pad;
          if (iy < 0 || iy >= h || ix < 0 || ix >= w) continue;
          xpix = (static_cast<long long>(b) * h + iy) * w + ix;
        }
        float f[8];
        load8(x + xpix * x_ld + x_coff + ch, f);
        for (int o = 0; o < cout; ++o) {
          const float gv = g[o];
#pragma unroll
          for (int k = 0; k < 8; ++k) acc[o][k] = fmaf(gv, f[k], acc[o][k]);
        }
      }
    }
    for (int o = 0; o < kMaxSmallN; ++o)
#pragma unroll
      for (int k = 0; k < 8; ++k) red[tid][o * 8 + k] = acc[o][k];
    __syncthreads();
    if (pl == 0 && cvec < cv) {
      for (int o = 0; o < cout; ++o)
        for (int k = 0; k < 8; ++k) {
          float a = 0.f;
          for (int p2 = 0; p2 < ppar; ++p2) a += red[p2 * cvl + cvi][o * 8 + k];
          atomicAdd(dw + (static_cast<long long>(o) * kh * kw + tap) * cin + ch + k, a);
        }
    }
    __syncthreads();
    if (dbias != nullptr && tap == 0 && cv0 == 0 && cvi == 0 && pl < ppar)
      for (int o = 0; o < cout; ++o)
        if (bacc[o] != 0.f) atomicAdd(dbias + o, bacc[o]);
  }
}


// ---------------------------------------------------------------------------------------------
// MaxPool2d(2, 2) on NHWC activations (VGG16 features of the perceptual loss, vae-gan.py:300-311); the backward
// routes each pooled gradient to the FIRST maximum of its window (PyTorch's tie rule), recomputed from the input
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void maxpool_fwd_kernel(const T* __restrict__ x, int x_ld, T* __restrict__ y, int n, int h, int w, int c) {
  const int cv = c / 8, ph = h / 2, pw = w / 2;
  const long long total = static_cast<long long>(n) * ph * pw * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long cell = i / cv;
    const int ch = static_cast<int>(i - cell * cv) * 8;
    const int pj = static_cast<int>(cell % pw);
    const long long r = cell / pw;
    const int pi = static_cast<int>(r % ph);
    const long long b = r / ph;
    const long long row0 = (b * h + 2 * pi) * w + 2 * pj;
    float m[8], f[8];
    load8(x + row0 * x_ld + ch, m);
    const long long offs[3] = {row0 + 1, row0 + w, row0 + w + 1};
#pragma unroll
    for (int u = 0; u < 3; ++u) {
      load8(x + offs[u] * x_ld + ch, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) m[k] = fmaxf(m[k], f[k]);
    }
    store8(y + cell * c + ch, m);
  }
}
template <typename T>
__global__ void maxpool_bwd_kernel(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, T* __restrict__ dx, int n,
                                   int h, int w, int c) {
  const int cv = c / 8, ph = h / 2, pw = w / 2;
  const long long total = static_cast<long long>(n) * ph * pw * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long cell = i / cv;
    const int ch = static_cast<int>(i - cell * cv) * 8;
    const int pj = static_cast<int>(cell % pw);
    const long long r = cell / pw;
    const int pi = static_cast<int>(r % ph);
    const long long b = r / ph;
    const long long row0 = (b * h + 2 * pi) * w + 2 * pj;
    const long long offs[4] = {row0, row0 + 1, row0 + w, row0 + w + 1};
    float f[4][8], g[8];
#pragma unroll
    for (int u = 0; u < 4; ++u) load8(x + offs[u] * x_ld + ch, f[u]);
    load8(dy + cell * c + ch, g);
    int best[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      best[k] = 0;
      float bv = f[0][k];
#pragma unroll
      for (int u = 1; u < 4; ++u)
        if (f[u][k] > bv) { bv = f[u][k]; best[k] = u; }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      float o[8];
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = best[k] == u ? g[k] : 0.f;
      store8(dx + offs[u] * c + ch, o);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// plain activation backward (for activations fused into a conv epilogue): dx = dy * act'(y)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void act_bwd_kernel(const T* __restrict__ y, int y_ld, const T* __restrict__ dy,
                               int dy_ld, T* __restrict__ dx, int dx_ld, long long rows, int c, int act) {
  const int cv = c / 8;
  const long long total = rows * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cv;
    const int ch = static_cast<int>(i % cv) * 8;
    float yv[8], d[8], o[8];
    load8(y + r * y_ld + ch, yv);
    load8(dy + r * dy_ld + ch, d);
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = d[k] * (yv[k] > 0.f ? 1.f : (act == 2 ? 0.2f : 0.f));
    store8(dx + r * dx_ld + ch, o);
  }
}
// in-place activation (used where it cannot be fused into the producing epilogue, e.g. split-K launches)
template <typename T>
__global__ void act_fwd_kernel(T* __restrict__ y, int y_ld, long long rows, int c, int act) {
  const int cv = c / 8;
  const long long total = rows * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cv;
    const int ch = static_cast<int>(i % cv) * 8;
    float f[8];
    load8(y + r * y_ld + ch, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = f[k] > 0.f ? f[k] : (act == 2 ? 0.2f * f[k] : 0.f);
    store8(y + r * y_ld + ch, f);
  }
}
// out[c] (=|+=) sum_r in[r][c]   (fp32; small matrices: bias gradients of the heads)
__global__ void colsum_f32_kernel(const float* __restrict__ in, long long rows, int cols, int ld, float* __restrict__ out,
                                  int accumulate) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= cols) return;
  float acc = 0.f;
  for (long long r = 0; r < rows; ++r) acc += in[r * ld + c];
  out[c] = accumulate ? out[c] + acc : acc;
}

// fp32 [rows][c] -> three bf16 planes hi | mid | lo (hi + mid + lo == x to ~2^-24): out[rows][3*cp], plane p at columns
// [p*cp, (p+1)*cp), columns c..cp zero.  Feeds the tensor core in the high-accuracy (split-bf16) mode.
__global__ void split3_kernel(const float* __restrict__ in, int ld_in, long long rows, int c, int cp,
                              __nv_bfloat16* __restrict__ out) {
  const long long total = rows * cp;
  const __nv_bfloat16 zero = __float2bfloat16(0.f);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int col = static_cast<int>(i % cp);
    const long long r = i / cp;
    __nv_bfloat16 h = zero, m = zero, l = zero;
    if (col < c) {
      const float x = in[r * ld_in + col];
      h = __float2bfloat16(x);
      const float r1 = x - __bfloat162float(h);
      m = __float2bfloat16(r1);
      l = __float2bfloat16(r1 - __bfloat162float(m));
    }
    __nv_bfloat16* o = out + r * 3 * cp + col;
    o[0] = h; o[cp] = m; o[2 * cp] = l;
  }
}

}  // namespace vg

using namespace vg;

// Host walk of the tiled plan on fp32 host buffers (same decode code as the kernel): lets the CPU test suite check
// the tiling logic without a GPU.  Returns the number of tiles (0 = the plan falls back to the generic kernel).
extern "C" long long vg_debug_copy_plan_host(const float* in, float* out, const long long* dims,
                                             const long long* in_strides, const long long* out_strides) {
  TiledCopyParams p;
  const long long tiles = plan_tiled_copy(dims, in_strides, out_strides, p);
  static float sm[kCopySmem];
  for (long long bid = 0; bid < tiles; ++bid) {
    int ext[5];
    const long long ibase = copy_origin(p.ld, static_cast<unsigned>(bid), ext);
    for (int l = 0; l < p.tile_elems; ++l) {
      int goff = 0, soff = 0;
      if (copy_decode(p.ld, ext, l, goff, soff)) sm[soff] = in[ibase + goff];
    }
    const long long obase = copy_origin(p.st, static_cast<unsigned>(bid), ext);
    for (int l = 0; l < p.tile_elems; ++l) {
      int goff = 0, soff = 0;
      if (copy_decode(p.st, ext, l, goff, soff)) out[obase + goff] = sm[soff];
    }
  }
  return tiles;
}

extern "C" int vg_strided_copy(const void* in, int in_dtype, void* out, int out_dtype, const long long* dims,
                               const long long* in_strides, const long long* out_strides, const float* scale,
                               int scale_inverse, int accumulate, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  CopyParams p;
  long long total = 1;
  for (int i = 0; i < 5; ++i) {
    p.dims[i] = dims[i]; p.is[i] = in_strides[i]; p.os[i] = out_strides[i];
    VG_CHECK(dims[i] >= 1, -1, "vg_strided_copy: dims must be >= 1");
    total *= dims[i];
  }
  VG_CHECK(in_dtype >= 0 && in_dtype <= 1 && out_dtype >= 0 && out_dtype <= 1, -1,
           "vg_strided_copy: dtype codes are 0 (fp32) and 1 (bf16)");
  long long dense_total = 0;
  if (same_order_dense(dims, in_strides, out_strides, &dense_total) && dense_total % 8 == 0 &&
      (reinterpret_cast<uintptr_t>(in) & 31) == 0 && (reinterpret_cast<uintptr_t>(out) & 31) == 0) {
    const long long n8 = dense_total / 8;
    const int gd = ew_grid(n8);
#define VG_CONTIG_DISPATCH(TI, TO)                                                                                      \
  contig_copy_kernel<TI, TO><<<gd, 256, 0, st>>>(static_cast<const TI*>(in), static_cast<TO*>(out), n8, scale,        \
                                                 scale_inverse, accumulate)
    if (in_dtype == 0 && out_dtype == 0) VG_CONTIG_DISPATCH(float, float);
    else if (in_dtype == 0 && out_dtype == 1) VG_CONTIG_DISPATCH(float, __nv_bfloat16);
    else if (in_dtype == 1 && out_dtype == 0) VG_CONTIG_DISPATCH(__nv_bfloat16, float);
    else VG_CONTIG_DISPATCH(__nv_bfloat16, __nv_bfloat16);
#undef VG_CONTIG_DISPATCH
    VG_LAUNCH_OK();
    return 0;
  }
  TiledCopyParams tp;
  const long long tiles = plan_tiled_copy(dims, in_strides, out_strides, tp);
  const int g = tiles > 0 ? static_cast<int>(tiles) : ew_grid(total);
#define VG_COPY_DISPATCH(TI, TO)                                                                                      \
  do {                                                                                                                \
    if (tiles > 0)                                                                                                    \
      tiled_copy_kernel<TI, TO><<<g, 256, 0, st>>>(static_cast<const TI*>(in), static_cast<TO*>(out), tp, scale,      \
                                                   scale_inverse, accumulate);                                        \
    else                                                                                                              \
      strided_copy_kernel<TI, TO><<<g, 256, 0, st>>>(static_cast<const TI*>(in), static_cast<TO*>(out), p, scale,     \
                                                     scale_inverse, accumulate);                                      \
  } while (0)
  if (in_dtype == 0 && out_dtype == 0) VG_COPY_DISPATCH(float, float);
  else if (in_dtype == 0 && out_dtype == 1) VG_COPY_DISPATCH(float, __nv_bfloat16);
  else if (in_dtype == 1 && out_dtype == 0) VG_COPY_DISPATCH(__nv_bfloat16, float);
  else VG_COPY_DISPATCH(__nv_bfloat16, __nv_bfloat16);
#undef VG_COPY_DISPATCH
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_film_fwd(const void* gb, const void* x, int x_ld, int x_coff, void* y, long long rows, int c,
                           int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1, "vg_film_fwd: channels must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (dtype == 0)
    film_fwd_kernel<__nv_bfloat16><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(gb), static_cast<const __nv_bfloat16*>(x), x_ld, x_coff,
        static_cast<__nv_bfloat16*>(y), rows, c);
  else
    film_fwd_kernel<float><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(static_cast<const float*>(gb), static_cast<const float*>(x),
                                                                  x_ld, x_coff, static_cast<float*>(y), rows, c);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_film_bwd(const void* gb, const void* x, int x_ld, int x_coff, const void* dy, void* dgb, void* dx,
                           int dx_ld, int dx_coff, long long rows, int c, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0 && dx_ld % 8 == 0 && dx_coff % 8 == 0, -1,
           "vg_film_bwd: channels must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (dtype == 0)
    film_bwd_kernel<__nv_bfloat16><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(gb), static_cast<const __nv_bfloat16*>(x), x_ld, x_coff,
        static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dgb), static_cast<__nv_bfloat16*>(dx), dx_ld,
        dx_coff, rows, c);
  else
    film_bwd_kernel<float><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(
        static_cast<const float*>(gb), static_cast<const float*>(x), x_ld, x_coff, static_cast<const float*>(dy),
        static_cast<float*>(dgb), static_cast<float*>(dx), dx_ld, dx_coff, rows, c);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_film_rows_fwd(const void* gb, int gh, const void* x, int x_ld, int x_coff, void* y, int n, int h, int w,
                                int c, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1, "vg_film_rows_fwd: channels must be multiples of 8");
  VG_CHECK(gh == 3 && h >= 3, -1, "vg_film_rows_fwd: the parameter map must have 3 rows and the image at least 3");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * h * w * (c / 8));
  if (dtype == 0)
    film_rows_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(gb), gh,
                                                            static_cast<const __nv_bfloat16*>(x), x_ld, x_coff,
                                                            static_cast<__nv_bfloat16*>(y), n, h, w, c);
  else
    film_rows_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(gb), gh, static_cast<const float*>(x), x_ld,
                                                    x_coff, static_cast<float*>(y), n, h, w, c);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_film_rows_bwd(const void* gb, int gh, const void* x, int x_ld, int x_coff, const void* dy, void* dgb,
                                void* dx, int dx_ld, int dx_coff, int n, int h, int w, int c, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0 && dx_ld % 8 == 0 && dx_coff % 8 == 0, -1,
           "vg_film_rows_bwd: channels must be multiples of 8");
  VG_CHECK(gh == 3 && h >= 3, -1, "vg_film_rows_bwd: the parameter map must have 3 rows and the image at least 3");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const long long items = static_cast<long long>(n) * w * (c / 8);
  const int grid = static_cast<int>((items + 127) / 128);
  if (dtype == 0)
    film_rows_bwd_kernel<__nv_bfloat16><<<grid, 128, 0, st>>>(
        static_cast<const __nv_bfloat16*>(gb), gh, static_cast<const __nv_bfloat16*>(x), x_ld, x_coff,
        static_cast<const __nv_bfloat16*>(dy), static_cast<__nv_bfloat16*>(dgb), static_cast<__nv_bfloat16*>(dx), dx_ld,
        dx_coff, n, h, w, c);
  else
    film_rows_bwd_kernel<float><<<grid, 128, 0, st>>>(static_cast<const float*>(gb), gh, static_cast<const float*>(x), x_ld,
                                                    x_coff, static_cast<const float*>(dy), static_cast<float*>(dgb),
                                                    static_cast<float*>(dx), dx_ld, dx_coff, n, h, w, c);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_upsample_w_fwd(const void* t, int t_ld, int t_coff, int n, int w0, int c, void* y, int h, int w,
                                 int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && t_ld % 8 == 0 && t_coff % 8 == 0, -1, "vg_upsample_w_fwd: channels must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int rpt = h >= 64 ? 8 : (h >= 16 ? 4 : 1);
  const int grid = ew_grid(static_cast<long long>(n) * ((h + rpt - 1) / rpt) * w * (c / 8));
  if (dtype == 0)
    upsample_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(t), t_ld, t_coff, n, w0, c,
                                                           static_cast<__nv_bfloat16*>(y), h, w, rpt);
  else
    upsample_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(t), t_ld, t_coff, n, w0, c,
                                                   static_cast<float*>(y), h, w, rpt);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_upsample_w_bwd(const void* dy, int n, int h, int w, int c, int w0, float* dt, int dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(c % 8 == 0, -1, "vg_upsample_w_bwd: channels must be multiples of 8");
  VG_CUDA(cudaMemsetAsync(dt, 0, sizeof(float) * static_cast<size_t>(n) * w0 * c, st));
  // rows per thread: the fp32 atomics that flush a thread's partial sums, not the loads, bound this kernel (shorter column runs
  // or fewer rows per thread = more atomics per byte read: 0.15 ... 0.66 of the copy bandwidth); 32 rows per thread: 0.79 at 128 x 128
  const int rpt = h >= 64 ? 32 : (h >= 32 ? 16 : 8);
  // a run of jlen columns spans at most jlen * w0 / w < 1 source steps plus the two-tap footprint: <= 3 sources
  const int jlen = std::max(1, w / (2 * w0));
  const long long items = static_cast<long long>(n) * ((w + jlen - 1) / jlen) * (c / 8) * ((h + rpt - 1) / rpt);
  if (dtype == 0)
    upsample_bwd_kernel<__nv_bfloat16><<<ew_grid(items), 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), n, h, w, c, w0, dt,
                                                                     rpt, jlen);
  else
    upsample_bwd_kernel<float><<<ew_grid(items), 256, 0, st>>>(static_cast<const float*>(dy), n, h, w, c, w0, dt, rpt, jlen);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_upsample_h_fwd(const void* t, int n, int h0, int w, int c, void* y, int h, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && h0 >= 1 && h >= 1, -1, "vg_upsample_h_fwd: channels must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * h * w * (c / 8));
  if (dtype == 0)
    upsample_h_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(t), n, h0, w, c,
                                                             static_cast<__nv_bfloat16*>(y), h);
  else
    upsample_h_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(t), n, h0, w, c, static_cast<float*>(y), h);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_upsample_h_bwd(const void* dy, int n, int h, int w, int c, int h0, float* dt, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && h0 >= 1 && h >= 1, -1, "vg_upsample_h_bwd: channels must be a multiple of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * h0 * w * (c / 8));
  if (dtype == 0)
    upsample_h_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dy), n, h, w, c, h0, dt);
  else
    upsample_h_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dy), n, h, w, c, h0, dt);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_im2col(const void* src, int n, int h, int w, int ld, int c, int kh, int kw, int stride, int pad,
                         void* col, int kpad, int dtype, void* stream_) {
  VG_CHECK(kh * kw * c <= kpad && kpad % 64 == 0, -1, "vg_im2col: kpad must be a multiple of 64 >= kh*kw*c");
  VG_CHECK(kpad <= kIm2colMaxK && kh < 256 && kw < 256 && c < 256, -1, "vg_im2col: kpad <= 1024, kernel extents and c < 256");
  VG_CHECK(static_cast<long long>(n) * h * w * ld < (1LL << 31), -1, "vg_im2col: image tensor too large");
  const int oh = (h + 2 * pad - kh) / stride + 1, ow = (w + 2 * pad - kw) / stride + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * oh * ow * (kpad / 8));
  if (dtype == 0)
    im2col_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(src), n, h, w, ld, c, kh, kw, stride,
                                                     pad, oh, ow, static_cast<__nv_bfloat16*>(col), kpad);
  else
    im2col_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(src), n, h, w, ld, c, kh, kw, stride, pad, oh, ow,
                                             static_cast<float*>(col), kpad);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_col2im(const void* dcol, int kpad, int n, int h, int w, int c, int kh, int kw, int stride, int pad,
                         float* dsrc_nchw, int dtype, void* stream_) {
  const int oh = (h + 2 * pad - kh) / stride + 1, ow = (w + 2 * pad - kw) / stride + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * c * h * w);
  if (dtype == 0)
    col2im_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(dcol), kpad, n, h, w, c, kh, kw,
                                                     stride, pad, oh, ow, dsrc_nchw);
  else
    col2im_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(dcol), kpad, n, h, w, c, kh, kw, stride, pad, oh, ow,
                                             dsrc_nchw);
  VG_LAUNCH_OK();
  return 0;
}

static int smalln_lanes(int kvecs) {
  int g = 1;
  while (g < 32 && g < kvecs) g <<= 1;
  return g;
}
extern "C" int vg_conv_smalln_fwd(const void* x, int x_ld, int x_coff, int n, int h, int w, int cin, const float* wt,
                                  const float* bias, int cout, int kh, int kw, int pad, float* out, int dtype,
                                  void* stream_) {
  VG_CHECK(cout >= 1 && cout <= kMaxSmallN && cin % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1,
           "vg_conv_smalln_fwd: cout must be 1..4 and channels multiples of 8");
  const int oh = h + 2 * pad - kh + 1, ow = w + 2 * pad - kw + 1;
  const int G = smalln_lanes(kh * kw * (cin / 8));
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * oh * ow * G);
  if (dtype == 0)
    smalln_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_ld, x_coff, n, h, w, cin, wt,
                                                         bias, cout, kh, kw, pad, oh, ow, out, G);
  else
    smalln_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), x_ld, x_coff, n, h, w, cin, wt, bias, cout,
                                                 kh, kw, pad, oh, ow, out, G);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_conv_smalln_dgrad(const float* dy, int n, int h, int w, int cin, const float* wt, int cout, int kh,
                                    int kw, int pad, void* dx, int dx_ld, int dx_coff, int dtype, void* stream_) {
  VG_CHECK(cout >= 1 && cout <= kMaxSmallN && cin % 8 == 0 && dx_ld % 8 == 0 && dx_coff % 8 == 0, -1,
           "vg_conv_smalln_dgrad: cout must be 1..4 and channels multiples of 8");
  const int oh = h + 2 * pad - kh + 1, ow = w + 2 * pad - kw + 1;
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * h * w * (cin / 8));
  if (dtype == 0)
    smalln_dgrad_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(dy, n, oh, ow, cout, wt, kh, kw, pad, h, w, cin,
                                                           static_cast<__nv_bfloat16*>(dx), dx_ld, dx_coff);
  else
    smalln_dgrad_kernel<float><<<grid, 256, 0, st>>>(dy, n, oh, ow, cout, wt, kh, kw, pad, h, w, cin, static_cast<float*>(dx),
                                                   dx_ld, dx_coff);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_conv_smalln_wgrad(const float* dy, const void* x, int x_ld, int x_coff, int n, int h, int w, int cin,
                                    int cout, int kh, int kw, int pad, float* dw, float* dbias, int dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(cout >= 1 && cout <= kMaxSmallN && cin % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1,
           "vg_conv_smalln_wgrad: cout must be 1..4 and channels multiples of 8");
  const int oh = h + 2 * pad - kh + 1, ow = w + 2 * pad - kw + 1;
  VG_CUDA(cudaMemsetAsync(dw, 0, sizeof(float) * static_cast<size_t>(cout) * kh * kw * cin, st));
  if (dbias) VG_CUDA(cudaMemsetAsync(dbias, 0, sizeof(float) * cout, st));
  const int cv = cin / 8, cvl = cv < 256 ? cv : 256, ppar = 256 / cvl;
  const long long pixels = static_cast<long long>(n) * oh * ow;
  if (kh == 1 && kw == 1 && pad == 0 && cv <= 32 && 256 % cv == 0) {
    const int gp = static_cast<int>(std::min<long long>((pixels + ppar * 8 - 1) / (ppar * 8), static_cast<long long>(num_sms()) * 8));
    if (dtype == 0)
      pw_wgrad_kernel<__nv_bfloat16><<<std::max(gp, 1), 256, 0, st>>>(dy, static_cast<const __nv_bfloat16*>(x), x_ld, x_coff, pixels,
                                                                    cin, cout, dw, dbias);
    else
      pw_wgrad_kernel<float><<<std::max(gp, 1), 256, 0, st>>>(dy, static_cast<const float*>(x), x_ld, x_coff, pixels, cin, cout,
                                                            dw, dbias);
    VG_LAUNCH_OK();
    return 0;
  }
  int gx = static_cast<int>(std::min<long long>((pixels + ppar * 16 - 1) / (ppar * 16), std::max(1, num_sms() * 4 / (kh * kw))));
  if (gx < 1) gx = 1;
  if (dtype == 0)
    smalln_wgrad_kernel<__nv_bfloat16><<<dim3(gx, kh * kw), 256, 0, st>>>(dy, static_cast<const __nv_bfloat16*>(x), x_ld, x_coff,
                                                                        n, h, w, cin, cout, kh, kw, pad, oh, ow, dw, dbias);
  else
    smalln_wgrad_kernel<float><<<dim3(gx, kh * kw), 256, 0, st>>>(dy, static_cast<const float*>(x), x_ld, x_coff, n, h, w, cin,
                                                                cout, kh, kw, pad, oh, ow, dw, dbias);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_maxpool2x2_fwd(const void* x, int x_ld, void* y, int n, int h, int w, int c, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && h % 2 == 0 && w % 2 == 0, -1, "vg_maxpool2x2_fwd: c % 8, even H and W");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8));
  if (dtype == 0)
    maxpool_fwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_ld,
                                                          static_cast<__nv_bfloat16*>(y), n, h, w, c);
  else
    maxpool_fwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), x_ld, static_cast<float*>(y), n, h, w, c);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_maxpool2x2_bwd(const void* x, int x_ld, const void* dy, void* dx, int n, int h, int w, int c, int dtype,
                                 void* stream_) {
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && h % 2 == 0 && w % 2 == 0, -1, "vg_maxpool2x2_bwd: c % 8, even H and W");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int grid = ew_grid(static_cast<long long>(n) * (h / 2) * (w / 2) * (c / 8));
  if (dtype == 0)
    maxpool_bwd_kernel<__nv_bfloat16><<<grid, 256, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_ld,
                                                          static_cast<const __nv_bfloat16*>(dy),
                                                          static_cast<__nv_bfloat16*>(dx), n, h, w, c);
  else
    maxpool_bwd_kernel<float><<<grid, 256, 0, st>>>(static_cast<const float*>(x), x_ld, static_cast<const float*>(dy),
                                                  static_cast<float*>(dx), n, h, w, c);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_act_bwd(const void* y, int y_ld, const void* dy, int dy_ld, void* dx, int dx_ld, long long rows, int c,
                          int act, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && y_ld % 8 == 0 && dy_ld % 8 == 0 && dx_ld % 8 == 0, -1, "vg_act_bwd: channels must be multiples of 8");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (dtype == 0)
    act_bwd_kernel<__nv_bfloat16><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(
        static_cast<const __nv_bfloat16*>(y), y_ld, static_cast<const __nv_bfloat16*>(dy), dy_ld,
        static_cast<__nv_bfloat16*>(dx), dx_ld, rows, c, act);
  else
    act_bwd_kernel<float><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(static_cast<const float*>(y), y_ld,
                                                                 static_cast<const float*>(dy), dy_ld,
                                                                 static_cast<float*>(dx), dx_ld, rows, c, act);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_colsum_f32(const float* in, long long rows, int cols, int ld, float* out, int accumulate, void* stream_) {
  colsum_f32_kernel<<<cdiv(cols, 128), 128, 0, static_cast<cudaStream_t>(stream_)>>>(in, rows, cols, ld, out, accumulate);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_split3(const float* in, int ld_in, long long rows, int c, int cp, void* out, void* stream_) {
  VG_CHECK(cp >= c && cp % 64 == 0, -1, "vg_split3: cp must be a multiple of 64 >= c");
  split3_kernel<<<ew_grid(rows * cp), 256, 0, static_cast<cudaStream_t>(stream_)>>>(in, ld_in, rows, c, cp,
                                                                                  static_cast<__nv_bfloat16*>(out));
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_act_fwd(void* y, int y_ld, long long rows, int c, int act, int dtype, void* stream_) {
  VG_CHECK(c % 8 == 0 && y_ld % 8 == 0 && (act == 1 || act == 2), -1, "vg_act_fwd: channels must be multiples of 8, act 1|2");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  if (dtype == 0) act_fwd_kernel<__nv_bfloat16><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(static_cast<__nv_bfloat16*>(y), y_ld, rows, c, act);
  else act_fwd_kernel<float><<<ew_grid(rows * (c / 8)), 256, 0, st>>>(static_cast<float*>(y), y_ld, rows, c, act);
  VG_LAUNCH_OK();
  return 0;
}
