// Normalisation + activation (+2x2 max-pool) kernels on NHWC bf16 activations, forward and backward.
//
// One family serves both normalisations of the reference:
//   * BatchNorm2d in training mode + ReLU  (generator; vae-gan.py:52-55,76-81; vae-gan-v2.py:171-177) --
//     statistics over all N*H*W rows ("groups = 1"), running stats updated with momentum 0.1 / unbiased var;
//   * InstanceNorm2d(affine) + LeakyReLU(0.2)  (discriminator; vae-gan.py:154-156) -- statistics per sample
//     ("groups = N"), no running stats.
// MaxPool2d(2,2) (vae-gan-v2.py:157-163) is fused into the apply pass: the kernel writes the full-resolution
// activation (the U-Net skip, possibly into a channel slice of the decoder's concat buffer) and the pooled
// tensor in one sweep; its backward routes the pooled gradient to the first maximum of each window.
//
// All kernels are HBM-bound.  Layout of the work: a thread owns ONE 8-channel vector (16-byte accesses) for its
// whole lifetime, so the per-channel constants (scale, shift, mean, rstd, ...) are loaded once into registers,
// and it walks rows with several independent loads in flight.  Reductions: registers -> shared memory across
// the rows of a block -> one fp32 atomic per (block, channel); grids are kept small (<= 2 blocks per SM per
// group) so that the atomics on one address do not serialise.
#include <stdlib.h>
#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

constexpr int kNT = 256;

struct RowMap {   // thread -> (row lane, channel vector) for C/8 channel vectors
  int cv, cvl, rows_par;
};
__host__ __device__ inline RowMap row_map(int c) {
  RowMap m;
  m.cv = c / 8;
  m.cvl = m.cv < kNT ? m.cv : kNT;
  m.rows_par = kNT / m.cvl;
  return m;
}

VG_DEVICE float act_fwd(float v, int act) {
  if (act == 1) return fmaxf(v, 0.f);
  if (act == 2) return v > 0.f ? v : 0.2f * v;
  return v;
}
VG_DEVICE float act_grad(float pre, int act) {
  if (act == 1) return pre > 0.f ? 1.f : 0.f;
  if (act == 2) return pre > 0.f ? 1.f : 0.2f;
  return 1.f;
}
// block-level reduction of 16 per-thread partials over the row lanes, then one atomic per channel
VG_DEVICE void reduce_rows_atomic(float (&red)[kNT][17], const float (&s)[8], const float (&q)[8], const RowMap& m,
                                  int rl, int cvi, int cvec, float* dst0, float* dst1) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < 8; ++i) { red[tid][i] = s[i]; red[tid][8 + i] = q[i]; }
  __syncthreads();
  // all 256 threads cooperate: thread t reduces partial (t % 16) of channel vector (t / 16) when it exists
  for (int item = tid; item < m.cvl * 16; item += kNT) {
    const int v = item >> 4, k = item & 15;
    if (cvec - cvi + v < m.cv) {
      float a = 0.f;
      for (int r = 0; r < m.rows_par; ++r) a += red[r * m.cvl + v][k];
      float* dst = (k < 8 ? dst0 : dst1) + (cvec - cvi + v) * 8 + (k & 7);
      atomicAdd(dst, a);
    }
  }
  __syncthreads();
  (void)rl;
}

// ---------------------------------------------------------------------------------------------
// statistics: sums[g][0][c] = sum x, sums[g][1][c] = sum x^2 over the rows of group g
// ---------------------------------------------------------------------------------------------
// Row classes ("virtual height"): a tensor of cls_h rows per image may stand for a taller one whose interior rows are
// all equal (the FiLM parameter maps, vae-gan-v2.py:138-145: a 3x3 conv of a row-constant input differs only in its
// first and last row).  Row 0 and row cls_h-1 then count once, every other row `wmid` times.
VG_DEVICE float row_class_weight(long long pixel, int cls_w, int cls_h, float wmid) {
  const int yy = static_cast<int>((pixel / cls_w) % cls_h);
  return (yy == 0 || yy == cls_h - 1) ? 1.f : wmid;
}

template <typename T>
__global__ void __launch_bounds__(kNT) stats_kernel(const T* __restrict__ x, int ld, int coff, int c,
                                                    long long rows_per_group, float* __restrict__ sums, int cls_w,
                                                    int cls_h, float wmid) {
  const RowMap m = row_map(c);
  const int g = blockIdx.y;
  const int tid = threadIdx.x;
  const int rl = tid / m.cvl, cvi = tid % m.cvl;
  __shared__ float red[kNT][17];
  for (int cv0 = 0; cv0 < m.cv; cv0 += m.cvl) {
    const int cvec = cv0 + cvi;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (rl < m.rows_par && cvec < m.cv) {
      const T* base = x + static_cast<long long>(g) * rows_per_group * ld + coff + cvec * 8;
      const long long stride = static_cast<long long>(gridDim.x) * m.rows_par;
      long long r = static_cast<long long>(blockIdx.x) * m.rows_par + rl;
      if (cls_h > 0) {      // weighted rows (small tensors only): plain loop
        for (; r < rows_per_group; r += stride) {
          float f[8];
          Raw8<T>::load(base + r * ld).unpack(f);
          const float wt = row_class_weight(r, cls_w, cls_h, wmid);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] = fmaf(wt, f[i], s[i]); q[i] = fmaf(wt * f[i], f[i], q[i]); }
        }
      }
      for (; r + 3 * stride < rows_per_group; r += 4 * stride) {
        Raw8<T> v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = Raw8<T>::load(base + (r + u * stride) * ld);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          v[u].unpack(f);
#pragma unroll
          for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
        }
      }
      for (; r < rows_per_group; r += stride) {
        float f[8];
        Raw8<T>::load(base + r * ld).unpack(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] = fmaf(f[i], f[i], q[i]); }
      }
    }
    float* dst = sums + static_cast<long long>(g) * 2 * c;
    reduce_rows_atomic(red, s, q, m, rl, cvi, cvec, dst, dst + c);
  }
}

// mean / rstd (+ BatchNorm running statistics); one thread per (group, channel)
__global__ void finalize_kernel(const float* __restrict__ sums, int groups, int c, long long rows_per_group, float eps,
                                float* __restrict__ mean_rstd, float momentum, float* running_mean, float* running_var,
                                long long* num_batches_tracked) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0 && num_batches_tracked != nullptr) *num_batches_tracked += 1;
  if (i >= groups * c) return;
  const int g = i / c, ch = i % c;
  const double n = static_cast<double>(rows_per_group);
  const double mean = sums[(g * 2 + 0) * c + ch] / n;
  double var = sums[(g * 2 + 1) * c + ch] / n - mean * mean;
  if (var < 0.0) var = 0.0;
  mean_rstd[(g * 2 + 0) * c + ch] = static_cast<float>(mean);
  mean_rstd[(g * 2 + 1) * c + ch] = static_cast<float>(1.0 / sqrt(var + static_cast<double>(eps)));
  if (running_mean != nullptr && groups == 1) {
    const double unbiased = n > 1.0 ? var * n / (n - 1.0) : var;
    running_mean[ch] = (1.f - momentum) * running_mean[ch] + momentum * static_cast<float>(mean);
    running_var[ch] = (1.f - momentum) * running_var[ch] + momentum * static_cast<float>(unbiased);
  }
}

// ---------------------------------------------------------------------------------------------
// apply: y = act(gamma * (x - mean) * rstd + beta), optional fused 2x2 max-pool output
// blockIdx.y = sample when per_sample (so the constants are block-uniform per channel), else 0
// ---------------------------------------------------------------------------------------------
template <typename T>
struct ApplyParams {
  const T* x; int x_ld, x_coff;
  int n, h, w, c;
  const float* mean_rstd; int per_sample;
  const float* gamma; const float* beta;
  int act;
  T* y; int y_ld, y_coff;
  T* pool; int p_ld, p_coff;
};

template <typename T, bool kPool>
__global__ void __launch_bounds__(kNT) apply_kernel(const ApplyParams<T> p) {
  const RowMap m = row_map(p.c);
  const int tid = threadIdx.x;
  const int rl = tid / m.cvl, cvi = tid % m.cvl;
  const int grp = blockIdx.y;
  const int n_count = p.per_sample ? 1 : p.n;
  const int ph = kPool ? p.h / 2 : p.h, pw = kPool ? p.w / 2 : p.w;
  const long long cells = static_cast<long long>(n_count) * ph * pw;
  const long long stride = static_cast<long long>(gridDim.x) * m.rows_par;
  for (int cv0 = 0; cv0 < m.cv; cv0 += m.cvl) {
    const int cvec = cv0 + cvi;
    if (rl >= m.rows_par || cvec >= m.cv) continue;
    const int ch = cvec * 8;
    const float* mr = p.mean_rstd + static_cast<long long>(p.per_sample ? grp : 0) * 2 * p.c;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = p.gamma ? p.gamma[ch + i] : 1.f, b = p.beta ? p.beta[ch + i] : 0.f;
      const float rstd = mr[p.c + ch + i];
      sc[i] = g * rstd;
      sh[i] = b - mr[ch + i] * g * rstd;
    }
    const T* xb = p.x + p.x_coff + ch;
    T* yb = p.y + p.y_coff + ch;
    const long long pix0 = static_cast<long long>(p.per_sample ? grp : 0) * p.h * p.w;
    if (!kPool) {
      long long r = static_cast<long long>(blockIdx.x) * m.rows_par + rl;
      for (; r + 3 * stride < cells; r += 4 * stride) {
        Raw8<T> v[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) v[u] = Raw8<T>::load(xb + (pix0 + r + u * stride) * p.x_ld);
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          v[u].unpack(f);
#pragma unroll
          for (int i = 0; i < 8; ++i) f[i] = act_fwd(fmaf(f[i], sc[i], sh[i]), p.act);
          store8(yb + (pix0 + r + u * stride) * p.y_ld, f);
        }
      }
      for (; r < cells; r += stride) {
        float f[8];
        Raw8<T>::load(xb + (pix0 + r) * p.x_ld).unpack(f);
#pragma unroll
        for (int i = 0; i < 8; ++i) f[i] = act_fwd(fmaf(f[i], sc[i], sh[i]), p.act);
        store8(yb + (pix0 + r) * p.y_ld, f);
      }
    } else {
      for (long long cell = static_cast<long long>(blockIdx.x) * m.rows_par + rl; cell < cells; cell += stride) {
        const int pj = static_cast<int>(cell % pw);
        const int pi = static_cast<int>((cell / pw) % ph);
        const long long nn = (p.per_sample ? grp : cell / (static_cast<long long>(pw) * ph));
        const long long row0 = (nn * p.h + 2 * pi) * p.w + 2 * pj;
        Raw8<T> v[4];
        v[0] = Raw8<T>::load(xb + row0 * p.x_ld);
        v[1] = Raw8<T>::load(xb + (row0 + 1) * p.x_ld);
        v[2] = Raw8<T>::load(xb + (row0 + p.w) * p.x_ld);
        v[3] = Raw8<T>::load(xb + (row0 + p.w + 1) * p.x_ld);
        float mx[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          float f[8];
          v[u].unpack(f);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            f[i] = act_fwd(fmaf(f[i], sc[i], sh[i]), p.act);
            mx[i] = fmaxf(mx[i], as_stored<T>(f[i]));
          }
          store8(yb + (row0 + (u >> 1) * p.w + (u & 1)) * p.y_ld, f);
        }
        const long long ppix = (nn * ph + pi) * pw + pj;
        store8(p.pool + ppix * p.p_ld + p.p_coff + ch, mx);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  g = (dy + routed pooled grad) * act'(pre);  pass 1 (kApply = false) reduces sum g and sum g*xhat,
// pass 2 (kApply = true) writes dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)).
// ---------------------------------------------------------------------------------------------
template <typename T>
struct BwdParams {
  const T* x; int x_ld, x_coff;           // raw (pre-normalisation) conv output
  const T* dy; int dy_ld, dy_coff;        // grad wrt full-resolution activation (nullable)
  const T* dpool; int dp_ld, dp_coff;     // grad wrt pooled activation (nullable)
  int n, h, w, c;
  const float* mean_rstd; int per_sample;
  const float* gamma; const float* beta;
  int act;
  float* sums;                                        // [groups][2][c]: sum g, sum g*xhat
  T* dx; int dx_ld, dx_coff;
  int virt_h;                                         // > 0: the h rows stand for virt_h rows (row classes, see above)
};

// Register budget is what decides the bandwidth of these passes: a thread keeps four per-channel constants for its 8
// channels (32 registers) and, while walking rows, U independent (x, dy) row pairs in flight.  Two 256-thread blocks per SM
// (<= 128 registers) x 8 sixteen-byte loads per thread = 64 KB in flight per SM, enough to cover HBM latency at full
// bandwidth; the first version of this kernel kept six constants per channel, took 148 registers = ONE block per SM with
// four loads per thread, and ran at 0.45-0.75 of the copy bandwidth (0.35 for the pooled variant, which spilled).
//   pre  = x * sc + sh                      (sc = gamma * rstd, sh = beta - mean * sc)      -> activation mask
//   xhat = x * ca + cb                      (reduce pass: ca = rstd, cb = -mean * rstd)
//   dx   = sc * g - wt * (x * ca + cb)      (apply pass:  ca = sc * k2 * rstd, cb = sc * (k1 - k2 * mean * rstd);
//                                            k1 = mean(g), k2 = mean(g * xhat); wt = 1 except for row classes)
template <typename T, bool kApply, bool kPool>
__global__ void __launch_bounds__(kNT, 2) bwd_kernel(const BwdParams<T> p) {
  constexpr int U = sizeof(T) == 2 ? 4 : 2;           // rows in flight per thread (non-pooled walk)
  const RowMap m = row_map(p.c);
  const int tid = threadIdx.x;
  const int rl = tid / m.cvl, cvi = tid % m.cvl;
  const int grp = blockIdx.y;                       // sample index when per_sample, else 0
  const int n_count = p.per_sample ? 1 : p.n;
  const int ph = kPool ? p.h / 2 : p.h, pw = kPool ? p.w / 2 : p.w;
  const long long cells = static_cast<long long>(n_count) * ph * pw;
  const long long stride = static_cast<long long>(gridDim.x) * m.rows_par;
  const float inv_rows = 1.f / (static_cast<float>(n_count) * (p.virt_h > 0 ? p.virt_h : p.h) * p.w);
  const bool virt = kApply && p.virt_h > 0;
  const float wmid = p.virt_h > 0 ? static_cast<float>(p.virt_h - 2) / static_cast<float>(p.h - 2) : 1.f;
  const long long pix0 = static_cast<long long>(p.per_sample ? grp : 0) * p.h * p.w;
  const float neg_slope = p.act == 1 ? 0.f : (p.act == 2 ? 0.2f : 1.f);      // act'(pre) for pre <= 0
  __shared__ float red[kApply ? 1 : kNT][17];
  for (int cv0 = 0; cv0 < m.cv; cv0 += m.cvl) {
    const int cvec = cv0 + cvi;
    const int ch = cvec * 8;
    const bool active = rl < m.rows_par && cvec < m.cv;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (active) {
      const float* mr = p.mean_rstd + static_cast<long long>(grp) * 2 * p.c;
      float sc[8], sh[8], ca[8], cb[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float ga = p.gamma ? p.gamma[ch + i] : 1.f, be = p.beta ? p.beta[ch + i] : 0.f;
        const float mean = mr[ch + i], rstd = mr[p.c + ch + i];
        sc[i] = ga * rstd;
        sh[i] = be - mean * sc[i];
        if (kApply) {
          const float* sm = p.sums + static_cast<long long>(grp) * 2 * p.c;
          const float k1 = sm[ch + i] * inv_rows, k2 = sm[p.c + ch + i] * inv_rows;
          ca[i] = sc[i] * k2 * rstd;
          cb[i] = sc[i] * (k1 - k2 * mean * rstd);
        } else {
          ca[i] = rstd;
          cb[i] = -mean * rstd;
        }
      }
      const T* xb = p.x + p.x_coff + ch;
      const T* gb = p.dy ? p.dy + p.dy_coff + ch : nullptr;
      T* ob = kApply ? p.dx + p.dx_coff + ch : nullptr;

      // one pixel: g = d * act'(pre); accumulate (reduce pass) or emit dx (apply pass)
      auto pixel = [&](const Raw8<T>& xv, float (&d)[8], long long pix) {
        float f[8];
        xv.unpack(f);
        if (kApply) {
          const float wt = virt ? row_class_weight(pix, p.w, p.h, wmid) : 1.f;
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float g = d[i] * (fmaf(f[i], sc[i], sh[i]) > 0.f ? 1.f : neg_slope);
            d[i] = fmaf(sc[i], g, -wt * fmaf(f[i], ca[i], cb[i]));
          }
          store8(ob + pix * p.dx_ld, d);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float g = d[i] * (fmaf(f[i], sc[i], sh[i]) > 0.f ? 1.f : neg_slope);
            s[i] += g;
            q[i] = fmaf(g, fmaf(f[i], ca[i], cb[i]), q[i]);
          }
        }
      };

      if (!kPool) {
        long long r = static_cast<long long>(blockIdx.x) * m.rows_par + rl;
        for (; r + (U - 1) * stride < cells; r += U * stride) {
          Raw8<T> xa[U], da[U];
#pragma unroll
          for (int u = 0; u < U; ++u) xa[u] = Raw8<T>::load(xb + (pix0 + r + u * stride) * p.x_ld);
#pragma unroll
          for (int u = 0; u < U; ++u) da[u] = Raw8<T>::load(gb + (pix0 + r + u * stride) * p.dy_ld);
#pragma unroll
          for (int u = 0; u < U; ++u) {
            float d[8];
            da[u].unpack(d);
            pixel(xa[u], d, pix0 + r + u * stride);
          }
        }
        for (; r < cells; r += stride) {
          const long long pa = pix0 + r;
          float d[8];
          Raw8<T>::load(gb + pa * p.dy_ld).unpack(d);
          pixel(Raw8<T>::load(xb + pa * p.x_ld), d, pa);
        }
      } else {
        for (long long cell = static_cast<long long>(blockIdx.x) * m.rows_par + rl; cell < cells; cell += stride) {
          const int pj = static_cast<int>(cell % pw);
          const int pi = static_cast<int>((cell / pw) % ph);
          const long long nn = (p.per_sample ? grp : cell / (static_cast<long long>(pw) * ph));
          const long long row0 = (nn * p.h + 2 * pi) * p.w + 2 * pj;
          Raw8<T> xv[4], dv[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) xv[u] = Raw8<T>::load(xb + (row0 + (u >> 1) * p.w + (u & 1)) * p.x_ld);
          if (gb != nullptr) {
#pragma unroll
            for (int u = 0; u < 4; ++u) dv[u] = Raw8<T>::load(gb + (row0 + (u >> 1) * p.w + (u & 1)) * p.dy_ld);
          }
          const long long ppix = (nn * ph + pi) * pw + pj;
          const Raw8<T> dpv = Raw8<T>::load(p.dpool + ppix * p.dp_ld + p.dp_coff + ch);
          // first maximum of the (storage-rounded) activations of the window, per channel: 2 bits per channel
          unsigned best = 0;
          {
            float bv[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) bv[i] = -INFINITY;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              float f[8];
              xv[u].unpack(f);
#pragma unroll
              for (int i = 0; i < 8; ++i) {
                const float yv = as_stored<T>(act_fwd(fmaf(f[i], sc[i], sh[i]), p.act));
                if (yv > bv[i]) { bv[i] = yv; best = (best & ~(3u << (2 * i))) | (static_cast<unsigned>(u) << (2 * i)); }
              }
            }
          }
          float dp[8];
          dpv.unpack(dp);
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            float d[8];
            if (gb != nullptr) dv[u].unpack(d);
            else {
#pragma unroll
              for (int i = 0; i < 8; ++i) d[i] = 0.f;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i) d[i] += (((best >> (2 * i)) & 3u) == static_cast<unsigned>(u)) ? dp[i] : 0.f;
            pixel(xv[u], d, row0 + (u >> 1) * p.w + (u & 1));
          }
        }
      }
    }
    if constexpr (!kApply) {
      float* dst = p.sums + static_cast<long long>(grp) * 2 * p.c;
      reduce_rows_atomic(red, s, q, m, rl, cvi, cvec, dst, dst + p.c);
    }
  }
}

// dgamma[c] (+)= sum_g sums[g][1][c], dbeta[c] (+)= sum_g sums[g][0][c]
__global__ void affine_grad_kernel(const float* __restrict__ sums, int groups, int c, float* dgamma, float* dbeta,
                                   int accumulate) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float a = 0.f, b = 0.f;
  for (int g = 0; g < groups; ++g) { b += sums[(g * 2 + 0) * c + ch]; a += sums[(g * 2 + 1) * c + ch]; }
  if (dgamma) dgamma[ch] = (accumulate ? dgamma[ch] : 0.f) + a;
  if (dbeta) dbeta[ch] = (accumulate ? dbeta[ch] : 0.f) + b;
}

// ---------------------------------------------------------------------------------------------
// Per-channel gate (GatedSkipConnection, vae-gan-oldv.py:226-231): y = x * s[c] written into a (slice of a) buffer;
// backward in ONE pass: dx = dy * s[c] and ds[c] += sum_rows dy * x.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(kNT) scale_fwd_kernel(const T* __restrict__ x, int x_ld, const float* __restrict__ s,
                                                        T* __restrict__ y, int y_ld, int y_coff, long long rows, int c) {
  const int cv = c / 8;
  const long long total = rows * cv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long r = i / cv;
    const int ch = static_cast<int>(i - r * cv) * 8;
    float f[8];
    Raw8<T>::load(x + r * x_ld + ch).unpack(f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= s[ch + k];
    store8(y + r * y_ld + y_coff + ch, f);
  }
}
template <typename T>
__global__ void __launch_bounds__(kNT) scale_bwd_kernel(const T* __restrict__ x, int x_ld, const T* __restrict__ dy, int dy_ld,
                                                        int dy_coff, const float* __restrict__ sc, T* __restrict__ dx, int dx_ld,
                                                        long long rows, int c, float* __restrict__ ds) {
  const RowMap m = row_map(c);
  const int tid = threadIdx.x;
  const int rl = tid / m.cvl, cvi = tid % m.cvl;
  __shared__ float red[kNT][17];
  for (int cv0 = 0; cv0 < m.cv; cv0 += m.cvl) {
    const int cvec = cv0 + cvi;
    float s[8], q[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (rl < m.rows_par && cvec < m.cv) {
      const int ch = cvec * 8;
      float g[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) g[i] = sc[ch + i];
      const long long stride = static_cast<long long>(gridDim.x) * m.rows_par;
      for (long long r = static_cast<long long>(blockIdx.x) * m.rows_par + rl; r < rows; r += stride) {
        float f[8], d[8], o[8];
        Raw8<T>::load(x + r * x_ld + ch).unpack(f);
        Raw8<T>::load(dy + r * dy_ld + dy_coff + ch).unpack(d);
#pragma unroll
        for (int i = 0; i < 8; ++i) { o[i] = d[i] * g[i]; s[i] = fmaf(d[i], f[i], s[i]); }
        store8(dx + r * dx_ld + ch, o);
      }
    }
    reduce_rows_atomic(red, s, q, m, rl, cvi, cvec, ds, ds + c);     // ds[c..2c) receives the (unused) zeros of q
  }
}

// grid.x for a row-walking kernel: enough blocks to cover the rows, at most `per_sm` blocks per SM (per group)
static int row_grid(long long rows, int rows_par, int groups, int per_sm, int rows_per_thread) {
  long long blocks = (rows + static_cast<long long>(rows_par) * rows_per_thread - 1) /
                     (static_cast<long long>(rows_par) * rows_per_thread);
  long long cap = static_cast<long long>(num_sms()) * per_sm / (groups > 0 ? groups : 1);
  if (cap < 1) cap = 1;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace vg

using namespace vg;

template <typename T>
static int norm_stats_impl(const void* x, int x_ld, int x_coff, int groups, long long rows_per_group, int c, float* sums,
                           cudaStream_t st, int cls_w = 0, int cls_h = 0, float wmid = 1.f) {
  const RowMap m = row_map(c);
  const int gx = row_grid(rows_per_group, m.rows_par, groups, 4, 8);
  stats_kernel<T><<<dim3(gx, groups), kNT, 0, st>>>(static_cast<const T*>(x), x_ld, x_coff, c, rows_per_group, sums, cls_w,
                                                    cls_h, wmid);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_norm_stats_rows(const void* x, int x_ld, int x_coff, int n, int h, int w, int c, int virt_h,
                                  float* sums, int dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1, "vg_norm_stats_rows: channels must be multiples of 8");
  VG_CHECK(dtype == 0 || dtype == 1, -1, "vg_norm_stats_rows: dtype must be 0 (bf16) or 1 (fp32)");
  VG_CHECK(h >= 3 && virt_h >= h, -1, "vg_norm_stats_rows: need h >= 3 rows standing for virt_h >= h rows");
  VG_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * c, st));
  const float wmid = static_cast<float>(virt_h - 2) / static_cast<float>(h - 2);
  const long long rows = static_cast<long long>(n) * h * w;
  return dtype == 0 ? norm_stats_impl<__nv_bfloat16>(x, x_ld, x_coff, 1, rows, c, sums, st, w, h, wmid)
                    : norm_stats_impl<float>(x, x_ld, x_coff, 1, rows, c, sums, st, w, h, wmid);
}

extern "C" int vg_norm_stats(const void* x, int x_ld, int x_coff, int groups, long long rows_per_group, int c,
                             float* sums, int dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1, "vg_norm_stats: channels must be multiples of 8");
  VG_CHECK(dtype == 0 || dtype == 1, -1, "vg_norm_stats: dtype must be 0 (bf16) or 1 (fp32)");
  VG_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * groups * c, st));
  return dtype == 0 ? norm_stats_impl<__nv_bfloat16>(x, x_ld, x_coff, groups, rows_per_group, c, sums, st)
                    : norm_stats_impl<float>(x, x_ld, x_coff, groups, rows_per_group, c, sums, st);
}

extern "C" int vg_norm_finalize(const float* sums, int groups, long long rows_per_group, int c, float eps,
                                float* mean_rstd, float momentum, float* running_mean, float* running_var,
                                long long* num_batches_tracked, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  finalize_kernel<<<cdiv(groups * c, 128), 128, 0, st>>>(sums, groups, c, rows_per_group, eps, mean_rstd, momentum,
                                                         running_mean, running_var, num_batches_tracked);
  VG_LAUNCH_OK();
  return 0;
}

template <typename T>
static int norm_apply_impl(const VgNormApply* d, cudaStream_t st) {
  ApplyParams<T> p;
  p.x = static_cast<const T*>(d->x); p.x_ld = d->x_ld; p.x_coff = d->x_coff;
  p.n = d->n; p.h = d->h; p.w = d->w; p.c = d->c;
  p.mean_rstd = d->mean_rstd; p.per_sample = d->per_sample; p.gamma = d->gamma; p.beta = d->beta; p.act = d->act;
  p.y = static_cast<T*>(d->y); p.y_ld = d->y_ld; p.y_coff = d->y_coff;
  p.pool = static_cast<T*>(d->pool); p.p_ld = d->p_ld; p.p_coff = d->p_coff;
  const int groups = d->per_sample ? d->n : 1;
  const RowMap m = row_map(d->c);
  const long long cells = static_cast<long long>(d->per_sample ? 1 : d->n) * (d->pool ? d->h / 2 : d->h) *
                          (d->pool ? d->w / 2 : d->w);
  if (d->pool) {
    apply_kernel<T, true><<<dim3(row_grid(cells, m.rows_par, groups, 8, 2), groups), kNT, 0, st>>>(p);
  } else {
    apply_kernel<T, false><<<dim3(row_grid(cells, m.rows_par, groups, 8, 8), groups), kNT, 0, st>>>(p);
  }
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_norm_apply(const VgNormApply* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(d->c % 8 == 0 && d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->y_ld % 8 == 0 && d->y_coff % 8 == 0, -1,
           "vg_norm_apply: channels / strides must be multiples of 8");
  VG_CHECK(d->pool == nullptr || (d->h % 2 == 0 && d->w % 2 == 0 && d->p_ld % 8 == 0 && d->p_coff % 8 == 0), -1,
           "vg_norm_apply: pooled output needs even H, W");
  VG_CHECK(d->dtype == 0 || d->dtype == 1, -1, "vg_norm_apply: dtype must be 0 (bf16) or 1 (fp32)");
  return d->dtype == 0 ? norm_apply_impl<__nv_bfloat16>(d, st) : norm_apply_impl<float>(d, st);
}

template <typename T>
static int norm_backward_impl(const VgNormBackward* d, cudaStream_t st) {
  BwdParams<T> p;
  p.x = static_cast<const T*>(d->x); p.x_ld = d->x_ld; p.x_coff = d->x_coff;
  p.dy = static_cast<const T*>(d->dy); p.dy_ld = d->dy_ld; p.dy_coff = d->dy_coff;
  p.dpool = static_cast<const T*>(d->dpool); p.dp_ld = d->dp_ld; p.dp_coff = d->dp_coff;
  p.n = d->n; p.h = d->h; p.w = d->w; p.c = d->c;
  p.mean_rstd = d->mean_rstd; p.per_sample = d->per_sample; p.gamma = d->gamma; p.beta = d->beta; p.act = d->act;
  p.sums = d->sums;
  p.dx = static_cast<T*>(d->dx); p.dx_ld = d->dx_ld; p.dx_coff = d->dx_coff;
  p.virt_h = d->virt_h;
  const int groups = d->per_sample ? d->n : 1;
  VG_CUDA(cudaMemsetAsync(d->sums, 0, sizeof(float) * 2 * groups * d->c, st));
  const RowMap m = row_map(d->c);
  const bool pool = d->dpool != nullptr;
  const long long cells = static_cast<long long>(d->per_sample ? 1 : d->n) * (pool ? d->h / 2 : d->h) *
                          (pool ? d->w / 2 : d->w);
  static const int red_per_sm = getenv("VG_NORM_RED_PER_SM") ? atoi(getenv("VG_NORM_RED_PER_SM")) : 2;
  static const int app_per_sm = getenv("VG_NORM_APP_PER_SM") ? atoi(getenv("VG_NORM_APP_PER_SM")) : 2;
  const dim3 g_red(row_grid(cells, m.rows_par, groups, red_per_sm, pool ? 2 : 8), groups);
  const dim3 g_app(row_grid(cells, m.rows_par, groups, app_per_sm, pool ? 2 : 8), groups);
  if (pool) {
    bwd_kernel<T, false, true><<<g_red, kNT, 0, st>>>(p);
    VG_LAUNCH_OK();
    bwd_kernel<T, true, true><<<g_app, kNT, 0, st>>>(p);
  } else {
    bwd_kernel<T, false, false><<<g_red, kNT, 0, st>>>(p);
    VG_LAUNCH_OK();
    bwd_kernel<T, true, false><<<g_app, kNT, 0, st>>>(p);
  }
  VG_LAUNCH_OK();
  if (d->dgamma != nullptr || d->dbeta != nullptr) {
    affine_grad_kernel<<<cdiv(d->c, 128), 128, 0, st>>>(d->sums, groups, d->c, d->dgamma, d->dbeta, d->accumulate);
    VG_LAUNCH_OK();
  }
  return 0;
}

extern "C" int vg_norm_backward(const VgNormBackward* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(d->c % 8 == 0 && d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->dx_ld % 8 == 0 && d->dx_coff % 8 == 0, -1,
           "vg_norm_backward: channels / strides must be multiples of 8");
  VG_CHECK(d->dy != nullptr || d->dpool != nullptr, -1, "vg_norm_backward: no incoming gradient");
  VG_CHECK(d->dpool == nullptr || (d->h % 2 == 0 && d->w % 2 == 0), -1, "vg_norm_backward: pooled grad needs even H, W");
  VG_CHECK(d->dtype == 0 || d->dtype == 1, -1, "vg_norm_backward: dtype must be 0 (bf16) or 1 (fp32)");
  VG_CHECK(d->virt_h == 0 || (d->virt_h >= d->h && d->h >= 3 && !d->per_sample && d->dpool == nullptr), -1,
           "vg_norm_backward: virt_h needs h >= 3 rows, batch statistics and no pooling");
  return d->dtype == 0 ? norm_backward_impl<__nv_bfloat16>(d, st) : norm_backward_impl<float>(d, st);
}

extern "C" int vg_channel_scale_fwd(const void* x, int x_ld, const float* scale, void* y, int y_ld, int y_coff, long long rows,
                                    int c, int dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && y_ld % 8 == 0 && y_coff % 8 == 0, -1, "vg_channel_scale_fwd: multiples of 8");
  const long long items = rows * (c / 8);
  const int grid = static_cast<int>(std::min<long long>((items + kNT - 1) / kNT, static_cast<long long>(num_sms()) * 16));
  if (dtype == 0)
    scale_fwd_kernel<__nv_bfloat16><<<std::max(grid, 1), kNT, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_ld, scale,
                                                                      static_cast<__nv_bfloat16*>(y), y_ld, y_coff, rows, c);
  else
    scale_fwd_kernel<float><<<std::max(grid, 1), kNT, 0, st>>>(static_cast<const float*>(x), x_ld, scale, static_cast<float*>(y),
                                                              y_ld, y_coff, rows, c);
  VG_LAUNCH_OK();
  return 0;
}

/* dscale: fp32 [2*c] scratch-and-result: the first c entries receive sum dy*x (zeroed here) */
extern "C" int vg_channel_scale_bwd(const void* x, int x_ld, const void* dy, int dy_ld, int dy_coff, const float* scale,
                                    void* dx, int dx_ld, long long rows, int c, float* dscale, int dtype, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && dy_ld % 8 == 0 && dy_coff % 8 == 0 && dx_ld % 8 == 0, -1,
           "vg_channel_scale_bwd: multiples of 8");
  VG_CUDA(cudaMemsetAsync(dscale, 0, sizeof(float) * 2 * c, st));
  const RowMap m = row_map(c);
  const int gx = row_grid(rows, m.rows_par, 1, 4, 8);
  if (dtype == 0)
    scale_bwd_kernel<__nv_bfloat16><<<gx, kNT, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_ld,
                                                       static_cast<const __nv_bfloat16*>(dy), dy_ld, dy_coff, scale,
                                                       static_cast<__nv_bfloat16*>(dx), dx_ld, rows, c, dscale);
  else
    scale_bwd_kernel<float><<<gx, kNT, 0, st>>>(static_cast<const float*>(x), x_ld, static_cast<const float*>(dy), dy_ld, dy_coff,
                                               scale, static_cast<float*>(dx), dx_ld, rows, c, dscale);
  VG_LAUNCH_OK();
  return 0;
}
