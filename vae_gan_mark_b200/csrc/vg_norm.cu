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
// All kernels are HBM-bound: 16-byte (8 x bf16) vector accesses, one channel-vector per thread, fp32 math,
// block-level partial sums in shared memory, one fp32 atomic per (block, channel).
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
VG_DEVICE float bf16_round(float v) { return __bfloat162float(__float2bfloat16(v)); }

// ---------------------------------------------------------------------------------------------
// statistics: sums[g][0][c] = sum x, sums[g][1][c] = sum x^2 over the rows of group g
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kNT) stats_kernel(const __nv_bfloat16* __restrict__ x, int ld, int coff, int c,
                                                    long long rows_per_group, float* __restrict__ sums) {
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
      const __nv_bfloat16* base = x + static_cast<long long>(g) * rows_per_group * ld + coff + cvec * 8;
      for (long long r = static_cast<long long>(blockIdx.x) * m.rows_par + rl; r < rows_per_group;
           r += static_cast<long long>(gridDim.x) * m.rows_par) {
        float f[8];
        load8(base + r * ld, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) { s[i] += f[i]; q[i] += f[i] * f[i]; }
      }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) { red[tid][i] = s[i]; red[tid][8 + i] = q[i]; }
    __syncthreads();
    if (rl == 0 && cvec < m.cv) {
      for (int k = 0; k < 16; ++k) {
        float a = 0.f;
        for (int r = 0; r < m.rows_par; ++r) a += red[r * m.cvl + cvi][k];
        const int ch = cvec * 8 + (k & 7);
        atomicAdd(sums + (static_cast<long long>(g) * 2 + (k >> 3)) * c + ch, a);
      }
    }
    __syncthreads();
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
// ---------------------------------------------------------------------------------------------
struct ApplyParams {
  const __nv_bfloat16* x; int x_ld, x_coff;
  int n, h, w, c;
  const float* mean_rstd; int per_sample;
  const float* gamma; const float* beta;
  int act;
  __nv_bfloat16* y; int y_ld, y_coff;
  __nv_bfloat16* pool; int p_ld, p_coff;
};

__global__ void __launch_bounds__(kNT) apply_kernel(const ApplyParams p) {
  const int cv = p.c / 8;
  const bool pooled = p.pool != nullptr;
  const int ph = pooled ? p.h / 2 : p.h, pw = pooled ? p.w / 2 : p.w;
  const long long cells = static_cast<long long>(p.n) * ph * pw;
  const long long total = cells * cv;
  for (long long idx = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; idx < total;
       idx += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int cvec = static_cast<int>(idx % cv);
    const long long cell = idx / cv;
    const int pj = static_cast<int>(cell % pw);
    const int pi = static_cast<int>((cell / pw) % ph);
    const int n = static_cast<int>(cell / (static_cast<long long>(pw) * ph));
    const int ch = cvec * 8;
    const float* mr = p.mean_rstd + static_cast<long long>(p.per_sample ? n : 0) * 2 * p.c;
    float sc[8], sh[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float g = p.gamma ? p.gamma[ch + i] : 1.f, b = p.beta ? p.beta[ch + i] : 0.f;
      const float rstd = mr[p.c + ch + i];
      sc[i] = g * rstd;
      sh[i] = b - mr[ch + i] * g * rstd;
    }
    const int reps = pooled ? 2 : 1;
    float mx[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) mx[i] = -INFINITY;
    for (int a = 0; a < reps; ++a)
      for (int b = 0; b < reps; ++b) {
        const int i_h = pooled ? 2 * pi + a : pi, i_w = pooled ? 2 * pj + b : pj;
        const long long pix = (static_cast<long long>(n) * p.h + i_h) * p.w + i_w;
        float f[8], o[8];
        load8(p.x + pix * p.x_ld + p.x_coff + ch, f);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          o[i] = act_fwd(fmaf(f[i], sc[i], sh[i]), p.act);
          mx[i] = fmaxf(mx[i], bf16_round(o[i]));
        }
        store8(p.y + pix * p.y_ld + p.y_coff + ch, o);
      }
    if (pooled) {
      const long long ppix = (static_cast<long long>(n) * ph + pi) * pw + pj;
      store8(p.pool + ppix * p.p_ld + p.p_coff + ch, mx);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward.  g = (dy + routed pooled grad) * act'(pre);  pass 1 reduces sum g and sum g*xhat,
// pass 2 writes dx = gamma*rstd*(g - mean(g) - xhat*mean(g*xhat)).
// ---------------------------------------------------------------------------------------------
struct BwdParams {
  const __nv_bfloat16* x; int x_ld, x_coff;           // raw (pre-normalisation) conv output
  const __nv_bfloat16* dy; int dy_ld, dy_coff;        // grad wrt full-resolution activation (nullable)
  const __nv_bfloat16* dpool; int dp_ld, dp_coff;     // grad wrt pooled activation (nullable)
  int n, h, w, c;
  const float* mean_rstd; int per_sample;
  const float* gamma; const float* beta;
  int act;
  float* sums;                                        // [groups][2][c]: sum g, sum g*xhat
  __nv_bfloat16* dx; int dx_ld, dx_coff;
};

// computes g[8] for the 1 or 4 pixels of a cell; returns through arrays indexed [pixel][i]
VG_DEVICE void cell_grads(const BwdParams& p, int n, int pi, int pj, int ch, const float* mr, bool pooled,
                          float (&xh)[4][8], float (&g)[4][8]) {
  float sc[8], sh[8], mean[8], rstd[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const float ga = p.gamma ? p.gamma[ch + i] : 1.f, be = p.beta ? p.beta[ch + i] : 0.f;
    mean[i] = mr[ch + i];
    rstd[i] = mr[p.c + ch + i];
    sc[i] = ga * rstd[i];
    sh[i] = be - mean[i] * ga * rstd[i];
  }
  const int reps = pooled ? 2 : 1;
  float pre[4][8], yv[4][8];
  for (int a = 0; a < reps; ++a)
    for (int b = 0; b < reps; ++b) {
      const int k = a * 2 + b;
      const int i_h = pooled ? 2 * pi + a : pi, i_w = pooled ? 2 * pj + b : pj;
      const long long pix = (static_cast<long long>(n) * p.h + i_h) * p.w + i_w;
      float f[8];
      load8(p.x + pix * p.x_ld + p.x_coff + ch, f);
      float d[8];
      if (p.dy != nullptr) load8(p.dy + pix * p.dy_ld + p.dy_coff + ch, d);
      else {
#pragma unroll
        for (int i = 0; i < 8; ++i) d[i] = 0.f;
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        xh[k][i] = (f[i] - mean[i]) * rstd[i];
        pre[k][i] = fmaf(f[i], sc[i], sh[i]);
        yv[k][i] = bf16_round(act_fwd(pre[k][i], p.act));
        g[k][i] = d[i];
      }
    }
  if (pooled && p.dpool != nullptr) {
    const int ph = p.h / 2, pw = p.w / 2;
    const long long ppix = (static_cast<long long>(n) * ph + pi) * pw + pj;
    float dp[8];
    load8(p.dpool + ppix * p.dp_ld + p.dp_coff + ch, dp);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      int best = 0;
      float bv = yv[0][i];
#pragma unroll
      for (int k = 1; k < 4; ++k)
        if (yv[k][i] > bv) { bv = yv[k][i]; best = k; }
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (k == best) g[k][i] += dp[i];
    }
  }
  const int npx = pooled ? 4 : 1;
  for (int k = 0; k < npx; ++k)
#pragma unroll
    for (int i = 0; i < 8; ++i) g[k][i] *= act_grad(pre[k][i], p.act);
}

template <bool kApply>
__global__ void __launch_bounds__(kNT) bwd_kernel(const BwdParams p) {
  const bool pooled = p.dpool != nullptr;
  const int ph = pooled ? p.h / 2 : p.h, pw = pooled ? p.w / 2 : p.w;
  const RowMap m = row_map(p.c);
  const int tid = threadIdx.x;
  const int rl = tid / m.cvl, cvi = tid % m.cvl;
  const int grp = blockIdx.y;                       // sample index when per_sample, else 0
  const int n_begin = p.per_sample ? grp : 0, n_count = p.per_sample ? 1 : p.n;
  const long long cells = static_cast<long long>(n_count) * ph * pw;
  const float inv_rows = 1.f / (static_cast<float>(n_count) * p.h * p.w);
  __shared__ float red[kApply ? 1 : kNT][17];
  for (int cv0 = 0; cv0 < m.cv; cv0 += m.cvl) {
    const int cvec = cv0 + cvi;
    const int ch = cvec * 8;
    const bool active = rl < m.rows_par && cvec < m.cv;
    const float* mr = p.mean_rstd + static_cast<long long>(grp) * 2 * p.c;
    float s[8], q[8], k1[8], k2[8], sc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) s[i] = q[i] = 0.f;
    if (kApply && active) {
      const float* sm = p.sums + static_cast<long long>(grp) * 2 * p.c;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        k1[i] = sm[ch + i] * inv_rows;
        k2[i] = sm[p.c + ch + i] * inv_rows;
        sc[i] = (p.gamma ? p.gamma[ch + i] : 1.f) * mr[p.c + ch + i];
      }
    }
    if (active) {
      for (long long cell = static_cast<long long>(blockIdx.x) * m.rows_par + rl; cell < cells;
           cell += static_cast<long long>(gridDim.x) * m.rows_par) {
        const int pj = static_cast<int>(cell % pw);
        const int pi = static_cast<int>((cell / pw) % ph);
        const int n = n_begin + static_cast<int>(cell / (static_cast<long long>(pw) * ph));
        float xh[4][8], g[4][8];
        cell_grads(p, n, pi, pj, ch, mr, pooled, xh, g);
        const int npx = pooled ? 4 : 1;
        for (int k = 0; k < npx; ++k) {
          if (kApply) {
            const int i_h = pooled ? 2 * pi + (k >> 1) : pi, i_w = pooled ? 2 * pj + (k & 1) : pj;
            const long long pix = (static_cast<long long>(n) * p.h + i_h) * p.w + i_w;
            float o[8];
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] = sc[i] * (g[k][i] - k1[i] - xh[k][i] * k2[i]);
            store8(p.dx + pix * p.dx_ld + p.dx_coff + ch, o);
          } else {
#pragma unroll
            for (int i = 0; i < 8; ++i) { s[i] += g[k][i]; q[i] += g[k][i] * xh[k][i]; }
          }
        }
      }
    }
    if (!kApply) {
#pragma unroll
      for (int i = 0; i < 8; ++i) { red[tid][i] = s[i]; red[tid][8 + i] = q[i]; }
      __syncthreads();
      if (rl == 0 && cvec < m.cv) {
        for (int k = 0; k < 16; ++k) {
          float a = 0.f;
          for (int r = 0; r < m.rows_par; ++r) a += red[r * m.cvl + cvi][k];
          atomicAdd(p.sums + (static_cast<long long>(grp) * 2 + (k >> 3)) * p.c + cvec * 8 + (k & 7), a);
        }
      }
      __syncthreads();
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

static int grid_for(long long work_items, int per_block) {
  long long blocks = (work_items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return static_cast<int>(blocks);
}

}  // namespace vg

using namespace vg;

extern "C" int vg_norm_stats(const void* x, int x_ld, int x_coff, int groups, long long rows_per_group, int c,
                             float* sums, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(c % 8 == 0 && x_ld % 8 == 0 && x_coff % 8 == 0, -1, "vg_norm_stats: channels must be multiples of 8");
  VG_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 2 * groups * c, st));
  const RowMap m = row_map(c);
  int gx = grid_for(rows_per_group, m.rows_par * 8);
  if (groups > 1) gx = max(1, min(gx, (num_sms() * 8) / groups));
  stats_kernel<<<dim3(gx, groups), kNT, 0, st>>>(static_cast<const __nv_bfloat16*>(x), x_ld, x_coff, c, rows_per_group,
                                                 sums);
  VG_LAUNCH_OK();
  return 0;
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

extern "C" int vg_norm_apply(const VgNormApply* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(d->c % 8 == 0 && d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->y_ld % 8 == 0 && d->y_coff % 8 == 0, -1,
           "vg_norm_apply: channels / strides must be multiples of 8");
  VG_CHECK(d->pool == nullptr || (d->h % 2 == 0 && d->w % 2 == 0 && d->p_ld % 8 == 0 && d->p_coff % 8 == 0), -1,
           "vg_norm_apply: pooled output needs even H, W");
  ApplyParams p;
  p.x = static_cast<const __nv_bfloat16*>(d->x); p.x_ld = d->x_ld; p.x_coff = d->x_coff;
  p.n = d->n; p.h = d->h; p.w = d->w; p.c = d->c;
  p.mean_rstd = d->mean_rstd; p.per_sample = d->per_sample; p.gamma = d->gamma; p.beta = d->beta; p.act = d->act;
  p.y = static_cast<__nv_bfloat16*>(d->y); p.y_ld = d->y_ld; p.y_coff = d->y_coff;
  p.pool = static_cast<__nv_bfloat16*>(d->pool); p.p_ld = d->p_ld; p.p_coff = d->p_coff;
  const long long cells = static_cast<long long>(d->n) * (d->pool ? d->h / 2 : d->h) * (d->pool ? d->w / 2 : d->w);
  apply_kernel<<<grid_for(cells * (d->c / 8), kNT), kNT, 0, st>>>(p);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_norm_backward(const VgNormBackward* d, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(d->c % 8 == 0 && d->x_ld % 8 == 0 && d->x_coff % 8 == 0 && d->dx_ld % 8 == 0 && d->dx_coff % 8 == 0, -1,
           "vg_norm_backward: channels / strides must be multiples of 8");
  VG_CHECK(d->dy != nullptr || d->dpool != nullptr, -1, "vg_norm_backward: no incoming gradient");
  VG_CHECK(d->dpool == nullptr || (d->h % 2 == 0 && d->w % 2 == 0), -1, "vg_norm_backward: pooled grad needs even H, W");
  BwdParams p;
  p.x = static_cast<const __nv_bfloat16*>(d->x); p.x_ld = d->x_ld; p.x_coff = d->x_coff;
  p.dy = static_cast<const __nv_bfloat16*>(d->dy); p.dy_ld = d->dy_ld; p.dy_coff = d->dy_coff;
  p.dpool = static_cast<const __nv_bfloat16*>(d->dpool); p.dp_ld = d->dp_ld; p.dp_coff = d->dp_coff;
  p.n = d->n; p.h = d->h; p.w = d->w; p.c = d->c;
  p.mean_rstd = d->mean_rstd; p.per_sample = d->per_sample; p.gamma = d->gamma; p.beta = d->beta; p.act = d->act;
  p.sums = d->sums;
  p.dx = static_cast<__nv_bfloat16*>(d->dx); p.dx_ld = d->dx_ld; p.dx_coff = d->dx_coff;
  const int groups = d->per_sample ? d->n : 1;
  VG_CUDA(cudaMemsetAsync(d->sums, 0, sizeof(float) * 2 * groups * d->c, st));
  const RowMap m = row_map(d->c);
  const long long cells = static_cast<long long>(d->per_sample ? 1 : d->n) * (d->dpool ? d->h / 2 : d->h) *
                          (d->dpool ? d->w / 2 : d->w);
  int gx = grid_for(cells, m.rows_par * 4);
  if (groups > 1) gx = max(1, min(gx, (num_sms() * 8) / groups));
  bwd_kernel<false><<<dim3(gx, groups), kNT, 0, st>>>(p);
  VG_LAUNCH_OK();
  bwd_kernel<true><<<dim3(gx, groups), kNT, 0, st>>>(p);
  VG_LAUNCH_OK();
  if (d->dgamma != nullptr || d->dbeta != nullptr) {
    affine_grad_kernel<<<cdiv(d->c, 128), 128, 0, st>>>(d->sums, groups, d->c, d->dgamma, d->dbeta, d->accumulate);
    VG_LAUNCH_OK();
  }
  return 0;
}
