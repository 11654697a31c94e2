// Loss, reparameterisation, spectral-norm and optimiser kernels (all HBM- or latency-bound, fp32).
//   * reparameterise z = mu + eps*exp(0.5*logvar) and the KL term, forward and backward fused
//     (vae-gan.py:133-136, :420);
//   * sigmoid output + L1 reconstruction loss with its gradient (vae-gan.py:82, :419);
//   * hinge losses of the discriminator / generator with their gradients (vae-gan.py:313-320);
//   * spectral normalisation: one power iteration, sigma, and the gradient through W/sigma
//     (torch.nn.utils.spectral_norm as applied at vae-gan.py:153-156);
//   * clip_grad_norm_ + Adam fused into one pass over flat parameter/gradient/moment buffers
//     (vae-gan.py:424, :541-542).
#include <algorithm>
#include <stdlib.h>

#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

static int lo_grid(long long items, int per_block = 256) {
  long long b = (items + per_block - 1) / per_block;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  return static_cast<int>(b < 1 ? 1 : (b > cap ? cap : b));
}

VG_DEVICE float block_sum(float v) {   // blockDim.x == 256
  __shared__ float sm[8];
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  float r = 0.f;
  if (threadIdx.x < 8) r = sm[threadIdx.x];
  if (threadIdx.x < 32) r = warp_sum(r);
  __syncthreads();
  return r;   // valid in thread 0 (and warp 0)
}

// ---------------------------------------------------------------------------------------------
// heads -> mu, logvar, z, KL
// ---------------------------------------------------------------------------------------------
// heads: fp32 [B][2z] (mu | logvar, bias not yet added); eps fp32 [B][z]
__global__ void reparam_fwd_kernel(const float* __restrict__ heads, const float* __restrict__ bias_mu,
                                   const float* __restrict__ bias_lv, const float* __restrict__ eps, int b, int z,
                                   float* __restrict__ mu, float* __restrict__ lv, float* __restrict__ zout,
                                   float* __restrict__ kl_out) {
  float acc = 0.f;
  const int total = b * z;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / z, c = i % z;
    const float m = heads[r * 2 * z + c] + bias_mu[c];
    const float l = heads[r * 2 * z + z + c] + bias_lv[c];
    mu[i] = m;
    lv[i] = l;
    zout[i] = m + eps[i] * expf(0.5f * l);
    acc += 1.f + l - m * m - expf(l);
  }
  const float s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(kl_out, -0.5f * s / static_cast<float>(total));
}
// d_heads[B][2z] = [dmu_ext + dz + klw*mu/(Bz) | dlv_ext + dz*eps*0.5*std + klw*(-0.5)(1-exp(lv))/(Bz)]
__global__ void reparam_bwd_kernel(const float* __restrict__ mu, const float* __restrict__ lv,
                                   const float* __restrict__ eps, const float* __restrict__ dz,
                                   const float* __restrict__ dmu_ext, const float* __restrict__ dlv_ext,
                                   const float* __restrict__ dkl, int b, int z, float* __restrict__ dheads,
                                   __nv_bfloat16* __restrict__ dheads_bf16, int bf_ld) {
  const int total = b * z;
  const float kw = dkl ? *dkl / static_cast<float>(total) : 0.f;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int r = i / z, c = i % z;
    const float g = dz ? dz[i] : 0.f;
    const float el = expf(lv[i]);
    const float dm = (dmu_ext ? dmu_ext[i] : 0.f) + g + kw * mu[i];
    const float dl = (dlv_ext ? dlv_ext[i] : 0.f) + g * eps[i] * 0.5f * sqrtf(el) - 0.5f * kw * (1.f - el);
    dheads[r * 2 * z + c] = dm;
    dheads[r * 2 * z + z + c] = dl;
    if (dheads_bf16) {
      dheads_bf16[static_cast<long long>(r) * bf_ld + c] = __float2bfloat16(dm);
      dheads_bf16[static_cast<long long>(r) * bf_ld + z + c] = __float2bfloat16(dl);
    }
  }
}

// ---------------------------------------------------------------------------------------------
// sigmoid + L1
// ---------------------------------------------------------------------------------------------
// pre: fp32 [N][H][W][C] (NHWC, C small); y: fp32 NCHW
__global__ void sigmoid_nhwc_to_nchw_kernel(const float* __restrict__ pre, int n, int c, int hw, float* __restrict__ y) {
  const long long total = static_cast<long long>(n) * c * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % hw);
    const int ch = static_cast<int>((i / hw) % c);
    const long long b = i / (static_cast<long long>(hw) * c);
    y[i] = 1.f / (1.f + expf(-pre[(b * hw + p) * c + ch]));
  }
}
// dpre (NHWC fp32) = dy (NCHW fp32) * y (1 - y)
__global__ void sigmoid_bwd_kernel(const float* __restrict__ y, const float* __restrict__ dy, int n, int c, int hw,
                                   float* __restrict__ dpre) {
  const long long total = static_cast<long long>(n) * c * hw;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int p = static_cast<int>(i % hw);
    const int ch = static_cast<int>((i / hw) % c);
    const long long b = i / (static_cast<long long>(hw) * c);
    const float v = y[i];
    dpre[(b * hw + p) * c + ch] = dy[i] * v * (1.f - v);
  }
}
__global__ void l1_fwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* out) {
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc += fabsf(a[i] - b[i]);
  const float s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, s / static_cast<float>(n));
}
// da (+)= gout * sign(a - b) / n
__global__ void l1_bwd_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                              const float* __restrict__ gout, float* __restrict__ da, int accumulate) {
  const float g = *gout / static_cast<float>(n);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float d = a[i] - b[i];
    const float v = d > 0.f ? g : (d < 0.f ? -g : 0.f);
    da[i] = accumulate ? da[i] + v : v;
  }
}

// ---------------------------------------------------------------------------------------------
// hinge: mode 1: mean(relu(1-p)); mode 0: mean(relu(1+p)); mode 2: -mean(p)
// ---------------------------------------------------------------------------------------------
__global__ void hinge_fwd_kernel(const float* __restrict__ p, long long n, int mode, float* out) {
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = p[i];
    acc += mode == 1 ? fmaxf(1.f - v, 0.f) : (mode == 0 ? fmaxf(1.f + v, 0.f) : -v);
  }
  const float s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, s / static_cast<float>(n));
}
__global__ void hinge_bwd_kernel(const float* __restrict__ p, long long n, int mode, const float* __restrict__ gout,
                                 float* __restrict__ dp) {
  const float g = *gout / static_cast<float>(n);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float v = p[i];
    dp[i] = mode == 1 ? (1.f - v > 0.f ? -g : 0.f) : (mode == 0 ? (1.f + v > 0.f ? g : 0.f) : -g);
  }
}

// ---------------------------------------------------------------------------------------------
// spectral norm.  W: fp32 [rows][cols] (OIHW flattened), u [rows], v [cols]
// ---------------------------------------------------------------------------------------------
// t[j] += sum_{i in this block's row chunk} W[i][j] u[i]   (t zeroed by the caller; blockIdx.y = row chunk).
// One thread per column over ALL rows was a chain of `rows` dependent loads on 16 blocks (27 us for 512 x 4096);
// splitting the rows over blockIdx.y fills the machine.  |t|^2 is taken by the next kernel.
__global__ void sn_wtu_kernel(const float* __restrict__ w, const float* __restrict__ u, int rows, int cols,
                              float* __restrict__ t, int rows_per_block) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  const int i0 = blockIdx.y * rows_per_block, i1 = min(rows, i0 + rows_per_block);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
  int i = i0;
  for (; i + 3 < i1; i += 4) {
    a0 = fmaf(w[static_cast<long long>(i) * cols + j], u[i], a0);
    a1 = fmaf(w[static_cast<long long>(i + 1) * cols + j], u[i + 1], a1);
    a2 = fmaf(w[static_cast<long long>(i + 2) * cols + j], u[i + 2], a2);
    a3 = fmaf(w[static_cast<long long>(i + 3) * cols + j], u[i + 3], a3);
  }
  for (; i < i1; ++i) a0 = fmaf(w[static_cast<long long>(i) * cols + j], u[i], a0);
  atomicAdd(t + j, (a0 + a1) + (a2 + a3));
}
// s[i] = sum_j W[i][j] * t[j] * tscale, tscale = 1/max(|t|, eps) (or 1 when nrm_t == nullptr: t is already a unit vector);
// every block takes |t|^2 itself from the t it reads anyway (block 0 publishes it in *nrm_t); nrm_s += sum s^2
__global__ void sn_wv_kernel(const float* __restrict__ w, const float* __restrict__ t, float* __restrict__ nrm_t,
                             float eps, int rows, int cols, float* __restrict__ s, float* __restrict__ nrm_s) {
  const int i = blockIdx.x;
  float acc = 0.f, tt = 0.f;
  for (int j = threadIdx.x; j < cols; j += blockDim.x) {
    const float tv = t[j];
    acc = fmaf(w[static_cast<long long>(i) * cols + j], tv, acc);
    tt = fmaf(tv, tv, tt);
  }
  float tscale = 1.f;
  if (nrm_t != nullptr) {
    const float t2 = block_sum(tt);          // identical in every block (same data, same order)
    tscale = 1.f / fmaxf(sqrtf(t2), eps);
    if (i == 0 && threadIdx.x == 0) *nrm_t = t2;
  }
  const float r = block_sum(acc) * tscale;
  if (threadIdx.x == 0) {
    s[i] = r;
    atomicAdd(nrm_s, r * r);
  }
}
// training: v = t/max(|t|,eps), u = s/max(|s|,eps), sigma = u . s.   eval: sigma = u_old . s (u, v untouched)
__global__ void sn_finish_kernel(const float* __restrict__ t, const float* __restrict__ s, const float* __restrict__ nrm,
                                 float eps, int rows, int cols, int training, float* __restrict__ u,
                                 float* __restrict__ v, float* __restrict__ sigma) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (training) {
    const float tn = fmaxf(sqrtf(nrm[0]), eps), sn = fmaxf(sqrtf(nrm[1]), eps);
    if (i < cols) v[i] = t[i] / tn;
    if (i < rows) u[i] = s[i] / sn;
    if (i == 0) *sigma = nrm[1] / sn;
  } else {
    float acc = 0.f;
    if (blockIdx.x == 0) {
      for (int k = threadIdx.x; k < rows; k += blockDim.x) acc += u[k] * s[k];
      const float r = block_sum(acc);
      if (threadIdx.x == 0) *sigma = r;
    }
  }
}
// dot += sum G .* W
__global__ void dot_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n, float* out) {
  float acc = 0.f;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc = fmaf(a[i], b[i], acc);
  const float s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
// dW_orig (+)= G/sigma - (dot/sigma^2) u v^T      (dot = <G, W_orig>)
__global__ void sn_bwd_kernel(const float* __restrict__ g, const float* __restrict__ u, const float* __restrict__ v,
                              const float* __restrict__ sigma, const float* __restrict__ dot, int rows, int cols,
                              float* __restrict__ dw, int accumulate) {
  const long long n = static_cast<long long>(rows) * cols;
  const float inv = 1.f / *sigma;
  const float k = *dot * inv * inv;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int r = static_cast<int>(i / cols), c = static_cast<int>(i % cols);
    const float val = g[i] * inv - k * u[r] * v[c];
    dw[i] = accumulate ? dw[i] + val : val;
  }
}

// ---------------------------------------------------------------------------------------------
// clip_grad_norm_ + Adam
// ---------------------------------------------------------------------------------------------
__global__ void sumsq_kernel(const float* __restrict__ g, long long n, float* out) {
  float acc = 0.f;
  const long long n4 = n / 4;
  const float4* g4 = reinterpret_cast<const float4*>(g);
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const float4 v = g4[i];
    acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
  }
  for (long long i = n4 * 4 + static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x)
    acc += g[i] * g[i];
  const float s = block_sum(acc);
  if (threadIdx.x == 0) atomicAdd(out, s);
}
__device__ __forceinline__ void adam_update(float& p, float& g, float& m, float& v, float clip, float step, float b1, float b2,
                                            float eps, float bc2_sqrt) {
  g *= clip;
  m = b1 * m + (1.f - b1) * g;
  v = b2 * v + (1.f - b2) * g * g;
  p -= step * m / (sqrtf(v) / bc2_sqrt + eps);
}
__global__ void adam_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                            long long n, float lr, float b1, float b2, float eps, float bc1, float bc2_sqrt,
                            const float* __restrict__ gnorm_sq, float max_norm, int write_back_grad) {
  float clip = 1.f;
  if (gnorm_sq != nullptr && max_norm > 0.f) clip = fminf(1.f, max_norm / (sqrtf(*gnorm_sq) + 1e-6f));
  const float step = lr / bc1;
  const long long tid = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  long long done = 0;
  // 16-byte accesses (the scalar loop below ran at 0.46 of the copy bandwidth: profiles/r02_ncu_hbm_bound_kernels.txt)
  if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
        reinterpret_cast<uintptr_t>(v)) & 15) == 0) {
    const long long n4 = n / 4;
    float4* p4 = reinterpret_cast<float4*>(p); float4* g4 = reinterpret_cast<float4*>(g);
    float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
    for (long long i = tid; i < n4; i += nth) {
      float4 pv = p4[i], gv = g4[i], mv = m4[i], vv = v4[i];
      adam_update(pv.x, gv.x, mv.x, vv.x, clip, step, b1, b2, eps, bc2_sqrt);
      adam_update(pv.y, gv.y, mv.y, vv.y, clip, step, b1, b2, eps, bc2_sqrt);
      adam_update(pv.z, gv.z, mv.z, vv.z, clip, step, b1, b2, eps, bc2_sqrt);
      adam_update(pv.w, gv.w, mv.w, vv.w, clip, step, b1, b2, eps, bc2_sqrt);
      p4[i] = pv; m4[i] = mv; v4[i] = vv;
      if (write_back_grad) g4[i] = gv;
    }
    done = n4 * 4;
  }
  for (long long i = done + tid; i < n; i += nth) {
    float pi = p[i], gi = g[i], mi = m[i], vi = v[i];
    adam_update(pi, gi, mi, vi, clip, step, b1, b2, eps, bc2_sqrt);
    p[i] = pi; m[i] = mi; v[i] = vi;
    if (write_back_grad) g[i] = gi;
  }
}

// ---- multi-tensor variants: one launch for a whole parameter list (table of VgAdamTensor on the device) ----
// state[0] = step count, state[1] = 1 - beta1^step, state[2] = sqrt(1 - beta2^step); advanced on the device so
// that a captured CUDA graph replays correctly.
__global__ void adam_prepare_kernel(float* state, float b1, float b2) {
  const double step = static_cast<double>(state[0]) + 1.0;
  state[0] = static_cast<float>(step);
  state[1] = static_cast<float>(1.0 - pow(static_cast<double>(b1), step));
  state[2] = static_cast<float>(sqrt(1.0 - pow(static_cast<double>(b2), step)));
}
// deterministic: per-block partials (fixed grid) + a single-block ordered final sum, so that data-parallel replicas
// holding identical gradients compute bit-identical norms (and therefore bit-identical clipped updates)
__global__ void final_sum_kernel(const float* __restrict__ partials, int n, float* out) {
  __shared__ double sm[256];
  double a = 0.0;
  for (int i = threadIdx.x; i < n; i += 256) a += static_cast<double>(partials[i]);
  sm[threadIdx.x] = a;
  __syncthreads();
  for (int s = 128; s > 0; s >>= 1) {
    if (threadIdx.x < s) sm[threadIdx.x] += sm[threadIdx.x + s];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = static_cast<float>(sm[0]);
}
// Most tensors of a parameter list are tiny (72 of the 109 tensors of the v2 generator hold < 64 K elements): with a fixed
// thread -> element map block 0 would walk the head of every one of them, one DRAM round trip after the other, while the
// other blocks idle through the list.  Tensor t is therefore started at block (t * 37) mod gridDim.x: the heads are spread
// over the grid (a fixed map: the per-block partial sums below stay deterministic and identical on every replica).
__device__ __forceinline__ long long rotated_tid(int t, int rotate) {
  const unsigned b = rotate ? (blockIdx.x + gridDim.x - (static_cast<unsigned>(t) * 37u) % gridDim.x) % gridDim.x : blockIdx.x;
  return static_cast<long long>(b) * blockDim.x + threadIdx.x;
}
__global__ void multi_sumsq_kernel(const VgAdamTensor* __restrict__ tab, int count, float* out, int rotate) {
  float acc = 0.f;
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int t = 0; t < count; ++t) {
    const float* g = tab[t].g;
    const long long n = tab[t].n;
    if (g == nullptr) continue;
    const long long tid = rotated_tid(t, rotate);
    if ((reinterpret_cast<uintptr_t>(g) & 15) == 0) {
      const long long n4 = n / 4;
      const float4* g4 = reinterpret_cast<const float4*>(g);
      for (long long i = tid; i < n4; i += nth) {
        const float4 v = g4[i];
        acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
      }
      for (long long i = n4 * 4 + tid; i < n; i += nth) acc += g[i] * g[i];
    } else {
      for (long long i = tid; i < n; i += nth) acc += g[i] * g[i];
    }
  }
  const float s = block_sum(acc);
  if (threadIdx.x == 0) out[blockIdx.x] = s;      // out = per-block partials
}
__global__ void multi_adam_kernel(const VgAdamTensor* __restrict__ tab, int count, float lr, float b1, float b2,
                                  float eps, const float* __restrict__ state, const float* __restrict__ gnorm_sq,
                                  float max_norm, int write_back_grad, int rotate) {
  float clip = 1.f;
  if (gnorm_sq != nullptr && max_norm > 0.f) clip = fminf(1.f, max_norm / (sqrtf(*gnorm_sq) + 1e-6f));
  const float step = (lr < 0.f ? state[3] : lr) / state[1];      // lr < 0: the rate lives in the device state block
  const float bc2_sqrt = state[2];
  const long long nth = static_cast<long long>(gridDim.x) * blockDim.x;
  for (int t = 0; t < count; ++t) {
    float* p = tab[t].p; float* g = tab[t].g; float* m = tab[t].m; float* v = tab[t].v;
    __nv_bfloat16* sh = static_cast<__nv_bfloat16*>(tab[t].shadow);
    const long long n = tab[t].n;
    if (g == nullptr) continue;
    const long long tid = rotated_tid(t, rotate);
    long long done = 0;
    if (((reinterpret_cast<uintptr_t>(p) | reinterpret_cast<uintptr_t>(g) | reinterpret_cast<uintptr_t>(m) |
          reinterpret_cast<uintptr_t>(v)) & 15) == 0 && (reinterpret_cast<uintptr_t>(sh) & 7) == 0) {
      const long long n4 = n / 4;
      float4* p4 = reinterpret_cast<float4*>(p); float4* g4 = reinterpret_cast<float4*>(g);
      float4* m4 = reinterpret_cast<float4*>(m); float4* v4 = reinterpret_cast<float4*>(v);
      for (long long i = tid; i < n4; i += nth) {
        float4 pv = p4[i], gv = g4[i], mv = m4[i], vv = v4[i];
        float* pp = &pv.x; float* gg = &gv.x; float* mm = &mv.x; float* vq = &vv.x;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float gi = gg[k] * clip;
          mm[k] = b1 * mm[k] + (1.f - b1) * gi;
          vq[k] = b2 * vq[k] + (1.f - b2) * gi * gi;
          pp[k] -= step * mm[k] / (sqrtf(vq[k]) / bc2_sqrt + eps);
          gg[k] = gi;
        }
        p4[i] = pv; m4[i] = mv; v4[i] = vv;
        if (write_back_grad) g4[i] = gv;
        if (sh != nullptr)      // bf16 copy of the updated weights = the tensor core's operand of the next step
          reinterpret_cast<uint2*>(sh)[i] = make_uint2(pack_bf16x2(pv.x, pv.y), pack_bf16x2(pv.z, pv.w));
      }
      done = n4 * 4;
    }
    for (long long i = done + tid; i < n; i += nth) {
      const float gi = g[i] * clip;
      const float mi = b1 * m[i] + (1.f - b1) * gi;
      const float vi = b2 * v[i] + (1.f - b2) * gi * gi;
      m[i] = mi;
      v[i] = vi;
      const float pn = p[i] - step * mi / (sqrtf(vi) / bc2_sqrt + eps);
      p[i] = pn;
      if (sh != nullptr) sh[i] = __float2bfloat16(pn);
      if (write_back_grad) g[i] = gi;
    }
  }
}

}  // namespace vg

using namespace vg;
#define ST static_cast<cudaStream_t>(stream_)
// development knob: VG_ADAM_ROTATE=0 restores the fixed thread -> element map of the multi-tensor kernels (rotated_tid)
static int adam_rotate() {
  static const int on = getenv("VG_ADAM_ROTATE") ? atoi(getenv("VG_ADAM_ROTATE")) : 1;
  return on;
}

extern "C" int vg_reparam_kl_fwd(const float* heads, const float* bias_mu, const float* bias_lv, const float* eps, int b,
                                 int z, float* mu, float* logvar, float* zout, float* kl_out, void* stream_) {
  VG_CUDA(cudaMemsetAsync(kl_out, 0, sizeof(float), ST));
  reparam_fwd_kernel<<<lo_grid(static_cast<long long>(b) * z), 256, 0, ST>>>(heads, bias_mu, bias_lv, eps, b, z, mu, logvar,
                                                                             zout, kl_out);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_reparam_kl_bwd(const float* mu, const float* logvar, const float* eps, const float* dz,
                                 const float* dmu_ext, const float* dlv_ext, const float* dkl, int b, int z,
                                 float* dheads, void* dheads_bf16, int bf_ld, void* stream_) {
  reparam_bwd_kernel<<<lo_grid(static_cast<long long>(b) * z), 256, 0, ST>>>(
      mu, logvar, eps, dz, dmu_ext, dlv_ext, dkl, b, z, dheads, static_cast<__nv_bfloat16*>(dheads_bf16), bf_ld);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_sigmoid_fwd(const float* pre_nhwc, int n, int c, int hw, float* y_nchw, void* stream_) {
  sigmoid_nhwc_to_nchw_kernel<<<lo_grid(static_cast<long long>(n) * c * hw), 256, 0, ST>>>(pre_nhwc, n, c, hw, y_nchw);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_sigmoid_bwd(const float* y_nchw, const float* dy_nchw, int n, int c, int hw, float* dpre_nhwc,
                              void* stream_) {
  sigmoid_bwd_kernel<<<lo_grid(static_cast<long long>(n) * c * hw), 256, 0, ST>>>(y_nchw, dy_nchw, n, c, hw, dpre_nhwc);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_l1_fwd(const float* a, const float* b, long long n, float* out, void* stream_) {
  VG_CUDA(cudaMemsetAsync(out, 0, sizeof(float), ST));
  l1_fwd_kernel<<<lo_grid(n, 1024), 256, 0, ST>>>(a, b, n, out);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_l1_bwd(const float* a, const float* b, long long n, const float* gout, float* da, int accumulate,
                         void* stream_) {
  l1_bwd_kernel<<<lo_grid(n), 256, 0, ST>>>(a, b, n, gout, da, accumulate);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_hinge_fwd(const float* p, long long n, int mode, float* out, void* stream_) {
  VG_CHECK(mode >= 0 && mode <= 2, -1, "vg_hinge_fwd: mode must be 0 (fake), 1 (real) or 2 (generator)");
  VG_CUDA(cudaMemsetAsync(out, 0, sizeof(float), ST));
  hinge_fwd_kernel<<<lo_grid(n, 1024), 256, 0, ST>>>(p, n, mode, out);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_hinge_bwd(const float* p, long long n, int mode, const float* gout, float* dp, void* stream_) {
  VG_CHECK(mode >= 0 && mode <= 2, -1, "vg_hinge_bwd: mode must be 0 (fake), 1 (real) or 2 (generator)");
  hinge_bwd_kernel<<<lo_grid(n), 256, 0, ST>>>(p, n, mode, gout, dp);
  VG_LAUNCH_OK();
  return 0;
}

// scratch: fp32 [rows + cols + 2]
extern "C" int vg_spectral_sigma(const float* w, int rows, int cols, float* u, float* v, int training, float eps,
                                 float* sigma, float* scratch, void* stream_) {
  float* t = scratch;          // [cols]
  float* s = scratch + cols;   // [rows]
  float* nrm = s + rows;       // [2]
  VG_CUDA(cudaMemsetAsync(nrm, 0, 2 * sizeof(float), ST));
  if (training) {
    VG_CUDA(cudaMemsetAsync(t, 0, sizeof(float) * cols, ST));
    const int rpb = std::max(8, cdiv(rows, 32));
    sn_wtu_kernel<<<dim3(cdiv(cols, 256), cdiv(rows, rpb)), 256, 0, ST>>>(w, u, rows, cols, t, rpb);
    VG_LAUNCH_OK();
    sn_wv_kernel<<<rows, 256, 0, ST>>>(w, t, nrm, eps, rows, cols, s, nrm + 1);
  } else {
    sn_wv_kernel<<<rows, 256, 0, ST>>>(w, v, nullptr, eps, rows, cols, s, nrm + 1);
  }
  VG_LAUNCH_OK();
  sn_finish_kernel<<<cdiv(std::max(rows, cols), 256), 256, 0, ST>>>(t, s, nrm, eps, rows, cols, training, u, v, sigma);
  VG_LAUNCH_OK();
  return 0;
}
// scratch: fp32 [1]
extern "C" int vg_spectral_bwd(const float* g, const float* w_orig, const float* u, const float* v, const float* sigma,
                               int rows, int cols, float* dw, int accumulate, float* scratch, void* stream_) {
  const long long n = static_cast<long long>(rows) * cols;
  VG_CUDA(cudaMemsetAsync(scratch, 0, sizeof(float), ST));
  dot_kernel<<<lo_grid(n, 1024), 256, 0, ST>>>(g, w_orig, n, scratch);
  VG_LAUNCH_OK();
  sn_bwd_kernel<<<lo_grid(n), 256, 0, ST>>>(g, u, v, sigma, scratch, rows, cols, dw, accumulate);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_sumsq(const float* g, long long n, float* out, int zero_first, void* stream_) {
  VG_CHECK((reinterpret_cast<uintptr_t>(g) & 15) == 0, -1, "vg_sumsq: buffer must be 16-byte aligned");
  if (zero_first) VG_CUDA(cudaMemsetAsync(out, 0, sizeof(float), ST));
  sumsq_kernel<<<lo_grid(n, 2048), 256, 0, ST>>>(g, n, out);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_adam_step(float* p, float* g, float* m, float* v, long long n, float lr, float beta1, float beta2,
                            float eps, int step, const float* gnorm_sq, float max_norm, int write_back_grad,
                            void* stream_) {
  VG_CHECK(step >= 1, -1, "vg_adam_step: step counts from 1");
  const float bc1 = static_cast<float>(1.0 - pow(static_cast<double>(beta1), step));
  const float bc2 = static_cast<float>(1.0 - pow(static_cast<double>(beta2), step));
  adam_kernel<<<lo_grid(n, 1024), 256, 0, ST>>>(p, g, m, v, n, lr, beta1, beta2, eps, bc1, sqrtf(bc2), gnorm_sq, max_norm,
                                                write_back_grad);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_adam_prepare(float* state, float beta1, float beta2, void* stream_) {
  adam_prepare_kernel<<<1, 1, 0, ST>>>(state, beta1, beta2);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_multi_sumsq(const VgAdamTensor* table, int count, float* out, float* scratch, int scratch_len,
                              void* stream_) {
  const int grid = std::min(num_sms() * 4, scratch_len);
  VG_CHECK(grid >= 1, -1, "vg_multi_sumsq: scratch must hold at least one float");
  multi_sumsq_kernel<<<grid, 256, 0, ST>>>(table, count, scratch, adam_rotate());
  VG_LAUNCH_OK();
  final_sum_kernel<<<1, 256, 0, ST>>>(scratch, grid, out);
  VG_LAUNCH_OK();
  return 0;
}
extern "C" int vg_multi_adam(const VgAdamTensor* table, int count, float lr, float beta1, float beta2, float eps,
                             const float* state, const float* gnorm_sq, float max_norm, int write_back_grad,
                             void* stream_) {
  multi_adam_kernel<<<num_sms() * 4, 256, 0, ST>>>(table, count, lr, beta1, beta2, eps, state, gnorm_sq, max_norm,
                                                   write_back_grad, adam_rotate());
  VG_LAUNCH_OK();
  return 0;
}
