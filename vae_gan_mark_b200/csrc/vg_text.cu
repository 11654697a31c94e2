// Front end of the character text encoder on the device (CharacterTokenEncoder, vae-gan-v2.py:65-114):
//   * tokenisation: code point -> vocabulary index through a lookup table (tokens_to_indices, :89-100, is a Python loop
//     with a dict lookup per character on the host in the reference);
//   * nn.Embedding forward (a row gather) and backward (a per-vocabulary-row sum of the token gradients: deterministic,
//     no sort, no atomics; the reference's embedding_dense_backward sorts the indices first);
//   * adaptive average pooling of the GRU output (B, L, C) along L into the NHWC text map [B][1][W/16][C]
//     (adaptive_pool + unsqueeze of :107-113 and the layout change to this package's activation format in one pass).
// All HBM-trivial (tens of KB); the point is that nothing of the text path runs on the host or in library kernels.
#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

__global__ void tokenize_kernel(const uint32_t* __restrict__ cp, long long n, const int* __restrict__ lut, int lut_size,
                                long long* __restrict__ idx) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const uint32_t c = cp[i];
    idx[i] = c < static_cast<uint32_t>(lut_size) ? lut[c] : 0;
  }
}

// out[t][:] = weight[idx[t]][:]
__global__ void embedding_fwd_kernel(const long long* __restrict__ idx, long long n, const float* __restrict__ weight, int v,
                                     int d, float* __restrict__ out) {
  const long long total = n * d;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const long long t = i / d;
    const int k = static_cast<int>(i - t * d);
    long long r = idx[t];
    if (r < 0 || r >= v) r = 0;
    out[i] = weight[r * d + k];
  }
}

// dw[r][k] = sum over tokens t with idx[t] == r of g[t][k]; row padding_idx stays zero.  One block per vocabulary row.
// Phase 1: all threads scan the tokens and mark the matches in a shared-memory bitmap (order-independent); phase 2:
// thread k walks the set bits in ascending token order and sums column k -- a fixed summation order, so the result is
// deterministic, without a sort and without atomics on the gradient.
constexpr int kEmbChunk = 8192;      // tokens per bitmap pass (1 KB of shared memory)
__global__ void embedding_bwd_kernel(const long long* __restrict__ idx, long long n, const float* __restrict__ g, int d,
                                     int padding_idx, float* __restrict__ dw) {
  __shared__ unsigned bits[kEmbChunk / 32];
  const int r = blockIdx.x;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};      // columns threadIdx.x + j * blockDim.x (d <= 4 * blockDim.x)
  if (r != padding_idx) {
    for (long long t0 = 0; t0 < n; t0 += kEmbChunk) {
      const int len = static_cast<int>(min(static_cast<long long>(kEmbChunk), n - t0));
      for (int w = threadIdx.x; w < kEmbChunk / 32; w += blockDim.x) bits[w] = 0u;
      __syncthreads();
      for (int t = threadIdx.x; t < len; t += blockDim.x)
        if (idx[t0 + t] == r) atomicOr(&bits[t >> 5], 1u << (t & 31));
      __syncthreads();
      for (int w = 0; w < (len + 31) / 32; ++w) {
        unsigned m = bits[w];
        while (m) {
          const int t = (w << 5) + __ffs(m) - 1;
          m &= m - 1;
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const int k = threadIdx.x + j * blockDim.x;
            if (k < d) acc[j] += g[(t0 + t) * d + k];
          }
        }
      }
      __syncthreads();
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const int k = threadIdx.x + j * blockDim.x;
    if (k < d) dw[static_cast<long long>(r) * d + k] = acc[j];
  }
}

// adaptive average pooling along L: bin j = [floor(j*L/W), ceil((j+1)*L/W))  (torch's AdaptiveAvgPool1d)
VG_DEVICE int bin_lo(int j, int l, int w) { return (j * l) / w; }
VG_DEVICE int bin_hi(int j, int l, int w) { return ((j + 1) * l + w - 1) / w; }

template <typename TI, typename TO>
__global__ void seqpool_fwd_kernel(const TI* __restrict__ seq, int in_ld, int b, int l, int c, int w, TO* __restrict__ out,
                                   int out_ld) {
  const long long total = static_cast<long long>(b) * w * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % c);
    const int j = static_cast<int>((i / c) % w);
    const long long n = i / (static_cast<long long>(c) * w);
    const int lo = bin_lo(j, l, w), hi = bin_hi(j, l, w);
    float a = 0.f;
    for (int t = lo; t < hi; ++t) a += static_cast<float>(seq[(n * l + t) * in_ld + k]);
    out[(n * w + j) * out_ld + k] = static_cast<TO>(a / static_cast<float>(hi - lo));
  }
}

// dseq[n][t][k] = sum over the bins j that contain t of dy[n][j][k] / len(j)
template <typename TI, typename TO>
__global__ void seqpool_bwd_kernel(const TI* __restrict__ dy, int dy_ld, int b, int l, int c, int w, TO* __restrict__ dseq,
                                   int out_ld) {
  const long long total = static_cast<long long>(b) * l * c;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    const int k = static_cast<int>(i % c);
    const int t = static_cast<int>((i / c) % l);
    const long long n = i / (static_cast<long long>(c) * l);
    // candidate bins: j with lo(j) <= t < hi(j); lo is non-decreasing in j, so scan a window around t*w/l
    int j0 = (t * w) / l;
    while (j0 > 0 && bin_hi(j0 - 1, l, w) > t) --j0;
    float a = 0.f;
    for (int j = j0; j < w && bin_lo(j, l, w) <= t; ++j) {
      const int lo = bin_lo(j, l, w), hi = bin_hi(j, l, w);
      if (t < hi) a += static_cast<float>(dy[(n * w + j) * dy_ld + k]) / static_cast<float>(hi - lo);
    }
    dseq[(n * l + t) * out_ld + k] = static_cast<TO>(a);
  }
}

static int text_grid(long long items) {
  long long blocks = (items + 255) / 256;
  const long long cap = static_cast<long long>(num_sms()) * 8;
  if (blocks > cap) blocks = cap;
  return static_cast<int>(blocks < 1 ? 1 : blocks);
}

}  // namespace vg

using namespace vg;

extern "C" int vg_tokenize(const uint32_t* codepoints, long long n, const int* lut, int lut_size, long long* idx,
                           void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(n >= 0 && lut_size >= 1, -1, "vg_tokenize: bad sizes");
  if (n == 0) return 0;
  tokenize_kernel<<<text_grid(n), 256, 0, st>>>(codepoints, n, lut, lut_size, idx);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_embedding_fwd(const long long* idx, long long n, const float* weight, int vocab, int dim, float* out,
                                void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(n >= 0 && vocab >= 1 && dim >= 1, -1, "vg_embedding_fwd: bad sizes");
  if (n == 0) return 0;
  embedding_fwd_kernel<<<text_grid(n * dim), 256, 0, st>>>(idx, n, weight, vocab, dim, out);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_embedding_bwd(const long long* idx, long long n, const float* g, int vocab, int dim, int padding_idx,
                                float* dw, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(n >= 0 && vocab >= 1 && dim >= 1 && dim <= 1024, -1, "vg_embedding_bwd: bad sizes (dim <= 1024)");
  embedding_bwd_kernel<<<vocab, 256, 0, st>>>(idx, n, g, dim, padding_idx, dw);
  VG_LAUNCH_OK();
  return 0;
}

template <typename TI, typename TO>
static int seqpool_fwd_impl(const void* seq, int in_ld, int b, int l, int c, int w, void* out, int out_ld, cudaStream_t st) {
  seqpool_fwd_kernel<TI, TO><<<text_grid(static_cast<long long>(b) * w * c), 256, 0, st>>>(
      static_cast<const TI*>(seq), in_ld, b, l, c, w, static_cast<TO*>(out), out_ld);
  VG_LAUNCH_OK();
  return 0;
}
template <typename TI, typename TO>
static int seqpool_bwd_impl(const void* dy, int dy_ld, int b, int l, int c, int w, void* dseq, int out_ld, cudaStream_t st) {
  seqpool_bwd_kernel<TI, TO><<<text_grid(static_cast<long long>(b) * l * c), 256, 0, st>>>(
      static_cast<const TI*>(dy), dy_ld, b, l, c, w, static_cast<TO*>(dseq), out_ld);
  VG_LAUNCH_OK();
  return 0;
}

extern "C" int vg_seqpool_fwd(const void* seq, int in_dtype, int in_ld, int b, int l, int c, int w, void* out, int out_dtype,
                              int out_ld, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(b >= 1 && l >= 1 && c >= 1 && w >= 1 && in_ld >= c && out_ld >= c, -1, "vg_seqpool_fwd: bad sizes");
  VG_CHECK((in_dtype == 0 || in_dtype == 1) && (out_dtype == 0 || out_dtype == 1), -1, "vg_seqpool_fwd: dtype codes are 0 (bf16) / 1 (fp32)");
  if (in_dtype == 1 && out_dtype == 0) return seqpool_fwd_impl<float, __nv_bfloat16>(seq, in_ld, b, l, c, w, out, out_ld, st);
  if (in_dtype == 1 && out_dtype == 1) return seqpool_fwd_impl<float, float>(seq, in_ld, b, l, c, w, out, out_ld, st);
  if (in_dtype == 0 && out_dtype == 0) return seqpool_fwd_impl<__nv_bfloat16, __nv_bfloat16>(seq, in_ld, b, l, c, w, out, out_ld, st);
  return seqpool_fwd_impl<__nv_bfloat16, float>(seq, in_ld, b, l, c, w, out, out_ld, st);
}

extern "C" int vg_seqpool_bwd(const void* dy, int dy_dtype, int dy_ld, int b, int l, int c, int w, void* dseq, int out_dtype,
                              int out_ld, void* stream_) {
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  VG_CHECK(b >= 1 && l >= 1 && c >= 1 && w >= 1 && dy_ld >= c && out_ld >= c, -1, "vg_seqpool_bwd: bad sizes");
  VG_CHECK((dy_dtype == 0 || dy_dtype == 1) && (out_dtype == 0 || out_dtype == 1), -1, "vg_seqpool_bwd: dtype codes are 0 (bf16) / 1 (fp32)");
  if (dy_dtype == 0 && out_dtype == 1) return seqpool_bwd_impl<__nv_bfloat16, float>(dy, dy_ld, b, l, c, w, dseq, out_ld, st);
  if (dy_dtype == 1 && out_dtype == 1) return seqpool_bwd_impl<float, float>(dy, dy_ld, b, l, c, w, dseq, out_ld, st);
  if (dy_dtype == 0 && out_dtype == 0) return seqpool_bwd_impl<__nv_bfloat16, __nv_bfloat16>(dy, dy_ld, b, l, c, w, dseq, out_ld, st);
  return seqpool_bwd_impl<float, __nv_bfloat16>(dy, dy_ld, b, l, c, w, dseq, out_ld, st);
}
