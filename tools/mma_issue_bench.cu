// Microbenchmark (dev tool): how fast does ONE thread get tcgen05.mma instructions of a given shape through the tensor pipe,
// depending on (a) the N extent, (b) how many TMEM accumulators the stream alternates between (dependent vs independent
// accumulation chains), (c) whether the A descriptor starts at a 1024-byte-aligned address or at a shifted row with
// 1280-byte group stride (the halo mode of the forward kernel), (d) K-major vs MN-major operands.
// Prints cycles per MMA; the floor for M = 128 is N / 2 cycles (8192 dense bf16 FLOP / cycle / SM).
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -o tools/bin/mma_issue_bench tools/mma_issue_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../vae_gan_mark_b200/csrc/vg_common.cuh"

namespace vg {
void set_error(const char*, ...) {}
unsigned long long g_launches = 0;
}  // namespace vg
using namespace vg;

struct Cfg {
  int n, nacc, shifted, mn_major, iters, mmas_per_commit, ctas_busy, vary_b, commits;
};

__global__ void __launch_bounds__(128, 1) bench_kernel(Cfg c, long long* cycles) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint64_t bar2[2];
  __shared__ uint32_t tmem_slot;
  // A: 48 KB region (enough for the shifted halo views), B: 32 KB
  uint8_t* sa = smem;
  uint8_t* sb = smem + 48 * 1024;      // up to 9 x 16 KB of weight tiles
  for (int i = threadIdx.x; i < 190 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0x3c003c00u + i;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    mbar_init(&bar2[0], 1 << 20);
    mbar_init(&bar2[1], 1 << 20);
    fence_barrier_init();
  }
  if (threadIdx.x < 32) { tmem_alloc(&tmem_slot, 512); tmem_relinquish(); }
  fence_proxy_async();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = tmem_slot;
  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
  if (warp == 0) {
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(128, c.n, c.mn_major, c.mn_major);
    const int acc_cols = 512 / c.nacc;
    uint32_t phase = 0;
    long long t0 = 0;
    for (int rep = 0; rep < 2; ++rep) {      // rep 0 = warm-up
      t0 = clock64();
      int acc = 0;
      // descriptors: the high word is constant per operand, only the 14-bit start-address field of the low word moves
      const uint64_t da0 = !c.mn_major ? umma_smem_desc_sw128(smem_u32(sa), 16, c.shifted ? 1280 : 1024)
                                       : umma_smem_desc_sw128(smem_u32(sa), 64 * 128, 1024);
      const uint64_t db0 = !c.mn_major ? umma_smem_desc_sw128(smem_u32(sb), 16, 1024)
                                       : umma_smem_desc_sw128(smem_u32(sb), 64 * 128, 1024);
      const uint32_t a_lo = static_cast<uint32_t>(da0), a_hi = static_cast<uint32_t>(da0 >> 32);
      const uint32_t b_lo = static_cast<uint32_t>(db0), b_hi = static_cast<uint32_t>(db0 >> 32);
      const uint32_t kstep = !c.mn_major ? 2u : (16u * 128u) >> 4;
      for (int it = 0; it < c.iters; ++it) {
#pragma unroll
        for (int j = 0; j < 36; ++j) {
          const int t = j / 4;
          const uint32_t shift = c.shifted ? static_cast<uint32_t>(((t / 3) * 10 + t % 3) * 128) >> 4 : 0u;
          const uint32_t bshift = c.vary_b ? static_cast<uint32_t>(t * c.n * 128) >> 4 : 0u;      // a different weight tile per tap
          if (leader) umma_bf16_lohi(tmem + acc * acc_cols, a_lo + shift + kstep * (j % 4), a_hi, b_lo + bshift + kstep * (j % 4), b_hi, idesc, j > 0 ? 1u : 0u);
          if (c.commits == 0) { if (++acc == c.nacc) acc = 0; }
        }
        if (c.commits) {      // the forward kernel's pattern: one accumulator per 36 MMAs, two commits after them
          if (leader) umma_commit(&bar2[0]);
          if (leader) umma_commit(&bar2[1]);
          if (++acc == c.nacc) acc = 0;
        }
      }
      if (leader) umma_commit(&bar);
      mbar_wait(&bar, phase);
      phase ^= 1;
    }
    const long long t1 = clock64();
    if (leader) cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x < 32) { tc_fence_after(); tmem_dealloc(tmem, 512); }
}

int main() {
  long long* d;
  cudaMalloc(&d, sizeof(long long) * 256);
  cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  printf("%5s %5s %8s %9s %6s | %12s %8s\n", "N", "nacc", "shifted", "mn_major", "ctas", "cyc/MMA", "floor");
  printf("%5s %5s %8s %9s %6s %7s %8s | %12s %8s\n", "N", "nacc", "shifted", "mn_major", "ctas", "vary_b", "commits", "cyc/MMA", "floor");
  for (int commits = 0; commits < 2; ++commits) {
    for (int vary_b = 0; vary_b < 2; ++vary_b) {
      for (int n : {64, 128, 256}) {
        for (int shifted = 0; shifted < 2; ++shifted) {
          for (int nacc : {1, 4}) {
            if (n * nacc > 512) continue;
            if (vary_b && n > 128) continue;      // 9 tiles of 32 KB do not fit
            Cfg c{n, nacc, shifted, 0, 200, 36, 148, vary_b, commits};
            bench_kernel<<<148, 128, 200 * 1024>>>(c, d);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            std::vector<long long> h(148);
            cudaMemcpy(h.data(), d, sizeof(long long) * 148, cudaMemcpyDeviceToHost);
            long long mx = 0;
            for (long long v : h) mx = v > mx ? v : mx;
            printf("%5d %5d %8d %9d %6d %7d %8d | %12.1f %8d\n", n, nacc, shifted, 0, 148, vary_b, commits, double(mx) / (200.0 * 36), n / 2);
          }
        }
      }
    }
  }
  return 0;
}
