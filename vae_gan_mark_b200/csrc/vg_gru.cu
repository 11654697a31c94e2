// Recurrent part of the bidirectional GRU text encoder (CharacterTokenEncoder, vae-gan-v2.py:65-114; torch.nn.GRU
// semantics, gate order r, z, n).  The stock path launches ~1000 tiny kernels per training step (one GEMM + one
// pointwise kernel per time step, layer and direction, forward and backward); here ONE launch walks all T time steps
// of one layer, both directions at once:
//
//   * a thread-block cluster of 8 CTAs owns one (direction, group of 8 batch rows); each CTA owns 32 of the 256
//     hidden units.  Independent (direction, batch group) pairs run on different clusters -- no sync between them.
//   * every thread keeps its 96-element slice of W_hh in REGISTERS for the whole sequence (the recurrent weights are
//     read from HBM once per launch, not once per time step);
//   * the state (forward: h_t; backward: d(hidden projection)) is exchanged between the 8 CTAs through distributed
//     shared memory: every thread pushes its value into all 8 CTAs with st.async, which signals the receiver's
//     mbarrier by byte count -- no cluster barrier and no release fence (that would also wait for the step's global
//     stores) on the per-step critical path, instead of a kernel boundary (~5 us) per step.
//
// The time-parallel GEMMs around the recurrence (x W_ih^T for all t, and the weight / input gradients) stay plain
// library GEMMs on the host side (layers.py GRULayerFn).  All arithmetic is fp32 FMA.
#include <cooperative_groups.h>

#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace cg = cooperative_groups;

namespace vg {

constexpr int kGruH = 256;    // hidden size
constexpr int kGruCS = 8;     // CTAs per cluster
constexpr int kGruHS = 32;    // hidden units per CTA
constexpr int kGruThreads = 256;

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }
__device__ __forceinline__ uint32_t map_cluster(uint32_t saddr, int rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
  return r;
}
// remote (or local) shared-memory store that completes `4 bytes` on the destination CTA's mbarrier
__device__ __forceinline__ void st_async_f32(uint32_t cluster_addr, float v, uint32_t cluster_bar) {
  asm volatile("st.async.shared::cluster.mbarrier::complete_tx::bytes.f32 [%0], %1, [%2];" ::"r"(cluster_addr), "f"(v),
               "r"(cluster_bar)
               : "memory");
}

// xproj [B][T][2][3H] (x W_ih^T + b_ih of both directions), w_hh [2][3H][H], b_hh [2][3H]
// out [B][T][2H] (direction d in columns [d*H, (d+1)*H)), gates [2][B][T][4][H] = r, z, n, (W_hn h + b_hn)
// BG = batch rows per cluster (8 or 16): the host picks the smallest BG whose clusters are all co-resident, because a
// second wave of clusters doubles the (latency-bound) run time.
template <int BG>
__global__ void __launch_bounds__(kGruThreads, 1)
gru_seq_fwd_kernel(const float* __restrict__ xproj, const float* __restrict__ w_hh, const float* __restrict__ b_hh,
                   float* __restrict__ out, float* __restrict__ gates, int B, int T) {
  constexpr int H = kGruH;
  constexpr int RB = BG / 8;                           // batch rows per thread in the pointwise phase
  extern __shared__ __align__(16) float gsm[];
  float (*hbuf)[BG][H] = reinterpret_cast<float (*)[BG][H]>(gsm);                                   // [2][BG][H]
  float (*red)[3][BG][kGruHS] = reinterpret_cast<float (*)[3][BG][kGruHS]>(gsm + 2 * BG * H);       // [8][3][BG][32]
  __shared__ __align__(8) uint64_t hbar[2];           // hbar[i]: all of hbuf[i] has arrived
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int cid = blockIdx.x / kGruCS;                 // cluster index = dir * nbg + bg
  const int nbg = (B + BG - 1) / BG;
  const int dir = cid / nbg, bg = cid % nbg;
  const int j0 = rank * kGruHS;
  const int tid = threadIdx.x;
  // phase-1 role: (kq, j): partial dot products over k in [kq*32, kq*32+32) for the 3 gate rows of hidden unit j0+j
  const int kq = tid >> 5, j = tid & 31;
  float w[3][32];
#pragma unroll
  for (int g = 0; g < 3; ++g) {
    const float4* src = reinterpret_cast<const float4*>(w_hh + (static_cast<size_t>(dir) * 3 * H + g * H + j0 + j) * H + kq * 32);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const float4 v = src[i];
      w[g][4 * i] = v.x; w[g][4 * i + 1] = v.y; w[g][4 * i + 2] = v.z; w[g][4 * i + 3] = v.w;
    }
  }
  // phase-2 role: (fb + 8*rb, fj): RB batch rows, one hidden unit
  const int fb = tid >> 5, fj = tid & 31;
  float bh[3];
#pragma unroll
  for (int g = 0; g < 3; ++g) bh[g] = b_hh[dir * 3 * H + g * H + j0 + fj];
  for (int i = tid; i < 2 * BG * H; i += kGruThreads) gsm[i] = 0.f;
  if (tid == 0) {
    mbar_init(&hbar[0], 1);
    mbar_init(&hbar[1], 1);
    fence_barrier_init();
  }
  uint32_t peer_h[kGruCS], peer_bar[kGruCS];
#pragma unroll
  for (int r = 0; r < kGruCS; ++r) {
    peer_h[r] = map_cluster(smem_u32(gsm), r);
    peer_bar[r] = map_cluster(smem_u32(&hbar[0]), r);
  }
  cluster.sync();      // every CTA of the cluster is running, has zeroed its state and initialised its barriers

  for (int step = 0; step < T; ++step) {
    const int t = dir ? T - 1 - step : step;
    const int cur = step & 1;
    // hbuf[cur^1] is filled during this step by all 8 x 256 threads of the cluster (4*RB bytes each per destination);
    // its previous fill was consumed two steps ago.  Nobody overwrites hbuf[cur] before every thread of every CTA
    // has sent its step-`step` values, i.e. has finished reading hbuf[cur].
    if (tid == 0 && step < T - 1) mbar_arrive_expect_tx(&hbar[cur ^ 1], BG * H * 4);
    float xg[RB][3];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int bglob = bg * BG + fb + 8 * rb;
#pragma unroll
      for (int g = 0; g < 3; ++g) xg[rb][g] = 0.f;
      if (bglob < B) {
        const float* xp = xproj + ((static_cast<size_t>(bglob) * T + t) * 2 + dir) * 3 * H + j0 + fj;
#pragma unroll
        for (int g = 0; g < 3; ++g) xg[rb][g] = __ldg(xp + g * H);
      }
    }
    if (step > 0) mbar_wait(&hbar[cur], ((step - 1) >> 1) & 1);
    float acc[3][BG];
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int b = 0; b < BG; ++b) acc[g][b] = 0.f;
    // 4 batch rows x 3 gates = 12 independent accumulation chains in flight (issue order written out explicitly:
    // back-to-back dependent FMAs would cost the full 4-cycle latency each)
#pragma unroll
    for (int b = 0; b < BG; b += 4) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float hq[4][4];
#pragma unroll
        for (int bb = 0; bb < 4; ++bb) {
          const float4 hv = *reinterpret_cast<const float4*>(&hbuf[cur][b + bb][kq * 32 + 4 * i]);
          hq[bb][0] = hv.x; hq[bb][1] = hv.y; hq[bb][2] = hv.z; hq[bb][3] = hv.w;
        }
#pragma unroll
        for (int c = 0; c < 4; ++c)
#pragma unroll
          for (int g = 0; g < 3; ++g)
#pragma unroll
            for (int bb = 0; bb < 4; ++bb) acc[g][b + bb] = fmaf(w[g][4 * i + c], hq[bb][c], acc[g][b + bb]);
      }
    }
#pragma unroll
    for (int g = 0; g < 3; ++g)
#pragma unroll
      for (int b = 0; b < BG; ++b) red[kq][g][b][j] = acc[g][b];
    __syncthreads();
    float s[RB][3], hprev[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
#pragma unroll
      for (int g = 0; g < 3; ++g) {
        float v = bh[g];
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[q][g][fb + 8 * rb][fj];
        s[rb][g] = v;
      }
      hprev[rb] = hbuf[cur][fb + 8 * rb][j0 + fj];
    }
    __syncthreads();      // red and hbuf[cur] are not read again in this step
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int bl = fb + 8 * rb, bglob = bg * BG + bl;
      const float r = sigmoidf_(xg[rb][0] + s[rb][0]);
      const float z = sigmoidf_(xg[rb][1] + s[rb][1]);
      const float n = tanhf(xg[rb][2] + r * s[rb][2]);
      const float hnew = (1.f - z) * n + z * hprev[rb];
      if (step < T - 1) {
        const uint32_t off = static_cast<uint32_t>(((cur ^ 1) * BG + bl) * H + j0 + fj) * 4u;
#pragma unroll
        for (int rr = 0; rr < kGruCS; ++rr) st_async_f32(peer_h[rr] + off, hnew, peer_bar[rr] + 8u * (cur ^ 1));
      }
      if (bglob < B) {
        out[(static_cast<size_t>(bglob) * T + t) * 2 * H + dir * H + j0 + fj] = hnew;
        if (gates != nullptr) {
          float* gp = gates + ((static_cast<size_t>(dir) * B + bglob) * T + t) * 4 * H + j0 + fj;
          gp[0] = r; gp[H] = z; gp[2 * H] = n; gp[3 * H] = s[rb][2];
        }
      }
    }
  }
  cluster.sync();        // no CTA leaves while stores into its shared memory may still be in flight
}

// dout [B][T][2H], out (forward output = h values) [B][T][2H], gates as above, w_hh [2][3H][H]
// dgx [B][T][2][3H] = gradient w.r.t. xproj;  dgh [2][B][T][3H] = gradient w.r.t. (h W_hh^T + b_hh)
template <int BG>
__global__ void __launch_bounds__(kGruThreads, 1)
gru_seq_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ out, const float* __restrict__ gates,
                   const float* __restrict__ w_hh, float* __restrict__ dgx, float* __restrict__ dgh, int B, int T) {
  constexpr int H = kGruH;
  constexpr int RB = BG / 8;
  extern __shared__ __align__(16) float gsm[];
  float (*gbuf)[BG][3 * H] = reinterpret_cast<float (*)[BG][3 * H]>(gsm);                        // [2][BG][768]
  float (*red)[BG][kGruHS] = reinterpret_cast<float (*)[BG][kGruHS]>(gsm + 2 * BG * 3 * H);      // [8][BG][32]
  __shared__ __align__(8) uint64_t gbar[2];           // gbar[i]: all of gbuf[i] has arrived
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = static_cast<int>(cluster.block_rank());
  const int cid = blockIdx.x / kGruCS;
  const int nbg = (B + BG - 1) / BG;
  const int dir = cid / nbg, bg = cid % nbg;
  const int j0 = rank * kGruHS;
  const int tid = threadIdx.x;
  // matvec role: (rq, k): partial sums over gate rows [rq*96, rq*96+96) for state column j0+k
  const int rq = tid >> 5, k = tid & 31;
  float w[96];
#pragma unroll
  for (int i = 0; i < 96; ++i) w[i] = w_hh[(static_cast<size_t>(dir) * 3 * H + rq * 96 + i) * H + j0 + k];
  const int fb = tid >> 5, fj = tid & 31;
  if (tid == 0) {
    mbar_init(&gbar[0], 1);
    mbar_init(&gbar[1], 1);
    fence_barrier_init();
  }
  uint32_t peer_g[kGruCS], peer_bar[kGruCS];
#pragma unroll
  for (int r = 0; r < kGruCS; ++r) {
    peer_g[r] = map_cluster(smem_u32(gsm), r);
    peer_bar[r] = map_cluster(smem_u32(&gbar[0]), r);
  }
  float carry[RB];        // d loss / d h_t flowing in from the later time step, for (fb + 8*rb, j0+fj)
#pragma unroll
  for (int rb = 0; rb < RB; ++rb) carry[rb] = 0.f;
  cluster.sync();

  for (int step = 0; step < T; ++step) {
    // walk the forward recurrence backwards: the forward pass of direction 0 ended at t = T-1, of direction 1 at t = 0
    const int t = dir ? step : T - 1 - step;
    const int tprev = dir ? t + 1 : t - 1;                  // time index of h_{prev} in the forward recurrence
    const bool has_prev = step < T - 1;                     // uniform across the cluster
    const int cur = step & 1;
    // gbuf[cur] was last read two steps ago; a peer reaches this step only after every thread of every CTA has sent
    // its values of the previous step, i.e. after all reads of gbuf[cur] of two steps ago
    if (tid == 0 && has_prev) mbar_arrive_expect_tx(&gbar[cur], BG * 3 * H * 4);
    float direct[RB];
#pragma unroll
    for (int rb = 0; rb < RB; ++rb) {
      const int bl = fb + 8 * rb, bglob = bg * BG + bl;
      float dgr = 0.f, dgz = 0.f, dgn = 0.f, dghn = 0.f;
      direct[rb] = 0.f;
      if (bglob < B) {
        const size_t o = (static_cast<size_t>(bglob) * T + t) * 2 * H + dir * H + j0 + fj;
        const float dh = __ldg(dout + o) + carry[rb];
        const float* gp = gates + ((static_cast<size_t>(dir) * B + bglob) * T + t) * 4 * H + j0 + fj;
        const float r = __ldg(gp), z = __ldg(gp + H), n = __ldg(gp + 2 * H), hn = __ldg(gp + 3 * H);
        const float hprev = has_prev ? __ldg(out + (static_cast<size_t>(bglob) * T + tprev) * 2 * H + dir * H + j0 + fj) : 0.f;
        const float dn = dh * (1.f - z);
        const float dz = dh * (hprev - n);
        direct[rb] = dh * z;
        dgn = dn * (1.f - n * n);
        dgz = dz * z * (1.f - z);
        dgr = dgn * hn * r * (1.f - r);
        dghn = dgn * r;
        float* xo = dgx + ((static_cast<size_t>(bglob) * T + t) * 2 + dir) * 3 * H + j0 + fj;
        xo[0] = dgr; xo[H] = dgz; xo[2 * H] = dgn;
        float* ho = dgh + ((static_cast<size_t>(dir) * B + bglob) * T + t) * 3 * H + j0 + fj;
        ho[0] = dgr; ho[H] = dgz; ho[2 * H] = dghn;
      }
      if (has_prev) {
        const uint32_t off = static_cast<uint32_t>((cur * BG + bl) * 3 * H + j0 + fj) * 4u;
#pragma unroll
        for (int rr = 0; rr < kGruCS; ++rr) {
          const uint32_t bar = peer_bar[rr] + 8u * cur;
          st_async_f32(peer_g[rr] + off, dgr, bar);
          st_async_f32(peer_g[rr] + off + H * 4u, dgz, bar);
          st_async_f32(peer_g[rr] + off + 2u * H * 4u, dghn, bar);
        }
      }
    }
    if (has_prev) {
      mbar_wait(&gbar[cur], (step >> 1) & 1);
      float acc[BG];
#pragma unroll
      for (int b = 0; b < BG; ++b) acc[b] = 0.f;
      // 8 batch rows = 8 independent accumulation chains in flight
#pragma unroll
      for (int b = 0; b < BG; b += 8) {
#pragma unroll
        for (int i = 0; i < 24; ++i) {
          float gq[8][4];
#pragma unroll
          for (int bb = 0; bb < 8; ++bb) {
            const float4 v = *reinterpret_cast<const float4*>(&gbuf[cur][b + bb][rq * 96 + 4 * i]);
            gq[bb][0] = v.x; gq[bb][1] = v.y; gq[bb][2] = v.z; gq[bb][3] = v.w;
          }
#pragma unroll
          for (int c = 0; c < 4; ++c)
#pragma unroll
            for (int bb = 0; bb < 8; ++bb) acc[b + bb] = fmaf(w[4 * i + c], gq[bb][c], acc[b + bb]);
        }
      }
#pragma unroll
      for (int b = 0; b < BG; ++b) red[rq][b][k] = acc[b];
      __syncthreads();
#pragma unroll
      for (int rb = 0; rb < RB; ++rb) {
        float v = direct[rb];
#pragma unroll
        for (int q = 0; q < 8; ++q) v += red[q][fb + 8 * rb][fj];
        carry[rb] = v;
      }
      __syncthreads();    // red is rewritten in the next step
    }
  }
  cluster.sync();         // no CTA leaves while stores into its shared memory may still be in flight
}

template <int BG>
static int gru_smem_fwd() { return (2 * BG * kGruH + 8 * 3 * BG * kGruHS) * static_cast<int>(sizeof(float)); }
template <int BG>
static int gru_smem_bwd() { return (2 * BG * 3 * kGruH + 8 * BG * kGruHS) * static_cast<int>(sizeof(float)); }

static void gru_config(cudaLaunchConfig_t& cfg, cudaLaunchAttribute* attr, int clusters, int smem, cudaStream_t st) {
  cfg = cudaLaunchConfig_t{};
  cfg.gridDim = dim3(static_cast<unsigned>(clusters * kGruCS), 1, 1);
  cfg.blockDim = dim3(kGruThreads, 1, 1);
  cfg.dynamicSmemBytes = static_cast<size_t>(smem);
  cfg.stream = st;
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kGruCS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
}
// opt in to the dynamic shared memory and ask how many clusters fit on the device at once
template <typename Kernel>
static int gru_query(Kernel kernel, int smem, int* max_clusters) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  gru_config(cfg, attr, 1, smem, nullptr);
  VG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  VG_CUDA(cudaOccupancyMaxActiveClusters(max_clusters, kernel, &cfg));
  return 0;
}
// launch `clusters` clusters of kGruCS CTAs
template <typename Kernel, typename... Args>
static int gru_launch(Kernel kernel, int clusters, int smem, cudaStream_t st, Args... args) {
  cudaLaunchConfig_t cfg;
  cudaLaunchAttribute attr[1];
  gru_config(cfg, attr, clusters, smem, st);
  VG_CUDA(cudaLaunchKernelEx(&cfg, kernel, args...));
  VG_LAUNCH_OK();
  return 0;
}

// smallest batch group whose clusters are all co-resident (one wave); cached per kernel direction
static int g_gru_max_clusters[64][2][2];   // [device][fwd|bwd][BG 8|16], 0 = not queried yet (stored as n + 1)

}  // namespace vg

extern "C" int vg_gru_max_active_clusters(int backward, int batch_group) {
  using namespace vg;
  if (batch_group != 8 && batch_group != 16) return -1;
  int& slot = g_gru_max_clusters[current_device()][backward ? 1 : 0][batch_group == 16 ? 1 : 0];
  if (slot == 0) {
    int n = 0, rc;
    if (!backward)
      rc = batch_group == 8 ? gru_query(gru_seq_fwd_kernel<8>, gru_smem_fwd<8>(), &n)
                            : gru_query(gru_seq_fwd_kernel<16>, gru_smem_fwd<16>(), &n);
    else
      rc = batch_group == 8 ? gru_query(gru_seq_bwd_kernel<8>, gru_smem_bwd<8>(), &n)
                            : gru_query(gru_seq_bwd_kernel<16>, gru_smem_bwd<16>(), &n);
    if (rc != 0) return rc;
    slot = n + 1;
  }
  return slot - 1;
}

static int gru_pick_group(int backward, int batch) {
  const int fit8 = vg_gru_max_active_clusters(backward, 8);
  if (fit8 < 0) return fit8;
  if (2 * ((batch + 7) / 8) <= fit8) return 8;
  const int fit16 = vg_gru_max_active_clusters(backward, 16);
  if (fit16 < 0) return fit16;
  return 16;
}

extern "C" int vg_gru_seq_fwd(const float* xproj, const float* w_hh, const float* b_hh, float* out, float* gates,
                              int batch, int steps, int hidden, void* stream_) {
  using namespace vg;
  VG_CHECK(hidden == kGruH, -1, "vg_gru_seq_fwd: hidden size must be %d (got %d)", kGruH, hidden);
  VG_CHECK(batch >= 1 && steps >= 1, -1, "vg_gru_seq_fwd: batch and steps must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int bg = gru_pick_group(0, batch);
  if (bg < 0) return bg;
  if (bg == 8)
    return gru_launch(gru_seq_fwd_kernel<8>, 2 * ((batch + 7) / 8), gru_smem_fwd<8>(), st, xproj, w_hh,
                      b_hh, out, gates, batch, steps);
  return gru_launch(gru_seq_fwd_kernel<16>, 2 * ((batch + 15) / 16), gru_smem_fwd<16>(), st, xproj, w_hh,
                    b_hh, out, gates, batch, steps);
}

extern "C" int vg_gru_seq_bwd(const float* dout, const float* out, const float* gates, const float* w_hh, float* dgx,
                              float* dgh, int batch, int steps, int hidden, void* stream_) {
  using namespace vg;
  VG_CHECK(hidden == kGruH, -1, "vg_gru_seq_bwd: hidden size must be %d (got %d)", kGruH, hidden);
  VG_CHECK(batch >= 1 && steps >= 1, -1, "vg_gru_seq_bwd: batch and steps must be >= 1");
  cudaStream_t st = static_cast<cudaStream_t>(stream_);
  const int bg = gru_pick_group(1, batch);
  if (bg < 0) return bg;
  if (bg == 8)
    return gru_launch(gru_seq_bwd_kernel<8>, 2 * ((batch + 7) / 8), gru_smem_bwd<8>(), st, dout, out,
                      gates, w_hh, dgx, dgh, batch, steps);
  return gru_launch(gru_seq_bwd_kernel<16>, 2 * ((batch + 15) / 16), gru_smem_bwd<16>(), st, dout, out,
                    gates, w_hh, dgx, dgh, batch, steps);
}
