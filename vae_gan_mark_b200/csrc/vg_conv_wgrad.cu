// Implicit-GEMM convolution weight gradient ("wgrad") on tcgen05 tensor cores, sm_100a.
//
//   dW[co, tap*Cin + ci] = sum_{pixels p} dY[p, co] * X[p shifted by tap, ci]
//
// The reduction (GEMM K) runs over output pixels, which is the *strided* dimension of both NHWC
// operands, so both tiles are MN-major: a TMA box {64 channels, TW, 1, TH, TN} (64 pixels per K step)
// lands as [pixel][64 ch] rows of 128 swizzled bytes, which is exactly the canonical MN-major
// SWIZZLE_128B layout (8-pixel groups 1024 B apart = SBO, 64-channel groups one box apart = LBO).
// A = dY^T (128 output channels = 2 boxes), B = X^T at up to 4 (tap, 64-channel) blocks = N up to 256.
// Split-K over pixel tiles fills the machine; partial tiles are combined with fp32 atomics.
//
// Serves Conv2d wgrad and (with the operand roles swapped by the caller) ConvTranspose2d wgrad of
// the reference layers (vae-gan.py:52-60,76-81,153-157; vae-gan-v2.py:123-127,168-176,199-241).
#include <algorithm>
#include <stdlib.h>

#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

constexpr int kWgBM = 128;        // output channels per tile
constexpr int kWgPix = 64;        // pixels per K step
constexpr int kWgThreads = 256;
constexpr int kWgBoxBytes = 64 * kWgPix * 2;   // 8 KB: 64 px x 64 ch bf16
constexpr int kWgEpiPitch = 144;               // staged row: 32 fp32 + 16 B pad
constexpr int kWgEpiBytes = 4 * 32 * kWgEpiPitch;

struct WgradParams {
  int m_n, m_h, m_w;              // output pixel grid (reduction space)
  int tn, th, tw;                 // pixel tile, product == 64
  int tiles_n, tiles_h, tiles_w;
  int cout, cin, num_taps;
  int g_coff;                     // channel offset of dY inside its buffer
  int bn, nb;                     // N tile (64*nb)
  int blocks_total;               // num_taps * cin/64
  int m_tiles, n_tiles, ksplit, stages;
  float* dw;                      // [cout][num_taps*cin] fp32
  int dw_ld;
  int atomic;
  long long split_stride;         // deterministic mode: split s stores its partial tile at dw + s * split_stride (a
                                  // workspace of ksplit x cout x dw_ld floats) and wgrad_reduce_kernel adds the splits in
                                  // a fixed order; 0 otherwise
  int4 taps[VG_MAX_TAPS];         // {c_base, dw, sh, dh}
  int ds;                         // dual-shift mode for <= 64 output channels (see wgrad_plan): A = dY at two pixel shifts
  int ncombos;                    // split-precision operand pairs per pixel tile (1 = plain bf16)
  int combo_g[8], combo_x[8];     // channel offsets of the pair's planes
};

// k2 = true: CTA-pair variant (cta_group::2).  A cluster of two CTAs computes a 256 x bn tile of dW with ONE stream of
// tcgen05.mma of M = 256: CTA r stages the dY boxes of ITS 128 output channels and HALF of the X blocks (nb / 2 boxes) in
// its own shared memory -- 32 KB per stage instead of 48 KB (6 stages instead of 4) and half the X bytes through L2 and
// shared memory per CTA, which is what bounds the MN-major operand path -- and drains its own 128 TMEM lanes.
template <bool k2>
__global__ void __launch_bounds__(kWgThreads, 1)
conv_wgrad_kernel(const __grid_constant__ CUtensorMap tmap_g, const __grid_constant__ CUtensorMap tmap_x,
                  const __grid_constant__ WgradParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = 2 * kWgBoxBytes;
  const int nb_cta = k2 ? p.nb / 2 : p.nb;                 // X boxes this CTA stages per K step
  const uint32_t stage_bytes = a_bytes + nb_cta * kWgBoxBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + static_cast<size_t>(p.stages) * stage_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full = bars + 2 * p.stages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bars) + 256;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // broadcast: provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  const uint32_t rank = k2 ? (blockIdx.x & 1u) : 0u;       // 0 = leader (issues the MMAs); clusters are (2, 1, 1)

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_g);
    tma_prefetch_desc(&tmap_x);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], k2 ? 256 : 128);           // the leader's barrier collects the epilogues of both CTAs
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (k2) { tmem_alloc_2cta(tmem_base_slot, 512); tmem_relinquish_2cta(); }
    else { tmem_alloc(tmem_base_slot, 512); tmem_relinquish(); }
  }
  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_base_slot;

  // K steps: (pixel tile, operand pair) with the pair index fastest
  const int pix_tiles = p.tiles_n * p.tiles_h * p.tiles_w * p.ncombos;
  const int m_units = k2 ? (p.m_tiles + 1) / 2 : p.m_tiles;      // 256-channel pairs of M tiles, or single M tiles
  const int total_tiles = m_units * p.n_tiles * p.ksplit;
  const int cchunks = p.cin / 64;
  const int worker = k2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nworkers = k2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The whole warp walks the K loop with uniform control flow; the up to six copies of a stage (2 dY boxes + up to 4 X
    // boxes) are issued one after the other by ONE elected lane with warp-uniform coordinates (a few uniform-datapath
    // instructions each).  Issued from inside a per-lane region -- one thread, or one lane per box -- every copy sits in
    // an ELECT / R2UR waterfall loop of ~20 instructions, which starved the tensor pipe.
    const bool leader = elect_one();
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = worker; tile < total_tiles; tile += nworkers) {
      const int n_t = tile % p.n_tiles;
      const int m_u = (tile / p.n_tiles) % m_units;
      const int m_t = k2 ? 2 * m_u + static_cast<int>(rank) : m_u;
      const int split = tile / (p.n_tiles * m_units);
      const int k_begin = static_cast<int>((static_cast<long long>(pix_tiles) * split) / p.ksplit);
      const int k_end = static_cast<int>((static_cast<long long>(pix_tiles) * (split + 1)) / p.ksplit);
      const int nblk = min(p.nb, p.blocks_total - n_t * p.nb);          // X blocks of this N tile that exist
      const int aboxes = p.ds ? 2 : max(0, min(2, (p.cout - m_t * kWgBM + 63) / 64));
      // X blocks staged by this CTA: the pair splits the N tile in halves [0, nb/2) | [nb/2, nb)
      const int b0 = k2 ? static_cast<int>(rank) * nb_cta : 0;
      const int myb = max(0, min(nb_cta, nblk - b0));
      uint32_t tx_bytes = static_cast<uint32_t>(aboxes + myb) * kWgBoxBytes;
      if (k2) {      // the leader's barrier also receives the peer's boxes
        const int ab1 = max(0, min(2, (p.cout - (2 * m_u + 1) * kWgBM + 63) / 64));
        const int ab0 = max(0, min(2, (p.cout - (2 * m_u) * kWgBM + 63) / 64));
        const int nb0 = max(0, min(nb_cta, nblk)), nb1 = max(0, min(nb_cta, nblk - nb_cta));
        tx_bytes = static_cast<uint32_t>(ab0 + ab1 + nb0 + nb1) * kWgBoxBytes;
      }
      // box coordinates that do not change along K: dY boxes i < aboxes, X boxes j < myb
      const int ca0 = p.g_coff + m_t * kWgBM;
      int cb[4], dwv[4], shv[4], dhv[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int blk = min(n_t * p.nb + b0 + j, p.blocks_total - 1);
        const int4 t = p.taps[blk / cchunks];
        cb[j] = t.x + (blk % cchunks) * 64; dwv[j] = t.y; shv[j] = t.z; dhv[j] = t.w;
      }
      int combo = k_begin % p.ncombos;
      const int pt0 = k_begin / p.ncombos;
      int tw_i = pt0 % p.tiles_w;
      int th_i = (pt0 / p.tiles_w) % p.tiles_h;
      int tn_i = pt0 / (p.tiles_w * p.tiles_h);
      for (int k = k_begin; k < k_end; ++k) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader && (!k2 || rank == 0)) mbar_arrive_expect_tx(&full_bar[stage], tx_bytes);
        uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
        const int ow0 = tw_i * p.tw, oh0 = th_i * p.th, n0 = tn_i * p.tn;
        const int cg = ca0 + p.combo_g[combo], cx = p.combo_x[combo];
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          if (i < aboxes) {
            // dual-shift mode: the second box holds the SAME 64 channels one pixel column to the left (dY[p - (0, 1)])
            const int ca = p.ds ? cg : cg + i * 64, owa = p.ds ? ow0 - i : ow0;
            if (leader) {
              if (k2) tma_load_5d_2cta(sa + i * kWgBoxBytes, &tmap_g, &full_bar[stage], ca, owa, 0, oh0, n0);
              else tma_load_5d(sa + i * kWgBoxBytes, &tmap_g, &full_bar[stage], ca, owa, 0, oh0, n0);
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (j < myb) {
            if (leader) {
              if (k2) tma_load_5d_2cta(sa + a_bytes + j * kWgBoxBytes, &tmap_x, &full_bar[stage], cb[j] + cx, ow0 + dwv[j], shv[j],
                                       oh0 + dhv[j], n0);
              else tma_load_5d(sa + a_bytes + j * kWgBoxBytes, &tmap_x, &full_bar[stage], cb[j] + cx, ow0 + dwv[j], shv[j],
                               oh0 + dhv[j], n0);
            }
          }
        }
        if (++combo == p.ncombos) {
          combo = 0;
          if (++tw_i == p.tiles_w) { tw_i = 0; if (++th_i == p.tiles_h) { th_i = 0; ++tn_i; } }
        }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA only in the pair variant) =====================
    // whole warp, uniform control flow and addresses; only the tcgen05 instructions sit under the elected lane (see the
    // forward kernel and elect_one(): an `if (lane == 0)` region costs ~135 cycles of issue overhead per MMA)
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(k2 ? 2 * kWgBM : kWgBM, p.bn, 1, 1);
    const uint64_t proto = umma_smem_desc_sw128(0, kWgBoxBytes, 1024);
    const uint32_t d_hi = static_cast<uint32_t>(proto >> 32), d_lo0 = static_cast<uint32_t>(proto);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < total_tiles; tile += nworkers) {
      const int split = tile / (p.n_tiles * m_units);
      const int k_begin = static_cast<int>((static_cast<long long>(pix_tiles) * split) / p.ksplit);
      const int k_end = static_cast<int>((static_cast<long long>(pix_tiles) * (split + 1)) / p.ksplit);
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * 256);
      for (int k = k_begin; k < k_end; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        const uint32_t a_lo = d_lo0 + (sa >> 4), b_lo = d_lo0 + (sb >> 4);
#pragma unroll
        for (int j = 0; j < kWgPix / 16; ++j) {
          // 16 pixels (= 2 groups of 8 rows, SBO apart) per MMA; 64-channel groups LBO (= one box) apart
          const uint32_t accum = (k > k_begin || j > 0) ? 1u : 0u;
          if (leader) {
            if (k2) umma_bf16_2cta_lohi(d_tmem, a_lo + j * (16 * 128 >> 4), d_hi, b_lo + j * (16 * 128 >> 4), d_hi, idesc, accum);
            else umma_bf16_lohi(d_tmem, a_lo + j * (16 * 128 >> 4), d_hi, b_lo + j * (16 * 128 >> 4), d_hi, idesc, accum);
          }
        }
        if (leader) { if (k2) umma_commit_2cta(&empty_bar[stage]); else umma_commit(&empty_bar[stage]); }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      if (leader) { if (k2) umma_commit_2cta(&tmem_full[acc]); else umma_commit(&tmem_full[acc]); }
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue =====================
    // TMEM -> registers (thread = one output-channel row, 32 fp32 columns at a time) -> per-warp staging tile in
    // shared memory -> coalesced 16-byte stores / vector reductions (8 lanes cover 128 contiguous bytes of a dW row).
    const int quad = warp & 3;
    uint8_t* stg = epi_smem + quad * (32 * kWgEpiPitch);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = worker; tile < total_tiles; tile += nworkers) {
      const int n_t = tile % p.n_tiles;
      const int m_u = (tile / p.n_tiles) % m_units;
      const int m_t = k2 ? 2 * m_u + static_cast<int>(rank) : m_u;
      const long long split_off = static_cast<long long>(tile / (p.n_tiles * m_units)) * p.split_stride;
      const int co_warp = m_t * kWgBM + quad * 32;           // first output channel handled by this warp
      const int ncols = min(p.bn, p.blocks_total * 64 - n_t * p.bn);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * 256);
      for (int c = 0; c < ncols; c += 32) {
        uint32_t r[32];
        tmem_ld_32x32(t_row + c, r);
        tmem_ld_wait();
        // dual-shift mode: accumulator rows [0, 64) belong to dY[p], rows [64, 128) to dY[p - (0, 1)]; the 64-column block
        // (dh, x in {-1, 0}, 64-channel chunk) of row half j is tap (dh, x + j): j = 1, x = -1 duplicates j = 0, x = 0
        int row0 = co_warp, col0 = n_t * p.bn + c;
        bool skip = false;
        if (p.ds) {
          const int j = quad >> 1, blk = n_t * p.nb + (c >> 6);
          const int t6 = blk / cchunks, cc = blk - t6 * cchunks;
          const int x = (t6 & 1) - 1;
          skip = (j == 1 && x == -1);
          row0 = (quad & 1) * 32;
          col0 = (((t6 >> 1) * 3) + (x + j + 1)) * p.cin + cc * 64 + (c & 63);
        }
        uint8_t* dst = stg + lane * kWgEpiPitch;
#pragma unroll
        for (int j = 0; j < 32; j += 4) *reinterpret_cast<uint4*>(dst + j * 4) = make_uint4(r[j], r[j + 1], r[j + 2], r[j + 3]);
        __syncwarp();
        const int seg = lane & 7, rsub = lane >> 3;
#pragma unroll
        for (int it = 0; it < 8; ++it) {
          const int rr = it * 4 + rsub;
          if (!skip && row0 + rr < p.cout) {
            const float4 v = *reinterpret_cast<const float4*>(stg + rr * kWgEpiPitch + seg * 16);
            float* o = p.dw + split_off + static_cast<long long>(row0 + rr) * p.dw_ld + col0 + seg * 4;
            if (p.atomic) {
              asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(o), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w)
                           : "memory");
            } else {
              *reinterpret_cast<float4*>(o) = v;
            }
          }
        }
        __syncwarp();
      }
      tc_fence_before();
      if (k2) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
      if (++acc == 2) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (k2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

// deterministic split-K: dw[i] = ws[0][i] + ws[1][i] + ... in this order, whatever order the splits finished in
__global__ void wgrad_reduce_kernel(const float4* __restrict__ ws, float4* __restrict__ dw, long long n4, int ksplit) {
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n4;
       i += static_cast<long long>(gridDim.x) * blockDim.x) {
    float4 a = ws[i];
    for (int s = 1; s < ksplit; ++s) {
      const float4 b = ws[static_cast<long long>(s) * n4 + i];
      a.x += b.x; a.y += b.y; a.z += b.z; a.w += b.w;
    }
    dw[i] = a;
  }
}

// CTA-pair (cta_group::2) variant on / off (vg_set_cta_pairs; on by default)
static int g_wgrad_pairs = 1;
static bool wgrad_pairs_enabled() { return g_wgrad_pairs != 0; }

// split factor of a launch: the largest split with tiles * ksplit <= #SMs wastes the least of the last wave (a second,
// partial wave costs a full tile time); keep >= ~8 K steps per split so the pipeline fills
static int wgrad_auto_split(int tiles, int pix_tiles, int sms) {
  const int max_split = max(1, min(64, pix_tiles / 8));
  double best = -1.0;
  int ksplit = 1;
  for (int s = 1; s <= max_split; ++s) {
    const long long work = static_cast<long long>(tiles) * s;
    const double eff = static_cast<double>(work) / (static_cast<double>((work + sms - 1) / sms) * sms);
    if (eff > best + 1e-9) { best = eff; ksplit = s; }
  }
  return ksplit;
}


// Tiling decisions of a launch, shared by vg_conv_wgrad and vg_conv_wgrad_workspace.
//
// Dual-shift mode (ds): with <= 64 output channels half of the M = 128 rows of every MMA would be empty (the 64 -> 64 3x3
// layers at full resolution ran the tensor pipe 94 % busy at 0.35 of the peak).  The empty half is filled with the SAME 64
// channels of dY shifted one pixel column to the left: with A_j[p] = dY[p - (0, j)] and B blocks X[p + (dh, x)], x in {-1, 0},
//     sum_p A_j[p] X[p + (dh, x)] = sum_p' dY[p'] X[p' + (dh, x + j)] = dW[tap (dh, x + j)],
// so six X blocks per 64-channel chunk (instead of nine) produce all nine taps ((j, x) = (0,-1), (0,0), (1,0); (1,-1) is a
// duplicate of (0,0) and is not stored): N = 384 per chunk on full M instead of N = 576 on half of M.  The pixel grid is
// extended by one column (p' = p - (0,1) must reach column W - 1), which is why the pixel tile is 8 x 8 here (+1 column of
// 8-wide tiles instead of +1 column of 64-wide ones); everything outside the image is TMA zero-fill on both operands.
struct WgPlan {
  int tw, th, tn, tiles_n, tiles_h, tiles_w, ncombos, pix_tiles, blocks_total, nb, m_tiles, m_units, ksplit, ds;
  bool pair;
};
static bool wgrad_plan(const VgConvWgrad* d, WgPlan* q) {
  static const int ds_env = getenv("VG_WGRAD_DS") ? atoi(getenv("VG_WGRAD_DS")) : 1;
  q->ncombos = d->num_combos > 1 ? d->num_combos : 1;
  bool ds = ds_env != 0 && d->cout <= 64 && d->num_taps == 9 && d->x_stride == 1 && q->ncombos == 1 && d->force_bn == 0 &&
            d->m_w >= 8 && d->m_h >= 8 && d->m_w == d->x_w && d->m_h == d->x_h;
  for (int i = 0; ds && i < 9; ++i)
    ds = d->taps[i][0] == d->taps[0][0] && d->taps[i][1] == i % 3 - 1 && d->taps[i][2] == 0 && d->taps[i][3] == i / 3 - 1;
  if (ds && static_cast<long long>(d->m_n) * d->m_h * d->m_w < 128 * kWgPix) ds = false;      // tiny: launch-latency bound
  q->ds = ds ? 1 : 0;
  if (ds) {
    q->tw = 8; q->th = 8; q->tn = 1;
    q->tiles_w = cdiv(d->m_w + 1, 8);
  } else {
    int w = 1;
    while (w < d->m_w && w < kWgPix) w <<= 1;
    int h = 1;
    while (h < d->m_h && w * h < kWgPix) h <<= 1;
    q->tw = w; q->th = h; q->tn = kWgPix / (w * h);
    q->tiles_w = cdiv(d->m_w, q->tw);
  }
  q->tiles_n = cdiv(d->m_n, q->tn); q->tiles_h = cdiv(d->m_h, q->th);
  q->pix_tiles = q->tiles_n * q->tiles_h * q->tiles_w * q->ncombos;
  q->blocks_total = (ds ? 6 : d->num_taps) * (d->cin / 64);
  q->m_tiles = ds ? 1 : cdiv(d->cout, kWgBM);
  int nb = min(4, q->blocks_total);
  if (d->force_bn == 64 || d->force_bn == 128 || d->force_bn == 192 || d->force_bn == 256) nb = min(nb, d->force_bn / 64);
  q->nb = nb;
  // CTA pairs: at least two 128-channel M tiles to pair up and an N tile that splits in halves
  // (tiny problems are launch-latency bound and gain nothing from the cluster launch)
  q->pair = wgrad_pairs_enabled() && q->m_tiles >= 2 && (nb == 2 || nb == 4) && q->pix_tiles >= 128;
  q->m_units = q->pair ? (q->m_tiles + 1) / 2 : q->m_tiles;
  const int workers = q->pair ? conv_sms() / 2 : conv_sms();
  q->ksplit = d->ksplit > 0 ? d->ksplit : wgrad_auto_split(q->m_units * cdiv(q->blocks_total, nb), q->pix_tiles, workers);
  return true;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_conv_wgrad(const VgConvWgrad* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VG_CHECK(d != nullptr, -1, "vg_conv_wgrad: null descriptor");
  VG_CHECK(d->cin > 0 && d->cin % 64 == 0, -1, "vg_conv_wgrad: cin (%d) must be a positive multiple of 64", d->cin);
  VG_CHECK(d->cout >= 1, -1, "vg_conv_wgrad: cout");
  VG_CHECK(d->num_taps >= 1 && d->num_taps <= VG_MAX_TAPS, -1, "vg_conv_wgrad: num_taps %d out of range", d->num_taps);
  VG_CHECK(d->x_stride == 1 || d->x_stride == 2, -1, "vg_conv_wgrad: x_stride must be 1 or 2");
  VG_CHECK(d->x_h % d->x_stride == 0 && d->x_w % d->x_stride == 0, -1, "vg_conv_wgrad: H,W must divide by the stride");
  VG_CHECK(d->x_ld % 8 == 0 && d->g_ld % 8 == 0 && d->dw_ld % 4 == 0, -1, "vg_conv_wgrad: leading dims alignment");
  VG_CHECK(d->dw_ld >= d->num_taps * d->cin, -1, "vg_conv_wgrad: dw_ld too small");

  WgPlan q;
  wgrad_plan(d, &q);
  WgradParams p;
  memset(&p, 0, sizeof(p));
  p.m_n = d->m_n; p.m_h = d->m_h; p.m_w = d->m_w;
  p.tw = q.tw; p.th = q.th; p.tn = q.tn;
  p.tiles_n = q.tiles_n; p.tiles_h = q.tiles_h; p.tiles_w = q.tiles_w;
  p.ncombos = q.ncombos;
  p.ds = q.ds;
  VG_CHECK(p.ncombos <= 8, -1, "vg_conv_wgrad: at most 8 operand pairs");
  for (int i = 0; i < 8; ++i) { p.combo_g[i] = d->num_combos > 1 ? d->combo_g[i] : 0; p.combo_x[i] = d->num_combos > 1 ? d->combo_x[i] : 0; }
  const int pix_tiles = q.pix_tiles;
  p.cout = d->cout; p.cin = d->cin; p.num_taps = q.ds ? 6 : d->num_taps; p.g_coff = d->g_coff;
  p.blocks_total = q.blocks_total;
  p.m_tiles = q.m_tiles;
  const int sms = conv_sms();
  const int nb = q.nb;
  p.nb = nb; p.bn = nb * 64;
  p.n_tiles = cdiv(p.blocks_total, nb);
  const bool pair = q.pair;
  const int workers = pair ? sms / 2 : sms;
  const int m_units = q.m_units;
  const int ksplit = q.ksplit;
  VG_CHECK(ksplit <= pix_tiles, -1, "vg_conv_wgrad: ksplit %d > pixel tiles %d", ksplit, pix_tiles);
  p.ksplit = ksplit;
  p.atomic = ksplit > 1 ? 1 : 0;
  const size_t dw_floats = static_cast<size_t>(d->cout) * d->dw_ld;
  const bool two_stage = ksplit > 1 && d->workspace != nullptr;
  if (two_stage) {
    VG_CHECK(d->workspace_bytes >= static_cast<long long>(ksplit * dw_floats * sizeof(float)), -1,
             "vg_conv_wgrad: workspace of %lld bytes is too small for %d splits of %zu floats (vg_conv_wgrad_workspace)",
             d->workspace_bytes, ksplit, dw_floats);
    VG_CHECK(dw_floats % 4 == 0 && (reinterpret_cast<uintptr_t>(d->workspace) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(d->dw) & 15) == 0, -1, "vg_conv_wgrad: workspace / dw must be 16-byte aligned");
    p.atomic = 0;
    p.split_stride = static_cast<long long>(dw_floats);
  }
  const int stage_bytes = (2 + (pair ? nb / 2 : nb)) * kWgBoxBytes;
  p.stages = min(8, (227 * 1024 - 1024 - 256 - kWgEpiBytes) / stage_bytes);
  p.dw = two_stage ? static_cast<float*>(d->workspace) : d->dw;
  p.dw_ld = d->dw_ld;
  for (int i = 0; i < d->num_taps; ++i) {
    p.taps[i] = make_int4(d->taps[i][0], d->taps[i][1], d->taps[i][2], d->taps[i][3]);
    VG_CHECK(d->taps[i][2] >= 0 && d->taps[i][2] < d->x_stride, -1, "vg_conv_wgrad: tap %d row parity out of range", i);
  }
  if (q.ds)      // the six X shifts (dh, x) of the dual-shift mode, block index t6 = (dh + 1) * 2 + (x + 1)
    for (int t6 = 0; t6 < 6; ++t6) p.taps[t6] = make_int4(d->taps[0][0], (t6 & 1) - 1, 0, (t6 >> 1) - 1);
  if (p.atomic)
    VG_CUDA(cudaMemsetAsync(d->dw, 0, dw_floats * sizeof(float), stream));

  CUtensorMap tmap_g, tmap_x;
  const uint32_t box[5] = {64, static_cast<uint32_t>(p.tw), 1, static_cast<uint32_t>(p.th), static_cast<uint32_t>(p.tn)};
  {
    const uint64_t ld = d->g_ld, W = d->m_w, H = d->m_h;
    uint64_t dims[5] = {ld, W, 1, H, static_cast<uint64_t>(d->m_n)};
    uint64_t strides[5] = {1, ld, W * ld, W * ld, H * W * ld};
    int rc = encode_tmap_bf16(&tmap_g, d->g, 5, dims, strides, box);
    if (rc) return rc;
  }
  {
    const int s = d->x_stride;
    const uint64_t ld = d->x_ld, W = d->x_w, H = d->x_h;
    uint64_t dims[5] = {ld * s, W / s, static_cast<uint64_t>(s), H / s, static_cast<uint64_t>(d->x_n)};
    uint64_t strides[5] = {1, ld * s, W * ld, W * ld * s, H * W * ld};
    int rc = encode_tmap_bf16(&tmap_x, d->x, 5, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + 1024 + 256 + kWgEpiBytes;
  static bool attr_set[64] = {false};      // per device: function attributes belong to the device's context
  const int dev = current_device();
  if (!attr_set[dev]) {
    VG_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(conv_wgrad_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  const int total_tiles = m_units * p.n_tiles * p.ksplit;
  if (pair) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(static_cast<unsigned>(2 * min(total_tiles, workers)), 1, 1);
    cfg.blockDim = dim3(kWgThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VG_CUDA(cudaLaunchKernelEx(&cfg, conv_wgrad_kernel<true>, tmap_g, tmap_x, p));
  } else {
    conv_wgrad_kernel<false><<<min(total_tiles, sms), kWgThreads, smem, stream>>>(tmap_g, tmap_x, p);
  }
  VG_LAUNCH_OK();
  if (two_stage) {
    const long long n4 = static_cast<long long>(dw_floats / 4);
    const int grid = static_cast<int>(std::min<long long>((n4 + 255) / 256, static_cast<long long>(num_sms()) * 8));
    wgrad_reduce_kernel<<<std::max(grid, 1), 256, 0, stream>>>(static_cast<const float4*>(d->workspace),
                                                             reinterpret_cast<float4*>(d->dw), n4, ksplit);
    VG_LAUNCH_OK();
  }
  return 0;
}

extern "C" int vg_set_cta_pairs(int wgrad_on) {
  g_wgrad_pairs = wgrad_on;
  return 0;
}

/* Bytes of scratch a deterministic (two-stage) launch of this descriptor needs: ksplit x cout x dw_ld floats, 0 when the
 * launch does not split.  Same split rule as vg_conv_wgrad. */
extern "C" long long vg_conv_wgrad_workspace(const VgConvWgrad* d) {
  if (d == nullptr || d->cin <= 0 || d->cin % 64 != 0 || d->cout < 1) return -1;
  WgPlan q;
  wgrad_plan(d, &q);
  return q.ksplit > 1 ? static_cast<long long>(q.ksplit) * d->cout * d->dw_ld * static_cast<long long>(sizeof(float)) : 0;
}
