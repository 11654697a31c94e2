// Implicit-GEMM convolution forward ("fprop") on tcgen05 tensor cores, sm_100a.
//
//   D[m, n] = sum_{tap, c} A[pixel(m) shifted by tap, c] * W[n, tap*Cin + c]
//
// * A (activations, NHWC bf16) is never materialised as an im2col matrix: each K step is one TMA
//   box {64 channels, TW, 1, TH, TN} of a 5-D view (s*ld, W/s, s, H/s, N) of the tensor, shifted by the
//   tap; TMA zero-fills out-of-bounds pixels, which implements the padding.  The stride-2 view
//   (s = 2) folds the column parity into the channel coordinate and the row parity into dim 2, so a
//   stride-2 conv (and the adjoint of a stride-2 ConvTranspose) is the same kernel.
// * W (bf16, [n_gemm][taps*Cin], K contiguous) is a 2-D TMA box {64, BN}.
// * Both land in shared memory in the 128-byte-swizzled K-major layout tcgen05.mma consumes.
// * b_mn_major: W is instead given as [K rows][N columns] (N contiguous) -- the forward operand [Cout][taps*Cin] of a
//   conv read as the B operand of its own DATA GRADIENT (K = Cout, N = Cin of one tap).  It is loaded as BN/64 boxes of
//   {64 N, 64 K} = the canonical MN-major SWIZZLE_128B layout, and the tensor core transposes it for free, so the data
//   gradient needs no second, transposed copy of the weights.
// * One elected thread issues tcgen05.mma (M=128, N=BN, K=16) into fp32 TMEM accumulators (2 x 256 columns, or
//   4 x 128 for BN <= 128); 8 epilogue warps drain TMEM with tcgen05.ld, add bias / activation, stage the tile in
//   XOR-swizzled shared memory and store bf16 or fp32 NHWC rows with 16-byte coalesced stores (optionally scattered
//   as a pixel shuffle for ConvTranspose, optionally into a channel slice of a wider buffer = fused concat,
//   optionally fp32 atomics for split-K).
// * Persistent: grid = min(tiles, #SMs); warp roles: 0 = TMA producer, 1 = MMA issuer, 2 = TMEM allocator,
//   4..11 = epilogue.
// * Per-launch variants: tile groups (the four parity classes of a stride-2 data gradient in one launch) and the
//   halo mode for narrow 3x3 layers (FpropParams).
//
// This one kernel serves Conv2d fprop, Conv2d dgrad (flipped weights), ConvTranspose2d fprop/dgrad and the
// GEMM-shaped layers (1x1 convs, full-kernel heads, bottleneck ConvT) of the reference models
// (vae-gan.py:52-60,76-81,153-157; vae-gan-v2.py:123-127,168-176,199-241).
#include <stdlib.h>
#include "vg_common.cuh"
#include "../../include/vaegan_b200.h"

namespace vg {

constexpr int kBM = 128;          // GEMM M tile (pixels) == TMEM lanes
constexpr int kBK = 64;           // K per stage: 64 bf16 = one 128-byte swizzle row
constexpr int kUmmaK = 16;
constexpr int kFpropThreads = 384;   // warps 0-3: TMA / MMA / TMEM alloc / spare; warps 4-11: epilogue
constexpr int kMaxTaps = VG_MAX_FPROP_TAPS;
constexpr int kHaloW = 10, kHaloH = 18, kHaloRows = kHaloW * kHaloH;   // halo of the 16 x 8 pixel tile
constexpr int kHaloBytes = 23 * 1024;                                    // 180 rows x 128 B, padded to a 1024 multiple
constexpr int kEpiStageBytes = 8 * 32 * 128;          // 8 epilogue warps x (32 rows x 128 B XOR-swizzled staging tile)
constexpr int kEpiBytes = kEpiStageBytes + 8 * 32 * 4;  // + per-warp 32-float bias window (read back as broadcast float4)

struct FpropParams {
  int m_n, m_h, m_w;            // output pixel grid
  int tn, th, tw;               // pixel tile, tn*th*tw == 128
  int tiles_n, tiles_h, tiles_w;
  int cin, num_taps, n_gemm, bn, n_tiles;
  int ksteps, ksplit, stages;
  // epilogue
  void* out;
  int out_kind;                 // 0 bf16, 1 fp32, 2 fp32 atomic add
  int out_h, out_w, out_ld, out_coff;
  int su_h, su_w, sub_h0, sub_w0, cout_per_sub;
  const float* bias;
  int act;                      // 0 none, 1 relu, 2 leaky relu 0.2
  int vec_ok;                   // destination allows 16-byte vector stores
  int b_mn;                     // B operand is MN-major (see header comment)
  // groups: independent problems sharing A, W, M, N and the destination, differing in their taps and in the
  // sub-pixel they write (the 4 output-parity classes of a stride-2 data gradient): one launch instead of four
  int ngroups, g_tap0[4], g_ntaps[4], g_sub_h0[4], g_sub_w0[4];
  // halo mode (3x3 stride-1 convs with few output channels, which are bound by the L2 -> smem traffic of re-reading
  // the activations once per tap): the pixel tile is 16 rows x 8 columns, its 18 x 10 halo is loaded ONCE per
  // 64-channel chunk, and the nine taps are nine UMMA descriptors into that one buffer (start shifted by
  // (dh+1)*10 + (dw+1) rows, 8-row groups 10 rows = 1280 bytes apart).  One pipeline stage = halo + the 9 weight tiles.
  int halo;                     // 0 off; 1: on.  (2: base_offset = swizzle phase of the start address -- WRONG on
                                // sm_100a, kept for the record: the hardware derives the phase from the absolute
                                // shared-memory address, so shifted starts need base_offset 0; checked in
                                // tools/gpu_halo_check.py)
  int a_bytes, b_bytes;         // bytes of A and of B in one stage
  int nacc, acc_cols;           // TMEM accumulators in flight: 2 x 256 columns, or 4 x 128 when bn <= 128 (narrow tiles are
                                // bound by the MMA -> epilogue -> MMA round trip per accumulator, not by the tensor pipe)
  int w_bytes;                  // halo mode with resident weights: all 9 x (cin/64) weight tiles live in shared memory for
                                // the whole kernel (loaded once per CTA), the stages hold only the activation halos
  float* stats;                 // optional fp32 [2][cout_per_sub]: per-channel sum and sum of squares of the values as
                                // stored (after bias / activation, rounded to the output type), over all valid pixels --
                                // the BatchNorm batch statistics of the layer that follows, taken from the epilogue
                                // instead of re-reading the tensor (vae-gan-v2.py:172-177: Conv -> BN -> ReLU)
  int dbg;                      // development only (VG_FPROP_DBG): 1 = the epilogue skips its work (what bounds the MMA side?)
  int4 taps[kMaxTaps];          // {c_base, dw, sh, dh}
  int wk[kMaxTaps];             // first weight column of each tap
};

// Flush one warp's per-lane column statistics (8 floats, see the epilogue) of N tile `n_t` into p.stats.
VG_DEVICE void stats_flush(const FpropParams& p, float (&st)[8], int n_t, int lane, int c_first, int c_step, int esz) {
  const int c2 = p.cout_per_sub;
  if (esz == 2) {
#pragma unroll
    for (int slot = 0; slot < 2; ++slot) {
      const int col = n_t * p.bn + c_first + slot * c_step + 2 * lane;
      if (c_first + slot * c_step < p.bn && col < p.n_gemm) {      // n_gemm % 8 == 0 on this path: col + 1 is valid too
        const int ch = col % c2;
        atomicAdd(p.stats + ch, st[slot * 4 + 0]);
        atomicAdd(p.stats + ch + 1, st[slot * 4 + 1]);
        atomicAdd(p.stats + c2 + ch, st[slot * 4 + 2]);
        atomicAdd(p.stats + c2 + ch + 1, st[slot * 4 + 3]);
      }
    }
  } else {
#pragma unroll
    for (int slot = 0; slot < 4; ++slot) {
      const int col = n_t * p.bn + c_first + slot * c_step + lane;
      if (c_first + slot * c_step < p.bn && col < p.n_gemm) {
        const int ch = col % c2;
        atomicAdd(p.stats + ch, st[slot * 2 + 0]);
        atomicAdd(p.stats + c2 + ch, st[slot * 2 + 1]);
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) st[i] = 0.f;
}

// Per-tile bookkeeping, kept cheap: the narrow layers have short tiles (K = 576: ~1800 cycles of MMAs), and the generic
// decode (three 32-bit divisions plus two 64-bit ones for the split-K range) cost ~1300 cycles per tile in EACH warp role --
// 40 % of such a tile in the MMA warp.  One group / one N tile / no split-K, the common case, needs no division here.
struct TileId { int grp, n_t, m_t, split; };
VG_DEVICE TileId decode_tile(const FpropParams& p, int tile, int tiles_per_group, int m_tiles) {
  TileId t;
  int tl = tile;
  if (p.ngroups == 1) t.grp = 0;
  else { t.grp = tile / tiles_per_group; tl = tile - t.grp * tiles_per_group; }
  int rest = tl;
  if (p.n_tiles == 1) t.n_t = 0;
  else { rest = tl / p.n_tiles; t.n_t = tl - rest * p.n_tiles; }
  if (p.ksplit == 1) { t.m_t = rest; t.split = 0; }
  else { t.split = rest / m_tiles; t.m_t = rest - t.split * m_tiles; }
  return t;
}
VG_DEVICE void k_range(int ksteps, int split, int ksplit, int& k_begin, int& k_end) {
  if (ksplit == 1) { k_begin = 0; k_end = ksteps; return; }
  // ksteps * ksplit stays far below 2^32 (ksteps <= 96 taps x 4096 channel chunks, ksplit <= #SMs)
  k_begin = static_cast<int>(static_cast<unsigned>(ksteps) * static_cast<unsigned>(split) / static_cast<unsigned>(ksplit));
  k_end = static_cast<int>(static_cast<unsigned>(ksteps) * static_cast<unsigned>(split + 1) / static_cast<unsigned>(ksplit));
}

// k2 = true: CTA-pair variant (cta_group::2; plain K-major mode with 256-wide N tiles only).  A cluster of two CTAs
// computes a 256-pixel x 256-column tile with ONE stream of M = 256 MMAs issued by the leader: each CTA stages its own
// 128-pixel activation box and HALF of the weight tile (128 columns), i.e. 32 KB per stage instead of 48 KB -- six
// pipeline stages instead of four in the same shared memory, and half the weight bytes from L2 per CTA -- and drains
// its own 128 TMEM lanes through the unchanged epilogue.
template <bool k2>
__global__ void __launch_bounds__(kFpropThreads, 1)
conv_fprop_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                  const __grid_constant__ FpropParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // carve: [stages][A 16 KB][B bn*128 B] then barriers
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const uint32_t a_bytes = static_cast<uint32_t>(p.a_bytes);
  const uint32_t b_bytes = static_cast<uint32_t>(p.b_bytes);
  const uint32_t stage_bytes = a_bytes + b_bytes;
  uint8_t* wres = smem + static_cast<size_t>(p.stages) * stage_bytes;            // resident weights (w_bytes, may be 0)
  uint64_t* bars = reinterpret_cast<uint64_t*>(wres + p.w_bytes);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + p.stages;
  uint64_t* tmem_full = bars + 2 * p.stages;
  uint64_t* tmem_empty = tmem_full + 4;
  uint64_t* w_bar = tmem_empty + 4;
  uint32_t* tmem_base_slot = reinterpret_cast<uint32_t*>(w_bar + 1);
  uint8_t* epi_smem = reinterpret_cast<uint8_t*>(bars) + 256;

  const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);      // broadcast: provably warp-uniform for the compiler
  const int lane = threadIdx.x & 31;
  // 0 = leader of the pair (issues the MMAs); clusters are (2, 1, 1), so the rank is the parity of blockIdx.x (uniform)
  const uint32_t rank = k2 ? (blockIdx.x & 1u) : 0u;

  if (threadIdx.x == 0) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
    for (int i = 0; i < p.stages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 4; ++i) {
      mbar_init(&tmem_full[i], 1);
      mbar_init(&tmem_empty[i], (p.nacc == 4 ? 128 : 256) * (k2 ? 2 : 1));   // pair: the leader's barrier collects both epilogues
    }
    mbar_init(w_bar, 1);
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

  // pair variant (one group, no split-K): work item = (pair of adjacent pixel tiles, N tile); this CTA's pixel tile is
  // m_t = 2 * unit + rank, expressed below as a virtual tile index in the single-CTA enumeration (n_t fastest); an odd
  // tile count leaves the last pair's second CTA with a tile beyond the grid: its loads are zero-filled, its rows masked
  const int m_tiles_real = p.tiles_n * p.tiles_h * p.tiles_w;
  const int m_tiles = k2 ? 2 * ((m_tiles_real + 1) / 2) : m_tiles_real;
  const int tiles_per_group = m_tiles * p.n_tiles * p.ksplit;
  const int total_tiles = tiles_per_group * p.ngroups;
  const int cchunks_all = p.cin / kBK;
  // single CTA: tiles blockIdx.x, + gridDim.x, ...   pair: units (blockIdx.x / 2), + gridDim.x / 2, ... each unit being the
  // two consecutive virtual tiles {2u * n_tiles + n_t + rank * n_tiles}
  const int worker = k2 ? static_cast<int>(blockIdx.x >> 1) : static_cast<int>(blockIdx.x);
  const int nworkers = k2 ? static_cast<int>(gridDim.x >> 1) : static_cast<int>(gridDim.x);
  const int total_work = k2 ? total_tiles / 2 : total_tiles;
  auto tile_of = [&](int work) -> int {
    if (!k2) return work;
    if (p.n_tiles == 1) return 2 * work + static_cast<int>(rank);
    const int unit = work / p.n_tiles, n_t = work - unit * p.n_tiles;
    return (2 * unit + static_cast<int>(rank)) * p.n_tiles + n_t;
  };

  // Producer: the whole warp walks the K loop with uniform control flow, and ONE elected lane issues the copies of a
  // stage (2 in the plain mode, up to 5 with an MN-major B operand, up to 10 in the halo mode) one after the other with
  // warp-uniform coordinates -- a few uniform-datapath instructions per copy.  (Issued from inside a per-lane region every
  // copy sits in an ELECT / R2UR waterfall loop of ~20 instructions; see the MMA issuer below.)
  if (warp == 0) {
    // ===================== TMA producer =====================
    const bool leader = elect_one();
    const int b_boxes = p.b_mn ? p.bn / 64 : 1;
    int stage = 0;
    uint32_t phase = 0;
    if (p.w_bytes > 0 && worker < total_work) {
      // resident weights: tile (cc, tap) at wres + (cc*9 + tap) * bn*128; one barrier for all of them
      if (leader) mbar_arrive_expect_tx(w_bar, p.w_bytes);
      for (int cc = 0; cc < cchunks_all; ++cc) {
#pragma unroll
        for (int t = 0; t < 9; ++t)
          if (leader) tma_load_2d(wres + static_cast<size_t>(cc * 9 + t) * (p.bn * 128), &tmap_b, w_bar, p.wk[t] + cc * kBK, 0);
      }
    }
    for (int work = worker; work < total_work; work += nworkers) {
      const int tile = tile_of(work);
      const TileId tid = decode_tile(p, tile, tiles_per_group, m_tiles);
      const int grp = tid.grp, n_t = tid.n_t, m_t = tid.m_t, split = tid.split;
      const int q1 = m_t / p.tiles_w, tw_i = m_t - q1 * p.tiles_w;
      const int tn_i = q1 / p.tiles_h, th_i = q1 - tn_i * p.tiles_h;
      const int ow0 = tw_i * p.tw, oh0 = th_i * p.th, n0 = tn_i * p.tn;
      const int cchunks = cchunks_all;
      const int ksteps = p.g_ntaps[grp] * cchunks;
      int k_begin, k_end;
      k_range(ksteps, split, p.ksplit, k_begin, k_end);
      if (p.halo) {
        // one stage per 64-channel chunk: the halo and (unless they are resident) the nine weight tiles
        for (int cc = 0; cc < cchunks; ++cc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (leader) mbar_arrive_expect_tx(&full_bar[stage], kHaloRows * 128 + (p.w_bytes > 0 ? 0 : 9 * p.bn * 128));
          uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
          if (leader) tma_load_5d(sa, &tmap_a, &full_bar[stage], p.taps[0].x + cc * kBK, ow0 - 1, 0, oh0 - 1, n0);
          if (p.w_bytes == 0) {
#pragma unroll
            for (int t = 0; t < 9; ++t)
              if (leader) tma_load_2d(sa + a_bytes + t * (p.bn * 128), &tmap_b, &full_bar[stage], p.wk[t] + cc * kBK, n_t * p.bn);
          }
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        continue;
      }
      int tap = p.g_tap0[grp], cc = 0;
      if (k_begin > 0) { tap += k_begin / cchunks; cc = k_begin % cchunks; }
      for (int k = k_begin; k < k_end; ++k) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (leader) {
          if (!k2) mbar_arrive_expect_tx(&full_bar[stage], stage_bytes);
          else if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * stage_bytes);      // both CTAs' boxes land on the leader's barrier
        }
        uint8_t* sa = smem + static_cast<size_t>(stage) * stage_bytes;
        uint8_t* sb = sa + a_bytes;
        const int4 t = p.taps[tap];
        const int wk = p.wk[tap];
        if (k2) {
          if (leader) tma_load_5d_2cta(sa, &tmap_a, &full_bar[stage], t.x + cc * kBK, ow0 + t.y, t.z, oh0 + t.w, n0);
          if (!p.b_mn) {
            if (leader) tma_load_2d_2cta(sb, &tmap_b, &full_bar[stage], wk + cc * kBK, n_t * p.bn + static_cast<int>(rank) * (p.bn / 2));
          } else {
            // MN-major operand of a pair: this CTA stages ITS half of the N columns = bn / 128 boxes of {64 N, 64 K}
            // (two for 256-wide tiles, one for 128-wide ones)
            const int half_boxes = p.bn / 128;
#pragma unroll
            for (int i = 0; i < 2; ++i)
              if (i < half_boxes) {
                if (leader)
                  tma_load_2d_2cta(sb + i * (64 * 128), &tmap_b, &full_bar[stage],
                                   wk + n_t * p.bn + (static_cast<int>(rank) * half_boxes + i) * 64, cc * kBK);
              }
          }
        } else {
          if (leader) tma_load_5d(sa, &tmap_a, &full_bar[stage], t.x + cc * kBK, ow0 + t.y, t.z, oh0 + t.w, n0);
          if (!p.b_mn) {
            if (leader) tma_load_2d(sb, &tmap_b, &full_bar[stage], wk + cc * kBK, n_t * p.bn);
          } else {
            // MN-major boxes: N columns [n_t*bn + 64*i, +64) of this tap, K rows [cc*64, +64)
#pragma unroll
            for (int i = 0; i < 4; ++i)
              if (i < b_boxes) {
                if (leader) tma_load_2d(sb + i * (64 * 128), &tmap_b, &full_bar[stage], wk + n_t * p.bn + i * 64, cc * kBK);
              }
          }
        }
        if (++cc == cchunks) { cc = 0; ++tap; }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1 && rank == 0) {
    // ===================== MMA issuer (leader CTA only in the pair variant) =====================
    // The WHOLE warp runs this loop (waits included) so that every address and descriptor is warp-uniform and lives in
    // uniform registers; only the tcgen05 instructions are predicated on one elected lane (see elect_one()).  Issued from
    // inside an `if (lane == 0)` region the same loop cost ~135 cycles per MMA, more than a 128 x 64 x 16 (32 cycles) or
    // 128 x 128 x 16 (64 cycles) MMA occupies the tensor pipe: the narrow layers ran at 0.25-0.5 of the tensor peak because
    // of the issue loop, not because of the tensor pipe (tools/mma_issue_bench.cu: 50 / 64 / 128 cycles per MMA at N = 64 /
    // 128 / 256 once the issue loop is lean, shifted halo views and MN-major operands included).
    const bool leader = elect_one();
    const uint32_t idesc = umma_idesc_bf16(k2 ? 2 * kBM : kBM, p.bn, 0, p.b_mn);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (p.w_bytes > 0 && worker < total_work) {
      mbar_wait(w_bar, 0);
      tc_fence_after();
    }
    // descriptor halves: the high word (SBO, version, swizzle) is constant per operand kind; the low word is the start
    // address (16-byte units) | LBO << 16
    const uint32_t a_hi = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, p.halo ? kHaloW * 128 : 1024) >> 32);
    const uint64_t b_proto = p.b_mn ? umma_smem_desc_sw128(0, 64 * 128, 1024) : umma_smem_desc_sw128(0, 16, 1024);
    const uint32_t b_hi = static_cast<uint32_t>(b_proto >> 32);
    const uint32_t a_lo0 = static_cast<uint32_t>(umma_smem_desc_sw128(0, 16, 1024));          // LBO field only
    const uint32_t b_lo0 = static_cast<uint32_t>(b_proto);
    const uint32_t b_step16 = p.b_mn ? (kUmmaK * 128) >> 4 : (kUmmaK * 2) >> 4;
    uint32_t halo_a16[9];      // start of each tap's view inside the halo buffer, in 16-byte units
#pragma unroll
    for (int t = 0; t < 9; ++t) halo_a16[t] = static_cast<uint32_t>(((p.taps[t].w + 1) * kHaloW + (p.taps[t].y + 1)) * 128) >> 4;
    const uint32_t b_tap16 = static_cast<uint32_t>(p.bn * 128) >> 4;
    long long dbg_t[4] = {0, 0, 0, 0};      // development (VG_FPROP_DBG=4): cycles in tmem_empty wait, full wait, issue; tiles
    const long long dbg_start = clock64();
    for (int work = worker; work < total_work; work += nworkers) {
      const int tile = tile_of(work);
      int grp = 0, split = 0;
      if (p.ngroups > 1 || p.ksplit > 1) {
        const TileId tid = decode_tile(p, tile, tiles_per_group, m_tiles);
        grp = tid.grp; split = tid.split;
      }
      const int ksteps = p.g_ntaps[grp] * cchunks_all;
      int k_begin, k_end;
      k_range(ksteps, split, p.ksplit, k_begin, k_end);
      const long long c0 = p.dbg == 4 ? clock64() : 0;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      if (p.dbg == 4) dbg_t[0] += clock64() - c0;
      const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * p.acc_cols);
      if (p.halo) {
        // the descriptors of one stage differ only in their start-address field: one base per operand plus the
        // (precomputed, 16-byte unit) tap and K offsets
        for (int cc = 0; cc < cchunks_all; ++cc) {
          const long long c1 = p.dbg == 4 ? clock64() : 0;
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const long long c2 = p.dbg == 4 ? clock64() : 0;
          if (p.dbg == 4) dbg_t[1] += c2 - c1;
          const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
          const uint32_t sb = p.w_bytes > 0 ? smem_u32(wres) + static_cast<uint32_t>(cc * 9 * p.bn * 128) : sa + a_bytes;
          const uint32_t a_lo = a_lo0 + (sa >> 4), b_lo = b_lo0 + (sb >> 4);
#pragma unroll
          for (int t = 0; t < 9; ++t) {
#pragma unroll
            for (int j = 0; j < kBK / kUmmaK; ++j)
              if (leader)
                umma_bf16_lohi(d_tmem, a_lo + halo_a16[t] + 2 * j, a_hi, b_lo + t * b_tap16 + 2 * j, b_hi, idesc,
                               (cc > 0 || t > 0 || j > 0) ? 1u : 0u);
          }
          if (leader) umma_commit(&empty_bar[stage]);
          if (p.dbg == 4) dbg_t[2] += clock64() - c2;
          if (++stage == p.stages) { stage = 0; phase ^= 1; }
        }
        if (leader) umma_commit(&tmem_full[acc]);
        if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
        if (p.dbg == 4) ++dbg_t[3];
        continue;
      }
      for (int k = k_begin; k < k_end; ++k) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * stage_bytes);
        const uint32_t sb = sa + a_bytes;
        // MN-major B: 16 K rows = 2 groups of 8 rows (SBO = 1024 B apart); 64-column N groups one box (8 KB) apart
        const uint32_t a_lo = a_lo0 + (sa >> 4), b_lo = b_lo0 + (sb >> 4);
#pragma unroll
        for (int j = 0; j < kBK / kUmmaK; ++j) {
          const uint32_t accum = (k > k_begin || j > 0) ? 1u : 0u;
          if (leader) {
            if (k2) umma_bf16_2cta_lohi(d_tmem, a_lo + 2 * j, a_hi, b_lo + j * b_step16, b_hi, idesc, accum);
            else umma_bf16_lohi(d_tmem, a_lo + 2 * j, a_hi, b_lo + j * b_step16, b_hi, idesc, accum);
          }
        }
        // frees the smem slot (in both CTAs of a pair) once these MMAs have read it
        if (leader) { if (k2) umma_commit_2cta(&empty_bar[stage]); else umma_commit(&empty_bar[stage]); }
        if (++stage == p.stages) { stage = 0; phase ^= 1; }
      }
      // accumulator complete -> epilogue (of both CTAs)
      if (leader) { if (k2) umma_commit_2cta(&tmem_full[acc]); else umma_commit(&tmem_full[acc]); }
      if (++acc == p.nacc) { acc = 0; acc_phase ^= 1; }
    }
    if (p.dbg == 4 && leader && blockIdx.x == 0)
      printf("fprop dbg: tiles %lld total %lld cyc | tmem_empty wait %lld  full wait %lld  issue %lld\n", dbg_t[3],
             clock64() - dbg_start, dbg_t[0], dbg_t[1], dbg_t[2]);
    __syncwarp();
  } else if (warp >= 4) {
    // ===================== epilogue (8 warps) =====================
    // TMEM -> registers (thread = one pixel row, 64 bf16 / 32 fp32 columns = 128 bytes per round) -> bias/activation
    // -> XOR-swizzled per-warp staging tile in shared memory -> coalesced 16-byte global stores (8 lanes cover 128
    // contiguous bytes of one output row).  Two warps share each TMEM lane quadrant and split the column chunks, so
    // every SM sub-partition has two epilogue warps to hide each other's latencies.
    const int quad = warp & 3;                 // TMEM lane quadrant this warp may access
    const int colhalf = (warp - 4) >> 2;       // which alternate column chunks this warp drains
    const int row = quad * 32 + lane;          // GEMM row inside the tile == pixel inside the tile
    const int r_w = row % p.tw;
    const int r_h = (row / p.tw) % p.th;
    const int r_n = row / (p.tw * p.th);
    uint8_t* stg = epi_smem + (warp - 4) * (32 * 128);
    float* bias_win = reinterpret_cast<float*>(epi_smem + kEpiStageBytes) + (warp - 4) * 32;
    const int esz = p.out_kind == 0 ? 2 : 4;
    const int chunk_cols = p.out_kind == 0 ? 64 : 32;
    const bool plain = (p.bias == nullptr) && (p.act == 0);
    // Narrow tiles (bn <= 128, four accumulators): the two warp groups (warps 4-7 / 8-11) drain ALTERNATE tiles, each
    // the whole tile width, so two epilogues are in flight -- their per-tile latency chain (TMEM load -> staging ->
    // stores), not their instruction count, is what bounds the 64-channel layers.  Wide tiles: both groups work on
    // the same tile and split its column chunks.
    const bool split = p.nacc == 4;
    const int c_first = split ? 0 : colhalf * chunk_cols, c_step = split ? chunk_cols : 2 * chunk_cols;
    // fused BatchNorm statistics: per-lane partial sums of the columns this lane owns in each chunk slot of the tile
    // (bf16 output: 2 slots x 2 columns x {sum, sum of squares}; fp32 output: 4 slots x 1 column), kept in registers
    // across the tiles of this CTA as long as they belong to the same N tile, then flushed with one atomic each
    float st[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) st[i] = 0.f;
    int st_nt = -1;
    int seq = 0;
    for (int work = worker; work < total_work; work += nworkers, ++seq) {
      const int tile = tile_of(work);
      if (split && (seq & 1) != colhalf) continue;
      const int acc = seq % p.nacc;
      const uint32_t acc_phase = static_cast<uint32_t>(seq / p.nacc) & 1u;
      const TileId tid = decode_tile(p, tile, tiles_per_group, m_tiles);
      const int grp = tid.grp, n_t = tid.n_t, m_t = tid.m_t;
      const int sub_h0 = p.g_sub_h0[grp], sub_w0 = p.g_sub_w0[grp];
      const int q1 = m_t / p.tiles_w, tw_i = m_t - q1 * p.tiles_w;
      const int tn_i = q1 / p.tiles_h, th_i = q1 - tn_i * p.tiles_h;
      const int ow = tw_i * p.tw + r_w, oh = th_i * p.th + r_h, n = tn_i * p.tn + r_n;
      const bool row_ok = (ow < p.m_w) && (oh < p.m_h) && (n < p.m_n);
      const long long my_base =
          row_ok ? ((static_cast<long long>(n) * p.out_h + oh * p.su_h) * p.out_w + ow * p.su_w) * p.out_ld : -1;
      const uint32_t vmask = __ballot_sync(0xffffffffu, row_ok);
      if (p.stats != nullptr && n_t != st_nt) {
        if (st_nt >= 0) stats_flush(p, st, st_nt, lane, c_first, c_step, esz);
        st_nt = n_t;
      }

      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t t_row = tmem_base + (static_cast<uint32_t>(quad * 32) << 16) + static_cast<uint32_t>(acc * p.acc_cols);
      if (p.dbg == 1) {
      } else if (p.vec_ok && p.out_kind != 2) {
        for (int c = c_first; c < p.bn; c += c_step) {
          const int ng0 = n_t * p.bn + c;        // first GEMM column of this chunk
          if (ng0 >= p.n_gemm) break;
          uint32_t r[64];
          tmem_ld_32x32(t_row + c, *reinterpret_cast<uint32_t(*)[32]>(&r[0]));
          if (esz == 2) tmem_ld_32x32(t_row + c + 32, *reinterpret_cast<uint32_t(*)[32]>(&r[32]));
          tmem_ld_wait();
          if (!plain) {
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) {
              if (hh == 1 && esz != 2) break;
              const int ngh = ng0 + hh * 32;
              const int ch0 = ngh % p.cout_per_sub;
              const int nvalid = min(32, p.n_gemm - ngh);
              // the 32 bias values of this column window: one coalesced load per warp, then broadcast float4 reads
              // (a per-element __ldg made bias-carrying layers LSU-bound: 64 loads per thread per 64 columns)
              if (p.bias != nullptr) {
                __syncwarp();
                bias_win[lane] = lane < nvalid ? __ldg(p.bias + ch0 + lane) : 0.f;
                __syncwarp();
              }
#pragma unroll
              for (int j4 = 0; j4 < 8; ++j4) {
                float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
                if (p.bias != nullptr) b4 = *reinterpret_cast<const float4*>(bias_win + 4 * j4);
                const float bb[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  float x = __uint_as_float(r[hh * 32 + 4 * j4 + e]) + bb[e];
                  if (p.act == 1) x = fmaxf(x, 0.f);
                  else if (p.act == 2) x = x > 0.f ? x : 0.2f * x;
                  r[hh * 32 + 4 * j4 + e] = __float_as_uint(x);
                }
              }
            }
          }
          if (p.dbg == 2) { if (r[0] == 0x7fc12345u) reinterpret_cast<uint32_t*>(p.out)[0] = r[63]; continue; }      // development: no staging, no stores
          // stage: 16-byte segment s of row `lane` goes to slot (s ^ (lane & 7)) -> conflict-free both ways
          uint8_t* dst = stg + lane * 128;
          if (esz == 2) {
#pragma unroll
            for (int sgm = 0; sgm < 8; ++sgm) {
              uint4 pk;
              pk.x = pack_bf16x2(__uint_as_float(r[sgm * 8 + 0]), __uint_as_float(r[sgm * 8 + 1]));
              pk.y = pack_bf16x2(__uint_as_float(r[sgm * 8 + 2]), __uint_as_float(r[sgm * 8 + 3]));
              pk.z = pack_bf16x2(__uint_as_float(r[sgm * 8 + 4]), __uint_as_float(r[sgm * 8 + 5]));
              pk.w = pack_bf16x2(__uint_as_float(r[sgm * 8 + 6]), __uint_as_float(r[sgm * 8 + 7]));
              *reinterpret_cast<uint4*>(dst + ((sgm ^ (lane & 7)) << 4)) = pk;
            }
          } else {
#pragma unroll
            for (int sgm = 0; sgm < 8; ++sgm)
              *reinterpret_cast<uint4*>(dst + ((sgm ^ (lane & 7)) << 4)) =
                  make_uint4(r[sgm * 4], r[sgm * 4 + 1], r[sgm * 4 + 2], r[sgm * 4 + 3]);
          }
          __syncwarp();
          const int seg = lane & 7, rsub = lane >> 3;
          const int half = esz == 2 ? (seg >> 2) : 0;
          const int ngh = ng0 + half * 32;
          const int valid_bytes = min(128, (p.n_gemm - ng0) * esz);
          const int sub = ngh / p.cout_per_sub;
          const int ch0 = ngh - sub * p.cout_per_sub;
          const long long delta =
              (static_cast<long long>(sub_h0 + sub / p.su_w) * p.out_w + (sub_w0 + sub % p.su_w)) * p.out_ld +
              p.out_coff + ch0 + (esz == 2 ? (seg & 3) * 8 : seg * 4);
          uint8_t* gout = reinterpret_cast<uint8_t*>(p.out) + delta * esz;
          const bool seg_ok = seg * 16 < valid_bytes;
#pragma unroll
          for (int it = 0; it < 8; ++it) {
            const int rr = it * 4 + rsub;
            const long long base = __shfl_sync(0xffffffffu, my_base, rr);
            if (seg_ok && base >= 0) {
              const uint4 v = *reinterpret_cast<const uint4*>(stg + rr * 128 + ((seg ^ (rr & 7)) << 4));
              if (p.dbg != 3 || v.x == 0x7fc12345u) *reinterpret_cast<uint4*>(gout + base * esz) = v;      // dbg 3: staging, no global stores
            }
          }
          if (p.stats != nullptr) {
            // column sums over the valid rows, read back from the staged tile (= the values as stored): lane owns the
            // 4-byte word `lane` of every 128-byte row -- two bf16 columns or one fp32 column; conflict-free
            const int wseg = lane >> 2, woff = (lane & 3) * 4;
            float a0 = 0.f, a1 = 0.f, q0 = 0.f, q1 = 0.f;
#pragma unroll 8
            for (int rr = 0; rr < 32; ++rr) {
              if ((vmask >> rr) & 1u) {
                const uint32_t wv = *reinterpret_cast<const uint32_t*>(stg + rr * 128 + ((wseg ^ (rr & 7)) << 4) + woff);
                if (esz == 2) {
                  const float2 f = unpack_bf16x2(wv);
                  a0 += f.x; a1 += f.y; q0 = fmaf(f.x, f.x, q0); q1 = fmaf(f.y, f.y, q1);
                } else {
                  const float f = __uint_as_float(wv);
                  a0 += f; q0 = fmaf(f, f, q0);
                }
              }
            }
            const int slot = (c - c_first) / c_step;
            if (esz == 2) {
#pragma unroll
              for (int j = 0; j < 2; ++j) {
                const bool on = slot == j;
                st[j * 4 + 0] += on ? a0 : 0.f; st[j * 4 + 1] += on ? a1 : 0.f;
                st[j * 4 + 2] += on ? q0 : 0.f; st[j * 4 + 3] += on ? q1 : 0.f;
              }
            } else {
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const bool on = slot == j;
                st[j * 2 + 0] += on ? a0 : 0.f; st[j * 2 + 1] += on ? q0 : 0.f;
              }
            }
          }
          __syncwarp();
        }
      } else if (split || colhalf == 0) {
        // generic path (odd alignments, split-K atomics): each thread stores its own row
        for (int c = 0; c < p.bn; c += 32) {
          uint32_t r[32];
          tmem_ld_32x32(t_row + c, r);
          tmem_ld_wait();
          const int ng0 = n_t * p.bn + c;
          if (row_ok && ng0 < p.n_gemm) {
            const int sub = ng0 / p.cout_per_sub;   // uniform over the chunk when cout_per_sub % 32 == 0
            const int ch0 = ng0 - sub * p.cout_per_sub;
            const int dh = sub_h0 + sub / p.su_w, dw = sub_w0 + sub % p.su_w;
            const long long off = my_base + (static_cast<long long>(dh) * p.out_w + dw) * p.out_ld + p.out_coff + ch0;
            const int nvalid = min(32, p.n_gemm - ng0);
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              if (j < nvalid) {
                float x = __uint_as_float(r[j]);
                if (p.bias != nullptr && (p.out_kind != 2 || tid.split == 0)) x += __ldg(p.bias + ch0 + j);   // split-K: bias once
                if (p.act == 1) x = fmaxf(x, 0.f);
                else if (p.act == 2) x = x > 0.f ? x : 0.2f * x;
                if (p.out_kind == 0) reinterpret_cast<__nv_bfloat16*>(p.out)[off + j] = __float2bfloat16(x);
                else if (p.out_kind == 1) reinterpret_cast<float*>(p.out)[off + j] = x;
                else atomicAdd(reinterpret_cast<float*>(p.out) + off + j, x);
              }
            }
          }
        }
      }
      tc_fence_before();
      if (k2) mbar_arrive_cluster(&tmem_empty[acc], 0); else mbar_arrive(&tmem_empty[acc]);
    }
    if (p.stats != nullptr && st_nt >= 0) stats_flush(p, st, st_nt, lane, c_first, c_step, esz);
  }

  tc_fence_before();
  if (k2) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (k2) tmem_dealloc_2cta(tmem_base, 512); else tmem_dealloc(tmem_base, 512);
  }
}

static int g_fprop_pairs = 1;      // CTA-pair variant on / off (vg_set_cta_pairs)

static int pick_pixel_tile(int m_n, int m_h, int m_w, int total, int* tn, int* th, int* tw) {
  // tw = largest power of two <= min(m_w rounded up to pow2, total); th likewise; tn = rest
  int w = 1;
  while (w < m_w && w < total) w <<= 1;
  int h = 1;
  while (h < m_h && w * h < total) h <<= 1;
  int n = total / (w * h);
  *tw = w; *th = h; *tn = n;
  (void)m_n;
  return 0;
}

}  // namespace vg

using namespace vg;

extern "C" int vg_set_fprop_cta_pairs(int on) {
  g_fprop_pairs = on;
  return 0;
}

extern "C" int vg_conv_fprop(const VgConvFprop* d, void* stream_) {
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  VG_CHECK(d != nullptr, -1, "vg_conv_fprop: null descriptor");
  VG_CHECK(d->cin > 0 && d->cin % 64 == 0, -1, "vg_conv_fprop: cin (%d) must be a positive multiple of 64", d->cin);
  VG_CHECK(d->num_taps >= 1 && d->num_taps <= kMaxTaps, -1, "vg_conv_fprop: num_taps %d out of range", d->num_taps);
  VG_CHECK(d->x_stride == 1 || d->x_stride == 2, -1, "vg_conv_fprop: x_stride must be 1 or 2");
  VG_CHECK(d->x_h % d->x_stride == 0 && d->x_w % d->x_stride == 0, -1, "vg_conv_fprop: H,W must divide by the stride");
  VG_CHECK(d->x_ld % 8 == 0 && d->w_ld % 8 == 0, -1, "vg_conv_fprop: leading dims must be multiples of 8");
  VG_CHECK(d->n_gemm >= 1, -1, "vg_conv_fprop: n_gemm");
  VG_CHECK(d->su_h >= 1 && d->su_w >= 1 && d->cout_per_sub >= 1, -1, "vg_conv_fprop: bad scatter");
  VG_CHECK(d->su_h * d->su_w == 1 || d->cout_per_sub % 32 == 0, -1,
           "vg_conv_fprop: pixel-shuffle epilogue needs cout_per_sub %% 32 == 0");
  VG_CHECK((reinterpret_cast<uintptr_t>(d->x) & 15) == 0 && (reinterpret_cast<uintptr_t>(d->w) & 15) == 0, -1,
           "vg_conv_fprop: operands must be 16-byte aligned");

  FpropParams p;
  memset(&p, 0, sizeof(p));
  p.m_n = d->m_n; p.m_h = d->m_h; p.m_w = d->m_w;
  // halo mode: a 3x3 stride-1 tap set (every offset of [-1,1]^2 once, one channel base), K-major weights, no groups,
  // no split-K; chosen for narrow N and many pixels, where the 9-fold re-read of A through L2 is the bound
  int halo = 0;
  if (d->halo_mode >= 0 && d->num_taps == 9 && d->x_stride == 1 && !d->b_mn_major && d->num_groups <= 1 && d->ksplit <= 1 &&
      d->out_kind != 2 && d->m_h >= 16 && d->m_w >= 8) {
    unsigned seen = 0;
    bool ok = true;
    for (int i = 0; i < 9; ++i) {
      const int dw = d->taps[i][1], dh = d->taps[i][3];
      if (dw < -1 || dw > 1 || dh < -1 || dh > 1 || d->taps[i][2] != 0 || d->taps[i][0] != d->taps[0][0]) { ok = false; break; }
      seen |= 1u << ((dh + 1) * 3 + (dw + 1));
    }
    const long long pixels = static_cast<long long>(d->m_n) * d->m_h * d->m_w;
    if (ok && seen == 0x1FFu && (d->halo_mode > 0 || (d->n_gemm <= 64 && pixels >= 65536))) halo = d->halo_mode > 0 ? d->halo_mode : 1;
  }
  p.halo = halo;
  if (halo) { p.tn = 1; p.th = 16; p.tw = 8; }
  else pick_pixel_tile(p.m_n, p.m_h, p.m_w, kBM, &p.tn, &p.th, &p.tw);
  p.tiles_n = cdiv(p.m_n, p.tn); p.tiles_h = cdiv(p.m_h, p.th); p.tiles_w = cdiv(p.m_w, p.tw);
  p.cin = d->cin; p.num_taps = d->num_taps; p.n_gemm = d->n_gemm;
  const int m_tiles = p.tiles_n * p.tiles_h * p.tiles_w;
  const int sms = conv_sms();
  // N tile: the widest that still leaves enough tiles to fill the machine
  int bn = 256;
  if (d->n_gemm <= 64) bn = 64;
  else if (d->n_gemm <= 128) bn = 128;
  else if (m_tiles * cdiv(d->n_gemm, 256) < sms && d->n_gemm % 256 != 0) bn = 128;
  else if (m_tiles * cdiv(d->n_gemm, 256) < sms / 2) bn = 128;
  if (d->force_bn == 64 || d->force_bn == 128 || d->force_bn == 256) bn = d->force_bn;
  if (halo) bn = 64;      // one stage holds the halo and all nine 64-wide weight tiles
  p.bn = bn;
  p.n_tiles = cdiv(d->n_gemm, bn);
  p.ksteps = d->num_taps * (d->cin / kBK);
  int ksplit = d->ksplit;
  if (ksplit <= 0) {
    ksplit = 1;
    const int tiles = m_tiles * p.n_tiles;
    if (d->out_kind == 2 && tiles < sms) ksplit = min(p.ksteps, max(1, sms / tiles));
  }
  VG_CHECK(ksplit == 1 || d->out_kind == 2, -1, "vg_conv_fprop: split-K needs the fp32 atomic output kind");
  VG_CHECK(ksplit <= p.ksteps, -1, "vg_conv_fprop: ksplit %d > k steps %d", ksplit, p.ksteps);
  VG_CHECK(ksplit == 1 || d->act == 0, -1, "vg_conv_fprop: an activation cannot be fused into a split-K launch");
  p.ksplit = ksplit;
  p.nacc = bn <= 128 ? 4 : 2;
  p.acc_cols = 512 / p.nacc;
  // CTA pairs: the plain mode (K-major or MN-major weights) with 256-wide N tiles, one group, no split-K, and enough pixel tiles that pairing
  // them does not leave SMs without work
  static const int mn_pairs = getenv("VG_FPROP_MN_PAIRS") ? atoi(getenv("VG_FPROP_MN_PAIRS")) : 1;
  static const int pair_min_bn = getenv("VG_FPROP_PAIR_MIN_BN") ? atoi(getenv("VG_FPROP_PAIR_MIN_BN")) : 128;
  const bool pair = g_fprop_pairs != 0 && !halo && (!d->b_mn_major || mn_pairs) && bn >= pair_min_bn && bn >= 128 && ksplit == 1 &&
                    d->num_groups <= 1 &&
                    (m_tiles / 2) * p.n_tiles >= sms / 2 && sms % 2 == 0;
  p.a_bytes = halo ? kHaloBytes : kBM * kBK * 2;
  p.b_bytes = halo ? 9 * bn * kBK * 2 : (pair ? bn / 2 : bn) * kBK * 2;
  const int smem_budget = 227 * 1024 - 1024 - 256 - kEpiBytes;
  if (halo && p.n_tiles == 1) {
    // weights resident when they leave room for at least two halo stages
    const int wb = 9 * (d->cin / kBK) * bn * kBK * 2;
    if (smem_budget - wb >= 2 * p.a_bytes) { p.w_bytes = wb; p.b_bytes = 0; }
  }
  const int stage_bytes = p.a_bytes + p.b_bytes;
  p.stages = min(8, (smem_budget - p.w_bytes) / stage_bytes);
  p.out = d->out; p.out_kind = d->out_kind;
  p.out_h = d->out_h; p.out_w = d->out_w; p.out_ld = d->out_ld; p.out_coff = d->out_coff;
  p.su_h = d->su_h; p.su_w = d->su_w; p.sub_h0 = d->sub_h0; p.sub_w0 = d->sub_w0; p.cout_per_sub = d->cout_per_sub;
  p.bias = d->bias; p.act = d->act;
  p.b_mn = d->b_mn_major ? 1 : 0;
  p.ngroups = d->num_groups > 1 ? d->num_groups : 1;
  VG_CHECK(p.ngroups <= 4, -1, "vg_conv_fprop: at most 4 groups");
  if (p.ngroups == 1) {
    p.g_tap0[0] = 0; p.g_ntaps[0] = d->num_taps; p.g_sub_h0[0] = d->sub_h0; p.g_sub_w0[0] = d->sub_w0;
  } else {
    int t0 = 0;
    for (int g = 0; g < p.ngroups; ++g) {
      VG_CHECK(d->group_ntaps[g] >= 1, -1, "vg_conv_fprop: group %d has no taps", g);
      p.g_tap0[g] = t0; p.g_ntaps[g] = d->group_ntaps[g];
      p.g_sub_h0[g] = d->group_sub[g][0]; p.g_sub_w0[g] = d->group_sub[g][1];
      t0 += d->group_ntaps[g];
    }
    VG_CHECK(t0 == d->num_taps, -1, "vg_conv_fprop: group tap counts (%d) do not add up to num_taps (%d)", t0, d->num_taps);
    VG_CHECK(ksplit == 1, -1, "vg_conv_fprop: groups cannot be combined with split-K");
  }
  VG_CHECK(!p.b_mn || d->w_rows >= 1, -1, "vg_conv_fprop: b_mn_major needs w_rows (number of K rows of the weight matrix)");
  {
    const int esz = d->out_kind == 0 ? 2 : 4;
    p.vec_ok = ((reinterpret_cast<uintptr_t>(d->out) & 15) == 0) && ((static_cast<long long>(d->out_ld) * esz) % 16 == 0) &&
               ((d->out_coff * esz) % 16 == 0) && ((d->cout_per_sub * esz) % 16 == 0) && (d->n_gemm % 8 == 0) &&
               (d->su_h * d->su_w == 1 || d->cout_per_sub % 32 == 0);
  }
  static const int dbg_env = getenv("VG_FPROP_DBG") ? atoi(getenv("VG_FPROP_DBG")) : 0;
  p.dbg = dbg_env;
  p.stats = d->stats;
  if (d->stats != nullptr) {
    VG_CHECK(p.vec_ok && d->out_kind != 2 && ksplit == 1, -1,
             "vg_conv_fprop: fused statistics need the vectorised store path (aligned destination, n_gemm %% 8 == 0), "
             "a plain (non-atomic) output and no split-K");
    VG_CHECK(d->cout_per_sub % 2 == 0, -1, "vg_conv_fprop: fused statistics need an even channel count");
    VG_CUDA(cudaMemsetAsync(d->stats, 0, sizeof(float) * 2 * d->cout_per_sub, stream));
  }
  for (int i = 0; i < d->num_taps; ++i) {
    p.taps[i] = make_int4(d->taps[i][0], d->taps[i][1], d->taps[i][2], d->taps[i][3]);
    p.wk[i] = d->use_wk ? d->wk[i] : i * d->cin;
    VG_CHECK(d->taps[i][2] >= 0 && d->taps[i][2] < d->x_stride, -1, "vg_conv_fprop: tap %d row parity out of range", i);
  }

  const int s = d->x_stride;
  CUtensorMap tmap_a, tmap_b;
  {
    const uint64_t ld = static_cast<uint64_t>(d->x_ld), W = d->x_w, H = d->x_h;
    uint64_t dims[5] = {ld * s, W / s, static_cast<uint64_t>(s), H / s, static_cast<uint64_t>(d->x_n)};
    uint64_t strides[5] = {1, ld * s, W * ld, W * ld * s, H * W * ld};
    uint32_t box[5] = {kBK, static_cast<uint32_t>(p.tw), 1, static_cast<uint32_t>(p.th), static_cast<uint32_t>(p.tn)};
    if (halo) { box[1] = kHaloW; box[3] = kHaloH; }
    int rc = encode_tmap_bf16(&tmap_a, d->x, 5, dims, strides, box);
    if (rc) return rc;
  }
  {
    uint64_t dims[2] = {static_cast<uint64_t>(d->w_ld), static_cast<uint64_t>(p.b_mn ? d->w_rows : d->n_gemm)};
    uint64_t strides[2] = {1, static_cast<uint64_t>(d->w_ld)};
    uint32_t box[2] = {kBK, static_cast<uint32_t>(p.b_mn ? 64 : (pair ? bn / 2 : bn))};
    int rc = encode_tmap_bf16(&tmap_b, d->w, 2, dims, strides, box);
    if (rc) return rc;
  }
  const size_t smem = static_cast<size_t>(p.stages) * stage_bytes + p.w_bytes + 1024 /*align*/ + 256 /*barriers*/ + kEpiBytes;
  static bool attr_set[64] = {false};      // per device: function attributes belong to the device's context
  const int dev = current_device();
  if (!attr_set[dev]) {
    VG_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    VG_CUDA(cudaFuncSetAttribute(conv_fprop_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    attr_set[dev] = true;
  }
  const int total_tiles = m_tiles * p.n_tiles * p.ksplit * p.ngroups;
  if (pair) {
    const int units = ((m_tiles + 1) / 2) * p.n_tiles;
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    cfg.gridDim = dim3(static_cast<unsigned>(2 * min(units, sms / 2)), 1, 1);
    cfg.blockDim = dim3(kFpropThreads, 1, 1);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VG_CUDA(cudaLaunchKernelEx(&cfg, conv_fprop_kernel<true>, tmap_a, tmap_b, p));
  } else {
    const int grid = min(total_tiles, sms);
    conv_fprop_kernel<false><<<grid, kFpropThreads, smem, stream>>>(tmap_a, tmap_b, p);
  }
  VG_LAUNCH_OK();
  return 0;
}
