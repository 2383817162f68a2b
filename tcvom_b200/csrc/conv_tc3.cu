// tcgen05 convolution kernel for NARROW layers (Cin in {8, 32, 64}, Cout = 32) -- "v3".
//
// Narrow layers at full / half resolution have almost no tensor work (K <= 576); with the v2 kernel they
// were bound by per-stage pipeline hand-shakes and by fetching the activation halo three times (once per
// horizontal tap offset, because 64B-swizzled operand tiles can only be shifted by whole 512 B atoms).
// This kernel stores the activation tile UN-swizzled in the canonical K-major "interleave" layout
// (8-row x 16-byte core matrices): rows (pixels) are 16 B apart inside a core matrix, 8-row groups are SBO
// apart, K-adjacent core matrices LBO apart -- so ANY 16-byte-aligned start address is a legal operand.
//   * one TMA box {8 ch, TW+2 px, 2*TH+2 rows, 1, Cin/8 chunks} per work item and K chunk serves all 9 taps:
//     tap (dy,dx) of accumulator i is the descriptor start ((i*TH+dy+1)*(TW+2) + dx+1)*16 B, SBO = one box
//     row, LBO = one 8-channel chunk plane;
//   * Cin = 8 (the network input: 3 image + 3 trimap + 2 pad channels): the K=16 MMA step spans TWO
//     neighbouring pixels (LBO = 16 B), i.e. the horizontal taps are folded into K: per vertical tap dy the
//     weights are packed as K = 32 = (dx=-1, 0, +1, zero) x 8 channels and two MMA steps cover them;
//   * tile = 16 rows x 8 px per accumulator (a core matrix = 8 consecutive pixels of one image row), two
//     accumulators stacked vertically, persistent CTAs, TMEM double buffering, weights resident in smem.
// Weights keep the 64B-swizzled layout (A and B descriptors carry independent layout types).
#include "tc_epilogue.cuh"

namespace tcv {

constexpr int V3_TH = 16, V3_TW = 8;
constexpr int V3_MAXT = 9;

struct V3Params {
  int gh, gw, tiles_x, tiles_y, total_work;
  int nchunk;          // Cin / 8 (1 for the folded Cin = 8 mode)
  int fold;            // Cin == 8: horizontal taps folded into K
  int kc_iters;        // K blocks of 32 per tap: nchunk / 4 (1 in fold mode)
  int ntaps, dy[V3_MAXT], dx[V3_MAXT], wtap[V3_MAXT];
  int dy_min, dx_min, box_w, box_rows;
  int b_resident;
  int a_slots, b_slots;   // runtime shared-memory carve-up (resident weight tiles, activation ring)
  uint32_t idesc;    // M=128, N=BN
  uint32_t idesc2;   // M=128, N=2*BN: A_hi x [B_hi ; B_lo] in one MMA
  EpiParams epi;
};

template <int BN>
struct V3Cfg {
  static constexpr int B_SLOT_BYTES = 2 * BN * 32 * 2;       // hi + lo, one tap x one 32-wide K block
  static constexpr int MAX_B_SLOTS = 18;                      // 9 taps x up to 2 K blocks stay resident (72 KB)
  static constexpr int A_SLOT_BYTES = 45056;                  // hi + lo of one 10x34-pixel x 32-channel box
  static constexpr int MAX_A_SLOTS = 3;
  static constexpr int STAGE_HALF = 2 * 2 * 128 * BN * 2;     // output staging: 2 accumulators x hi/lo x 128 px x BN ch
  static constexpr int STAGE_BYTES = 2 * STAGE_HALF;          // double-buffered (the epilogue bounds these layers)
  static constexpr int BAR_BYTES = 512;
  static constexpr int TMEM_COLS = 8 * BN;                    // 2 buffers x 2 accumulators x (hi.hi | hi.lo) halves
  static int smem_bytes(int a_slots, int b_slots) {
    return a_slots * A_SLOT_BYTES + b_slots * B_SLOT_BYTES + STAGE_BYTES + BAR_BYTES + 1024;
  }
};

// un-swizzled K-major descriptor: LBO = byte distance between K-adjacent core matrices, SBO = between
// consecutive 8-row groups
__device__ __forceinline__ uint64_t smem_desc_ns(uint32_t addr, uint32_t lbo, uint32_t sbo) {
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)((lbo >> 4) & 0x3FFF) << 16) |
         ((uint64_t)((sbo >> 4) & 0x3FFF) << 32) | (1ull << 46);
}

template <int BN>
__global__ void __launch_bounds__(320, 1) conv_tc3_kernel(const __grid_constant__ CUtensorMap mapA_hi,
                                                          const __grid_constant__ CUtensorMap mapA_lo,
                                                          const __grid_constant__ CUtensorMap mapB_hi,
                                                          const __grid_constant__ CUtensorMap mapB_lo,
                                                          const __grid_constant__ CUtensorMap mapY_hi,
                                                          const __grid_constant__ CUtensorMap mapY_lo,
                                                          const __grid_constant__ V3Params p) {
  using Cfg = V3Cfg<BN>;
  const int SA = p.a_slots, SB = p.b_slots;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + SA * Cfg::A_SLOT_BYTES;
  const uint32_t stage_base = b_base + SB * Cfg::B_SLOT_BYTES;      // 1 KB aligned (slot sizes are 1 KB multiples)
  const uint32_t bar_base = stage_base + Cfg::STAGE_BYTES;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (SA + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (2 * SA + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (2 * SA + SB + s); };
  auto accFull = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + a); };
  auto accEmpty = [&](int a) { return bar_base + 8u * (2 * SA + 2 * SB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * SA + 2 * SB + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accFull(a), 1); mbar_init(accEmpty(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_lo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();

  const int box_px = p.box_w * p.box_rows;                  // pixels per chunk plane
  const uint32_t chunk_bytes = (uint32_t)box_px * 16u;      // one 8-channel chunk plane of the box
  const uint32_t a_plane_bytes = chunk_bytes * (uint32_t)(p.fold ? 1 : 4);   // one 32-channel K block (or the folded box)
  const uint32_t a_lo_off = (a_plane_bytes + 127u) & ~127u;  // TMA destinations must be 128-byte aligned
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  auto decode = [&](int work, int& img, int& h0, int& w0) {
    const int t = work % tiles_per_img;
    img = work / tiles_per_img;
    const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
    h0 = ty * 2 * V3_TH;
    w0 = tx * V3_TW;
  };

  if (warp == 0) {
    // ================================ TMA producer ================================
    int ia = 0, ib = 0;
    for (int work = blockIdx.x; work < p.total_work; work += gridDim.x) {
      int img, h0, w0;
      decode(work, img, h0, w0);
      const bool load_b = !p.b_resident || work == (int)blockIdx.x;
      for (int kc = 0; kc < p.kc_iters; ++kc, ++ia) {
        const int sa = ia % SA;
        mbar_wait(emptyA(sa), ((uint32_t)(ia / SA) & 1u) ^ 1u);
        const uint32_t adst = smem_base + sa * Cfg::A_SLOT_BYTES;
        if (elect_one()) {
          mbar_expect_tx(fullA(sa), 2 * a_plane_bytes);
          // box {8 ch, box_w px, box_rows, 1 img, 4 chunks (1 when folded)} at (0, w0+dx_min, h0+dy_min, img, 4*kc)
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
              ::"r"(adst), "l"(&mapA_hi), "r"(fullA(sa)), "r"(0), "r"(w0 + p.dx_min), "r"(h0 + p.dy_min), "r"(img),
              "r"(4 * kc)
              : "memory");
          asm volatile(
              "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6, %7}], [%2];"
              ::"r"(adst + a_lo_off), "l"(&mapA_lo), "r"(fullA(sa)), "r"(0), "r"(w0 + p.dx_min),
              "r"(h0 + p.dy_min), "r"(img), "r"(4 * kc)
              : "memory");
        }
        __syncwarp();
        if (!load_b) continue;
        for (int t = 0; t < p.ntaps; ++t, ++ib) {
          const int sb = ib % SB;
          mbar_wait(emptyB(sb), ((uint32_t)(ib / SB) & 1u) ^ 1u);
          const uint32_t bdst = b_base + sb * Cfg::B_SLOT_BYTES;
          if (elect_one()) {
            mbar_expect_tx(fullB(sb), Cfg::B_SLOT_BYTES);
            tma_load_3d(bdst, &mapB_hi, fullB(sb), kc * 32, 0, p.wtap[t]);
            tma_load_3d(bdst + Cfg::B_SLOT_BYTES / 2, &mapB_lo, fullB(sb), kc * 32, 0, p.wtap[t]);
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    int ia = 0, ib = 0, iw = 0;
    const uint32_t row_bytes = (uint32_t)p.box_w * 16u;                       // SBO: next image row of the box
    const uint32_t lbo = p.fold ? 16u : chunk_bytes;                          // K-adjacent core matrix
    // Resident weights (every layer shape of the network): ONE issue block per K block -- all taps back to back, operand
    // descriptors formed from low words (address >> 4) with constant high words.  The per-tap version below needed an
    // elect / descriptor round of ~0.3 us for 0.1 us of tensor work per tap (measured: 165 us with the epilogue's memory
    // operations switched off, 100 us with the MMAs off as well, for 32 -> 32 @ 544 x 960 x 3).
    const bool fast = p.b_resident && !(p.epi.dbg & (64 | 256));
    const uint32_t a_lo16 = a_lo_off >> 4, th16 = (uint32_t)(V3_TH * p.box_w);      // (pixel = 16 bytes)
    const uint32_t kstep16 = p.fold ? 2u : (2u * chunk_bytes) >> 4;
    const uint32_t ns_hi = ((row_bytes >> 4) & 0x3FFFu) | (1u << 14);
    const uint32_t ns_lbo = ((lbo >> 4) & 0x3FFFu) << 16;
    for (int work = blockIdx.x; work < p.total_work; work += gridDim.x, ++iw) {
      const int buf = iw & 1;
      mbar_wait(accEmpty(buf), ((uint32_t)(iw >> 1) & 1u) ^ 1u);
      tc_fence_after();
      const uint32_t d0 = tmem_base + (uint32_t)(buf * 4 * BN), d1 = d0 + (uint32_t)(2 * BN);
      bool first = true;
      int il = 0;
      if (fast) {
        if (iw == 0) {                       // the weights arrive once
          for (int s = 0; s < p.ntaps * p.kc_iters; ++s) mbar_wait(fullB(s), 0u);
        }
        for (int kc = 0; kc < p.kc_iters; ++kc, ++ia) {
          const int sa = ia % SA;
          const uint32_t a32 = (((smem_base + sa * Cfg::A_SLOT_BYTES) & 0x3FFFFu) >> 4) | ns_lbo;
          const uint32_t b32 = smem_desc_lo(b_base + (uint32_t)(kc * p.ntaps) * Cfg::B_SLOT_BYTES);
          mbar_wait(fullA(sa), (uint32_t)(ia / SA) & 1u);
          tc_fence_after();
          if (elect_one()) {
            for (int t = 0; t < p.ntaps; ++t) {
              const uint32_t px0 = (uint32_t)((p.dy[t] - p.dy_min) * p.box_w + (p.fold ? 0 : p.dx[t] - p.dx_min));
              const uint32_t ah32 = a32 + px0, bt32 = b32 + (uint32_t)t * (Cfg::B_SLOT_BYTES >> 4);
#pragma unroll
              for (int ks = 0; ks < ((p.epi.dbg & 1) ? 0 : 2); ++ks) {
                const uint32_t a0 = ah32 + ks * kstep16;
                const uint64_t ah0 = ((uint64_t)ns_hi << 32) | a0, ah1 = ((uint64_t)ns_hi << 32) | (a0 + th16);
                const uint64_t al0 = ((uint64_t)ns_hi << 32) | (a0 + a_lo16), al1 = ((uint64_t)ns_hi << 32) | (a0 + a_lo16 + th16);
                const uint64_t bh = smem_desc_join<32>(bt32 + 2 * ks);
                const uint32_t acc = (kc == 0 && t == 0 && ks == 0) ? 0u : 1u;
                tc_mma(d0, ah0, bh, p.idesc2, acc);
                tc_mma(d1, ah1, bh, p.idesc2, acc);
                tc_mma(d0, al0, bh, p.idesc, 1u);
                tc_mma(d1, al1, bh, p.idesc, 1u);
              }
            }
            tc_commit(emptyA(sa));
            if (kc == p.kc_iters - 1) tc_commit(accFull(buf));
          }
          __syncwarp();
        }
        continue;
      }
      for (int kc = 0; kc < p.kc_iters; ++kc, ++ia) {
        const int sa = ia % SA;
        mbar_wait(fullA(sa), (uint32_t)(ia / SA) & 1u);
        tc_fence_after();
        const uint32_t a_hi = smem_base + sa * Cfg::A_SLOT_BYTES, a_lo = a_hi + a_lo_off;
        for (int t = 0; t < p.ntaps; ++t, ++ib, ++il) {
          const int sb = p.b_resident ? il : ib % SB;
          if (!p.b_resident || iw == 0) {
            mbar_wait(fullB(sb), p.b_resident ? 0u : ((uint32_t)(ib / SB) & 1u));
            tc_fence_after();
          }
          const uint32_t b_hi = b_base + sb * Cfg::B_SLOT_BYTES, b_lo = b_hi + Cfg::B_SLOT_BYTES / 2;
          if (elect_one()) {
            // pixel offset of the tap inside the box (fold mode: dx is folded into K, start at the box's left edge)
            const uint32_t px0 = (uint32_t)((p.dy[t] - p.dy_min) * p.box_w + (p.fold ? 0 : p.dx[t] - p.dx_min));
            const uint32_t off0 = px0 * 16u, off1 = off0 + (uint32_t)(V3_TH * p.box_w) * 16u;
#pragma unroll
            for (int ks = 0; ks < ((p.epi.dbg & 1) ? 0 : 2); ++ks) {
              // K step ks covers chunks 2ks, 2ks+1 (fold mode: pixels 2ks, 2ks+1 of the row)
              const uint32_t kofs = p.fold ? (uint32_t)ks * 32u : (uint32_t)ks * 2u * chunk_bytes;
              const uint64_t ah0 = smem_desc_ns(a_hi + off0 + kofs, lbo, row_bytes);
              const uint64_t al0 = smem_desc_ns(a_lo + off0 + kofs, lbo, row_bytes);
              const uint64_t ah1 = smem_desc_ns(a_hi + off1 + kofs, lbo, row_bytes);
              const uint64_t al1 = smem_desc_ns(a_lo + off1 + kofs, lbo, row_bytes);
              const uint64_t bh = smem_desc<32>(b_hi + ks * 32);
              const uint32_t acc = (first && ks == 0) ? 0u : 1u;
              // The hi and lo weight planes are adjacent in the slot, so [B_hi ; B_lo] is ONE 2*BN-row operand:
              // A_hi is read from shared memory once for both products (the narrow layers are bound by the
              // operand reads of N=32 MMAs).  Columns [0,BN) get A_hi.B_hi (+ A_lo.B_hi below), [BN,2BN) A_hi.B_lo;
              // the epilogue adds the two halves.
              tc_mma(d0, ah0, bh, p.idesc2, acc);
              tc_mma(d1, ah1, bh, p.idesc2, acc);
              if (!(p.epi.dbg & 64)) {   // (measurement switch: skip the A_lo term)
                tc_mma(d0, al0, bh, p.idesc, 1u);
                tc_mma(d1, al1, bh, p.idesc, 1u);
              }
            }
            if (!p.b_resident) tc_commit(emptyB(sb));
            if (t == p.ntaps - 1) {
              tc_commit(emptyA(sa));
              if (kc == p.kc_iters - 1) tc_commit(accFull(buf));
            }
          }
          __syncwarp();
          first = false;
        }
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    // TMEM -> registers -> fused epilogue -> 64B-swizzled smem tile -> TMA store: the output leaves the SM
    // as whole 64-byte pixel rows written by the copy engine instead of 32 scattered 16-byte sectors per
    // store instruction (the epilogue, not the MMA, bounded the 32-channel layers).
    const int e = warp - 2;
    const int i = e >> 2;
    const int q = warp & 3;
    const int r = q * 32 + lane;
    StoreCtx stc;
    stc.stage_hi = stage_base + (uint32_t)i * (2u * 128u * BN * 2u);
    stc.stage_lo = stc.stage_hi + 128u * BN * 2u;
    stc.row = r;
    stc.bar = 1 + i;
    stc.issuer = (e & 3) == 0 && lane == 0;
    stc.map_hi = &mapY_hi;
    stc.map_lo = &mapY_lo;
    stc.alt_bytes = Cfg::STAGE_HALF;
    const bool issuer = stc.issuer;
    int iw = 0;
    for (int work = blockIdx.x; work < p.total_work; work += gridDim.x, ++iw) {
      int img, h0, w0;
      decode(work, img, h0, w0);
      const int buf = iw & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 4 * BN + i * 2 * BN);
      const int ty = r / V3_TW, tx = r - ty * V3_TW;
      const int gy = h0 + i * V3_TH + ty, gx = w0 + tx;
      stc.cx = w0;
      stc.cy = h0 + i * V3_TH;
      conv_epilogue<BN>(p.epi, taddr, gy < p.gh && gx < p.gw, img, gy, gx, 0, accFull(buf), (uint32_t)(iw >> 1) & 1u,
                        accEmpty(buf), lane, stc, /*split_halves=*/true);
    }
    if (issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
int conv2d_tc3_supported(const tcv_conv_desc& d) {
  if (d.stride != 1 || d.pad_mode != TCV_PAD_ZERO) return 0;
  if (d.cout != 32 || !d.y || d.y_f32) return 0;
  if (d.x_img_stride != (long long)d.ih * d.iw * d.cin) return 0;
  if (d.ntaps > V3_MAXT) return 0;
  for (int t = 0; t < d.ntaps; ++t)
    if (d.dy[t] < -1 || d.dy[t] > 1 || d.dx[t] < -1 || d.dx[t] > 1) return 0;
  if (d.cin == 8) {
    if (!d.w_tc_fold || d.ntaps != 9) return 0;
    for (int t = 0; t < 9; ++t)
      if (d.dy[t] != t / 3 - 1 || d.dx[t] != t % 3 - 1 || d.wtap[t] != t) return 0;
    return 1;
  }
  if (!d.w_tc) return 0;
  return (d.cin == 32 || d.cin == 64) ? 1 : 0;
}

template <int BN>
static int conv_tc3_bn(const tcv_conv_desc& d, cudaStream_t st) {
  using Cfg = V3Cfg<BN>;
  V3Params p;
  memset(&p, 0, sizeof(p));
  p.fold = d.cin == 8 ? 1 : 0;
  p.nchunk = p.fold ? 1 : d.cin / 8;
  p.kc_iters = p.fold ? 1 : d.cin / 32;
  int dymin = 1, dymax = -1, dxmin = 1, dxmax = -1;
  for (int t = 0; t < d.ntaps; ++t) {
    dymin = d.dy[t] < dymin ? d.dy[t] : dymin; dymax = d.dy[t] > dymax ? d.dy[t] : dymax;
    dxmin = d.dx[t] < dxmin ? d.dx[t] : dxmin; dxmax = d.dx[t] > dxmax ? d.dx[t] : dxmax;
  }
  if (p.fold) {
    // one MMA row of K=32 per vertical tap: (dx=-1, 0, +1, zero-weight pixel) x 8 channels
    p.ntaps = 3;
    for (int t = 0; t < 3; ++t) { p.dy[t] = t - 1; p.dx[t] = -1; p.wtap[t] = t; }
    dymin = -1; dymax = 1; dxmin = -1; dxmax = 2;
  } else {
    p.ntaps = d.ntaps;
    for (int t = 0; t < d.ntaps; ++t) { p.dy[t] = d.dy[t]; p.dx[t] = d.dx[t]; p.wtap[t] = d.wtap[t]; }
  }
  p.dy_min = dymin; p.dx_min = dxmin;
  p.box_w = V3_TW + (dxmax - dxmin);
  p.box_rows = 2 * V3_TH + (dymax - dymin);
  const int plane_bytes = p.box_w * p.box_rows * 16 * (p.fold ? 1 : 4);
  if (((plane_bytes + 127) & ~127) + plane_bytes > Cfg::A_SLOT_BYTES) return fail(TCV_ERR_UNSUPPORTED, "conv_tc3: activation box too large");
  p.gh = d.gh; p.gw = d.gw;
  p.tiles_x = (d.gw + V3_TW - 1) / V3_TW;
  p.tiles_y = (d.gh + 2 * V3_TH - 1) / (2 * V3_TH);
  p.total_work = p.tiles_x * p.tiles_y * d.n;
  p.b_slots = p.ntaps * p.kc_iters;                 // every weight tile stays resident
  if (p.b_slots > Cfg::MAX_B_SLOTS) return fail(TCV_ERR_UNSUPPORTED, "conv_tc3: too many weight tiles");
  p.b_resident = 1;
  p.a_slots = Cfg::MAX_A_SLOTS;
  while (p.a_slots > 2 && Cfg::smem_bytes(p.a_slots, p.b_slots) > 227 * 1024) --p.a_slots;
  p.idesc = instr_desc(BN, false);
  p.idesc2 = instr_desc(2 * BN, false);
  fill_epi(p.epi, d, g_debug_flags.load());

  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mY_hi, mY_lo;
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(d.x);
  {
    // output view over the compute grid: pixel (gy, gx) of image n lives at y[n][gy*oy_mul+oy_off][gx*ox_mul+ox_off]
    const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(d.y) + ((long long)d.oy_off * d.ow + d.ox_off) * d.cout;
    cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)d.gw, (cuuint64_t)d.gh, (cuuint64_t)d.n};
    cuuint64_t str[3] = {(cuuint64_t)d.ox_mul * d.cout * 2, (cuuint64_t)d.oy_mul * d.ow * d.cout * 2,
                         (cuuint64_t)d.oh * d.ow * d.cout * 2};
    cuuint32_t box[4] = {(cuuint32_t)BN, (cuuint32_t)V3_TW, (cuuint32_t)V3_TH, 1};
    int rc = make_map(&mY_hi, y, 4, dims, str, box, 32);
    if (rc) return rc;
    rc = make_map(&mY_lo, y + (long long)d.n * d.oh * d.ow * d.cout, 4, dims, str, box, 32);
    if (rc) return rc;
  }
  {
    // 5-D view of the NHWC tensor: {8 ch, W, H, N, Cin/8 chunks}; chunk stride 16 B
    EncodeTiledFn enc = get_encode();
    if (!enc) return fail(TCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[5] = {8, (cuuint64_t)d.iw, (cuuint64_t)d.ih, (cuuint64_t)d.n, (cuuint64_t)(d.cin / 8)};
    cuuint64_t str[4] = {(cuuint64_t)d.cin * 2, (cuuint64_t)d.iw * d.cin * 2, (cuuint64_t)d.x_img_stride * 2, 16};
    cuuint32_t box[5] = {8, (cuuint32_t)p.box_w, (cuuint32_t)p.box_rows, 1, (cuuint32_t)(p.fold ? 1 : 4)};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    for (int pl = 0; pl < 2; ++pl) {
      CUresult r = enc(pl ? &mA_lo : &mA_hi, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 5,
                       const_cast<__nv_bfloat16*>(a + (pl ? d.x_plane : 0)), dims, str, box, estr,
                       CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                       CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) return fail(TCV_ERR_CUDA, "conv_tc3: cuTensorMapEncodeTiled (activation view) failed (%d)", (int)r);
    }
  }
  {
    const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(p.fold ? d.w_tc_fold : d.w_tc);
    const int kdim = p.fold ? 32 : d.cin, wt = p.fold ? 3 : d.w_tc_taps;
    cuuint64_t dims[3] = {(cuuint64_t)kdim, (cuuint64_t)d.cout, (cuuint64_t)wt};
    cuuint64_t str[2] = {(cuuint64_t)kdim * 2, (cuuint64_t)d.cout * kdim * 2};
    cuuint32_t box[3] = {32, (cuuint32_t)BN, 1};
    int rc = make_map(&mB_hi, b, 3, dims, str, box, 32);
    if (rc) return rc;
    rc = make_map(&mB_lo, b + (long long)wt * d.cout * kdim, 3, dims, str, box, 32);
    if (rc) return rc;
  }
  auto kern = conv_tc3_kernel<BN>;
  const int smem = Cfg::smem_bytes(p.a_slots, p.b_slots);
  TCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  int dev = 0, sms = 0;
  TCV_CUDA(cudaGetDevice(&dev));
  TCV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.total_work < sms ? p.total_work : sms;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)grid, 1, 1);
  cfg.blockDim = dim3(320, 1, 1);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  cfg.numAttrs = tc_launch_attrs(attr, 1);
  TCV_CUDA(cudaLaunchKernelEx(&cfg, kern, mA_hi, mA_lo, mB_hi, mB_lo, mY_hi, mY_lo, p));
  return launched("conv_tc3_kernel");
}

int conv2d_tc3(const tcv_conv_desc& d, cudaStream_t st) { return conv_tc3_bn<32>(d, st); }

}  // namespace tcv
