// CTA-pair version of the persistent tcgen05 convolution kernel (conv_tc2.cu) for the wide layers (Cout >= 128).
//
// Why: with both operands in shared memory an M=128,N=128,K=16 MMA reads 8 KB per 64 cycles -- exactly the 128 B/clk a
// SM's shared memory delivers -- so in conv_tc2 the MMA stream and the TMA writes of the next stage (37 B/clk) serialise
// (measured: MMA alone 14.5 us, TMA alone 11.7 us, together 27 us per work item of a 128->128 layer).  A CTA pair
// (cluster of 2, tcgen05.mma.cta_group::2, M = 256) halves the B-operand traffic per SM:
//   * the pair owns the same 2*TH x TW pixel super-tile as one conv_tc2 CTA, but CTA rank r stages only ITS half
//     (TH + halo rows) of the activation box and HALF (BN/2 rows) of every weight tile;
//   * one MMA of M=256, N=BN (up to 256) covers both pixel tiles: per SM it reads 4 KB of A + BN*16 B of B per
//     BN/2 cycles (N=256: 64 B/clk), TMA writes ~30 B/clk;
//   * each CTA holds ONE accumulator [128 px x BN ch], double-buffered in TMEM (2 x BN columns); its 8 epilogue warps
//     split the columns in two halves (4 warps = 4 TMEM lane quarters each) and run the same fused epilogue
//     (tc_epilogue.cuh) with the TMA-store staging of conv_tc2.
// Barrier protocol as in gemm_tc2.cu: fullA/fullB live in the leader (its producer expects both CTAs' bytes, both CTAs'
// TMA loads complete_tx there), emptyA/emptyB/accFull per CTA (multicast commit), accEmpty in the leader (16 arrivals).
#include "tc_epilogue.cuh"

namespace tcv {

constexpr int VP_MAXG = 3;
constexpr int VP_MAXDY = 3;
constexpr int VP_A_ROWS = 192;   // rows (pixels) of the largest (TH + halo) x TW activation box

struct VPParams {
  int gh, gw, TH, TW, tiles_x, tiles_y, n_tiles_n, total_work;
  int kc_iters;
  int ngroups, group_dx[VP_MAXG], ndy[VP_MAXG], dy[VP_MAXG][VP_MAXDY], wtap[VP_MAXG][VP_MAXDY];
  int dy_min, box_rows;
  int tma_store;
  int dbg;           // measurement switches (tcv_set_debug_flags): 1 no MMA, 4 activations loaded once, 8 weights loaded once
  int b_resident;    // all weight tiles of a work item fit the B ring: staged by the first work item only
  uint32_t idesc;
  uint32_t idesc2;   // STACK: the N = 2*BN descriptor of the stacked MMA (idesc stays N = BN)
  EpiParams epi;
};

// STACK (BN = 64 only): the hi.hi and hi.lo products of a K step come from ONE MMA of N = 2*BN over the stacked operand
// [B_hi ; B_lo] (columns BN..2BN-1 of the accumulator are added in the epilogue), then A_lo.B_hi with N = BN: two reads of
// the 4 KB activation tile per K step instead of three -- a 64-channel layer is bound by exactly those reads (160 B/clk per
// SM with three MMAs of N = 64; 117 B/clk stacked).  In cta_group::2 the N rows of an operand are split between the CTAs, so
// the leader stages B_hi (all BN rows) and the peer B_lo at slot offset 0; the N = BN operand of the second MMA sits at
// slot offset 2*BN*64 B: rows 0..BN/2-1 of B_hi in the leader, rows BN/2..BN-1 in the peer.
// BK: K elements (channels) per stage = bytes per shared-memory row / 2: 32 -> SWIZZLE_64B, 64 -> SWIZZLE_128B.  BK = 64 halves
// the number of stages per work item: the per-stage cost of the issuing thread (barrier wait, descriptors, commits: ~0.26 us,
// measured with the MMAs switched off) exceeds the 0.2 us of tensor work of a BK = 32 stage of the <= 128-channel layers.
template <int BN, bool STACK = false, int BK_ = 32>
struct VPCfg {
  static constexpr int BK = BK_;
  static constexpr int A_SLOT_BYTES = 2 * VP_A_ROWS * BK * 2;   // hi + lo planes
  static constexpr int B_ROWS = BN / 2;                         // rows of a weight tile this CTA stages
  static constexpr int B_SLOT_BYTES = STACK ? 3 * B_ROWS * BK * 2 : 2 * B_ROWS * BK * 2;   // hi + lo (STACK: 2 + 1 blocks)
  static constexpr int ACC_COLS = STACK ? 2 * BN : BN;
  // narrow layers: all (<= 3) vertical taps of a horizontal-offset group share ONE weight stage (one full/empty barrier round
  // trip per 3 taps): a 64-channel stage is only 4 MMAs = 0.1 us of tensor work, the per-stage hand-shake cost more
  static constexpr int TPS = BN <= 64 ? VP_MAXDY : 1;
  static constexpr int B_TAP_BYTES = B_SLOT_BYTES;
  // activation ring depth: one box feeds only ndy x 2 (x 3) MMAs, i.e. 0.3 us of tensor work for a 64-channel layer against
  // ~1 us of TMA latency -- ncu shows those layers neither DRAM- (9 %), L2- (20 %) nor tensor-bound (23 %).
  // Measured (round 2): SIX / FOUR slots made every layer slower (64 -> 64: 149 -> 171 us, 128 -> 128: 93 -> 104 us), and so
  // did a 220 KB instead of a 200 KB shared-memory budget (256 -> 256 with one more weight slot: 91 -> 147 us): the
  // epilogue's residual / affine loads live in what the carve-out leaves to L1.  Three slots, <= VP_SMEM_BUDGET.
#ifndef VP_A_SLOTS
#define VP_A_SLOTS 3
#endif
#ifndef VP_SMEM_BUDGET
#define VP_SMEM_BUDGET (200 * 1024)
#endif
  static constexpr int A_SLOTS = BK == 64 ? 2 : VP_A_SLOTS;     // (a BK = 64 slot carries twice the K)
  static constexpr int A_BYTES = A_SLOTS * A_SLOT_BYTES;        // 72 / 96 KB
  static constexpr int STAGE_BYTES = 2 * 2 * 128 * 64;          // TMA-store staging: 2 column groups x hi/lo x 128 rows x 64 B
  // (weights that fit the ring stay resident -- b_resident; measured: no gain for the 64-channel layers, so the ring is small)
  static constexpr int B_STAGE_BYTES = TPS * B_TAP_BYTES;
  static constexpr int B_CAP = 8;
  static constexpr int B_SLOTS = (VP_SMEM_BUDGET - A_BYTES - STAGE_BYTES) / B_STAGE_BYTES > B_CAP
                                     ? B_CAP : (VP_SMEM_BUDGET - A_BYTES - STAGE_BYTES) / B_STAGE_BYTES;
  static constexpr int SMEM = A_BYTES + B_SLOTS * B_STAGE_BYTES + STAGE_BYTES + 1024 + 512;
  static constexpr int TMEM_COLS = 2 * ACC_COLS < 32 ? 32 : 2 * ACC_COLS;   // 2 buffers x one accumulator
};

__device__ __forceinline__ uint32_t vp_cluster_rank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void vp_cluster_sync() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t vp_mapa(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void vp_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void vp_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2,
                                          int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void vp_tma_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar_cluster, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar_cluster), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void vp_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void vp_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
               ::"r"(bar), "h"((uint16_t)3)
               : "memory");
}

template <int BN, bool STACK, int BK, bool STATS>
__global__ void __launch_bounds__(320, 1) conv_tc2p_kernel(const __grid_constant__ CUtensorMap mapA_hi,
                                                           const __grid_constant__ CUtensorMap mapA_lo,
                                                           const __grid_constant__ CUtensorMap mapB_hi,
                                                           const __grid_constant__ CUtensorMap mapB_lo,
                                                           const __grid_constant__ CUtensorMap mapY_hi,
                                                           const __grid_constant__ CUtensorMap mapY_lo,
                                                           const __grid_constant__ VPParams p) {
  using Cfg = VPCfg<BN, STACK, BK>;
  constexpr int SA = Cfg::A_SLOTS, SB = Cfg::B_SLOTS;
  constexpr int VP_BK = BK;
  constexpr int VP_A_SLOT_BYTES = Cfg::A_SLOT_BYTES;
  constexpr uint32_t BLK = Cfg::B_ROWS * VP_BK * 2;     // one plane of B_ROWS weight rows
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + Cfg::A_BYTES;
  const uint32_t stage_base = b_base + SB * Cfg::B_STAGE_BYTES;
  constexpr int TPS = Cfg::TPS;
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
  const uint32_t rank = vp_cluster_rank();
  const bool leader = rank == 0;
  const int pair = blockIdx.x >> 1, npairs = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < SA; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accFull(a), 1); mbar_init(accEmpty(a), 16); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB_lo) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot),
                 "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  vp_cluster_sync();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot_ptr;
  pdl_wait();        // everything above overlapped the previous kernel's tail; from here on its results are read

  const int a_plane_bytes = p.box_rows * p.TW * (VP_BK * 2);   // one plane of this CTA's A box
  const int tiles_per_img = p.tiles_x * p.tiles_y;

  // work item -> (image, origin of the pair's 2*TH x TW super-tile, first output channel)
  auto decode = [&](int work, int& img, int& h0, int& w0, int& n0) {
    const int nt = work % p.n_tiles_n;
    int r = work / p.n_tiles_n;
    const int t = r % tiles_per_img;
    img = r / tiles_per_img;
    const int ty = t / p.tiles_x, tx = t - ty * p.tiles_x;
    h0 = ty * 2 * p.TH;
    w0 = tx * p.TW;
    n0 = nt * BN;
  };

  if (warp == 0) {
    // ================================ TMA producer (every CTA) ================================
    int ia = 0, ib = 0;
    for (int work = pair; work < p.total_work; work += npairs) {
      int img, h0, w0, n0;
      decode(work, img, h0, w0, n0);
      const int hr = h0 + (int)rank * p.TH + p.dy_min;      // first box row of this CTA's pixel tile
      const int nb = n0 + (int)rank * Cfg::B_ROWS;          // this CTA's half of the weight tile
      const bool load_b = !p.b_resident || work == pair;    // resident weights: first work item only
      for (int kc = 0; kc < p.kc_iters; ++kc) {
        for (int g = 0; g < p.ngroups; ++g, ++ia) {
          const int sa = ia % SA;
          mbar_wait(emptyA(sa), ((uint32_t)(ia / SA) & 1u) ^ 1u);
          const uint32_t adst = smem_base + sa * VP_A_SLOT_BYTES;
          if (elect_one()) {
            if ((p.dbg & 4) && ia >= SA) {          // measurement switch: activations loaded once (results are garbage)
              if (leader) mbar_arrive(fullA(sa));
            } else {
              if (leader) mbar_expect_tx(fullA(sa), 4u * a_plane_bytes);
              const uint32_t fb = vp_mapa(fullA(sa), 0);
              vp_tma_4d(adst, &mapA_hi, fb, kc * VP_BK, w0 + p.group_dx[g], hr, img);
              vp_tma_4d(adst + a_plane_bytes, &mapA_lo, fb, kc * VP_BK, w0 + p.group_dx[g], hr, img);
            }
          }
          __syncwarp();
          if (!load_b) continue;
          for (int j0 = 0; j0 < p.ndy[g]; j0 += TPS, ++ib) {
            const int sb = ib % SB;          // resident mode: ib < SB, slot == stage index within the work item
            const int nt = p.ndy[g] - j0 < TPS ? p.ndy[g] - j0 : TPS;     // taps in this weight stage
            int nt0 = 1;
            mbar_wait(emptyB(sb), ((uint32_t)(ib / SB) & 1u) ^ 1u);
            if (elect_one()) {
              if ((p.dbg & 8) && ib >= SB) {        // measurement switch: weights loaded once
                if (leader) mbar_arrive(fullB(sb));
                nt0 = 0;
              } else if (leader) mbar_expect_tx(fullB(sb), 2u * (uint32_t)nt * Cfg::B_TAP_BYTES);
              const uint32_t fb = vp_mapa(fullB(sb), 0);
              for (int jj = 0; jj < (nt0 ? nt : 0); ++jj) {
                const int j = j0 + jj;
                const uint32_t bdst = b_base + sb * Cfg::B_STAGE_BYTES + jj * Cfg::B_TAP_BYTES;
                if constexpr (STACK) {
                  // leader: B_hi rows 0..BN-1 (stacked operand, its N rows 0..BN-1) + B_hi rows 0..BN/2-1 again;
                  // peer:   B_lo rows 0..BN-1 (N rows BN..2BN-1)                  + B_hi rows BN/2..BN-1
                  const CUtensorMap* m0 = leader ? &mapB_hi : &mapB_lo;
                  vp_tma_3d(bdst, m0, fb, kc * VP_BK, n0, p.wtap[g][j]);
                  vp_tma_3d(bdst + BLK, m0, fb, kc * VP_BK, n0 + Cfg::B_ROWS, p.wtap[g][j]);
                  vp_tma_3d(bdst + 2 * BLK, &mapB_hi, fb, kc * VP_BK, nb, p.wtap[g][j]);
                } else {
                  vp_tma_3d(bdst, &mapB_hi, fb, kc * VP_BK, nb, p.wtap[g][j]);
                  vp_tma_3d(bdst + Cfg::B_TAP_BYTES / 2, &mapB_lo, fb, kc * VP_BK, nb, p.wtap[g][j]);
                }
              }
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ================================
    // The issue path is latency-critical (measured: a handful of extra instructions per stage cost the 128-channel layers
    // 10 %): every operand descriptor of a stage is formed from precomputed low words BEFORE the stage's barrier is awaited,
    // so that the MMAs go out back to back the moment the data has landed.
    if (leader) {
      int ia = 0, ib = 0, iw = 0;
      const uint32_t plane16 = (uint32_t)a_plane_bytes >> 4;
      const uint32_t row16 = (uint32_t)(p.TW * (VP_BK * 2)) >> 4;          // one image row of the box, in 16-byte units
      for (int work = pair; work < p.total_work; work += npairs, ++iw) {
        const int buf = iw & 1;
        mbar_wait(accEmpty(buf), ((uint32_t)(iw >> 1) & 1u) ^ 1u);   // both CTAs' epilogues have drained this buffer
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)(buf * Cfg::ACC_COLS);
        uint32_t acc_first = 0u;
        int il = 0;   // weight stage index within this work item
        for (int kc = 0; kc < p.kc_iters; ++kc) {
          for (int g = 0; g < p.ngroups; ++g, ++ia) {
            const int sa = ia % SA;
            const uint32_t a_lo32 = smem_desc_lo(smem_base + sa * VP_A_SLOT_BYTES);
            const int ndy = p.ndy[g];
            const bool last_group = kc == p.kc_iters - 1 && g == p.ngroups - 1;
            mbar_wait(fullA(sa), (uint32_t)(ia / SA) & 1u);
            for (int j = 0; j < ndy; ++j) {
              const bool stage_first = (j % TPS) == 0, stage_last = ((j + 1) % TPS) == 0 || j == ndy - 1;
              const int sb = p.b_resident ? il : ib % SB;
              // rows of the accumulator for vertical tap dy start (dy - dy_min) image rows into the box (same offset in
              // the peer's shared memory, whose box starts TH image rows lower)
              const uint32_t ah32 = a_lo32 + (uint32_t)(p.dy[g][j] - p.dy_min) * row16, al32 = ah32 + plane16;
              const uint32_t bh32 = smem_desc_lo(b_base + sb * Cfg::B_STAGE_BYTES + (j % TPS) * Cfg::B_TAP_BYTES);
              if (stage_first && (!p.b_resident || iw == 0))   // resident weights: only the first work item waits for them
                mbar_wait(fullB(sb), p.b_resident ? 0u : ((uint32_t)(ib / SB) & 1u));
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int ks = 0; ks < ((p.dbg & 1) ? 0 : VP_BK / 16); ++ks) {
                  const uint64_t ah = smem_desc_join<VP_BK>(ah32 + 2 * ks), al = smem_desc_join<VP_BK>(al32 + 2 * ks);
                  if constexpr (STACK) {
                    const uint64_t bs = smem_desc_join<VP_BK>(bh32 + 2 * ks);
                    const uint64_t bh2 = smem_desc_join<VP_BK>(bh32 + ((2 * BLK) >> 4) + 2 * ks);
                    vp_mma(d, ah, bs, p.idesc2, acc_first | (uint32_t)ks);        // cols [0,BN): hi.hi ; [BN,2BN): hi.lo
                    vp_mma(d, al, bh2, p.idesc, 1u);                              // cols [0,BN) += lo.hi
                  } else {
                    const uint64_t bh = smem_desc_join<VP_BK>(bh32 + 2 * ks);
                    const uint64_t bl = smem_desc_join<VP_BK>(bh32 + (Cfg::B_TAP_BYTES >> 5) + 2 * ks);
                    vp_mma(d, ah, bh, p.idesc, acc_first | (uint32_t)ks);
                    vp_mma(d, ah, bl, p.idesc, 1u);
                    vp_mma(d, al, bh, p.idesc, 1u);
                  }
                }
                if (!p.b_resident && stage_last) vp_commit(emptyB(sb));
                if (j == ndy - 1) {
                  vp_commit(emptyA(sa));
                  if (last_group) vp_commit(accFull(buf));
                }
              }
              __syncwarp();
              acc_first = 1u;
              if (stage_last) { ++ib; ++il; }
            }
          }
        }
      }
    }
  } else {
    // ================================ epilogue (warps 2..9, every CTA) ================================
    constexpr int HN = BN / 2;           // columns per 4-warp group
    const int e = warp - 2;              // 0..7
    const int grp = e >> 2;              // column half of this CTA's accumulator
    const int q = warp & 3;              // TMEM lane quarter accessible to this warp
    const int r = q * 32 + lane;         // accumulator row = pixel of this CTA's TH x TW tile
    StoreCtx stc;
    if (p.tma_store) {
      stc.stage_hi = stage_base + (uint32_t)grp * (2u * 128u * 64u);
      stc.stage_lo = stc.stage_hi + 128u * 64u;
      stc.row = r;
      stc.bar = 1 + grp;
      stc.issuer = (e & 3) == 0 && lane == 0;
      stc.map_hi = &mapY_hi;
      stc.map_lo = &mapY_lo;
    }
    const uint32_t ae_leader = vp_mapa(accEmpty(0), 0);
    int iw = 0;
    for (int work = pair; work < p.total_work; work += npairs, ++iw) {
      int img, h0, w0, n0;
      decode(work, img, h0, w0, n0);
      const int buf = iw & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * Cfg::ACC_COLS + grp * HN);
      const int ty = r / p.TW, tx = r - ty * p.TW;
      const int gy = h0 + (int)rank * p.TH + ty, gx = w0 + tx;
      stc.cx = w0;
      stc.cy = h0 + (int)rank * p.TH;
      conv_epilogue<HN, BN, STATS>(p.epi, taddr, gy < p.gh && gx < p.gw, img, gy, gx, n0 + grp * HN, accFull(buf),
                            (uint32_t)(iw >> 1) & 1u, ae_leader + 8u * buf, lane, stc, /*split_halves=*/STACK,
                            /*remote_empty=*/true);
    }
    if (stc.issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  vp_cluster_sync();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
static bool vp_build_groups(const tcv_conv_desc& d, VPParams& p) {
  p.ngroups = 0;
  int dymin = 1 << 30, dymax = -(1 << 30);
  for (int t = 0; t < d.ntaps; ++t) {
    int g = -1;
    for (int k = 0; k < p.ngroups; ++k)
      if (p.group_dx[k] == d.dx[t]) g = k;
    if (g < 0) {
      if (p.ngroups == VP_MAXG) return false;
      g = p.ngroups++;
      p.group_dx[g] = d.dx[t];
      p.ndy[g] = 0;
    }
    if (p.ndy[g] == VP_MAXDY) return false;
    p.dy[g][p.ndy[g]] = d.dy[t];
    p.wtap[g][p.ndy[g]] = d.wtap[t];
    p.ndy[g]++;
    dymin = d.dy[t] < dymin ? d.dy[t] : dymin;
    dymax = d.dy[t] > dymax ? d.dy[t] : dymax;
  }
  p.dy_min = dymin;
  p.box_rows = dymax - dymin;   // halo rows; TH added once the tile is chosen
  return true;
}

static void vp_pick_tile(int gh, int gw, int halo, int VP_BK, int VP_A_SLOT_BYTES, int* TH, int* TW) {
  long long best = -1;
  const int cand[3][2] = {{8, 16}, {4, 32}, {16, 8}};
  for (auto& c : cand) {
    const int th = c[0], tw = c[1];
    if ((th + halo) * tw * VP_BK * 2 > VP_A_SLOT_BYTES / 2) continue;
    // cost model: pixels covered by the pair tiles, plus the halo rows each CTA fetches on top of its TH rows
    const long long tiles = (long long)((gh + 2 * th - 1) / (2 * th)) * ((gw + tw - 1) / tw);
    const long long cost = tiles * 2 * (th + halo) * tw;
    if (best < 0 || cost < best) { best = cost; *TH = th; *TW = tw; }
  }
}

int conv2d_tc2p_supported(const tcv_conv_desc& d) {
  if (!d.w_tc) return 0;
  if (d.stride != 1 || d.pad_mode != TCV_PAD_ZERO) return 0;
  if (d.cin % 32 != 0 || d.cout % 64 != 0) return 0;
  if (d.cout % 128 != 0 && (g_debug_flags.load() & 8192)) return 0;    // A/B switch: 64-channel layers on conv_tc2 instead
  if (d.x_img_stride != (long long)d.ih * d.iw * d.cin) return 0;
  if (!d.y || d.y_f32) return 0;                       // TMA-store epilogue only
  VPParams p;
  return vp_build_groups(d, p) ? 1 : 0;
}

template <int BN, bool STACK = false, int BK = 32, bool STATS = false>
static int conv_tc2p_bn(const tcv_conv_desc& d, cudaStream_t st) {
  using Cfg = VPCfg<BN, STACK, BK>;
  constexpr int VP_BK = BK;
  VPParams p;
  memset(&p, 0, sizeof(p));
  if (!vp_build_groups(d, p)) return fail(TCV_ERR_UNSUPPORTED, "conv_tc2p: tap pattern not supported");
  const int halo = p.box_rows;
  vp_pick_tile(d.gh, d.gw, halo, BK, Cfg::A_SLOT_BYTES, &p.TH, &p.TW);
  p.box_rows = p.TH + halo;
  p.gh = d.gh; p.gw = d.gw;
  p.tiles_x = (d.gw + p.TW - 1) / p.TW;
  p.tiles_y = (d.gh + 2 * p.TH - 1) / (2 * p.TH);
  p.n_tiles_n = d.cout / BN;
  p.total_work = p.tiles_x * p.tiles_y * p.n_tiles_n * d.n;
  p.kc_iters = d.cin / VP_BK;
  {
    int stages = 0;      // weight stages per work item
    for (int g = 0; g < p.ngroups; ++g) stages += (p.ndy[g] + Cfg::TPS - 1) / Cfg::TPS;
    stages *= p.kc_iters;
    p.b_resident = (p.n_tiles_n == 1 && stages <= Cfg::B_SLOTS && !(g_debug_flags.load() & 32768)) ? 1 : 0;
  }
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  p.idesc2 = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)((2 * BN) >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
  p.dbg = g_debug_flags.load() & (1 | 4 | 8);
  fill_epi(p.epi, d, g_debug_flags.load() & (2 | 32 | 128 | 512));
  p.tma_store = 1;

  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mY_hi, mY_lo;
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(d.x);
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(d.w_tc);
  {
    const __nv_bfloat16* y = reinterpret_cast<const __nv_bfloat16*>(d.y) + ((long long)d.oy_off * d.ow + d.ox_off) * d.cout;
    cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)d.gw, (cuuint64_t)d.gh, (cuuint64_t)d.n};
    cuuint64_t str[3] = {(cuuint64_t)d.ox_mul * d.cout * 2, (cuuint64_t)d.oy_mul * d.ow * d.cout * 2,
                         (cuuint64_t)d.oh * d.ow * d.cout * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    int rc = make_map(&mY_hi, y, 4, dims, str, box, 32);
    if (rc) return rc;
    rc = make_map(&mY_lo, y + (long long)d.n * d.oh * d.ow * d.cout, 4, dims, str, box, 32);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.iw, (cuuint64_t)d.ih, (cuuint64_t)d.n};
    cuuint64_t str[3] = {(cuuint64_t)d.cin * 2, (cuuint64_t)d.iw * d.cin * 2, (cuuint64_t)d.x_img_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)VP_BK, (cuuint32_t)p.TW, (cuuint32_t)p.box_rows, 1};
    int rc = make_map(&mA_hi, a, 4, dims, str, box, VP_BK);
    if (rc) return rc;
    rc = make_map(&mA_lo, a + d.x_plane, 4, dims, str, box, VP_BK);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.cout, (cuuint64_t)d.w_tc_taps};
    cuuint64_t str[2] = {(cuuint64_t)d.cin * 2, (cuuint64_t)d.cout * d.cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)VP_BK, (cuuint32_t)Cfg::B_ROWS, 1};
    int rc = make_map(&mB_hi, b, 3, dims, str, box, VP_BK);
    if (rc) return rc;
    rc = make_map(&mB_lo, b + (long long)d.w_tc_taps * d.cout * d.cin, 3, dims, str, box, VP_BK);
    if (rc) return rc;
  }
  auto kern = conv_tc2p_kernel<BN, STACK, BK, STATS>;
  TCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  int dev = 0, sms = 0;
  TCV_CUDA(cudaGetDevice(&dev));
  TCV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  int npairs = sms / 2;
  if (npairs > p.total_work) npairs = p.total_work;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3((unsigned)(2 * npairs), 1, 1);
  cfg.blockDim = dim3(320, 1, 1);
  cfg.dynamicSmemBytes = Cfg::SMEM;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  cfg.attrs = attr;
  cfg.numAttrs = tc_launch_attrs(attr, 2);
  TCV_CUDA(cudaLaunchKernelEx(&cfg, kern, mA_hi, mA_lo, mB_hi, mB_lo, mY_hi, mY_lo, p));
  return launched("conv_tc2p_kernel");
}

int conv2d_tc2p(const tcv_conv_desc& d, cudaStream_t st) {
  const bool bk64 = d.cin % 64 == 0 && !(g_debug_flags.load() & 65536);       // A/B switch 65536: BK = 32 everywhere
  if (d.stats) {
    // epilogue with per-channel output statistics (tcv_conv_desc.stats): separate instantiations, default tile shapes
    if (d.cout % 256 == 0) return conv_tc2p_bn<256, false, 32, true>(d, st);
    if (d.cout % 128 == 0) return bk64 ? conv_tc2p_bn<128, false, 64, true>(d, st) : conv_tc2p_bn<128, false, 32, true>(d, st);
    return bk64 ? conv_tc2p_bn<64, true, 64, true>(d, st) : conv_tc2p_bn<64, true, 32, true>(d, st);
  }
  if (d.cout % 256 == 0)
    // N = 256: BK = 32 (a stage is already 0.4 us of tensor work; BK = 64 leaves only two weight slots: measured 2-5 % slower)
    return (bk64 && (g_debug_flags.load() & 131072)) ? conv_tc2p_bn<256, false, 64>(d, st) : conv_tc2p_bn<256>(d, st);
  if (d.cout % 128 == 0) return bk64 ? conv_tc2p_bn<128, false, 64>(d, st) : conv_tc2p_bn<128>(d, st);
  if (g_debug_flags.load() & 16384) return conv_tc2p_bn<64, false>(d, st);   // A/B switch: three N = 64 MMAs per K step
  return bk64 ? conv_tc2p_bn<64, true, 64>(d, st) : conv_tc2p_bn<64, true>(d, st);
}

}  // namespace tcv
