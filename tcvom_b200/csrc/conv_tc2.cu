// Persistent tcgen05 convolution kernel with shared-halo activation tiles ("v2").
//
// Why: the first implicit-GEMM kernel (igemm_tc.cu) reloads the 128-pixel A tile once per filter
// tap (9x for a 3x3) and the weight tile once per 128 pixels; at 1080p every conv layer was bound
// by L2->SM bandwidth, not by the tensor pipe (measured: 128->128 3x3 @136x240 ran at the time the
// 1.18 MB/tile of TMA traffic takes at ~6 TB/s).  This kernel cuts that traffic ~2.2-2.7x:
//
//   * one CTA owns a 2*TH x TW pixel super-tile = two M=128 accumulators that share every weight
//     tile (weights are fetched once per 256 output pixels);
//   * for each horizontal tap offset dx the producer loads ONE activation box of (2*TH + halo)
//     rows x TW pixels x 32 channels; the vertical taps dy are then just 1024-byte-aligned row
//     offsets into that box (TW % 8 == 0 keeps every offset on a 64B-swizzle atom boundary), so a
//     3x3 needs 3 boxes of 18 rows instead of 9 boxes of 8 rows per accumulator;
//   * persistent CTAs (grid = #SMs) with a double-buffered TMEM accumulator: the 8 epilogue warps
//     drain super-tile k while TMA/MMA already work on super-tile k+1.
//
// Warp roles (320 threads): warp 0 TMA producer, warp 1 MMA issuer + TMEM owner,
// warps 2..5 epilogue of accumulator 0, warps 6..9 epilogue of accumulator 1.
// Precision: bf16x3 (Ahi.Bhi + Ahi.Blo + Alo.Bhi), fp32 accumulation in TMEM.
#include "tc_epilogue.cuh"

namespace tcv {

extern std::atomic<int> g_debug_flags;

constexpr int V2_BK = 32;
constexpr int V2_MAXG = 3;   // distinct dx values
constexpr int V2_MAXDY = 3;  // taps per dx group
constexpr int V2_A_SLOT_BYTES = 2 * 20480;  // hi + lo planes, up to 320 rows x 64 B each

struct V2Params {
  int gh, gw, TH, TW, tiles_x, tiles_y, n_tiles_n, total_work;
  int kc_iters;
  int ngroups, group_dx[V2_MAXG], ndy[V2_MAXG], dy[V2_MAXG][V2_MAXDY], wtap[V2_MAXG][V2_MAXDY];
  int dy_min, box_rows;
  int tma_store;   // epilogue stages its output in smem and stores it with TMA (needs y, no fp32 side output)
  int dbg;         // measurement switches (tcv_set_debug_flags): 1 no MMA, 2 no epilogue memory ops, 4 A loaded once, 8 B loaded once
  int b_resident;  // all weight tiles of a work item fit the B ring: load them once, never recycle
  uint32_t idesc;
  EpiParams epi;   // fused epilogue (same contract as tcv_conv_desc)
};

template <int BN>
struct V2Cfg {
  static constexpr int B_SLOT_BYTES = 2 * BN * V2_BK * 2;  // hi + lo
  // narrow layers are latency-bound on the activation stream: give them a deeper A ring, and enough
  // B slots to keep all 9 taps of a 32-channel layer resident (weights are then loaded once per CTA)
  static constexpr int B_SLOTS = BN >= 128 ? 5 : (BN >= 64 ? 6 : 9);
  static constexpr int A_SLOTS = BN >= 128 ? 2 : 3;
  static constexpr int A_BYTES = A_SLOTS * V2_A_SLOT_BYTES;
  static constexpr int STAGE_BYTES = 2 * 2 * 128 * 64;        // TMA-store staging: 2 accumulators x hi/lo x 128 rows x 64 B
  static constexpr int SMEM = A_BYTES + B_SLOTS * B_SLOT_BYTES + STAGE_BYTES + 1024 + 512;
  static constexpr int TMEM_COLS = 4 * BN < 32 ? 32 : 4 * BN;  // 2 buffers x 2 accumulators
};

template <int BN>
__global__ void __launch_bounds__(320, 1) conv_tc2_kernel(const __grid_constant__ CUtensorMap mapA_hi,
                                                          const __grid_constant__ CUtensorMap mapA_lo,
                                                          const __grid_constant__ CUtensorMap mapB_hi,
                                                          const __grid_constant__ CUtensorMap mapB_lo,
                                                          const __grid_constant__ CUtensorMap mapY_hi,
                                                          const __grid_constant__ CUtensorMap mapY_lo,
                                                          const __grid_constant__ V2Params p) {
  using Cfg = V2Cfg<BN>;
  constexpr int SB = Cfg::B_SLOTS;
  constexpr int V2_A_SLOTS = Cfg::A_SLOTS;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t b_base = smem_base + Cfg::A_BYTES;
  const uint32_t stage_base = b_base + SB * Cfg::B_SLOT_BYTES;
  const uint32_t bar_base = stage_base + Cfg::STAGE_BYTES;
  auto fullA = [&](int s) { return bar_base + 8u * s; };
  auto emptyA = [&](int s) { return bar_base + 8u * (V2_A_SLOTS + s); };
  auto fullB = [&](int s) { return bar_base + 8u * (2 * V2_A_SLOTS + s); };
  auto emptyB = [&](int s) { return bar_base + 8u * (2 * V2_A_SLOTS + SB + s); };
  auto accFull = [&](int a) { return bar_base + 8u * (2 * V2_A_SLOTS + 2 * SB + a); };
  auto accEmpty = [&](int a) { return bar_base + 8u * (2 * V2_A_SLOTS + 2 * SB + 2 + a); };
  const uint32_t tmem_slot = bar_base + 8u * (2 * V2_A_SLOTS + 2 * SB + 4);
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < V2_A_SLOTS; ++s) { mbar_init(fullA(s), 1); mbar_init(emptyA(s), 1); }
    for (int s = 0; s < SB; ++s) { mbar_init(fullB(s), 1); mbar_init(emptyB(s), 1); }
    for (int a = 0; a < 2; ++a) { mbar_init(accFull(a), 1); mbar_init(accEmpty(a), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB_lo) : "memory");
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

  const int a_plane_bytes = p.box_rows * p.TW * (V2_BK * 2);   // one plane of an A box
  const int tiles_per_img = p.tiles_x * p.tiles_y;

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
    // ================================ TMA producer ================================
    {
      int ia = 0, ib = 0;
      for (int work = blockIdx.x; work < p.total_work; work += gridDim.x) {
        int img, h0, w0, n0;
        decode(work, img, h0, w0, n0);
        const bool load_b = !p.b_resident || work == (int)blockIdx.x;   // resident weights: first item only
        for (int kc = 0; kc < p.kc_iters; ++kc) {
          for (int g = 0; g < p.ngroups; ++g, ++ia) {
            const int sa = ia % V2_A_SLOTS;
            mbar_wait(emptyA(sa), ((uint32_t)(ia / V2_A_SLOTS) & 1u) ^ 1u);
            const uint32_t adst = smem_base + sa * V2_A_SLOT_BYTES;
            if (elect_one()) {
              if ((p.dbg & 4) && ia >= V2_A_SLOTS) {
                mbar_arrive(fullA(sa));
              } else {
                mbar_expect_tx(fullA(sa), 2 * a_plane_bytes);
                tma_load_4d(adst, &mapA_hi, fullA(sa), kc * V2_BK, w0 + p.group_dx[g], h0 + p.dy_min, img);
                tma_load_4d(adst + a_plane_bytes, &mapA_lo, fullA(sa), kc * V2_BK, w0 + p.group_dx[g], h0 + p.dy_min, img);
              }
            }
            __syncwarp();
            if (!load_b) continue;
            for (int j = 0; j < p.ndy[g]; ++j, ++ib) {
              const int sb = ib % SB;   // resident mode: ib < SB, slot == stage index within the work item
              mbar_wait(emptyB(sb), ((uint32_t)(ib / SB) & 1u) ^ 1u);
              const uint32_t bdst = b_base + sb * Cfg::B_SLOT_BYTES;
              if (elect_one()) {
                if ((p.dbg & 8) && ib >= SB) {
                  mbar_arrive(fullB(sb));
                } else {
                  mbar_expect_tx(fullB(sb), Cfg::B_SLOT_BYTES);
                  tma_load_3d(bdst, &mapB_hi, fullB(sb), kc * V2_BK, n0, p.wtap[g][j]);
                  tma_load_3d(bdst + Cfg::B_SLOT_BYTES / 2, &mapB_lo, fullB(sb), kc * V2_BK, n0, p.wtap[g][j]);
                }
              }
              __syncwarp();
            }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      int ia = 0, ib = 0, iw = 0;
      for (int work = blockIdx.x; work < p.total_work; work += gridDim.x, ++iw) {
        const int buf = iw & 1;
        mbar_wait(accEmpty(buf), ((uint32_t)(iw >> 1) & 1u) ^ 1u);   // epilogue has drained this buffer
        tc_fence_after();
        const uint32_t acc0 = tmem_base + (uint32_t)(buf * 2 * BN);
        bool first = true;
        int il = 0;   // weight stage index within this work item
        for (int kc = 0; kc < p.kc_iters; ++kc) {
          for (int g = 0; g < p.ngroups; ++g, ++ia) {
            const int sa = ia % V2_A_SLOTS;
            mbar_wait(fullA(sa), (uint32_t)(ia / V2_A_SLOTS) & 1u);
            tc_fence_after();
            const uint32_t a_hi = smem_base + sa * V2_A_SLOT_BYTES, a_lo = a_hi + a_plane_bytes;
            for (int j = 0; j < p.ndy[g]; ++j, ++ib, ++il) {
              const int sb = p.b_resident ? il : ib % SB;
              if (!p.b_resident || iw == 0) {   // resident weights: only the first work item has to wait for them
                mbar_wait(fullB(sb), p.b_resident ? 0u : ((uint32_t)(ib / SB) & 1u));
                tc_fence_after();
              }
              const uint32_t b_hi = b_base + sb * Cfg::B_SLOT_BYTES, b_lo = b_hi + Cfg::B_SLOT_BYTES / 2;
              if (elect_one()) {
                // rows of accumulator i for vertical tap dy start (i*TH + dy - dy_min) image rows into the box
                const uint32_t roff0 = (uint32_t)((p.dy[g][j] - p.dy_min) * p.TW) * (V2_BK * 2);
                const uint32_t roff1 = roff0 + (uint32_t)(p.TH * p.TW) * (V2_BK * 2);
                const uint32_t d0 = acc0, d1 = acc0 + (uint32_t)BN;
                // consecutive MMAs alternate between the two accumulators: dependent (same-accumulator)
                // MMAs issued back to back do not pipeline as well as independent ones
#pragma unroll
                for (int ks = 0; ks < ((p.dbg & 1) ? 0 : V2_BK / 16); ++ks) {
                  const uint64_t ah0 = smem_desc<V2_BK>(a_hi + roff0 + ks * 32), al0 = smem_desc<V2_BK>(a_lo + roff0 + ks * 32);
                  const uint64_t ah1 = smem_desc<V2_BK>(a_hi + roff1 + ks * 32), al1 = smem_desc<V2_BK>(a_lo + roff1 + ks * 32);
                  const uint64_t bh = smem_desc<V2_BK>(b_hi + ks * 32), bl = smem_desc<V2_BK>(b_lo + ks * 32);
                  const uint32_t acc = (first && ks == 0) ? 0u : 1u;
                  tc_mma(d0, ah0, bh, p.idesc, acc);
                  tc_mma(d1, ah1, bh, p.idesc, acc);
                  tc_mma(d0, ah0, bl, p.idesc, 1u);
                  tc_mma(d1, ah1, bl, p.idesc, 1u);
                  tc_mma(d0, al0, bh, p.idesc, 1u);
                  tc_mma(d1, al1, bh, p.idesc, 1u);
                }
                if (!p.b_resident) tc_commit(emptyB(sb));
                if (j == p.ndy[g] - 1) {
                  tc_commit(emptyA(sa));
                  if (kc == p.kc_iters - 1 && g == p.ngroups - 1) tc_commit(accFull(buf));
                }
              }
              __syncwarp();
              first = false;
            }
          }
        }
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    const int e = warp - 2;            // 0..7
    const int i = e >> 2;              // accumulator (upper / lower tile of the super-tile)
    const int q = warp & 3;            // TMEM lane quarter accessible to this warp
    const int r = q * 32 + lane;
    StoreCtx stc;
    if (p.tma_store) {
      stc.stage_hi = stage_base + (uint32_t)i * (2u * 128u * 64u);
      stc.stage_lo = stc.stage_hi + 128u * 64u;
      stc.row = r;
      stc.bar = 1 + i;
      stc.issuer = (e & 3) == 0 && lane == 0;
      stc.map_hi = &mapY_hi;
      stc.map_lo = &mapY_lo;
    }
    int iw = 0;
    for (int work = blockIdx.x; work < p.total_work; work += gridDim.x, ++iw) {
      int img, h0, w0, n0;
      decode(work, img, h0, w0, n0);
      const int buf = iw & 1;
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(buf * 2 * BN + i * BN);
      const int ty = r / p.TW, tx = r - ty * p.TW;
      const int gy = h0 + i * p.TH + ty, gx = w0 + tx;
      stc.cx = w0;
      stc.cy = h0 + i * p.TH;
      conv_epilogue<BN>(p.epi, taddr, gy < p.gh && gx < p.gw, img, gy, gx, n0, accFull(buf), (uint32_t)(iw >> 1) & 1u,
                        accEmpty(buf), lane, stc);
    }
    if (stc.issuer) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TMEM_COLS)
                 : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
static bool build_groups(const tcv_conv_desc& d, V2Params& p) {
  p.ngroups = 0;
  int dymin = 1 << 30, dymax = -(1 << 30);
  for (int t = 0; t < d.ntaps; ++t) {
    int g = -1;
    for (int k = 0; k < p.ngroups; ++k)
      if (p.group_dx[k] == d.dx[t]) g = k;
    if (g < 0) {
      if (p.ngroups == V2_MAXG) return false;
      g = p.ngroups++;
      p.group_dx[g] = d.dx[t];
      p.ndy[g] = 0;
    }
    if (p.ndy[g] == V2_MAXDY) return false;
    p.dy[g][p.ndy[g]] = d.dy[t];
    p.wtap[g][p.ndy[g]] = d.wtap[t];
    p.ndy[g]++;
    dymin = d.dy[t] < dymin ? d.dy[t] : dymin;
    dymax = d.dy[t] > dymax ? d.dy[t] : dymax;
  }
  p.dy_min = dymin;
  p.box_rows = dymax - dymin;  // halo rows; 2*TH added once the tile is chosen
  return true;
}

static void pick_tile2(int gh, int gw, int halo, int* TH, int* TW) {
  long long best = -1;
  const int cand[3][2] = {{8, 16}, {4, 32}, {16, 8}};
  for (auto& c : cand) {
    const int th = c[0], tw = c[1];
    if ((2 * th + halo) * tw * V2_BK * 2 > V2_A_SLOT_BYTES / 2) continue;
    const long long cover = (long long)((gh + 2 * th - 1) / (2 * th)) * 2 * th * ((gw + tw - 1) / tw) * tw;
    // prefer less wasted work; among equals the first candidate (8x16: smallest halo overhead)
    if (best < 0 || cover < best) { best = cover; *TH = th; *TW = tw; }
  }
}

int conv2d_tc2_supported(const tcv_conv_desc& d) {
  if (!d.w_tc) return 0;
  if (d.stride != 1 || d.pad_mode != TCV_PAD_ZERO) return 0;
  if (d.cin % 32 != 0 || d.cout % 32 != 0) return 0;
  if (d.x_img_stride != (long long)d.ih * d.iw * d.cin) return 0;
  V2Params p;
  return build_groups(d, p) ? 1 : 0;
}

template <int BN>
static int conv_tc2_bn(const tcv_conv_desc& d, cudaStream_t st) {
  using Cfg = V2Cfg<BN>;
  V2Params p;
  memset(&p, 0, sizeof(p));
  if (!build_groups(d, p)) return fail(TCV_ERR_UNSUPPORTED, "conv_tc2: tap pattern not supported");
  const int halo = p.box_rows;
  pick_tile2(d.gh, d.gw, halo, &p.TH, &p.TW);
  p.box_rows = 2 * p.TH + halo;
  p.gh = d.gh; p.gw = d.gw;
  p.tiles_x = (d.gw + p.TW - 1) / p.TW;
  p.tiles_y = (d.gh + 2 * p.TH - 1) / (2 * p.TH);
  p.n_tiles_n = d.cout / BN;
  p.total_work = p.tiles_x * p.tiles_y * p.n_tiles_n * d.n;
  p.kc_iters = d.cin / V2_BK;
  p.dbg = g_debug_flags.load();
  p.b_resident = (p.n_tiles_n == 1 && d.ntaps * p.kc_iters <= Cfg::B_SLOTS) ? 1 : 0;
  p.idesc = instr_desc(BN, false);
  fill_epi(p.epi, d, p.dbg);

  CUtensorMap mA_hi, mA_lo, mB_hi, mB_lo, mY_hi, mY_lo;
  const __nv_bfloat16* a = reinterpret_cast<const __nv_bfloat16*>(d.x);
  const __nv_bfloat16* b = reinterpret_cast<const __nv_bfloat16*>(d.w_tc);
  p.tma_store = (d.y != nullptr && d.y_f32 == nullptr) ? 1 : 0;
  {
    // output view over the compute grid (also expresses the strided placement of a transposed-conv phase)
    const __nv_bfloat16* y0 = reinterpret_cast<const __nv_bfloat16*>(d.y ? d.y : d.x);
    const __nv_bfloat16* y = y0 + (d.y ? ((long long)d.oy_off * d.ow + d.ox_off) * d.cout : 0);
    cuuint64_t dims[4] = {(cuuint64_t)d.cout, (cuuint64_t)d.gw, (cuuint64_t)d.gh, (cuuint64_t)d.n};
    cuuint64_t str[3] = {(cuuint64_t)d.ox_mul * d.cout * 2, (cuuint64_t)d.oy_mul * d.ow * d.cout * 2,
                         (cuuint64_t)d.oh * d.ow * d.cout * 2};
    cuuint32_t box[4] = {32, (cuuint32_t)p.TW, (cuuint32_t)p.TH, 1};
    int rc = make_map(&mY_hi, y, 4, dims, str, box, 32);
    if (rc) return rc;
    rc = make_map(&mY_lo, d.y ? y + (long long)d.n * d.oh * d.ow * d.cout : y, 4, dims, str, box, 32);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[4] = {(cuuint64_t)d.cin, (cuuint64_t)d.iw, (cuuint64_t)d.ih, (cuuint64_t)d.n};
    cuuint64_t str[3] = {(cuuint64_t)d.cin * 2, (cuuint64_t)d.iw * d.cin * 2, (cuuint64_t)d.x_img_stride * 2};
    cuuint32_t box[4] = {(cuuint32_t)V2_BK, (cuuint32_t)p.TW, (cuuint32_t)p.box_rows, 1};
    int rc = make_map(&mA_hi, a, 4, dims, str, box, V2_BK);
    if (rc) return rc;
    rc = make_map(&mA_lo, a + d.x_plane, 4, dims, str, box, V2_BK);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)d.cin, (cuuint64_t)d.cout, (cuuint64_t)d.w_tc_taps};
    cuuint64_t str[2] = {(cuuint64_t)d.cin * 2, (cuuint64_t)d.cout * d.cin * 2};
    cuuint32_t box[3] = {(cuuint32_t)V2_BK, (cuuint32_t)BN, 1};
    int rc = make_map(&mB_hi, b, 3, dims, str, box, V2_BK);
    if (rc) return rc;
    rc = make_map(&mB_lo, b + (long long)d.w_tc_taps * d.cout * d.cin, 3, dims, str, box, V2_BK);
    if (rc) return rc;
  }
  auto kern = conv_tc2_kernel<BN>;
  TCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));
  int dev = 0, sms = 0;
  TCV_CUDA(cudaGetDevice(&dev));
  TCV_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int grid = p.total_work < sms ? p.total_work : sms;
  kern<<<grid, 320, Cfg::SMEM, st>>>(mA_hi, mA_lo, mB_hi, mB_lo, mY_hi, mY_lo, p);
  return launched("conv_tc2_kernel");
}

// Tuning switch (tcv_set_debug_flags bit 8 = 256, results unchanged, default off; to be A/B-measured): prefer N = 64
// tiles over N = 128 when that shortens the schedule.  A persistent grid of 148 CTAs runs ceil(items / 148) rounds; the
// deep layers have few items (512->512 @34x60x3: 96 items = 0.65 rounds, 256->256 @68x120x3: 192 items = 1.3 rounds
// -> 2), so halving the item size can cut the quantisation loss although a narrower tile re-reads the activations.
static bool prefer_bn64(const tcv_conv_desc& d) {
  V2Params p;
  memset(&p, 0, sizeof(p));
  if (!build_groups(d, p)) return false;
  int th = 8, tw = 16;
  pick_tile2(d.gh, d.gw, p.box_rows, &th, &tw);
  const long long tiles = (long long)((d.gw + tw - 1) / tw) * ((d.gh + 2 * th - 1) / (2 * th)) * d.n;
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  const long long i128 = tiles * (d.cout / 128), i64 = tiles * (d.cout / 64);
  const double t128 = (double)((i128 + sms - 1) / sms);          // rounds x cost of an N = 128 item (1.0)
  const double t64 = (double)((i64 + sms - 1) / sms) * 0.56;     // an N = 64 item: half the MMAs, same activation boxes
  return t64 < t128;
}

int conv2d_tc2(const tcv_conv_desc& d, cudaStream_t st) {
  if (d.cout % 128 == 0) {
    if ((g_debug_flags.load() & 256) && prefer_bn64(d)) return conv_tc2_bn<64>(d, st);
    return conv_tc2_bn<128>(d, st);
  }
  if (d.cout % 64 == 0) return conv_tc2_bn<64>(d, st);
  return conv_tc2_bn<32>(d, st);
}

}  // namespace tcv
