// tcgen05 implicit-GEMM kernel for sm_100a: one kernel template serves
//   * the 3x3 / 1x1 / (4 phases of the) 4x4-stride-2-transposed convolutions  (EPI_CONV)
//   * the two guided-contextual-attention GEMMs  scores = Q.Kn^T, O = P.Vt^T  (EPI_F32 / EPI_BF16)
//
// Structure (one 128 x BN output tile per CTA, 320 threads):
//   warp 0      TMA producer   cp.async.bulk.tensor (4-D NHWC activation boxes, 3-D weight boxes)
//                              -> 64B/128B-swizzled K-major shared-memory tiles, mbarrier ring
//   warp 1      MMA issuer     one elected thread issues tcgen05.mma (M=128, N=BN, K=16, bf16 ->
//                              fp32 accumulators in TMEM); tcgen05.commit frees smem stages
//   warps 2..9  epilogue       tcgen05.ld (32 lanes x 32 columns per warp) -> fused affine /
//                              residual / activation -> split-bf16 NHWC (or fp32 / bf16 GEMM output)
//
// Precision ("bf16x3"): activations and weights are stored as bf16 hi/lo pairs; with NSPLIT==3 each
// K step issues Ahi.Bhi + Ahi.Blo + Alo.Bhi, which keeps ~16 mantissa bits per operand (measured
// 2e-5 max-abs on the alpha matte vs 1.6e-2 for plain bf16).  NSPLIT==1 is
// plain bf16 (used for P.V where the operands are probabilities / averaged values).
//
// The A operand of a convolution is never materialised (no im2col): for filter tap (dy,dx) the
// producer loads the box {BK channels, TW, TH, 1} at (c0, w0+dx, h0+dy, n); out-of-image
// coordinates are zero-filled by TMA, which implements the zero padding.
#include "tc_common.cuh"

namespace tcv {

extern std::atomic<int> g_debug_flags;
int gemm_tc2(const void* A, long long a_plane, const void* B, long long b_plane, float* C, int M, int N, int K, long long ldc,
             long long c_batch_stride, int batch, int mode, cudaStream_t st);

enum { EPI_CONV = 0, EPI_F32 = 1, EPI_BF16 = 2 };

struct TcParams {
  int gh, gw, tiles_x, TH, TW;
  int ntaps, kc_iters;
  int stride;  // input pixels per output pixel (TMA traversal stride)
  int dy[TCV_MAX_TAPS], dx[TCV_MAX_TAPS], wtap[TCV_MAX_TAPS];
  uint32_t idesc;  // tcgen05 instruction descriptor (operand format bf16 or fp16)
  int b_batched;  // GEMM mode: third weight-map coordinate = blockIdx.z instead of the tap index
  // split-K GEMM mode (weight gradients): blockIdx.z selects the K range [z*split_k, (z+1)*split_k) of ONE
  // [rows][K] operand pair instead of a batch entry; koff shifts the A operand along K (filter-tap offset in the
  // zero-ringed channel-major activation; TMA zero-fills coordinates outside [0, K))
  int split_k, koff;
  // stacked-B mode (weight gradients of <= 64-channel layers): the B tile is nstack sub-tiles of sub_rows rows, sub-tile
  // j loaded at K offset -bkoff[j] (vertical filter taps), so one MMA of N = nstack*sub_rows columns serves nstack taps
  int nstack, sub_rows, bkoff[3];
  uint32_t stage_tx;  // bytes per pipeline stage when they differ from the full tile (0: Cfg::STAGE_BYTES)
  // conv epilogue
  int n_imgs, oh, ow, cout, oy_mul, oy_off, ox_mul, ox_off;
  __nv_bfloat16* y;
  float* y_f32;
  const float *s1, *b1, *s2, *b2;
  const __nv_bfloat16 *res1, *res2;
  long long res1_plane, res2_plane;
  int res1_shift, act;
  // GEMM epilogue: C[batch][M][ldc]
  void* c;
  long long ldc, c_batch_stride;
  int M, N;
};

template <int BN, int BK, int NSPLIT>
struct TcCfg {
  static constexpr int A_BYTES = 128 * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int PLANES = NSPLIT == 6 ? 3 : (NSPLIT == 3 ? 2 : 1);
  static constexpr int STAGE_BYTES = PLANES * (A_BYTES + B_BYTES);
  // two CTAs per SM when the tile is small enough (their mainloop/epilogue phases overlap)
  static constexpr int BUDGET = BN <= 128 ? 100 * 1024 : 200 * 1024;
  static constexpr int STAGES_RAW = BUDGET / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : (STAGES_RAW < 2 ? 2 : STAGES_RAW);
  static constexpr int SMEM = STAGES * STAGE_BYTES + 1024 /*align*/ + 256 /*barriers*/;
};

template <int BN, int BK, int NSPLIT, int EPI>
__global__ void __launch_bounds__(320) igemm_tc_kernel(const __grid_constant__ CUtensorMap mapA_hi,
                                                       const __grid_constant__ CUtensorMap mapA_lo,
                                                       const __grid_constant__ CUtensorMap mapA_p2,
                                                       const __grid_constant__ CUtensorMap mapB_hi,
                                                       const __grid_constant__ CUtensorMap mapB_lo,
                                                       const __grid_constant__ CUtensorMap mapB_p2,
                                                       const __grid_constant__ TcParams p) {
  using Cfg = TcCfg<BN, BK, NSPLIT>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = smem_base + STAGES * Cfg::STAGE_BYTES;  // full[STAGES], empty[STAGES], accum, tmem ptr
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * STAGES);
  const uint32_t tmem_slot = accum_bar + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // GEMM mode rasterises N-tiles fastest: consecutive CTAs then share one A row panel (the 4 MB
  // probability panel of P.V stays in L2 while the 16 value tiles stream past it; with M fastest the
  // whole 268 MB P matrix was re-read from HBM once per N-tile -- ncu: 4.2 GB dram reads per launch)
  const int tile = p.b_batched ? blockIdx.y : blockIdx.x;
  const int tile_y = tile / p.tiles_x, tile_x = tile - tile_y * p.tiles_x;
  const int h0 = tile_y * p.TH, w0 = tile_x * p.TW;
  const int n0 = (p.b_batched ? blockIdx.x : blockIdx.y) * BN;
  const int img = blockIdx.z;
  const int total_iters = p.ntaps * p.kc_iters;
  const int kbase = p.split_k ? img * p.split_k : 0;
  const int aimg = p.split_k ? 0 : img;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapB_hi) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    {
      int it = 0;
      for (int t = 0; t < p.ntaps; ++t) {
        const int cw = w0 * p.stride + p.dx[t], ch = h0 * p.stride + p.dy[t];
        const int bz = p.b_batched ? aimg : p.wtap[t];
        for (int kc = 0; kc < p.kc_iters; ++kc, ++it) {
          const int s = it % STAGES;
          const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
          mbar_wait(empty_bar(s), ph ^ 1u);
          const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
          if (elect_one()) {
            mbar_expect_tx(full_bar(s), p.stage_tx ? p.stage_tx : (uint32_t)Cfg::STAGE_BYTES);
            const int ka = kbase + kc * BK + p.koff, kb = kbase + kc * BK;
            tma_load_4d(st, &mapA_hi, full_bar(s), ka, cw, ch, aimg);
            if (NSPLIT >= 3) tma_load_4d(st + Cfg::A_BYTES, &mapA_lo, full_bar(s), ka, cw, ch, aimg);
            if (p.nstack > 1) {
              const uint32_t sub = (uint32_t)p.sub_rows * BK * 2;
              for (int j = 0; j < p.nstack; ++j) {
                tma_load_3d(st + Cfg::PLANES * Cfg::A_BYTES + j * sub, &mapB_hi, full_bar(s), kb - p.bkoff[j], 0, 0);
                if (NSPLIT >= 3)
                  tma_load_3d(st + Cfg::PLANES * Cfg::A_BYTES + Cfg::B_BYTES + j * sub, &mapB_lo, full_bar(s),
                              kb - p.bkoff[j], 0, 0);
              }
            } else {
              tma_load_3d(st + Cfg::PLANES * Cfg::A_BYTES, &mapB_hi, full_bar(s), kb, n0, bz);
              if (NSPLIT >= 3)
                tma_load_3d(st + Cfg::PLANES * Cfg::A_BYTES + Cfg::B_BYTES, &mapB_lo, full_bar(s), kb, n0, bz);
            }
            if (NSPLIT == 6) {
              tma_load_4d(st + 2 * Cfg::A_BYTES, &mapA_p2, full_bar(s), ka, cw, ch, aimg);
              tma_load_3d(st + 3 * Cfg::A_BYTES + 2 * Cfg::B_BYTES, &mapB_p2, full_bar(s), kb, n0, bz);
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    {
      const uint32_t idesc = p.idesc;
      for (int it = 0; it < total_iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t st = smem_base + s * Cfg::STAGE_BYTES;
        const uint32_t a_hi = st, a_lo = st + Cfg::A_BYTES;
        const uint32_t b_hi = st + Cfg::PLANES * Cfg::A_BYTES, b_lo = b_hi + Cfg::B_BYTES;
        if (elect_one()) {
#pragma unroll
        for (int ks = 0; ks < BK / 16; ++ks) {
          const uint64_t ah = smem_desc<BK>(a_hi + ks * 32), bh = smem_desc<BK>(b_hi + ks * 32);
          tc_mma(tmem_d, ah, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
          if (NSPLIT >= 3) {
            const uint64_t al = smem_desc<BK>(a_lo + ks * 32), bl = smem_desc<BK>(b_lo + ks * 32);
            tc_mma(tmem_d, ah, bl, idesc, 1u);
            tc_mma(tmem_d, al, bh, idesc, 1u);
            if (NSPLIT == 6) {   // three bf16 planes per operand (~24 mantissa bits): + h.l2, l2.h, m.m
              const uint64_t a2 = smem_desc<BK>(a_lo + Cfg::A_BYTES + ks * 32);
              const uint64_t b2 = smem_desc<BK>(b_lo + Cfg::B_BYTES + ks * 32);
              tc_mma(tmem_d, ah, b2, idesc, 1u);
              tc_mma(tmem_d, a2, bh, idesc, 1u);
              tc_mma(tmem_d, al, bl, idesc, 1u);
            }
          }
        }
        tc_commit(empty_bar(s));  // frees the smem stage once these MMAs have read it
        if (it == total_iters - 1) tc_commit(accum_bar);  // accumulator complete
        }
        __syncwarp();
      }
    }
  } else {
    // ================================ epilogue (warps 2..9) ================================
    // two warps per TMEM lane quarter, each draining half of the accumulator columns (the GEMM epilogues
    // write 64 KB of fp32 per tile; with four warps the epilogue, not the MMA, bounded the scores GEMM)
    const int q = warp & 3;                 // TMEM lane quarter this warp may access
    const int r = q * 32 + lane;            // accumulator row = tile pixel
    constexpr int CHALF = BN / 2 >= 32 ? BN / 2 : 32;
    const int c_begin = ((warp - 2) >> 2) * CHALF;
    const int c_end = c_begin + CHALF < BN ? c_begin + CHALF : BN;
    mbar_wait(accum_bar, 0);
    tc_fence_after();
    const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
    if (EPI == EPI_CONV) {
      const int ty = r / p.TW, tx = r - ty * p.TW;
      const int gy = h0 + ty, gx = w0 + tx;
      const bool valid = gy < p.gh && gx < p.gw;
      const int oy = gy * p.oy_mul + p.oy_off, ox = gx * p.ox_mul + p.ox_off;
      const long long oplane = (long long)p.n_imgs * p.oh * p.ow * p.cout;
      const long long obase = (((long long)img * p.oh + oy) * p.ow + ox) * p.cout + n0;
      const int rh = p.oh >> p.res1_shift, rw = p.ow >> p.res1_shift;
      const long long r1base = (((long long)img * rh + (oy >> p.res1_shift)) * rw + (ox >> p.res1_shift)) * p.cout + n0;
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        tc_ld32(taddr + c0, v);
        if (!valid) continue;
        float f[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
        if (p.s1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] *= __ldg(p.s1 + n0 + c0 + j);
        }
        if (p.b1) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] += __ldg(p.b1 + n0 + c0 + j);
        }
        if (p.res1) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float g[8];
            load8(p.res1 + r1base + c0 + j, p.res1_plane, g);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[j + k] += g[k];
          }
        }
        apply_act_n<32>(f, p.act);
        if (p.s2) {
#pragma unroll
          for (int j = 0; j < 32; ++j) f[j] = f[j] * __ldg(p.s2 + n0 + c0 + j) + __ldg(p.b2 + n0 + c0 + j);
        }
        if (p.res2) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) {
            float g[8];
            load8(p.res2 + obase + c0 + j, p.res2_plane, g);
#pragma unroll
            for (int k = 0; k < 8; ++k) f[j + k] += g[k];
          }
        }
        if (p.y) {
#pragma unroll
          for (int j = 0; j < 32; j += 8) store8(p.y + obase + c0 + j, oplane, f + j);
        }
        if (p.y_f32) {
#pragma unroll
          for (int j = 0; j < 32; j += 4)
            *reinterpret_cast<float4*>(p.y_f32 + obase + c0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
        }
      }
    } else {
      const int m = w0 + r;  // GEMM mode: TH == 1, tile rows are consecutive M indices
      const bool valid = m < p.M;
#pragma unroll 1
      for (int c0 = c_begin; c0 < c_end; c0 += 32) {
        uint32_t v[32];
        tc_ld32(taddr + c0, v);
        if (!valid) continue;
        const int nb = n0 + c0;
        if (nb >= p.N) continue;
        if (EPI == EPI_F32) {
          float* c = reinterpret_cast<float*>(p.c) + (long long)img * p.c_batch_stride + (long long)m * p.ldc + nb;
          if (nb + 32 <= p.N) {
#pragma unroll
            for (int j = 0; j < 32; j += 4)
              *reinterpret_cast<uint4*>(c + j) = make_uint4(v[j], v[j + 1], v[j + 2], v[j + 3]);
          } else {
            for (int j = 0; j < 32 && nb + j < p.N; ++j) c[j] = __uint_as_float(v[j]);
          }
        } else {
          __nv_bfloat16* c =
              reinterpret_cast<__nv_bfloat16*>(p.c) + (long long)img * p.c_batch_stride + (long long)m * p.ldc + nb;
          for (int j = 0; j < 32 && nb + j < p.N; ++j) c[j] = __float2bfloat16_rn(__uint_as_float(v[j]));
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"((uint32_t)BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------ host
static void pick_tile(int gh, int gw, int stride, int* TH, int* TW) {
  long long best = -1;
  for (int tw = 128; tw >= 4; tw >>= 1) {
    if (tw * stride > 256 || (128 / tw) * stride > 256) continue;  // TMA box dimension limit
    const int th = 128 / tw;
    const long long cover = (long long)((gh + th - 1) / th) * th * ((gw + tw - 1) / tw) * tw;
    if (best < 0 || cover < best) {
      best = cover;
      *TH = th;
      *TW = tw;
    }
  }
}

struct TcOperands {
  const __nv_bfloat16* a;   // activation / A matrix, hi plane
  long long a_plane;        // elements hi->lo
  int n, h, w, c;           // NHWC dims of A
  long long a_img_stride;   // elements
  const __nv_bfloat16* b;   // weights / B matrix [bz][rows][c], hi plane
  long long b_plane;
  int b_rows, b_z;
  bool fp16;                // operands are IEEE half instead of bf16 (NSPLIT == 1 GEMMs only)
  int grid_z = 0;           // split-K mode: number of K slices (grid.z)
};

template <int BN, int BK, int NSPLIT, int EPI>
static int launch_tc(const TcOperands& o, TcParams& p, cudaStream_t st) {
  using Cfg = TcCfg<BN, BK, NSPLIT>;
  CUtensorMap mA_hi, mA_lo, mA_p2, mB_hi, mB_lo, mB_p2;
  {
    cuuint64_t dims[4] = {(cuuint64_t)o.c, (cuuint64_t)o.w, (cuuint64_t)o.h, (cuuint64_t)o.n};
    cuuint64_t str[3] = {(cuuint64_t)o.c * 2, (cuuint64_t)o.w * o.c * 2, (cuuint64_t)o.a_img_stride * 2};
    // with a traversal stride s the box spans s*T input pixels and delivers T of them
    cuuint32_t box[4] = {(cuuint32_t)BK, (cuuint32_t)(p.TW * p.stride), (cuuint32_t)(p.TH * p.stride), 1};
    int rc = make_map(&mA_hi, o.a, 4, dims, str, box, BK, o.fp16, p.stride);
    if (rc) return rc;
    rc = make_map(&mA_lo, NSPLIT >= 3 ? o.a + o.a_plane : o.a, 4, dims, str, box, BK, o.fp16, p.stride);
    if (rc) return rc;
    rc = make_map(&mA_p2, NSPLIT == 6 ? o.a + 2 * o.a_plane : o.a, 4, dims, str, box, BK, o.fp16, p.stride);
    if (rc) return rc;
  }
  {
    cuuint64_t dims[3] = {(cuuint64_t)o.c, (cuuint64_t)o.b_rows, (cuuint64_t)o.b_z};
    cuuint64_t str[2] = {(cuuint64_t)o.c * 2, (cuuint64_t)o.b_rows * o.c * 2};
    cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)(p.nstack > 1 ? p.sub_rows : BN), 1};
    int rc = make_map(&mB_hi, o.b, 3, dims, str, box, BK, o.fp16);
    if (rc) return rc;
    rc = make_map(&mB_lo, NSPLIT >= 3 ? o.b + o.b_plane : o.b, 3, dims, str, box, BK);
    if (rc) return rc;
    rc = make_map(&mB_p2, NSPLIT == 6 ? o.b + 2 * o.b_plane : o.b, 3, dims, str, box, BK);
    if (rc) return rc;
  }
  p.idesc = instr_desc(p.nstack > 1 ? p.nstack * p.sub_rows : BN, o.fp16);
  if (p.nstack > 1)
    p.stage_tx = (uint32_t)(Cfg::PLANES * (Cfg::A_BYTES + p.nstack * p.sub_rows * BK * 2));
  auto kern = igemm_tc_kernel<BN, BK, NSPLIT, EPI>;
  TCV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM));  // per device
  p.tiles_x = (p.gw + p.TW - 1) / p.TW;
  const int tiles_y = (p.gh + p.TH - 1) / p.TH;
  const int nrows = EPI == EPI_CONV ? p.cout : p.N;
  const int gz = p.split_k ? o.grid_z : o.n;
  dim3 grid(p.tiles_x * tiles_y, (nrows + BN - 1) / BN, gz);
  if (p.b_batched) grid = dim3((nrows + BN - 1) / BN, p.tiles_x * tiles_y, gz);
  kern<<<grid, 320, Cfg::SMEM, st>>>(mA_hi, mA_lo, mA_p2, mB_hi, mB_lo, mB_p2, p);
  return launched("igemm_tc_kernel");
}

// ---- convolution entry -----------------------------------------------------------------------
// The packed tensor-core weights live next to the fp32 packed weights: tcv_conv_desc.w points at
// fp32 [taps][cin][cout]; the bf16 hi/lo [taps][cout][cin] copy is passed through w_tc (see cabi).
int conv2d_tc_supported(const tcv_conv_desc& d) {
  if (!d.w_tc) return 0;
  if ((d.stride != 1 && d.stride != 2) || d.pad_mode != TCV_PAD_ZERO) return 0;
  if (d.cin % 32 != 0 || d.cout % 32 != 0) return 0;
  if (d.x_img_stride != (long long)d.ih * d.iw * d.cin) return 0;
  if (!d.y && d.y_f32 == nullptr) return 0;
  return 1;
}

template <int BN>
static int conv_tc_bn(const tcv_conv_desc& d, cudaStream_t st) {
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.gh = d.gh; p.gw = d.gw;
  pick_tile(d.gh, d.gw, d.stride, &p.TH, &p.TW);
  p.ntaps = d.ntaps;
  p.stride = d.stride;
  p.kc_iters = d.cin / 32;
  for (int t = 0; t < d.ntaps; ++t) { p.dy[t] = d.dy[t]; p.dx[t] = d.dx[t]; p.wtap[t] = d.wtap[t]; }
  p.b_batched = 0;
  p.n_imgs = d.n; p.oh = d.oh; p.ow = d.ow; p.cout = d.cout;
  p.oy_mul = d.oy_mul; p.oy_off = d.oy_off; p.ox_mul = d.ox_mul; p.ox_off = d.ox_off;
  p.y = reinterpret_cast<__nv_bfloat16*>(d.y);
  p.y_f32 = d.y_f32;
  p.s1 = d.s1; p.b1 = d.b1; p.s2 = d.s2; p.b2 = d.b2;
  p.res1 = reinterpret_cast<const __nv_bfloat16*>(d.res1);
  p.res2 = reinterpret_cast<const __nv_bfloat16*>(d.res2);
  p.res1_plane = d.res1_plane; p.res2_plane = d.res2_plane; p.res1_shift = d.res1_shift; p.act = d.act;
  TcOperands o;
  o.a = reinterpret_cast<const __nv_bfloat16*>(d.x);
  o.a_plane = d.x_plane;
  o.n = d.n; o.h = d.ih; o.w = d.iw; o.c = d.cin; o.a_img_stride = d.x_img_stride;
  o.b = reinterpret_cast<const __nv_bfloat16*>(d.w_tc);
  o.b_rows = d.cout;
  o.b_z = d.w_tc_taps;
  o.fp16 = false;
  o.b_plane = (long long)d.w_tc_taps * d.cout * d.cin;
  return launch_tc<BN, 32, 3, EPI_CONV>(o, p, st);
}

int conv2d_tc(const tcv_conv_desc& d, cudaStream_t st) {
  if (d.cout % 128 == 0) return conv_tc_bn<128>(d, st);
  if (d.cout % 64 == 0) return conv_tc_bn<64>(d, st);
  return conv_tc_bn<32>(d, st);
}

}  // namespace tcv

using namespace tcv;

// C[b] = A[b] * B[b]^T on tensor cores.  A: bf16 [batch][M][K] (hi plane; lo plane a_plane elements
// later when nsplit == 3), B: bf16 [batch][N][K]; C: fp32 (out_bf16 == 0) or bf16 [batch][M][ldc].
// K % 64 == 0.  in_fp16: operands are IEEE half (nsplit == 1 only).
extern "C" int tcv_gemm_tn_tc(const void* A, long long a_plane, const void* B, long long b_plane, void* C, int M,
                              int N, int K, long long ldc, long long c_batch_stride, int batch, int nsplit,
                              int out_bf16, int in_fp16, tcv_stream_t stream) {
  TCV_REQUIRE(A && B && C, "gemm_tn_tc: null pointer");
  TCV_REQUIRE(M > 0 && N > 0 && K > 0 && K % 64 == 0, "gemm_tn_tc: K must be a positive multiple of 64");
  const int bk = nsplit >= 3 ? 32 : 64;
  TCV_REQUIRE(nsplit == 1 || nsplit == 3 || nsplit == 6, "gemm_tn_tc: nsplit must be 1, 3 or 6");
  TCV_REQUIRE(!in_fp16 || nsplit == 1, "gemm_tn_tc: fp16 operands only with nsplit == 1");
  TCV_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0,
              "gemm_tn_tc: pointers must be 16-byte aligned");
  TCV_REQUIRE(out_bf16 || ldc % 4 == 0, "gemm_tn_tc: ldc must be a multiple of 4 for fp32 output");
  {
    // large fp32-output bf16x3 products (the GCA scores / aggregation GEMMs): persistent 256-wide tiles, by default on
    // CTA pairs (gemm_tc2.cu).  tcv_set_debug_flags: 1024 = single-CTA 128 x 256 tiles, 2048 = the per-tile kernel below.
    const int flags = g_debug_flags.load();
    if (nsplit == 3 && !out_bf16 && !in_fp16 && M >= 512 && N >= 256 && !(flags & 2048))
      return gemm_tc2(A, a_plane, B, b_plane, reinterpret_cast<float*>(C), M, N, K, ldc, c_batch_stride, batch,
                      (flags & 1024) ? 1 : 2, S(stream));
  }
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.gh = 1; p.gw = M; p.TH = 1; p.TW = 128;
  p.ntaps = 1; p.kc_iters = K / bk; p.stride = 1;
  p.b_batched = 1;
  p.c = C; p.ldc = ldc; p.c_batch_stride = c_batch_stride; p.M = M; p.N = N;
  TcOperands o;
  o.a = reinterpret_cast<const __nv_bfloat16*>(A);
  o.a_plane = a_plane;
  o.n = batch; o.h = 1; o.w = M; o.c = K; o.a_img_stride = (long long)M * K;
  o.b = reinterpret_cast<const __nv_bfloat16*>(B);
  o.b_plane = b_plane;
  o.b_rows = N; o.b_z = batch;
  o.fp16 = in_fp16 != 0;
  cudaStream_t st = S(stream);
  if (nsplit == 6) {
    TCV_REQUIRE(!out_bf16, "gemm_tn_tc: nsplit == 6 writes fp32");
    return launch_tc<128, 32, 6, EPI_F32>(o, p, st);
  }
  if (nsplit == 3) {
    if (out_bf16) return launch_tc<128, 32, 3, EPI_BF16>(o, p, st);
    return launch_tc<128, 32, 3, EPI_F32>(o, p, st);
  }
  if (out_bf16) return launch_tc<256, 64, 1, EPI_BF16>(o, p, st);
  return launch_tc<256, 64, 1, EPI_F32>(o, p, st);
}


// ---- weight gradient on the tensor cores ---------------------------------------------------------------
// dw[wtap[t]][ci][co] += sum_p XT[ci][p + dy[t]*row + dx[t]] * ZT[co][p]: per filter tap one split-K GEMM over the
// zero-ringed channel-major copies of the activation and of the output gradient (tcv_transpose_pad), bf16x3,
// partial sums per K slice in `partial`, then one reduction kernel.
namespace tcv {
// partial [nsplit][cin][nstack*cout] -> dw[wt[j]][ci][co] += sum over slices
struct WgTaps { int wt[3]; };
__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int nsplit, int nstack, int cin, int cout,
                                    WgTaps taps, float* __restrict__ dw, int dw_cout) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  const int ncol = nstack * cout;
  if (i >= cin * ncol) return;
  const int ci = i / ncol, col = i - ci * ncol;
  const int j = col / cout, co = col - j * cout;
  float s = 0.f;
  for (int k = 0; k < nsplit; ++k) s += partial[((long long)k * cin + ci) * ncol + col];
  dw[((long long)taps.wt[j] * cin + ci) * dw_cout + co] += s;
}
}  // namespace tcv

extern "C" int tcv_wgrad_tc(const void* xt, long long xt_plane, const void* zt, long long zt_plane, int cin, int cout,
                            long long ktot, int row_stride, int ntaps, const int* dy, const int* dx, const int* wtap,
                            float* partial, int nsplit, float* dw, int dw_cout, tcv_stream_t stream) {
  TCV_REQUIRE(xt && zt && partial && dw && dy && dx && wtap, "wgrad_tc: null pointer");
  TCV_REQUIRE(cin > 0 && cout > 0 && cout % 4 == 0 && ktot > 0 && ktot % 8 == 0 && ktot < (1LL << 31),
              "wgrad_tc: bad geometry (cout %% 4, ktot %% 8)");
  TCV_REQUIRE(ntaps >= 1 && ntaps <= TCV_MAX_TAPS && nsplit >= 1, "wgrad_tc: bad ntaps / nsplit");
  for (int t = 0; t < ntaps; ++t)
    TCV_REQUIRE((dy[t] * row_stride + dx[t]) % 8 == 0, "wgrad_tc: tap %d offset %d not a multiple of 8 elements", t,
                dy[t] * row_stride + dx[t]);
  TCV_REQUIRE(((uintptr_t)xt & 15) == 0 && ((uintptr_t)zt & 15) == 0 && xt_plane % 8 == 0 && zt_plane % 8 == 0,
              "wgrad_tc: operands must be 16-byte aligned");
  cudaStream_t st = S(stream);
  long long chunk = (ktot + nsplit - 1) / nsplit;
  chunk = (chunk + 31) / 32 * 32;
  const int slices = (int)((ktot + chunk - 1) / chunk);
  // <= 64 output channels: the (up to 3) taps of this call are stacked along N (one GEMM, A streamed once)
  const bool stack = ntaps > 1 && ntaps <= 3 && cout % 8 == 0 && ntaps * cout <= 256 && (ntaps * cout) % 16 == 0;
  for (int t = 0; t < ntaps; t += stack ? ntaps : 1) {
    TcParams p;
    memset(&p, 0, sizeof(p));
    p.gh = 1; p.gw = cin; p.TH = 1; p.TW = 128;
    p.ntaps = 1; p.kc_iters = (int)(chunk / 32); p.stride = 1;
    p.b_batched = 1;
    p.split_k = (int)chunk;
    WgTaps wt;
    wt.wt[0] = wt.wt[1] = wt.wt[2] = wtap[t];
    const int ns = stack ? ntaps : 1;
    if (stack) {
      p.nstack = ntaps; p.sub_rows = cout;
      for (int j = 0; j < ntaps; ++j) { p.bkoff[j] = dy[j] * row_stride + dx[j]; wt.wt[j] = wtap[j]; }
    } else {
      p.koff = dy[t] * row_stride + dx[t];
    }
    const int ncol = ns * cout;
    p.c = partial; p.ldc = ncol; p.c_batch_stride = (long long)cin * ncol; p.M = cin; p.N = ncol;
    TcOperands o;
    o.a = reinterpret_cast<const __nv_bfloat16*>(xt);
    o.a_plane = xt_plane;
    o.n = 1; o.h = 1; o.w = cin; o.c = (int)ktot; o.a_img_stride = (long long)cin * ktot;
    o.b = reinterpret_cast<const __nv_bfloat16*>(zt);
    o.b_plane = zt_plane;
    o.b_rows = cout; o.b_z = 1;
    o.fp16 = false;
    o.grid_z = slices;
    int rc;
    if (stack) {
      if (ncol > 128) rc = launch_tc<256, 32, 3, EPI_F32>(o, p, st);
      else if (ncol > 64) rc = launch_tc<128, 32, 3, EPI_F32>(o, p, st);
      else if (ncol > 32) rc = launch_tc<64, 32, 3, EPI_F32>(o, p, st);
      else rc = launch_tc<32, 32, 3, EPI_F32>(o, p, st);
    } else if (cout % 128 == 0) rc = launch_tc<128, 32, 3, EPI_F32>(o, p, st);
    else if (cout % 64 == 0) rc = launch_tc<64, 32, 3, EPI_F32>(o, p, st);
    else rc = launch_tc<32, 32, 3, EPI_F32>(o, p, st);
    if (rc) return rc;
    wgrad_reduce_kernel<<<(cin * ncol + 255) / 256, 256, 0, st>>>(partial, slices, ns, cin, cout, wt, dw, dw_cout);
    rc = launched("wgrad_reduce_kernel");
    if (rc) return rc;
  }
  return TCV_OK;
}
