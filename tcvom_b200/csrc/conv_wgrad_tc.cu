// Weight gradient on the tensor cores, read straight from the NHWC activations (no transposed copies):
//   dw[wtap[t]][ci][co] += sum_{img,y,x} x[img, y+dy[t], x+dx[t], ci] * dz[img, y*mul+oy, x*mul+ox, co]
// is a GEMM with M = ci, N = co and the reduction over pixels.  In NHWC a pixel's channels are contiguous, so a
// TMA box {64 channels, TW, TH, 1} lands in shared memory as [K = pixel][MN = channel] rows of 128 B: exactly the
// canonical MN-major SWIZZLE_128B operand layout of tcgen05.mma (cute::UMMA make_umma_desc<Major::MN>:
// ((8,n),(8,k)):((1,LBO),(8,SBO)) in 16-byte units -- 64 channels contiguous, 8 pixels per 1024-byte atom,
// SBO = 1024 B between pixel groups, LBO = bytes between 64-channel blocks).  Both operands are MN-major
// (instruction-descriptor bits 15/16).  The filter tap is a shift of the x box in (W, H); TMA zero-fills
// out-of-image coordinates, which is the convolution's zero padding; stride-2 convs read x and the deconv phases
// read dz with a TMA traversal stride of 2.  Precision: bf16x3 (hi.hi + hi.lo + lo.hi) like every other tensor-core kernel here.
//
// One CTA = one (filter tap, 128x{64,128} tile of the (ci, co) plane, slice of the pixel tiles); 192 threads:
// warp 0 TMA producer, warp 1 MMA issuer, warps 2..5 epilogue (TMEM -> fp32 atomics into dw).
#include <algorithm>

#include "tc_common.cuh"

namespace tcv {

constexpr int WG_KT = 64;                 // pixels per pipeline stage
constexpr int WG_MAX_STAGES = 8;
constexpr int WG_SMEM_BUDGET = 192 * 1024;

struct WgParams {
  int n, gh, gw, TW, TH, tiles_x, tiles_y, total_tiles, tiles_per_slice;
  int ntaps, dy[TCV_MAX_TAPS], dx[TCV_MAX_TAPS], wtap[TCV_MAX_TAPS];
  int cin, cout, dz_c, mul, oy, ox;
  int xmul;               // input pixels per output pixel (conv stride): x is read with a traversal stride
  // operand geometry: a channel block is one TMA box of row_bytes/2 channels (128 B rows: SWIZZLE_128B, 64 channels;
  // 64 B rows: SWIZZLE_64B, 32 channels -- layers with <= 32 channels on both sides move half the bytes).  The MMA is
  // always M = 128: with a_nblk == 1 the leading-dimension offset of the A descriptor is 0, so the rows beyond the first
  // block alias it (their accumulator rows are duplicates and never stored) instead of streaming zero-filled boxes.
  int row_bytes, a_nblk, nblk, stages;
  int n_tiles_n;
  // stacked-tap mode (both sides <= 64 channels): the roles are swapped -- A = dz (M side = output channels, one
  // block, aliased up to M = 128), B = the x boxes of `tpg` filter taps stacked along N (LBO = one block), so one
  // MMA of N = tpg * cblk columns serves tpg taps.  An MN-major MMA costs ~100 cycles per K = 16 step almost
  // independently of N (measured), so the narrow layers want few, wide MMAs.
  int swap, tpg, ngroups;
  // rows mode (3x3 unit-spaced taps, stride 1, both sides <= 32 channels): ONE MMA per K step serves all nine taps.
  // A = dz of the three output rows that meet x row y (blocks b = 0..2: row y - (dy0 + b); the MMA's 4th block aliases
  // the next buffer, its accumulator rows are never read), B = x row y at the three horizontal shifts (blocks c = 0..2:
  // column + dx0 + c).  Per 64 pixels that is 6 boxes and 12 MMAs instead of 11 boxes and 24 MMAs in the two CTAs of the
  // stacked-tap mode.  fold: an 8-channel x (the network input) is read through an overlapping tensor map whose
  // "channels" are 3 neighbouring pixels x 8 channels, so one box holds all three horizontal shifts (nbx = 1).
  int rows, fold, nbx, dy0, dx0, y_first, wtap3[9];
  uint32_t idesc;
  float* dw;
};

// MN-major shared-memory descriptor: LBO between channel blocks, SBO = 8 pixel rows; layout 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t smem_desc_mn(uint32_t addr, uint32_t lbo_bytes, uint32_t row_bytes) {
  const uint64_t layout = row_bytes == 128 ? 2 : 4;
  return (uint64_t)((addr & 0x3FFFF) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)((8 * row_bytes) >> 4) << 32) |
         (1ull << 46) | (layout << 61);
}

__global__ void __launch_bounds__(192) conv_wgrad_tc_kernel(const __grid_constant__ CUtensorMap mapX_hi,
                                                            const __grid_constant__ CUtensorMap mapX_lo,
                                                            const __grid_constant__ CUtensorMap mapZ_hi,
                                                            const __grid_constant__ CUtensorMap mapZ_lo,
                                                            const __grid_constant__ WgParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const int WG_BLK = WG_KT * p.row_bytes;                // bytes of one channel block of a stage
  const int WG_STAGES = p.stages;
  const int cblk = p.row_bytes / 2;                      // channels per block
  // wide: A = x (a_nblk blocks), B = dz (nblk blocks); stacked: A = dz (1 block), B = x boxes of tpg taps
  const int nb_a = p.rows ? 3 : (p.swap ? 1 : p.a_nblk), nb_b = p.rows ? p.nbx : (p.swap ? p.tpg : p.nblk);
  const int stage_bytes = 2 * (nb_a + nb_b) * WG_BLK;
  const uint32_t bar_base = smem_base + WG_STAGES * stage_bytes;
  auto full_bar = [&](int s) { return bar_base + 8u * s; };
  auto empty_bar = [&](int s) { return bar_base + 8u * (WG_MAX_STAGES + s); };
  const uint32_t accum_bar = bar_base + 8u * (2 * WG_MAX_STAGES);
  const uint32_t tmem_slot = accum_bar + 8u;
  uint32_t* tmem_slot_ptr = reinterpret_cast<uint32_t*>(smem_raw + (tmem_slot - smem_u32(smem_raw)));

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = blockIdx.x;                                // filter tap / tap group (fastest: they share L2 lines)
  const int t0 = p.swap ? t * p.tpg : t;                   // first tap of this CTA
  const int nt_here = p.swap ? min(p.tpg, p.ntaps - t0) : 1;
  const int slice = blockIdx.y;
  const int mt = blockIdx.z / p.n_tiles_n, nt = blockIdx.z % p.n_tiles_n;
  const int ci0 = mt * 128, co0 = nt * (cblk * p.nblk);
  const int tile_begin = slice * p.tiles_per_slice;
  const int tile_end = min(tile_begin + p.tiles_per_slice, p.total_tiles);
  const int iters = tile_end - tile_begin;
  // (rows == 2: two accumulators side by side)
  const uint32_t acc_cols = (uint32_t)(cblk * nb_b);
  const uint32_t ncols_used = p.rows == 2 ? 2 * acc_cols : acc_cols;
  uint32_t ncols = 32;                                     // TMEM allocation: power of two >= columns used
  while (ncols < ncols_used) ncols <<= 1;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < WG_STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(accum_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapX_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapZ_hi) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_d = *tmem_slot_ptr;

  if (iters > 0) {
    if (warp == 0) {
      // ================================ TMA producer ================================
      for (int it = 0; it < iters; ++it) {
        const int tile = tile_begin + it;
        const int tx = tile % p.tiles_x;
        const int ty = (tile / p.tiles_x) % p.tiles_y;
        const int img = tile / (p.tiles_x * p.tiles_y);
        const int x0 = tx * p.TW, y0 = ty * p.TH;
        const int s = it % WG_STAGES;
        const uint32_t ph = (uint32_t)(it / WG_STAGES) & 1u;
        mbar_wait(empty_bar(s), ph ^ 1u);
        const uint32_t st = smem_base + s * stage_bytes;
        if (elect_one()) {
          const int zx = x0 * p.mul + p.ox, zy = y0 * p.mul + p.oy;   // dz (sub-sampled by mul for the deconv phases)
          const uint32_t bst = st + 2 * nb_a * WG_BLK;
          if (p.rows) {
            mbar_expect_tx(full_bar(s), (uint32_t)stage_bytes);
            const int xy = p.y_first + y0;                   // x row of this tile (tiles run over the rows of x)
            for (int b = 0; b < 3; ++b) {
              tma_load_4d(st + b * WG_BLK, &mapZ_hi, full_bar(s), 0, x0, xy - (p.dy0 + b), img);
              tma_load_4d(st + (3 + b) * WG_BLK, &mapZ_lo, full_bar(s), 0, x0, xy - (p.dy0 + b), img);
            }
            for (int c = 0; c < nb_b; ++c) {
              tma_load_4d(bst + c * WG_BLK, &mapX_hi, full_bar(s), 0, x0 + p.dx0 + c, xy, img);
              tma_load_4d(bst + (nb_b + c) * WG_BLK, &mapX_lo, full_bar(s), 0, x0 + p.dx0 + c, xy, img);
            }
          } else if (p.swap) {
            mbar_expect_tx(full_bar(s), (uint32_t)(2 * (1 + nt_here) * WG_BLK));
            tma_load_4d(st, &mapZ_hi, full_bar(s), 0, zx, zy, img);
            tma_load_4d(st + WG_BLK, &mapZ_lo, full_bar(s), 0, zx, zy, img);
            for (int j = 0; j < nt_here; ++j) {              // x shifted by each tap of the group
              const int ax = x0 * p.xmul + p.dx[t0 + j], ay = y0 * p.xmul + p.dy[t0 + j];
              tma_load_4d(bst + j * WG_BLK, &mapX_hi, full_bar(s), 0, ax, ay, img);
              tma_load_4d(bst + (nb_b + j) * WG_BLK, &mapX_lo, full_bar(s), 0, ax, ay, img);
            }
          } else {
            mbar_expect_tx(full_bar(s), (uint32_t)stage_bytes);
            // A = x shifted by the tap: blocks (channels ci0.., ci0+cblk..) x planes (hi, lo)
            const int ax = x0 * p.xmul + p.dx[t], ay = y0 * p.xmul + p.dy[t];
            for (int b = 0; b < p.a_nblk; ++b) {
              tma_load_4d(st + b * WG_BLK, &mapX_hi, full_bar(s), ci0 + cblk * b, ax, ay, img);
              tma_load_4d(st + (p.a_nblk + b) * WG_BLK, &mapX_lo, full_bar(s), ci0 + cblk * b, ax, ay, img);
            }
            for (int b = 0; b < p.nblk; ++b) {
              tma_load_4d(bst + b * WG_BLK, &mapZ_hi, full_bar(s), co0 + cblk * b, zx, zy, img);
              tma_load_4d(bst + (p.nblk + b) * WG_BLK, &mapZ_lo, full_bar(s), co0 + cblk * b, zx, zy, img);
            }
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ================================ MMA issuer ================================
      for (int it = 0; it < iters; ++it) {
        const int s = it % WG_STAGES;
        const uint32_t ph = (uint32_t)(it / WG_STAGES) & 1u;
        mbar_wait(full_bar(s), ph);
        tc_fence_after();
        const uint32_t st = smem_base + s * stage_bytes;
        const uint32_t a_hi = st, a_lo = st + nb_a * WG_BLK;
        const uint32_t b_hi = st + 2 * nb_a * WG_BLK, b_lo = b_hi + nb_b * WG_BLK;
        const uint32_t a_lbo = nb_a > 1 ? (uint32_t)WG_BLK : 0u, rb = (uint32_t)p.row_bytes;
        // the last tap group may hold fewer taps: its MMAs are narrower (idesc N field, bits 17..22)
        const uint32_t idesc = p.swap ? ((p.idesc & ~(0x3Fu << 17)) | ((uint32_t)((nt_here * cblk) >> 3) << 17)) : p.idesc;
        if (elect_one()) {
#pragma unroll
          for (int ks = 0; ks < WG_KT / 16; ++ks) {
            const uint32_t koff = ks * 16 * rb;            // 16 pixels further along K
            const uint64_t ah = smem_desc_mn(a_hi + koff, a_lbo, rb), al = smem_desc_mn(a_lo + koff, a_lbo, rb);
            const uint64_t bh = smem_desc_mn(b_hi + koff, WG_BLK, rb), bl = smem_desc_mn(b_lo + koff, WG_BLK, rb);
            tc_mma(tmem_d, ah, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
            tc_mma(tmem_d, ah, bl, idesc, 1u);
            tc_mma(tmem_d, al, bh, idesc, 1u);
            if (p.rows == 2) {
              // 64-channel blocks: M = 128 holds two of the three dz row blocks; the third one (its upper half aliases
              // the next buffer, never read) accumulates into a second accumulator
              const uint64_t ah2 = smem_desc_mn(a_hi + 2 * WG_BLK + koff, a_lbo, rb);
              const uint64_t al2 = smem_desc_mn(a_lo + 2 * WG_BLK + koff, a_lbo, rb);
              tc_mma(tmem_d + acc_cols, ah2, bh, idesc, (it > 0 || ks > 0) ? 1u : 0u);
              tc_mma(tmem_d + acc_cols, ah2, bl, idesc, 1u);
              tc_mma(tmem_d + acc_cols, al2, bh, idesc, 1u);
            }
          }
          tc_commit(empty_bar(s));
          if (it == iters - 1) tc_commit(accum_bar);
        }
        __syncwarp();
      }
    } else {
      // ================================ epilogue (warps 2..5) ================================
      const int q = warp & 3;                      // TMEM lane quarter this warp may access
      const int r = q * 32 + lane;                 // accumulator row = input channel within the tile
      mbar_wait(accum_bar, 0);
      tc_fence_after();
      const uint32_t taddr = tmem_d + ((uint32_t)(q * 32) << 16);
      if (p.rows == 2) {
        // 64-channel blocks.  accumulator 0: rows 0..63 = dz row block 0, 64..127 = block 1; accumulator 1: rows 0..63 =
        // block 2.  columns = (horizontal shift c, input channel): c0 = c * 64 + ci
        const int co = (q & 1) * 32 + lane;
#pragma unroll 1
        for (int acc = 0; acc < 2; ++acc) {
          const int blk = acc * 2 + (q >> 1);
          if (blk > 2) continue;
#pragma unroll 1
          for (int c0 = 0; c0 < (int)acc_cols; c0 += 32) {
            uint32_t v[32];
            tc_ld32(taddr + acc * acc_cols + c0, v);
            if (co >= p.dz_c) continue;
            const int c = c0 >> 6, ci0 = c0 & 63;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const float f = __uint_as_float(v[i]);
              if (ci0 + i < p.cin && f != 0.f)
                atomicAdd(p.dw + ((long long)p.wtap3[blk * 3 + c] * p.cin + ci0 + i) * p.dz_c + co, f);
            }
          }
        }
      } else if (p.rows) {
        // accumulator rows = (output row block b = this warp, output channel), columns = (horizontal shift, input channel)
        if (q < 3) {
#pragma unroll 1
          for (int c0 = 0; c0 < (int)ncols_used; c0 += 32) {
            uint32_t v[32];
            tc_ld32(taddr + c0, v);
            if (lane >= p.dz_c) continue;
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              const int c = p.fold ? (i >> 3) : (c0 >> 5), ci = p.fold ? (i & 7) : i;
              const float f = __uint_as_float(v[i]);
              if (c < 3 && ci < p.cin && f != 0.f)
                atomicAdd(p.dw + ((long long)p.wtap3[q * 3 + c] * p.cin + ci) * p.dz_c + lane, f);
            }
          }
        }
      } else if (p.swap) {
        // rows = output channels (duplicates beyond the block are skipped), columns = (tap of the group, input channel)
        const bool row_ok = r < cblk && r < p.dz_c;
#pragma unroll 1
        for (int c0 = 0; c0 < nt_here * cblk; c0 += 32) {
          uint32_t v[32];
          tc_ld32(taddr + c0, v);
          if (!row_ok) continue;
          const int j = c0 / cblk, cc0 = c0 - j * cblk;
          float* out = p.dw + ((long long)p.wtap[t0 + j] * p.cin + cc0) * p.dz_c + r;
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float f = __uint_as_float(v[i]);
            if (cc0 + i < p.cin && f != 0.f) atomicAdd(out + (long long)i * p.dz_c, f);
          }
        }
      } else {
        const int ci = ci0 + r;
        float* out = p.dw + ((long long)p.wtap[t] * p.cin + ci) * p.dz_c + co0;
#pragma unroll 1
        for (int c0 = 0; c0 < (int)ncols_used; c0 += 32) {
          uint32_t v[32];
          tc_ld32(taddr + c0, v);
          if (ci >= p.cin) continue;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float f = __uint_as_float(v[j]);
            if (co0 + c0 + j < p.dz_c && f != 0.f) atomicAdd(out + c0 + j, f);
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_d), "r"(ncols) : "memory");
  }
}

// c_mem / w_mem: the channels and width of the tensor in memory when they differ from the extents (c, w) the map exposes
// (fold mode: a pixel of the map is 3 neighbouring 8-channel pixels, pixel stride still 8 channels)
static int make_nhwc_map(CUtensorMap* m, const __nv_bfloat16* base, int c, int w, int h, int n, int tw, int th, int trav,
                         int cblk, int c_mem = 0, int w_mem = 0) {
  if (c_mem == 0) c_mem = c;
  if (w_mem == 0) w_mem = w;
  cuuint64_t dims[4] = {(cuuint64_t)c, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t str[3] = {(cuuint64_t)c_mem * 2, (cuuint64_t)w_mem * c_mem * 2, (cuuint64_t)h * w_mem * c_mem * 2};
  cuuint32_t box[4] = {(cuuint32_t)cblk, (cuuint32_t)(tw * trav), (cuuint32_t)(th * trav), 1};
  return make_map(m, base, 4, dims, str, box, /*bk: 64 -> SWIZZLE_128B, 32 -> SWIZZLE_64B*/ cblk, false, trav);
}

int conv2d_wgrad_tc_supported(const tcv_conv_desc& d, int dz_c) {
  if ((d.stride != 1 && d.stride != 2) || d.pad_mode != TCV_PAD_ZERO) return 0;
  if (d.cin % 8 != 0 || dz_c % 8 != 0) return 0;
  if (d.x_img_stride != (long long)d.ih * d.iw * d.cin) return 0;
  if (d.oy_mul != d.ox_mul || (d.oy_mul != 1 && d.oy_mul != 2)) return 0;
  if (d.stride == 2 && d.oy_mul != 1) return 0;
  return 1;
}

int conv2d_wgrad_tc(const tcv_conv_desc& d, const __nv_bfloat16* dz, long long dz_plane, int dz_c, float* dw,
                    cudaStream_t st) {
  WgParams p;
  memset(&p, 0, sizeof(p));
  p.n = d.n; p.gh = d.gh; p.gw = d.gw;
  if (d.gw >= 64) { p.TW = 64; p.TH = 1; }
  else if (d.gw >= 32) { p.TW = 32; p.TH = 2; }
  else if (d.gw >= 16) { p.TW = 16; p.TH = 4; }
  else { p.TW = 8; p.TH = 8; }
  p.tiles_x = (d.gw + p.TW - 1) / p.TW;
  p.tiles_y = (d.gh + p.TH - 1) / p.TH;
  p.total_tiles = d.n * p.tiles_x * p.tiles_y;
  p.ntaps = d.ntaps;
  for (int t = 0; t < d.ntaps; ++t) { p.dy[t] = d.dy[t]; p.dx[t] = d.dx[t]; p.wtap[t] = d.wtap[t]; }
  p.cin = d.cin; p.cout = d.cout; p.dz_c = dz_c;
  p.mul = d.oy_mul; p.oy = d.oy_off; p.ox = d.ox_off;
  p.xmul = d.stride;
  const bool narrow = d.cin <= 32 && dz_c <= 32;
  p.row_bytes = narrow ? 64 : 128;
  const int cblk = p.row_bytes / 2;
  p.nblk = dz_c > cblk ? 2 : 1;
  p.a_nblk = d.cin > cblk ? 2 : 1;
  const int ntile = cblk * p.nblk;
  p.n_tiles_n = (dz_c + ntile - 1) / ntile;
  const int n_tiles_m = (d.cin + 127) / 128;
  // D = f32, A = B = bf16, both MN-major (bits 15, 16), M = 128, N = ntile
  p.idesc = (1u << 4) | (1u << 7) | (1u << 10) | (1u << 15) | (1u << 16) | ((uint32_t)(ntile >> 3) << 17) |
            ((uint32_t)(128 >> 4) << 24);
  p.dw = dw;
  int mn = n_tiles_m * p.n_tiles_n;
  int grid_x = d.ntaps;
  if (d.cin <= cblk && dz_c <= cblk) {
    p.swap = 1;
    p.ngroups = (d.ntaps * cblk + 255) / 256;
    p.tpg = (d.ntaps + p.ngroups - 1) / p.ngroups;
    p.ngroups = (d.ntaps + p.tpg - 1) / p.tpg;
    grid_x = p.ngroups;
    mn = 1;
    p.idesc = (p.idesc & ~(0x3Fu << 17)) | ((uint32_t)((p.tpg * cblk) >> 3) << 17);
  }
  // rows mode: nine unit-spaced taps of a stride-1 narrow layer in one MMA (see WgParams)
  int x_w = d.iw, x_c = d.cin;                 // geometry of the x tensor map (fold: overlapping 24-"channel" pixels)
  const bool mid = !narrow && d.cin <= 64 && dz_c <= 64;       // 64-channel blocks: rows mode with two accumulators
  if ((narrow || (mid && !(g_debug_flags.load() & (1 << 21)))) && d.stride == 1 && p.mul == 1 && p.oy == 0 && p.ox == 0 &&
      d.ntaps == 9 && !(g_debug_flags.load() & (1 << 20))) {
    int dy0 = d.dy[0], dx0 = d.dx[0];
    for (int t = 1; t < 9; ++t) { dy0 = std::min(dy0, d.dy[t]); dx0 = std::min(dx0, d.dx[t]); }
    bool grid3 = true;
    int seen = 0;
    for (int t = 0; t < 9; ++t) {
      const int b = d.dy[t] - dy0, c = d.dx[t] - dx0;
      if (b < 0 || b > 2 || c < 0 || c > 2) { grid3 = false; break; }
      seen |= 1 << (b * 3 + c);
      p.wtap3[b * 3 + c] = d.wtap[t];
    }
    if (grid3 && seen == 0x1FF) {
      p.rows = mid ? 2 : 1; p.swap = 0;
      p.dy0 = dy0; p.dx0 = dx0;
      p.y_first = std::max(0, dy0);
      const int y_count = std::min(d.ih, d.gh + dy0 + 2) - p.y_first;
      p.tiles_y = (y_count + p.TH - 1) / p.TH;
      p.total_tiles = d.n * p.tiles_x * p.tiles_y;
      p.fold = (!mid && d.cin == 8 && dx0 >= 0) ? 1 : 0;
      p.nbx = p.fold ? 1 : 3;
      if (p.fold) { x_w = d.iw - 2; x_c = 24; }
      p.idesc = (p.idesc & ~(0x3Fu << 17)) | ((uint32_t)((p.nbx * cblk) >> 3) << 17);
      grid_x = 1;
      mn = 1;
    }
  }
  int slices = (4 * 148 + grid_x * mn - 1) / (grid_x * mn);
  if (p.rows) slices = 148;                    // 48 KB stages: one resident CTA per SM
  const int max_slices = (p.total_tiles + 3) / 4;          // at least 4 pixel tiles (256 pixels) per CTA
  if (slices > max_slices) slices = max_slices;
  if (slices < 1) slices = 1;
  p.tiles_per_slice = (p.total_tiles + slices - 1) / slices;
  slices = (p.total_tiles + p.tiles_per_slice - 1) / p.tiles_per_slice;

  CUtensorMap mX_hi, mX_lo, mZ_hi, mZ_lo;
  const __nv_bfloat16* x = reinterpret_cast<const __nv_bfloat16*>(d.x);
  int rc = make_nhwc_map(&mX_hi, x, x_c, x_w, d.ih, d.n, p.TW, p.TH, d.stride, cblk, d.cin, d.iw);
  if (rc) return rc;
  if ((rc = make_nhwc_map(&mX_lo, x + d.x_plane, x_c, x_w, d.ih, d.n, p.TW, p.TH, d.stride, cblk, d.cin, d.iw))) return rc;
  if ((rc = make_nhwc_map(&mZ_hi, dz, dz_c, d.ow, d.oh, d.n, p.TW, p.TH, p.mul, cblk))) return rc;
  if ((rc = make_nhwc_map(&mZ_lo, dz + dz_plane, dz_c, d.ow, d.oh, d.n, p.TW, p.TH, p.mul, cblk))) return rc;
  const int stage_bytes = 2 * (p.rows ? 3 + p.nbx : (p.swap ? 1 + p.tpg : p.a_nblk + p.nblk)) * WG_KT * p.row_bytes;
  p.stages = WG_SMEM_BUDGET / stage_bytes;
  if (p.stages > WG_MAX_STAGES) p.stages = WG_MAX_STAGES;
  const int smem = p.stages * stage_bytes + 1024 + 256;
  TCV_CUDA(cudaFuncSetAttribute(conv_wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  dim3 grid(grid_x, slices, mn);
  conv_wgrad_tc_kernel<<<grid, 192, smem, st>>>(mX_hi, mX_lo, mZ_hi, mZ_lo, p);
  return launched("conv_wgrad_tc_kernel");
}

}  // namespace tcv

using namespace tcv;

extern "C" int tcv_conv2d_wgrad_nhwc_tc(const tcv_conv_desc* dp, const void* dz, long long dz_plane, int dz_c, float* dw,
                                        tcv_stream_t stream) {
  TCV_REQUIRE(dp && dz && dw, "conv2d_wgrad_nhwc_tc: null pointer");
  tcv_conv_desc d = *dp;
  if (d.x_plane == 0) d.x_plane = (long long)d.n * d.ih * d.iw * d.cin;
  if (d.x_img_stride == 0) d.x_img_stride = (long long)d.ih * d.iw * d.cin;
  if (dz_plane == 0) dz_plane = (long long)d.n * d.oh * d.ow * dz_c;
  TCV_REQUIRE(conv2d_wgrad_tc_supported(d, dz_c), "conv2d_wgrad_nhwc_tc: shape not supported (zero-padded / pre-padded stride 1 or 2 only)");
  TCV_REQUIRE(((uintptr_t)d.x & 15) == 0 && ((uintptr_t)dz & 15) == 0 && d.x_plane % 8 == 0 && dz_plane % 8 == 0,
              "conv2d_wgrad_nhwc_tc: operands must be 16-byte aligned");
  return conv2d_wgrad_tc(d, reinterpret_cast<const __nv_bfloat16*>(dz), dz_plane, dz_c, dw, S(stream));
}
