// Fused convolution epilogue shared by the persistent tcgen05 conv kernels: TMEM accumulator ->
// per-channel affine (BatchNorm / bias) -> residual -> activation -> affine -> residual -> split-bf16 NHWC
// (and/or fp32) store.  One warp handles 32 accumulator rows (= 32 output pixels) x BN channels.
#pragma once
#include "tc_common.cuh"

namespace tcv {

// 32 bytes per lane in one request (LDG.E.256, sm_100): the residual operand of a 32-channel chunk is 64 B per pixel and plane
// at a pixel pitch of cout * 2 B, so a warp-wide 16-byte load touches 32 sectors and uses half of each; measured with the
// MMAs switched off, those loads were 17 of the 27 us the epilogue adds to a 128 -> 128 layer.
__device__ __forceinline__ void ldg256(const void* p, uint4& a, uint4& b) {
  asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
               : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w), "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w)
               : "l"(p));
}

struct EpiParams {
  int n_imgs, oh, ow, cout, oy_mul, oy_off, ox_mul, ox_off;
  __nv_bfloat16* y;
  float* y_f32;
  const float *s1, *b1, *s2, *b2;
  const __nv_bfloat16 *res1, *res2;
  long long res1_plane, res2_plane;
  int res1_shift, act;
  int dbg;
  double* stats;      // optional per-(image group, channel) sum / sum of squares of the output (tcv_conv_desc.stats)
  int stats_groups, stats_copies;
};

// Sum over the 32 lanes of a warp of 32 per-lane values: afterwards lane j holds the total of element j.  Recursive halving:
// 16 + 8 + 4 + 2 + 1 = 31 shuffles instead of 32 x 5.
__device__ __forceinline__ float warp_transpose_sum32(const float* f, int lane) {
  float t[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) {
    const bool up = lane & 16;
    const float mine = up ? f[16 + i] : f[i], give = up ? f[i] : f[16 + i];
    t[i] = mine + __shfl_xor_sync(0xffffffffu, give, 16);
  }
#pragma unroll
  for (int w = 8; w >= 1; w >>= 1) {
#pragma unroll
    for (int i = 0; i < w; ++i) {
      const bool up = lane & w;
      const float mine = up ? t[w + i] : t[i], give = up ? t[i] : t[w + i];
      t[i] = mine + __shfl_xor_sync(0xffffffffu, give, w);
    }
  }
  return t[0];
}

static inline void fill_epi(EpiParams& e, const tcv_conv_desc& d, int dbg) {
  e.n_imgs = d.n; e.oh = d.oh; e.ow = d.ow; e.cout = d.cout;
  e.oy_mul = d.oy_mul; e.oy_off = d.oy_off; e.ox_mul = d.ox_mul; e.ox_off = d.ox_off;
  e.y = reinterpret_cast<__nv_bfloat16*>(d.y);
  e.y_f32 = d.y_f32;
  e.s1 = d.s1; e.b1 = d.b1; e.s2 = d.s2; e.b2 = d.b2;
  e.res1 = reinterpret_cast<const __nv_bfloat16*>(d.res1);
  e.res2 = reinterpret_cast<const __nv_bfloat16*>(d.res2);
  e.res1_plane = d.res1_plane; e.res2_plane = d.res2_plane; e.res1_shift = d.res1_shift; e.act = d.act;
  e.dbg = dbg;
  e.stats = d.stats;
  e.stats_groups = d.stats_groups > 0 ? d.stats_groups : 1;
  e.stats_copies = d.stats_copies > 0 ? d.stats_copies : 1;
}

// per-channel sum / sum of squares of the 32 pixels of this warp (rows outside the image were zeroed by the caller);
// lane j ends up with channel j of the chunk and adds it to its fp64 accumulator pair
static __device__ __noinline__ void epi_stats(float* g, int lane, double* acc) {
  const float s = warp_transpose_sum32(g, lane);
#pragma unroll
  for (int j = 0; j < 32; ++j) g[j] *= g[j];
  const float q = warp_transpose_sum32(g, lane);
  atomicAdd(acc, (double)s);
  atomicAdd(acc + 1, (double)q);
}

// TMA-store path: the 4 warps that drain one accumulator stage each 32-channel chunk in a 64B-swizzled
// shared-memory tile (row r at stage + 64*r) and one elected thread hands it to the copy engine, so the output
// leaves the SM as whole pixel rows instead of 32 scattered 16-byte pieces per store instruction.
struct StoreCtx {
  uint32_t stage_hi = 0, stage_lo = 0;   // smem tiles (128 rows x 64 B each); 0: plain global stores
  int row = 0;                           // this thread's accumulator row
  int bar = 0;                           // named barrier pair (bar, bar + 2) private to the 4 warps
  bool issuer = false;                   // the one thread that issues / waits for the bulk stores
  const CUtensorMap* map_hi = nullptr;   // output views {cout, gw, gh, n} of the hi / lo plane
  const CUtensorMap* map_lo = nullptr;
  int cx = 0, cy = 0;                    // tile origin in the compute grid
  // double-buffered staging (alt_bytes != 0): chunk k stages at stage_* + (k & 1) * alt_bytes, so that the bulk store of the
  // previous chunk may still be reading its tile while this one is being written (wait_group.read 1 instead of 0)
  uint32_t alt_bytes = 0;
  uint32_t chunk = 0;                    // running chunk counter of this warp group (kept across work items)
};

// gy/gx: position of this thread's pixel in the compute grid; in_grid: inside it.  The warp waits on
// `acc_full` (parity given), drains its 32 TMEM lanes starting at `taddr`, and arrives on `acc_empty` as soon
// as its last tcgen05.ld has completed.
// split_halves: the accumulator holds a second partial product SPLIT_OFF columns further right; the two are summed first.
template <int BN, int SPLIT_OFF = BN, bool STATS = false>
__device__ __forceinline__ void conv_epilogue(const EpiParams& p, uint32_t taddr, bool in_grid, int img, int gy, int gx,
                                              int n0, uint32_t acc_full, uint32_t full_parity, uint32_t acc_empty,
                                              int lane, StoreCtx& st, bool split_halves = false,
                                              bool remote_empty = false) {
  const bool valid = in_grid && !(p.dbg & 2);
  const int oy = gy * p.oy_mul + p.oy_off, ox = gx * p.ox_mul + p.ox_off;
  const long long oplane = (long long)p.n_imgs * p.oh * p.ow * p.cout;
  const long long obase = (((long long)img * p.oh + oy) * p.ow + ox) * p.cout + n0;
  const int rh = p.oh >> p.res1_shift, rw = p.ow >> p.res1_shift;
  const long long r1base = (((long long)img * rh + (oy >> p.res1_shift)) * rw + (ox >> p.res1_shift)) * p.cout + n0;
  const bool has1 = valid && p.res1 != nullptr && !(p.dbg & 32), has2 = valid && p.res2 != nullptr && !(p.dbg & 32);

  // Residual operands are fetched one 32-channel chunk AHEAD of the accumulator chunk that consumes them,
  // and the first chunk is requested before this warp waits for the MMAs to finish.
  uint4 ra[2][8], rb[2][8];   // [double buffer][4 x hi, 4 x lo] for res1 / res2
  auto fetch = [&](int c0, int slot) {
    if (has1) {
      const __nv_bfloat16* h = p.res1 + r1base + c0;
      const __nv_bfloat16* l = h + p.res1_plane;
      ldg256(h, ra[slot][0], ra[slot][1]); ldg256(h + 16, ra[slot][2], ra[slot][3]);
      ldg256(l, ra[slot][4], ra[slot][5]); ldg256(l + 16, ra[slot][6], ra[slot][7]);
    }
    if (has2) {
      const __nv_bfloat16* h = p.res2 + obase + c0;
      const __nv_bfloat16* l = h + p.res2_plane;
      ldg256(h, rb[slot][0], rb[slot][1]); ldg256(h + 16, rb[slot][2], rb[slot][3]);
      ldg256(l, rb[slot][4], rb[slot][5]); ldg256(l + 16, rb[slot][6], rb[slot][7]);
    }
  };
  auto add_res = [&](const uint4* rr, float* f) {
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const uint32_t hw[4] = {rr[k].x, rr[k].y, rr[k].z, rr[k].w};
      const uint32_t lw[4] = {rr[4 + k].x, rr[4 + k].y, rr[4 + k].z, rr[4 + k].w};
#pragma unroll
      for (int m = 0; m < 4; ++m) {
        f[8 * k + 2 * m] += __uint_as_float(hw[m] << 16) + __uint_as_float(lw[m] << 16);
        f[8 * k + 2 * m + 1] += __uint_as_float(hw[m] & 0xffff0000u) + __uint_as_float(lw[m] & 0xffff0000u);
      }
    }
  };
  fetch(0, 0);
  if (BN / 32 == 2) fetch(32, 1);     // two chunks: both residual operands are in flight while the MMAs still run
  mbar_wait(acc_full, full_parity);
  tc_fence_after();
#pragma unroll
  for (int ci = 0; ci < BN / 32; ++ci) {
    const int c0 = ci * 32;
    const int slot = ci & 1;
    uint32_t v[32];
    tc_ld32(taddr + c0, v);
    if (split_halves) {   // BN == 32: columns [32,64) hold the A_hi.B_lo partial product
      uint32_t v2[32];
      tc_ld32(taddr + SPLIT_OFF + c0, v2);
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
    }
    if (ci == BN / 32 - 1) {
      // all TMEM reads of this warp are complete: hand the buffer back to the MMA warp early
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        // remote_empty: `acc_empty` is a shared::cluster address (the leader CTA's barrier of a CTA pair)
        if (remote_empty)
          asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty) : "memory");
        else
          mbar_arrive(acc_empty);
      }
    } else if (BN / 32 != 2) {
      fetch(c0 + 32, slot ^ 1);
    }
    float f[32];
    if (valid) {
#pragma unroll
      for (int j = 0; j < 32; ++j) f[j] = __uint_as_float(v[j]);
      if (p.s1 && p.b1 && !(p.dbg & 128)) {           // BatchNorm affine: one FMA per element
        const float4* sv = reinterpret_cast<const float4*>(p.s1 + n0 + c0);
        const float4* bv = reinterpret_cast<const float4*>(p.b1 + n0 + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(sv + j), u = __ldg(bv + j);
          f[4 * j] = fmaf(f[4 * j], t.x, u.x); f[4 * j + 1] = fmaf(f[4 * j + 1], t.y, u.y);
          f[4 * j + 2] = fmaf(f[4 * j + 2], t.z, u.z); f[4 * j + 3] = fmaf(f[4 * j + 3], t.w, u.w);
        }
      } else if (!(p.dbg & 128)) {
        if (p.s1) {
          const float4* sv = reinterpret_cast<const float4*>(p.s1 + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(sv + j);
            f[4 * j] *= t.x; f[4 * j + 1] *= t.y; f[4 * j + 2] *= t.z; f[4 * j + 3] *= t.w;
          }
        }
        if (p.b1) {
          const float4* sv = reinterpret_cast<const float4*>(p.b1 + n0 + c0);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 t = __ldg(sv + j);
            f[4 * j] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
          }
        }
      }
      if (has1) add_res(ra[slot], f);
      apply_act_n<32>(f, p.act);
      if (p.s2 && !(p.dbg & 128)) {
        const float4* sv = reinterpret_cast<const float4*>(p.s2 + n0 + c0);
        const float4* bv = reinterpret_cast<const float4*>(p.b2 + n0 + c0);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const float4 t = __ldg(sv + j), u = __ldg(bv + j);
          f[4 * j] = fmaf(f[4 * j], t.x, u.x); f[4 * j + 1] = fmaf(f[4 * j + 1], t.y, u.y);
          f[4 * j + 2] = fmaf(f[4 * j + 2], t.z, u.z); f[4 * j + 3] = fmaf(f[4 * j + 3], t.w, u.w);
        }
      }
      if (has2) add_res(rb[slot], f);
    }
    if constexpr (STATS) {
      // (compiled only into the STATS instantiations of the kernels: the common path pays nothing for it)
      float g[32];
#pragma unroll
      for (int j = 0; j < 32; ++j) g[j] = valid ? f[j] : 0.f;
      const int copy = (int)((blockIdx.x * 8u + (threadIdx.x >> 5)) % (unsigned)p.stats_copies);
      epi_stats(g, lane, p.stats + (((long long)copy * p.stats_groups + img % p.stats_groups) * p.cout + n0 + c0 + lane) * 2);
    }
    if (st.stage_hi) {
      // the previous bulk store into THIS staging tile must have finished READING it
      const uint32_t alt = (st.chunk & 1u) * st.alt_bytes;
      const uint32_t stage_hi = st.stage_hi + alt, stage_lo = st.stage_lo + alt;
      ++st.chunk;
      if (st.issuer) {
        if (st.alt_bytes) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
        else asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
      }
      asm volatile("bar.sync %0, 128;" ::"r"(st.bar) : "memory");
      if (valid) {
        const uint32_t sw = (uint32_t)((st.row >> 1) & 3);
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          uint32_t h[4], l[4];
#pragma unroll
          for (int m = 0; m < 4; ++m) split2_bf16(f[8 * k + 2 * m], f[8 * k + 2 * m + 1], h[m], l[m]);
          const uint32_t off = (uint32_t)st.row * 64u + (((uint32_t)k ^ sw) << 4);
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_hi + off), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stage_lo + off), "r"(l[0]), "r"(l[1]), "r"(l[2]), "r"(l[3]) : "memory");
        }
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      asm volatile("bar.sync %0, 128;" ::"r"(st.bar + 2) : "memory");
      if (st.issuer && !(p.dbg & (2 | 512))) {      // (512: measurement switch, staging written but never stored)
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(st.map_hi), "r"(stage_hi), "r"(n0 + c0), "r"(st.cx), "r"(st.cy), "r"(img) : "memory");
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];"
                     ::"l"(st.map_lo), "r"(stage_lo), "r"(n0 + c0), "r"(st.cx), "r"(st.cy), "r"(img) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
      continue;
    }
    if (!valid) continue;
    if (p.y && !((p.dbg & 16) && f[0] != 12345.678f)) {
#pragma unroll
      for (int j = 0; j < 32; j += 8) store8(p.y + obase + c0 + j, oplane, f + j);
    }
    if (p.y_f32) {
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(p.y_f32 + obase + c0 + j) = make_float4(f[j], f[j + 1], f[j + 2], f[j + 3]);
    }
  }
}

}  // namespace tcv
