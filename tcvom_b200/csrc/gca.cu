// Guided contextual attention (models/GCA/ops.py:106-229) -- operand preparation, masked
// softmax and overlap-add fold.  The two GEMMs (scores = Q.Kn^T, aggregation = P.Vt^T) run in
// gemm_f32.cu (exact path) or gemm_tc.cu (tcgen05 path).
//
//   hh = h/2, ww = w/2 (OS16 grid), P = hh*ww patches, D = 9*64 = 576, DV = 16*128 = 2048.
//   Q[p][(kh*3+kw)*64 + c]  = g_reflect[py+kh-1][px+kw-1][c]                    (ops.py:125-130)
//   Kn[p]                   = Q[p] / max(|Q[p]|, 1e-4) * (mm[p] ? s_u : s_k)    (ops.py:174-175,186)
//   mm[p]                   = any(unknown_os16 in the reflect 3x3 window of p)   (ops.py:148-156)
//   Vt[(ty*4+tx)*128+c][p]  = feat_reflect[2py+ty-1][2px+tx-1][c]               (ops.py:112-118)
// The (kh,kw,c) ordering of the 576/2048 axes differs from the reference's (c,kh,kw); inner
// products and the fold are invariant to that permutation as long as both sides agree.
#include <cuda_fp16.h>

#include "common.cuh"

namespace tcv {

constexpr int GC = 64;     // guidance channels after guidance_conv
constexpr int FC = 128;    // alpha-feature channels
constexpr int QD = 9 * GC; // 576
constexpr int VD = 16 * FC;// 2048

// scales[n] = (clamp(sqrt(um/(1-um)),0.1,10), clamp(sqrt((1-um)/um),0.1,10)), um = mean(unknown[::2,::2])
__global__ void gca_scales_kernel(const float* __restrict__ unknown, int h, int w, float* __restrict__ scales) {
  __shared__ float part[32];
  const int n = blockIdx.x;
  const int hh = h / 2, ww = w / 2;
  const float* u = unknown + (long long)n * h * w;
  float s = 0.f;
  for (int i = threadIdx.x; i < hh * ww; i += blockDim.x) {
    const int y = i / ww, x = i - y * ww;
    s += u[(2 * y) * w + 2 * x];
  }
  s = warp_sum(s);
  if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x < 32) {
    float t = threadIdx.x < (blockDim.x >> 5) ? part[threadIdx.x] : 0.f;
    t = warp_sum(t);
    if (threadIdx.x == 0) {
      const float um = t / (float)(hh * ww);
      const float km = 1.0f - um;
      scales[2 * n + 0] = fminf(fmaxf(sqrtf(um / km), 0.1f), 10.f);
      scales[2 * n + 1] = fminf(fmaxf(sqrtf(km / um), 0.1f), 10.f);
    }
  }
}

// one warp per patch
template <int SPLIT>   // 0: fp32, 2: bf16 hi/lo planes, 3: bf16 hi/mid/lo planes
__global__ void gca_prep_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ unknown, int n,
                                int h, int w, const float* __restrict__ scales, void* __restrict__ Qv,
                                void* __restrict__ Knv, float* __restrict__ mm) {
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n * P) return;
  const int img = warp / P, p = warp - img * P;
  const int py = p / ww, px = p - py * ww;
  const long long gplane = (long long)n * P * GC;
  const __nv_bfloat16* gi = g + (long long)img * P * GC;
  const float* u = unknown + (long long)img * h * w;

  float q[18];  // 576 / 32 : lane owns 2 channels (2*lane, 2*lane+1) of each of the 9 taps
  float ss = 0.f;
  float usum = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = reflect(py + t / 3 - 1, hh), xx = reflect(px + t % 3 - 1, ww);
    const __nv_bfloat16* src = gi + ((long long)yy * ww + xx) * GC + 2 * lane;
    const uint32_t a = *reinterpret_cast<const uint32_t*>(src);
    const uint32_t b = *reinterpret_cast<const uint32_t*>(src + gplane);
    q[2 * t] = __uint_as_float(a << 16) + __uint_as_float(b << 16);
    q[2 * t + 1] = __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u);
    ss += q[2 * t] * q[2 * t] + q[2 * t + 1] * q[2 * t + 1];
    usum += u[(2 * yy) * w + 2 * xx];
  }
  ss = warp_sum(ss);
  const float m = usum > 0.f ? 1.f : 0.f;
  const float scale = m > 0.f ? scales[2 * img] : scales[2 * img + 1];
  const float inv = scale / fmaxf(sqrtf(ss), 1e-4f);
  const long long off = ((long long)img * P + p) * QD;
  if (SPLIT) {
    __nv_bfloat16* qo = reinterpret_cast<__nv_bfloat16*>(Qv) + off;
    __nv_bfloat16* ko = reinterpret_cast<__nv_bfloat16*>(Knv) + off;
    const long long plane = (long long)n * P * QD;
    auto put = [&](__nv_bfloat16* dst, float x0, float x1) {
#pragma unroll
      for (int pl = 0; pl < SPLIT; ++pl) {
        const __nv_bfloat16 b0 = __float2bfloat16_rn(x0), b1 = __float2bfloat16_rn(x1);
        *reinterpret_cast<uint32_t*>(dst + pl * plane) = pack2(b0, b1);
        x0 -= __bfloat162float(b0);
        x1 -= __bfloat162float(b1);
      }
    };
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      put(qo + t * GC + 2 * lane, q[2 * t], q[2 * t + 1]);
      put(ko + t * GC + 2 * lane, q[2 * t] * inv, q[2 * t + 1] * inv);
    }
  } else {
    float* qo = reinterpret_cast<float*>(Qv) + off;
    float* ko = reinterpret_cast<float*>(Knv) + off;
#pragma unroll
    for (int t = 0; t < 9; ++t) {
      *reinterpret_cast<float2*>(qo + t * GC + 2 * lane) = make_float2(q[2 * t], q[2 * t + 1]);
      *reinterpret_cast<float2*>(ko + t * GC + 2 * lane) = make_float2(q[2 * t] * inv, q[2 * t + 1] * inv);
    }
  }
  if (lane == 0) mm[(long long)img * P + p] = m;
}

// store one value in the GEMM operand format `MODE`: 0 fp32, 1 bf16, 2 split-bf16 (lo plane `plane`
// elements later), 3 fp16
template <int MODE>
__device__ __forceinline__ void store_operand(void* base, long long i, long long plane, float v) {
  if (MODE == 0) reinterpret_cast<float*>(base)[i] = v;
  else if (MODE == 1) reinterpret_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v);
  else if (MODE == 2) store1(reinterpret_cast<__nv_bfloat16*>(base) + i, plane, v);
  else reinterpret_cast<__half*>(base)[i] = __float2half_rn(v);
}

// Block = (64 consecutive patches p, one of the 16 taps, image): pixel rows are read coalesced (128
// channels = 256 B per plane), transposed through shared memory, and written as 64-patch row segments.
template <int MODE>
__global__ void __launch_bounds__(256) gca_values_kernel(const __nv_bfloat16* __restrict__ feat, int n, int h, int w,
                                                         int P_pad, void* __restrict__ Vt) {
  __shared__ float tile[64][FC + 1];
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  const int p0 = blockIdx.x * 64, t = blockIdx.y, img = blockIdx.z;
  const long long fplane = (long long)n * h * w * FC;
  for (int i = threadIdx.x; i < 64 * (FC / 8); i += 256) {
    const int pp = i / (FC / 8), c8 = (i % (FC / 8)) * 8;
    const int p = p0 + pp;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p < P) {
      const int py = p / ww, px = p - py * ww;
      const int yy = reflect(2 * py + t / 4 - 1, h), xx = reflect(2 * px + t % 4 - 1, w);
      load8(feat + (((long long)img * h + yy) * w + xx) * FC + c8, fplane, f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[pp][c8 + k] = f[k];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)n * VD * P_pad;
  for (int c = warp; c < FC; c += 8) {
    const long long o = ((long long)img * VD + t * FC + c) * P_pad + p0;
    store_operand<MODE>(Vt, o + lane, total, tile[lane][c]);
    store_operand<MODE>(Vt, o + lane + 32, total, tile[lane + 32][c]);
  }
}

// one CTA per (row q, image): masked softmax over keys with the row cached in shared memory
// (one HBM read of the fp32 logits, one write of the operand-format probabilities)
template <int MODE>
__global__ void __launch_bounds__(256) gca_softmax_kernel(float* __restrict__ S, const float* __restrict__ mm, int P,
                                                          int P_pad, void* __restrict__ Pb, long long plane) {
  extern __shared__ float srow[];   // P_pad floats
  __shared__ float red[32];
  __shared__ float bcast;
  const int q = blockIdx.x, img = blockIdx.y;
  float* row = S + ((long long)img * P + q) * P_pad;
  const float diag = -1e4f * mm[(long long)img * P + q];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int p4 = threadIdx.x * 4; p4 < P_pad; p4 += 1024) {
    float4 v = *reinterpret_cast<const float4*>(row + p4);
    float e[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int p = p4 + k;
      if (p == q) e[k] += diag;
      if (p >= P) e[k] = -INFINITY;
      mx = fmaxf(mx, e[k]);
    }
    *reinterpret_cast<float4*>(srow + p4) = make_float4(e[0], e[1], e[2], e[3]);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float t = lane < 8 ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) bcast = t;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int p4 = threadIdx.x * 4; p4 < P_pad; p4 += 1024) {
    float4 v = *reinterpret_cast<const float4*>(srow + p4);
    v.x = expf(v.x - mx); v.y = expf(v.y - mx); v.z = expf(v.z - mx); v.w = expf(v.w - mx);   // exp(-inf) = 0 on pads
    sum += (v.x + v.y) + (v.z + v.w);
    *reinterpret_cast<float4*>(srow + p4) = v;
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float t = lane < 8 ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) bcast = t;
  }
  __syncthreads();
  const float inv = 1.0f / bcast;
  const long long o = ((long long)img * P + q) * P_pad;
  for (int p4 = threadIdx.x * 4; p4 < P_pad; p4 += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(srow + p4);
    const float e[4] = {v.x * inv, v.y * inv, v.z * inv, v.w * inv};
    if (MODE == 0) {
      *reinterpret_cast<float4*>(row + p4) = make_float4(e[0], e[1], e[2], e[3]);
    } else if (MODE == 2) {
      __nv_bfloat16 hb[4], lb[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) split_bf16(e[k], hb[k], lb[k]);
      __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(Pb) + o + p4;
      *reinterpret_cast<uint2*>(dst) = make_uint2(pack2(hb[0], hb[1]), pack2(hb[2], hb[3]));
      *reinterpret_cast<uint2*>(dst + plane) = make_uint2(pack2(lb[0], lb[1]), pack2(lb[2], lb[3]));
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) store_operand<MODE>(Pb, o + p4 + k, plane, e[k]);
    }
  }
}

// Y[y][x][c] = 1/4 * sum over (ty,tx) with (y+1-ty), (x+1-tx) even and in range of O[q][(ty*4+tx)*128+c]
__global__ void gca_fold_kernel(const float* __restrict__ O, int n, int h, int w, __nv_bfloat16* __restrict__ Y) {
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  const long long total = (long long)n * h * w * (FC / 4);
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (FC / 4)) * 4;
  long long t = i / (FC / 4);
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int img = (int)(t / h);
  float acc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int a = 0; a < 2; ++a) {
    const int ty = ((y + 1) & 1) + 2 * a;
    const int qy = (y + 1 - ty) / 2;
    if (y + 1 - ty < 0 || qy >= hh) continue;
#pragma unroll
    for (int b = 0; b < 2; ++b) {
      const int tx = ((x + 1) & 1) + 2 * b;
      const int qx = (x + 1 - tx) / 2;
      if (x + 1 - tx < 0 || qx >= ww) continue;
      const float4 o = *reinterpret_cast<const float4*>(O + ((long long)img * P + qy * ww + qx) * VD +
                                                        (ty * 4 + tx) * FC + c);
      acc[0] += o.x; acc[1] += o.y; acc[2] += o.z; acc[3] += o.w;
    }
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) acc[k] *= 0.25f;
  store4(Y + (((long long)img * h + y) * w + x) * FC + c, (long long)n * h * w * FC, acc);
}

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_gca_prep(const void* g, const float* unknown, int n, int h, int w, void* Q, void* Kn, float* mm,
                 float* scales, int bf16_split, tcv_stream_t stream) {
  TCV_REQUIRE(g && unknown && Q && Kn && mm && scales, "gca_prep: null pointer");
  TCV_REQUIRE(n > 0 && h >= 4 && w >= 4 && h % 2 == 0 && w % 2 == 0, "gca_prep: h,w must be even and >= 4");
  gca_scales_kernel<<<n, 256, 0, S(stream)>>>(unknown, h, w, scales);
  int rc = launched("gca_scales_kernel");
  if (rc) return rc;
  const long long warps = (long long)n * (h / 2) * (w / 2);
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  auto G = reinterpret_cast<const __nv_bfloat16*>(g);
  TCV_REQUIRE(bf16_split == 0 || bf16_split == 2 || bf16_split == 3 || bf16_split == 1, "gca_prep: planes must be 0, 2 or 3");
  if (bf16_split == 3)
    gca_prep_kernel<3><<<grid, 256, 0, S(stream)>>>(G, unknown, n, h, w, scales, Q, Kn, mm);
  else if (bf16_split)
    gca_prep_kernel<2><<<grid, 256, 0, S(stream)>>>(G, unknown, n, h, w, scales, Q, Kn, mm);
  else
    gca_prep_kernel<0><<<grid, 256, 0, S(stream)>>>(G, unknown, n, h, w, scales, Q, Kn, mm);
  return launched("gca_prep_kernel");
}

int tcv_gca_values(const void* feat, int n, int h, int w, void* Vt, int mode, tcv_stream_t stream) {
  TCV_REQUIRE(feat && Vt, "gca_values: null pointer");
  TCV_REQUIRE(h % 2 == 0 && w % 2 == 0 && h >= 4 && w >= 4, "gca_values: h,w must be even and >= 4");
  const int P = (h / 2) * (w / 2), P_pad = (P + 63) / 64 * 64;
  const long long total = (long long)n * VD * P_pad;
  (void)total;
  const dim3 grid(P_pad / 64, 16, n);
  TCV_REQUIRE(mode >= 0 && mode <= 3, "gca_values: mode must be 0..3");
  auto F = reinterpret_cast<const __nv_bfloat16*>(feat);
  switch (mode) {
    case 0: gca_values_kernel<0><<<grid, 256, 0, S(stream)>>>(F, n, h, w, P_pad, Vt); break;
    case 1: gca_values_kernel<1><<<grid, 256, 0, S(stream)>>>(F, n, h, w, P_pad, Vt); break;
    case 2: gca_values_kernel<2><<<grid, 256, 0, S(stream)>>>(F, n, h, w, P_pad, Vt); break;
    default: gca_values_kernel<3><<<grid, 256, 0, S(stream)>>>(F, n, h, w, P_pad, Vt); break;
  }
  return launched("gca_values_kernel");
}

int tcv_gca_softmax(float* Sm, const float* mm, int n, int P, int P_pad, void* P_out, int mode,
                    tcv_stream_t stream) {
  TCV_REQUIRE(Sm && mm, "gca_softmax: null pointer");
  TCV_REQUIRE(P > 0 && P_pad >= P, "gca_softmax: bad P");
  dim3 grid(P, n);
  TCV_REQUIRE(mode >= 0 && mode <= 3 && (mode == 0 || P_out), "gca_softmax: bad mode / missing output");
  TCV_REQUIRE(P_pad % 4 == 0 && (size_t)P_pad * 4 <= 200 * 1024, "gca_softmax: row of %d keys does not fit shared memory", P_pad);
  const long long plane = (long long)n * P * P_pad;
  const size_t smem = (size_t)P_pad * sizeof(float);
#define TCV_SM_LAUNCH(M, OUT)                                                                               \
  do {                                                                                                      \
    TCV_CUDA(cudaFuncSetAttribute(gca_softmax_kernel<M>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
    gca_softmax_kernel<M><<<grid, 256, smem, S(stream)>>>(Sm, mm, P, P_pad, OUT, plane);                      \
  } while (0)
  switch (mode) {
    case 0: TCV_SM_LAUNCH(0, nullptr); break;
    case 1: TCV_SM_LAUNCH(1, P_out); break;
    case 2: TCV_SM_LAUNCH(2, P_out); break;
    default: TCV_SM_LAUNCH(3, P_out); break;
  }
#undef TCV_SM_LAUNCH
  return launched("gca_softmax_kernel");
}

int tcv_gca_fold(const float* O, int n, int h, int w, void* Y, tcv_stream_t stream) {
  TCV_REQUIRE(O && Y, "gca_fold: null pointer");
  const long long total = (long long)n * h * w * (FC / 4);
  gca_fold_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(O, n, h, w,
                                                                         reinterpret_cast<__nv_bfloat16*>(Y));
  return launched("gca_fold_kernel");
}

}  // extern "C"

// =====================================================================================================================
// Shift-sum form of the aggregation  Y = fold(A.V; k4,s2,p1)/4  (ops.py:112-118,204).
//
// Tap k = 2a + r (a, r in {0,1} per axis) of patch q = (qy,qx) lands on output row 2(qy+ay)+ry-1 and reads, through A[q,p],
// feature row 2(py+ay)+ry-1 of the reflect-padded feature.  With m = q + a and p' = p + a on the (hh+1) x (ww+1) grid:
//     Y[2my+ry-1][2mx+rx-1][c] = 1/4 * sum_p' A2[m][p'] * F_r[p'][c]
//     A2[m][p'] = sum_{a in {0,1}^2} A[m-a][p'-a]            (terms outside the hh x ww patch grid are zero)
//     F_r[p'][c] = feat_reflect[2p'y+ry-1][2p'x+rx-1][c]     (4 parities x 128 channels = 512 columns)
// i.e. ONE [Pk x Pk].[Pk x 512] GEMM (Pk = (hh+1)(ww+1)) instead of [P x P].[P x 2048] followed by an overlap-add:
// 3.8x fewer FLOPs, every output pixel written exactly once.  Exact (a re-association of the same sums).
//
// To make "p' - a" a plain column shift the KEYS live on the padded grid: Kn row index = py*(ww+1)+px with zero rows at
// px == ww / py == hh, so S = Q.Kn^T is [P][ld] (ld = Pk rounded up to 64) and A2[m][j] = sum_a A[m-a][j - (ay*(ww+1)+ax)].
// =====================================================================================================================
namespace tcv {

// zero rows of Kn at the pad positions of the (hh+1) x (ww+1) key grid: one warp per (pad row, plane, image)
__global__ void gca_kn_pad_zero_kernel(__nv_bfloat16* __restrict__ Kn, int n, int hh, int ww, int planes) {
  const int Pk = (hh + 1) * (ww + 1), npad = hh + ww + 1;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= npad * planes * n) return;
  const int i = warp % npad, pl = (warp / npad) % planes, img = warp / (npad * planes);
  const int row = i < hh ? i * (ww + 1) + ww : hh * (ww + 1) + (i - hh);
  uint4* dst = reinterpret_cast<uint4*>(Kn + ((long long)(pl * n + img) * Pk + row) * QD);
  for (int k = lane; k < QD * 2 / 16; k += 32) dst[k] = make_uint4(0, 0, 0, 0);
}

// one warp per patch; like gca_prep_kernel<2> but Kn rows on the padded key grid
__global__ void gca_prep_grid_kernel(const __nv_bfloat16* __restrict__ g, const float* __restrict__ unknown, int n, int h,
                                     int w, const float* __restrict__ scales, __nv_bfloat16* __restrict__ Qo,
                                     __nv_bfloat16* __restrict__ Ko, float* __restrict__ mm) {
  const int hh = h / 2, ww = w / 2, P = hh * ww, Pk = (hh + 1) * (ww + 1);
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= n * P) return;
  const int img = warp / P, p = warp - img * P;
  const int py = p / ww, px = p - py * ww;
  const long long gplane = (long long)n * P * GC;
  const __nv_bfloat16* gi = g + (long long)img * P * GC;
  const float* u = unknown + (long long)img * h * w;
  float q[18];
  float ss = 0.f, usum = 0.f;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    const int yy = reflect(py + t / 3 - 1, hh), xx = reflect(px + t % 3 - 1, ww);
    const __nv_bfloat16* src = gi + ((long long)yy * ww + xx) * GC + 2 * lane;
    const uint32_t a = *reinterpret_cast<const uint32_t*>(src);
    const uint32_t b = *reinterpret_cast<const uint32_t*>(src + gplane);
    q[2 * t] = __uint_as_float(a << 16) + __uint_as_float(b << 16);
    q[2 * t + 1] = __uint_as_float(a & 0xffff0000u) + __uint_as_float(b & 0xffff0000u);
    ss += q[2 * t] * q[2 * t] + q[2 * t + 1] * q[2 * t + 1];
    usum += u[(2 * yy) * w + 2 * xx];
  }
  ss = warp_sum(ss);
  const float m = usum > 0.f ? 1.f : 0.f;
  const float scale = m > 0.f ? scales[2 * img] : scales[2 * img + 1];
  const float inv = scale / fmaxf(sqrtf(ss), 1e-4f);
  __nv_bfloat16* qo = Qo + ((long long)img * P + p) * QD;
  __nv_bfloat16* ko = Ko + ((long long)img * Pk + py * (ww + 1) + px) * QD;
  const long long qplane = (long long)n * P * QD, kplane = (long long)n * Pk * QD;
#pragma unroll
  for (int t = 0; t < 9; ++t) {
    uint32_t hi, lo;
    split2_bf16(q[2 * t], q[2 * t + 1], hi, lo);
    *reinterpret_cast<uint32_t*>(qo + t * GC + 2 * lane) = hi;
    *reinterpret_cast<uint32_t*>(qo + qplane + t * GC + 2 * lane) = lo;
    split2_bf16(q[2 * t] * inv, q[2 * t + 1] * inv, hi, lo);
    *reinterpret_cast<uint32_t*>(ko + t * GC + 2 * lane) = hi;
    *reinterpret_cast<uint32_t*>(ko + kplane + t * GC + 2 * lane) = lo;
  }
  if (lane == 0) mm[(long long)img * P + p] = m;
}

// Ft[img][(ry*2+rx)*128 + c][p'] = feat_reflect[2p'y+ry-1][2p'x+rx-1][c] on the (hh+1)x(ww+1) grid, split-bf16 planes
// [2][n][512][ld].  Block = (64 consecutive p', parity r, image): coalesced pixel reads, smem transpose, 64-p' row segments.
__global__ void __launch_bounds__(256) gca_values_parity_kernel(const __nv_bfloat16* __restrict__ feat, int n, int h,
                                                                int w, int ld, __nv_bfloat16* __restrict__ Ft) {
  __shared__ float tile[64][FC + 1];
  const int hh = h / 2, ww = w / 2, Pk = (hh + 1) * (ww + 1);
  const int p0 = blockIdx.x * 64, r = blockIdx.y, img = blockIdx.z;
  const int ry = r >> 1, rx = r & 1;
  const long long fplane = (long long)n * h * w * FC;
  for (int i = threadIdx.x; i < 64 * (FC / 8); i += 256) {
    const int pp = i / (FC / 8), c8 = (i % (FC / 8)) * 8;
    const int p = p0 + pp;
    float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (p < Pk) {
      const int py = p / (ww + 1), px = p - py * (ww + 1);
      const int yy = reflect(2 * py + ry - 1, h), xx = reflect(2 * px + rx - 1, w);
      load8(feat + (((long long)img * h + yy) * w + xx) * FC + c8, fplane, f);
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) tile[pp][c8 + k] = f[k];
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long long total = (long long)n * 4 * FC * ld;
  for (int c = warp; c < FC; c += 8) {
    const long long o = ((long long)img * 4 * FC + r * FC + c) * ld + p0;
    uint32_t hi, lo;
    split2_bf16(tile[2 * lane][c], tile[2 * lane + 1][c], hi, lo);
    *reinterpret_cast<uint32_t*>(Ft + o + 2 * lane) = hi;
    *reinterpret_cast<uint32_t*>(Ft + total + o + 2 * lane) = lo;
  }
}

// validity of key-grid column j (a real patch, not a pad position) and the logit with the self-mask applied
struct KeyGrid {
  int ww1, Pv;   // ww + 1 ; hh * (ww+1): columns >= Pv are the pad row / beyond the grid
  __device__ __forceinline__ bool valid(int j) const { return j >= 0 && j < Pv && (j % ww1) != ww1 - 1; }
};

// stats[img][q] = (row max, 1 / sum exp) of S[q][:] - 1e4*[col == q]*mm[q] over the valid key columns; one CTA per row
template <bool NORMALISE>
__global__ void __launch_bounds__(256) gca_rowstats_kernel(const float* __restrict__ S, const float* __restrict__ mm,
                                                           int hh, int ww, int ld, float2* __restrict__ stats) {
  extern __shared__ float srow[];   // ld floats
  __shared__ float red[8];
  __shared__ float bcast;
  const int P = hh * ww;
  const int q = blockIdx.x, img = blockIdx.y;
  const KeyGrid kg{ww + 1, hh * (ww + 1)};
  const int qj = (q / ww) * (ww + 1) + (q % ww);
  const float* row = S + ((long long)img * P + q) * ld;
  const float diag = -1e4f * mm[(long long)img * P + q];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float mx = -INFINITY;
  for (int j4 = threadIdx.x * 4; j4 < ld; j4 += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(row + j4);
    float e[4] = {v.x, v.y, v.z, v.w};
    int col = j4 % kg.ww1;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int j = j4 + k;
      if (j == qj) e[k] += diag;
      if (j >= kg.Pv || col == ww) e[k] = -INFINITY;
      if (++col == kg.ww1) col = 0;
      mx = fmaxf(mx, e[k]);
    }
    *reinterpret_cast<float4*>(srow + j4) = make_float4(e[0], e[1], e[2], e[3]);
  }
  mx = warp_max(mx);
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (warp == 0) {
    float t = lane < 8 ? red[lane] : -INFINITY;
    t = warp_max(t);
    if (lane == 0) bcast = t;
  }
  __syncthreads();
  mx = bcast;
  float sum = 0.f;
  for (int j4 = threadIdx.x * 4; j4 < ld; j4 += 1024) {
    float4 v = *reinterpret_cast<const float4*>(srow + j4);
    v.x = expf(v.x - mx); v.y = expf(v.y - mx); v.z = expf(v.z - mx); v.w = expf(v.w - mx);   // exp(-inf) = 0 at the pads
    sum += (v.x + v.y) + (v.z + v.w);
    if (NORMALISE) *reinterpret_cast<float4*>(srow + j4) = v;     // each thread re-reads only what it wrote itself
  }
  sum = warp_sum(sum);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (warp == 0) {
    float t = lane < 8 ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) {
      stats[(long long)img * P + q] = make_float2(mx, 1.0f / t);
      bcast = 1.0f / t;
    }
  }
  if (!NORMALISE) return;
  __syncthreads();
  const float inv = bcast;
  float* out = const_cast<float*>(row);
  for (int j4 = threadIdx.x * 4; j4 < ld; j4 += 1024) {
    const float4 v = *reinterpret_cast<const float4*>(srow + j4);
    *reinterpret_cast<float4*>(out + j4) = make_float4(v.x * inv, v.y * inv, v.z * inv, v.w * inv);
  }
}

// A2[img][m][j] = sum_a A[m - a][j - shift_a] from the normalised probabilities (gca_rowstats_kernel<true> wrote them in
// place of S, zeros at the pad columns): 4 coalesced row reads (L2: a row is re-read by its 4 consumers within 2 key-grid
// rows), one split-bf16 write.  One CTA per (row m, image).
__global__ void __launch_bounds__(256) gca_shift_add_kernel(const float* __restrict__ A, int n, int hh, int ww, int ld,
                                                            __nv_bfloat16* __restrict__ A2) {
  const int P = hh * ww, ww1 = ww + 1, Pk = (hh + 1) * ww1;
  const int m = blockIdx.x, img = blockIdx.y;
  const int my = m / ww1, mx = m - my * ww1;
  const float* rows[4];
  int sh[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int qy = my - (a >> 1), qx = mx - (a & 1);
    const bool ok = qy >= 0 && qy < hh && qx >= 0 && qx < ww;
    sh[a] = (a >> 1) * ww1 + (a & 1);
    rows[a] = ok ? A + ((long long)img * P + qy * ww + qx) * ld - sh[a] : nullptr;
  }
  const long long plane = (long long)n * Pk * ld;
  __nv_bfloat16* out = A2 + ((long long)img * Pk + m) * ld;
  // four consecutive outputs per thread; a source row shifted by sh is read with the widest loads its alignment allows
  // (sh % 4 == 0: one 16-byte load; == 2: two 8-byte loads; odd: 4 + 8 + 4 bytes)
#pragma unroll 2
  for (int j = threadIdx.x * 4; j < ld; j += 1024) {
    float v[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (rows[a] == nullptr) continue;
      const float* src = rows[a] + j;               // element for output column j (source column j - sh)
      if (j >= sh[a]) {
        const int al = sh[a] & 3;
        if (al == 0) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src));
          v[0] += t.x; v[1] += t.y; v[2] += t.z; v[3] += t.w;
        } else if (al == 2) {
          const float2 t0 = __ldg(reinterpret_cast<const float2*>(src)), t1 = __ldg(reinterpret_cast<const float2*>(src + 2));
          v[0] += t0.x; v[1] += t0.y; v[2] += t1.x; v[3] += t1.y;
        } else {
          const float t0 = __ldg(src);
          const float2 t1 = __ldg(reinterpret_cast<const float2*>(src + 1));
          const float t3 = __ldg(src + 3);
          v[0] += t0; v[1] += t1.x; v[2] += t1.y; v[3] += t3;
        }
      } else {                                      // the first columns of the row: part of the window is left of the row
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e >= sh[a]) v[e] += __ldg(src + e);
      }
    }
    uint32_t h0, l0, h1, l1;
    split2_bf16(v[0], v[1], h0, l0);
    split2_bf16(v[2], v[3], h1, l1);
    *reinterpret_cast<uint2*>(out + j) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(out + plane + j) = make_uint2(l0, l1);
  }
}

// A2[img][m][j] = sum_a softmax(S)[m - a][j - shift_a], split-bf16 planes [2][n][Pk][ld], straight from the logits and the
// row statistics (no normalised copy of S in HBM).  Same streaming structure as gca_shift_add_kernel -- one CTA per
// (row m, image), threads stride over the columns, the <= 4 source rows are read as long contiguous streams (L2: a row is
// re-read by its 4 consumers within 2 key-grid rows) -- plus the exponential of every operand.
__global__ void __launch_bounds__(256) gca_softmax_shift_kernel(const float* __restrict__ S, const float2* __restrict__ stats,
                                                                const float* __restrict__ mm, int n, int hh, int ww, int ld,
                                                                __nv_bfloat16* __restrict__ A2) {
  const int P = hh * ww, ww1 = ww + 1, Pk = (hh + 1) * ww1, Pv = hh * ww1;
  const int m = blockIdx.x, img = blockIdx.y;
  const int my = m / ww1, mx = m - my * ww1;
  const float* rows[4];
  float mxv[4], inv[4], dg[4];
  int sh[4], qj[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int qy = my - (a >> 1), qx = mx - (a & 1);
    const bool ok = qy >= 0 && qy < hh && qx >= 0 && qx < ww;
    sh[a] = (a >> 1) * ww1 + (a & 1);
    rows[a] = nullptr;
    mxv[a] = inv[a] = dg[a] = 0.f;
    qj[a] = -1;
    if (ok) {
      const long long q = (long long)img * P + qy * ww + qx;
      rows[a] = S + q * ld - sh[a];
      const float2 st = __ldg(stats + q);
      mxv[a] = st.x;
      inv[a] = st.y;
      dg[a] = -1e4f * __ldg(mm + q);
      qj[a] = qy * ww1 + qx + sh[a];          // output column whose operand from row a is the self-masked logit
    }
  }
  const long long plane = (long long)n * Pk * ld;
  __nv_bfloat16* out = A2 + ((long long)img * Pk + m) * ld;
#pragma unroll 4
  for (int j = threadIdx.x * 2; j < ld; j += 512) {
    float v0 = 0.f, v1 = 0.f;
    // column validity of the operand at output column c shifted by sh: c - sh is a real key
    const int c0 = j % ww1;                 // (j - sh[a]) % ww1 takes the values c0, c0 - 1 (mod ww1) only
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (rows[a] == nullptr) continue;
      const int s0 = j - sh[a], s1 = s0 + 1;
      const int col0 = (a & 1) ? (c0 == 0 ? ww : c0 - 1) : c0;
      const int col1 = col0 == ww ? 0 : col0 + 1;
      if (s0 >= 0 && s0 < Pv && col0 != ww)
        v0 += __expf(__ldg(rows[a] + j) + (j == qj[a] ? dg[a] : 0.f) - mxv[a]) * inv[a];
      if (s1 >= 0 && s1 < Pv && col1 != ww)
        v1 += __expf(__ldg(rows[a] + j + 1) + (j + 1 == qj[a] ? dg[a] : 0.f) - mxv[a]) * inv[a];
    }
    uint32_t hi, lo;
    split2_bf16(v0, v1, hi, lo);
    *reinterpret_cast<uint32_t*>(out + j) = hi;
    *reinterpret_cast<uint32_t*>(out + plane + j) = lo;
  }
}

// Y[img][2my+ry-1][2mx+rx-1][c] = O2[img][m][(ry*2+rx)*128 + c] / 4   (every output pixel exactly once)
__global__ void gca_unfold_parity_kernel(const float* __restrict__ O2, int n, int h, int w, __nv_bfloat16* __restrict__ Y) {
  const int ww1 = w / 2 + 1, Pk = (h / 2 + 1) * ww1;
  const long long total = (long long)n * h * w * (FC / 4);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (FC / 4)) * 4;
  long long t = i / (FC / 4);
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int img = (int)(t / h);
  const int ry = (y + 1) & 1, rx = (x + 1) & 1;
  const int my = (y + 1 - ry) >> 1, mx = (x + 1 - rx) >> 1;
  const float4 o = *reinterpret_cast<const float4*>(O2 + ((long long)img * Pk + my * ww1 + mx) * (4 * FC) +
                                                    (ry * 2 + rx) * FC + c);
  const float acc[4] = {o.x * 0.25f, o.y * 0.25f, o.z * 0.25f, o.w * 0.25f};
  store4(Y + (((long long)img * h + y) * w + x) * FC + c, (long long)n * h * w * FC, acc);
}

}  // namespace tcv

extern "C" {

int tcv_gca_prep_grid(const void* g, const float* unknown, int n, int h, int w, void* Q, void* Kn, float* mm,
                      float* scales, tcv_stream_t stream) {
  TCV_REQUIRE(g && unknown && Q && Kn && mm && scales, "gca_prep_grid: null pointer");
  TCV_REQUIRE(n > 0 && h >= 4 && w >= 4 && h % 2 == 0 && w % 2 == 0, "gca_prep_grid: h,w must be even and >= 4");
  gca_scales_kernel<<<n, 256, 0, S(stream)>>>(unknown, h, w, scales);
  int rc = launched("gca_scales_kernel");
  if (rc) return rc;
  const int hh = h / 2, ww = w / 2;
  const long long pad_warps = (long long)(hh + ww + 1) * 2 * n;
  gca_kn_pad_zero_kernel<<<(unsigned)((pad_warps * 32 + 255) / 256), 256, 0, S(stream)>>>(
      reinterpret_cast<__nv_bfloat16*>(Kn), n, hh, ww, 2);
  rc = launched("gca_kn_pad_zero_kernel");
  if (rc) return rc;
  const long long warps = (long long)n * hh * ww;
  gca_prep_grid_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, S(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(g), unknown, n, h, w, scales, reinterpret_cast<__nv_bfloat16*>(Q),
      reinterpret_cast<__nv_bfloat16*>(Kn), mm);
  return launched("gca_prep_grid_kernel");
}

int tcv_gca_values_parity(const void* feat, int n, int h, int w, int ld, void* Ft, tcv_stream_t stream) {
  TCV_REQUIRE(feat && Ft, "gca_values_parity: null pointer");
  TCV_REQUIRE(n > 0 && h % 2 == 0 && w % 2 == 0 && h >= 4 && w >= 4, "gca_values_parity: h,w must be even and >= 4");
  const int Pk = (h / 2 + 1) * (w / 2 + 1);
  TCV_REQUIRE(ld % 64 == 0 && ld >= Pk, "gca_values_parity: ld must be a multiple of 64 and >= (h/2+1)*(w/2+1)");
  gca_values_parity_kernel<<<dim3(ld / 64, 4, n), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(feat), n, h, w,
                                                                      ld, reinterpret_cast<__nv_bfloat16*>(Ft));
  return launched("gca_values_parity_kernel");
}

int tcv_gca_rowstats(float* Sm, const float* mm, int n, int h, int w, int ld, float* stats, int normalise,
                     tcv_stream_t stream) {
  TCV_REQUIRE(Sm && mm && stats, "gca_rowstats: null pointer");
  const int hh = h / 2, ww = w / 2;
  TCV_REQUIRE(n > 0 && hh > 0 && ww > 0 && ld % 4 == 0 && ld >= (hh + 1) * (ww + 1), "gca_rowstats: bad geometry");
  TCV_REQUIRE((size_t)ld * 4 <= 200 * 1024, "gca_rowstats: row of %d keys does not fit shared memory", ld);
  const size_t smem = (size_t)ld * sizeof(float);
  TCV_CUDA(cudaFuncSetAttribute(gca_rowstats_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  TCV_CUDA(cudaFuncSetAttribute(gca_rowstats_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  if (normalise)
    gca_rowstats_kernel<true><<<dim3(hh * ww, n), 256, smem, S(stream)>>>(Sm, mm, hh, ww, ld, reinterpret_cast<float2*>(stats));
  else
    gca_rowstats_kernel<false><<<dim3(hh * ww, n), 256, smem, S(stream)>>>(Sm, mm, hh, ww, ld, reinterpret_cast<float2*>(stats));
  return launched("gca_rowstats_kernel");
}

int tcv_gca_shift_add(const float* A, int n, int h, int w, int ld, void* A2, tcv_stream_t stream) {
  TCV_REQUIRE(A && A2, "gca_shift_add: null pointer");
  const int hh = h / 2, ww = w / 2, Pk = (hh + 1) * (ww + 1);
  TCV_REQUIRE(n > 0 && hh > 0 && ww > 0 && ld % 4 == 0 && ld >= Pk, "gca_shift_add: bad geometry (ld %% 4)");
  gca_shift_add_kernel<<<dim3(Pk, n), 256, 0, S(stream)>>>(A, n, hh, ww, ld, reinterpret_cast<__nv_bfloat16*>(A2));
  return launched("gca_shift_add_kernel");
}

int tcv_gca_softmax_shift(const float* Sm, const float* stats, const float* mm, int n, int h, int w, int ld, void* A2,
                          tcv_stream_t stream) {
  TCV_REQUIRE(Sm && stats && mm && A2, "gca_softmax_shift: null pointer");
  const int hh = h / 2, ww = w / 2, Pk = (hh + 1) * (ww + 1);
  TCV_REQUIRE(n > 0 && hh > 0 && ww > 0 && ld % 64 == 0 && ld >= Pk, "gca_softmax_shift: bad geometry (ld %% 64)");
  gca_softmax_shift_kernel<<<dim3(Pk, n), 256, 0, S(stream)>>>(Sm, reinterpret_cast<const float2*>(stats), mm, n, hh, ww, ld,
                                                               reinterpret_cast<__nv_bfloat16*>(A2));
  return launched("gca_softmax_shift_kernel");
}

int tcv_gca_unfold_parity(const float* O2, int n, int h, int w, void* Y, tcv_stream_t stream) {
  TCV_REQUIRE(O2 && Y, "gca_unfold_parity: null pointer");
  TCV_REQUIRE(n > 0 && h % 2 == 0 && w % 2 == 0, "gca_unfold_parity: h,w must be even");
  const long long total = (long long)n * h * w * (FC / 4);
  gca_unfold_parity_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(O2, n, h, w,
                                                                                  reinterpret_cast<__nv_bfloat16*>(Y));
  return launched("gca_unfold_parity_kernel");
}

}  // extern "C"
