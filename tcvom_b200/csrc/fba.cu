// FBA base-network kernels: weight standardisation, GroupNorm, pooling, bilinear
// resize, channel concatenation, the distance-transform trimap encoding, the fusion head and the EvalModel tail.
// Everything here is HBM-bound byte shuffling around the tcgen05 convolutions (tcv_conv2d): one work item per
// 16-byte channel vector (coalesced NHWC access), no shared-memory staging needed except for the reductions.
// The per-work-item bodies live in fba_body.h so that the CPU test-suite can execute the same index arithmetic.
#include "common.cuh"
#include "fba_body.h"

namespace tcv {

using namespace tcv_fba;

template <typename P, void (*BODY)(ll, const P&)>
__global__ void __launch_bounds__(256) body_kernel(const P p, const ll total) {
  const ll i = (ll)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < total) BODY(i, p);
}

template <typename P, void (*BODY)(ll, const P&)>
static int launch_body(const P& p, ll total, cudaStream_t st, const char* what) {
  if (total <= 0) return TCV_OK;
  const ll blocks = (total + 255) / 256;
  if (blocks > 0x7fffffffLL) return fail(TCV_ERR_INVALID, "%s: too many work items", what);
  body_kernel<P, BODY><<<(unsigned)blocks, 256, 0, st>>>(p, total);
  return launched(what);
}

static bool pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

// ---------------------------------------------------------------------------------- weight standardisation + packing
// one block per output channel (rows co >= cout write zeros)
__global__ void __launch_bounds__(256) ws_pack_kernel(const float* __restrict__ w, int cout, int cin, int taps,
                                                      int standardize, int cin_pad, int cout_pad,
                                                      float* __restrict__ packed) {
  const int co = blockIdx.x;
  const int cnt = cin * taps;
  __shared__ double red[256];
  __shared__ double s_mean, s_inv;
  double mean = 0.0, inv = 1.0;
  if (co < cout && standardize) {
    const float* row = w + (ll)co * cnt;
    double s = 0.0;
    for (int i = threadIdx.x; i < cnt; i += 256) s += row[i];
    red[threadIdx.x] = s;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) s_mean = red[0] / cnt;
    __syncthreads();
    mean = s_mean;
    double ss = 0.0;
    for (int i = threadIdx.x; i < cnt; i += 256) {
      const double d = (double)row[i] - mean;
      ss += d * d;
    }
    red[threadIdx.x] = ss;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    // torch.var: unbiased (cnt - 1); std = sqrt(var + 1e-12) + 1e-5   (layers_WS.py:19-20)
    if (threadIdx.x == 0) s_inv = 1.0 / (sqrt(red[0] / (cnt > 1 ? cnt - 1 : 1) + 1e-12) + 1e-5);
    __syncthreads();
    inv = s_inv;
  }
  // packed[t][ci][co] = w[co][ci][t]
  for (int i = threadIdx.x; i < cin_pad * taps; i += 256) {
    const int ci = i / taps, t = i - ci * taps;
    float v = 0.f;
    if (co < cout && ci < cin) v = (float)(((double)w[((ll)co * cin + ci) * taps + t] - mean) * inv);
    packed[((ll)t * cin_pad + ci) * cout_pad + co] = v;
  }
}

// ---------------------------------------------------------------------------------- GroupNorm statistics
// grid (pixel chunks, images); 256 threads = (256/cv) pixel lanes x cv channel vectors, cv = c/8 <= 256
__global__ void __launch_bounds__(256) gn_stats_kernel(const uint16_t* __restrict__ x, ll x_plane, ll pixels, int c,
                                                       int chunk, double* __restrict__ sums) {
  __shared__ double red[256][17];
  const int cv = c >> 3;
  const int lanes = 256 / cv;
  const int vec = threadIdx.x % cv, lane = threadIdx.x / cv;
  const int img = blockIdx.y;
  const ll p0 = (ll)blockIdx.x * chunk;
  ll p1 = p0 + chunk;
  if (p1 > pixels) p1 = pixels;
  double s[8], ss[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) s[k] = ss[k] = 0.0;
#pragma unroll 4   // several 16-byte load pairs in flight per thread (the loop is latency-bound otherwise)
  for (ll px = p0 + lane; px < p1; px += lanes) {
    float f[8];
    ld8(x + ((ll)img * pixels + px) * c + vec * 8, x_plane, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s[k] += (double)f[k];
      ss[k] += (double)f[k] * (double)f[k];
    }
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red[threadIdx.x][k] = s[k];
    red[threadIdx.x][8 + k] = ss[k];
  }
  __syncthreads();
  // thread t < cv*16 sums one (vector, slot) column over the pixel lanes
  for (int t = threadIdx.x; t < cv * 16; t += 256) {
    const int v = t / 16, slot = t % 16;
    double a = 0.0;
    for (int l = 0; l < lanes; ++l) a += red[l * cv + v][slot];
    const int ch = v * 8 + (slot & 7);
    atomicAdd(sums + ((ll)img * c + ch) * 2 + (slot >> 3), a);
  }
}

// ---------------------------------------------------------------------------------- adaptive average pooling
// grid (s*s, images, c/cw); 256 threads = (256 / (cw/8)) pixel lanes x cw/8 channel vectors.  cw = 64 channels per CTA, or 16
// when the grid would otherwise leave most SMs idle (the 1x1 and 2x2 bins of the pyramid pooling: 96 / 384 CTAs, each walking
// 8 / 2 MB alone -- they took most of the 1.2 ms the four pooling launches of an FBA window cost)
__global__ void __launch_bounds__(256) adaptive_avgpool_kernel(const uint16_t* __restrict__ x, ll x_plane, int n, int h,
                                                               int w, int c, int x_c, int x_off, int s, int cw,
                                                               uint16_t* __restrict__ y) {
  __shared__ float red[256][9];
  const int bi = blockIdx.x / s, bj = blockIdx.x % s;
  const int img = blockIdx.y;
  const int vecs = cw >> 3, lanes = 256 / vecs;
  const int vec = threadIdx.x % vecs, lane = threadIdx.x / vecs;
  const int ch = blockIdx.z * cw + vec * 8;
  const int y0 = bin_start(bi, h, s), y1 = bin_end(bi, h, s);
  const int x0 = bin_start(bj, w, s), x1 = bin_end(bj, w, s);
  const int rw = x1 - x0, cnt = (y1 - y0) * rw;
  float acc[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] = 0.f;
#pragma unroll 4
  for (int i = lane; i < cnt; i += lanes) {
    const int yy = y0 + i / rw, xx = x0 + i % rw;
    float f[8];
    ld8(x + (((ll)img * h + yy) * w + xx) * x_c + x_off + ch, x_plane, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) acc[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = acc[k];
  __syncthreads();
  if (threadIdx.x < cw) {
    const int v = threadIdx.x >> 3, k = threadIdx.x & 7;
    float a = 0.f;
    for (int l = 0; l < lanes; ++l) a += red[l * vecs + v][k];
    const ll yplane = (ll)n * s * s * c;
    st1(y + (((ll)img * s + bi) * s + bj) * c + blockIdx.z * cw + v * 8 + k, yplane, a / (float)cnt);
  }
}

}  // namespace tcv

using namespace tcv;
using namespace tcv_fba;

#define U16(p) reinterpret_cast<uint16_t*>(p)
#define CU16(p) reinterpret_cast<const uint16_t*>(p)

extern "C" {

int tcv_ws_pack(const float* w, int cout, int cin, int kh, int kw, int standardize, int cin_pad, int cout_pad,
                float* packed, tcv_stream_t stream) {
  TCV_REQUIRE(w && packed, "ws_pack: null pointer");
  TCV_REQUIRE(cout > 0 && cin > 0 && kh > 0 && kw > 0 && cin_pad >= cin && cout_pad >= cout, "ws_pack: bad dims");
  ws_pack_kernel<<<cout_pad, 256, 0, S(stream)>>>(w, cout, cin, kh * kw, standardize, cin_pad, cout_pad, packed);
  return launched("ws_pack_kernel");
}

__global__ void __launch_bounds__(256) gn_finalize_warp_kernel(GnFinalizeP p) {
  const ll warp = ((ll)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= (ll)p.n * p.groups) return;
  const int img = (int)(warp / p.groups), g = (int)(warp % p.groups);
  const int cpg = p.c / p.groups;
  double s = 0.0, ss = 0.0;
  for (int cp = 0; cp < p.copies; ++cp)
    for (int k = lane; k < cpg; k += 32) {
      double* q = const_cast<double*>(p.sums) + (((ll)cp * p.n + img) * p.c + g * cpg + k) * 2;
      s += q[0];
      ss += q[1];
      if (p.clear) q[0] = q[1] = 0.0;
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    s += __shfl_xor_sync(0xffffffffu, s, o);
    ss += __shfl_xor_sync(0xffffffffu, ss, o);
  }
  const double cnt = (double)cpg * (double)p.pixels;
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;  // biased, as nn.GroupNorm
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + (double)p.eps);
  for (int k = lane; k < cpg; k += 32) {
    const int ch = g * cpg + k;
    const double sc = (double)p.gamma[ch] * invstd;
    p.scale[(ll)img * p.c + ch] = (float)sc;
    p.shift[(ll)img * p.c + ch] = (float)((double)p.beta[ch] - mean * sc);
  }
}

int tcv_gn_stats(const void* x, long long x_plane, int n, long long pixels, int c, double* sums, tcv_stream_t stream) {
  TCV_REQUIRE(x && sums && n > 0 && pixels > 0, "gn_stats: bad arguments");
  TCV_REQUIRE(c % 8 == 0 && pow2(c / 8) && c / 8 <= 256, "gn_stats: c/8 must be a power of two <= 256 (c=%d)", c);
  if (x_plane == 0) x_plane = (long long)n * pixels * c;
  TCV_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * (size_t)n * c, S(stream)));
  // ~4 blocks per SM over all images; every block owns one contiguous pixel chunk of one image
  int dev = 0, sms = 148;
  if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  long long chunks = ((long long)sms * 4 + n - 1) / n;
  const int lanes = 256 / (c / 8);
  if (chunks * lanes > pixels) chunks = (pixels + lanes - 1) / lanes;
  if (chunks < 1) chunks = 1;
  const int chunk = (int)((pixels + chunks - 1) / chunks);
  dim3 grid((unsigned)((pixels + chunk - 1) / chunk), n);
  gn_stats_kernel<<<grid, 256, 0, S(stream)>>>(CU16(x), x_plane, pixels, c, chunk, sums);
  return launched("gn_stats_kernel");
}

int tcv_gn_finalize(const double* sums, int n, long long pixels, int c, int groups, const float* gamma,
                    const float* beta, float eps, float* scale, float* shift, tcv_stream_t stream) {
  TCV_REQUIRE(sums && gamma && beta && scale && shift, "gn_finalize: null pointer");
  TCV_REQUIRE(n > 0 && pixels > 0 && groups > 0 && c % groups == 0, "gn_finalize: bad dims");
  GnFinalizeP p{sums, n, c, groups, pixels, gamma, beta, eps, scale, shift, 1, 0};
  // one WARP per (image, group): lanes stride the group's channels (the per-work-item body needs 2 x 64 dependent fp64
  // additions in one thread: 15 us per launch, 0.77 ms per window)
  const ll warps = (ll)n * groups;
  gn_finalize_warp_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, S(stream)>>>(p);
  return launched("gn_finalize_kernel");
}

int tcv_gn_finalize_acc(double* sums, int copies, int clear, int n, long long pixels, int c, int groups, const float* gamma,
                        const float* beta, float eps, float* scale, float* shift, tcv_stream_t stream) {
  TCV_REQUIRE(sums && gamma && beta && scale && shift, "gn_finalize_acc: null pointer");
  TCV_REQUIRE(copies > 0 && n > 0 && pixels > 0 && groups > 0 && c % groups == 0, "gn_finalize_acc: bad dims");
  GnFinalizeP p{sums, n, c, groups, pixels, gamma, beta, eps, scale, shift, copies, clear};
  const ll warps = (ll)n * groups;
  gn_finalize_warp_kernel<<<(unsigned)((warps * 32 + 255) / 256), 256, 0, S(stream)>>>(p);
  return launched("gn_finalize_kernel");
}

int tcv_gn_apply(const void* x, long long x_plane, int n, long long pixels, int c, const float* scale,
                 const float* shift, const void* res, long long res_plane, int act, void* y, long long y_plane,
                 int y_c, int y_off, tcv_stream_t stream) {
  TCV_REQUIRE(x && scale && shift && y, "gn_apply: null pointer");
  TCV_REQUIRE(n > 0 && pixels > 0 && c % 8 == 0 && y_c % 8 == 0 && y_off % 8 == 0 && y_off + c <= y_c,
              "gn_apply: bad dims");
  TCV_REQUIRE(act >= TCV_ACT_NONE && act <= TCV_ACT_RELU6, "gn_apply: unknown activation");
  if (x_plane == 0) x_plane = (long long)n * pixels * c;
  if (res && res_plane == 0) res_plane = (long long)n * pixels * c;
  if (y_plane == 0) y_plane = (long long)n * pixels * y_c;
  GnApplyP p{CU16(x), x_plane, n, c, pixels, scale, shift, CU16(res), res_plane, act, U16(y), y_plane, y_c, y_off};
  return launch_body<GnApplyP, gn_apply_body>(p, (ll)n * pixels * (c / 8), S(stream), "gn_apply_kernel");
}

int tcv_maxpool3s2(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && c % 8 == 0, "maxpool3s2: bad arguments");
  PoolP p{CU16(x), n, h, w, c, (h - 1) / 2 + 1, (w - 1) / 2 + 1, U16(y)};
  return launch_body<PoolP, maxpool3s2_body>(p, (ll)n * p.oh * p.ow * (c / 8), S(stream), "maxpool3s2_kernel");
}

int tcv_adaptive_avgpool(const void* x, long long x_plane, int n, int h, int w, int c, int x_c, int x_off, int s,
                         void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && s > 0, "adaptive_avgpool: bad arguments");
  TCV_REQUIRE(c % 64 == 0 && x_c % 8 == 0 && x_off % 8 == 0 && x_off + c <= x_c, "adaptive_avgpool: bad channels");
  if (x_plane == 0) x_plane = (long long)n * h * w * x_c;
  const int cw = ((ll)s * s * n * (c / 64) >= 8 * 148) ? 64 : 16;
  dim3 grid(s * s, n, c / cw);
  adaptive_avgpool_kernel<<<grid, 256, 0, S(stream)>>>(CU16(x), x_plane, n, h, w, c, x_c, x_off, s, cw, U16(y));
  return launched("adaptive_avgpool_kernel");
}

int tcv_bilinear(const void* x, int n, int ih, int iw, int c, void* y, long long y_plane, int oh, int ow, int y_c,
                 int y_off, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && n > 0 && ih > 0 && iw > 0 && oh > 0 && ow > 0, "bilinear: bad arguments");
  TCV_REQUIRE(c % 8 == 0 && y_c % 8 == 0 && y_off % 8 == 0 && y_off + c <= y_c, "bilinear: bad channels");
  if (y_plane == 0) y_plane = (long long)n * oh * ow * y_c;
  BilinearP p{CU16(x), n, ih, iw, c, U16(y), y_plane, oh, ow, y_c, y_off};
  return launch_body<BilinearP, bilinear_body>(p, (ll)n * oh * ow * (c / 8), S(stream), "bilinear_kernel");
}

int tcv_copy_channels(const void* x, long long x_plane, int x_c, int x_off, void* y, long long y_plane, int y_c,
                      int y_off, int c, long long pixels, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && pixels > 0, "copy_channels: bad arguments");
  TCV_REQUIRE(c % 8 == 0 && x_c % 8 == 0 && y_c % 8 == 0 && x_off % 8 == 0 && y_off % 8 == 0 && x_off + c <= x_c &&
                  y_off + c <= y_c, "copy_channels: bad channels");
  if (x_plane == 0) x_plane = pixels * x_c;
  if (y_plane == 0) y_plane = pixels * y_c;
  CopyP p{CU16(x), x_plane, x_c, x_off, U16(y), y_plane, y_c, y_off, c, pixels};
  return launch_body<CopyP, copy_channels_body>(p, pixels * (c / 8), S(stream), "copy_channels_kernel");
}

int tcv_fba_encode_inputs(const void* imgs, const void* tris, int is_u8, int frames, int h, int w, void* x16,
                          tcv_stream_t stream) {
  TCV_REQUIRE(imgs && tris && x16 && frames > 0 && h > 0 && w > 0, "fba_encode_inputs: bad arguments");
  EncodeP p{imgs, tris, is_u8, frames, h, w, U16(x16)};
  return launch_body<EncodeP, fba_encode_body>(p, (ll)frames * h * w, S(stream), "fba_encode_kernel");
}

int tcv_fba_edt_cols(const void* x16, int frames, int h, int w, int* g, tcv_stream_t stream) {
  TCV_REQUIRE(x16 && g && frames > 0 && h > 0 && w > 0, "fba_edt_cols: bad arguments");
  TCV_REQUIRE(h < TCV_EDT_INF && w < TCV_EDT_INF, "fba_edt_cols: image too large");
  EdtP p{U16(const_cast<void*>(x16)), frames, h, w, g};
  return launch_body<EdtP, fba_edt_cols_body>(p, (ll)frames * 2 * w, S(stream), "fba_edt_cols_kernel");
}

// Row pass of the exact Euclidean distance transform, same arithmetic as fba_edt_rows_body (the CPU test double runs that
// body; all comparisons are on integers, so both give identical d^2), organised for the GPU: one CTA per image row, the
// row of vertical distances staged ONCE in shared memory (the body version re-read it from L1/L2 up to 760 times per
// pixel: 2.98 ms per 1080p window), 32-bit arithmetic (r^2 <= 145 000, g^2 <= h^2 < 2^31 / 2).
__global__ void __launch_bounds__(256) fba_edt_rows_smem_kernel(EdtP p) {
  extern __shared__ int grow[];   // w ints
  const int y = blockIdx.x, k = blockIdx.y, f = blockIdx.z;
  const ll hw = (ll)p.h * p.w;
  const int* row = p.g + (((ll)f * 2 + k) * p.h + y) * p.w;
  for (int x = threadIdx.x; x < p.w; x += 256) grow[x] = row[x];
  __syncthreads();
  const int cap2 = TCV_EDT_CAP2;
  const ll plane = (ll)p.frames * hw * 16;
  for (int x = threadIdx.x; x < p.w; x += 256) {
    const int g0 = grow[x];
    // "infinite" (no seed in this column) and anything beyond the cap behave alike: the features are exact zeros
    int best = g0 >= TCV_EDT_INF || g0 > 46340 ? 0x7fffffff : g0 * g0;
    for (int r = 1; r < p.w; ++r) {
      const int r2 = r * r;
      if (r2 >= best || r2 > cap2) break;
      if (x - r >= 0) {
        const int gv = grow[x - r];
        if (gv < 46340) { const int v = r2 + gv * gv; best = v < best ? v : best; }
      }
      if (x + r < p.w) {
        const int gv = grow[x + r];
        if (gv < 46340) { const int v = r2 + gv * gv; best = v < best ? v : best; }
      }
    }
    uint16_t* o = p.x16 + ((ll)f * hw + (ll)y * p.w + x) * 16 + 3 + 3 * k;
    float e[3] = {0.f, 0.f, 0.f};
    if (best <= cap2) {
      const float d = sqrtf((float)best);
      const float m = -(d * d);
      e[0] = expf(m / 81.92f);
      e[1] = expf(m / 1310.72f);
      e[2] = expf(m / 5242.88f);
    }
    for (int j = 0; j < 3; ++j) st1(o + j, plane, e[j]);
  }
}

int tcv_fba_edt_rows(const int* g, int frames, int h, int w, void* x16, tcv_stream_t stream) {
  TCV_REQUIRE(x16 && g && frames > 0 && h > 0 && w > 0, "fba_edt_rows: bad arguments");
  TCV_REQUIRE(h < 32768 && (size_t)w * 4 <= 160 * 1024, "fba_edt_rows: image too large");
  EdtP p{U16(x16), frames, h, w, const_cast<int*>(g)};
  const size_t smem = (size_t)w * sizeof(int);
  if (smem > 48 * 1024)
    TCV_CUDA(cudaFuncSetAttribute(fba_edt_rows_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  fba_edt_rows_smem_kernel<<<dim3(h, 2, frames), 256, smem, S(stream)>>>(p);
  return launched("fba_edt_rows_kernel");
}

int tcv_fba_cat_inputs(const void* x16, long long x16_plane, long long pixels, void* y, long long y_plane, int y_c,
                       int y_off, tcv_stream_t stream) {
  TCV_REQUIRE(x16 && y && pixels > 0 && y_c % 8 == 0 && y_off % 8 == 0 && y_off + 32 <= y_c,
              "fba_cat_inputs: bad arguments");
  if (y_plane == 0) y_plane = pixels * y_c;
  if (x16_plane == 0) x16_plane = pixels * 16;
  CatP p{CU16(x16), x16_plane, pixels, U16(y), y_plane, y_c, y_off};
  return launch_body<CatP, fba_cat_inputs_body>(p, pixels, S(stream), "fba_cat_inputs_kernel");
}

int tcv_fba_fusion(const void* o8, const void* x16, long long x16_plane, long long x16_img_stride, int n, int h, int w,
                   float* pred, tcv_stream_t stream) {
  TCV_REQUIRE(o8 && x16 && pred && n > 0 && h > 0 && w > 0 && x16_plane > 0, "fba_fusion: bad arguments");
  if (x16_img_stride == 0) x16_img_stride = (long long)h * w * 16;
  FusionP p{CU16(o8), CU16(x16), x16_plane, x16_img_stride, n, h, w, pred};
  return launch_body<FusionP, fba_fusion_body>(p, (ll)n * h * w, S(stream), "fba_fusion_kernel");
}

// One thread per four consecutive pixels: the trimap quad is tested first (most of a frame is known foreground / background
// and costs one 4-byte load), the per-pixel body then reads its bytes from lines the warp already pulled into L1.  Sums are
// kept in double per thread, reduced over the warp and the CTA, and leave as one atomic per CTA and slot.
__global__ void __launch_bounds__(256) frame_metrics_kernel(MetricP p, double* __restrict__ out, int vec) {
  __shared__ float lut[256];
  lut[threadIdx.x] = metric_u8((uint8_t)threadIdx.x);
  __syncthreads();
  p.lut = lut;
  double acc[7] = {0, 0, 0, 0, 0, 0, 0};
  const int total = p.h * p.w, quads = (total + 3) >> 2;
  for (int q = blockIdx.x * blockDim.x + threadIdx.x; q < quads; q += gridDim.x * blockDim.x) {
    const int i0 = q << 2;
    if (vec && i0 + 3 < total) {
      const uchar4 t = __ldg(reinterpret_cast<const uchar4*>(p.tri) + q);
      const unsigned char tk[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (tk[k] > 0 && tk[k] < 255) metric_body(i0 + k, p, acc);
    } else {
      for (int k = 0; k < 4 && i0 + k < total; ++k) metric_body(i0 + k, p, acc);
    }
  }
  __shared__ double red[8][7];
#pragma unroll
  for (int k = 0; k < 7; ++k) {
    double v = acc[k];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5][k] = v;
  }
  __syncthreads();
  if (threadIdx.x < 7) {
    double v = 0.0;
#pragma unroll
    for (int wp = 0; wp < 8; ++wp) v += red[wp][threadIdx.x];
    if (v != 0.0) atomicAdd(out + threadIdx.x, v);
  }
}

int tcv_frame_metrics(const uint8_t* alpha, const uint8_t* gt, const uint8_t* tri, const uint8_t* next_alpha,
                      const uint8_t* next_gt, const float* flow, int h, int w, double* out, tcv_stream_t stream) {
  TCV_REQUIRE(alpha && gt && tri && out && h > 1 && w > 1, "frame_metrics: bad arguments");
  TCV_REQUIRE((next_alpha == nullptr) == (next_gt == nullptr) && (!flow || next_alpha), "frame_metrics: the next frame needs both images");
  TCV_CUDA(cudaMemsetAsync(out, 0, 7 * sizeof(double), S(stream)));
  MetricP p{alpha, gt, tri, next_alpha, next_gt, flow, h, w, nullptr};
  TCV_REQUIRE((ll)h * w < (1ll << 31) - 4, "frame_metrics: frame too large");
  const int quads = (h * w + 3) / 4;
  const int blocks = std::min((quads + 255) / 256, 8 * 148);
  frame_metrics_kernel<<<blocks, 256, 0, S(stream)>>>(p, out, (reinterpret_cast<uintptr_t>(tri) & 3) == 0);
  return launched("frame_metrics_kernel");
}

int tcv_dwconv3x3(const void* x, int n, int h, int w, int c, int dil, const float* wt, const float* scale, const float* shift,
                  const float* border, int act, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && wt && scale && shift && y && n > 0 && h > 0 && w > 0 && c % 8 == 0 && dil >= 1, "dwconv3x3: bad arguments");
  TCV_REQUIRE(act >= TCV_ACT_NONE && act <= TCV_ACT_RELU6, "dwconv3x3: unknown activation");
  DwConvP p{CU16(x), n, h, w, c, dil, wt, scale, shift, border, act, U16(y)};
  return launch_body<DwConvP, dwconv3x3_body>(p, (ll)n * h * w * (c / 8), S(stream), "dwconv3x3_kernel");
}

int tcv_index_finish(const void* b0, const void* b1, const void* b2, const void* b3, int n, int h2, int w2, int c,
                     void* idx_en, void* idx_de, tcv_stream_t stream) {
  TCV_REQUIRE(b0 && b1 && b2 && b3 && idx_en && idx_de && n > 0 && h2 > 0 && w2 > 0 && c % 8 == 0, "index_finish: bad arguments");
  IndexFinishP p{{CU16(b0), CU16(b1), CU16(b2), CU16(b3)}, n, h2, w2, c, U16(idx_en), U16(idx_de)};
  return launch_body<IndexFinishP, index_finish_body>(p, (ll)n * h2 * w2 * (c / 8), S(stream), "index_finish_kernel");
}

int tcv_index_pool(const void* x, const void* idx_en, int n, int h, int w, int c, void* masked, void* pooled,
                   tcv_stream_t stream) {
  TCV_REQUIRE(x && idx_en && masked && pooled && n > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "index_pool: bad arguments");
  IndexPoolP p{CU16(x), CU16(idx_en), n, h, w, c, U16(masked), U16(pooled)};
  return launch_body<IndexPoolP, index_pool_body>(p, (ll)n * (h / 2) * (w / 2) * (c / 8), S(stream), "index_pool_kernel");
}

int tcv_index_upcat(const void* dec, int dec_c, int dec_real, int up, const void* idx, int idx_c, long long idx_plane,
                    const void* low, int low_c, long long low_plane, int low_real, int n, int h, int w, int cat_c, void* cat,
                    tcv_stream_t stream) {
  TCV_REQUIRE(dec && low && cat && n > 0 && h > 0 && w > 0, "index_upcat: bad arguments");
  TCV_REQUIRE(dec_c % 8 == 0 && dec_real % 8 == 0 && low_c % 8 == 0 && low_real % 8 == 0 && cat_c % 8 == 0 &&
              dec_real <= dec_c && low_real <= low_c && dec_real + low_real <= cat_c, "index_upcat: bad channel counts");
  TCV_REQUIRE((up == 0 || up == 1) && (up == 0 || (h % 2 == 0 && w % 2 == 0)) && (!idx || (idx_c % 8 == 0 && idx_c >= dec_real)),
              "index_upcat: bad geometry");
  if (idx_plane == 0) idx_plane = (ll)n * h * w * idx_c;
  if (low_plane == 0) low_plane = (ll)n * h * w * low_c;
  IndexUpcatP p{CU16(dec), CU16(idx), CU16(low), n, h, w, up, dec_c, dec_real, idx_c, low_c, low_real, cat_c, U16(cat),
                idx_plane, low_plane};
  return launch_body<IndexUpcatP, index_upcat_body>(p, (ll)n * h * w * (cat_c / 8), S(stream), "index_upcat_kernel");
}

int tcv_maxpool2_idx(const void* x, int n, int h, int w, int c, void* y, uint8_t* idx, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && idx && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "maxpool2_idx: bad arguments");
  Pool2P p{CU16(x), n, h, w, c, U16(y), idx};
  return launch_body<Pool2P, maxpool2_idx_body>(p, (ll)n * (h / 2) * (w / 2) * (c / 8), S(stream), "maxpool2_idx_kernel");
}

int tcv_maxunpool2(const void* x, const uint8_t* idx, int n, int h, int w, int c, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && idx && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "maxunpool2: bad arguments");
  Unpool2P p{CU16(x), idx, n, h, w, c, U16(y)};
  return launch_body<Unpool2P, maxunpool2_body>(p, (ll)n * h * w * (c / 8), S(stream), "maxunpool2_kernel");
}

int tcv_dim_fix_inputs(const void* tris, int is_u8, int frames, int h, int w, void* x8, tcv_stream_t stream) {
  TCV_REQUIRE(tris && x8 && frames > 0 && h > 0 && w > 0, "dim_fix_inputs: bad arguments");
  DimFixP p{tris, is_u8, (ll)frames * h * w, U16(x8)};
  return launch_body<DimFixP, dim_fix_inputs_body>(p, p.pixels, S(stream), "dim_fix_inputs_kernel");
}

int tcv_space_to_depth2(const void* x, long long x_plane, int n, int h, int w, int c, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 8 == 0,
              "space_to_depth2: bad arguments");
  if (x_plane == 0) x_plane = (long long)n * h * w * c;
  S2dP p{CU16(x), x_plane, n, h, w, c, U16(y)};
  return launch_body<S2dP, space_to_depth2_body>(p, (ll)n * h * w * (c / 8), S(stream), "space_to_depth2_kernel");
}

int tcv_s2d_pack_stem(const float* w49, int cin_pad, int cout, float* out, tcv_stream_t stream) {
  TCV_REQUIRE(w49 && out && cin_pad > 0 && cout > 0, "s2d_pack_stem: bad arguments");
  S2dPackP p{w49, cin_pad, cout, out};
  return launch_body<S2dPackP, s2d_pack_stem_body>(p, (ll)16 * 4 * cin_pad * cout, S(stream), "s2d_pack_stem_kernel");
}

static int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }

int tcv_s2d_pack(const float* src, int k, int pad, int cin_src, int cout_src, int cin_dst, int cout_dst, float* out,
                 tcv_stream_t stream) {
  TCV_REQUIRE(src && out && k >= 1 && pad >= 0 && pad < k, "s2d_pack: bad arguments");
  TCV_REQUIRE(cin_src > 0 && cout_src > 0 && cin_dst >= cin_src && cout_dst >= cout_src, "s2d_pack: bad channel counts");
  const int t0 = floordiv2(-pad), T = floordiv2(k - 1 - pad) - t0 + 1;
  S2dPackGenP p{src, k, pad, t0, T, cin_src, cout_src, cin_dst, cout_dst, out};
  return launch_body<S2dPackGenP, s2d_pack_body>(p, (ll)T * T * 4 * cin_dst * cout_dst, S(stream), "s2d_pack_kernel");
}

int tcv_postprocess_eval_fba(const float* pred, const void* imgs, const void* tris, int is_u8, const float* trimask,
                             int batch, int frames, int h, int w, float* alphas, float* Fs, float* Bs,
                             tcv_stream_t stream) {
  TCV_REQUIRE(pred && imgs && tris && trimask && alphas && Fs && Bs, "postprocess_eval_fba: null pointer");
  TCV_REQUIRE(batch > 0 && frames >= 3 && h > 0 && w > 0, "postprocess_eval_fba: bad dims");
  PostP p{pred, imgs, tris, is_u8, trimask, batch, frames, h, w, alphas, Fs, Bs};
  return launch_body<PostP, postprocess_fba_body>(p, (ll)batch * frames * h * w, S(stream), "postprocess_fba_kernel");
}

}  // extern "C"
