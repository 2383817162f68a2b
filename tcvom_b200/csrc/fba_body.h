// Per-work-item bodies of the FBA-path kernels (fba.cu) as host/device inline functions.
//
// One work item = what one CUDA thread does.  fba.cu wraps every body in a __global__ kernel
// (`i = blockIdx.x * blockDim.x + threadIdx.x; if (i < total) body(i, p)`); the test double of the C ABI under
// tests/host_emul/ compiles THIS header with g++ and runs `for (i = 0; i < total; ++i) body(i, p)`, so the index
// arithmetic of these kernels is exercised by the CPU test-suite against the oracle (there is no GPU in the build
// container).  Only the 16-byte vector load/store differs between the two builds (ld8 / st8 below).
//
// Layout: split-bf16 NHWC, hi plane at p, lo plane `plane` elements later, value = hi + lo.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#ifdef __CUDACC__
#define TCV_HD __host__ __device__ __forceinline__
#else
#define TCV_HD static inline
#endif

namespace tcv_fba {

typedef long long ll;

TCV_HD float bf_to_f(uint16_t b) {
  const uint32_t u = (uint32_t)b << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}
// round to nearest even, the rounding of __float2bfloat16_rn (finite values; NaN stays NaN)
TCV_HD uint16_t f_to_bf(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}
TCV_HD float ld1(const uint16_t* p, ll plane) { return bf_to_f(p[0]) + bf_to_f(p[plane]); }
TCV_HD void st1(uint16_t* p, ll plane, float v) {
  const uint16_t h = f_to_bf(v);
  p[0] = h;
  p[plane] = f_to_bf(v - bf_to_f(h));
}
// 8 consecutive channels (16-byte aligned)
TCV_HD void ld8(const uint16_t* p, ll plane, float* f) {
#ifdef __CUDA_ARCH__
  const uint4 a = *reinterpret_cast<const uint4*>(p);
  const uint4 b = *reinterpret_cast<const uint4*>(p + plane);
  const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    f[2 * i] = __uint_as_float(aw[i] << 16) + __uint_as_float(bw[i] << 16);
    f[2 * i + 1] = __uint_as_float(aw[i] & 0xffff0000u) + __uint_as_float(bw[i] & 0xffff0000u);
  }
#else
  for (int i = 0; i < 8; ++i) f[i] = ld1(p + i, plane);
#endif
}
// eight consecutive fp32 parameters (per-channel scale / shift, 32-byte aligned: channel offsets are multiples of 8) as two
// 16-byte loads.  Eight scalar loads cost a warp 8 requests x 32 sectors (the lanes are 32 bytes apart): on a 64-byte-per-
// thread streaming kernel that L1 traffic, not DRAM, set the speed (ncu: 30 sectors per load request in gn_apply).
TCV_HD void ldf8(const float* p, float* f) {
#ifdef __CUDA_ARCH__
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
#else
  for (int i = 0; i < 8; ++i) f[i] = p[i];
#endif
}

TCV_HD void st8(uint16_t* p, ll plane, const float* f) {
#ifdef __CUDA_ARCH__
  uint32_t h[4], l[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint16_t h0 = f_to_bf(f[2 * i]), h1 = f_to_bf(f[2 * i + 1]);
    const uint16_t l0 = f_to_bf(f[2 * i] - bf_to_f(h0)), l1 = f_to_bf(f[2 * i + 1] - bf_to_f(h1));
    h[i] = (uint32_t)h0 | ((uint32_t)h1 << 16);
    l[i] = (uint32_t)l0 | ((uint32_t)l1 << 16);
  }
  *reinterpret_cast<uint4*>(p) = make_uint4(h[0], h[1], h[2], h[3]);
  *reinterpret_cast<uint4*>(p + plane) = make_uint4(l[0], l[1], l[2], l[3]);
#else
  for (int i = 0; i < 8; ++i) st1(p + i, plane, f[i]);
#endif
}

TCV_HD float act_fn(float t, int act) {
  switch (act) {
    case 1: return t > 0.f ? t : 0.f;                      // TCV_ACT_RELU
    case 2: return t > 0.f ? t : 0.2f * t;                 // TCV_ACT_LEAKY02
    case 3: return (tanhf(t) + 1.0f) * 0.5f;               // TCV_ACT_TANH01
    case 4: return t > 0.f ? t : 0.01f * t;                // TCV_ACT_LEAKY001
    case 5: return t < 0.f ? 0.f : (t > 1.f ? 1.f : t);    // TCV_ACT_CLAMP01
    case 6: return t < 0.f ? 0.f : (t > 6.f ? 6.f : t);    // TCV_ACT_RELU6
    default: return t;
  }
}

// ------------------------------------------------------------------------------------------ GroupNorm
struct GnFinalizeP {
  const double* sums;  // [copies][n][c][2]
  int n, c, groups;
  ll pixels;
  const float* gamma;
  const float* beta;
  float eps;
  float* scale;  // [n][c]
  float* shift;
  int copies, clear;   // accumulator copies to sum (>= 1); clear: zero them after reading
};
// work item = (image, group); total = n * groups
TCV_HD void gn_finalize_body(ll i, const GnFinalizeP& p) {
  const int img = (int)(i / p.groups), g = (int)(i % p.groups);
  const int cpg = p.c / p.groups;
  double s = 0.0, ss = 0.0;
  for (int cp = 0; cp < p.copies; ++cp)
    for (int k = 0; k < cpg; ++k) {
      double* q = const_cast<double*>(p.sums) + (((ll)cp * p.n + img) * p.c + g * cpg + k) * 2;
      s += q[0];
      ss += q[1];
      if (p.clear) q[0] = q[1] = 0.0;
    }
  const double cnt = (double)cpg * (double)p.pixels;
  const double mean = s / cnt;
  double var = ss / cnt - mean * mean;  // biased, as nn.GroupNorm
  if (var < 0.0) var = 0.0;
  const double invstd = 1.0 / sqrt(var + (double)p.eps);
  for (int k = 0; k < cpg; ++k) {
    const int ch = g * cpg + k;
    const double sc = (double)p.gamma[ch] * invstd;
    p.scale[(ll)img * p.c + ch] = (float)sc;
    p.shift[(ll)img * p.c + ch] = (float)((double)p.beta[ch] - mean * sc);
  }
}

struct GnApplyP {
  const uint16_t* x;
  ll x_plane;
  int n, c;
  ll pixels;
  const float* scale;
  const float* shift;
  const uint16_t* res;
  ll res_plane;
  int act;
  uint16_t* y;
  ll y_plane;
  int y_c, y_off;
};
// work item = 8 channels of one pixel; total = n * pixels * c / 8
TCV_HD void gn_apply_body(ll i, const GnApplyP& p) {
  const int cv = p.c / 8;
  ll pix;
  int ch, img;
  if ((ll)p.n * p.pixels * cv < (1ll << 32)) {
    // 32-bit index arithmetic: three 64-bit divisions (~100 instructions each on the GPU) made this 64-byte-per-thread
    // streaming kernel instruction-bound at 55 % of the copy bandwidth
    const unsigned iu = (unsigned)i, pu = iu / (unsigned)cv;
    ch = (int)(iu - pu * (unsigned)cv) * 8;
    img = (int)(pu / (unsigned)p.pixels);
    pix = pu;
  } else {
    pix = i / cv;
    ch = (int)(i % cv) * 8;
    img = (int)(pix / p.pixels);
  }
  float f[8];
  ld8(p.x + pix * p.c + ch, p.x_plane, f);
  float sc[8], sh[8];
  ldf8(p.scale + (ll)img * p.c + ch, sc);
  ldf8(p.shift + (ll)img * p.c + ch, sh);
  for (int k = 0; k < 8; ++k) f[k] = f[k] * sc[k] + sh[k];
  if (p.res) {
    float r[8];
    ld8(p.res + pix * p.c + ch, p.res_plane, r);
    for (int k = 0; k < 8; ++k) f[k] += r[k];
  }
  for (int k = 0; k < 8; ++k) f[k] = act_fn(f[k], p.act);
  st8(p.y + pix * p.y_c + p.y_off + ch, p.y_plane, f);
}

// ------------------------------------------------------------------------------------------ pooling / resize / copies
struct PoolP {
  const uint16_t* x;
  int n, h, w, c, oh, ow;
  uint16_t* y;
};
// 3x3 / stride 2 / pad 1 max pooling; work item = 8 channels of one output pixel; total = n*oh*ow*c/8
TCV_HD void maxpool3s2_body(ll i, const PoolP& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int ox = (int)(t % p.ow);
  t /= p.ow;
  const int oy = (int)(t % p.oh);
  const int img = (int)(t / p.oh);
  const ll xplane = (ll)p.n * p.h * p.w * p.c, yplane = (ll)p.n * p.oh * p.ow * p.c;
  float m[8];
  for (int k = 0; k < 8; ++k) m[k] = -INFINITY;
  for (int dy = -1; dy <= 1; ++dy) {
    const int iy = 2 * oy + dy;
    if (iy < 0 || iy >= p.h) continue;
    for (int dx = -1; dx <= 1; ++dx) {
      const int ix = 2 * ox + dx;
      if (ix < 0 || ix >= p.w) continue;
      float f[8];
      ld8(p.x + (((ll)img * p.h + iy) * p.w + ix) * p.c + ch, xplane, f);
      for (int k = 0; k < 8; ++k) m[k] = f[k] > m[k] ? f[k] : m[k];
    }
  }
  st8(p.y + (((ll)img * p.oh + oy) * p.ow + ox) * p.c + ch, yplane, m);
}

// bins of nn.AdaptiveAvgPool2d: [floor(i*in/out), ceil((i+1)*in/out))
TCV_HD int bin_start(int i, int in, int out) { return (int)(((ll)i * in) / out); }
TCV_HD int bin_end(int i, int in, int out) { return (int)((((ll)i + 1) * in + out - 1) / out); }

struct BilinearP {
  const uint16_t* x;
  int n, ih, iw, c;
  uint16_t* y;
  ll y_plane;
  int oh, ow, y_c, y_off;
};
// torch's area_pixel_compute_source_index + guard_index_and_lambda (align_corners=False, float32 arithmetic)
TCV_HD void src_index(int dst, int in, int out, int* i0, int* i1, float* l0, float* l1) {
  const float scale = (float)in / (float)out;
  float s = scale * ((float)dst + 0.5f) - 0.5f;
  if (s < 0.f) s = 0.f;
  int a = (int)s;
  if (a > in - 1) a = in - 1;
  float lam = s - (float)a;
  lam = lam < 0.f ? 0.f : (lam > 1.f ? 1.f : lam);
  *i0 = a;
  *i1 = a + (a < in - 1 ? 1 : 0);
  *l1 = lam;
  *l0 = 1.f - lam;
}
// work item = 8 channels of one output pixel; total = n*oh*ow*c/8
TCV_HD void bilinear_body(ll i, const BilinearP& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int ox = (int)(t % p.ow);
  t /= p.ow;
  const int oy = (int)(t % p.oh);
  const int img = (int)(t / p.oh);
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  src_index(oy, p.ih, p.oh, &y0, &y1, &ly0, &ly1);
  src_index(ox, p.iw, p.ow, &x0, &x1, &lx0, &lx1);
  const ll xplane = (ll)p.n * p.ih * p.iw * p.c;
  const uint16_t* b = p.x + (ll)img * p.ih * p.iw * p.c + ch;
  float a00[8], a01[8], a10[8], a11[8], o[8];
  ld8(b + ((ll)y0 * p.iw + x0) * p.c, xplane, a00);
  ld8(b + ((ll)y0 * p.iw + x1) * p.c, xplane, a01);
  ld8(b + ((ll)y1 * p.iw + x0) * p.c, xplane, a10);
  ld8(b + ((ll)y1 * p.iw + x1) * p.c, xplane, a11);
  for (int k = 0; k < 8; ++k)
    o[k] = ly0 * (lx0 * a00[k] + lx1 * a01[k]) + ly1 * (lx0 * a10[k] + lx1 * a11[k]);
  st8(p.y + (((ll)img * p.oh + oy) * p.ow + ox) * p.y_c + p.y_off + ch, p.y_plane, o);
}

struct CopyP {
  const uint16_t* x;
  ll x_plane;
  int x_c, x_off;
  uint16_t* y;
  ll y_plane;
  int y_c, y_off, c;
  ll pixels;
};
// work item = 8 channels of one pixel; total = pixels * c / 8.  Bit copy of both planes.
TCV_HD void copy_channels_body(ll i, const CopyP& p) {
  const int cv = p.c / 8;
  const ll pix = i / cv;
  const int ch = (int)(i % cv) * 8;
  const uint16_t* s = p.x + pix * p.x_c + p.x_off + ch;
  uint16_t* d = p.y + pix * p.y_c + p.y_off + ch;
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
  *reinterpret_cast<uint4*>(d + p.y_plane) = *reinterpret_cast<const uint4*>(s + p.x_plane);
#else
  for (int k = 0; k < 8; ++k) {
    d[k] = s[k];
    d[k + p.y_plane] = s[k + p.x_plane];
  }
#endif
}

// ------------------------------------------------------------------------------------------ FBA input encoding
struct EncodeP {
  const void* imgs;  // [F,3,H,W] BGR
  const void* tris;  // [F,1,H,W]
  int is_u8, frames, h, w;
  uint16_t* x16;
};
TCV_HD float in_val(const void* p, int is_u8, ll idx) {
  return is_u8 ? (float)reinterpret_cast<const uint8_t*>(p)[idx] : reinterpret_cast<const float*>(p)[idx];
}
// work item = one pixel of one frame; total = F*H*W
TCV_HD void fba_encode_body(ll i, const EncodeP& p) {
  const ll hw = (ll)p.h * p.w;
  const int f = (int)(i / hw);
  const ll pix = i % hw;
  const ll plane = (ll)p.frames * hw * 16;
  uint16_t* o = p.x16 + i * 16;
  const float scale = 1.0f / 255;  // IMG_SCALE (models/model.py:34)
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int c = 0; c < 3; ++c) {
    // .flip([2]): output channel c (RGB) reads input channel 2-c (BGR)   (models/model.py:366-367)
    const float s = in_val(p.imgs, p.is_u8, ((ll)f * 3 + (2 - c)) * hw + pix) * scale;
    st1(o + c, plane, (s - mean[c]) / stdv[c]);
    st1(o + 11 + c, plane, s);
  }
  const float t = in_val(p.tris, p.is_u8, (ll)f * hw + pix) * scale;
  st1(o + 9, plane, t == 0.f ? 1.f : 0.f);   // trimap2b (models/model.py:381)
  st1(o + 10, plane, t == 1.f ? 1.f : 0.f);  // trimap2f (models/model.py:380)
  st1(o + 14, plane, 0.f);
  st1(o + 15, plane, 0.f);
}

#define TCV_EDT_INF (1 << 20)
#define TCV_EDT_CAP2 145000
struct EdtP {
  uint16_t* x16;
  int frames, h, w;
  int* g;  // [F][2][H][W]
};
// work item = one column of one (frame, k); total = F*2*W.  Two sweeps: distance to the nearest seed above, below.
TCV_HD void fba_edt_cols_body(ll i, const EdtP& p) {
  const int x = (int)(i % p.w);
  const int k = (int)((i / p.w) % 2);
  const int f = (int)(i / (2 * (ll)p.w));
  const ll hw = (ll)p.h * p.w;
  const uint16_t* src = p.x16 + (ll)f * hw * 16 + 9 + k;  // hi plane is exact for 0 / 1
  int* g = p.g + ((ll)f * 2 + k) * hw;
  int d = TCV_EDT_INF;
  for (int y = 0; y < p.h; ++y) {
    const bool seed = bf_to_f(src[((ll)y * p.w + x) * 16]) != 0.f;
    d = seed ? 0 : (d >= TCV_EDT_INF ? TCV_EDT_INF : d + 1);
    g[(ll)y * p.w + x] = d;
  }
  d = TCV_EDT_INF;
  for (int y = p.h - 1; y >= 0; --y) {
    const int cur = g[(ll)y * p.w + x];
    d = cur == 0 ? 0 : (d >= TCV_EDT_INF ? TCV_EDT_INF : d + 1);
    if (d < cur) g[(ll)y * p.w + x] = d;
  }
}
// work item = one pixel of one (frame, k); total = F*2*H*W.  Exact lower envelope by outward search:
// a column x' can only improve the minimum while (x-x')^2 < best.
TCV_HD void fba_edt_rows_body(ll i, const EdtP& p) {
  const ll hw = (ll)p.h * p.w;
  const int x = (int)(i % p.w);
  const int y = (int)((i / p.w) % p.h);
  const int k = (int)((i / hw) % 2);
  const int f = (int)(i / (2 * hw));
  const int* row = p.g + (((ll)f * 2 + k) * p.h + y) * p.w;
  const ll inf2 = (ll)TCV_EDT_INF * TCV_EDT_INF;
  // beyond d^2 = TCV_EDT_CAP2 even the widest feature exp(-d^2 / 5242.88) is below 1e-12 (exp(-27.7)): such pixels
  // are written as exact zeros and the search stops there (columns further out cannot bring d^2 under the cap)
  const ll cap2 = TCV_EDT_CAP2;
  ll best = row[x] >= TCV_EDT_INF ? inf2 : (ll)row[x] * row[x];
  for (int r = 1; r < p.w; ++r) {
    if ((ll)r * r >= best || (ll)r * r > cap2) break;
    if (x - r >= 0) {
      const int gv = row[x - r];
      if (gv < TCV_EDT_INF) {
        const ll v = (ll)r * r + (ll)gv * gv;
        best = v < best ? v : best;
      }
    }
    if (x + r < p.w) {
      const int gv = row[x + r];
      if (gv < TCV_EDT_INF) {
        const ll v = (ll)r * r + (ll)gv * gv;
        best = v < best ? v : best;
      }
    }
  }
  uint16_t* o = p.x16 + ((ll)f * hw + (ll)y * p.w + x) * 16 + 3 + 3 * k;
  const ll plane = (ll)p.frames * hw * 16;
  float e[3] = {0.f, 0.f, 0.f};
  if (best <= cap2) {
    const float d = sqrtf((float)best);  // cv2.distanceTransform(DIST_L2, precise) returns the float32 distance
    const float m = -(d * d);           // -dt(...)**2   (utils/utils.py:33)
    // 2*(f*L)^2 with L = 320: 81.92, 1310.72, 5242.88 (utils/utils.py:34-37)
    e[0] = expf(m / 81.92f);
    e[1] = expf(m / 1310.72f);
    e[2] = expf(m / 5242.88f);
  }
  for (int j = 0; j < 3; ++j) st1(o + j, plane, e[j]);
}

struct CatP {
  const uint16_t* x16;
  ll x16_plane;
  ll pixels;
  uint16_t* y;
  ll y_plane;
  int y_c, y_off;
};
// work item = one pixel; channels y_off.. = (normalised RGB, RGB/255, trimap2 bg, fg), then 24 zero channels
TCV_HD void fba_cat_inputs_body(ll i, const CatP& p) {
  const ll xplane = p.x16_plane;
  const uint16_t* s = p.x16 + i * 16;
  const int src[8] = {0, 1, 2, 11, 12, 13, 9, 10};
  float f[8];
  for (int k = 0; k < 8; ++k) f[k] = ld1(s + src[k], xplane);
  uint16_t* d = p.y + i * p.y_c + p.y_off;
  st8(d, p.y_plane, f);
  for (int k = 0; k < 8; ++k) f[k] = 0.f;
  st8(d + 8, p.y_plane, f);
  st8(d + 16, p.y_plane, f);
  st8(d + 24, p.y_plane, f);
}

struct FusionP {
  const uint16_t* o8;
  const uint16_t* x16;
  ll x16_plane, x16_img_stride;
  int n, h, w;
  float* pred;  // [n,7,H,W]
};
TCV_HD float clamp01(float v) { return v < 0.f ? 0.f : (v > 1.f ? 1.f : v); }
// work item = one pixel; VMN_FBA.py:51-57 + FBA/models.py:246-255
TCV_HD void fba_fusion_body(ll i, const FusionP& p) {
  const ll hw = (ll)p.h * p.w;
  const int img_i = (int)(i / hw);
  const ll pix = i % hw;
  float o[8];
  ld8(p.o8 + i * 8, (ll)p.n * hw * 8, o);
  const uint16_t* xs = p.x16 + (ll)img_i * p.x16_img_stride + pix * 16 + 11;
  float img[3], F[3], B[3];
  for (int c = 0; c < 3; ++c) {
    img[c] = ld1(xs + c, p.x16_plane);
    F[c] = 1.0f / (1.0f + expf(-o[1 + c]));
    B[c] = 1.0f / (1.0f + expf(-o[4 + c]));
  }
  float alpha = clamp01(o[0]);
  float num = 0.f, den = 0.f;
  for (int c = 0; c < 3; ++c) {
    const float Fn = alpha * img[c] + (1.f - alpha * alpha) * F[c] - alpha * (1.f - alpha) * B[c];
    const float Bn = (1.f - alpha) * img[c] + (2.f * alpha - alpha * alpha) * B[c] - alpha * (1.f - alpha) * Fn;
    F[c] = clamp01(Fn);
    B[c] = clamp01(Bn);
  }
  for (int c = 0; c < 3; ++c) {
    num += (img[c] - B[c]) * (F[c] - B[c]);
    den += (F[c] - B[c]) * (F[c] - B[c]);
  }
  const float la = 0.1f;
  alpha = clamp01((alpha * la + num) / (den + la));
  float* out = p.pred + (ll)img_i * 7 * hw + pix;
  out[0] = alpha;
  for (int c = 0; c < 3; ++c) {
    out[(1 + c) * hw] = F[c];
    out[(4 + c) * hw] = B[c];
  }
}

struct PostP {
  const float* pred;  // [B*(S-2),7,H,W]
  const void* imgs;
  const void* tris;
  int is_u8;
  const float* trimask;  // [B,S,H,W]
  int batch, frames, h, w;
  float* alphas;  // [B,S,1,H,W]
  float* Fs;      // [B,S,3,H,W]
  float* Bs;
};
// work item = one pixel of one frame; total = B*S*H*W   (models/model.py:426-446)
TCV_HD void postprocess_fba_body(ll i, const PostP& p) {
  const ll hw = (ll)p.h * p.w;
  const ll fr = i / hw, pix = i % hw;
  const int s = (int)(fr % p.frames), b = (int)(fr / p.frames);
  float a = 0.f, F[3] = {0.f, 0.f, 0.f}, B[3] = {0.f, 0.f, 0.f};
  if (s > 0 && s < p.frames - 1) {
    const float scale = 1.0f / 255;
    const bool m = p.trimask[i] != 0.f;
    const float* pr = p.pred + ((ll)b * (p.frames - 2) + (s - 1)) * 7 * hw + pix;
    a = m ? pr[0] : in_val(p.tris, p.is_u8, i) * scale;
    for (int c = 0; c < 3; ++c) {
      const float im = in_val(p.imgs, p.is_u8, (fr * 3 + (2 - c)) * hw + pix) * scale;
      F[c] = m ? pr[(1 + c) * hw] : im;
      B[c] = m ? pr[(4 + c) * hw] : im;
    }
  }
  p.alphas[i] = a;
  for (int c = 0; c < 3; ++c) {
    p.Fs[(fr * 3 + c) * hw + pix] = F[c];
    p.Bs[(fr * 3 + c) * hw + pix] = B[c];
  }
}

// ------------------------------------------------------------------------------------------ 7x7 / stride-2 stem on the tensor cores
// A k=7, s=2, p=3 convolution over c channels equals a k=4, s=1 convolution (taps -2..1) over the 2x2
// space-to-depth image with 4c channels: input row 2*oy + ky - 3 = 2*(oy + ty) + py with ty = floor((ky-3)/2),
// py = (ky-3) mod 2.  For the 16-channel FBA input that is ONE 16-tap, K = 1024 tcgen05 launch instead of four
// CUDA-core launches.
struct S2dP {
  const uint16_t* x;
  ll x_plane;
  int n, h, w, c;  // h, w even
  uint16_t* y;     // dense [n, h/2, w/2, 4c], channel (py*2+px)*c + ch
};
// work item = 8 channels of one input pixel; total = n*h*w*c/8
TCV_HD void space_to_depth2_body(ll i, const S2dP& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int x = (int)(t % p.w);
  t /= p.w;
  const int y = (int)(t % p.h);
  const int img = (int)(t / p.h);
  const int oh = p.h / 2, ow = p.w / 2;
  const ll yplane = (ll)p.n * oh * ow * 4 * p.c;
  const uint16_t* s = p.x + (((ll)img * p.h + y) * p.w + x) * p.c + ch;
  uint16_t* d = p.y + (((ll)img * oh + (y >> 1)) * ow + (x >> 1)) * (4 * p.c) + ((y & 1) * 2 + (x & 1)) * p.c + ch;
#ifdef __CUDA_ARCH__
  *reinterpret_cast<uint4*>(d) = *reinterpret_cast<const uint4*>(s);
  *reinterpret_cast<uint4*>(d + yplane) = *reinterpret_cast<const uint4*>(s + p.x_plane);
#else
  for (int k = 0; k < 8; ++k) {
    d[k] = s[k];
    d[k + yplane] = s[k + p.x_plane];
  }
#endif
}

struct S2dPackP {
  const float* w49;  // packed 7x7 weights [49][cin_pad][cout]
  int cin_pad, cout;
  float* out;        // [16][4*cin_pad][cout], tap t = (ty+2)*4 + (tx+2)
};
// work item = one output element; total = 16 * 4*cin_pad * cout
TCV_HD void s2d_pack_stem_body(ll i, const S2dPackP& p) {
  const int co = (int)(i % p.cout);
  const int ch = (int)((i / p.cout) % (4 * p.cin_pad));
  const int t = (int)(i / ((ll)p.cout * 4 * p.cin_pad));
  const int ty = t / 4 - 2, tx = t % 4 - 2;
  const int q = ch / p.cin_pad, c = ch % p.cin_pad;
  const int ky = 2 * ty + (q >> 1) + 3, kx = 2 * tx + (q & 1) + 3;
  float v = 0.f;
  if (ky >= 0 && ky < 7 && kx >= 0 && kx < 7) v = p.w49[((ll)(ky * 7 + kx) * p.cin_pad + c) * p.cout + co];
  p.out[i] = v;
}

// generic form of the rewrite: a k x k / stride-2 convolution with zero padding `pad` (or pad = 0 on a pre-padded input)
// equals a T x T / stride-1 convolution with tap offsets t0 .. t0+T-1 over the 2x2 space-to-depth image:
// input row 2*oy + ky - pad = 2*(oy + ty) + py  =>  ky = 2*ty + py + pad,  t0 = floor(-pad / 2),  T = floor((k-1-pad)/2) - t0 + 1
struct S2dPackGenP {
  const float* src;  // packed k x k weights [k*k][cin_src][cout_src]
  int k, pad, t0, T, cin_src, cout_src, cin_dst, cout_dst;
  float* out;        // [T*T][4*cin_dst][cout_dst], tap (ty-t0)*T + (tx-t0), row (py*2+px)*cin_dst + c; zero padding elsewhere
};
// work item = one output element; total = T*T * 4*cin_dst * cout_dst
TCV_HD void s2d_pack_body(ll i, const S2dPackGenP& p) {
  const int co = (int)(i % p.cout_dst);
  const int ch = (int)((i / p.cout_dst) % (4 * p.cin_dst));
  const int t = (int)(i / ((ll)p.cout_dst * 4 * p.cin_dst));
  const int ty = t / p.T + p.t0, tx = t % p.T + p.t0;
  const int q = ch / p.cin_dst, c = ch % p.cin_dst;
  const int ky = 2 * ty + (q >> 1) + p.pad, kx = 2 * tx + (q & 1) + p.pad;
  float v = 0.f;
  if (ky >= 0 && ky < p.k && kx >= 0 && kx < p.k && c < p.cin_src && co < p.cout_src)
    v = p.src[((ll)(ky * p.k + kx) * p.cin_src + c) * p.cout_src + co];
  p.out[i] = v;
}

// ------------------------------------------------------------------------------------------ DIM (VMN_DIM.py)
struct Pool2P {
  const uint16_t* x;  // dense [n, h, w, c]
  int n, h, w, c;
  uint16_t* y;        // dense [n, h/2, w/2, c]
  uint8_t* idx;       // [n, h/2, w/2, c]: ky*2 + kx of the first maximum (torch scans the window row-major, keeps on ties)
};
// work item = 8 channels of one pooled pixel; total = n*(h/2)*(w/2)*c/8
TCV_HD void maxpool2_idx_body(ll i, const Pool2P& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int oh = p.h / 2, ow = p.w / 2;
  const int x = (int)(t % ow);
  t /= ow;
  const int y = (int)(t % oh);
  const int img = (int)(t / oh);
  const ll xplane = (ll)p.n * p.h * p.w * p.c, yplane = (ll)p.n * oh * ow * p.c;
  float best[8];
  int bi[8];
  for (int k = 0; k < 4; ++k) {
    const uint16_t* s = p.x + (((ll)img * p.h + 2 * y + (k >> 1)) * p.w + 2 * x + (k & 1)) * p.c + ch;
    float f[8];
    ld8(s, xplane, f);
    for (int j = 0; j < 8; ++j)
      if (k == 0 || f[j] > best[j] || (f[j] != f[j] && best[j] == best[j])) {   // torch: (val > max) || isnan(val)
        best[j] = f[j];
        bi[j] = k;
      }
  }
  const ll o = (((ll)img * oh + y) * ow + x) * p.c + ch;
#ifdef __CUDA_ARCH__
  // hi + lo is exact in fp32, so re-splitting stores the same VALUE (the pair itself can differ in a rounding tie)
  st8(p.y + o, yplane, best);
  uint32_t lo4 = 0, hi4 = 0;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    lo4 |= (uint32_t)bi[j] << (8 * j);
    hi4 |= (uint32_t)bi[j + 4] << (8 * j);
  }
  *reinterpret_cast<uint2*>(p.idx + o) = make_uint2(lo4, hi4);
#else
  // the maximum is one of the stored values: copy its two planes
  for (int j = 0; j < 8; ++j) {
    const uint16_t* s = p.x + (((ll)img * p.h + 2 * y + (bi[j] >> 1)) * p.w + 2 * x + (bi[j] & 1)) * p.c + ch + j;
    p.y[o + j] = s[0];
    p.y[o + j + yplane] = s[xplane];
    p.idx[o + j] = (uint8_t)bi[j];
  }
#endif
}

struct Unpool2P {
  const uint16_t* x;   // dense [n, h/2, w/2, c]
  const uint8_t* idx;  // [n, h/2, w/2, c]
  int n, h, w, c;      // OUTPUT size
  uint16_t* y;         // dense [n, h, w, c]
};
// work item = 8 channels of one OUTPUT pixel; total = n*h*w*c/8
TCV_HD void maxunpool2_body(ll i, const Unpool2P& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int x = (int)(t % p.w);
  t /= p.w;
  const int y = (int)(t % p.h);
  const int img = (int)(t / p.h);
  const int ih = p.h / 2, iw = p.w / 2;
  const ll xplane = (ll)p.n * ih * iw * p.c, yplane = (ll)p.n * p.h * p.w * p.c;
  const ll s = (((ll)img * ih + (y >> 1)) * iw + (x >> 1)) * p.c + ch;
  const ll o = (((ll)img * p.h + y) * p.w + x) * p.c + ch;
  const int me = (y & 1) * 2 + (x & 1);
#ifdef __CUDA_ARCH__
  const uint2 id = *reinterpret_cast<const uint2*>(p.idx + s);
  const uint4 a = *reinterpret_cast<const uint4*>(p.x + s), b = *reinterpret_cast<const uint4*>(p.x + s + xplane);
  uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t k = ((j < 4 ? id.x : id.y) >> (8 * (j & 3))) & 0xffu;
    if ((int)k != me) {
      const uint32_t keep = (j & 1) ? 0x0000ffffu : 0xffff0000u;
      aw[j >> 1] &= keep;
      bw[j >> 1] &= keep;
    }
  }
  *reinterpret_cast<uint4*>(p.y + o) = make_uint4(aw[0], aw[1], aw[2], aw[3]);
  *reinterpret_cast<uint4*>(p.y + o + yplane) = make_uint4(bw[0], bw[1], bw[2], bw[3]);
#else
  for (int j = 0; j < 8; ++j) {
    const bool hit = p.idx[s + j] == me;
    p.y[o + j] = hit ? p.x[s + j] : (uint16_t)0;
    p.y[o + j + yplane] = hit ? p.x[s + j + xplane] : (uint16_t)0;
  }
#endif
}

struct DimFixP {
  const void* tris;   // [F,1,H,W] fp32 or uint8
  int is_u8;
  ll pixels;          // F*H*W
  uint16_t* x8;       // dense [F,H,W,8]
};
// work item = one pixel: channel 3 := tri/255 (models/model.py:368: tri.float() * IMG_SCALE), channels 4..7 := 0
TCV_HD void dim_fix_inputs_body(ll i, const DimFixP& p) {
  const float tv = p.is_u8 ? (float)((const uint8_t*)p.tris)[i] : ((const float*)p.tris)[i];
  const ll plane = p.pixels * 8;
  uint16_t* d = p.x8 + i * 8;
  st1(d + 3, plane, tv * (1.0f / 255));
  for (int k = 4; k < 8; ++k) {
    d[k] = 0;
    d[k + plane] = 0;
  }
}

// ------------------------------------------------------------------------------------------ IndexNet (models/Index)
struct DwConvP {
  const uint16_t* x;   // dense [n, h, w, c]
  int n, h, w, c, dil;
  const float* wt;     // [9][c] (tap-major, tap = ky*3+kx)
  const float* scale;  // [c] BatchNorm scale / shift (eval fold)
  const float* shift;
  const float* border; // [c] value of an out-of-image tap (NULL: zero padding).  MobileNetV2 blocks of the reference pad
                       // the block INPUT and run the 1x1 expansion + BN + ReLU6 over the padded tensor (net.py:79-83), so
                       // the depthwise conv sees relu6(BN shift) -- not zero -- in its one-pixel border.
  int act;
  uint16_t* y;         // dense [n, h, w, c]
};
// work item = 8 channels of one output pixel: depthwise 3x3 (dilation dil, padding dil) + affine + activation
TCV_HD void dwconv3x3_body(ll i, const DwConvP& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int x = (int)(t % p.w);
  t /= p.w;
  const int y = (int)(t % p.h);
  const int img = (int)(t / p.h);
  const ll plane = (ll)p.n * p.h * p.w * p.c;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 9; ++k) {
    const int yy = y + (k / 3 - 1) * p.dil, xx = x + (k % 3 - 1) * p.dil;
    float f[8];
    if (yy >= 0 && yy < p.h && xx >= 0 && xx < p.w) {
      ld8(p.x + (((ll)img * p.h + yy) * p.w + xx) * p.c + ch, plane, f);
    } else {
      if (p.border) ldf8(p.border + ch, f);
      else
        for (int j = 0; j < 8; ++j) f[j] = 0.f;
    }
    float wv[8];
#ifdef __CUDA_ARCH__
    {   // (ch and c are multiples of 8: two aligned 16-byte loads instead of eight scalar ones)
      const float4 w0 = __ldg(reinterpret_cast<const float4*>(p.wt + (ll)k * p.c + ch));
      const float4 w1 = __ldg(reinterpret_cast<const float4*>(p.wt + (ll)k * p.c + ch + 4));
      wv[0] = w0.x; wv[1] = w0.y; wv[2] = w0.z; wv[3] = w0.w; wv[4] = w1.x; wv[5] = w1.y; wv[6] = w1.z; wv[7] = w1.w;
    }
#else
    for (int j = 0; j < 8; ++j) wv[j] = p.wt[(ll)k * p.c + ch + j];
#endif
    for (int j = 0; j < 8; ++j) acc[j] += f[j] * wv[j];
  }
  float sc[8], sh[8];
  ldf8(p.scale + ch, sc);
  ldf8(p.shift + ch, sh);
  for (int j = 0; j < 8; ++j) acc[j] = act_fn(acc[j] * sc[j] + sh[j], p.act);
  st8(p.y + (((ll)img * p.h + y) * p.w + x) * p.c + ch, plane, acc);
}

struct IndexFinishP {
  const uint16_t* b[4];  // the four branch outputs, dense [n, h2, w2, c] each
  int n, h2, w2, c;
  uint16_t* idx_en;      // dense [n, 2*h2, 2*w2, c]: softmax over the branches of sigmoid(branch)   (hlindex.py:155-166)
  uint16_t* idx_de;      //                           sigmoid(branch); branch k lands on sub-pixel (k / 2, k % 2)
};
// work item = 8 channels of one LOW-resolution pixel
TCV_HD void index_finish_body(ll i, const IndexFinishP& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int x = (int)(t % p.w2);
  t /= p.w2;
  const int y = (int)(t % p.h2);
  const int img = (int)(t / p.h2);
  const ll iplane = (ll)p.n * p.h2 * p.w2 * p.c, oplane = iplane * 4;
  const ll src = (((ll)img * p.h2 + y) * p.w2 + x) * p.c + ch;
  float s[4][8], e[4][8], sum[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) {
    ld8(p.b[k] + src, iplane, s[k]);
    for (int j = 0; j < 8; ++j) {
      s[k][j] = 1.0f / (1.0f + expf(-s[k][j]));
      e[k][j] = expf(s[k][j]);
      sum[j] += e[k][j];
    }
  }
  for (int k = 0; k < 4; ++k) {
    const ll dst = (((ll)img * 2 * p.h2 + 2 * y + (k >> 1)) * (2 * p.w2) + 2 * x + (k & 1)) * p.c + ch;
    float z[8];
    for (int j = 0; j < 8; ++j) z[j] = e[k][j] / sum[j];
    st8(p.idx_en + dst, oplane, z);
    st8(p.idx_de + dst, oplane, s[k]);
  }
}

struct IndexPoolP {
  const uint16_t* x;       // dense [n, h, w, c]
  const uint16_t* idx_en;  // dense [n, h, w, c]
  int n, h, w, c;
  uint16_t* masked;        // dense [n, h, w, c] = idx_en * x                         (net.py:193,201,...)
  uint16_t* pooled;        // dense [n, h/2, w/2, c] = 4 * avg_pool2(idx_en * x) = sum over the 2x2 window
};
// work item = 8 channels of one POOLED pixel
TCV_HD void index_pool_body(ll i, const IndexPoolP& p) {
  const int cv = p.c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int oh = p.h / 2, ow = p.w / 2;
  const int x = (int)(t % ow);
  t /= ow;
  const int y = (int)(t % oh);
  const int img = (int)(t / oh);
  const ll plane = (ll)p.n * p.h * p.w * p.c, pplane = (ll)p.n * oh * ow * p.c;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < 4; ++k) {
    const ll o = (((ll)img * p.h + 2 * y + (k >> 1)) * p.w + 2 * x + (k & 1)) * p.c + ch;
    float a[8], b[8];
    ld8(p.x + o, plane, a);
    ld8(p.idx_en + o, plane, b);
    for (int j = 0; j < 8; ++j) {
      a[j] *= b[j];
      acc[j] += a[j];
    }
    st8(p.masked + o, plane, a);
  }
  // the reference pools the STORED product (4 * avg_pool2d of l = idx_en * l); summing the fp32 products differs from that
  // only by the rounding of the stored tensor
  st8(p.pooled + (((ll)img * oh + y) * ow + x) * p.c + ch, pplane, acc);
}

struct IndexUpcatP {
  const uint16_t* dec;   // dense [n, h >> up, w >> up, dec_c]: decoder feature, first dec_real channels are used
  const uint16_t* idx;   // dense [n, h, w, idx_c] decoder indices (NULL: no index guidance)
  const uint16_t* low;   // dense [n, h, w, low_c]: encoder feature, first low_real channels are used
  int n, h, w, up;       // up = 1: nearest x2 upsampling of dec (only together with idx), 0: same resolution
  int dec_c, dec_real, idx_c, low_c, low_real, cat_c;
  uint16_t* cat;         // dense [n, h, w, cat_c] = [ idx * up(dec) | low | 0 ]     (hldecoder.py:121-127)
  ll idx_plane, low_plane;   // elements between the hi and lo plane of idx / low (they may be image slices of a larger tensor)
};
// work item = 8 channels of one output pixel
TCV_HD void index_upcat_body(ll i, const IndexUpcatP& p) {
  const int cv = p.cat_c / 8;
  const int ch = (int)(i % cv) * 8;
  ll t = i / cv;
  const int x = (int)(t % p.w);
  t /= p.w;
  const int y = (int)(t % p.h);
  const int img = (int)(t / p.h);
  const ll px = ((ll)img * p.h + y) * p.w + x;
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (ch < p.dec_real) {
    const int dh = p.h >> p.up, dw = p.w >> p.up;
    ld8(p.dec + (((ll)img * dh + (y >> p.up)) * dw + (x >> p.up)) * p.dec_c + ch, (ll)p.n * dh * dw * p.dec_c, f);
    if (p.idx) {
      float g[8];
      ld8(p.idx + px * p.idx_c + ch, p.idx_plane, g);
      for (int j = 0; j < 8; ++j) f[j] *= g[j];
    }
  } else if (ch < p.dec_real + p.low_real) {
    ld8(p.low + px * p.low_c + (ch - p.dec_real), p.low_plane, f);
  }
  st8(p.cat + px * p.cat_c + ch, (ll)p.n * p.h * p.w * p.cat_c, f);
}

// ------------------------------------------------------------------------------------------ calc_metric.py
struct MetricP {
  const uint8_t* a;    // predicted alpha, uint8 [h, w]            (calc_metric.py:49-62: PNGs, value / 255.0 -> float32)
  const uint8_t* g;    // ground-truth alpha
  const uint8_t* tri;  // trimap; the metrics run over 0 < tri < 255
  const uint8_t* ha;   // the NEXT frame's prediction / ground truth (NULL: single-frame metrics only)
  const uint8_t* hg;
  const float* flow;   // [h, w, 2] optical flow current -> next in pixels, NaN = invalid (NULL: no warped metric)
  int h, w;
  const float* lut;    // lut[v] = metric_u8(v), 256 entries (shared memory on the device): no fp64 divide per load
};
TCV_HD float metric_u8(uint8_t v) { return (float)((double)v / 255.0); }
// bilinear sample with zero padding at pixel coordinates (F.grid_sample(align_corners=True) through utils.grid_sampler,
// utils/utils.py:76-90: the coordinates make the round trip through the normalised grid in fp32)
TCV_HD float metric_sample(const uint8_t* img, const float* lut, int h, int w, float px, float py) {
  const float gx = 2.0f * px / (float)(w - 1) - 1.0f, gy = 2.0f * py / (float)(h - 1) - 1.0f;
  const float ix = (gx + 1.0f) * 0.5f * (float)(w - 1), iy = (gy + 1.0f) * 0.5f * (float)(h - 1);
  const float x0f = floorf(ix), y0f = floorf(iy);
  const int x0 = (int)x0f, y0 = (int)y0f;
  const float wx[2] = {x0f + 1.0f - ix, ix - x0f}, wy[2] = {y0f + 1.0f - iy, iy - y0f};
  float acc = 0.f;
  for (int dy = 0; dy < 2; ++dy)
    for (int dx = 0; dx < 2; ++dx) {
      const int xx = x0 + dx, yy = y0 + dy;
      if (xx < 0 || xx >= w || yy < 0 || yy >= h) continue;
      acc += lut[img[(ll)yy * w + xx]] * (wx[dx] * wy[dy]);
    }
  return acc;
}
// contributions of pixel i to out[7] = (pixel count, sum |a-g|, sum (a-g)^2, sum ((a-ha)-(g-hg))^2,
//                                       sum |(a-g)-(pa-pg)|, sum |(a-g)^2-(pa-pg)^2|, flow pixel count)
TCV_HD void metric_body(ll i, const MetricP& p, double* out) {
  const uint8_t t = p.tri[i];
  if (!(t > 0 && t < 255)) return;
  const float a = p.lut[p.a[i]], g = p.lut[p.g[i]];
  const float d = a - g;
  out[0] += 1.0;
  out[1] += (double)fabsf(d);
  out[2] += (double)(d * d);
  if (!p.ha) return;
  const float dd = (a - p.lut[p.ha[i]]) - (g - p.lut[p.hg[i]]);
  out[3] += (double)(dd * dd);
  if (!p.flow) return;
  const float fx = p.flow[2 * i];
  float fy = p.flow[2 * i + 1];
  if (fx != fx) return;                                   // NaN: no flow for this pixel (calc_metric.py:64-70); the
  if (fy != fy) fy = 0.f;                                 // validity mask is the x channel's (utils/utils.py:109-113)
  const int y = (int)((unsigned)i / (unsigned)p.w), x = (int)i - y * p.w;   // h * w < 2^31 (checked by the entry point)
  const float e = metric_sample(p.ha, p.lut, p.h, p.w, (float)x + fx, (float)y + fy) -
                  metric_sample(p.hg, p.lut, p.h, p.w, (float)x + fx, (float)y + fy);
  out[4] += (double)fabsf(d - e);
  out[5] += (double)fabsf(d * d - e * e);
  out[6] += 1.0;
}

}  // namespace tcv_fba
