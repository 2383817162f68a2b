// TEST DOUBLE of the tcvom_b200 C ABI for the CPU test-suite (tests/test_host_emul_fba.py).  NOT a product path and
// not a fallback: nothing under tcvom_b200/ can load it; the tests substitute it for libtcvom_b200.so by
// monkey-patching tcvom_b200._cabi._lib so that the HOST program of FbaVmnEngine (which kernels run, in which
// order, on which buffers, with which descriptors) can be executed on host memory and compared with the oracle
// in the GPU-less build container.
//
// * the FBA element-wise kernels execute the very same per-work-item bodies as the CUDA kernels
//   (tcvom_b200/csrc/fba_body.h), looped over the work items;
// * tcv_conv2d, tcv_tam_attend, tcv_preprocess_eval, tcv_ws_pack, tcv_gn_stats, tcv_adaptive_avgpool are independent
//   naive restatements of the semantics documented in include/tcvom_b200.h.
// Only the entry points the FBA eval program uses are provided.
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/tcvom_b200.h"
#include "../../tcvom_b200/csrc/fba_body.h"

using namespace tcv_fba;

static long long g_launches = 0;
static const char* g_err = "";

#define REQ(c, msg) do { if (!(c)) { g_err = msg; return TCV_ERR_INVALID; } } while (0)

template <typename P, void (*BODY)(ll, const P&)>
static int run_body(const P& p, ll total) {
#pragma omp parallel for schedule(static)
  for (ll i = 0; i < total; ++i) BODY(i, p);
  ++g_launches;
  return 0;
}

static std::vector<float> to_float(const uint16_t* x, ll plane, ll img_stride, int n, ll img_elems) {
  std::vector<float> f((size_t)n * img_elems);
#pragma omp parallel for
  for (ll i = 0; i < (ll)n * img_elems; ++i) {
    const ll img = i / img_elems, e = i % img_elems;
    f[i] = ld1(x + img * img_stride + e, plane);
  }
  return f;
}

static int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * n - 2 - i;
  return i;
}

template <typename T>
static int preprocess_eval_t(const T* imgs, const T* tris, int frames, int h, int w, int dilate, void* x8v,
                             float* trimask) {
  const ll hw = (ll)h * w;
  std::vector<uint8_t> m((size_t)frames * hw);
  for (ll i = 0; i < frames * hw; ++i) {
    const float s = (float)tris[i] * (1.0f / 255);
    m[i] = (s > 0.f && s < 1.f) ? 1 : 0;
  }
  uint16_t* x8 = (uint16_t*)x8v;
  const ll plane = (ll)frames * hw * 8;
  const float mean[3] = {0.485f, 0.456f, 0.406f}, stdv[3] = {0.229f, 0.224f, 0.225f};
  for (int f = 0; f < frames; ++f)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        uint8_t v = 0;
        if (dilate > 0) {
          for (int yy = y - dilate; yy <= y + dilate && !v; ++yy)
            for (int xx = x - dilate; xx <= x + dilate && !v; ++xx)
              if (yy >= 0 && yy < h && xx >= 0 && xx < w && m[f * hw + (ll)yy * w + xx]) v = 1;
        } else {
          v = m[f * hw + (ll)y * w + x];
        }
        const ll i = f * hw + (ll)y * w + x;
        trimask[i] = v ? 1.f : 0.f;
        // EvalModel.preprocess, TRIMAP_CHANNEL == 3 (models/model.py:366-379): normalised RGB + one-hot {bg, unknown, fg}
        float o[8] = {0, 0, 0, 0, 0, 0, 0, 0};
        for (int c = 0; c < 3; ++c) {
          const float sc = (float)imgs[((ll)f * 3 + (2 - c)) * hw + (ll)y * w + x] * (1.0f / 255);
          o[c] = (sc - mean[c]) / stdv[c];
        }
        const float st = (float)tris[i] * (1.0f / 255);
        const int cls = v ? 1 : (int)(2.0f * st);
        o[3 + cls] = 1.f;
        for (int c = 0; c < 8; ++c) st1(x8 + i * 8 + c, plane, o[c]);
      }
  ++g_launches;
  return 0;
}

template <typename T>
static int postprocess_eval_t(const float* pred, const T* tris, const float* trimask, int batch, int frames, int h, int w,
                              float* alphas) {
  const ll hw = (ll)h * w;
  for (ll i = 0; i < (ll)batch * frames * hw; ++i) {
    const ll f = i / hw, p = i % hw;
    const int b = (int)(f / frames), s = (int)(f % frames);
    float a = 0.f;
    if (s > 0 && s < frames - 1) {
      const float pr = pred[((ll)b * (frames - 2) + (s - 1)) * hw + p];
      a = trimask[i] != 0.f ? pr : (float)tris[i] * (1.0f / 255);
    }
    alphas[i] = a;
  }
  ++g_launches;
  return 0;
}

extern "C" {

int tcv_version(void) { return 100; }
const char* tcv_last_error(void) { return g_err; }
long long tcv_launch_count(void) { return g_launches; }
// roughly the shapes the persistent tcgen05 kernels take on the GPU (so that host logic keyed on the path -- statistics
// fused into the conv epilogue -- is exercised here as well)
int tcv_conv2d_path(const tcv_conv_desc* d) {
  return (d->w_tc && d->cout % 32 == 0 && d->cin % 32 == 0 && d->stride == 1 && d->pad_mode == TCV_PAD_ZERO && d->y && !d->y_f32) ? (d->cout % 64 == 0 ? 4 : 2) : 0;
}
int tcv_pack_weight_tc(const float*, int, int, int, void*, tcv_stream_t) { return 0; }

int tcv_conv2d(const tcv_conv_desc* dp, tcv_stream_t) {
  tcv_conv_desc d = *dp;
  REQ(d.x && d.w && (d.y || d.y_f32), "conv2d: null");
  REQ(d.cin % 8 == 0 && (d.cin <= 32 || d.cin % 32 == 0), "conv2d: cin");
  REQ(d.ntaps >= 1 && d.ntaps <= TCV_MAX_TAPS, "conv2d: ntaps");
  REQ(d.stride == 1 || d.stride == 2, "conv2d: stride");
  REQ((d.gh - 1) * d.oy_mul + d.oy_off < d.oh && (d.gw - 1) * d.ox_mul + d.ox_off < d.ow, "conv2d: grid");
  const ll img_elems = (ll)d.ih * d.iw * d.cin;
  if (d.x_plane == 0) d.x_plane = (ll)d.n * img_elems;
  if (d.x_img_stride == 0) d.x_img_stride = img_elems;
  if (d.res1 && d.res1_plane == 0) d.res1_plane = (ll)d.n * (d.oh >> d.res1_shift) * (d.ow >> d.res1_shift) * d.cout;
  if (d.res2 && d.res2_plane == 0) d.res2_plane = (ll)d.n * d.oh * d.ow * d.cout;
  const std::vector<float> X = to_float((const uint16_t*)d.x, d.x_plane, d.x_img_stride, d.n, img_elems);
  const ll oplane = (ll)d.n * d.oh * d.ow * d.cout;
#pragma omp parallel for collapse(2) schedule(static)
  for (int n = 0; n < d.n; ++n)
    for (int gy = 0; gy < d.gh; ++gy) {
      std::vector<float> acc(d.cout);
      for (int gx = 0; gx < d.gw; ++gx) {
        for (int j = 0; j < d.cout; ++j) acc[j] = 0.f;
        for (int t = 0; t < d.ntaps; ++t) {
          int iy = gy * d.stride + d.dy[t], ix = gx * d.stride + d.dx[t];
          if (d.pad_mode == TCV_PAD_REFLECT) { iy = reflect_idx(iy, d.ih); ix = reflect_idx(ix, d.iw); }
          else if (iy < 0 || iy >= d.ih || ix < 0 || ix >= d.iw) continue;
          const float* xp = X.data() + (ll)n * img_elems + ((ll)iy * d.iw + ix) * d.cin;
          const float* wp = d.w + (ll)d.wtap[t] * d.cin * d.cout;
          for (int ci = 0; ci < d.cin; ++ci) {
            const float xv = xp[ci];
            const float* wr = wp + (ll)ci * d.cout;
            for (int j = 0; j < d.cout; ++j) acc[j] += xv * wr[j];
          }
        }
        const int oy = gy * d.oy_mul + d.oy_off, ox = gx * d.ox_mul + d.ox_off;
        const ll obase = (((ll)n * d.oh + oy) * d.ow + ox) * d.cout;
        for (int j = 0; j < d.cout; ++j) {
          float a = acc[j];
          if (d.s1) a *= d.s1[j];
          if (d.b1) a += d.b1[j];
          if (d.res1) {
            const int rh = d.oh >> d.res1_shift, rw = d.ow >> d.res1_shift;
            a += ld1((const uint16_t*)d.res1 + (((ll)n * rh + (oy >> d.res1_shift)) * rw + (ox >> d.res1_shift)) * d.cout + j,
                     d.res1_plane);
          }
          a = act_fn(a, d.act);
          if (d.s2) a = a * d.s2[j] + d.b2[j];
          if (d.res2) a += ld1((const uint16_t*)d.res2 + obase + j, d.res2_plane);
          if (d.y) st1((uint16_t*)d.y + obase + j, oplane, a);
          if (d.y_f32) d.y_f32[obase + j] = a;
          if (d.stats) {
            double* acc2 = d.stats + ((ll)(n % (d.stats_groups > 0 ? d.stats_groups : 1)) * d.cout + j) * 2;
#pragma omp atomic
            acc2[0] += (double)a;
#pragma omp atomic
            acc2[1] += (double)a * (double)a;
          }
        }
      }
    }
  ++g_launches;
  return 0;
}

int tcv_zero_bytes(void* p, long long bytes, tcv_stream_t) {
  memset(p, 0, (size_t)bytes);
  ++g_launches;
  return 0;
}

int tcv_preprocess_eval(const float* imgs, const float* tris, int frames, int h, int w, int dilate, void* x8,
                        float* trimask, uint8_t*, tcv_stream_t) {
  return preprocess_eval_t<float>(imgs, tris, frames, h, w, dilate, x8, trimask);
}
int tcv_preprocess_eval_u8(const uint8_t* imgs, const uint8_t* tris, int frames, int h, int w, int dilate, void* x8,
                           float* trimask, uint8_t*, tcv_stream_t) {
  return preprocess_eval_t<uint8_t>(imgs, tris, frames, h, w, dilate, x8, trimask);
}
int tcv_postprocess_eval(const float* pred, const float* tris, const float* trimask, int batch, int frames, int h, int w,
                         float* alphas, tcv_stream_t) {
  return postprocess_eval_t<float>(pred, tris, trimask, batch, frames, h, w, alphas);
}
int tcv_postprocess_eval_u8(const float* pred, const uint8_t* tris, const float* trimask, int batch, int frames, int h,
                            int w, float* alphas, tcv_stream_t) {
  return postprocess_eval_t<uint8_t>(pred, tris, trimask, batch, frames, h, w, alphas);
}

// ------------------------------------------------------------------------------------------ vmn_gca entry points
// (naive restatements of the semantics documented in include/tcvom_b200.h; fp32 operand formats only: the engine is
// driven with use_tc_attn = False in the host-logic tests)
int tcv_pack_weight_fold(const float*, int, void*, tcv_stream_t) { return 0; }

int tcv_sn_fold_pack(const float* w_bar, const float* u, const float* v, int cout, int cin, int kh, int kw,
                     int transposed, int cin_pad, float* packed, float* sigma_out, tcv_stream_t) {
  const int taps = kh * kw;
  double sigma = 1.0;
  if (u) {   // sigma = u^T W v, W = w_bar viewed [shape[0], rest]  (GCA/ops.py:38-45)
    const int rows = transposed ? cin : cout, cols = (transposed ? cout : cin) * taps;
    sigma = 0.0;
    for (int r = 0; r < rows; ++r) {
      double sr = 0.0;
      for (int c = 0; c < cols; ++c) sr += (double)w_bar[(ll)r * cols + c] * v[c];
      sigma += sr * u[r];
    }
  }
  if (sigma_out) *sigma_out = (float)sigma;
  for (int t = 0; t < taps; ++t)
    for (int ci = 0; ci < cin_pad; ++ci)
      for (int co = 0; co < cout; ++co) {
        float val = 0.f;
        if (ci < cin) {
          const ll src = transposed ? (((ll)ci * cout + co) * taps + t) : (((ll)co * cin + ci) * taps + t);
          val = u ? w_bar[src] / (float)sigma : w_bar[src];
        }
        packed[((ll)t * cin_pad + ci) * cout + co] = val;
      }
  ++g_launches;
  return 0;
}

int tcv_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int c,
                float* scale, float* shift, tcv_stream_t) {
  for (int i = 0; i < c; ++i) {
    const float s = gamma[i] / sqrtf(var[i] + eps);
    scale[i] = s;
    shift[i] = beta[i] - mean[i] * s;
  }
  ++g_launches;
  return 0;
}

int tcv_avgpool2(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t) {
  const int oh = h / 2, ow = w / 2;
  const ll ip = (ll)n * h * w * c, op = (ll)n * oh * ow * c;
  for (int img = 0; img < n; ++img)
    for (int oy = 0; oy < oh; ++oy)
      for (int ox = 0; ox < ow; ++ox)
        for (int ch = 0; ch < c; ++ch) {
          float a = 0.f;
          for (int dy = 0; dy < 2; ++dy)
            for (int dx = 0; dx < 2; ++dx)
              a += ld1((const uint16_t*)x + (((ll)img * h + 2 * oy + dy) * w + 2 * ox + dx) * c + ch, ip);
          st1((uint16_t*)y + (((ll)img * oh + oy) * ow + ox) * c + ch, op, a * 0.25f);
        }
  ++g_launches;
  return 0;
}

int tcv_pad_reflect1(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t) {
  const int oh = h + 2, ow = w + 2;
  const ll ip = (ll)n * h * w * c, op = (ll)n * oh * ow * c;
  for (int img = 0; img < n; ++img)
    for (int oy = 0; oy < oh; ++oy)
      for (int ox = 0; ox < ow; ++ox) {
        const int iy = reflect_idx(oy - 1, h), ix = reflect_idx(ox - 1, w);
        const uint16_t* src = (const uint16_t*)x + (((ll)img * h + iy) * w + ix) * c;
        uint16_t* dst = (uint16_t*)y + (((ll)img * oh + oy) * ow + ox) * c;
        for (int ch = 0; ch < c; ++ch) { dst[ch] = src[ch]; dst[ch + op] = src[ch + ip]; }
      }
  ++g_launches;
  return 0;
}

int tcv_unknown_os8(const void* x8, int n, int h, int w, float* unknown, tcv_stream_t) {
  const int oh = h / 8, ow = w / 8;
  for (int img = 0; img < n; ++img)
    for (int oy = 0; oy < oh; ++oy)
      for (int ox = 0; ox < ow; ++ox)
        unknown[((ll)img * oh + oy) * ow + ox] = bf_to_f(((const uint16_t*)x8)[(((ll)img * h + oy * 8) * w + ox * 8) * 8 + 4]);
  ++g_launches;
  return 0;
}

// GCA/ops.py:106-229 (see the header): g [n,h/2,w/2,64] split-bf16 (guidance_conv at stride 2), unknown fp32 [n,h,w]
int tcv_gca_prep(const void* g, const float* unknown, int n, int h, int w, void* Qv, void* Knv, float* mm, float* scales,
                 int bf16_split, tcv_stream_t) {
  REQ(bf16_split == 0 || bf16_split == 2 || bf16_split == 3, "gca_prep: planes must be 0, 2 or 3");
  const int hh = h / 2, ww = w / 2, P = hh * ww, GC = 64, QD = 576;
  const ll gplane = (ll)n * P * GC;
  std::vector<float> qbuf, kbuf;
  if (bf16_split) { qbuf.resize((size_t)n * P * QD); kbuf.resize((size_t)n * P * QD); }
  float* Q = bf16_split ? qbuf.data() : (float*)Qv;
  float* Kn = bf16_split ? kbuf.data() : (float*)Knv;
  for (int img = 0; img < n; ++img) {
    const float* u = unknown + (ll)img * h * w;
    double s = 0.0;
    for (int y = 0; y < hh; ++y)
      for (int x = 0; x < ww; ++x) s += u[(2 * y) * w + 2 * x];
    const float um = (float)(s / (hh * ww)), km = 1.0f - um;
    scales[2 * img] = fminf(fmaxf(sqrtf(um / km), 0.1f), 10.f);
    scales[2 * img + 1] = fminf(fmaxf(sqrtf(km / um), 0.1f), 10.f);
    for (int p = 0; p < P; ++p) {
      const int py = p / ww, px = p % ww;
      float* q = Q + ((ll)img * P + p) * QD;
      float ss = 0.f, usum = 0.f;
      for (int t = 0; t < 9; ++t) {
        const int yy = reflect_idx(py + t / 3 - 1, hh), xx = reflect_idx(px + t % 3 - 1, ww);
        for (int c = 0; c < GC; ++c) {
          const float val = ld1((const uint16_t*)g + ((ll)img * P + (ll)yy * ww + xx) * GC + c, gplane);
          q[t * GC + c] = val;
          ss += val * val;
        }
        usum += u[(2 * yy) * w + 2 * xx];
      }
      const float m = usum > 0.f ? 1.f : 0.f;
      const float inv = (m > 0.f ? scales[2 * img] : scales[2 * img + 1]) / fmaxf(sqrtf(ss), 1e-4f);
      float* k = Kn + ((ll)img * P + p) * QD;
      for (int i = 0; i < QD; ++i) k[i] = q[i] * inv;
      mm[(ll)img * P + p] = m;
    }
  }
  if (bf16_split) {   // bf16 planes [split][n][P][576]: hi, (mid,) lo
    const ll plane = (ll)n * P * QD;
    for (ll i = 0; i < plane; ++i) {
      float a = Q[i], b = Kn[i];
      for (int pl = 0; pl < bf16_split; ++pl) {
        const uint16_t ha = f_to_bf(a), hb = f_to_bf(b);
        ((uint16_t*)Qv)[pl * plane + i] = ha;
        ((uint16_t*)Knv)[pl * plane + i] = hb;
        a -= bf_to_f(ha);
        b -= bf_to_f(hb);
      }
    }
  }
  ++g_launches;
  return 0;
}

int tcv_gca_values(const void* feat, int n, int h, int w, void* Vtv, int mode, tcv_stream_t) {
  REQ(mode >= 0 && mode <= 2, "gca_values: the test double provides fp32 / bf16 / split-bf16 operands (not fp16)");
  const int hh = h / 2, ww = w / 2, P = hh * ww, P_pad = (P + 63) / 64 * 64, FC = 128, VD = 2048;
  const ll fplane = (ll)n * h * w * FC;
  std::vector<float> vbuf;
  if (mode) vbuf.resize((size_t)n * VD * P_pad);
  float* Vt = mode ? vbuf.data() : (float*)Vtv;
  for (int img = 0; img < n; ++img)
    for (int t = 0; t < 16; ++t)
      for (int c = 0; c < FC; ++c)
        for (int p = 0; p < P_pad; ++p) {
          float val = 0.f;
          if (p < P) {
            const int py = p / ww, px = p % ww;
            const int yy = reflect_idx(2 * py + t / 4 - 1, h), xx = reflect_idx(2 * px + t % 4 - 1, w);
            val = ld1((const uint16_t*)feat + (((ll)img * h + yy) * w + xx) * FC + c, fplane);
          }
          Vt[((ll)img * VD + t * FC + c) * P_pad + p] = val;
        }
  if (mode) {
    const ll total = (ll)n * VD * P_pad;
    for (ll i = 0; i < total; ++i) {
      if (mode == 1) ((uint16_t*)Vtv)[i] = f_to_bf(Vt[i]);
      else st1((uint16_t*)Vtv + i, total, Vt[i]);
    }
  }
  ++g_launches;
  return 0;
}

int tcv_gemm_tn_f32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb, int ldc,
                    long long strideA, long long strideB, long long strideC, int batch, tcv_stream_t) {
#pragma omp parallel for collapse(2) schedule(static)
  for (int b = 0; b < batch; ++b)
    for (int m = 0; m < M; ++m)
      for (int nn = 0; nn < N; ++nn) {
        const float* a = A + b * strideA + (ll)m * lda;
        const float* bb = B + b * strideB + (ll)nn * ldb;
        float acc = 0.f;
        for (int k = 0; k < K; ++k) acc += a[k] * bb[k];
        C[b * strideC + (ll)m * ldc + nn] = acc;
      }
  ++g_launches;
  return 0;
}

int tcv_gca_softmax(float* S, const float* mm, int n, int P, int P_pad, void* P_out, int mode, tcv_stream_t) {
  REQ(mode >= 0 && mode <= 2 && (mode == 0 || P_out), "gca_softmax: fp32 / bf16 / split-bf16 outputs only");
  for (int img = 0; img < n; ++img)
    for (int q = 0; q < P; ++q) {
      float* row = S + ((ll)img * P + q) * P_pad;
      row[q] += -1e4f * mm[(ll)img * P + q];
      float mx = -INFINITY;
      for (int p = 0; p < P; ++p) mx = fmaxf(mx, row[p]);
      float sum = 0.f;
      for (int p = 0; p < P; ++p) { row[p] = expf(row[p] - mx); sum += row[p]; }
      for (int p = 0; p < P; ++p) row[p] /= sum;
      for (int p = P; p < P_pad; ++p) row[p] = 0.f;
      if (mode) {
        const ll plane = (ll)n * P * P_pad, o = ((ll)img * P + q) * P_pad;
        for (int p = 0; p < P_pad; ++p) {
          if (mode == 1) ((uint16_t*)P_out)[o + p] = f_to_bf(row[p]);
          else st1((uint16_t*)P_out + o + p, plane, row[p]);
        }
      }
    }
  ++g_launches;
  return 0;
}

// C[b] = A[b] * B[b]^T with bf16 operand planes (see tcv_gemm_tn_tc in the header): nsplit 1: Ahi.Bhi; 3: hi.hi + hi.lo +
// lo.hi; 6: three planes, hh + hm + mh + hl + lh + mm.  fp32 accumulation, fp32 output.
int tcv_gemm_tn_tc(const void* A, long long a_plane, const void* B, long long b_plane, void* C, int M, int N, int K,
                   long long ldc, long long c_batch_stride, int batch, int nsplit, int out_bf16, int in_fp16, tcv_stream_t) {
  REQ(!out_bf16 && !in_fp16, "gemm_tn_tc: the test double provides bf16 operands and fp32 output only");
  REQ(nsplit == 1 || nsplit == 3 || nsplit == 6, "gemm_tn_tc: nsplit");
  REQ(K % 64 == 0, "gemm_tn_tc: K % 64");
  const int np = nsplit == 1 ? 1 : (nsplit == 3 ? 2 : 3);
  // which (plane of A, plane of B) products are summed
  const int pairs[6][2] = {{0, 0}, {0, 1}, {1, 0}, {0, 2}, {2, 0}, {1, 1}};
  std::vector<float> af((size_t)np * M * K), bf((size_t)np * N * K);
  for (int b = 0; b < batch; ++b) {
    for (int pl = 0; pl < np; ++pl) {
      for (ll i = 0; i < (ll)M * K; ++i) af[(ll)pl * M * K + i] = bf_to_f(((const uint16_t*)A)[pl * a_plane + (ll)b * M * K + i]);
      for (ll i = 0; i < (ll)N * K; ++i) bf[(ll)pl * N * K + i] = bf_to_f(((const uint16_t*)B)[pl * b_plane + (ll)b * N * K + i]);
    }
#pragma omp parallel for schedule(static)
    for (int m = 0; m < M; ++m)
      for (int nn = 0; nn < N; ++nn) {
        float acc = 0.f;
        for (int t = 0; t < nsplit; ++t) {
          const float* a = af.data() + (ll)pairs[t][0] * M * K + (ll)m * K;
          const float* bb = bf.data() + (ll)pairs[t][1] * N * K + (ll)nn * K;
          float s2 = 0.f;
          for (int k = 0; k < K; ++k) s2 += a[k] * bb[k];
          acc += s2;
        }
        ((float*)C)[b * c_batch_stride + (ll)m * ldc + nn] = acc;
      }
  }
  ++g_launches;
  return 0;
}

int tcv_gca_fold(const float* O, int n, int h, int w, void* Y, tcv_stream_t) {
  const int hh = h / 2, ww = w / 2, P = hh * ww, FC = 128, VD = 2048;
  const ll yplane = (ll)n * h * w * FC;
  for (int img = 0; img < n; ++img)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x)
        for (int c = 0; c < FC; ++c) {
          float acc = 0.f;
          // output pixel (y, x) is covered by patch (qy, qx) at tap (ty, tx) iff y = 2*qy + ty - 1, x = 2*qx + tx - 1
          for (int ty = 0; ty < 4; ++ty) {
            if ((y + 1 - ty) % 2 != 0 || y + 1 - ty < 0) continue;
            const int qy = (y + 1 - ty) / 2;
            if (qy >= hh) continue;
            for (int tx = 0; tx < 4; ++tx) {
              if ((x + 1 - tx) % 2 != 0 || x + 1 - tx < 0) continue;
              const int qx = (x + 1 - tx) / 2;
              if (qx >= ww) continue;
              acc += O[((ll)img * P + qy * ww + qx) * VD + (ty * 4 + tx) * FC + c];
            }
          }
          st1((uint16_t*)Y + (((ll)img * h + y) * w + x) * FC + c, yplane, acc * 0.25f);
        }
  ++g_launches;
  return 0;
}

// ---- shift-sum form of the aggregation (see include/tcvom_b200.h): independent naive restatements
int tcv_gca_prep_grid(const void* g, const float* unknown, int n, int h, int w, void* Qv, void* Knv, float* mm, float* scales,
                      tcv_stream_t st) {
  const int hh = h / 2, ww = w / 2, P = hh * ww, Pk = (hh + 1) * (ww + 1), QD = 576;
  std::vector<uint16_t> kn((size_t)2 * n * P * QD);
  int rc = tcv_gca_prep(g, unknown, n, h, w, Qv, kn.data(), mm, scales, 2, st);
  if (rc) return rc;
  uint16_t* K = (uint16_t*)Knv;
  memset(K, 0, (size_t)2 * n * Pk * QD * sizeof(uint16_t));
  for (int pl = 0; pl < 2; ++pl)
    for (int img = 0; img < n; ++img)
      for (int p = 0; p < P; ++p)
        memcpy(K + (((ll)pl * n + img) * Pk + (p / ww) * (ww + 1) + p % ww) * QD, kn.data() + (((ll)pl * n + img) * P + p) * QD,
               QD * sizeof(uint16_t));
  return 0;
}

int tcv_gca_values_parity(const void* feat, int n, int h, int w, int ld, void* Ftv, tcv_stream_t) {
  const int hh = h / 2, ww = w / 2, Pk = (hh + 1) * (ww + 1), FC = 128;
  REQ(ld % 64 == 0 && ld >= Pk, "gca_values_parity: ld");
  const ll fplane = (ll)n * h * w * FC, total = (ll)n * 4 * FC * ld;
  for (int img = 0; img < n; ++img)
    for (int r = 0; r < 4; ++r)
      for (int c = 0; c < FC; ++c)
        for (int p = 0; p < ld; ++p) {
          float val = 0.f;
          if (p < Pk) {
            const int py = p / (ww + 1), px = p % (ww + 1);
            const int yy = reflect_idx(2 * py + (r >> 1) - 1, h), xx = reflect_idx(2 * px + (r & 1) - 1, w);
            val = ld1((const uint16_t*)feat + (((ll)img * h + yy) * w + xx) * FC + c, fplane);
          }
          st1((uint16_t*)Ftv + ((ll)img * 4 * FC + r * FC + c) * ld + p, total, val);
        }
  ++g_launches;
  return 0;
}

static bool key_valid(int j, int hh, int ww) { return j >= 0 && j < hh * (ww + 1) && j % (ww + 1) != ww; }

int tcv_gca_rowstats(float* S, const float* mm, int n, int h, int w, int ld, float* stats, int normalise, tcv_stream_t) {
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  for (int img = 0; img < n; ++img)
    for (int q = 0; q < P; ++q) {
      const float* row = S + ((ll)img * P + q) * ld;
      const int qj = (q / ww) * (ww + 1) + q % ww;
      const float diag = -1e4f * mm[(ll)img * P + q];
      float mx = -INFINITY;
      for (int j = 0; j < ld; ++j)
        if (key_valid(j, hh, ww)) mx = fmaxf(mx, row[j] + (j == qj ? diag : 0.f));
      float sum = 0.f;
      for (int j = 0; j < ld; ++j)
        if (key_valid(j, hh, ww)) sum += expf(row[j] + (j == qj ? diag : 0.f) - mx);
      stats[2 * ((ll)img * P + q)] = mx;
      stats[2 * ((ll)img * P + q) + 1] = 1.0f / sum;
      if (normalise) {
        float* out = S + ((ll)img * P + q) * ld;
        for (int j = 0; j < ld; ++j)
          out[j] = key_valid(j, hh, ww) ? expf(out[j] + (j == qj ? diag : 0.f) - mx) * (1.0f / sum) : 0.f;
      }
    }
  ++g_launches;
  return 0;
}

int tcv_gca_shift_add(const float* A, int n, int h, int w, int ld, void* A2v, tcv_stream_t) {
  const int hh = h / 2, ww = w / 2, P = hh * ww, ww1 = ww + 1, Pk = (hh + 1) * ww1;
  const ll plane = (ll)n * Pk * ld;
  for (int img = 0; img < n; ++img)
    for (int m = 0; m < Pk; ++m)
      for (int j = 0; j < ld; ++j) {
        float acc = 0.f;
        for (int a = 0; a < 4; ++a) {
          const int qy = m / ww1 - (a >> 1), qx = m % ww1 - (a & 1);
          const int js = j - ((a >> 1) * ww1 + (a & 1));
          if (qy < 0 || qy >= hh || qx < 0 || qx >= ww || js < 0) continue;
          acc += A[((ll)img * P + qy * ww + qx) * ld + js];
        }
        st1((uint16_t*)A2v + ((ll)img * Pk + m) * ld + j, plane, acc);
      }
  ++g_launches;
  return 0;
}

int tcv_gca_softmax_shift(const float* S, const float* stats, const float* mm, int n, int h, int w, int ld, void* A2v,
                          tcv_stream_t) {
  const int hh = h / 2, ww = w / 2, P = hh * ww, ww1 = ww + 1, Pk = (hh + 1) * ww1;
  const ll plane = (ll)n * Pk * ld;
  for (int img = 0; img < n; ++img)
    for (int m = 0; m < Pk; ++m)
      for (int j = 0; j < ld; ++j) {
        float acc[4] = {0, 0, 0, 0};
        for (int a = 0; a < 4; ++a) {
          const int qy = m / ww1 - (a >> 1), qx = m % ww1 - (a & 1);
          const int js = j - ((a >> 1) * ww1 + (a & 1));
          if (qy < 0 || qy >= hh || qx < 0 || qx >= ww || !key_valid(js, hh, ww)) continue;
          const ll q = (ll)img * P + qy * ww + qx;
          const float t = S[q * ld + js] + (js == qy * ww1 + qx ? -1e4f * mm[q] : 0.f);
          acc[a] = expf(t - stats[2 * q]) * stats[2 * q + 1];
        }
        st1((uint16_t*)A2v + ((ll)img * Pk + m) * ld + j, plane, (acc[0] + acc[1]) + (acc[2] + acc[3]));
      }
  ++g_launches;
  return 0;
}

int tcv_gca_unfold_parity(const float* O2, int n, int h, int w, void* Y, tcv_stream_t) {
  const int ww1 = w / 2 + 1, Pk = (h / 2 + 1) * ww1, FC = 128;
  const ll yplane = (ll)n * h * w * FC;
  for (int img = 0; img < n; ++img)
    for (int y = 0; y < h; ++y)
      for (int x = 0; x < w; ++x) {
        const int ry = (y + 1) % 2, rx = (x + 1) % 2, my = (y + 1 - ry) / 2, mx = (x + 1 - rx) / 2;
        for (int c = 0; c < FC; ++c)
          st1((uint16_t*)Y + (((ll)img * h + y) * w + x) * FC + c, yplane,
              O2[((ll)img * Pk + my * ww1 + mx) * 4 * FC + (ry * 2 + rx) * FC + c] * 0.25f);
      }
  ++g_launches;
  return 0;
}

int tcv_head_conv5_clamp01(const void* x, long long x_plane, int n, int h, int w, const float* wt, const float* bias,
                           float* pred, tcv_stream_t) {
  if (x_plane == 0) x_plane = (ll)n * h * w * 64;
  for (int img = 0; img < n; ++img)
    for (int y = 0; y < h; ++y)
      for (int xx = 0; xx < w; ++xx) {
        float acc = bias ? bias[0] : 0.f;
        for (int t = 0; t < 25; ++t) {
          const int yy = y + t / 5 - 2, xc = xx + t % 5 - 2;
          if (yy < 0 || yy >= h || xc < 0 || xc >= w) continue;
          for (int c = 0; c < 64; ++c)
            acc += ld1((const uint16_t*)x + (((ll)img * h + yy) * w + xc) * 64 + c, x_plane) * wt[t * 64 + c];
        }
        pred[((ll)img * h + y) * w + xx] = acc < 0.f ? 0.f : (acc > 1.f ? 1.f : acc);
      }
  ++g_launches;
  return 0;
}

int tcv_head_conv_tanh01(const void* x, long long x_plane, int n, int h, int w, const float* wt, const float* bias, float* pred,
                         tcv_stream_t) {
  if (x_plane == 0) x_plane = (ll)n * h * w * 32;
  for (int img = 0; img < n; ++img)
    for (int y = 0; y < h; ++y)
      for (int xx = 0; xx < w; ++xx) {
        float acc = bias ? bias[0] : 0.f;
        for (int t = 0; t < 9; ++t) {
          const int yy = y + t / 3 - 1, xc = xx + t % 3 - 1;
          if (yy < 0 || yy >= h || xc < 0 || xc >= w) continue;
          for (int c = 0; c < 32; ++c)
            acc += ld1((const uint16_t*)x + (((ll)img * h + yy) * w + xc) * 32 + c, x_plane) * wt[t * 32 + c];
        }
        pred[((ll)img * h + y) * w + xx] = (tanhf(acc) + 1.0f) * 0.5f;
      }
  ++g_launches;
  return 0;
}

int tcv_head_tanh01(const void* x, long long x_plane, long long pixels, int c, float* pred, tcv_stream_t) {
  if (x_plane == 0) x_plane = pixels * c;
  for (ll i = 0; i < pixels; ++i) pred[i] = (tanhf(ld1((const uint16_t*)x + i * c, x_plane)) + 1.0f) * 0.5f;
  ++g_launches;
  return 0;
}

int tcv_split_to_nchw(const void* x, int n, int c, int h, int w, int c_pad, long long x_plane, float* y, tcv_stream_t) {
  const ll hw = (ll)h * w;
  if (x_plane == 0) x_plane = (ll)n * hw * c_pad;
  for (ll i = 0; i < (ll)n * c * hw; ++i) {
    const ll p = i % hw, img = i / (hw * c);
    const int cc = (int)((i / hw) % c);
    y[i] = ld1((const uint16_t*)x + (img * hw + p) * c_pad + cc, x_plane);
  }
  ++g_launches;
  return 0;
}

int tcv_nchw_to_split(const float* x, int n, int c, int h, int w, int c_pad, void* y, long long y_plane, tcv_stream_t) {
  const ll hw = (ll)h * w;
  if (y_plane == 0) y_plane = (ll)n * hw * c_pad;
  for (ll i = 0; i < (ll)n * hw * c_pad; ++i) {
    const int cc = (int)(i % c_pad);
    const ll p = (i / c_pad) % hw, img = i / (c_pad * hw);
    st1((uint16_t*)y + i, y_plane, cc < c ? x[(img * c + cc) * hw + p] : 0.f);
  }
  ++g_launches;
  return 0;
}

// VMN_model.py:27-68 after the q/k/v convolutions (see tcv_tam_attend in the header)
int tcv_tam_attend(const void* q, const void* v, const void* kb, const void* kf, const float* mask,
                   long long mask_stride, int mh, int mw, int batch, int h, int w, int c, int window, void* out,
                   float* attb, float* attf, uint8_t* small_mask, tcv_stream_t) {
  const ll N = (ll)h * w, plane = (ll)batch * N * c;
  const int w2 = window * window, r = window / 2;
  const float inv = 1.0f / sqrtf((float)c);
#pragma omp parallel for schedule(static)
  for (ll gp = 0; gp < batch * N; ++gp) {
    const int b = (int)(gp / N), pix = (int)(gp % N), y = pix / w, x = pix % w;
    const int my = (int)(((ll)y * mh) / h), mx = (int)(((ll)x * mw) / w);
    const bool m = mask[(ll)b * mask_stride + (ll)my * mw + mx] != 0.f;
    small_mask[gp] = m ? 1 : 0;
    std::vector<float> o(c), qv(c), logit(w2);
    const ll base = gp * c;
    for (int ch = 0; ch < c; ++ch) o[ch] = ld1((const uint16_t*)v + base + ch, plane);
    for (int nb = 0; nb < 2; ++nb) {
      const uint16_t* k = (const uint16_t*)(nb == 0 ? kb : kf);
      float* att = nb == 0 ? attb : attf;
      if (!m) {
        for (int j = 0; j < w2; ++j) att[((ll)b * w2 + j) * N + pix] = 0.f;
        continue;
      }
      for (int ch = 0; ch < c; ++ch) qv[ch] = ld1((const uint16_t*)q + base + ch, plane);
      float mxv = -INFINITY;
      for (int j = 0; j < w2; ++j) {
        const int yy = y + j / window - r, xx = x + j % window - r;
        float dsum = 0.f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w)
          for (int ch = 0; ch < c; ++ch) dsum += qv[ch] * ld1(k + ((ll)b * N + (ll)yy * w + xx) * c + ch, plane);
        logit[j] = dsum * inv;
        att[((ll)b * w2 + j) * N + pix] = logit[j];
        mxv = fmaxf(mxv, logit[j]);
      }
      float den = 0.f;
      for (int j = 0; j < w2; ++j) den += expf(logit[j] - mxv);
      for (int j = 0; j < w2; ++j) {
        const int yy = y + j / window - r, xx = x + j % window - r;
        if (yy < 0 || yy >= h || xx < 0 || xx >= w) continue;
        const float a = expf(logit[j] - mxv) / den;
        for (int ch = 0; ch < c; ++ch) o[ch] += a * ld1(k + ((ll)b * N + (ll)yy * w + xx) * c + ch, plane);
      }
    }
    for (int ch = 0; ch < c; ++ch) st1((uint16_t*)out + base + ch, plane, o[ch]);
  }
  ++g_launches;
  return 0;
}

// ------------------------------------------------------------------------------------------ FBA entry points
int tcv_ws_pack(const float* w, int cout, int cin, int kh, int kw, int standardize, int cin_pad, int cout_pad,
                float* packed, tcv_stream_t) {
  const int taps = kh * kw, cnt = cin * taps;
  memset(packed, 0, sizeof(float) * (size_t)taps * cin_pad * cout_pad);
  for (int co = 0; co < cout; ++co) {
    double mean = 0.0, inv = 1.0;
    if (standardize) {
      for (int i = 0; i < cnt; ++i) mean += w[(ll)co * cnt + i];
      mean /= cnt;
      double ss = 0.0;
      for (int i = 0; i < cnt; ++i) { const double dd = w[(ll)co * cnt + i] - mean; ss += dd * dd; }
      inv = 1.0 / (sqrt(ss / (cnt - 1) + 1e-12) + 1e-5);
    }
    for (int ci = 0; ci < cin; ++ci)
      for (int t = 0; t < taps; ++t)
        packed[((ll)t * cin_pad + ci) * cout_pad + co] = (float)((w[((ll)co * cin + ci) * taps + t] - mean) * inv);
  }
  ++g_launches;
  return 0;
}

int tcv_gn_stats(const void* x, long long x_plane, int n, long long pixels, int c, double* sums, tcv_stream_t) {
  if (x_plane == 0) x_plane = (ll)n * pixels * c;
  for (ll i = 0; i < (ll)n * c * 2; ++i) sums[i] = 0.0;
  for (int img = 0; img < n; ++img)
    for (ll p = 0; p < pixels; ++p)
      for (int ch = 0; ch < c; ++ch) {
        const double v = ld1((const uint16_t*)x + ((ll)img * pixels + p) * c + ch, x_plane);
        sums[((ll)img * c + ch) * 2] += v;
        sums[((ll)img * c + ch) * 2 + 1] += v * v;
      }
  ++g_launches;
  return 0;
}

int tcv_gn_finalize(const double* sums, int n, long long pixels, int c, int groups, const float* gamma,
                    const float* beta, float eps, float* scale, float* shift, tcv_stream_t) {
  GnFinalizeP p{sums, n, c, groups, pixels, gamma, beta, eps, scale, shift, 1, 0};
  return run_body<GnFinalizeP, gn_finalize_body>(p, (ll)n * groups);
}

int tcv_gn_finalize_acc(double* sums, int copies, int clear, int n, long long pixels, int c, int groups, const float* gamma,
                        const float* beta, float eps, float* scale, float* shift, tcv_stream_t) {
  GnFinalizeP p{sums, n, c, groups, pixels, gamma, beta, eps, scale, shift, copies, clear};
  return run_body<GnFinalizeP, gn_finalize_body>(p, (ll)n * groups);
}

int tcv_gn_apply(const void* x, long long x_plane, int n, long long pixels, int c, const float* scale,
                 const float* shift, const void* res, long long res_plane, int act, void* y, long long y_plane,
                 int y_c, int y_off, tcv_stream_t) {
  REQ(c % 8 == 0 && y_c % 8 == 0 && y_off % 8 == 0 && y_off + c <= y_c, "gn_apply: dims");
  if (x_plane == 0) x_plane = (ll)n * pixels * c;
  if (res && res_plane == 0) res_plane = (ll)n * pixels * c;
  if (y_plane == 0) y_plane = (ll)n * pixels * y_c;
  GnApplyP p{(const uint16_t*)x, x_plane, n, c, pixels, scale, shift, (const uint16_t*)res, res_plane, act,
             (uint16_t*)y, y_plane, y_c, y_off};
  return run_body<GnApplyP, gn_apply_body>(p, (ll)n * pixels * (c / 8));
}

int tcv_maxpool3s2(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t) {
  PoolP p{(const uint16_t*)x, n, h, w, c, (h - 1) / 2 + 1, (w - 1) / 2 + 1, (uint16_t*)y};
  return run_body<PoolP, maxpool3s2_body>(p, (ll)n * p.oh * p.ow * (c / 8));
}

int tcv_adaptive_avgpool(const void* x, long long x_plane, int n, int h, int w, int c, int x_c, int x_off, int s,
                         void* y, tcv_stream_t) {
  REQ(c % 64 == 0 && x_off + c <= x_c, "adaptive_avgpool: dims");
  if (x_plane == 0) x_plane = (ll)n * h * w * x_c;
  const ll yplane = (ll)n * s * s * c;
  for (int img = 0; img < n; ++img)
    for (int i = 0; i < s; ++i)
      for (int j = 0; j < s; ++j) {
        const int y0 = bin_start(i, h, s), y1 = bin_end(i, h, s);   // the bin arithmetic the CUDA kernel uses
        const int x0 = bin_start(j, w, s), x1 = bin_end(j, w, s);
        for (int ch = 0; ch < c; ++ch) {
          float a = 0.f;
          for (int yy = y0; yy < y1; ++yy)
            for (int xx = x0; xx < x1; ++xx)
              a += ld1((const uint16_t*)x + (((ll)img * h + yy) * w + xx) * x_c + x_off + ch, x_plane);
          st1((uint16_t*)y + (((ll)img * s + i) * s + j) * c + ch, yplane, a / (float)((y1 - y0) * (x1 - x0)));
        }
      }
  ++g_launches;
  return 0;
}

int tcv_bilinear(const void* x, int n, int ih, int iw, int c, void* y, long long y_plane, int oh, int ow, int y_c,
                 int y_off, tcv_stream_t) {
  REQ(c % 8 == 0 && y_c % 8 == 0 && y_off % 8 == 0 && y_off + c <= y_c, "bilinear: dims");
  if (y_plane == 0) y_plane = (ll)n * oh * ow * y_c;
  BilinearP p{(const uint16_t*)x, n, ih, iw, c, (uint16_t*)y, y_plane, oh, ow, y_c, y_off};
  return run_body<BilinearP, bilinear_body>(p, (ll)n * oh * ow * (c / 8));
}

int tcv_copy_channels(const void* x, long long x_plane, int x_c, int x_off, void* y, long long y_plane, int y_c,
                      int y_off, int c, long long pixels, tcv_stream_t) {
  REQ(c % 8 == 0 && x_off + c <= x_c && y_off + c <= y_c, "copy_channels: dims");
  if (x_plane == 0) x_plane = pixels * x_c;
  if (y_plane == 0) y_plane = pixels * y_c;
  CopyP p{(const uint16_t*)x, x_plane, x_c, x_off, (uint16_t*)y, y_plane, y_c, y_off, c, pixels};
  return run_body<CopyP, copy_channels_body>(p, pixels * (c / 8));
}

int tcv_fba_encode_inputs(const void* imgs, const void* tris, int is_u8, int frames, int h, int w, void* x16,
                          tcv_stream_t) {
  EncodeP p{imgs, tris, is_u8, frames, h, w, (uint16_t*)x16};
  return run_body<EncodeP, fba_encode_body>(p, (ll)frames * h * w);
}

int tcv_fba_edt_cols(const void* x16, int frames, int h, int w, int* g, tcv_stream_t) {
  EdtP p{(uint16_t*)x16, frames, h, w, g};
  return run_body<EdtP, fba_edt_cols_body>(p, (ll)frames * 2 * w);
}

int tcv_fba_edt_rows(const int* g, int frames, int h, int w, void* x16, tcv_stream_t) {
  EdtP p{(uint16_t*)x16, frames, h, w, (int*)g};
  return run_body<EdtP, fba_edt_rows_body>(p, (ll)frames * 2 * h * w);
}

int tcv_fba_cat_inputs(const void* x16, long long x16_plane, long long pixels, void* y, long long y_plane, int y_c,
                       int y_off, tcv_stream_t) {
  REQ(y_off + 32 <= y_c, "fba_cat_inputs: dims");
  if (y_plane == 0) y_plane = pixels * y_c;
  if (x16_plane == 0) x16_plane = pixels * 16;
  CatP p{(const uint16_t*)x16, x16_plane, pixels, (uint16_t*)y, y_plane, y_c, y_off};
  return run_body<CatP, fba_cat_inputs_body>(p, pixels);
}

int tcv_fba_fusion(const void* o8, const void* x16, long long x16_plane, long long x16_img_stride, int n, int h, int w,
                   float* pred, tcv_stream_t) {
  if (x16_img_stride == 0) x16_img_stride = (ll)h * w * 16;
  FusionP p{(const uint16_t*)o8, (const uint16_t*)x16, x16_plane, x16_img_stride, n, h, w, pred};
  return run_body<FusionP, fba_fusion_body>(p, (ll)n * h * w);
}

int tcv_frame_metrics(const uint8_t* alpha, const uint8_t* gt, const uint8_t* tri, const uint8_t* next_alpha,
                      const uint8_t* next_gt, const float* flow, int h, int w, double* out, tcv_stream_t) {
  float lut[256];
  for (int v = 0; v < 256; ++v) lut[v] = metric_u8((uint8_t)v);
  MetricP p{alpha, gt, tri, next_alpha, next_gt, flow, h, w, lut};
  for (int k = 0; k < 7; ++k) out[k] = 0.0;
  for (ll i = 0; i < (ll)h * w; ++i) metric_body(i, p, out);
  ++g_launches;
  return 0;
}

int tcv_dwconv3x3(const void* x, int n, int h, int w, int c, int dil, const float* wt, const float* scale, const float* shift,
                  const float* border, int act, void* y, tcv_stream_t) {
  REQ(c % 8 == 0 && dil >= 1, "dwconv3x3: dims");
  DwConvP p{(const uint16_t*)x, n, h, w, c, dil, wt, scale, shift, border, act, (uint16_t*)y};
  return run_body<DwConvP, dwconv3x3_body>(p, (ll)n * h * w * (c / 8));
}

int tcv_index_finish(const void* b0, const void* b1, const void* b2, const void* b3, int n, int h2, int w2, int c,
                     void* idx_en, void* idx_de, tcv_stream_t) {
  IndexFinishP p{{(const uint16_t*)b0, (const uint16_t*)b1, (const uint16_t*)b2, (const uint16_t*)b3}, n, h2, w2, c,
                 (uint16_t*)idx_en, (uint16_t*)idx_de};
  return run_body<IndexFinishP, index_finish_body>(p, (ll)n * h2 * w2 * (c / 8));
}

int tcv_index_pool(const void* x, const void* idx_en, int n, int h, int w, int c, void* masked, void* pooled, tcv_stream_t) {
  IndexPoolP p{(const uint16_t*)x, (const uint16_t*)idx_en, n, h, w, c, (uint16_t*)masked, (uint16_t*)pooled};
  return run_body<IndexPoolP, index_pool_body>(p, (ll)n * (h / 2) * (w / 2) * (c / 8));
}

int tcv_index_upcat(const void* dec, int dec_c, int dec_real, int up, const void* idx, int idx_c, long long idx_plane,
                    const void* low, int low_c, long long low_plane, int low_real, int n, int h, int w, int cat_c, void* cat,
                    tcv_stream_t) {
  if (idx_plane == 0) idx_plane = (ll)n * h * w * idx_c;
  if (low_plane == 0) low_plane = (ll)n * h * w * low_c;
  IndexUpcatP p{(const uint16_t*)dec, (const uint16_t*)idx, (const uint16_t*)low, n, h, w, up, dec_c, dec_real, idx_c, low_c,
                low_real, cat_c, (uint16_t*)cat, idx_plane, low_plane};
  return run_body<IndexUpcatP, index_upcat_body>(p, (ll)n * h * w * (cat_c / 8));
}

int tcv_maxpool2_idx(const void* x, int n, int h, int w, int c, void* y, uint8_t* idx, tcv_stream_t) {
  REQ(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "maxpool2_idx: dims");
  Pool2P p{(const uint16_t*)x, n, h, w, c, (uint16_t*)y, idx};
  return run_body<Pool2P, maxpool2_idx_body>(p, (ll)n * (h / 2) * (w / 2) * (c / 8));
}

int tcv_maxunpool2(const void* x, const uint8_t* idx, int n, int h, int w, int c, void* y, tcv_stream_t) {
  REQ(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "maxunpool2: dims");
  Unpool2P p{(const uint16_t*)x, idx, n, h, w, c, (uint16_t*)y};
  return run_body<Unpool2P, maxunpool2_body>(p, (ll)n * h * w * (c / 8));
}

int tcv_dim_fix_inputs(const void* tris, int is_u8, int frames, int h, int w, void* x8, tcv_stream_t) {
  DimFixP p{tris, is_u8, (ll)frames * h * w, (uint16_t*)x8};
  return run_body<DimFixP, dim_fix_inputs_body>(p, p.pixels);
}

int tcv_space_to_depth2(const void* x, long long x_plane, int n, int h, int w, int c, void* y, tcv_stream_t) {
  REQ(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "space_to_depth2: dims");
  if (x_plane == 0) x_plane = (ll)n * h * w * c;
  S2dP p{(const uint16_t*)x, x_plane, n, h, w, c, (uint16_t*)y};
  return run_body<S2dP, space_to_depth2_body>(p, (ll)n * h * w * (c / 8));
}

int tcv_s2d_pack_stem(const float* w49, int cin_pad, int cout, float* out, tcv_stream_t) {
  S2dPackP p{w49, cin_pad, cout, out};
  return run_body<S2dPackP, s2d_pack_stem_body>(p, (ll)16 * 4 * cin_pad * cout);
}

static int floordiv2(int a) { return a >= 0 ? a / 2 : -((-a + 1) / 2); }
int tcv_s2d_pack(const float* src, int k, int pad, int cin_src, int cout_src, int cin_dst, int cout_dst, float* out,
                 tcv_stream_t) {
  const int t0 = floordiv2(-pad), T = floordiv2(k - 1 - pad) - t0 + 1;
  S2dPackGenP p{src, k, pad, t0, T, cin_src, cout_src, cin_dst, cout_dst, out};
  return run_body<S2dPackGenP, s2d_pack_body>(p, (ll)T * T * 4 * cin_dst * cout_dst);
}

int tcv_postprocess_eval_fba(const float* pred, const void* imgs, const void* tris, int is_u8, const float* trimask,
                             int batch, int frames, int h, int w, float* alphas, float* Fs, float* Bs, tcv_stream_t) {
  PostP p{pred, imgs, tris, is_u8, trimask, batch, frames, h, w, alphas, Fs, Bs};
  return run_body<PostP, postprocess_fba_body>(p, (ll)batch * frames * h * w);
}

}  // extern "C"
