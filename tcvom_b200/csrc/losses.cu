// FullModel / FullModel_VMD task-wrapper kernels (models/model.py:54-127, 258-357):
//   * train-side preprocess: composite image, trimap synthesis (threshold, unknown band, per-sample
//     max-pool dilation), one-hot trimap encoding, visualisation tensors
//   * fused forward of the three losses: L_im (masked L1, loss_func.py:9-22), L_tc (dtSSD) and
//     L_af (BCE-with-logits on the TAM logits against the 7x7 alpha-difference targets)
// All are single-pass, HBM-bound, with warp-shuffle + double atomics for the reductions.
#include "common.cuh"

namespace tcv {

static inline unsigned nblocks(long long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

__device__ __forceinline__ float thr_alpha(float g, float eps) {
  float a = g < eps ? 0.f : g;
  a = a > 1.0f - eps ? 1.0f : a;
  return a;
}

__global__ void train_trimask_raw_kernel(const float* __restrict__ a, long long total, float eps,
                                         uint8_t* __restrict__ m) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float al = thr_alpha(a[i] * (1.0f / 255), eps);
  m[i] = (al > 0.f) & (al < 1.0f);
}

// separable max filter with a per-sample radius (radii[b], b = frame / S)
__global__ void dilate_row_rad_kernel(const uint8_t* __restrict__ in, int frames, int S, int h, int w,
                                      const int* __restrict__ radii, uint8_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)frames * h * w) return;
  const int x = (int)(i % w);
  const int r = radii[(int)(i / ((long long)h * w)) / S];
  const long long row = i - x;
  uint8_t v = 0;
  for (int k = max(x - r, 0); k <= min(x + r, w - 1); ++k) v |= in[row + k];
  out[i] = v;
}
__global__ void dilate_col_rad_kernel(const uint8_t* __restrict__ in, int frames, int S, int h, int w,
                                      const int* __restrict__ radii, uint8_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)frames * h * w) return;
  const int x = (int)(i % w);
  const int y = (int)((i / w) % h);
  const int r = radii[(int)(i / ((long long)h * w)) / S];
  const long long img = i - (long long)y * w - x;
  uint8_t v = 0;
  for (int k = max(y - r, 0); k <= min(y + r, h - 1); ++k) v |= in[img + (long long)k * w + x];
  out[i] = v;
}

__global__ void train_preprocess_kernel(const float* __restrict__ a, const float* __restrict__ fg,
                                        const float* __restrict__ bg, const uint8_t* __restrict__ mask, int frames,
                                        int h, int w, float eps, __nv_bfloat16* __restrict__ x8,
                                        float* __restrict__ trimask, float* __restrict__ gts, float* __restrict__ fgs,
                                        float* __restrict__ bgs, float* __restrict__ imgs, float* __restrict__ tris_vis) {
  const long long hw = (long long)h * w;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * hw) return;
  const long long f = i / hw, p = i - f * hw;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  const float g = a[i] * (1.0f / 255);
  float v[8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const float F = fg[(f * 3 + (2 - c)) * hw + p] * (1.0f / 255);   // flip([2]): BGR -> RGB
    const float B = bg[(f * 3 + (2 - c)) * hw + p] * (1.0f / 255);
    const float im = F * g + B * (1.0f - g);
    fgs[(f * 3 + c) * hw + p] = F;
    bgs[(f * 3 + c) * hw + p] = B;
    imgs[(f * 3 + c) * hw + p] = im;
    v[c] = (im - mean[c]) / stdv[c];
  }
  const bool m = mask[i] != 0;
  const float al = thr_alpha(g, eps);
  const int cls = m ? 1 : (int)(2.0f * al);
  v[3] = cls == 0 ? 1.f : 0.f;
  v[4] = cls == 1 ? 1.f : 0.f;
  v[5] = cls == 2 ? 1.f : 0.f;
  v[6] = 0.f;
  v[7] = 0.f;
  store8(x8 + i * 8, frames * hw * 8, v);
  trimask[i] = m ? 1.f : 0.f;
  gts[i] = g;
  tris_vis[i] = m ? 128.0f * (1.0f / 255) : g;
}

// ---- L_im + L_tc + visual outputs --------------------------------------------------------------------
// acc layout (double): [0, S) sum|refine-gt|*m per frame, [S, 2S) count(m) per frame,
//                      [2S, 3S) sum|dadt-dgtdt|*m per frame c (pair c, c+1), [3S,4S) L_af sum (b), [4S,5S) L_af sum (f),
//                      [5S, 6S) unknown count at OS8 per frame
__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

__global__ void loss_im_tc_kernel(const float* __restrict__ pred, const float* __restrict__ trimask,
                                  const float* __restrict__ gts, const float* __restrict__ fgs,
                                  const float* __restrict__ bgs, int B, int S, int h, int w,
                                  float* __restrict__ alphas, float* __restrict__ comps, double* __restrict__ acc) {
  const long long hw = (long long)h * w;
  const long long total = (long long)B * S * hw;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  // whole warps share (b, s): hw is a multiple of 32 (H, W multiples of 32)
  const long long f = min(i, total - 1) / hw, p = min(i, total - 1) - f * hw;
  const int b = (int)(f / S), s = (int)(f % S);
  double l1 = 0, cnt = 0, ldt = 0;
  if (i < total) {
    float al = 0.f;
    float cm[3] = {0.f, 0.f, 0.f};
    if (s > 0 && s < S - 1) {
      const float m = trimask[i];
      const float g = gts[i];
      const float pr = pred[((long long)b * (S - 2) + (s - 1)) * hw + p];
      const float refine = m != 0.f ? pr : g;                       // model.py:102
      l1 = fabsf(refine - g) * m;
      cnt = m > 1.001e-5f ? 1.0 : 0.0;
#pragma unroll
      for (int c = 0; c < 3; ++c)
        cm[c] = fgs[(f * 3 + c) * hw + p] * refine + bgs[(f * 3 + c) * hw + p] * (1.0f - refine);
      al = fminf(fmaxf(refine, 0.f), 1.f);
      if (S >= 5 && s < S - 2) {                                     // _dtSSD pairs (c, c+1), c in [1, S-3]
        const long long j = i + hw;
        const float m1 = trimask[j], g1 = gts[j];
        const float pr1 = pred[((long long)b * (S - 2) + s) * hw + p];
        const float al1 = fminf(fmaxf(m1 != 0.f ? pr1 : g1, 0.f), 1.f);
        ldt = fabsf((al - al1) - (g - g1)) * m;
      }
    }
    alphas[i] = al;
#pragma unroll
    for (int c = 0; c < 3; ++c) comps[(f * 3 + c) * hw + p] = fminf(fmaxf(cm[c], 0.f), 1.f);
  }
  l1 = warp_sum_d(l1);
  cnt = warp_sum_d(cnt);
  ldt = warp_sum_d(ldt);
  if ((threadIdx.x & 31) == 0 && i < total + 31) {
    if (l1 != 0) atomicAdd(acc + s, l1);
    if (cnt != 0) atomicAdd(acc + S + s, cnt);
    if (ldt != 0) atomicAdd(acc + 2 * S + s, ldt);
  }
}

__global__ void avgpool8_kernel(const float* __restrict__ x, int frames, int h, int w, float* __restrict__ y) {
  const int oh = h / 8, ow = w / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)frames * oh * ow) return;
  const int ox = (int)(i % ow), oy = (int)((i / ow) % oh);
  const long long f = i / ((long long)ow * oh);
  const float* src = x + (f * h + oy * 8) * w + ox * 8;
  float s = 0.f;
  for (int dy = 0; dy < 8; ++dy) {
    const float4 a = *reinterpret_cast<const float4*>(src + (long long)dy * w);
    const float4 b = *reinterpret_cast<const float4*>(src + (long long)dy * w + 4);
    s += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
  }
  y[i] = s * (1.0f / 64);
}

// L_af: one thread per (b, centre c, OS8 pixel, window offset j).  att* fp32 [B, S-2, w2, N8] raw logits,
// small_mask uint8 [B, S-2, N8], gt8 fp32 [B, S, h8, w8].
__global__ void loss_af_kernel(const float* __restrict__ attb, const float* __restrict__ attf,
                               const uint8_t* __restrict__ small_mask, const float* __restrict__ gt8, int B, int S,
                               int h8, int w8, int window, float thres, float smooth, double* __restrict__ acc) {
  const int N8 = h8 * w8, w2 = window * window, r = window / 2;
  const long long total = (long long)B * (S - 2) * w2 * N8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long ii = min(i, total - 1);
  const int n = (int)(ii % N8);
  const int j = (int)((ii / N8) % w2);
  const int ci = (int)((ii / ((long long)N8 * w2)) % (S - 2));
  const int b = (int)(ii / ((long long)N8 * w2 * (S - 2)));
  const int s = ci + 1;
  double lb = 0, lf = 0, cnt = 0;
  if (i < total && small_mask[((long long)b * (S - 2) + ci) * N8 + n]) {
    const int y = n / w8, x = n - y * w8;
    const int yy = y + j / window - r, xx = x + j % window - r;
    const bool in = yy >= 0 && yy < h8 && xx >= 0 && xx < w8;
    const float* g = gt8 + (long long)b * S * N8;
    const float cg = g[(long long)s * N8 + n];
    const float bgv = in ? g[(long long)(s - 1) * N8 + yy * w8 + xx] : 0.f;   // F.unfold zero padding
    const float fgv = in ? g[(long long)(s + 1) * N8 + yy * w8 + xx] : 0.f;
    const float tb = fabsf(cg - bgv) < thres ? 1.0f - smooth : 0.f;
    const float tf = fabsf(cg - fgv) < thres ? 1.0f - smooth : 0.f;
    const float xb = attb[ii], xf = attf[ii];
    // BCEWithLogits: max(x,0) - x*t + log(1 + exp(-|x|))
    lb = fmaxf(xb, 0.f) - xb * tb + log1pf(expf(-fabsf(xb)));
    lf = fmaxf(xf, 0.f) - xf * tf + log1pf(expf(-fabsf(xf)));
    cnt = j == 0 ? 1.0 : 0.0;
  }
  // a warp may straddle (c, b) boundaries only if w2*N8 is not a multiple of 32: reduce per thread group safely
  const int key = b * (S - 2) + ci;
  const int key0 = __shfl_sync(0xffffffffu, key, 0);
  if (__all_sync(0xffffffffu, key == key0)) {
    lb = warp_sum_d(lb); lf = warp_sum_d(lf); cnt = warp_sum_d(cnt);
    if ((threadIdx.x & 31) == 0) {
      if (lb != 0) atomicAdd(acc + 3 * S + s, lb);
      if (lf != 0) atomicAdd(acc + 4 * S + s, lf);
      if (cnt != 0) atomicAdd(acc + 5 * S + s, cnt);
    }
  } else if (i < total) {
    if (lb != 0) atomicAdd(acc + 3 * S + s, lb);
    if (lf != 0) atomicAdd(acc + 4 * S + s, lf);
    if (cnt != 0) atomicAdd(acc + 5 * S + s, cnt);
  }
}

__global__ void loss_finalize_kernel(const double* __restrict__ acc, int B, int S, int h, int w, int w2, float mult,
                                     float* __restrict__ out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  const double eps = 1.001e-5, cap = (double)B * h * w + 1.0;
  double la = 0, ldt = 0, laf = 0;
  for (int c = 1; c < S - 1; ++c) {
    const double safe = fmin(fmax(acc[S + c], eps), cap);
    la += acc[c] / safe;
    if (S >= 5 && c < S - 2) ldt += acc[2 * S + c] / safe;
    const double u = acc[5 * S + c];                       // unknown OS8 pixels (over the batch)
    if (u > 0) laf += 0.5 * (acc[3 * S + c] + acc[4 * S + c]) / (u * w2);
  }
  out[0] = (float)(la / (S - 2));       // L_alpha
  out[1] = 0.f;                         // L_comp  (GCA: alpha loss only, model.py:112-114)
  out[2] = 0.f;                         // L_grad
  out[3] = S >= 5 ? (float)(ldt / (S - 3)) : 0.f;
  out[4] = (float)(laf / (S - 2)) * mult;
}


// ---- loss gradients (autograd of L_im / L_tc / L_af) ----------------------------------------------------
__device__ __forceinline__ float sgn(float v) { return v > 0.f ? 1.f : (v < 0.f ? -1.f : 0.f); }
__device__ __forceinline__ double safe_count(const double* acc, int S, int c, int B, int h, int w) {
  return fmin(fmax(acc[S + c], 1.001e-5), (double)B * h * w + 1.0);
}

// dpred[b, ci, p] for the inner frame s = ci + 1 (only unknown-region pixels carry gradient, model.py:102)
__global__ void loss_im_tc_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ trimask,
                                      const float* __restrict__ gts, const double* __restrict__ acc,
                                      const float* __restrict__ gl, int B, int S, int h, int w,
                                      float* __restrict__ dpred) {
  const long long hw = (long long)h * w;
  const long long total = (long long)B * (S - 2) * hw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i % hw;
  const int ci = (int)((i / hw) % (S - 2));
  const int b = (int)(i / (hw * (S - 2)));
  const int s = ci + 1;
  const long long f = ((long long)b * S + s) * hw + p;
  const float m = trimask[f];
  float grad = 0.f;
  if (m != 0.f) {
    const float pr = pred[i], g = gts[f];
    const float g_alpha = gl[0], g_dt = gl[3];
    grad = g_alpha * sgn(pr - g) * m / (float)(safe_count(acc, S, s, B, h, w) * (S - 2));
    if (S >= 5 && pr >= 0.f && pr <= 1.f) {
      const float al = fminf(fmaxf(pr, 0.f), 1.f);
      if (s <= S - 3) {                                   // pair (s, s+1), masked by frame s
        const float m1 = trimask[f + hw], g1 = gts[f + hw];
        const float al1 = fminf(fmaxf(m1 != 0.f ? pred[i + hw] : g1, 0.f), 1.f);
        const float D = (al - al1) - (g - g1);
        grad += g_dt * sgn(D) * m / (float)(safe_count(acc, S, s, B, h, w) * (S - 3));
      }
      if (s - 1 >= 1) {                                   // pair (s-1, s), masked by frame s-1
        const float m0 = trimask[f - hw], g0 = gts[f - hw];
        if (m0 != 0.f) {
          const float al0 = fminf(fmaxf(pred[i - hw], 0.f), 1.f);   // m0 != 0: refine = pred
          const float D = (al0 - al) - (g0 - g);
          grad -= g_dt * sgn(D) * m0 / (float)(safe_count(acc, S, s - 1, B, h, w) * (S - 3));
        }
      }
    }
  }
  dpred[i] = grad;
}

__global__ void loss_af_bwd_kernel(const float* __restrict__ attb, const float* __restrict__ attf,
                                   const uint8_t* __restrict__ small_mask, const float* __restrict__ gt8,
                                   const double* __restrict__ acc, const float* __restrict__ gl, int B, int S, int h8,
                                   int w8, int window, float thres, float smooth, float mult,
                                   float* __restrict__ dattb, float* __restrict__ dattf) {
  const int N8 = h8 * w8, w2 = window * window, r = window / 2;
  const long long total = (long long)B * (S - 2) * w2 * N8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int n = (int)(i % N8);
  const int j = (int)((i / N8) % w2);
  const int ci = (int)((i / ((long long)N8 * w2)) % (S - 2));
  const int b = (int)(i / ((long long)N8 * w2 * (S - 2)));
  const int s = ci + 1;
  float gb = 0.f, gf = 0.f;
  if (small_mask[((long long)b * (S - 2) + ci) * N8 + n]) {
    const int y = n / w8, x = n - y * w8;
    const int yy = y + j / window - r, xx = x + j % window - r;
    const bool in = yy >= 0 && yy < h8 && xx >= 0 && xx < w8;
    const float* g = gt8 + (long long)b * S * N8;
    const float cg = g[(long long)s * N8 + n];
    const float bgv = in ? g[(long long)(s - 1) * N8 + yy * w8 + xx] : 0.f;
    const float fgv = in ? g[(long long)(s + 1) * N8 + yy * w8 + xx] : 0.f;
    const float tb = fabsf(cg - bgv) < thres ? 1.0f - smooth : 0.f;
    const float tf = fabsf(cg - fgv) < thres ? 1.0f - smooth : 0.f;
    const float xb = attb[i], xf = attf[i];
    // d/dx BCEWithLogits(mean over u*w2 elements) = (sigmoid(x) - t) / (u*w2); L_att = mean_c (Lb + Lf)/2 * mult
    const float k = gl[4] * mult * 0.5f / (float)(acc[5 * S + s] * w2 * (S - 2));
    gb = k * (1.0f / (1.0f + expf(-xb)) - tb);
    gf = k * (1.0f / (1.0f + expf(-xf)) - tf);
  }
  dattb[i] = gb;
  dattf[i] = gf;
}

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_preprocess_train(const float* a, const float* fg, const float* bg, int batch, int frames_per_sample, int h,
                         int w, float eps, const int* radii, void* x8, float* trimask, float* gts, float* fgs,
                         float* bgs, float* imgs, float* tris_vis, uint8_t* tmp, tcv_stream_t stream) {
  TCV_REQUIRE(a && fg && bg && radii && x8 && trimask && gts && fgs && bgs && imgs && tris_vis && tmp,
              "preprocess_train: null pointer");
  const int frames = batch * frames_per_sample;
  const long long total = (long long)frames * h * w;
  uint8_t* m0 = tmp;
  uint8_t* m1 = tmp + total;
  train_trimask_raw_kernel<<<nblocks(total), 256, 0, S(stream)>>>(a, total, eps, m0);
  int rc = launched("train_trimask_raw_kernel");
  if (rc) return rc;
  dilate_row_rad_kernel<<<nblocks(total), 256, 0, S(stream)>>>(m0, frames, frames_per_sample, h, w, radii, m1);
  if ((rc = launched("dilate_row_rad_kernel"))) return rc;
  dilate_col_rad_kernel<<<nblocks(total), 256, 0, S(stream)>>>(m1, frames, frames_per_sample, h, w, radii, m0);
  if ((rc = launched("dilate_col_rad_kernel"))) return rc;
  train_preprocess_kernel<<<nblocks(total), 256, 0, S(stream)>>>(a, fg, bg, m0, frames, h, w, eps,
                                                                reinterpret_cast<__nv_bfloat16*>(x8), trimask, gts,
                                                                fgs, bgs, imgs, tris_vis);
  return launched("train_preprocess_kernel");
}

int tcv_losses_vmd(const float* pred, const float* trimask, const float* gts, const float* fgs, const float* bgs,
                   const float* attb, const float* attf, const uint8_t* small_mask, int batch, int frames_per_sample,
                   int h, int w, int window, float att_thres, float label_smooth, float att_multiplier,
                   float* alphas, float* comps, float* gt8, double* acc, float* losses, tcv_stream_t stream) {
  TCV_REQUIRE(pred && trimask && gts && fgs && bgs && alphas && comps && gt8 && acc && losses,
              "losses_vmd: null pointer");
  TCV_REQUIRE(frames_per_sample >= 3 && h % 32 == 0 && w % 32 == 0, "losses_vmd: need S >= 3 and H,W %% 32 == 0");
  const int Sn = frames_per_sample;
  TCV_CUDA(cudaMemsetAsync(acc, 0, sizeof(double) * 6 * Sn, S(stream)));
  const long long total = (long long)batch * Sn * h * w;
  loss_im_tc_kernel<<<nblocks(total), 256, 0, S(stream)>>>(pred, trimask, gts, fgs, bgs, batch, Sn, h, w, alphas, comps,
                                                          acc);
  int rc = launched("loss_im_tc_kernel");
  if (rc) return rc;
  if (attb && attf && small_mask) {
    const int h8 = h / 8, w8 = w / 8;
    const long long t8 = (long long)batch * Sn * h8 * w8;
    avgpool8_kernel<<<nblocks(t8), 256, 0, S(stream)>>>(gts, batch * Sn, h, w, gt8);
    if ((rc = launched("avgpool8_kernel"))) return rc;
    const long long ta = (long long)batch * (Sn - 2) * window * window * h8 * w8;
    loss_af_kernel<<<nblocks(ta), 256, 0, S(stream)>>>(attb, attf, small_mask, gt8, batch, Sn, h8, w8, window, att_thres,
                                                      label_smooth, acc);
    if ((rc = launched("loss_af_kernel"))) return rc;
  }
  loss_finalize_kernel<<<1, 32, 0, S(stream)>>>(acc, batch, Sn, h, w, window * window, att_multiplier, losses);
  return launched("loss_finalize_kernel");
}

int tcv_losses_vmd_bwd(const float* pred, const float* trimask, const float* gts, const float* attb,
                       const float* attf, const uint8_t* small_mask, const float* gt8, const double* acc,
                       const float* gl, int batch, int frames_per_sample, int h, int w, int window, float att_thres,
                       float label_smooth, float att_multiplier, float* dpred, float* dattb, float* dattf,
                       tcv_stream_t stream) {
  TCV_REQUIRE(pred && trimask && gts && acc && gl && dpred, "losses_vmd_bwd: null pointer");
  TCV_REQUIRE(frames_per_sample >= 3, "losses_vmd_bwd: need S >= 3");
  const int Sn = frames_per_sample;
  const long long total = (long long)batch * (Sn - 2) * h * w;
  loss_im_tc_bwd_kernel<<<nblocks(total), 256, 0, S(stream)>>>(pred, trimask, gts, acc, gl, batch, Sn, h, w, dpred);
  int rc = launched("loss_im_tc_bwd_kernel");
  if (rc) return rc;
  if (attb && attf && small_mask && dattb && dattf) {
    TCV_REQUIRE(gt8, "losses_vmd_bwd: gt8 workspace of the forward required");
    const int h8 = h / 8, w8 = w / 8;
    const long long ta = (long long)batch * (Sn - 2) * window * window * h8 * w8;
    loss_af_bwd_kernel<<<nblocks(ta), 256, 0, S(stream)>>>(attb, attf, small_mask, gt8, acc, gl, batch, Sn, h8, w8,
                                                          window, att_thres, label_smooth, att_multiplier, dattb, dattf);
    rc = launched("loss_af_bwd_kernel");
  }
  return rc;
}

}  // extern "C"
