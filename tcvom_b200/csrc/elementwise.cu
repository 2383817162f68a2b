// HBM-bound elementwise / small reduction kernels: pre/post-processing, pooling, weight folding,
// layout adapters.  All are single-pass, coalesced, one launch each.
#include "common.cuh"

namespace tcv {

// ---------------------------------------------------------------------------------------------
// EvalModel.preprocess (models/model.py:360-387, TRIMAP_CHANNEL==3 branch)
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void trimask_raw_kernel(const T* __restrict__ tris, long long total, uint8_t* __restrict__ m) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const float st = (float)tris[i] * (1.0f / 255);
  m[i] = (st > 0.f) & (st < 1.f);
}

// separable max filter of radius r (F.max_pool2d(k=2r+1, stride 1, pad r) on a 0/1 mask)
__global__ void dilate_row_kernel(const uint8_t* __restrict__ in, int frames, int h, int w, int r,
                                  uint8_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)frames * h * w) return;
  const int x = (int)(i % w);
  const long long row = i - x;
  uint8_t v = 0;
  const int lo = max(x - r, 0), hi = min(x + r, w - 1);
  for (int k = lo; k <= hi; ++k) v |= in[row + k];
  out[i] = v;
}
__global__ void dilate_col_kernel(const uint8_t* __restrict__ in, int frames, int h, int w, int r,
                                  uint8_t* __restrict__ out) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)frames * h * w) return;
  const int x = (int)(i % w);
  const int y = (int)((i / w) % h);
  const long long img = i - (long long)y * w - x;
  uint8_t v = 0;
  const int lo = max(y - r, 0), hi = min(y + r, h - 1);
  for (int k = lo; k <= hi; ++k) v |= in[img + (long long)k * w + x];
  out[i] = v;
}

template <typename T>
__global__ void preprocess_kernel(const T* __restrict__ imgs, const T* __restrict__ tris,
                                  const uint8_t* __restrict__ mask, int frames, int h, int w,
                                  __nv_bfloat16* __restrict__ x8, float* __restrict__ trimask) {
  const long long hw = (long long)h * w;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= frames * hw) return;
  const long long f = i / hw, p = i - f * hw;
  const float mean[3] = {0.485f, 0.456f, 0.406f};
  const float stdv[3] = {0.229f, 0.224f, 0.225f};
  float v[8];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    // flip([2]) : output channel c (RGB) <- input channel 2-c (BGR)
    const float s = (float)imgs[(f * 3 + (2 - c)) * hw + p] * (1.0f / 255);
    v[c] = (s - mean[c]) / stdv[c];
  }
  const float st = (float)tris[i] * (1.0f / 255);
  const bool m = mask[i] != 0;
  const int cls = m ? 1 : (int)(2.0f * st);  // .long() truncation (model.py:379)
  v[3] = cls == 0 ? 1.f : 0.f;
  v[4] = cls == 1 ? 1.f : 0.f;
  v[5] = cls == 2 ? 1.f : 0.f;
  v[6] = 0.f;
  v[7] = 0.f;
  store8(x8 + i * 8, frames * hw * 8, v);
  trimask[i] = m ? 1.f : 0.f;
}

template <typename T>
__global__ void postprocess_kernel(const float* __restrict__ pred, const T* __restrict__ tris,
                                   const float* __restrict__ trimask, int batch, int frames, int h, int w,
                                   float* __restrict__ alphas) {
  const long long hw = (long long)h * w;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)batch * frames * hw) return;
  const long long f = i / hw, p = i - f * hw;
  const int b = (int)(f / frames), s = (int)(f % frames);
  float a = 0.f;
  if (s > 0 && s < frames - 1) {
    const float pr = pred[((long long)b * (frames - 2) + (s - 1)) * hw + p];
    a = trimask[i] != 0.f ? pr : (float)tris[i] * (1.0f / 255);
  }
  alphas[i] = a;
}

// ---------------------------------------------------------------------------------------------
__global__ void avgpool2_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c,
                                __nv_bfloat16* __restrict__ y) {
  const int oh = h / 2, ow = w / 2, c8 = c / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * oh * ow * c8;
  if (i >= total) return;
  const int cc = (int)(i % c8) * 8;
  long long t = i / c8;
  const int ox = (int)(t % ow);
  t /= ow;
  const int oy = (int)(t % oh);
  const int img = (int)(t / oh);
  const long long iplane = (long long)n * h * w * c;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
  for (int dy = 0; dy < 2; ++dy)
#pragma unroll
    for (int dx = 0; dx < 2; ++dx) {
      float f[8];
      load8(x + (((long long)img * h + 2 * oy + dy) * w + 2 * ox + dx) * c + cc, iplane, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) acc[k] *= 0.25f;
  store8(y + (((long long)img * oh + oy) * ow + ox) * c + cc, (long long)n * oh * ow * c, acc);
}

// 1-pixel reflect border (nn.ReflectionPad2d(1)) materialised once so that the following stride-2 conv can
// run on the TMA/tcgen05 path (TMA zero-fills out-of-range coordinates, it cannot reflect)
__global__ void pad_reflect1_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c,
                                    __nv_bfloat16* __restrict__ y) {
  const int oh = h + 2, ow = w + 2, c8 = c / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * oh * ow * c8;
  if (i >= total) return;
  const int cc = (int)(i % c8) * 8;
  long long t = i / c8;
  const int ox = (int)(t % ow);
  t /= ow;
  const int oy = (int)(t % oh);
  const int img = (int)(t / oh);
  const int iy = reflect(oy - 1, h), ix = reflect(ox - 1, w);
  const long long ip = (long long)n * h * w * c, op = (long long)n * oh * ow * c;
  const __nv_bfloat16* src = x + (((long long)img * h + iy) * w + ix) * c + cc;
  __nv_bfloat16* dst = y + (((long long)img * oh + oy) * ow + ox) * c + cc;
  *reinterpret_cast<uint4*>(dst) = *reinterpret_cast<const uint4*>(src);
  *reinterpret_cast<uint4*>(dst + op) = *reinterpret_cast<const uint4*>(src + ip);
}

__global__ void unknown_os8_kernel(const __nv_bfloat16* __restrict__ x8, int n, int h, int w,
                                   float* __restrict__ unk) {
  const int oh = h / 8, ow = w / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n * oh * ow) return;
  const int ox = (int)(i % ow);
  const int oy = (int)((i / ow) % oh);
  const int img = (int)(i / ((long long)ow * oh));
  unk[i] = __bfloat162float(x8[(((long long)img * h + oy * 8) * w + ox * 8) * 8 + 4]);
}

// ---------------------------------------------------------------------------------------------
// SpectralNorm fold (GCA/ops.py:38-45): sigma = u^T (W v); packed = W/sigma in [tap][cin_pad][cout]
// ---------------------------------------------------------------------------------------------
__global__ void sn_sigma_kernel(const float* __restrict__ w, const float* __restrict__ u,
                                const float* __restrict__ v, int rows, int cols, float* __restrict__ sigma) {
  // single CTA; each warp handles rows round-robin
  __shared__ float part[32];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
  float acc = 0.f;
  for (int r = warp; r < rows; r += nw) {
    float s = 0.f;
    for (int c = lane; c < cols; c += 32) s += w[(long long)r * cols + c] * v[c];
    s = warp_sum(s);
    acc += s * u[r];
  }
  if (lane == 0) part[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float t = lane < nw ? part[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) *sigma = t;
  }
}

__global__ void pack_weight_kernel(const float* __restrict__ w, const float* __restrict__ sigma, int cout,
                                   int cin, int kh, int kw, int transposed, int cin_pad,
                                   float* __restrict__ packed) {
  const long long total = (long long)kh * kw * cin_pad * cout;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int co = (int)(i % cout);
  const int ci = (int)((i / cout) % cin_pad);
  const int t = (int)(i / ((long long)cout * cin_pad));
  float val = 0.f;
  if (ci < cin) {
    const long long src = transposed ? (((long long)ci * cout + co) * kh * kw + t)
                                     : (((long long)co * cin + ci) * kh * kw + t);
    val = w[src];
    if (sigma) val = val / *sigma;
  }
  packed[i] = val;
}

__global__ void pack_weight_tc_kernel(const float* __restrict__ packed, int taps, int cin, int cout,
                                      __nv_bfloat16* __restrict__ wtc) {
  const long long total = (long long)taps * cout * cin;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ci = (int)(i % cin);
  const int co = (int)((i / cin) % cout);
  const int t = (int)(i / ((long long)cin * cout));
  store1(wtc + i, total, packed[((long long)t * cin + ci) * cout + co]);
}

__global__ void pack_weight_fold_kernel(const float* __restrict__ packed, int cout, __nv_bfloat16* __restrict__ wf) {
  const int total = 3 * cout * 32;
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int k = i % 32, co = (i / 32) % cout, dy = i / (32 * cout);
  const int dxi = k / 8, c = k % 8;
  const float v = dxi < 3 ? packed[((long long)(dy * 3 + dxi) * 8 + c) * cout + co] : 0.f;
  store1(wf + i, total, v);
}

__global__ void bn_fold_kernel(const float* g, const float* b, const float* m, const float* v, float eps, int c,
                               float* scale, float* shift) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= c) return;
  const float s = g[i] / sqrtf(v[i] + eps);
  scale[i] = s;
  shift[i] = b[i] - m[i] * s;
}

// ---------------------------------------------------------------------------------------------
__global__ void nchw_to_split_kernel(const float* __restrict__ x, int n, int c, int h, int w, int c_pad,
                                     __nv_bfloat16* __restrict__ y, long long plane) {
  const long long hw = (long long)h * w;
  const long long total = (long long)n * hw * c_pad;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cc = (int)(i % c_pad);
  const long long p = (i / c_pad) % hw;
  const long long img = i / (c_pad * hw);
  const float v = cc < c ? x[(img * c + cc) * hw + p] : 0.f;
  store1(y + i, plane, v);
}
__global__ void split_to_nchw_kernel(const __nv_bfloat16* __restrict__ x, int n, int c, int h, int w, int c_pad,
                                     long long plane, float* __restrict__ y) {
  const long long hw = (long long)h * w;
  const long long total = (long long)n * c * hw;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const long long p = i % hw;
  const int cc = (int)((i / hw) % c);
  const long long img = i / (hw * c);
  y[i] = load1(x + (img * hw + p) * c_pad + cc, plane);
}

static inline unsigned blocks(long long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

}  // namespace tcv

using namespace tcv;

extern "C" {

}  // extern "C"

template <typename T>
static int preprocess_eval_t(const T* imgs, const T* tris, int frames, int h, int w, int dilate, void* x8,
                             float* trimask, uint8_t* tmp, tcv_stream_t stream) {
  TCV_REQUIRE(imgs && tris && x8 && trimask && tmp, "preprocess_eval: null pointer");
  TCV_REQUIRE(frames > 0 && h > 0 && w > 0, "preprocess_eval: bad dims");
  const long long total = (long long)frames * h * w;
  // tmp holds two byte planes: [0,total) mask, [total, 2*total) scratch for the separable dilation
  uint8_t* m0 = tmp;
  uint8_t* m1 = tmp + total;
  trimask_raw_kernel<T><<<blocks(total), 256, 0, S(stream)>>>(tris, total, m0);
  int rc = launched("trimask_raw_kernel");
  if (rc) return rc;
  if (dilate > 0) {
    dilate_row_kernel<<<blocks(total), 256, 0, S(stream)>>>(m0, frames, h, w, dilate, m1);
    if ((rc = launched("dilate_row_kernel"))) return rc;
    dilate_col_kernel<<<blocks(total), 256, 0, S(stream)>>>(m1, frames, h, w, dilate, m0);
    if ((rc = launched("dilate_col_kernel"))) return rc;
  }
  preprocess_kernel<T><<<blocks(total), 256, 0, S(stream)>>>(imgs, tris, m0, frames, h, w,
                                                             reinterpret_cast<__nv_bfloat16*>(x8), trimask);
  return launched("preprocess_kernel");
}

extern "C" {

int tcv_preprocess_eval(const float* imgs, const float* tris, int frames, int h, int w, int dilate, void* x8,
                        float* trimask, uint8_t* tmp, tcv_stream_t stream) {
  return preprocess_eval_t<float>(imgs, tris, frames, h, w, dilate, x8, trimask, tmp, stream);
}

int tcv_preprocess_eval_u8(const uint8_t* imgs, const uint8_t* tris, int frames, int h, int w, int dilate, void* x8,
                           float* trimask, uint8_t* tmp, tcv_stream_t stream) {
  return preprocess_eval_t<uint8_t>(imgs, tris, frames, h, w, dilate, x8, trimask, tmp, stream);
}

int tcv_postprocess_eval(const float* pred, const float* tris, const float* trimask, int batch, int frames,
                         int h, int w, float* alphas, tcv_stream_t stream) {
  TCV_REQUIRE(pred && tris && trimask && alphas, "postprocess_eval: null pointer");
  TCV_REQUIRE(frames >= 3, "postprocess_eval: need at least 3 frames");
  const long long total = (long long)batch * frames * h * w;
  postprocess_kernel<float><<<blocks(total), 256, 0, S(stream)>>>(pred, tris, trimask, batch, frames, h, w, alphas);
  return launched("postprocess_kernel");
}

int tcv_postprocess_eval_u8(const float* pred, const uint8_t* tris, const float* trimask, int batch, int frames,
                            int h, int w, float* alphas, tcv_stream_t stream) {
  TCV_REQUIRE(pred && tris && trimask && alphas, "postprocess_eval_u8: null pointer");
  TCV_REQUIRE(frames >= 3, "postprocess_eval_u8: need at least 3 frames");
  const long long total = (long long)batch * frames * h * w;
  postprocess_kernel<uint8_t><<<blocks(total), 256, 0, S(stream)>>>(pred, tris, trimask, batch, frames, h, w, alphas);
  return launched("postprocess_kernel");
}

int tcv_avgpool2(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y, "avgpool2: null pointer");
  TCV_REQUIRE(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "avgpool2: h,w must be even and c%%8==0");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  avgpool2_kernel<<<blocks(total), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c,
                                                        reinterpret_cast<__nv_bfloat16*>(y));
  return launched("avgpool2_kernel");
}

int tcv_pad_reflect1(const void* x, int n, int h, int w, int c, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y, "pad_reflect1: null pointer");
  TCV_REQUIRE(h >= 2 && w >= 2 && c % 8 == 0, "pad_reflect1: need h,w >= 2 and c %% 8 == 0");
  const long long total = (long long)n * (h + 2) * (w + 2) * (c / 8);
  pad_reflect1_kernel<<<blocks(total), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c,
                                                            reinterpret_cast<__nv_bfloat16*>(y));
  return launched("pad_reflect1_kernel");
}

int tcv_unknown_os8(const void* x8, int n, int h, int w, float* unknown, tcv_stream_t stream) {
  TCV_REQUIRE(x8 && unknown, "unknown_os8: null pointer");
  TCV_REQUIRE(h % 8 == 0 && w % 8 == 0, "unknown_os8: h,w must be multiples of 8");
  const long long total = (long long)n * (h / 8) * (w / 8);
  unknown_os8_kernel<<<blocks(total), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x8), n, h, w,
                                                           unknown);
  return launched("unknown_os8_kernel");
}

int tcv_sn_fold_pack(const float* w_bar, const float* u, const float* v, int cout, int cin, int kh, int kw,
                     int transposed, int cin_pad, float* packed, float* sigma_out, tcv_stream_t stream) {
  TCV_REQUIRE(w_bar && packed, "sn_fold_pack: null pointer");
  TCV_REQUIRE(cin_pad >= cin, "sn_fold_pack: cin_pad < cin");
  TCV_REQUIRE((u == nullptr) == (v == nullptr), "sn_fold_pack: u and v must both be given or both null");
  TCV_REQUIRE(!u || sigma_out, "sn_fold_pack: sigma_out workspace required with spectral norm");
  if (u) {
    const int rows = transposed ? cin : cout;
    const int cols = (transposed ? cout : cin) * kh * kw;
    sn_sigma_kernel<<<1, 1024, 0, S(stream)>>>(w_bar, u, v, rows, cols, sigma_out);
    int rc = launched("sn_sigma_kernel");
    if (rc) return rc;
  }
  const long long total = (long long)kh * kw * cin_pad * cout;
  pack_weight_kernel<<<blocks(total), 256, 0, S(stream)>>>(w_bar, u ? sigma_out : nullptr, cout, cin, kh, kw,
                                                           transposed, cin_pad, packed);
  return launched("pack_weight_kernel");
}

int tcv_pack_weight_tc(const float* packed, int taps, int cin, int cout, void* w_tc, tcv_stream_t stream) {
  TCV_REQUIRE(packed && w_tc && taps > 0 && cin > 0 && cout > 0, "pack_weight_tc: bad arguments");
  const long long total = (long long)taps * cin * cout;
  pack_weight_tc_kernel<<<blocks(total), 256, 0, S(stream)>>>(packed, taps, cin, cout,
                                                              reinterpret_cast<__nv_bfloat16*>(w_tc));
  return launched("pack_weight_tc_kernel");
}

int tcv_pack_weight_fold(const float* packed, int cout, void* w_fold, tcv_stream_t stream) {
  TCV_REQUIRE(packed && w_fold && cout > 0, "pack_weight_fold: bad arguments");
  pack_weight_fold_kernel<<<blocks(3 * cout * 32), 256, 0, S(stream)>>>(packed, cout,
                                                                        reinterpret_cast<__nv_bfloat16*>(w_fold));
  return launched("pack_weight_fold_kernel");
}

int tcv_bn_fold(const float* gamma, const float* beta, const float* mean, const float* var, float eps, int c,
                float* scale, float* shift, tcv_stream_t stream) {
  TCV_REQUIRE(gamma && beta && mean && var && scale && shift, "bn_fold: null pointer");
  bn_fold_kernel<<<blocks(c), 256, 0, S(stream)>>>(gamma, beta, mean, var, eps, c, scale, shift);
  return launched("bn_fold_kernel");
}

int tcv_nchw_to_split(const float* x, int n, int c, int h, int w, int c_pad, void* y, long long y_plane,
                      tcv_stream_t stream) {
  TCV_REQUIRE(x && y && c_pad >= c, "nchw_to_split: bad arguments");
  const long long total = (long long)n * h * w * c_pad;
  nchw_to_split_kernel<<<blocks(total), 256, 0, S(stream)>>>(x, n, c, h, w, c_pad,
                                                             reinterpret_cast<__nv_bfloat16*>(y),
                                                             y_plane ? y_plane : total);
  return launched("nchw_to_split_kernel");
}

int tcv_split_to_nchw(const void* x, int n, int c, int h, int w, int c_pad, long long x_plane, float* y,
                      tcv_stream_t stream) {
  TCV_REQUIRE(x && y && c_pad >= c, "split_to_nchw: bad arguments");
  const long long total = (long long)n * c * h * w;
  split_to_nchw_kernel<<<blocks(total), 256, 0, S(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), n, c, h, w, c_pad,
      x_plane ? x_plane : (long long)n * h * w * c_pad, y);
  return launched("split_to_nchw_kernel");
}

}  // extern "C"

