// Temporal Attention Module core (models/VMN/VMN_model.py:27-68): for every unknown-region pixel,
// 7x7 local cross-frame attention of the centre-frame query against the previous/next frame
// keys, softmax over the window, aggregation of the *keys*, plus the value branch.
//
// The reference materialises F.unfold(k) = [C,49,N] fp32 (819 MB per neighbour at 1080p) and
// gathers unknown pixels with torch.nonzero (host sync).  Here one warp owns one pixel: lanes
// split the channel axis, the 49 logits live in registers (lane j holds logit j and j+32), and
// both neighbours are handled in the same pass.  Out-of-image window positions contribute a
// zero key: logit 0, still part of the softmax (zero padding of F.unfold, VMN_model.py:35-36).
#include "common.cuh"

namespace tcv {

template <int CPL>  // channels per lane (C = 32*CPL)
__global__ void __launch_bounds__(256) tam_attend_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ v, const __nv_bfloat16* __restrict__ kb,
    const __nv_bfloat16* __restrict__ kf, const float* __restrict__ mask, long long mask_stride, int mh, int mw,
    int batch, int h, int w, int window, __nv_bfloat16* __restrict__ out, float* __restrict__ attb,
    float* __restrict__ attf, uint8_t* __restrict__ small_mask) {
  constexpr int C = 32 * CPL;
  const int N = h * w;
  const long long gw = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (gw >= (long long)batch * N) return;
  const int b = (int)(gw / N), pix = (int)(gw - (long long)b * N);
  const int y = pix / w, x = pix - y * w;
  const long long plane = (long long)batch * N * C;
  const long long base = ((long long)b * N + pix) * C + lane * CPL;
  const int w2 = window * window, r = window / 2;

  // nearest-neighbour mask sample: src = floor(dst * in/out)  (F.interpolate 'nearest', VMN_model.py:22)
  const int my = (int)(((long long)y * mh) / h), mx = (int)(((long long)x * mw) / w);
  const bool m = mask[(long long)b * mask_stride + (long long)my * mw + mx] != 0.f;
  if (lane == 0) small_mask[(long long)b * N + pix] = m ? 1 : 0;

  float o[CPL];
  if (CPL == 1) o[0] = load1(v + base, plane); else if (CPL == 4) load4(v + base, plane, o); else load8(v + base, plane, o);

  if (!m) {
    for (int j = lane; j < w2; j += 32) {
      attb[((long long)b * w2 + j) * N + pix] = 0.f;
      attf[((long long)b * w2 + j) * N + pix] = 0.f;
    }
  } else {
    float qv[CPL];
    if (CPL == 1) qv[0] = load1(q + base, plane); else if (CPL == 4) load4(q + base, plane, qv); else load8(q + base, plane, qv);
    const float inv_sqrt_c = 1.0f / sqrtf((float)C);
#pragma unroll 1
    for (int nb = 0; nb < 2; ++nb) {
      const __nv_bfloat16* k = nb == 0 ? kb : kf;
      float* att = nb == 0 ? attb : attf;
      float l0 = -INFINITY, l1 = -INFINITY;  // logits j = lane and j = lane + 32
      for (int j = 0; j < w2; ++j) {
        const int yy = y + j / window - r, xx = x + j % window - r;
        float d = 0.f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          float kv[CPL];
          const __nv_bfloat16* kp = k + ((long long)b * N + (long long)yy * w + xx) * C + lane * CPL;
          if (CPL == 1) kv[0] = load1(kp, plane); else if (CPL == 4) load4(kp, plane, kv); else load8(kp, plane, kv);
#pragma unroll
          for (int c = 0; c < CPL; ++c) d = fmaf(qv[c], kv[c], d);
          d = warp_sum(d);
        }
        d *= inv_sqrt_c;
        if ((j & 31) == lane) {
          if (j < 32) l0 = d; else l1 = d;
        }
      }
      for (int j = lane; j < w2; j += 32) att[((long long)b * w2 + j) * N + pix] = j < 32 ? l0 : l1;
      const float mxv = warp_max(fmaxf(l0, l1));
      const float e0 = lane < w2 ? expf(l0 - mxv) : 0.f;
      const float e1 = lane + 32 < w2 ? expf(l1 - mxv) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      for (int j = 0; j < w2; ++j) {
        const int yy = y + j / window - r, xx = x + j % window - r;
        const float a = __shfl_sync(0xffffffffu, j < 32 ? e0 : e1, j & 31) * inv;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          float kv[CPL];
          const __nv_bfloat16* kp = k + ((long long)b * N + (long long)yy * w + xx) * C + lane * CPL;
          if (CPL == 1) kv[0] = load1(kp, plane); else if (CPL == 4) load4(kp, plane, kv); else load8(kp, plane, kv);
#pragma unroll
          for (int c = 0; c < CPL; ++c) o[c] = fmaf(a, kv[c], o[c]);
        }
      }
    }
  }
  if (CPL == 1) store1(out + base, plane, o[0]); else if (CPL == 4) store4(out + base, plane, o); else store8(out + base, plane, o);
}

}  // namespace tcv

using namespace tcv;

extern "C" int tcv_tam_attend(const void* q, const void* v, const void* kb, const void* kf, const float* mask,
                              long long mask_stride, int mh, int mw, int batch, int h, int w, int c, int window,
                              void* out, float* attb, float* attf, uint8_t* small_mask, tcv_stream_t stream) {
  TCV_REQUIRE(q && v && kb && kf && mask && out && attb && attf && small_mask, "tam_attend: null pointer");
  TCV_REQUIRE(window >= 1 && window % 2 == 1 && window * window <= 64, "tam_attend: window must be odd and <= 7");
  TCV_REQUIRE(c == 32 || c == 128 || c == 256, "tam_attend: channels must be 32, 128 or 256 (the TAM widths of the reference's four base networks)");
  const long long warps = (long long)batch * h * w;
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  auto Q = reinterpret_cast<const __nv_bfloat16*>(q);
  auto V = reinterpret_cast<const __nv_bfloat16*>(v);
  auto KB = reinterpret_cast<const __nv_bfloat16*>(kb);
  auto KF = reinterpret_cast<const __nv_bfloat16*>(kf);
  auto O = reinterpret_cast<__nv_bfloat16*>(out);
  if (c == 32)
    tam_attend_kernel<1><<<grid, 256, 0, S(stream)>>>(Q, V, KB, KF, mask, mask_stride, mh, mw, batch, h, w,
                                                      window, O, attb, attf, small_mask);
  else if (c == 128)
    tam_attend_kernel<4><<<grid, 256, 0, S(stream)>>>(Q, V, KB, KF, mask, mask_stride, mh, mw, batch, h, w,
                                                      window, O, attb, attf, small_mask);
  else
    tam_attend_kernel<8><<<grid, 256, 0, S(stream)>>>(Q, V, KB, KF, mask, mask_stride, mh, mw, batch, h, w,
                                                      window, O, attb, attf, small_mask);
  return launched("tam_attend_kernel");
}
