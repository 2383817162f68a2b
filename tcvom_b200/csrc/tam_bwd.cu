// Backward of the Temporal Attention Module core (autograd of models/VMN/VMN_model.py:27-68).
//
// Forward (tam.cu) for an unknown-region pixel n and neighbour t in {prev, next}:
//   logit_j = <q_n, k_t[n+j]>/sqrt(C) ; att = softmax_j(logit) ; out_n = v_n + sum_t sum_j att_j k_t[n+j]
//   returned logits: m * logit (consumed by L_af)
// Backward, one warp per unknown pixel (lanes split the channel axis, the 49 logits / gradients live in
// registers: lane j holds entries j and j+32):
//   datt_j = <dout_n, k_j> ; dlogit_j = att_j (datt_j - sum_i att_i datt_i) + dlogit_ext_j
//   dq_n   = sum_t sum_j dlogit_j k_j / sqrt(C)
//   dk_t[n+j] += att_j dout_n + dlogit_j q_n / sqrt(C)        (fp32 atomics: windows of neighbouring pixels overlap)
//   dv = dout (aliased by the host)
#include "common.cuh"

namespace tcv {

template <int CPL>
__global__ void __launch_bounds__(256) tam_attend_bwd_kernel(
    const __nv_bfloat16* __restrict__ q, const __nv_bfloat16* __restrict__ kb, const __nv_bfloat16* __restrict__ kf,
    const float* __restrict__ mask, long long mask_stride, int mh, int mw, int batch, int h, int w, int window,
    const __nv_bfloat16* __restrict__ dout, const float* __restrict__ dattb, const float* __restrict__ dattf,
    __nv_bfloat16* __restrict__ dq, float* __restrict__ dkb, float* __restrict__ dkf) {
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
  const int my = (int)(((long long)y * mh) / h), mx = (int)(((long long)x * mw) / w);
  const bool m = mask[(long long)b * mask_stride + (long long)my * mw + mx] != 0.f;

  float dqv[CPL];
#pragma unroll
  for (int c = 0; c < CPL; ++c) dqv[c] = 0.f;
  if (m) {
    float qv[CPL], go[CPL];
    if (CPL == 1) { qv[0] = load1(q + base, plane); go[0] = load1(dout + base, plane); }
    else if (CPL == 4) { load4(q + base, plane, qv); load4(dout + base, plane, go); }
    else { load8(q + base, plane, qv); load8(dout + base, plane, go); }
    const float inv_sqrt_c = 1.0f / sqrtf((float)C);
#pragma unroll 1
    for (int nb = 0; nb < 2; ++nb) {
      const __nv_bfloat16* k = nb == 0 ? kb : kf;
      float* dk = nb == 0 ? dkb : dkf;
      const float* dext = nb == 0 ? dattb : dattf;
      float l0 = -INFINITY, l1 = -INFINITY, g0 = 0.f, g1 = 0.f;   // logits and datt for j = lane, lane + 32
      for (int j = 0; j < w2; ++j) {
        const int yy = y + j / window - r, xx = x + j % window - r;
        float d = 0.f, g = 0.f;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          float kv[CPL];
          const __nv_bfloat16* kp = k + ((long long)b * N + (long long)yy * w + xx) * C + lane * CPL;
          if (CPL == 1) kv[0] = load1(kp, plane); else if (CPL == 4) load4(kp, plane, kv); else load8(kp, plane, kv);
#pragma unroll
          for (int c = 0; c < CPL; ++c) { d = fmaf(qv[c], kv[c], d); g = fmaf(go[c], kv[c], g); }
          d = warp_sum(d);
          g = warp_sum(g);
        }
        d *= inv_sqrt_c;
        if ((j & 31) == lane) {
          if (j < 32) { l0 = d; g0 = g; } else { l1 = d; g1 = g; }
        }
      }
      const float mxv = warp_max(fmaxf(l0, l1));
      const float e0 = lane < w2 ? expf(l0 - mxv) : 0.f;
      const float e1 = lane + 32 < w2 ? expf(l1 - mxv) : 0.f;
      const float inv = 1.0f / warp_sum(e0 + e1);
      const float a0 = e0 * inv, a1 = e1 * inv;
      const float sdot = warp_sum(a0 * g0 + a1 * g1);
      float dl0 = a0 * (g0 - sdot), dl1 = a1 * (g1 - sdot);
      if (dext) {
        if (lane < w2) dl0 += dext[((long long)b * w2 + lane) * N + pix];
        if (lane + 32 < w2) dl1 += dext[((long long)b * w2 + lane + 32) * N + pix];
      }
      for (int j = 0; j < w2; ++j) {
        const int yy = y + j / window - r, xx = x + j % window - r;
        const float a = __shfl_sync(0xffffffffu, j < 32 ? a0 : a1, j & 31);
        const float dl = __shfl_sync(0xffffffffu, j < 32 ? dl0 : dl1, j & 31) * inv_sqrt_c;
        if (yy >= 0 && yy < h && xx >= 0 && xx < w) {
          float kv[CPL];
          const long long ko = ((long long)b * N + (long long)yy * w + xx) * C + lane * CPL;
          if (CPL == 1) kv[0] = load1(k + ko, plane); else if (CPL == 4) load4(k + ko, plane, kv); else load8(k + ko, plane, kv);
#pragma unroll
          for (int c = 0; c < CPL; ++c) {
            dqv[c] = fmaf(dl, kv[c], dqv[c]);
            atomicAdd(dk + ko + c, a * go[c] + dl * qv[c]);
          }
        }
      }
    }
  }
  if (CPL == 1) store1(dq + base, plane, dqv[0]); else if (CPL == 4) store4(dq + base, plane, dqv); else store8(dq + base, plane, dqv);
}

}  // namespace tcv

using namespace tcv;

extern "C" int tcv_tam_attend_bwd(const void* q, const void* kb, const void* kf, const float* mask,
                                  long long mask_stride, int mh, int mw, int batch, int h, int w, int c, int window,
                                  const void* dout, const float* dattb, const float* dattf, void* dq, float* dkb,
                                  float* dkf, tcv_stream_t stream) {
  TCV_REQUIRE(q && kb && kf && mask && dout && dq && dkb && dkf, "tam_attend_bwd: null pointer");
  TCV_REQUIRE(window >= 1 && window % 2 == 1 && window * window <= 64, "tam_attend_bwd: window must be odd and <= 7");
  TCV_REQUIRE(c == 32 || c == 128 || c == 256, "tam_attend_bwd: channels must be 32, 128 or 256");
  const long long elems = (long long)batch * h * w * c;
  TCV_CUDA(cudaMemsetAsync(dkb, 0, sizeof(float) * elems, S(stream)));
  TCV_CUDA(cudaMemsetAsync(dkf, 0, sizeof(float) * elems, S(stream)));
  const long long warps = (long long)batch * h * w;
  const unsigned grid = (unsigned)((warps * 32 + 255) / 256);
  auto Q = reinterpret_cast<const __nv_bfloat16*>(q);
  auto KB = reinterpret_cast<const __nv_bfloat16*>(kb);
  auto KF = reinterpret_cast<const __nv_bfloat16*>(kf);
  auto DO = reinterpret_cast<const __nv_bfloat16*>(dout);
  auto DQ = reinterpret_cast<__nv_bfloat16*>(dq);
  if (c == 32)
    tam_attend_bwd_kernel<1><<<grid, 256, 0, S(stream)>>>(Q, KB, KF, mask, mask_stride, mh, mw, batch, h, w, window,
                                                          DO, dattb, dattf, DQ, dkb, dkf);
  else if (c == 128)
    tam_attend_bwd_kernel<4><<<grid, 256, 0, S(stream)>>>(Q, KB, KF, mask, mask_stride, mh, mw, batch, h, w, window,
                                                          DO, dattb, dattf, DQ, dkb, dkf);
  else
    tam_attend_bwd_kernel<8><<<grid, 256, 0, S(stream)>>>(Q, KB, KF, mask, mask_stride, mh, mw, batch, h, w, window,
                                                          DO, dattb, dattf, DQ, dkb, dkf);
  return launched("tam_attend_bwd_kernel");
}
