// Weight gradient of the gather-form convolution (autograd of nn.Conv2d / nn.ConvTranspose2d w.r.t. weight):
//   dw[wtap[t]][ci][co] += sum_{n, gy, gx} x[n, gy*stride+dy[t], gx*stride+dx[t], ci] * dz[n, oy, ox, co]
// i.e. per tap a GEMM with M = cin, N = cout and the reduction over all output pixels (K up to millions).
// CUDA-core fp32 kernel (first training-path version): CTA tile TMxTN of the (ci, co) plane for one tap and one
// slice of the pixel range (split-K), pixel chunks of 16 staged through shared memory as fp32 (hi+lo summed at
// load time, 16-byte NHWC vector loads), partial results added to dw with fp32 atomics.
#include "common.cuh"

namespace tcv {

constexpr int WG_BK = 16;

template <int TM, int TN>   // CTA tile (TM = TN = 64: 4x4 per thread; 32: 2x2 per thread), 256 threads
__global__ void __launch_bounds__(256) conv_wgrad_kernel(const tcv_conv_desc d, const __nv_bfloat16* __restrict__ dz,
                                                         long long dz_plane, int dz_c, float* __restrict__ dw,
                                                         long long px_per_slice) {
  constexpr int RM = TM / 16, RN = TN / 16;
  __shared__ float As[WG_BK][TM + 4];
  __shared__ float Bs[WG_BK][TN + 4];
  const int tiles_n = (dz_c + TN - 1) / TN;
  const int ci0 = (blockIdx.x / tiles_n) * TM, co0 = (blockIdx.x % tiles_n) * TN;
  const int t = blockIdx.y;
  const long long total_px = (long long)d.n * d.gh * d.gw;
  const long long k_begin = (long long)blockIdx.z * px_per_slice;
  const long long k_end = min(k_begin + px_per_slice, total_px);
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(d.x);
  const int tdy = d.dy[t], tdx = d.dx[t];

  float acc[RM][RN];
#pragma unroll
  for (int i = 0; i < RM; ++i)
#pragma unroll
    for (int j = 0; j < RN; ++j) acc[i][j] = 0.f;

  // loader mapping: 16 pixels x (TM/4) float4 channel groups; with TM = 64 that is 256 slots (one per thread),
  // with TM = 32 half of the threads load A and all load... keep it simple: loop over slots
  for (long long k0 = k_begin; k0 < k_end; k0 += WG_BK) {
    __syncthreads();
    for (int s = threadIdx.x; s < WG_BK * (TM / 4); s += 256) {
      const int kk = s / (TM / 4), c4 = (s % (TM / 4)) * 4;
      const long long k = k0 + kk;
      float f[4] = {0, 0, 0, 0};
      if (k < k_end && ci0 + c4 < d.cin) {
        const int gx = (int)(k % d.gw);
        const int gy = (int)((k / d.gw) % d.gh);
        const int img = (int)(k / ((long long)d.gw * d.gh));
        int iy = gy * d.stride + tdy, ix = gx * d.stride + tdx;
        bool ok = true;
        if (d.pad_mode == TCV_PAD_REFLECT) { iy = reflect(iy, d.ih); ix = reflect(ix, d.iw); }
        else ok = iy >= 0 && iy < d.ih && ix >= 0 && ix < d.iw;
        if (ok) load4(xin + (long long)img * d.x_img_stride + ((long long)iy * d.iw + ix) * d.cin + ci0 + c4, d.x_plane, f);
      }
      As[kk][c4 + 0] = f[0]; As[kk][c4 + 1] = f[1]; As[kk][c4 + 2] = f[2]; As[kk][c4 + 3] = f[3];
    }
    for (int s = threadIdx.x; s < WG_BK * (TN / 4); s += 256) {
      const int kk = s / (TN / 4), c4 = (s % (TN / 4)) * 4;
      const long long k = k0 + kk;
      float f[4] = {0, 0, 0, 0};
      if (k < k_end && co0 + c4 < dz_c) {
        const int gx = (int)(k % d.gw);
        const int gy = (int)((k / d.gw) % d.gh);
        const int img = (int)(k / ((long long)d.gw * d.gh));
        const int oy = gy * d.oy_mul + d.oy_off, ox = gx * d.ox_mul + d.ox_off;
        load4(dz + (((long long)img * d.oh + oy) * d.ow + ox) * dz_c + co0 + c4, dz_plane, f);
      }
      Bs[kk][c4 + 0] = f[0]; Bs[kk][c4 + 1] = f[1]; Bs[kk][c4 + 2] = f[2]; Bs[kk][c4 + 3] = f[3];
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < WG_BK; ++kk) {
      float a[RM], b[RN];
#pragma unroll
      for (int i = 0; i < RM; ++i) a[i] = As[kk][ty * RM + i];
#pragma unroll
      for (int j = 0; j < RN; ++j) b[j] = Bs[kk][tx * RN + j];
#pragma unroll
      for (int i = 0; i < RM; ++i)
#pragma unroll
        for (int j = 0; j < RN; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
  float* out = dw + (long long)d.wtap[t] * d.cin * dz_c;
#pragma unroll
  for (int i = 0; i < RM; ++i) {
    const int ci = ci0 + ty * RM + i;
    if (ci >= d.cin) continue;
#pragma unroll
    for (int j = 0; j < RN; ++j) {
      const int co = co0 + tx * RN + j;
      if (co < dz_c && acc[i][j] != 0.f) atomicAdd(out + (long long)ci * dz_c + co, acc[i][j]);
    }
  }
}

}  // namespace tcv

using namespace tcv;

extern "C" int tcv_conv2d_wgrad(const tcv_conv_desc* dp, const void* dz, long long dz_plane, int dz_c, float* dw,
                                tcv_stream_t stream) {
  TCV_REQUIRE(dp && dz && dw, "conv2d_wgrad: null pointer");
  tcv_conv_desc d = *dp;
  TCV_REQUIRE(d.x && d.n > 0 && d.gh > 0 && d.gw > 0, "conv2d_wgrad: bad descriptor");
  TCV_REQUIRE(d.cin % 8 == 0 && dz_c % 8 == 0 && dz_c >= d.cout, "conv2d_wgrad: cin and dz_c must be multiples of 8");
  TCV_REQUIRE(d.ntaps >= 1 && d.ntaps <= TCV_MAX_TAPS, "conv2d_wgrad: ntaps out of range");
  if (d.x_plane == 0) d.x_plane = (long long)d.n * d.ih * d.iw * d.cin;
  if (d.x_img_stride == 0) d.x_img_stride = (long long)d.ih * d.iw * d.cin;
  if (dz_plane == 0) dz_plane = (long long)d.n * d.oh * d.ow * dz_c;
  const long long total_px = (long long)d.n * d.gh * d.gw;
  const bool small = d.cin <= 32 && dz_c <= 32;
  const int T = small ? 32 : 64;
  const int tiles = ((d.cin + T - 1) / T) * ((dz_c + T - 1) / T);
  // split-K: aim at >= ~4 CTAs per SM overall, at least 256 pixels per slice
  long long want = (148LL * 4 + (long long)tiles * d.ntaps - 1) / ((long long)tiles * d.ntaps);
  long long max_slices = (total_px + 255) / 256;
  long long slices = want < 1 ? 1 : want;
  if (slices > max_slices) slices = max_slices;
  if (slices > 65535) slices = 65535;
  long long per = (total_px + slices - 1) / slices;
  per = (per + WG_BK - 1) / WG_BK * WG_BK;
  slices = (total_px + per - 1) / per;
  dim3 grid((unsigned)tiles, (unsigned)d.ntaps, (unsigned)slices);
  auto DZ = reinterpret_cast<const __nv_bfloat16*>(dz);
  if (small) conv_wgrad_kernel<32, 32><<<grid, 256, 0, S(stream)>>>(d, DZ, dz_plane, dz_c, dw, per);
  else conv_wgrad_kernel<64, 64><<<grid, 256, 0, S(stream)>>>(d, DZ, dz_plane, dz_c, dw, per);
  return launched("conv_wgrad_kernel");
}
