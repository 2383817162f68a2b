// CUDA-core gather-form convolution with fused epilogue (fp32 accumulate on split-bf16 NHWC).
//
// This is the general-shape path: small channel counts (3/6/16 -> N), the Cout=1 alpha head,
// reflect padding, stride-2 and the 4 phases of the 4x4/stride-2 transposed conv.  These
// layers are HBM-bound, so the kernel is organised around coalesced
// NHWC access: one thread owns one output pixel and COT consecutive output channels, reads
// its input pixels as 16-byte vectors (8 channels per plane) and receives the weights as
// warp-wide shared-memory broadcasts.
//
// Replaces nn.Conv2d / nn.ConvTranspose2d + BatchNorm2d + activation + residual call sites,
// see include/tcvom_b200.h (tcv_conv2d).
#include "common.cuh"

namespace tcv {

constexpr int CIC = 32;  // input channels staged per weight chunk

// PX output pixels (consecutive in x) per thread (PX = 2 halves the shared-memory weight traffic per FMA but
// measured slower than PX = 1 on B200, see launch_conv).
template <int COT, int PX>
__global__ void __launch_bounds__(256) conv_direct_kernel(const tcv_conv_desc d) {
  extern __shared__ float wsm[];  // [ntaps][CIC][COT]
  const int cic = d.cin < CIC ? d.cin : CIC;
  const int gwp = d.gw / PX;                       // pixel groups per row (gw % PX == 0)
  const int grp = blockIdx.x * 256 + threadIdx.x;
  const int co0 = blockIdx.y * COT;
  const int n = blockIdx.z;
  const bool valid = grp < d.gh * gwp;
  const int gy = valid ? grp / gwp : 0;
  const int gx0 = valid ? (grp - gy * gwp) * PX : 0;
  const long long xplane = d.x_plane;
  const __nv_bfloat16* xin = reinterpret_cast<const __nv_bfloat16*>(d.x) + (long long)n * d.x_img_stride;

  float acc[PX][COT];
#pragma unroll
  for (int q = 0; q < PX; ++q)
#pragma unroll
    for (int j = 0; j < COT; ++j) acc[q][j] = 0.f;

  for (int ci0 = 0; ci0 < d.cin; ci0 += cic) {
    __syncthreads();
    const int chunk = d.ntaps * cic * COT;
    for (int i = threadIdx.x; i < chunk; i += 256) {
      const int j = i % COT;
      const int c = (i / COT) % cic;
      const int t = i / (COT * cic);
      wsm[i] = d.w[((long long)d.wtap[t] * d.cin + ci0 + c) * d.cout + co0 + j];
    }
    __syncthreads();
    if (!valid) continue;
    for (int t = 0; t < d.ntaps; ++t) {
      int iy = gy * d.stride + d.dy[t];
      bool rowok = true;
      if (d.pad_mode == TCV_PAD_REFLECT) iy = reflect(iy, d.ih);
      else rowok = iy >= 0 && iy < d.ih;
      if (!rowok) continue;
      const __nv_bfloat16* prow = xin + (long long)iy * d.iw * d.cin + ci0;
      const float* wt = wsm + t * cic * COT;
      int ixs[PX];
      bool ok[PX];
#pragma unroll
      for (int q = 0; q < PX; ++q) {
        int ix = (gx0 + q) * d.stride + d.dx[t];
        if (d.pad_mode == TCV_PAD_REFLECT) { ix = reflect(ix, d.iw); ok[q] = true; }
        else { ok[q] = ix >= 0 && ix < d.iw; ix = ok[q] ? ix : 0; }
        ixs[q] = ix;
      }
      for (int c8 = 0; c8 < cic; c8 += 8) {
        float f[PX][8];
#pragma unroll
        for (int q = 0; q < PX; ++q) {
          if (ok[q]) load8(prow + (long long)ixs[q] * d.cin + c8, xplane, f[q]);
          else {
#pragma unroll
            for (int k = 0; k < 8; ++k) f[q][k] = 0.f;
          }
        }
        if constexpr (COT == 1) {
          // single output channel (the alpha head): the 8 weights of this channel group are two float4
          // broadcasts instead of eight scalar shared-memory loads
          const float4 wa = *reinterpret_cast<const float4*>(wt + c8), wb = *reinterpret_cast<const float4*>(wt + c8 + 4);
          const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
#pragma unroll
          for (int q = 0; q < PX; ++q)
#pragma unroll
            for (int k = 0; k < 8; ++k) acc[q][0] = fmaf(f[q][k], wv[k], acc[q][0]);
        } else {
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          const float* wp = wt + (c8 + k) * COT;
          if (COT >= 4) {
#pragma unroll
            for (int j = 0; j < COT / 4; ++j) {
              const float4 w4 = reinterpret_cast<const float4*>(wp)[j];
#pragma unroll
              for (int q = 0; q < PX; ++q) {
                const float xv = f[q][k];
                acc[q][4 * j + 0] = fmaf(xv, w4.x, acc[q][4 * j + 0]);
                acc[q][4 * j + 1] = fmaf(xv, w4.y, acc[q][4 * j + 1]);
                acc[q][4 * j + 2] = fmaf(xv, w4.z, acc[q][4 * j + 2]);
                acc[q][4 * j + 3] = fmaf(xv, w4.w, acc[q][4 * j + 3]);
              }
            }
          } else {
#pragma unroll
            for (int j = 0; j < COT; ++j)
#pragma unroll
              for (int q = 0; q < PX; ++q) acc[q][j] = fmaf(f[q][k], wp[j], acc[q][j]);
          }
        }
        }
      }
    }
  }
  if (!valid) return;

  // ---- fused epilogue
  const long long oplane = (long long)d.n * d.oh * d.ow * d.cout;
#pragma unroll
  for (int q = 0; q < PX; ++q) {
    float* a = acc[q];
    const int oy = gy * d.oy_mul + d.oy_off;
    const int ox = (gx0 + q) * d.ox_mul + d.ox_off;
    const long long obase = (((long long)n * d.oh + oy) * d.ow + ox) * d.cout + co0;
    if (d.s1) {
#pragma unroll
      for (int j = 0; j < COT; ++j) a[j] = a[j] * d.s1[co0 + j];
    }
    if (d.b1) {
#pragma unroll
      for (int j = 0; j < COT; ++j) a[j] += d.b1[co0 + j];
    }
    if (d.res1) {
      const int rh = d.oh >> d.res1_shift, rw = d.ow >> d.res1_shift;
      const long long rplane = d.res1_plane;
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(d.res1) +
                               (((long long)n * rh + (oy >> d.res1_shift)) * rw + (ox >> d.res1_shift)) * d.cout + co0;
      if (COT >= 8) {
#pragma unroll
        for (int j = 0; j < COT; j += 8) {
          float f[8];
          load8(r + j, rplane, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) a[j + k] += f[k];
        }
      } else {
#pragma unroll
        for (int j = 0; j < COT; ++j) a[j] += load1(r + j, rplane);
      }
    }
    apply_act_n<COT>(a, d.act);
    if (d.s2) {
#pragma unroll
      for (int j = 0; j < COT; ++j) a[j] = a[j] * d.s2[co0 + j] + d.b2[co0 + j];
    }
    if (d.res2) {
      const __nv_bfloat16* r = reinterpret_cast<const __nv_bfloat16*>(d.res2) + obase;
      if (COT >= 8) {
#pragma unroll
        for (int j = 0; j < COT; j += 8) {
          float f[8];
          load8(r + j, d.res2_plane, f);
#pragma unroll
          for (int k = 0; k < 8; ++k) a[j + k] += f[k];
        }
      } else {
#pragma unroll
        for (int j = 0; j < COT; ++j) a[j] += load1(r + j, d.res2_plane);
      }
    }
    if (d.y) {
      __nv_bfloat16* y = reinterpret_cast<__nv_bfloat16*>(d.y) + obase;
      if (COT >= 8) {
#pragma unroll
        for (int j = 0; j < COT; j += 8) store8(y + j, oplane, a + j);
      } else {
#pragma unroll
        for (int j = 0; j < COT; ++j) store1(y + j, oplane, a[j]);
      }
    }
    if (d.y_f32) {
      float* y = d.y_f32 + obase;
#pragma unroll
      for (int j = 0; j < COT; ++j) y[j] = a[j];
    }
  }
}

template <int COT>
static int launch_conv(const tcv_conv_desc& d, cudaStream_t st) {
  const int cic = d.cin < CIC ? d.cin : CIC;
  const size_t smem = (size_t)d.ntaps * cic * COT * sizeof(float);
  // PX = 2 measured SLOWER on B200 (3.3 vs 2.6 ms per 1080p window over the six direct layers): disabled
  if (false && d.gw % 2 == 0) {
    dim3 grid((d.gh * (d.gw / 2) + 255) / 256, d.cout / COT, d.n);
    conv_direct_kernel<COT, 2><<<grid, 256, smem, st>>>(d);
  } else {
    dim3 grid((d.gh * d.gw + 255) / 256, d.cout / COT, d.n);
    conv_direct_kernel<COT, 1><<<grid, 256, smem, st>>>(d);
  }
  return launched("conv_direct_kernel");
}

int conv2d_direct(const tcv_conv_desc& d, cudaStream_t st) {
  if (d.cout % 32 == 0) return launch_conv<32>(d, st);
  if (d.cout % 16 == 0) return launch_conv<16>(d, st);
  if (d.cout % 8 == 0) return launch_conv<8>(d, st);
  return launch_conv<1>(d, st);
}

}  // namespace tcv
