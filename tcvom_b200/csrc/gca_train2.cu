// Training side of the shift-sum aggregation of guided contextual attention (forward: csrc/gca.cu, "Shift-sum form";
// reference: GCA/ops.py:112-118,204 and its autograd).  With m = q + a, p' = p + a on the (hh+1) x (ww+1) grid:
//     Y  = unfold_parity( A2 . F ) / 4           A2[m][p'] = sum_a A[m-a][p'-a]        F_r[p'] = feat_reflect[2p'+r-1]
//     dO2 = unfold_parity^T(dY) / 4              dA2 = dO2 . F^T      dF = A2^T . dO2
//     dA[q][p'] = sum_a dA2[q+a][p'+sh_a]        dfeat = values_parity^T(dF)
// i.e. the two backward GEMMs shrink from [P x 2048 x P] to [Pk x 512 x Pk] like the forward one (3.8x fewer FLOPs each).
// Scores, softmax and shift-add run on the padded key grid with the inference kernels (gca.cu); what is new here:
//   unfold_parity_bwd   dY split-bf16 [n,h,w,128]          -> dO2 planes [2][n][Pk][512] (x 1/4)
//   softmax_bwd_grid    A fp32 [n][P][ld], dA2 [n][Pk][ld] -> dS split-bf16 [2][n][P][ld]   (gather + row dot + softmax backward)
//   values_parity_bwd   dF fp32 [n][Pk][512]               -> dfeat split-bf16 [n,h,w,128]
#include "common.cuh"

namespace tcv {
constexpr int T2_FC = 128;

// work item = 8 channels of one (image, m, parity)
__global__ void gca_unfold_parity_bwd_kernel(const __nv_bfloat16* __restrict__ dY, int n, int h, int w,
                                             __nv_bfloat16* __restrict__ dO2) {
  const int ww1 = w / 2 + 1, Pk = (h / 2 + 1) * ww1;
  const long long total = (long long)n * Pk * 4 * (T2_FC / 8);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (T2_FC / 8)) * 8;
  long long t = i / (T2_FC / 8);
  const int r = (int)(t % 4);
  t /= 4;
  const int m = (int)(t % Pk);
  const int img = (int)(t / Pk);
  const int my = m / ww1, mx = m - my * ww1;
  const int y = 2 * my + (r >> 1) - 1, x = 2 * mx + (r & 1) - 1;
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (y >= 0 && y < h && x >= 0 && x < w) {
    load8(dY + (((long long)img * h + y) * w + x) * T2_FC + c, (long long)n * h * w * T2_FC, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= 0.25f;
  }
  store8(dO2 + ((long long)img * Pk + m) * (4 * T2_FC) + r * T2_FC + c, (long long)n * Pk * 4 * T2_FC, f);
}

// work item = 4 channels of one feature pixel: the (<= 2 x 2) positions of the reflect-padded parity grid that read it
__global__ void gca_values_parity_bwd_kernel(const float* __restrict__ dF, int n, int h, int w, __nv_bfloat16* __restrict__ dfeat) {
  const int ww1 = w / 2 + 1, Pk = (h / 2 + 1) * ww1;
  const long long total = (long long)n * h * w * (T2_FC / 4);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (T2_FC / 4)) * 4;
  long long t = i / (T2_FC / 4);
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int img = (int)(t / h);
  int us[2], vs[2], nu = 1, nv = 1;      // unreflected coordinates u in [-1, h], v in [-1, w] that reflect onto (y, x)
  us[0] = y; vs[0] = x;
  if (y == 1) us[nu++] = -1;
  if (y == h - 2) us[nu++] = h;
  if (x == 1) vs[nv++] = -1;
  if (x == w - 2) vs[nv++] = w;
  // (h, w >= 4: y == 1 and y == h-2 never coincide with a third source)
  float acc[4] = {0, 0, 0, 0};
  for (int a = 0; a < nu && a < 2; ++a)
    for (int b = 0; b < nv && b < 2; ++b) {
      const int u = us[a], v = vs[b];
      const int ry = (u + 1) & 1, rx = (v + 1) & 1;
      const int py = (u + 1 - ry) >> 1, px = (v + 1 - rx) >> 1;
      const float4 g = *reinterpret_cast<const float4*>(dF + ((long long)img * Pk + py * ww1 + px) * (4 * T2_FC) +
                                                        (ry * 2 + rx) * T2_FC + c);
      acc[0] += g.x; acc[1] += g.y; acc[2] += g.z; acc[3] += g.w;
    }
  store4(dfeat + (((long long)img * h + y) * w + x) * T2_FC + c, (long long)n * h * w * T2_FC, acc);
}

// Fused backward of shift-add + softmax on the padded key grid (forward: gca_rowstats_kernel<true> + gca_shift_add_kernel):
//   dA[q][j] = sum_a dA2[q + a][j + sh_a]      delta = <A[q], dA[q]>      dS[q][j] = A[q][j] (dA[q][j] - delta)
// One CTA per (query row q, image); the gathered row stays in shared memory between the two passes, so HBM sees one read
// of A, one (L2-shared, 4x) read of dA2 and one split-bf16 write of dS instead of the gather / row-dot / softmax-backward
// round trips.  Columns with A == 0 (pad keys, columns >= Pk whose dA2 the GEMM never wrote) give dS = 0 exactly.
__global__ void __launch_bounds__(256) gca_softmax_bwd_grid_kernel(const float* __restrict__ A, const float* __restrict__ dA2,
                                                                   int n, int hh, int ww, int ld,
                                                                   __nv_bfloat16* __restrict__ dS) {
  extern __shared__ float t2_smem[];   // g[ld]
  __shared__ float red[8];
  float* gs = t2_smem;
  const int P = hh * ww, ww1 = ww + 1, Pk = (hh + 1) * ww1;
  const int q = blockIdx.x, img = blockIdx.y;
  const int qy = q / ww, qx = q - qy * ww;
  const float* rows[4];
  int sh[4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    sh[a] = (a >> 1) * ww1 + (a & 1);
    rows[a] = dA2 + ((long long)img * Pk + (qy + (a >> 1)) * ww1 + qx + (a & 1)) * ld + sh[a];
  }
  const float* arow = A + ((long long)img * P + q) * ld;
  float dot = 0.f;
  // the loads of an iteration do not depend on each other (a zero probability masks its column afterwards)
#pragma unroll 2
  for (int j = threadIdx.x * 4; j < ld; j += 1024) {
    const float4 av = __ldg(reinterpret_cast<const float4*>(arow + j));
    float g[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      const float* src = rows[a] + j;
      if (j + 4 + sh[a] <= ld) {
        const int al = sh[a] & 3;
        if (al == 0) {
          const float4 t = __ldg(reinterpret_cast<const float4*>(src));
          g[0] += t.x; g[1] += t.y; g[2] += t.z; g[3] += t.w;
        } else if (al == 2) {
          const float2 t0 = __ldg(reinterpret_cast<const float2*>(src)), t1 = __ldg(reinterpret_cast<const float2*>(src + 2));
          g[0] += t0.x; g[1] += t0.y; g[2] += t1.x; g[3] += t1.y;
        } else {
          const float t0 = __ldg(src);
          const float2 t1 = __ldg(reinterpret_cast<const float2*>(src + 1));
          const float t3 = __ldg(src + 3);
          g[0] += t0; g[1] += t1.x; g[2] += t1.y; g[3] += t3;
        }
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (j + e + sh[a] < ld) g[e] += __ldg(src + e);
      }
    }
    const float a4[4] = {av.x, av.y, av.z, av.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (a4[e] == 0.f) g[e] = 0.f;            // pad keys and columns >= Pk: dA2 there is unwritten memory
      dot = fmaf(a4[e], g[e], dot);
    }
    *reinterpret_cast<float4*>(gs + j) = make_float4(g[0], g[1], g[2], g[3]);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  float delta = 0.f;
#pragma unroll
  for (int k = 0; k < 8; ++k) delta += red[k];
  const long long plane = (long long)n * P * ld;
  __nv_bfloat16* out = dS + ((long long)img * P + q) * ld;
#pragma unroll 2
  for (int j = threadIdx.x * 4; j < ld; j += 1024) {   // each thread re-reads only what it wrote itself; A from L1/L2
    const float4 g = *reinterpret_cast<const float4*>(gs + j);
    const float4 a = __ldg(reinterpret_cast<const float4*>(arow + j));
    uint32_t h0, l0, h1, l1;
    split2_bf16(a.x * (g.x - delta), a.y * (g.y - delta), h0, l0);
    split2_bf16(a.z * (g.z - delta), a.w * (g.w - delta), h1, l1);
    *reinterpret_cast<uint2*>(out + j) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(out + plane + j) = make_uint2(l0, l1);
  }
}

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_gca_unfold_parity_bwd(const void* dY, int n, int h, int w, void* dO2, tcv_stream_t stream) {
  TCV_REQUIRE(dY && dO2, "gca_unfold_parity_bwd: null pointer");
  TCV_REQUIRE(n > 0 && h % 2 == 0 && w % 2 == 0 && h >= 4 && w >= 4, "gca_unfold_parity_bwd: h,w must be even and >= 4");
  const long long total = (long long)n * (h / 2 + 1) * (w / 2 + 1) * 4 * (T2_FC / 8);
  gca_unfold_parity_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(dY), n, h, w, reinterpret_cast<__nv_bfloat16*>(dO2));
  return launched("gca_unfold_parity_bwd_kernel");
}

int tcv_gca_values_parity_bwd(const float* dF, int n, int h, int w, void* dfeat, tcv_stream_t stream) {
  TCV_REQUIRE(dF && dfeat, "gca_values_parity_bwd: null pointer");
  TCV_REQUIRE(n > 0 && h % 2 == 0 && w % 2 == 0 && h >= 4 && w >= 4, "gca_values_parity_bwd: h,w must be even and >= 4");
  const long long total = (long long)n * h * w * (T2_FC / 4);
  gca_values_parity_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(dF, n, h, w,
                                                                                      reinterpret_cast<__nv_bfloat16*>(dfeat));
  return launched("gca_values_parity_bwd_kernel");
}

int tcv_gca_softmax_bwd_grid(const float* A, const float* dA2, int n, int h, int w, int ld, void* dS, tcv_stream_t stream) {
  TCV_REQUIRE(A && dA2 && dS, "gca_softmax_bwd_grid: null pointer");
  const int hh = h / 2, ww = w / 2;
  TCV_REQUIRE(n > 0 && hh > 0 && ww > 0 && ld % 4 == 0 && ld >= (hh + 1) * (ww + 1), "gca_softmax_bwd_grid: bad geometry");
  const size_t smem = (size_t)ld * sizeof(float);
  TCV_REQUIRE(smem <= 200 * 1024, "gca_softmax_bwd_grid: row of %d keys does not fit shared memory", ld);
  TCV_CUDA(cudaFuncSetAttribute(gca_softmax_bwd_grid_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  gca_softmax_bwd_grid_kernel<<<dim3(hh * ww, n), 256, smem, S(stream)>>>(A, dA2, n, hh, ww, ld,
                                                                         reinterpret_cast<__nv_bfloat16*>(dS));
  return launched("gca_softmax_bwd_grid_kernel");
}

}  // extern "C"
