// Backward of guided contextual attention (autograd of models/GCA/ops.py:106-229), exact fp32 path.
//
// Forward (gca.cu):  g -> Q (3x3 reflect patches) ; Kn = Q/max(|Q|,1e-4)*scale_p ; S = Q.Kn^T - 1e4*diag*mm ;
//                    A = softmax_p(S) ; O = A.V (V = 4x4/stride-2 reflect patches of feat) ; Y = fold(O)/4.
// Backward:          dO = unfold(dY)/4 ; delta_q = <dO_q, O_q> ; dA = dO.V^T ; dS = A*(dA - delta) ;
//                    dV = A^T.dO -> dfeat ; dQ = dS.Kn ; dKn = dS^T.Q ; dQ += d(normalise)(dKn) ; dg = fold3x3(dQ).
// The four GEMMs run on tcv_gemm_f32_strided (below): operands are addressed with explicit (row, k) strides so
// that transposed uses need no materialised transpose.  At the training crop (512x512) P = 1024 and the
// attention is ~3 % of the step; a tcgen05 version is the next step for 1080p training.
#include "common.cuh"

namespace tcv {

constexpr int GC = 64, FC = 128, QD = 9 * GC, VD = 16 * FC;

// ---- generic strided fp32 GEMM: 64x64x16 tiles, 4x4 per thread --------------------------------------
constexpr int GB = 64, GK = 16;

__global__ void __launch_bounds__(256) gemm_f32_strided_kernel(const float* __restrict__ A, long long sam,
                                                               long long sak, const float* __restrict__ B,
                                                               long long sbn, long long sbk, float* __restrict__ C,
                                                               long long ldc, int M, int N, int K, long long bsA,
                                                               long long bsB, long long bsC, int accumulate) {
  __shared__ float As[GK][GB + 4];
  __shared__ float Bs[GK][GB + 4];
  A += (long long)blockIdx.z * bsA;
  B += (long long)blockIdx.z * bsB;
  C += (long long)blockIdx.z * bsC;
  const int m0 = blockIdx.y * GB, n0 = blockIdx.x * GB;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  for (int k0 = 0; k0 < K; k0 += GK) {
    __syncthreads();
    // 64 x 16 tile = 1024 elements, 4 per thread; the thread index runs along the unit-stride axis
    for (int s = threadIdx.x; s < GB * GK; s += 256) {
      int m, k;
      if (sak == 1) { k = s % GK; m = s / GK; } else { m = s % GB; k = s / GB; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < K) ? A[(long long)gm * sam + (long long)gk * sak] : 0.f;
    }
    for (int s = threadIdx.x; s < GB * GK; s += 256) {
      int n, k;
      if (sbk == 1) { k = s % GK; n = s / GK; } else { n = s % GB; k = s / GB; }
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < K) ? B[(long long)gn * sbn + (long long)gk * sbk] : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < GK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b4 = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w}, b[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= N) continue;
      float* c = C + (long long)m * ldc + n;
      *c = accumulate ? *c + acc[i][j] : acc[i][j];
    }
  }
}

// dO[q][(ty*4+tx)*128 + c] = dY[2qy+ty-1][2qx+tx-1][c] / 4 (zero outside); delta[q] = <dO[q], O[q]>
// one CTA of 128 threads per patch q: thread = channel, loop over the 16 taps
__global__ void __launch_bounds__(128) gca_fold_bwd_kernel(const __nv_bfloat16* __restrict__ dY,
                                                           const float* __restrict__ O, int n, int h, int w,
                                                           float* __restrict__ dO, float* __restrict__ delta,
                                                           __nv_bfloat16* __restrict__ dO_split) {
  __shared__ float red[4];
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  const int q = blockIdx.x, img = blockIdx.y;
  const int qy = q / ww, qx = q - qy * ww;
  const int c = threadIdx.x;
  const long long plane = (long long)n * h * w * FC;
  const long long row = ((long long)img * P + q) * VD;
  float dot = 0.f;
#pragma unroll
  for (int t = 0; t < 16; ++t) {
    const int y = 2 * qy + t / 4 - 1, x = 2 * qx + t % 4 - 1;
    float v = 0.f;
    if (y >= 0 && y < h && x >= 0 && x < w)
      v = 0.25f * load1(dY + (((long long)img * h + y) * w + x) * FC + c, plane);
    dO[row + t * FC + c] = v;
    if (dO_split) store1(dO_split + row + t * FC + c, (long long)n * P * VD, v);
    dot = fmaf(v, O[row + t * FC + c], dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) delta[(long long)img * P + q] = red[0] + red[1] + red[2] + red[3];
}

__global__ void gca_softmax_bwd_kernel(const float* __restrict__ A, float* __restrict__ dA,
                                       const float* __restrict__ delta, int P, int P_pad, long long rows,
                                       __nv_bfloat16* __restrict__ dS_split) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= rows * P_pad) return;
  const long long r = i / P_pad;
  const int p = (int)(i - r * P_pad);
  const float v = p < P ? A[i] * (dA[i] - delta[r]) : 0.f;
  dA[i] = v;
  if (dS_split) store1(dS_split + i, rows * P_pad, v);
}

// out[b][c][r] = in[b][r][c] for both planes of a split-bf16 matrix (raw 16-bit moves), r >= rows zero-filled up to ld_out.
// 64 x 64 tiles, 8-byte global accesses on both sides (rows of 128 B); blockIdx.z = batch * 2 + plane.
__global__ void __launch_bounds__(256) transpose_planes_kernel(const uint16_t* __restrict__ in, long long in_plane,
                                                               int rows, int cols, long long ld_in, long long bs_in,
                                                               uint16_t* __restrict__ out, long long out_plane,
                                                               long long ld_out, long long bs_out) {
  __shared__ uint16_t tile[64][66];       // pitch 33 words: a column read of the 64 rows is two-way bank-conflicted at worst
  const int b = blockIdx.z >> 1, pl = blockIdx.z & 1;
  const int r0 = blockIdx.y * 64, c0 = blockIdx.x * 64;
  const uint16_t* src = in + (long long)pl * in_plane + (long long)b * bs_in;
  uint16_t* dst = out + (long long)pl * out_plane + (long long)b * bs_out;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;       // 16 threads x 4 elements per 64-element row
  const bool vec_in = (ld_in & 3) == 0 && ((reinterpret_cast<uintptr_t>(src) & 7) == 0);
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int rr = pass * 16 + ty, r = r0 + rr, c = c0 + tx * 4;
    uint16_t v[4] = {0, 0, 0, 0};
    if (r < rows) {
      const uint16_t* p = src + (long long)r * ld_in + c;
      if (vec_in && c + 3 < cols) {
        const uint2 t = *reinterpret_cast<const uint2*>(p);
        v[0] = (uint16_t)(t.x & 0xFFFF); v[1] = (uint16_t)(t.x >> 16); v[2] = (uint16_t)(t.y & 0xFFFF); v[3] = (uint16_t)(t.y >> 16);
      } else {
#pragma unroll
        for (int e = 0; e < 4; ++e)
          if (c + e < cols) v[e] = p[e];
      }
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) tile[rr][tx * 4 + e] = v[e];
  }
  __syncthreads();
  const bool vec_out = (ld_out & 3) == 0 && ((reinterpret_cast<uintptr_t>(dst) & 7) == 0);
#pragma unroll
  for (int pass = 0; pass < 4; ++pass) {
    const int cc = pass * 16 + ty, c = c0 + cc, r = r0 + tx * 4;
    if (c >= cols || r >= ld_out) continue;
    uint16_t v[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) v[e] = tile[tx * 4 + e][cc];
    uint16_t* p = dst + (long long)c * ld_out + r;
    if (vec_out && r + 3 < ld_out) {
      *reinterpret_cast<uint2*>(p) = make_uint2((uint32_t)v[0] | ((uint32_t)v[1] << 16), (uint32_t)v[2] | ((uint32_t)v[3] << 16));
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e)
        if (r + e < ld_out) p[e] = v[e];
    }
  }
}

// dfeat[y][x][c] = sum over (py,ty),(px,tx) with reflect(2py+ty-1) = y, reflect(2px+tx-1) = x of dV[p][(ty*4+tx)*128+c]
__device__ __forceinline__ int value_sources(int y, int h, int hh, int* py, int* ty) {
  int cnt = 0;
#pragma unroll
  for (int a = 0; a < 2; ++a) {           // direct hits: ty = (y+1) & 1 (+2)
    const int t = ((y + 1) & 1) + 2 * a;
    const int r = y + 1 - t;
    if (r >= 0 && r / 2 < hh) { py[cnt] = r / 2; ty[cnt] = t; ++cnt; }
  }
  if (y == 1) { py[cnt] = 0; ty[cnt] = 0; ++cnt; }              // padded row -1 mirrors row 1
  if (y == h - 2) { py[cnt] = hh - 1; ty[cnt] = 3; ++cnt; }     // padded row h mirrors row h-2
  return cnt;
}

__global__ void gca_values_bwd_kernel(const float* __restrict__ dV, int n, int h, int w,
                                      __nv_bfloat16* __restrict__ dfeat) {
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  const long long total = (long long)n * h * w * (FC / 4);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (FC / 4)) * 4;
  long long t = i / (FC / 4);
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int img = (int)(t / h);
  int pys[4], tys[4], pxs[4], txs[4];
  const int ny = value_sources(y, h, hh, pys, tys), nx = value_sources(x, w, ww, pxs, txs);
  float acc[4] = {0, 0, 0, 0};
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(dV + ((long long)img * P + pys[a] * ww + pxs[b]) * VD +
                                                        (tys[a] * 4 + txs[b]) * FC + c);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
  store4(dfeat + (((long long)img * h + y) * w + x) * FC + c, (long long)n * h * w * FC, acc);
}

// dQ[p] += d(Kn)/d(Q) applied to dKn[p]; one warp per patch (576 = 18 per lane)
__global__ void gca_qgrad_kernel(float* __restrict__ dQ, const float* __restrict__ dKn, const float* __restrict__ Q,
                                 const float* __restrict__ mm, const float* __restrict__ scales, int P, long long rows,
                                 int ww, int Pk) {
  const long long r = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  const int img = (int)(r / P);
  const float* q = Q + r * QD;
  // ww > 0: dKn lives on the padded (hh+1) x (ww+1) key grid of the shift-sum form (rows py*(ww+1) + px, Pk per image)
  const int pl = (int)(r - (long long)img * P);
  const float* dk = dKn + (ww > 0 ? (long long)img * Pk + (pl / ww) * (ww + 1) + pl % ww : r) * QD;
  float qv[18], dv[18], ss = 0.f, qd = 0.f;
#pragma unroll
  for (int j = 0; j < 18; ++j) {
    qv[j] = q[j * 32 + lane];
    dv[j] = dk[j * 32 + lane];
    ss = fmaf(qv[j], qv[j], ss);
    qd = fmaf(qv[j], dv[j], qd);
  }
  ss = warp_sum(ss);
  qd = warp_sum(qd);
  const float scale = mm[r] > 0.f ? scales[2 * img] : scales[2 * img + 1];
  const float nrm = sqrtf(ss);
  float* out = dQ + r * QD;
  if (nrm > 1e-4f) {
    // Kn = s*Q/|Q|  ->  dQ = s*(dKn/|Q| - Q*<Q,dKn>/|Q|^3)
    const float a = scale / nrm, b = scale * qd / (nrm * ss);
#pragma unroll
    for (int j = 0; j < 18; ++j) out[j * 32 + lane] += a * dv[j] - b * qv[j];
  } else {
    const float a = scale / 1e-4f;       // clamped norm: Kn = s*Q/1e-4
#pragma unroll
    for (int j = 0; j < 18; ++j) out[j * 32 + lane] += a * dv[j];
  }
}

__device__ __forceinline__ int patch_sources(int y, int hh, int* py, int* kh) {
  int cnt = 0;
#pragma unroll
  for (int k = 0; k < 3; ++k) {           // py + k - 1 = y
    const int p = y + 1 - k;
    if (p >= 0 && p < hh) { py[cnt] = p; kh[cnt] = k; ++cnt; }
  }
  if (y == 1) { py[cnt] = 0; kh[cnt] = 0; ++cnt; }              // padded row -1 (py = 0, k = 0) mirrors row 1
  if (y == hh - 2) { py[cnt] = hh - 1; kh[cnt] = 2; ++cnt; }    // padded row hh (py = hh-1, k = 2) mirrors row hh-2
  return cnt;
}

// dg[y][x][c] = sum over taps of dQ[p][(kh*3+kw)*64 + c] with reflect(py+kh-1) = y, reflect(px+kw-1) = x
__global__ void gca_patch_bwd_kernel(const float* __restrict__ dQ, int n, int hh, int ww,
                                     __nv_bfloat16* __restrict__ dg) {
  const int P = hh * ww;
  const long long total = (long long)n * P * (GC / 4);
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = (int)(i % (GC / 4)) * 4;
  long long t = i / (GC / 4);
  const int x = (int)(t % ww);
  t /= ww;
  const int y = (int)(t % hh);
  const int img = (int)(t / hh);
  int pys[5], khs[5], pxs[5], kws[5];
  const int ny = patch_sources(y, hh, pys, khs), nx = patch_sources(x, ww, pxs, kws);
  float acc[4] = {0, 0, 0, 0};
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      const float4 v = *reinterpret_cast<const float4*>(dQ + ((long long)img * P + pys[a] * ww + pxs[b]) * QD +
                                                        (khs[a] * 3 + kws[b]) * GC + c);
      acc[0] += v.x; acc[1] += v.y; acc[2] += v.z; acc[3] += v.w;
    }
  store4(dg + (((long long)img * hh + y) * ww + x) * GC + c, (long long)n * P * GC, acc);
}

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_gemm_f32_strided(const float* A, long long sam, long long sak, const float* B, long long sbn, long long sbk,
                         float* C, long long ldc, int M, int N, int K, long long strideA, long long strideB,
                         long long strideC, int batch, int accumulate, tcv_stream_t stream) {
  TCV_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && batch > 0, "gemm_f32_strided: bad arguments");
  TCV_REQUIRE((sam == 1 || sak == 1) && (sbn == 1 || sbk == 1), "gemm_f32_strided: each operand needs a unit stride");
  dim3 grid((N + GB - 1) / GB, (M + GB - 1) / GB, batch);
  gemm_f32_strided_kernel<<<grid, 256, 0, S(stream)>>>(A, sam, sak, B, sbn, sbk, C, ldc, M, N, K, strideA, strideB,
                                                      strideC, accumulate);
  return launched("gemm_f32_strided_kernel");
}

int tcv_transpose_planes(const void* in, long long in_plane, int rows, int cols, long long ld_in, long long bs_in,
                         void* out, long long out_plane, long long ld_out, long long bs_out, int batch,
                         tcv_stream_t stream) {
  TCV_REQUIRE(in && out && rows > 0 && cols > 0 && ld_in >= cols && ld_out >= rows && batch > 0,
              "transpose_planes: bad arguments");
  dim3 grid((cols + 63) / 64, (unsigned)((ld_out + 63) / 64), batch * 2);
  transpose_planes_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const uint16_t*>(in), in_plane, rows, cols, ld_in,
                                                      bs_in, reinterpret_cast<uint16_t*>(out), out_plane, ld_out, bs_out);
  return launched("transpose_planes_kernel");
}

int tcv_gca_fold_bwd(const void* dY, const float* O, int n, int h, int w, float* dO, float* delta, void* dO_split,
                     tcv_stream_t stream) {
  TCV_REQUIRE(dY && O && dO && delta, "gca_fold_bwd: null pointer");
  TCV_REQUIRE(h % 2 == 0 && w % 2 == 0 && h >= 4 && w >= 4, "gca_fold_bwd: h,w must be even and >= 4");
  dim3 grid((h / 2) * (w / 2), n);
  gca_fold_bwd_kernel<<<grid, 128, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dY), O, n, h, w, dO, delta,
                                                   reinterpret_cast<__nv_bfloat16*>(dO_split));
  return launched("gca_fold_bwd_kernel");
}

int tcv_gca_softmax_bwd(const float* A, float* dA, const float* delta, int n, int P, int P_pad, void* dS_split,
                        tcv_stream_t stream) {
  TCV_REQUIRE(A && dA && delta && P > 0 && P_pad >= P, "gca_softmax_bwd: bad arguments");
  const long long rows = (long long)n * P;
  const long long total = rows * P_pad;
  gca_softmax_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(A, dA, delta, P, P_pad, rows,
                                                                                reinterpret_cast<__nv_bfloat16*>(dS_split));
  return launched("gca_softmax_bwd_kernel");
}

int tcv_gca_values_bwd(const float* dV, int n, int h, int w, void* dfeat, tcv_stream_t stream) {
  TCV_REQUIRE(dV && dfeat, "gca_values_bwd: null pointer");
  TCV_REQUIRE(h % 2 == 0 && w % 2 == 0 && h >= 4 && w >= 4, "gca_values_bwd: h,w must be even and >= 4");
  const long long total = (long long)n * h * w * (FC / 4);
  gca_values_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(dV, n, h, w,
                                                                               reinterpret_cast<__nv_bfloat16*>(dfeat));
  return launched("gca_values_bwd_kernel");
}

static int gca_prep_bwd_impl(float* dQ, const float* dKn, const float* Q, const float* mm, const float* scales, int n,
                             int h, int w, void* dg, int key_grid, tcv_stream_t stream) {
  TCV_REQUIRE(dQ && dKn && Q && mm && scales && dg, "gca_prep_bwd: null pointer");
  TCV_REQUIRE(h % 2 == 0 && w % 2 == 0 && h >= 8 && w >= 8, "gca_prep_bwd: h,w must be even and >= 8");
  const int hh = h / 2, ww = w / 2, P = hh * ww;
  const long long rows = (long long)n * P;
  gca_qgrad_kernel<<<(unsigned)((rows * 32 + 255) / 256), 256, 0, S(stream)>>>(dQ, dKn, Q, mm, scales, P, rows,
                                                                              key_grid ? ww : 0, (hh + 1) * (ww + 1));
  int rc = launched("gca_qgrad_kernel");
  if (rc) return rc;
  const long long total = rows * (GC / 4);
  gca_patch_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(dQ, n, hh, ww,
                                                                              reinterpret_cast<__nv_bfloat16*>(dg));
  return launched("gca_patch_bwd_kernel");
}

int tcv_gca_prep_bwd(float* dQ, const float* dKn, const float* Q, const float* mm, const float* scales, int n,
                     int h, int w, void* dg, tcv_stream_t stream) {
  return gca_prep_bwd_impl(dQ, dKn, Q, mm, scales, n, h, w, dg, 0, stream);
}

int tcv_gca_prep_bwd_grid(float* dQ, const float* dKn_grid, const float* Q, const float* mm, const float* scales, int n,
                          int h, int w, void* dg, tcv_stream_t stream) {
  return gca_prep_bwd_impl(dQ, dKn_grid, Q, mm, scales, n, h, w, dg, 1, stream);
}

}  // extern "C"
