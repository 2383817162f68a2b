// fp32 CUDA-core GEMM  C = A * B^T  (both operands K-contiguous).  Used as the exact-precision
// path for the guided-contextual-attention GEMMs (scores Q.K^T and aggregation P.V) and as the
// on-device fp32 cross-check of the tcgen05 kernels.  128x128x8 CTA tile, 8x8 register tile,
// double-buffered shared memory.
#include "common.cuh"

namespace tcv {

constexpr int BM = 128, BN = 128, BK = 8;

__global__ void __launch_bounds__(256) gemm_tn_f32_kernel(const float* __restrict__ A, const float* __restrict__ B,
                                                          float* __restrict__ C, int M, int N, int K, int lda,
                                                          int ldb, int ldc, long long sA, long long sB,
                                                          long long sC) {
  __shared__ float As[2][BK][BM + 4];
  __shared__ float Bs[2][BK][BN + 4];
  A += (long long)blockIdx.z * sA;
  B += (long long)blockIdx.z * sB;
  C += (long long)blockIdx.z * sC;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;  // 16 x 16 threads, each 8x8 outputs (strided by 16... no: contiguous 4+4)
  // loader mapping: each thread loads one float4 (4 consecutive k) of one row for A and B
  const int lrow = tid >> 1;        // 0..127
  const int lk = (tid & 1) * 4;     // 0 or 4
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  auto load_tile = [&](int k0, float4& ra, float4& rb) {
    const int am = m0 + lrow, bn = n0 + lrow;
    ra = am < M ? *reinterpret_cast<const float4*>(A + (long long)am * lda + k0 + lk) : make_float4(0, 0, 0, 0);
    rb = bn < N ? *reinterpret_cast<const float4*>(B + (long long)bn * ldb + k0 + lk) : make_float4(0, 0, 0, 0);
  };
  auto store_tile = [&](int buf, const float4& ra, const float4& rb) {
    As[buf][lk + 0][lrow] = ra.x; As[buf][lk + 1][lrow] = ra.y; As[buf][lk + 2][lrow] = ra.z; As[buf][lk + 3][lrow] = ra.w;
    Bs[buf][lk + 0][lrow] = rb.x; Bs[buf][lk + 1][lrow] = rb.y; Bs[buf][lk + 2][lrow] = rb.z; Bs[buf][lk + 3][lrow] = rb.w;
  };

  float4 ra, rb;
  load_tile(0, ra, rb);
  store_tile(0, ra, rb);
  __syncthreads();
  const int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < nk) load_tile((kt + 1) * BK, ra, rb);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_tile(buf ^ 1, ra, rb);
      __syncthreads();
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      const int n = n0 + jh * 64 + tx * 4;
      float* c = C + (long long)m * ldc + n;
      if (n + 3 < N && (ldc & 3) == 0) {
        *reinterpret_cast<float4*>(c) = make_float4(acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2],
                                                    acc[i][jh * 4 + 3]);
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (n + j < N) c[j] = acc[i][jh * 4 + j];
      }
    }
  }
}

}  // namespace tcv

using namespace tcv;

extern "C" int tcv_gemm_tn_f32(const float* A, const float* B, float* C, int M, int N, int K, int lda, int ldb,
                               int ldc, long long strideA, long long strideB, long long strideC, int batch,
                               tcv_stream_t stream) {
  TCV_REQUIRE(A && B && C, "gemm_tn_f32: null pointer");
  TCV_REQUIRE(M > 0 && N > 0 && K > 0 && K % 8 == 0, "gemm_tn_f32: K must be a positive multiple of 8");
  TCV_REQUIRE(lda % 4 == 0 && ldb % 4 == 0, "gemm_tn_f32: lda/ldb must be multiples of 4");
  TCV_REQUIRE(((uintptr_t)A & 15) == 0 && ((uintptr_t)B & 15) == 0 && ((uintptr_t)C & 15) == 0,
              "gemm_tn_f32: pointers must be 16-byte aligned");
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, batch);
  gemm_tn_f32_kernel<<<grid, 256, 0, S(stream)>>>(A, B, C, M, N, K, lda, ldb, ldc, strideA, strideB, strideC);
  return launched("gemm_tn_f32_kernel");
}
