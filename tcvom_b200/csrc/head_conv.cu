// Alpha head of the decoder: Conv2d(32 -> 1, 3x3, zero pad, bias) + (tanh + 1) / 2  (resnet_dec.py:80, VMN_GCA.py:46-47).
// HBM-bound (one read of the 32-channel full-resolution tensor, 4 B per output pixel).  Tile = 8 rows x 32 px per CTA: the
// (8+2) x (32+2) pixel halo arrives as two TMA boxes (bf16 hi / lo planes, no swizzle, out-of-image pixels zero-filled =
// the zero padding); thread = (pixel column, group of 4 rows, 8-channel chunk), so every shared-memory read is a
// conflict-free 16-byte access that serves 4 output rows; the 4 chunk lanes of a pixel are summed with two shuffles.
#include "tc_common.cuh"

namespace tcv {
constexpr int HD_TW = 32, HD_TH = 8, HD_C = 32;
constexpr int HD_PLANE = (HD_TH + 2) * (HD_TW + 2) * HD_C * 2;   // bytes of one plane of the halo tile (21 760)

__global__ void __launch_bounds__(256) head_conv_tanh01_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                               const __grid_constant__ CUtensorMap map_lo, int h, int w,
                                                               const float* __restrict__ wt, const float* __restrict__ bias,
                                                               float* __restrict__ pred) {
  __shared__ __align__(128) uint4 tile[2][HD_PLANE / 16];   // [plane][pixel][chunk of 8 channels]
  __shared__ float wsm[9 * HD_C];
  __shared__ __align__(8) unsigned long long bar;
  const int x0 = blockIdx.x * HD_TW, y0 = blockIdx.y * HD_TH, img = blockIdx.z;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_a, 2u * HD_PLANE);
    tma_load_4d(smem_u32(&tile[0][0]), &map_hi, bar_a, 0, x0 - 1, y0 - 1, img);
    tma_load_4d(smem_u32(&tile[1][0]), &map_lo, bar_a, 0, x0 - 1, y0 - 1, img);
  }
  for (int i = threadIdx.x; i < 9 * HD_C; i += 256) wsm[i] = wt[i];
  __syncthreads();                                    // barrier initialised, weights staged
  mbar_wait(bar_a, 0);
  const int ck = threadIdx.x & 3, tx = (threadIdx.x >> 2) & 31, rg = threadIdx.x >> 7;   // rows rg*4 .. rg*4+3
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int ry = 0; ry < 6; ++ry) {          // tile row rg*4 + ry = image row y0 + rg*4 + ry - 1
#pragma unroll
    for (int dx = 0; dx < 3; ++dx) {
      const int idx = ((rg * 4 + ry) * (HD_TW + 2) + tx + dx) * 4 + ck;
      const uint4 a = tile[0][idx], b = tile[1][idx];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
      float f[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
        f[2 * k + 1] = __uint_as_float(aw[k] & 0xffff0000u) + __uint_as_float(bw[k] & 0xffff0000u);
      }
#pragma unroll
      for (int r = 0; r < 4; ++r) {         // output row rg*4 + r sees this input row as vertical tap dy = ry - r
        const int dy = ry - r;
        if (dy < 0 || dy > 2) continue;
        const float* wv = wsm + (dy * 3 + dx) * HD_C + ck * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r] = fmaf(f[k], wv[k], acc[r]);
      }
    }
  }
  const float b0 = bias ? bias[0] : 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    float v = acc[r];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    const int gy = y0 + rg * 4 + r, gx = x0 + tx;
    if (ck == 0 && gy < h && gx < w) pred[((long long)img * h + gy) * w + gx] = (tanhf(v + b0) + 1.0f) * 0.5f;
  }
}
}  // namespace tcv

extern "C" int tcv_head_conv_tanh01(const void* x, long long x_plane, int n, int h, int w, const float* wt, const float* bias,
                                    float* pred, tcv_stream_t stream) {
  using namespace tcv;
  TCV_REQUIRE(x && wt && pred, "head_conv_tanh01: null pointer");
  TCV_REQUIRE(n > 0 && h > 0 && w > 0, "head_conv_tanh01: bad dims");
  if (x_plane == 0) x_plane = (long long)n * h * w * HD_C;
  TCV_REQUIRE(((uintptr_t)x & 15) == 0 && x_plane % 8 == 0, "head_conv_tanh01: x must be 16-byte aligned");
  CUtensorMap m_hi, m_lo;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  cuuint64_t dims[4] = {(cuuint64_t)HD_C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t str[3] = {(cuuint64_t)HD_C * 2, (cuuint64_t)w * HD_C * 2, (cuuint64_t)h * w * HD_C * 2};
  cuuint32_t box[4] = {(cuuint32_t)HD_C, (cuuint32_t)(HD_TW + 2), (cuuint32_t)(HD_TH + 2), 1};
  int rc = make_map(&m_hi, xb, 4, dims, str, box, 0);
  if (rc) return rc;
  rc = make_map(&m_lo, xb + x_plane, 4, dims, str, box, 0);
  if (rc) return rc;
  const dim3 grid((w + HD_TW - 1) / HD_TW, (h + HD_TH - 1) / HD_TH, n);
  head_conv_tanh01_kernel<<<grid, 256, 0, S(stream)>>>(m_hi, m_lo, h, w, wt, bias, pred);
  return launched("head_conv_tanh01_kernel");
}

// ---------------------------------------------------------------------------------------------------------------------
// Alpha head of the DIM decoder: Conv2d(64 -> 1, 5x5, zero pad 2, bias).clamp(0, 1)  (VMN_DIM.py:97,135).  Same scheme as
// above with a (8+4) x (32+4) pixel halo tile of 64 channels (two 55 KB planes in dynamic shared memory); thread = (pixel
// column, 8-channel chunk), all 8 rows of the tile: one 16-byte shared-memory read serves up to 5 output rows.
namespace tcv {
constexpr int H5_TW = 32, H5_TH = 8, H5_C = 64, H5_K = 5;
constexpr int H5_PLANE = (H5_TH + H5_K - 1) * (H5_TW + H5_K - 1) * H5_C * 2;   // 55 296 bytes

__global__ void __launch_bounds__(256) head_conv5_clamp01_kernel(const __grid_constant__ CUtensorMap map_hi,
                                                                 const __grid_constant__ CUtensorMap map_lo, int h, int w,
                                                                 const float* __restrict__ wt, const float* __restrict__ bias,
                                                                 float* __restrict__ pred) {
  extern __shared__ __align__(128) uint8_t h5_smem[];
  uint4* tile0 = reinterpret_cast<uint4*>(h5_smem);
  uint4* tile1 = reinterpret_cast<uint4*>(h5_smem + H5_PLANE);
  float* wsm = reinterpret_cast<float*>(h5_smem + 2 * H5_PLANE);            // [25][64]
  __shared__ __align__(8) unsigned long long bar;
  const int x0 = blockIdx.x * H5_TW, y0 = blockIdx.y * H5_TH, img = blockIdx.z;
  const uint32_t bar_a = smem_u32(&bar);
  if (threadIdx.x == 0) {
    mbar_init(bar_a, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_expect_tx(bar_a, 2u * H5_PLANE);
    tma_load_4d(smem_u32(tile0), &map_hi, bar_a, 0, x0 - 2, y0 - 2, img);
    tma_load_4d(smem_u32(tile1), &map_lo, bar_a, 0, x0 - 2, y0 - 2, img);
  }
  for (int i = threadIdx.x; i < H5_K * H5_K * H5_C; i += 256) wsm[i] = wt[i];
  __syncthreads();
  mbar_wait(bar_a, 0);
  const int ck = threadIdx.x & 7, tx = threadIdx.x >> 3;     // 8 chunks of 8 channels x 32 pixel columns
  float acc[H5_TH];
#pragma unroll
  for (int r = 0; r < H5_TH; ++r) acc[r] = 0.f;
#pragma unroll
  for (int ry = 0; ry < H5_TH + H5_K - 1; ++ry) {            // tile row ry = image row y0 + ry - 2
#pragma unroll
    for (int dx = 0; dx < H5_K; ++dx) {
      const int idx = (ry * (H5_TW + H5_K - 1) + tx + dx) * (H5_C / 8) + ck;
      const uint4 a = tile0[idx], b = tile1[idx];
      const uint32_t aw[4] = {a.x, a.y, a.z, a.w}, bw[4] = {b.x, b.y, b.z, b.w};
      float f[8];
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        f[2 * k] = __uint_as_float(aw[k] << 16) + __uint_as_float(bw[k] << 16);
        f[2 * k + 1] = __uint_as_float(aw[k] & 0xffff0000u) + __uint_as_float(bw[k] & 0xffff0000u);
      }
#pragma unroll
      for (int r = 0; r < H5_TH; ++r) {                      // output row r sees this input row as vertical tap dy = ry - r
        const int dy = ry - r;
        if (dy < 0 || dy >= H5_K) continue;
        const float* wv = wsm + (dy * H5_K + dx) * H5_C + ck * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) acc[r] = fmaf(f[k], wv[k], acc[r]);
      }
    }
  }
  const float b0 = bias ? bias[0] : 0.f;
#pragma unroll
  for (int r = 0; r < H5_TH; ++r) {
    float v = acc[r];
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 4);
    const int gy = y0 + r, gx = x0 + tx;
    if (ck == 0 && gy < h && gx < w) pred[((long long)img * h + gy) * w + gx] = fminf(fmaxf(v + b0, 0.f), 1.f);
  }
}
}  // namespace tcv

extern "C" int tcv_head_conv5_clamp01(const void* x, long long x_plane, int n, int h, int w, const float* wt, const float* bias,
                                      float* pred, tcv_stream_t stream) {
  using namespace tcv;
  TCV_REQUIRE(x && wt && pred, "head_conv5_clamp01: null pointer");
  TCV_REQUIRE(n > 0 && h > 0 && w > 0, "head_conv5_clamp01: bad dims");
  if (x_plane == 0) x_plane = (long long)n * h * w * H5_C;
  TCV_REQUIRE(((uintptr_t)x & 15) == 0 && x_plane % 8 == 0, "head_conv5_clamp01: x must be 16-byte aligned");
  CUtensorMap m_hi, m_lo;
  const __nv_bfloat16* xb = reinterpret_cast<const __nv_bfloat16*>(x);
  cuuint64_t dims[4] = {(cuuint64_t)H5_C, (cuuint64_t)w, (cuuint64_t)h, (cuuint64_t)n};
  cuuint64_t str[3] = {(cuuint64_t)H5_C * 2, (cuuint64_t)w * H5_C * 2, (cuuint64_t)h * w * H5_C * 2};
  cuuint32_t box[4] = {(cuuint32_t)H5_C, (cuuint32_t)(H5_TW + H5_K - 1), (cuuint32_t)(H5_TH + H5_K - 1), 1};
  int rc = make_map(&m_hi, xb, 4, dims, str, box, 0);
  if (rc) return rc;
  rc = make_map(&m_lo, xb + x_plane, 4, dims, str, box, 0);
  if (rc) return rc;
  const int smem = 2 * H5_PLANE + H5_K * H5_K * H5_C * (int)sizeof(float);      // 117 KB
  TCV_CUDA(cudaFuncSetAttribute(head_conv5_clamp01_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
  const dim3 grid((w + H5_TW - 1) / H5_TW, (h + H5_TH - 1) / H5_TH, n);
  head_conv5_clamp01_kernel<<<grid, 256, smem, S(stream)>>>(m_hi, m_lo, h, w, wt, bias, pred);
  return launched("head_conv5_clamp01_kernel");
}
