// Training-path utility kernels: frame gather/scatter for the decoder tail, gradient accumulation,
// pooling / padding / tanh gradients, bias gradients, weight-layout transposes for the data-gradient
// convolutions, SpectralNorm power iteration (models/GCA/ops.py:25-36) and the unpacking of packed weight
// gradients into the torch parameter layout including the d(sigma) term of W_bar / sigma.
// All HBM-bound single-pass kernels on split-bf16 NHWC tensors (see include/tcvom_b200.h).
#include <cooperative_groups.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace tcv {

static inline unsigned nb(long long total, int bs = 256) { return (unsigned)((total + bs - 1) / bs); }

__global__ void copy_images_kernel(const __nv_bfloat16* __restrict__ from, long long from_plane,
                                   __nv_bfloat16* __restrict__ to, long long to_plane, long long elems8, int n_pairs,
                                   int group, int g_from, int off_from, int g_to, int off_to, int accumulate) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)n_pairs * elems8) return;
  const int pair = (int)(i / elems8);
  const long long e = (i - (long long)pair * elems8) * 8;
  const int b = pair / group, j = pair - b * group;
  const long long src = ((long long)b * g_from + j + off_from) * elems8 * 8 + e;
  const long long dst = ((long long)b * g_to + j + off_to) * elems8 * 8 + e;
  if (accumulate) {
    float a[8], c[8];
    load8(from + src, from_plane, a);
    load8(to + dst, to_plane, c);
#pragma unroll
    for (int k = 0; k < 8; ++k) c[k] += a[k];
    store8(to + dst, to_plane, c);
  } else {
    *reinterpret_cast<uint4*>(to + dst) = *reinterpret_cast<const uint4*>(from + src);
    *reinterpret_cast<uint4*>(to + dst + to_plane) = *reinterpret_cast<const uint4*>(from + src + from_plane);
  }
}

__global__ void add_split_kernel(const __nv_bfloat16* __restrict__ x, long long x_plane, __nv_bfloat16* __restrict__ y,
                                 long long y_plane, long long elems8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= elems8) return;
  float a[8], c[8];
  load8(x + i * 8, x_plane, a);
  load8(y + i * 8, y_plane, c);
#pragma unroll
  for (int k = 0; k < 8; ++k) c[k] += a[k];
  store8(y + i * 8, y_plane, c);
}

// block = 256 threads = (256 / c8) pixel lanes x c8 channel groups; PIX_PER_BLOCK pixels per block
constexpr int CS_PIX = 1024;
__global__ void __launch_bounds__(256) channel_sum_kernel(const __nv_bfloat16* __restrict__ x, long long plane,
                                                          long long pixels, int c, float* __restrict__ out) {
  __shared__ float red[256][9];
  const int c8 = c / 8, lanes = 256 / c8;
  const int cg = threadIdx.x % c8, lane = threadIdx.x / c8;
  const long long p0 = (long long)blockIdx.x * CS_PIX;
  const long long p1 = min(p0 + CS_PIX, pixels);
  float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (long long p = p0 + lane; p < p1; p += lanes) {
    float f[8];
    load8(x + p * c + cg * 8, plane, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) s[k] += f[k];
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) red[threadIdx.x][k] = s[k];
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l)
#pragma unroll
      for (int k = 0; k < 8; ++k) s[k] += red[l * c8 + cg][k];
#pragma unroll
    for (int k = 0; k < 8; ++k) atomicAdd(out + cg * 8 + k, s[k]);
  }
}

__global__ void pool2_scaled_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, float scale,
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
  for (int k = 0; k < 8; ++k) acc[k] *= scale;
  store8(y + (((long long)img * oh + oy) * ow + ox) * c + cc, (long long)n * oh * ow * c, acc);
}

__global__ void upsample2_scaled_kernel(const __nv_bfloat16* __restrict__ x, int n, int h, int w, int c, float scale,
                                        __nv_bfloat16* __restrict__ y) {
  const int oh = 2 * h, ow = 2 * w, c8 = c / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * oh * ow * c8;
  if (i >= total) return;
  const int cc = (int)(i % c8) * 8;
  long long t = i / c8;
  const int ox = (int)(t % ow);
  t /= ow;
  const int oy = (int)(t % oh);
  const int img = (int)(t / oh);
  float f[8];
  load8(x + (((long long)img * h + oy / 2) * w + ox / 2) * c + cc, (long long)n * h * w * c, f);
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] *= scale;
  store8(y + (((long long)img * oh + oy) * ow + ox) * c + cc, (long long)n * oh * ow * c, f);
}

// dx[y][x] = sum of dy over the padded positions that read (y, x): (y+1, x+1) plus the mirrored border rows/cols
__global__ void pad_reflect1_bwd_kernel(const __nv_bfloat16* __restrict__ dy, int n, int h, int w, int c,
                                        __nv_bfloat16* __restrict__ dx) {
  const int ph = h + 2, pw = w + 2, c8 = c / 8;
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)n * h * w * c8;
  if (i >= total) return;
  const int cc = (int)(i % c8) * 8;
  long long t = i / c8;
  const int x = (int)(t % w);
  t /= w;
  const int y = (int)(t % h);
  const int img = (int)(t / h);
  int ys[2], xs[2], ny = 1, nx = 1;
  ys[0] = y + 1;
  xs[0] = x + 1;
  if (y == 1) ys[ny++] = 0;               // padded row 0 mirrors row 1
  if (y == h - 2) ys[ny++] = h + 1;       // padded row h+1 mirrors row h-2
  if (x == 1) xs[nx++] = 0;
  if (x == w - 2) xs[nx++] = w + 1;
  // h == 3 (or w == 3): y == 1 is both; ny can reach 3 only then -- guard by requiring h, w >= 4 on the host
  const long long pplane = (long long)n * ph * pw * c;
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int a = 0; a < ny; ++a)
    for (int b = 0; b < nx; ++b) {
      float f[8];
      load8(dy + (((long long)img * ph + ys[a]) * pw + xs[b]) * c + cc, pplane, f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  store8(dx + (((long long)img * h + y) * w + x) * c + cc, (long long)n * h * w * c, acc);
}

__global__ void tanh01_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ dpred, long long pixels,
                                  __nv_bfloat16* __restrict__ dz8) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  const float y = pred[i];
  float f[8] = {dpred[i] * 2.0f * y * (1.0f - y), 0, 0, 0, 0, 0, 0, 0};
  store8(dz8 + i * 8, pixels * 8, f);
}

// alpha head on the tensor-core path: the 32 -> 1 conv runs as a 32 -> 32 conv with zero-padded output channels;
// pred = (tanh(channel 0) + 1) / 2  (VMN_GCA.py:47), and its gradient goes back into channel 0 only
__global__ void head_tanh01_kernel(const __nv_bfloat16* __restrict__ x, long long plane, long long pixels, int c,
                                   float* __restrict__ pred) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels) return;
  pred[i] = (tanhf(load1(x + i * c, plane)) + 1.0f) * 0.5f;
}

__global__ void head_tanh01_bwd_kernel(const float* __restrict__ pred, const float* __restrict__ dpred, long long pixels,
                                       int c8, __nv_bfloat16* __restrict__ dz) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= pixels * c8) return;
  const long long px = i / c8;
  const int g = (int)(i - px * c8);
  float f[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (g == 0) {
    const float y = pred[px];
    f[0] = dpred[px] * 2.0f * y * (1.0f - y);
  }
  store8(dz + i * 8, pixels * c8 * 8, f);
}

__global__ void f32_to_split_kernel(const float* __restrict__ x, long long count4, __nv_bfloat16* __restrict__ y,
                                    long long plane) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  const float f[4] = {v.x, v.y, v.z, v.w};
  store4(y + i * 4, plane, f);
}

__global__ void split_to_f32_kernel(const __nv_bfloat16* __restrict__ x, long long plane, long long count4,
                                    float* __restrict__ y) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count4) return;
  float f[4];
  load4(x + i * 4, plane, f);
  reinterpret_cast<float4*>(y)[i] = make_float4(f[0], f[1], f[2], f[3]);
}

__global__ void transpose_packed_kernel(const float* __restrict__ in, int taps, int cin, int cout, int cout_pad,
                                        float* __restrict__ out) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long total = (long long)taps * cout_pad * cin;
  if (i >= total) return;
  const int ci = (int)(i % cin);
  const int co = (int)((i / cin) % cout_pad);
  const int t = (int)(i / ((long long)cin * cout_pad));
  out[i] = co < cout ? in[((long long)t * cin + ci) * cout + co] : 0.f;
}


// NHWC split-bf16 -> channel-major, zero-ringed: xt[ch][(img*(gh+2) + gy+1)*row + gx+1+shift] = x[img][gy*mul+oy][gx*mul+ox][ch]
// (the K-major operand layout of the tensor-core weight-gradient GEMM; the ring makes every 3x3 tap a pure shift;
// rows are padded to a multiple of 8 elements and horizontal tap offsets are baked in as `shift`, because TMA only
// accepts 16-byte aligned coordinates in the innermost dimension -- measured: unaligned ones fault).
// 32 pixels x 32 channels per block through shared memory, planes moved as raw 16-bit elements.
__global__ void __launch_bounds__(256) transpose_pad_kernel(const uint16_t* __restrict__ x, long long x_plane, int h, int w,
                                                            int c, int mul, int oy, int ox, int gh, int gw, int cblocks,
                                                            int row, int shift, uint16_t* __restrict__ xt,
                                                            long long xt_plane, long long ktot) {
  __shared__ uint16_t tile[2][32][34];
  const int img = blockIdx.z / cblocks, cb = blockIdx.z % cblocks;
  const int gy = blockIdx.y, gx0 = blockIdx.x * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;   // 32 x 8
  for (int pp = ty; pp < 32; pp += 8) {
    const int gx = gx0 + pp, ch = cb * 32 + tx;
    uint16_t a = 0, b = 0;
    if (gx < gw && ch < c) {
      const long long src = (((long long)img * h + (gy * mul + oy)) * w + (gx * mul + ox)) * c + ch;
      a = x[src];
      b = x[src + x_plane];
    }
    tile[0][pp][tx] = a;
    tile[1][pp][tx] = b;
  }
  __syncthreads();
  for (int cc = ty; cc < 32; cc += 8) {
    const int gx = gx0 + tx, ch = cb * 32 + cc;
    if (gx < gw && ch < c) {
      const long long dst = (long long)ch * ktot + ((long long)img * (gh + 2) + gy + 1) * row + gx + 1 + shift;
      xt[dst] = tile[0][tx][cc];
      xt[dst + xt_plane] = tile[1][tx][cc];
    }
  }
}

// ---- SpectralNorm power iteration --------------------------------------------------
__device__ float block_sum_1024(float v, float* red) {
  v = warp_sum(v);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  float t = (threadIdx.x < (blockDim.x >> 5)) ? red[threadIdx.x] : 0.f;
  if (threadIdx.x < 32) {
    t = warp_sum(t);
    if (threadIdx.x == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// One thread-block CLUSTER of 8 CTAs per layer (a single CTA streaming a 512x4608 weight ten times took 5.7 ms for
// the 70 layers, 8 % of the training step): phase 1 splits the columns of v = W^T u over the CTAs, phase 2 the rows of
// t = W v; the two scalar norms are reduced through distributed shared memory; u / v travel through global memory
// (cluster.sync() orders it).
constexpr int SN_CLUSTER = 8;
constexpr int SN_MAX_COLS_SLICE = 1024;   // 512x(512*4*4) deconv weight: 8192 / 8
constexpr int SN_MAX_ROWS_SLICE = 64;     // 512 / 8

__device__ __forceinline__ float cluster_sum(cg::cluster_group& cluster, float* slot, float v, float* red) {
  v = block_sum_1024(v, red);
  if (threadIdx.x == 0) *slot = v;
  cluster.sync();
  float t = 0.f;
  for (int r = 0; r < SN_CLUSTER; ++r) t += *cluster.map_shared_rank(slot, r);
  return t;
}

__global__ void __cluster_dims__(SN_CLUSTER, 1, 1) __launch_bounds__(1024)
    sn_power_iter_kernel(const tcv_sn_desc* __restrict__ descs) {
  __shared__ float red[33];
  __shared__ float part[4];                         // [iteration parity][phase]: a slot is rewritten two syncs later
  __shared__ float vacc[SN_MAX_COLS_SLICE];
  __shared__ float tloc[SN_MAX_ROWS_SLICE];
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const tcv_sn_desc d = descs[blockIdx.x / SN_CLUSTER];
  const int rows = d.rows, cols = d.cols;
  const float* W = d.w_bar;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarp = blockDim.x >> 5;
  const int cs = (cols + SN_CLUSTER - 1) / SN_CLUSTER, c0 = min(rank * cs, cols), c1 = min(c0 + cs, cols);
  const int rs = (rows + SN_CLUSTER - 1) / SN_CLUSTER, r0 = min(rank * rs, rows), r1 = min(r0 + rs, rows);
  for (int k = 0; k < d.calls; ++k) {
    const float* u_prev = k == 0 ? d.u : d.u_hist + (long long)(k - 1) * rows;
    float* vk = d.v_hist + (long long)k * cols;
    float* uk = d.u_hist + (long long)k * rows;
    // ---- phase 1: v[c0:c1] = sum_r W[r][c] u[r]  (lanes = 32 consecutive columns, warps stride the rows)
    for (int c = threadIdx.x; c < c1 - c0; c += blockDim.x) vacc[c] = 0.f;
    __syncthreads();
    for (int cb = c0; cb < c1; cb += 32) {
      const int c = cb + lane;
      float a = 0.f;
      if (c < c1)
        for (int r = warp; r < rows; r += nwarp) a = fmaf(W[(long long)r * cols + c], u_prev[r], a);
      if (c < c1) atomicAdd(&vacc[c - c0], a);
    }
    __syncthreads();
    float ss = 0.f;
    for (int c = threadIdx.x; c < c1 - c0; c += blockDim.x) ss += vacc[c] * vacc[c];
    ss = cluster_sum(cluster, &part[(k & 1) * 2], ss, red);
    const float invv = 1.0f / (sqrtf(ss) + 1e-12f);
    for (int c = threadIdx.x; c < c1 - c0; c += blockDim.x) vk[c0 + c] = vacc[c] * invv;
    cluster.sync();                                 // v complete and visible to the whole cluster
    // ---- phase 2: t[r0:r1] = W[r] . v  (one warp per row)
    float tt = 0.f;
    for (int r = r0 + warp; r < r1; r += nwarp) {
      float a = 0.f;
      for (int c = lane; c < cols; c += 32) a = fmaf(W[(long long)r * cols + c], vk[c], a);
      a = warp_sum(a);
      if (lane == 0) { tloc[r - r0] = a; tt += a * a; }
    }
    tt = cluster_sum(cluster, &part[(k & 1) * 2 + 1], tt, red);
    const float nrm = sqrtf(tt);
    const float invu = 1.0f / (nrm + 1e-12f);
    for (int r = threadIdx.x; r < r1 - r0; r += blockDim.x) uk[r0 + r] = tloc[r] * invu;
    if (rank == 0 && threadIdx.x == 0) {
      const float sigma = tt * invu;                // u . (W v) = |Wv|^2 / (|Wv| + 1e-12)
      d.sigma[k] = sigma;
      d.inv_sigma[k] = 1.0f / sigma;
    }
    cluster.sync();                                 // u complete and visible
  }
  if (d.calls > 0) {
    const float* ul = d.u_hist + (long long)(d.calls - 1) * rows;
    const float* vl = d.v_hist + (long long)(d.calls - 1) * cols;
    for (int r = r0 + threadIdx.x; r < r1; r += blockDim.x) d.u[r] = ul[r];
    for (int c = c0 + threadIdx.x; c < c1; c += blockDim.x) d.v[c] = vl[c];
  }
}

__global__ void weight_grad_unpack_kernel(const float* __restrict__ dw, int cout, int cin, int kh, int kw,
                                          int transposed, int cin_pad, int cout_pad, const float* __restrict__ u_hist,
                                          const float* __restrict__ v_hist, const float* __restrict__ sigma,
                                          const double* __restrict__ zdot, int calls, float* __restrict__ grad) {
  const long long total = (long long)cout * cin * kh * kw;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int kx = (int)(i % kw);
  const int ky = (int)((i / kw) % kh);
  const int d1 = (int)((i / ((long long)kw * kh)) % (transposed ? cout : cin));
  const int d0 = (int)(i / ((long long)kw * kh * (transposed ? cout : cin)));
  const int co = transposed ? d1 : d0, ci = transposed ? d0 : d1;
  float g = dw[((long long)(ky * kw + kx) * cin_pad + ci) * cout_pad + co];
  if (calls > 0) {
    const int rows = transposed ? cin : cout;
    const long long cols = total / rows;
    const int r = (int)(i / cols);
    const long long cidx = i - (long long)r * cols;
    float corr = 0.f;
    for (int k = 0; k < calls; ++k)
      corr += (float)(zdot[k] / (double)sigma[k]) * u_hist[(long long)k * rows + r] * v_hist[(long long)k * cols + cidx];
    g -= corr;
  }
  grad[i] = g;
}

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_copy_images(const void* from, long long from_plane, void* to, long long to_plane, long long img_elems,
                    int n_pairs, int group, int g_from, int off_from, int g_to, int off_to, int accumulate,
                    tcv_stream_t stream) {
  TCV_REQUIRE(from && to, "copy_images: null pointer");
  TCV_REQUIRE(img_elems % 8 == 0 && n_pairs > 0 && group > 0 && n_pairs % group == 0, "copy_images: bad geometry");
  TCV_REQUIRE(off_from >= 0 && off_to >= 0 && group + off_from <= g_from && group + off_to <= g_to,
              "copy_images: frame window outside the sample");
  const long long e8 = img_elems / 8;
  copy_images_kernel<<<nb((long long)n_pairs * e8), 256, 0, S(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(from), from_plane, reinterpret_cast<__nv_bfloat16*>(to), to_plane, e8,
      n_pairs, group, g_from, off_from, g_to, off_to, accumulate);
  return launched("copy_images_kernel");
}

int tcv_add_split(const void* x, long long x_plane, void* y, long long y_plane, long long elems,
                  tcv_stream_t stream) {
  TCV_REQUIRE(x && y && elems % 8 == 0, "add_split: null pointer or elems %% 8 != 0");
  add_split_kernel<<<nb(elems / 8), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_plane,
                                                        reinterpret_cast<__nv_bfloat16*>(y), y_plane, elems / 8);
  return launched("add_split_kernel");
}

static bool pow2_8_512(int c) { return c >= 8 && c <= 512 && (c & (c - 1)) == 0; }

int tcv_channel_sum(const void* x, long long x_plane, long long pixels, int c, float* out, tcv_stream_t stream) {
  TCV_REQUIRE(x && out && pixels > 0, "channel_sum: null pointer");
  TCV_REQUIRE(pow2_8_512(c), "channel_sum: c=%d must be a power of two in [8,512]", c);
  channel_sum_kernel<<<nb(pixels, CS_PIX), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_plane,
                                                               pixels, c, out);
  return launched("channel_sum_kernel");
}

int tcv_pool2_scaled(const void* x, int n, int h, int w, int c, float scale, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y, "pool2_scaled: null pointer");
  TCV_REQUIRE(h % 2 == 0 && w % 2 == 0 && c % 8 == 0, "pool2_scaled: h,w must be even and c%%8==0");
  const long long total = (long long)n * (h / 2) * (w / 2) * (c / 8);
  pool2_scaled_kernel<<<nb(total), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c, scale,
                                                       reinterpret_cast<__nv_bfloat16*>(y));
  return launched("pool2_scaled_kernel");
}

int tcv_upsample2_scaled(const void* x, int n, int h, int w, int c, float scale, void* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && c % 8 == 0, "upsample2_scaled: null pointer or c%%8 != 0");
  const long long total = (long long)n * (2 * h) * (2 * w) * (c / 8);
  upsample2_scaled_kernel<<<nb(total), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), n, h, w, c,
                                                           scale, reinterpret_cast<__nv_bfloat16*>(y));
  return launched("upsample2_scaled_kernel");
}

int tcv_pad_reflect1_bwd(const void* dy, int n, int h, int w, int c, void* dx, tcv_stream_t stream) {
  TCV_REQUIRE(dy && dx, "pad_reflect1_bwd: null pointer");
  TCV_REQUIRE(h >= 4 && w >= 4 && c % 8 == 0, "pad_reflect1_bwd: need h,w >= 4 and c %% 8 == 0");
  const long long total = (long long)n * h * w * (c / 8);
  pad_reflect1_bwd_kernel<<<nb(total), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(dy), n, h, w, c,
                                                           reinterpret_cast<__nv_bfloat16*>(dx));
  return launched("pad_reflect1_bwd_kernel");
}

int tcv_tanh01_bwd(const float* pred, const float* dpred, long long pixels, void* dz8, tcv_stream_t stream) {
  TCV_REQUIRE(pred && dpred && dz8 && pixels > 0, "tanh01_bwd: null pointer");
  tanh01_bwd_kernel<<<nb(pixels), 256, 0, S(stream)>>>(pred, dpred, pixels, reinterpret_cast<__nv_bfloat16*>(dz8));
  return launched("tanh01_bwd_kernel");
}

int tcv_head_tanh01(const void* x, long long x_plane, long long pixels, int c, float* pred, tcv_stream_t stream) {
  TCV_REQUIRE(x && pred && pixels > 0 && c >= 1, "head_tanh01: bad arguments");
  head_tanh01_kernel<<<nb(pixels), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x),
                                                       x_plane ? x_plane : pixels * c, pixels, c, pred);
  return launched("head_tanh01_kernel");
}

int tcv_head_tanh01_bwd(const float* pred, const float* dpred, long long pixels, int c, void* dz, tcv_stream_t stream) {
  TCV_REQUIRE(pred && dpred && dz && pixels > 0 && c % 8 == 0, "head_tanh01_bwd: bad arguments");
  head_tanh01_bwd_kernel<<<nb(pixels * (c / 8)), 256, 0, S(stream)>>>(pred, dpred, pixels, c / 8,
                                                                     reinterpret_cast<__nv_bfloat16*>(dz));
  return launched("head_tanh01_bwd_kernel");
}

int tcv_f32_to_split(const float* x, long long count, void* y, long long y_plane, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && count % 4 == 0, "f32_to_split: null pointer or count %% 4 != 0");
  f32_to_split_kernel<<<nb(count / 4), 256, 0, S(stream)>>>(x, count / 4, reinterpret_cast<__nv_bfloat16*>(y), y_plane);
  return launched("f32_to_split_kernel");
}

int tcv_split_to_f32(const void* x, long long x_plane, long long count, float* y, tcv_stream_t stream) {
  TCV_REQUIRE(x && y && count % 4 == 0, "split_to_f32: null pointer or count %% 4 != 0");
  split_to_f32_kernel<<<nb(count / 4), 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(x), x_plane,
                                                           count / 4, y);
  return launched("split_to_f32_kernel");
}

int tcv_transpose_packed(const float* packed, int taps, int cin, int cout, int cout_pad, float* out,
                         tcv_stream_t stream) {
  TCV_REQUIRE(packed && out && cout_pad >= cout, "transpose_packed: bad arguments");
  const long long total = (long long)taps * cout_pad * cin;
  transpose_packed_kernel<<<nb(total), 256, 0, S(stream)>>>(packed, taps, cin, cout, cout_pad, out);
  return launched("transpose_packed_kernel");
}

int tcv_sn_power_iter(const tcv_sn_desc* descs_device, int n, tcv_stream_t stream) {
  TCV_REQUIRE(descs_device && n > 0, "sn_power_iter: no layers");
  sn_power_iter_kernel<<<n * SN_CLUSTER, 1024, 0, S(stream)>>>(descs_device);   // one cluster of 8 CTAs per layer
  return launched("sn_power_iter_kernel");
}

int tcv_weight_grad_unpack(const float* dw, int cout, int cin, int kh, int kw, int transposed, int cin_pad,
                           int cout_pad, const float* u_hist, const float* v_hist, const float* sigma,
                           const double* zdot, int calls, float* grad, tcv_stream_t stream) {
  TCV_REQUIRE(dw && grad && cin_pad >= cin && cout_pad >= cout, "weight_grad_unpack: bad arguments");
  TCV_REQUIRE(calls == 0 || (u_hist && v_hist && sigma && zdot), "weight_grad_unpack: spectral-norm history missing");
  const long long total = (long long)cout * cin * kh * kw;
  weight_grad_unpack_kernel<<<nb(total), 256, 0, S(stream)>>>(dw, cout, cin, kh, kw, transposed, cin_pad, cout_pad,
                                                             u_hist, v_hist, sigma, zdot, calls, grad);
  return launched("weight_grad_unpack_kernel");
}

int tcv_transpose_pad(const void* x, long long x_plane, int n, int h, int w, int c, int mul, int off_y, int off_x,
                      int row_stride, int shift, void* xt, long long xt_plane, long long ktot, tcv_stream_t stream) {
  TCV_REQUIRE(x && xt && n > 0 && c > 0 && mul >= 1 && h % mul == 0 && w % mul == 0, "transpose_pad: bad arguments");
  const int gh = h / mul, gw = w / mul;
  TCV_REQUIRE(row_stride >= gw + 2 && shift >= -1 && shift <= 1, "transpose_pad: row_stride < gw+2 or |shift| > 1");
  TCV_REQUIRE(ktot >= (long long)n * (gh + 2) * row_stride && xt_plane >= (long long)c * ktot, "transpose_pad: output too small");
  if (x_plane == 0) x_plane = (long long)n * h * w * c;
  TCV_CUDA(cudaMemsetAsync(xt, 0, sizeof(uint16_t) * (size_t)c * ktot, S(stream)));
  TCV_CUDA(cudaMemsetAsync(reinterpret_cast<uint16_t*>(xt) + xt_plane, 0, sizeof(uint16_t) * (size_t)c * ktot, S(stream)));
  const int cblocks = (c + 31) / 32;
  dim3 grid((gw + 31) / 32, gh, n * cblocks);
  transpose_pad_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const uint16_t*>(x), x_plane, h, w, c, mul, off_y,
                                                   off_x, gh, gw, cblocks, row_stride, shift,
                                                   reinterpret_cast<uint16_t*>(xt), xt_plane, ktot);
  return launched("transpose_pad_kernel");
}

}  // extern "C"
