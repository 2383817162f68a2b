// Train-mode BatchNorm2d (batch statistics per frame group, running-stat update, backward), fused with the
// per-call SpectralNorm scale 1/sigma, the activation and the residual adds that surround it in the reference
// blocks (resnet_enc.py:33-49, resnet_dec.py:43-59, res_gca_enc.py:20-33,47-55, ops.py:227).
//
// The reference runs the encoder / decoder once per frame (VMN_model.py:93-98,107-110), so every BatchNorm
// sees a batch of B images per call; here all frames are one NHWC tensor and image i belongs to statistics
// group i % groups.  nn.SyncBatchNorm is obtained by all-reducing the `sums` buffers between the two halves of
// the forward (stats -> finalize/apply) and of the backward (reduce -> apply); the host does that with NCCL.
//
// HBM-bound: every kernel reads its operands once.  Block = 256 threads = (256 / c8) pixel lanes x c8 channel
// groups of 8; per-channel partial sums are reduced through shared memory and added with double atomics.
#include "common.cuh"

namespace tcv {

constexpr int BN_PIX_MAX = 2048;  // most pixels of one image per block (fewer when the tensor is small: the grid
                                  // should cover the 148 SMs several times, see bn_pix())

struct BnGeom {
  int c8, lanes, cg, lane;
  int p0, p1;
  int img, g;
};

__device__ __forceinline__ BnGeom bn_geom(const tcv_bn_desc& d, int pix) {
  BnGeom q;
  q.c8 = d.c / 8;
  q.lanes = 256 / q.c8;
  q.cg = threadIdx.x % q.c8;
  q.lane = threadIdx.x / q.c8;
  const int hw = d.h * d.w;
  q.p0 = blockIdx.x * pix;
  q.p1 = min(q.p0 + pix, hw);
  q.img = blockIdx.y;
  q.g = q.img % d.groups;
  return q;
}

__device__ __forceinline__ float act_grad(float t, int act) {
  if (act == TCV_ACT_RELU) return t > 0.f ? 1.f : 0.f;
  if (act == TCV_ACT_LEAKY02) return t > 0.f ? 1.f : 0.2f;
  return 1.f;
}

// reduces s0[8], s1[8] over the pixel lanes of the block and adds them to sums[(g*c + ch)*2 + {0,1}]
__device__ __forceinline__ void bn_block_reduce(const BnGeom& q, float* s0, float* s1, int c, double* sums,
                                                float (*red)[17]) {
#pragma unroll
  for (int k = 0; k < 8; ++k) { red[threadIdx.x][k] = s0[k]; red[threadIdx.x][8 + k] = s1[k]; }
  __syncthreads();
  if (q.lane == 0) {
    for (int l = 1; l < q.lanes; ++l)
#pragma unroll
      for (int k = 0; k < 8; ++k) { s0[k] += red[l * q.c8 + q.cg][k]; s1[k] += red[l * q.c8 + q.cg][8 + k]; }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      double* dst = sums + ((long long)q.g * c + q.cg * 8 + k) * 2;
      atomicAdd(dst, (double)s0[k]);
      atomicAdd(dst + 1, (double)s1[k]);
    }
  }
}

__global__ void __launch_bounds__(256) bn_stats_kernel(const tcv_bn_desc d, double* __restrict__ sums, int pix) {
  __shared__ float red[256][17];
  const BnGeom q = bn_geom(d, pix);
  const __nv_bfloat16* z = reinterpret_cast<const __nv_bfloat16*>(d.z) + (long long)q.img * d.h * d.w * d.c + q.cg * 8;
  const float is = d.inv_sigma ? d.inv_sigma[q.g] : 1.f;
  float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = q.p0 + q.lane; p < q.p1; p += q.lanes) {
    float f[8];
    load8(z + (long long)p * d.c, d.z_plane, f);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float t = f[k] * is;
      if (d.mode == 2) t = apply_act(t, d.act);
      s0[k] += t;
      s1[k] = fmaf(t, t, s1[k]);
    }
  }
  bn_block_reduce(q, s0, s1, d.c, sums, red);
}

__global__ void bn_finalize_kernel(const double* __restrict__ sums, double count, double ucount, int groups, int c, float eps,
                                   float momentum, float* __restrict__ mean, float* __restrict__ invstd,
                                   float* __restrict__ running_mean, float* __restrict__ running_var) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  float rm = running_mean ? running_mean[ch] : 0.f, rv = running_var ? running_var[ch] : 0.f;
  for (int g = 0; g < groups; ++g) {
    const double m = sums[((long long)g * c + ch) * 2] / count;
    double var = sums[((long long)g * c + ch) * 2 + 1] / count - m * m;
    var = var < 0 ? 0 : var;
    mean[g * c + ch] = (float)m;
    invstd[g * c + ch] = (float)(1.0 / sqrt(var + (double)eps));
    const double unbiased = ucount > 1 ? var * ucount / (ucount - 1) : var;
    rm = (1.f - momentum) * rm + momentum * (float)m;
    rv = (1.f - momentum) * rv + momentum * (float)unbiased;
  }
  if (running_mean) running_mean[ch] = rm;
  if (running_var) running_var[ch] = rv;
}

// value of up(res1) at output pixel (y, x): res1 is [n, h>>shift, w>>shift, c]
__device__ __forceinline__ void load_res1(const tcv_bn_desc& d, int img, int y, int x, int ch, float* r) {
  const int rh = d.h >> d.res1_shift, rw = d.w >> d.res1_shift;
  const __nv_bfloat16* p = reinterpret_cast<const __nv_bfloat16*>(d.res1) +
                           (((long long)img * rh + (y >> d.res1_shift)) * rw + (x >> d.res1_shift)) * d.c + ch;
  load8(p, d.res1_plane, r);
}

__global__ void __launch_bounds__(256) bn_apply_kernel(const tcv_bn_desc d, const bool vec) {
  const int c8 = d.c / 8;
  const long long total = (long long)d.n * d.h * d.w * c8;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int ch = (int)(i % c8) * 8;
  const long long pix = i / c8;
  const int x = (int)(pix % d.w);
  const int y = (int)((pix / d.w) % d.h);
  const int img = (int)(pix / ((long long)d.w * d.h));
  const int g = img % d.groups;
  const float is = d.inv_sigma ? d.inv_sigma[g] : 1.f;
  float f[8];
  load8(reinterpret_cast<const __nv_bfloat16*>(d.z) + pix * d.c + ch, d.z_plane, f);
  // per-channel parameters as 16-byte loads: 32 scalar loads per thread (lanes 32 bytes apart -> 32 sectors per request)
  // made the L1, not DRAM, the limit of this kernel
  float mu[8], iv[8], ga[8], be[8];
  if (vec) {
    ldf8(d.mean + g * d.c + ch, mu);
    ldf8(d.invstd + g * d.c + ch, iv);
    ldf8(d.gamma + ch, ga);
    ldf8(d.beta + ch, be);
  } else {          // a parameter tensor that is not 16-byte aligned (a view into a flat buffer)
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      mu[k] = d.mean[g * d.c + ch + k];
      iv[k] = d.invstd[g * d.c + ch + k];
      ga[k] = d.gamma[ch + k];
      be[k] = d.beta[ch + k];
    }
  }
  if (d.mode == 1) {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = (f[k] * is - mu[k]) * iv[k] * ga[k] + be[k];
    if (d.res1) {
      float r[8];
      load_res1(d, img, y, x, ch, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
    apply_act_n<8>(f, d.act);
    if (d.res2) {
      float r[8];
      load8(reinterpret_cast<const __nv_bfloat16*>(d.res2) + pix * d.c + ch, d.res2_plane, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) f[k] += r[k];
    }
  } else {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] *= is;
    apply_act_n<8>(f, d.act);
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = (f[k] - mu[k]) * iv[k] * ga[k] + be[k];
  }
  store8(reinterpret_cast<__nv_bfloat16*>(d.y) + pix * d.c + ch, d.y_plane, f);
}

__global__ void __launch_bounds__(256) bn_bwd_reduce_kernel(const tcv_bn_desc d, const __nv_bfloat16* __restrict__ dy,
                                                            long long dy_plane, __nv_bfloat16* __restrict__ e,
                                                            long long e_plane, double* __restrict__ sums, int pix) {
  __shared__ float red[256][17];
  const BnGeom q = bn_geom(d, pix);
  const long long ibase = (long long)q.img * d.h * d.w;
  const int ch = q.cg * 8;
  const float is = d.inv_sigma ? d.inv_sigma[q.g] : 1.f;
  float mu[8], iv[8], ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mu[k] = d.mean[q.g * d.c + ch + k];
    iv[k] = d.invstd[q.g * d.c + ch + k];
    ga[k] = d.gamma[ch + k];
    be[k] = d.beta[ch + k];
  }
  float s0[8] = {0, 0, 0, 0, 0, 0, 0, 0}, s1[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int p = q.p0 + q.lane; p < q.p1; p += q.lanes) {
    const long long pix = ibase + p;
    float f[8], gy[8], xh[8];
    load8(reinterpret_cast<const __nv_bfloat16*>(d.z) + pix * d.c + ch, d.z_plane, f);
    load8(dy + pix * d.c + ch, dy_plane, gy);
    if (d.mode == 1) {
      float r[8] = {0, 0, 0, 0, 0, 0, 0, 0};
      if (d.res1) load_res1(d, q.img, p / d.w, p % d.w, ch, r);
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        xh[k] = (f[k] * is - mu[k]) * iv[k];
        const float tpre = xh[k] * ga[k] + be[k] + r[k];
        gy[k] *= act_grad(tpre, d.act);
      }
      if (e) store8(e + pix * d.c + ch, e_plane, gy);
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) xh[k] = (apply_act(f[k] * is, d.act) - mu[k]) * iv[k];
    }
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      s0[k] += gy[k];
      s1[k] = fmaf(gy[k], xh[k], s1[k]);
    }
  }
  bn_block_reduce(q, s0, s1, d.c, sums, red);
}

__global__ void bn_param_grads_kernel(const double* __restrict__ sums, int groups, int c, float* __restrict__ dgamma,
                                      float* __restrict__ dbeta) {
  const int ch = blockIdx.x * blockDim.x + threadIdx.x;
  if (ch >= c) return;
  double a = 0, b = 0;
  for (int g = 0; g < groups; ++g) {
    b += sums[((long long)g * c + ch) * 2];
    a += sums[((long long)g * c + ch) * 2 + 1];
  }
  dgamma[ch] += (float)a;
  dbeta[ch] += (float)b;
}

__global__ void __launch_bounds__(256) bn_bwd_apply_kernel(const tcv_bn_desc d, const __nv_bfloat16* __restrict__ e,
                                                           long long e_plane, const double* __restrict__ sums,
                                                           double count, __nv_bfloat16* __restrict__ dz,
                                                           long long dz_plane, double* __restrict__ zdot, int pix,
                                                           int e_is_dy) {
  __shared__ double dred[8];
  const BnGeom q = bn_geom(d, pix);
  const long long ibase = (long long)q.img * d.h * d.w;
  const int ch = q.cg * 8;
  const float is = d.inv_sigma ? d.inv_sigma[q.g] : 1.f;
  float mu[8], iv[8], m1[8], m2[8], gi[8], ga[8], be[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mu[k] = d.mean[q.g * d.c + ch + k];
    iv[k] = d.invstd[q.g * d.c + ch + k];
    ga[k] = d.gamma[ch + k];
    be[k] = d.beta[ch + k];
    gi[k] = d.gamma[ch + k] * iv[k];
    m1[k] = (float)(sums[((long long)q.g * d.c + ch + k) * 2] / count);
    m2[k] = (float)(sums[((long long)q.g * d.c + ch + k) * 2 + 1] / count);
  }
  float dot = 0.f;
  for (int p = q.p0 + q.lane; p < q.p1; p += q.lanes) {
    const long long pix = ibase + p;
    float f[8], ge[8], o[8];
    load8(reinterpret_cast<const __nv_bfloat16*>(d.z) + pix * d.c + ch, d.z_plane, f);
    load8(e + pix * d.c + ch, e_plane, ge);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const float t0 = f[k] * is;
      if (d.mode == 1) {
        const float xh = (t0 - mu[k]) * iv[k];
        if (e_is_dy) ge[k] *= act_grad(xh * ga[k] + be[k], d.act);     // same expression as bn_bwd_reduce (no res1)
        o[k] = gi[k] * (ge[k] - m1[k] - xh * m2[k]) * is;
      } else {
        const float xh = (apply_act(t0, d.act) - mu[k]) * iv[k];
        o[k] = gi[k] * (ge[k] - m1[k] - xh * m2[k]) * act_grad(t0, d.act) * is;
      }
      dot = fmaf(f[k], o[k], dot);
    }
    store8(dz + pix * d.c + ch, dz_plane, o);
  }
  if (zdot) {
    dot = warp_sum(dot);
    if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = (double)dot;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0;
      for (int i = 0; i < 8; ++i) t += dred[i];
      atomicAdd(zdot + q.g, t);
    }
  }
}

__global__ void __launch_bounds__(256) group_dot_kernel(const __nv_bfloat16* __restrict__ z, long long z_plane,
                                                        const __nv_bfloat16* __restrict__ dz, long long dz_plane,
                                                        long long img_elems8, int groups, double* __restrict__ zdot) {
  __shared__ double dred[8];
  const int img = blockIdx.y;
  float dot = 0.f;
  const long long e0 = (long long)blockIdx.x * 4096;
  const long long e1 = min(e0 + 4096, img_elems8);
  for (long long i = e0 + threadIdx.x; i < e1; i += 256) {
    float a[8], b[8];
    load8(z + ((long long)img * img_elems8 + i) * 8, z_plane, a);
    load8(dz + ((long long)img * img_elems8 + i) * 8, dz_plane, b);
#pragma unroll
    for (int k = 0; k < 8; ++k) dot = fmaf(a[k], b[k], dot);
  }
  dot = warp_sum(dot);
  if ((threadIdx.x & 31) == 0) dred[threadIdx.x >> 5] = (double)dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0;
    for (int i = 0; i < 8; ++i) t += dred[i];
    atomicAdd(zdot + img % groups, t);
  }
}

static int check_desc(const tcv_bn_desc& d, const char* who) {
  if (!d.z || d.n <= 0 || d.h <= 0 || d.w <= 0) return fail(TCV_ERR_INVALID, "%s: bad tensor", who);
  if (!(d.c >= 8 && d.c <= 512 && (d.c & (d.c - 1)) == 0))
    return fail(TCV_ERR_INVALID, "%s: c=%d must be a power of two in [8,512]", who, d.c);
  if (d.groups <= 0 || d.n % d.groups) return fail(TCV_ERR_INVALID, "%s: n=%d not a multiple of groups=%d", who, d.n, d.groups);
  if (d.mode != 1 && d.mode != 2) return fail(TCV_ERR_INVALID, "%s: mode must be 1 or 2", who);
  if (d.mode == 2 && (d.res1 || d.res2)) return fail(TCV_ERR_INVALID, "%s: mode 2 takes no residuals", who);
  if (d.act < TCV_ACT_NONE || d.act > TCV_ACT_LEAKY02) return fail(TCV_ERR_INVALID, "%s: activation %d not supported", who, d.act);
  if (!d.mean || !d.invstd || !d.gamma || !d.beta) return fail(TCV_ERR_INVALID, "%s: null statistics / affine pointer", who);
  return TCV_OK;
}

static tcv_bn_desc with_defaults(const tcv_bn_desc* dp) {
  tcv_bn_desc d = *dp;
  const long long full = (long long)d.n * d.h * d.w * d.c;
  if (d.z_plane == 0) d.z_plane = full;
  if (d.y_plane == 0) d.y_plane = full;
  if (d.res2 && d.res2_plane == 0) d.res2_plane = full;
  if (d.res1 && d.res1_plane == 0) d.res1_plane = (long long)d.n * (d.h >> d.res1_shift) * (d.w >> d.res1_shift) * d.c;
  return d;
}

// pixels per block: as many as BN_PIX_MAX, but few enough that the grid covers the 148 SMs ~4 times (the OS8..OS32
// tensors have only a few thousand pixels per image; with a fixed 2048 their grids were 10-40 blocks)
static int bn_pix(const tcv_bn_desc& d) {
  const long long total = (long long)d.n * d.h * d.w;
  long long pix = (total + 4 * 148 - 1) / (4 * 148);
  const int lanes = 256 / (d.c / 8);
  const long long lo = 2LL * lanes;
  if (pix < lo) pix = lo;
  if (pix > BN_PIX_MAX) pix = BN_PIX_MAX;
  return (int)((pix + lanes - 1) / lanes * lanes);
}

static dim3 bn_grid(const tcv_bn_desc& d, int pix) {
  return dim3((unsigned)(((long long)d.h * d.w + pix - 1) / pix), (unsigned)d.n);
}

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_bn_stats(const tcv_bn_desc* dp, double* sums, tcv_stream_t stream) {
  TCV_REQUIRE(dp && sums, "bn_stats: null pointer");
  const tcv_bn_desc d = with_defaults(dp);
  int rc = check_desc(d, "bn_stats");
  if (rc) return rc;
  TCV_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * d.groups * d.c, S(stream)));
  const int pix = bn_pix(d);
  bn_stats_kernel<<<bn_grid(d, pix), 256, 0, S(stream)>>>(d, sums, pix);
  return launched("bn_stats_kernel");
}

int tcv_bn_finalize(const double* sums, double count, double unbiased_count, int groups, int c, float eps,
                    float momentum, float* mean, float* invstd, float* running_mean, float* running_var,
                    tcv_stream_t stream) {
  TCV_REQUIRE(sums && mean && invstd && count > 0 && groups > 0 && c > 0, "bn_finalize: bad arguments");
  bn_finalize_kernel<<<(c + 127) / 128, 128, 0, S(stream)>>>(sums, count, unbiased_count, groups, c, eps, momentum, mean, invstd,
                                                           running_mean, running_var);
  return launched("bn_finalize_kernel");
}

int tcv_bn_apply(const tcv_bn_desc* dp, tcv_stream_t stream) {
  TCV_REQUIRE(dp, "bn_apply: null descriptor");
  const tcv_bn_desc d = with_defaults(dp);
  int rc = check_desc(d, "bn_apply");
  if (rc) return rc;
  TCV_REQUIRE(d.y, "bn_apply: null output");
  const long long total = (long long)d.n * d.h * d.w * (d.c / 8);
  const bool vec = ((reinterpret_cast<uintptr_t>(d.mean) | reinterpret_cast<uintptr_t>(d.invstd) |
                     reinterpret_cast<uintptr_t>(d.gamma) | reinterpret_cast<uintptr_t>(d.beta)) & 15) == 0;
  bn_apply_kernel<<<(unsigned)((total + 255) / 256), 256, 0, S(stream)>>>(d, vec);
  return launched("bn_apply_kernel");
}

int tcv_bn_bwd_reduce(const tcv_bn_desc* dp, const void* dy, long long dy_plane, void* e, long long e_plane,
                      double* sums, tcv_stream_t stream) {
  TCV_REQUIRE(dp && dy && sums, "bn_bwd_reduce: null pointer");
  const tcv_bn_desc d = with_defaults(dp);
  int rc = check_desc(d, "bn_bwd_reduce");
  if (rc) return rc;
  TCV_REQUIRE(d.mode == 2 || e || !d.res1, "bn_bwd_reduce: mode 1 with a residual input needs the e output");
  const long long full = (long long)d.n * d.h * d.w * d.c;
  TCV_CUDA(cudaMemsetAsync(sums, 0, sizeof(double) * 2 * d.groups * d.c, S(stream)));
  const int pix = bn_pix(d);
  bn_bwd_reduce_kernel<<<bn_grid(d, pix), 256, 0, S(stream)>>>(d, reinterpret_cast<const __nv_bfloat16*>(dy),
                                                              dy_plane ? dy_plane : full,
                                                              reinterpret_cast<__nv_bfloat16*>(e),
                                                              e_plane ? e_plane : full, sums, pix);
  return launched("bn_bwd_reduce_kernel");
}

int tcv_bn_param_grads(const double* sums, int groups, int c, float* dgamma, float* dbeta, tcv_stream_t stream) {
  TCV_REQUIRE(sums && dgamma && dbeta, "bn_param_grads: null pointer");
  bn_param_grads_kernel<<<(c + 127) / 128, 128, 0, S(stream)>>>(sums, groups, c, dgamma, dbeta);
  return launched("bn_param_grads_kernel");
}

int tcv_bn_bwd_apply(const tcv_bn_desc* dp, const void* e, long long e_plane, const double* sums, double count,
                     void* dz, long long dz_plane, double* zdot, int e_is_dy, tcv_stream_t stream) {
  TCV_REQUIRE(dp && e && sums && dz && count > 0, "bn_bwd_apply: null pointer");
  const tcv_bn_desc d = with_defaults(dp);
  int rc = check_desc(d, "bn_bwd_apply");
  if (rc) return rc;
  TCV_REQUIRE(!e_is_dy || (d.mode == 1 && !d.res1), "bn_bwd_apply: e_is_dy is for mode 1 without a residual input");
  const long long full = (long long)d.n * d.h * d.w * d.c;
  const int pix = bn_pix(d);
  bn_bwd_apply_kernel<<<bn_grid(d, pix), 256, 0, S(stream)>>>(d, reinterpret_cast<const __nv_bfloat16*>(e),
                                                             e_plane ? e_plane : full, sums, count,
                                                             reinterpret_cast<__nv_bfloat16*>(dz),
                                                             dz_plane ? dz_plane : full, zdot, pix, e_is_dy);
  return launched("bn_bwd_apply_kernel");
}

int tcv_group_dot(const void* z, long long z_plane, const void* dz, long long dz_plane, int n, long long img_elems,
                  int groups, double* zdot, tcv_stream_t stream) {
  TCV_REQUIRE(z && dz && zdot && n > 0 && groups > 0 && img_elems % 8 == 0, "group_dot: bad arguments");
  const long long e8 = img_elems / 8;
  dim3 grid((unsigned)((e8 + 4095) / 4096), (unsigned)n);
  group_dot_kernel<<<grid, 256, 0, S(stream)>>>(reinterpret_cast<const __nv_bfloat16*>(z),
                                               z_plane ? z_plane : (long long)n * img_elems,
                                               reinterpret_cast<const __nv_bfloat16*>(dz),
                                               dz_plane ? dz_plane : (long long)n * img_elems, e8, groups, zdot);
  return launched("group_dot_kernel");
}

}  // extern "C"
