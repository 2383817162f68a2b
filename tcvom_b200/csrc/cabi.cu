// C-ABI plumbing: error reporting, launch accounting and the tcv_conv2d dispatcher.
#include <stdarg.h>
#include <string.h>

#include "common.cuh"

namespace tcv {

thread_local char g_err[512] = "";
std::atomic<long long> g_launches{0};

int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int launched(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return fail(TCV_ERR_CUDA, "%s launch failed: %s", what, cudaGetErrorString(e));
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return TCV_OK;
}

int conv2d_direct(const tcv_conv_desc& d, cudaStream_t st);
int conv2d_tc_supported(const tcv_conv_desc& d);
int conv2d_tc(const tcv_conv_desc& d, cudaStream_t st);
int conv2d_tc2_supported(const tcv_conv_desc& d);
int conv2d_tc2(const tcv_conv_desc& d, cudaStream_t st);
int conv2d_tc2p_supported(const tcv_conv_desc& d);
int conv2d_tc2p(const tcv_conv_desc& d, cudaStream_t st);
int conv2d_tc3_supported(const tcv_conv_desc& d);
int conv2d_tc3(const tcv_conv_desc& d, cudaStream_t st);
std::atomic<int> g_conv_tc_version{3};
std::atomic<int> g_debug_flags{0};

}  // namespace tcv

using namespace tcv;

extern "C" {

int tcv_version(void) { return 100; }
const char* tcv_last_error(void) { return g_err; }
long long tcv_launch_count(void) { return g_launches.load(); }

int tcv_conv2d_path(const tcv_conv_desc* dp) {
  if (!dp) return -1;
  tcv_conv_desc d = *dp;
  if (d.x_plane == 0) d.x_plane = (long long)d.n * d.ih * d.iw * d.cin;
  if (d.x_img_stride == 0) d.x_img_stride = (long long)d.ih * d.iw * d.cin;
  if (g_conv_tc_version.load() >= 3 && conv2d_tc3_supported(d)) return 3;
  if (g_conv_tc_version.load() >= 2 && !(g_debug_flags.load() & 4096) && conv2d_tc2p_supported(d)) return 4;
  if (g_conv_tc_version.load() >= 2 && conv2d_tc2_supported(d)) return 2;
  return conv2d_tc_supported(d) ? 1 : 0;
}

int tcv_set_debug_flags(int flags) { return g_debug_flags.exchange(flags); }

int tcv_set_conv_tc_version(int v) {
  const int old = g_conv_tc_version.exchange(v);
  return old;
}

int tcv_conv2d(const tcv_conv_desc* dp, tcv_stream_t stream) {
  TCV_REQUIRE(dp, "conv2d: null descriptor");
  tcv_conv_desc d = *dp;
  TCV_REQUIRE(d.x && d.w && (d.y || d.y_f32), "conv2d: null tensor pointer");
  if (d.x_plane == 0) d.x_plane = (long long)d.n * d.ih * d.iw * d.cin;
  if (d.x_img_stride == 0) d.x_img_stride = (long long)d.ih * d.iw * d.cin;
  TCV_REQUIRE(d.n > 0 && d.ih > 0 && d.iw > 0 && d.oh > 0 && d.ow > 0 && d.gh > 0 && d.gw > 0, "conv2d: bad dims");
  TCV_REQUIRE(d.cin % 8 == 0 && (d.cin <= 32 || d.cin % 32 == 0), "conv2d: cin=%d must be 8,16,24,32 or a multiple of 32", d.cin);
  TCV_REQUIRE(d.cout >= 1, "conv2d: bad cout");
  TCV_REQUIRE(d.ntaps >= 1 && d.ntaps <= TCV_MAX_TAPS, "conv2d: ntaps=%d out of range", d.ntaps);
  TCV_REQUIRE(d.stride == 1 || d.stride == 2, "conv2d: stride must be 1 or 2");
  TCV_REQUIRE(d.act >= TCV_ACT_NONE && d.act <= TCV_ACT_RELU6, "conv2d: unknown activation %d", d.act);
  TCV_REQUIRE((d.s2 == nullptr) == (d.b2 == nullptr), "conv2d: s2 and b2 go together");
  TCV_REQUIRE((d.gh - 1) * d.oy_mul + d.oy_off < d.oh && (d.gw - 1) * d.ox_mul + d.ox_off < d.ow,
              "conv2d: compute grid does not fit the output tensor");
  if (d.pad_mode == TCV_PAD_REFLECT) {
    for (int t = 0; t < d.ntaps; ++t) {
      const int ymin = d.dy[t], ymax = (d.gh - 1) * d.stride + d.dy[t];
      const int xmin = d.dx[t], xmax = (d.gw - 1) * d.stride + d.dx[t];
      TCV_REQUIRE(ymin > -d.ih && ymax < 2 * d.ih - 1 && xmin > -d.iw && xmax < 2 * d.iw - 1,
                  "conv2d: reflect padding wider than the image");
    }
  }
  if (d.x_plane == 0) d.x_plane = (long long)d.n * d.ih * d.iw * d.cin;
  if (d.x_img_stride == 0) d.x_img_stride = (long long)d.ih * d.iw * d.cin;
  if (d.res1 && d.res1_plane == 0)
    d.res1_plane = (long long)d.n * (d.oh >> d.res1_shift) * (d.ow >> d.res1_shift) * d.cout;
  if (d.res2 && d.res2_plane == 0) d.res2_plane = (long long)d.n * d.oh * d.ow * d.cout;
  if (d.stats) {
    TCV_REQUIRE(g_conv_tc_version.load() >= 2 && !(g_debug_flags.load() & 4096) && conv2d_tc2p_supported(d),
                "conv2d: output statistics need the CTA-pair tcgen05 kernel (tcv_conv2d_path 4)");
    return conv2d_tc2p(d, S(stream));
  }
  if (g_conv_tc_version.load() >= 3 && conv2d_tc3_supported(d)) return conv2d_tc3(d, S(stream));
  // wide layers (Cout >= 128) on CTA pairs (conv_tc2p.cu); tcv_set_debug_flags bit 4096 falls back to the single-CTA kernel
  if (g_conv_tc_version.load() >= 2 && !(g_debug_flags.load() & 4096) && conv2d_tc2p_supported(d))
    return conv2d_tc2p(d, S(stream));
  if (g_conv_tc_version.load() >= 2 && conv2d_tc2_supported(d)) return conv2d_tc2(d, S(stream));
  if (conv2d_tc_supported(d)) return conv2d_tc(d, S(stream));
  return conv2d_direct(d, S(stream));
}

int tcv_zero_bytes(void* p, long long bytes, tcv_stream_t stream) {
  TCV_REQUIRE(p && bytes > 0, "zero_bytes: bad arguments");
  TCV_CUDA(cudaMemsetAsync(p, 0, (size_t)bytes, S(stream)));
  return launched("zero_bytes");
}

}  // extern "C"
