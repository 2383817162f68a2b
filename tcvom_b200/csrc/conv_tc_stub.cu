// placeholder until the tcgen05 implicit-GEMM path lands
#include "common.cuh"
namespace tcv {
int conv2d_tc_supported(const tcv_conv_desc&) { return 0; }
int conv2d_tc(const tcv_conv_desc&, cudaStream_t) { return fail(TCV_ERR_UNSUPPORTED, "tc conv not built"); }
}
