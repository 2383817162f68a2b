// Shared tcgen05 / TMA / mbarrier PTX wrappers and tensor-map helpers for the tensor-core kernels.
#pragma once
#include <cuda.h>
#include <string.h>

#include <mutex>

#include "common.cuh"

namespace tcv {

// ------------------------------------------------------------------------------------------ PTX
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded wait: a protocol bug must not hang the GPU box -- trap after ~2 s instead.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  if (mbar_try_wait(bar, parity)) return;
  const long long t0 = clock64();
  while (!mbar_try_wait(bar, parity)) {
    if (clock64() - t0 > 4000000000LL) {
      printf("tcvom_b200: mbarrier wait timed out (block %d,%d,%d thread %d)\n", blockIdx.x, blockIdx.y, blockIdx.z,
             threadIdx.x);
      __trap();
    }
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* v) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
        "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
        "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
        "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
      : "r"(taddr)
      : "memory");
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major shared-memory matrix descriptor (cute::UMMA::SmemDescriptor layout): rows of BK bf16
// (64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B), 8-row groups SBO apart.
template <int BK>
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  constexpr uint64_t row_bytes = BK * 2;
  constexpr uint64_t sbo = (8 * row_bytes) >> 4;
  constexpr uint64_t layout = row_bytes == 128 ? 2 : 4;  // SWIZZLE_128B : SWIZZLE_64B
  return (uint64_t)((addr & 0x3FFFF) >> 4) | (1ull << 16) | (sbo << 32) | (1ull << 46) | (layout << 61);
}

// The same descriptor in two halves, for issue loops that must stay short: the high word is a compile-time constant and
// the low word is (address >> 4) | LBO, so an operand `off` bytes further on is `lo + (off >> 4)` (shared-memory addresses
// are < 2^18, the sum never carries into the LBO field).
template <int BK>
__device__ __forceinline__ constexpr uint32_t smem_desc_hi() {
  return (uint32_t)((8u * BK * 2u) >> 4) | (1u << 14) | ((BK * 2 == 128 ? 2u : 4u) << 29);
}
__device__ __forceinline__ uint32_t smem_desc_lo(uint32_t addr) { return ((addr & 0x3FFFFu) >> 4) | (1u << 16); }
template <int BK>
__device__ __forceinline__ uint64_t smem_desc_join(uint32_t lo) {
  return ((uint64_t)smem_desc_hi<BK>() << 32) | lo;
}

// kind::f16 instruction descriptor: D=f32, A=B=bf16 (format 1) or fp16 (format 0), both K-major, M=128, N=bn
static inline uint32_t instr_desc(int bn, bool fp16) {
  const uint32_t fmt = fp16 ? 0u : 1u;
  return (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
}


// Programmatic dependent launch (PDL): a kernel launched with cudaLaunchAttributeProgrammaticStreamSerialization may start
// while its predecessor in the stream is still draining.  pdl_launch_dependents() at the top of a kernel lets the NEXT kernel's
// CTAs be scheduled as soon as SMs free up; pdl_wait() blocks until the PREVIOUS kernel has completed and flushed -- it must
// precede the first access to memory the predecessor may have written.  Between the two sits the prologue (barrier init, TMEM
// allocation, tensor-map prefetch, cluster sync), which thereby overlaps the predecessor's tail.  Both are no-ops for a
// kernel launched without the attribute.  Opt-in through tcv_set_debug_flags bit 524288 (measured: no gain inside a CUDA graph).
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// One lane of a fully converged warp.  Issuing TMA / tcgen05 instructions under this predicate (with
// all loop control kept warp-uniform) lets ptxas emit them directly; under a plain `lane == 0` branch
// it wraps every UTCHMMA / UTMALDG in a vote-and-elect serialisation loop (measured: ~16 extra
// instructions per MMA, which made the single issuing thread the bottleneck).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "elect.sync _|p, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(pred));
  return pred != 0;
}

__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

extern std::atomic<int> g_debug_flags;
// launch attributes of the persistent tensor-core kernels: optional cluster of two CTAs, programmatic dependent launch
static inline int tc_launch_attrs(cudaLaunchAttribute* attr, int cluster) {
  int n = 0;
  if (cluster > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster;
    attr[n].val.clusterDim.y = 1;
    attr[n].val.clusterDim.z = 1;
    ++n;
  }
  if (g_debug_flags.load() & 524288) {      // opt-in: measured on B200 (CUDA-graph replay of the window): no gain
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  return n;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static inline EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(f);
  });
  return fn;
}

// bf16 tensor map, dims innermost-first; strides[i] = byte stride of dim i+1; bk: 64 -> SWIZZLE_128B, 0 -> no swizzle,
// anything else -> SWIZZLE_64B
static inline int make_map(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                    const cuuint32_t* box, int bk, bool fp16 = false, int trav = 1) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(TCV_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  // traversal stride on the two spatial dims (W, H) of a 4-D activation map: every `trav`-th pixel
  cuuint32_t estr[5] = {1, (cuuint32_t)trav, (cuuint32_t)trav, 1, 1};
  CUresult r = enc(m, fp16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, rank, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE,
                   bk == 64 ? CU_TENSOR_MAP_SWIZZLE_128B : (bk == 0 ? CU_TENSOR_MAP_SWIZZLE_NONE : CU_TENSOR_MAP_SWIZZLE_64B),
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(TCV_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return TCV_OK;
}


}  // namespace tcv
