// SyncBatchNorm statistic exchange over NVLink peer memory (train_ddp.py:273 nn.SyncBatchNorm: one tiny reduction per
// BatchNorm layer in the forward and one in the backward pass, <= 5 x 512 x 2 doubles each, 144 per training step).
// NCCL needs ~25 us per such call (its own stream, two event hand-offs, a latency-bound ring / tree); this kernel does
// the whole all-reduce in ONE launch on the compute stream:
//   every rank owns a symmetric buffer (same layout on every GPU, mapped into every peer):  [flags 2 x 16 x u64][slot 0][slot 1]
//   1. publish: copy the local values into the own slot (epoch parity), __threadfence_system
//   2. signal:  thread r stores `epoch` into rank r's flag word [slot][my rank]            (remote store over NVLink)
//   3. wait:    thread r spins on the own flag word [slot][r] until it holds `epoch`        (local loads)
//   4. reduce:  every element is summed over the ranks' slots IN RANK ORDER                 (remote loads) -> identical bits
//               on every rank, and run-to-run deterministic
// Two slots suffice: a rank can start call k+1 only after every peer has arrived at call k, i.e. finished call k-1, so slot
// (k+1) % 2 is no longer read by anybody.  Bounded spin (~2 s) traps instead of hanging the GPU if a peer never arrives.
#include "common.cuh"

namespace tcv {
constexpr int PR_MAXW = 16;                 // ranks per node supported by the flag block
constexpr int PR_FLAG_BYTES = 2 * PR_MAXW * 8;

__global__ void __launch_bounds__(1024) peer_allreduce_f64_kernel(double* __restrict__ data, int count,
                                                                  void* const* __restrict__ peers, int rank, int world,
                                                                  unsigned long long epoch, long long slot_doubles) {
  const int slot = (int)(epoch & 1ull);
  auto slot_of = [&](int r) {
    return reinterpret_cast<double*>(reinterpret_cast<char*>(peers[r]) + PR_FLAG_BYTES) + (long long)slot * slot_doubles;
  };
  double* mine = slot_of(rank);
  for (int i = threadIdx.x; i < count; i += blockDim.x) mine[i] = data[i];
  __threadfence_system();
  __syncthreads();
  if ((int)threadIdx.x < world) {
    unsigned long long* remote = reinterpret_cast<unsigned long long*>(peers[threadIdx.x]) + slot * PR_MAXW + rank;
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(remote), "l"(epoch) : "memory");
    const unsigned long long* local = reinterpret_cast<const unsigned long long*>(peers[rank]) + slot * PR_MAXW + threadIdx.x;
    unsigned long long v = 0;
    const long long t0 = clock64();
    while (true) {
      asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(local) : "memory");
      if (v >= epoch) break;
      if (clock64() - t0 > 4000000000LL) {
        printf("tcvom_b200: peer all-reduce timed out waiting for rank %d (epoch %llu, have %llu)\n", (int)threadIdx.x, epoch, v);
        __trap();
      }
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < count; i += blockDim.x) {
    double s = 0.0;
    for (int r = 0; r < world; ++r) {
      double v;
      asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(slot_of(r) + i) : "memory");
      s += v;
    }
    data[i] = s;
  }
}
}  // namespace tcv

extern "C" int tcv_peer_allreduce_f64(double* data, int count, void* const* peers_dev, int rank, int world,
                                      unsigned long long epoch, long long slot_doubles, tcv_stream_t stream) {
  using namespace tcv;
  TCV_REQUIRE(data && peers_dev, "peer_allreduce_f64: null pointer");
  TCV_REQUIRE(world >= 1 && world <= PR_MAXW && rank >= 0 && rank < world, "peer_allreduce_f64: world must be 1..16");
  TCV_REQUIRE(count > 0 && count <= slot_doubles, "peer_allreduce_f64: %d values do not fit a slot of %lld", count, slot_doubles);
  TCV_REQUIRE(epoch > 0, "peer_allreduce_f64: epochs start at 1");
  peer_allreduce_f64_kernel<<<1, 1024, 0, S(stream)>>>(data, count, peers_dev, rank, world, epoch, slot_doubles);
  return launched("peer_allreduce_f64_kernel");
}

extern "C" int tcv_peer_buffer_bytes(long long slot_doubles, long long* bytes) {
  TCV_REQUIRE(bytes && slot_doubles > 0, "peer_buffer_bytes: bad arguments");
  *bytes = tcv::PR_FLAG_BYTES + 2 * slot_doubles * 8;
  return TCV_OK;
}
