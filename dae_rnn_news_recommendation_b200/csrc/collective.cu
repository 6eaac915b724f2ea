// Data-parallel gradient all-reduce over NVSwitch multicast memory ("one all-reduce of the flat [dW|dbh|dbv] buffer").
//
// Reference op replaced: none -- the reference is single-process; this is the exchange step the row-sharded training adds between
// the gradient kernels and dae_optimizer_step.  The NCCL all-reduce (torch.distributed) is the default; this kernel is the
// in-graph alternative (TrainEngine, DAE_ALLREDUCE=multimem): the gradient buffer lives in symmetric memory that every rank maps,
// bound to one multicast address, and each rank reduces 1/P of it IN THE SWITCH and broadcasts the sums:
//
//   barrier (all ranks' gradients complete)  ->  multimem.ld_reduce.add.v4.f32 on my slice (the switch pulls the P copies and adds)
//   ->  multimem.st.v4.f32 of the sums (the switch writes all P copies)  ->  barrier (all sums landed everywhere)
//
// Per GPU that is N(P-1)/P bytes out and in over NVLink, no staging buffer, no intermediate HBM round trip, and -- being an
// ordinary kernel on the step's stream -- it is captured inside the step's CUDA graph (NCCL's launch has to stay outside).
// Barriers are per CTA: CTA b of rank r writes the exchange's epoch into flag[b][r] on every peer (one remote store each) and polls its
// own flag[b][0..P-1] in local memory; the epoch is a per-CTA device counter the kernel advances itself, so it stays in step with graph
// replays.
#include "common.cuh"

namespace dae {

__device__ __forceinline__ void st_release_sys(unsigned* addr, unsigned value) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(addr), "r"(value) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* addr) {
  unsigned v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(addr) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long now_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
// A peer that never arrives (crashed rank, mismatched launch) must not leave this GPU spinning forever: after 5 s the kernel traps
// and the error surfaces on the host at the next synchronisation.
constexpr unsigned long long kBarrierTimeoutNs = 5000000000ull;

// flags[p] = base of rank p's flag words (peer-mapped); word index = (phase * n_blocks + block) * world + source rank.
// A flag carries the EPOCH of the exchange (a per-CTA counter kept in local memory and advanced by the kernel itself, so it stays in
// step with graph replays): arriving is ONE fire-and-forget remote store per peer, waiting polls local memory -- no read-modify-write
// round trips over NVLink and nothing to re-arm.
__device__ __forceinline__ void peer_barrier(unsigned* const* flags, int rank, int world, int phase, unsigned epoch) {
  __syncthreads();
  if ((int)threadIdx.x < world) {
    const int64_t slot = ((int64_t)phase * gridDim.x + blockIdx.x) * world;
    __threadfence_system();
    st_release_sys(flags[threadIdx.x] + slot + rank, epoch);      // "rank `rank`, CTA b has reached `epoch`" on peer threadIdx.x
    const unsigned* mine = flags[rank] + slot + threadIdx.x;       // wait for peer threadIdx.x's CTA b
    const unsigned long long t0 = now_ns();
    while ((int)(ld_acquire_sys(mine) - epoch) < 0) { if (now_ns() - t0 > kBarrierTimeoutNs) __trap(); }
  }
  __syncthreads();
}

__global__ void __launch_bounds__(512) allreduce_multimem_kernel(float* __restrict__ mc, unsigned* const* __restrict__ flags,
                                                                 unsigned* __restrict__ epochs, int rank, int world, int64_t n) {
  const unsigned epoch = epochs[blockIdx.x] + 1u;     // every rank runs the same sequence of exchanges: the counters agree
  peer_barrier(flags, rank, world, 0, epoch);
  const int64_t n4 = n >> 2;                                 // whole float4 packets; the <= 3 trailing floats go to rank 0
  const int64_t per = (n4 + world - 1) / world;
  const int64_t lo = (int64_t)rank * per;
  const int64_t hi = (lo + per < n4) ? lo + per : n4;
  // kUnroll independent in-switch reductions in flight per thread (each is a round trip GPU -> switch -> P GPUs -> switch -> GPU),
  // then their broadcasts
  constexpr int kUnroll = 4;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i0 = lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i0 < hi; i0 += stride * kUnroll) {
    float v[kUnroll][4];
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int64_t i = i0 + k * stride;
      if (i < hi)
        asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0, %1, %2, %3}, [%4];"
                     : "=f"(v[k][0]), "=f"(v[k][1]), "=f"(v[k][2]), "=f"(v[k][3]) : "l"(mc + (i << 2)) : "memory");
    }
#pragma unroll
    for (int k = 0; k < kUnroll; ++k) {
      const int64_t i = i0 + k * stride;
      if (i < hi)
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(mc + (i << 2)), "f"(v[k][0]), "f"(v[k][1]), "f"(v[k][2]),
                     "f"(v[k][3]) : "memory");
    }
  }
  if (rank == 0 && blockIdx.x == 0 && (n4 << 2) + (int64_t)threadIdx.x < n) {
    float* p = mc + (n4 << 2) + threadIdx.x;
    float a;
    asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(a) : "l"(p) : "memory");
    asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(p), "f"(a) : "memory");
  }
  peer_barrier(flags, rank, world, 1, epoch);
  if (threadIdx.x == 0) epochs[blockIdx.x] = epoch;
}

}  // namespace dae

extern "C" int dae_allreduce_multimem(float* multicast_grad, void* const* peer_flags, uint32_t* epochs, int32_t rank, int32_t world,
                                      int64_t n, int32_t n_blocks, void* stream) {
  using namespace dae;
  DAE_REQUIRE(multicast_grad && peer_flags && epochs && world >= 2 && world <= 32 && rank >= 0 && rank < world && n > 0 && n_blocks > 0 && n_blocks <= 1024,
              "dae_allreduce_multimem: bad arguments (world=%d rank=%d n=%lld blocks=%d)", world, rank, (long long)n, n_blocks);
  DAE_REQUIRE((reinterpret_cast<uintptr_t>(multicast_grad) & 15) == 0, "dae_allreduce_multimem: the multicast buffer must be 16-byte aligned");
  allreduce_multimem_kernel<<<n_blocks, 512, 0, (cudaStream_t)stream>>>(multicast_grad, (unsigned* const*)peer_flags, epochs, rank, world, n);
  DAE_CHECK_LAUNCH("dae_allreduce_multimem");
  return DAE_OK;
}
