// Shared helpers for libdae_sm100.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include "../../include/dae_sm100.h"

namespace dae {

void set_error(const char* fmt, ...);

#define DAE_REQUIRE(cond, ...)                     \
  do {                                             \
    if (!(cond)) {                                 \
      dae::set_error(__VA_ARGS__);                 \
      return DAE_ERR_BAD_ARG;                      \
    }                                              \
  } while (0)

#define DAE_CHECK_LAUNCH(name)                                              \
  do {                                                                      \
    cudaError_t e__ = cudaGetLastError();                                   \
    if (e__ != cudaSuccess) {                                               \
      dae::set_error("%s: %s", name, cudaGetErrorString(e__));              \
      return DAE_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

#define DAE_CUDA(call)                                                      \
  do {                                                                      \
    cudaError_t e__ = (call);                                               \
    if (e__ != cudaSuccess) {                                               \
      dae::set_error("%s: %s", #call, cudaGetErrorString(e__));             \
      return DAE_ERR_CUDA;                                                  \
    }                                                                       \
  } while (0)

constexpr float kEps = 1e-16f;

template <int ACT>
__device__ __forceinline__ float act_fwd(float x) {
  if (ACT == DAE_ACT_SIGMOID) return 1.0f / (1.0f + expf(-x));
  if (ACT == DAE_ACT_TANH) return tanhf(x);
  return x;
}
// derivative expressed through the activation value y = f(x)
template <int ACT>
__device__ __forceinline__ float act_grad_from_y(float y) {
  if (ACT == DAE_ACT_SIGMOID) return y * (1.0f - y);
  if (ACT == DAE_ACT_TANH) return 1.0f - y * y;
  return 1.0f;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result valid in thread 0 (and broadcast through smem to all). blockDim <= 1024.
template <typename T>
__device__ __forceinline__ T block_sum(T v, T* smem32) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) smem32[wid] = v;
  __syncthreads();
  const int nw = (blockDim.x + 31) >> 5;
  T r = (threadIdx.x < nw) ? smem32[threadIdx.x] : T(0);
  if (wid == 0) {
    r = warp_sum(r);
    if (lane == 0) smem32[0] = r;
  }
  __syncthreads();
  return smem32[0];
}

#define DAE_DISPATCH_ACT(act, ACT, ...)                      \
  switch (act) {                                             \
    case DAE_ACT_SIGMOID: { constexpr int ACT = DAE_ACT_SIGMOID; __VA_ARGS__; } break; \
    case DAE_ACT_TANH:    { constexpr int ACT = DAE_ACT_TANH;    __VA_ARGS__; } break; \
    default:              { constexpr int ACT = DAE_ACT_NONE;    __VA_ARGS__; } break; \
  }

}  // namespace dae
