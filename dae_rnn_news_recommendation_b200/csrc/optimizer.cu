// K6: fused optimizer over the flat parameter buffer, and the masking-noise kernel.
//
// Reference ops replaced: tf.train.{GradientDescent,Adagrad,Momentum,Adam}Optimizer.minimize's apply step
// (autoencoder/autoencoder.py:451-472, TF-1.12 update rules) and utils.masking_noise (autoencoder/utils.py:94-115).
#include <cuda_bf16.h>
#include "common.cuh"

namespace dae {

template <int OPT>
__global__ void __launch_bounds__(256) optimizer_kernel(float* __restrict__ theta, const float* __restrict__ grad,
                                                        float* __restrict__ slot1, float* __restrict__ slot2, int64_t n, float lr,
                                                        float momentum, float gscale, float lr_t, __nv_bfloat16* __restrict__ w_hi,
                                                        __nv_bfloat16* __restrict__ w_lo, int64_t n_w, int H, int64_t ld_split,
                                                        const int64_t* __restrict__ ctl) {
  if (OPT == DAE_OPT_ADAM && ctl) {  // device-resident step counter (CUDA-graph replay): lr_t = lr sqrt(1-b2^t)/(1-b1^t)
    const double t = (double)ctl[2];
    lr_t = (float)((double)lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
  }
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float g = grad[i] * gscale;
    float p = theta[i];
    if (OPT == DAE_OPT_SGD) {
      p -= lr * g;
    } else if (OPT == DAE_OPT_ADAGRAD) {   // accum += g^2 ; var -= lr * g * rsqrt(accum)   (initial accum 0.1, no epsilon)
      const float a = slot1[i] + g * g;
      slot1[i] = a;
      p -= lr * g / sqrtf(a);
    } else if (OPT == DAE_OPT_MOMENTUM) {  // accum = mu * accum + g ; var -= lr * accum
      const float a = momentum * slot1[i] + g;
      slot1[i] = a;
      p -= lr * a;
    } else {                               // Adam: m, v; var -= lr_t * m / (sqrt(v) + 1e-8)
      const float m = 0.9f * slot1[i] + (1.0f - 0.9f) * g;
      const float v = 0.999f * slot2[i] + (1.0f - 0.999f) * g * g;
      slot1[i] = m;
      slot2[i] = v;
      p -= lr_t * m / (sqrtf(v) + 1e-8f);
    }
    theta[i] = p;
    if (w_hi != nullptr && i < n_w) {  // refresh the bf16 hi/lo operand copy of W consumed by the tensor-core GEMMs
      const int64_t r = i / H;
      const int64_t o = r * ld_split + (i - r * H);
      const __nv_bfloat16 h = __float2bfloat16_rn(p);
      w_hi[o] = h;
      w_lo[o] = __float2bfloat16_rn(p - __bfloat162float(h));
    }
  }
}

// Philox4x32-10 (Salmon et al. 2011)
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  const uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  const uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  const uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__global__ void __launch_bounds__(256) mask_values_kernel(const float* __restrict__ values, const uint8_t* __restrict__ keep,
                                                          int64_t nnz, float corr_frac, uint64_t seed, uint64_t epoch,
                                                          float* __restrict__ out) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; q * 4 < nnz; q += stride) {
    float u[4];
    if (!keep) {
      uint32_t c[4] = {(uint32_t)q, (uint32_t)(q >> 32), (uint32_t)epoch, (uint32_t)(epoch >> 32)};
      uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
      for (int r = 0; r < 10; ++r) philox_round(c, k);
#pragma unroll
      for (int e = 0; e < 4; ++e) u[e] = (float)(c[e] >> 8) * (1.0f / 16777216.0f);  // [0,1)
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int64_t p = q * 4 + e;
      if (p < nnz) {
        const bool k1 = keep ? (keep[p] != 0) : (u[e] >= corr_frac);  // np.random.rand(nnz) >= v  (utils.py:111)
        out[p] = k1 ? values[p] : 0.0f;
      }
    }
  }
}

}  // namespace dae

extern "C" int dae_optimizer_step(float* theta, const float* grad, float* slot1, float* slot2, int64_t n, int32_t opt, float lr,
                                  float momentum, float grad_scale, int32_t step, const int64_t* ctl, void* w_hi, void* w_lo,
                                  int32_t F, int32_t H, int64_t ld_split, void* stream) {
  using namespace dae;
  DAE_REQUIRE(theta && grad && n > 0, "dae_optimizer_step: bad arguments");
  DAE_REQUIRE(opt == DAE_OPT_SGD || slot1, "dae_optimizer_step: slot1 required");
  DAE_REQUIRE(opt != DAE_OPT_ADAM || slot2, "dae_optimizer_step: slot2 required for adam");
  DAE_REQUIRE(!w_hi || (w_lo && F > 0 && H > 0 && ld_split >= H && (int64_t)F * H <= n), "dae_optimizer_step: bad split arguments");
  cudaStream_t st = (cudaStream_t)stream;
  __nv_bfloat16* wh = (__nv_bfloat16*)w_hi; __nv_bfloat16* wl = (__nv_bfloat16*)w_lo;
  const int64_t n_w = w_hi ? (int64_t)F * H : 0;
  const int blocks = (int)((n + 255) / 256 < 148 * 8 ? (n + 255) / 256 : 148 * 8);
  float lr_t = lr;
  if (opt == DAE_OPT_ADAM) {
    const double t = (double)(step < 1 ? 1 : step);
    lr_t = (float)((double)lr * sqrt(1.0 - pow(0.999, t)) / (1.0 - pow(0.9, t)));
  }
  switch (opt) {
    case DAE_OPT_SGD: optimizer_kernel<DAE_OPT_SGD><<<blocks, 256, 0, st>>>(theta, grad, slot1, slot2, n, lr, momentum, grad_scale, lr_t, wh, wl, n_w, H, ld_split, ctl); break;
    case DAE_OPT_ADAGRAD: optimizer_kernel<DAE_OPT_ADAGRAD><<<blocks, 256, 0, st>>>(theta, grad, slot1, slot2, n, lr, momentum, grad_scale, lr_t, wh, wl, n_w, H, ld_split, ctl); break;
    case DAE_OPT_MOMENTUM: optimizer_kernel<DAE_OPT_MOMENTUM><<<blocks, 256, 0, st>>>(theta, grad, slot1, slot2, n, lr, momentum, grad_scale, lr_t, wh, wl, n_w, H, ld_split, ctl); break;
    case DAE_OPT_ADAM: optimizer_kernel<DAE_OPT_ADAM><<<blocks, 256, 0, st>>>(theta, grad, slot1, slot2, n, lr, momentum, grad_scale, lr_t, wh, wl, n_w, H, ld_split, ctl); break;
    default: set_error("dae_optimizer_step: unknown optimizer %d", opt); return DAE_ERR_BAD_ARG;
  }
  DAE_CHECK_LAUNCH("dae_optimizer_step");
  return DAE_OK;
}

extern "C" int dae_mask_values(const float* values, const uint8_t* keep, int64_t nnz, float corr_frac, uint64_t seed, uint64_t epoch,
                               float* values_out, void* stream) {
  using namespace dae;
  DAE_REQUIRE(values && values_out && nnz >= 0, "dae_mask_values: bad arguments");
  if (nnz == 0) return DAE_OK;
  const int64_t quads = (nnz + 3) / 4;
  const int blocks = (int)((quads + 255) / 256 < 148 * 16 ? (quads + 255) / 256 : 148 * 16);
  mask_values_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(values, keep, nnz, corr_frac, seed, epoch, values_out);
  DAE_CHECK_LAUNCH("dae_mask_values");
  return DAE_OK;
}
