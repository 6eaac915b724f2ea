// Reconstruction loss + dZ (elementwise part of the decode step), column sums, and the per-step scalar epilogue.
//
// Reference ops replaced: tf.sparse.to_dense(input) and the weighted per-row CE / MSE / cosine loss
// (autoencoder/triplet_loss_utils.py:262-277) on D = g(E.W^T + bv) (autoencoder/autoencoder.py:411), plus their autodiff.
// The clean target row is densified on the fly in shared memory; dense X is never materialised in HBM.
#include "common.cuh"

namespace dae {

constexpr int kLossThreads = 256;
constexpr int kFChunk = 8192;  // floats of the densified target row kept in smem at a time

template <int ACT, int LOSS>
__global__ void __launch_bounds__(kLossThreads) decode_loss_kernel(
    const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices, const float* __restrict__ values,
    const int32_t* __restrict__ rows, int F, const float* __restrict__ bv, const float* __restrict__ weight,
    const double* __restrict__ stats, float* __restrict__ Z, int64_t ldz, float* __restrict__ row_loss) {
  __shared__ float xs[kFChunk];
  __shared__ float red[32];
  const int tid = threadIdx.x;
  const int r = blockIdx.x;
  const int64_t row = rows ? (int64_t)rows[r] : (int64_t)r;
  const int64_t p0 = indptr[row], p1 = indptr[row + 1];
  const float sum_w = (float)stats[DAE_STAT_SUM_W];
  const float sc = (weight ? weight[r] : 1.0f) / (sum_w + kEps);  // d(L_ae)/d(l_r)   (triplet_loss_utils.py:275)
  float* z = Z + (int64_t)r * ldz;

  float rx = 0.0f, rd = 0.0f, sxd = 0.0f, rd3 = 0.0f;
  if (LOSS == DAE_LOSS_COSINE) {
    // pass 0: sum x^2 (CSR), sum d^2, sum x.d   (tf.nn.l2_normalize: x * rsqrt(max(sum x^2, 1e-12)))
    float sxx = 0.0f, sdd = 0.0f, sxd_p = 0.0f;
    for (int64_t p = p0 + tid; p < p1; p += kLossThreads) { const float v = values[p]; sxx += v * v; }
    for (int f0 = 0; f0 < F; f0 += kFChunk) {
      const int fn = min(kFChunk, F - f0);
      for (int f = tid; f < fn; f += kLossThreads) xs[f] = 0.0f;
      __syncthreads();
      for (int64_t p = p0 + tid; p < p1; p += kLossThreads) { const int c = indices[p] - f0; if (c >= 0 && c < fn) xs[c] = values[p]; }
      __syncthreads();
      for (int f = tid; f < fn; f += kLossThreads) {
        const float d = act_fwd<ACT>(z[f0 + f] + __ldg(bv + f0 + f));
        sdd += d * d;
        sxd_p += xs[f] * d;
      }
      __syncthreads();
    }
    sxx = block_sum(sxx, red);
    sdd = block_sum(sdd, red);
    sxd = block_sum(sxd_p, red);
    rx = rsqrtf(fmaxf(sxx, 1e-12f));
    rd = rsqrtf(fmaxf(sdd, 1e-12f));
    rd3 = (sdd >= 1e-12f) ? rd * rd * rd : 0.0f;  // clamped branch of max() has zero gradient
  }

  float lsum = 0.0f;
  for (int f0 = 0; f0 < F; f0 += kFChunk) {
    const int fn = min(kFChunk, F - f0);
    for (int f = tid; f < fn; f += kLossThreads) xs[f] = 0.0f;
    __syncthreads();
    for (int64_t p = p0 + tid; p < p1; p += kLossThreads) { const int c = indices[p] - f0; if (c >= 0 && c < fn) xs[c] = values[p]; }
    __syncthreads();
    for (int f = tid; f < fn; f += kLossThreads) {
      const float x = xs[f];
      const float d = act_fwd<ACT>(z[f0 + f] + __ldg(bv + f0 + f));
      const float gp = act_grad_from_y<ACT>(d);
      float dl;  // dl_r / dD
      if (LOSS == DAE_LOSS_CE) {
        const float a = d + kEps;           // decode + 1e-16
        const float b = (1.0f - d) + kEps;  // 1. - decode + 1e-16, left to right (triplet_loss_utils.py:269)
        lsum -= x * logf(a) + (1.0f - x) * logf(b);
        dl = -(x / a - (1.0f - x) / b);
      } else if (LOSS == DAE_LOSS_MSE) {
        const float e = x - d;
        lsum += e * e;
        dl = -2.0f * e;
      } else {
        dl = -rx * (x * rd - sxd * rd3 * d);
      }
      z[f0 + f] = sc * dl * gp;
    }
    __syncthreads();
  }
  if (LOSS == DAE_LOSS_COSINE) {
    if (tid == 0) row_loss[r] = -sxd * rx * rd;
  } else {
    lsum = block_sum(lsum, red);
    if (tid == 0) row_loss[r] = lsum;
  }
}

// out[f] = sum_r M[r, f]; grid.x over column blocks of 128, grid.y over row slabs; atomics across slabs.
__global__ void colsum_kernel(const float* __restrict__ M, int n_rows, int n_cols, int64_t ld, int rows_per_slab,
                              float* __restrict__ out) {
  const int f = blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= n_cols) return;
  const int r0 = blockIdx.y * rows_per_slab, r1 = min(n_rows, r0 + rows_per_slab);
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  int r = r0;
  for (; r + 3 < r1; r += 4) {
    s0 += M[(int64_t)r * ld + f];
    s1 += M[(int64_t)(r + 1) * ld + f];
    s2 += M[(int64_t)(r + 2) * ld + f];
    s3 += M[(int64_t)(r + 3) * ld + f];
  }
  for (; r < r1; ++r) s0 += M[(int64_t)r * ld + f];
  atomicAdd(out + f, (s0 + s1) + (s2 + s3));
}

__global__ void reduce_parts_kernel(const float* __restrict__ parts, int n_parts, int n, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.0f;
  for (int p = 0; p < n_parts; ++p) s += parts[(int64_t)p * n + i];
  out[i] = s;
}

// Single CTA: deterministic reduction of the weighted row losses + the step's scalars.
__global__ void __launch_bounds__(1024) step_finalize_kernel(const float* __restrict__ row_loss, const float* __restrict__ parts,
                                                             int n_parts, const float* __restrict__ weight, int B, int strategy,
                                                             float alpha, double* __restrict__ stats, double* __restrict__ stats_log,
                                                             const int64_t* __restrict__ ctl) {
  if (stats_log && ctl) stats_log += ctl[1] * DAE_STAT_SLOTS;  // device-resident log cursor (CUDA-graph replay)
  __shared__ double red[32];
  double s = 0.0;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float l;
    if (parts) {  // per-tile partial row losses of the fused decode epilogue, summed in a fixed order
      l = 0.0f;
#pragma unroll 8
      for (int q = 0; q < n_parts; ++q) l += parts[(int64_t)q * B + i];
    } else {
      l = row_loss[i];
    }
    s += (double)l * (double)(weight ? weight[i] : 1.0f);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) {
    const double sum_w = stats[DAE_STAT_SUM_W];
    const double ae = s / (sum_w + 1e-16);
    double tl = 0.0, frac = 0.0, num = 0.0;
    if (strategy == DAE_TRIPLET_BATCH_ALL) {
      const double nv = stats[DAE_STAT_N_VALID];
      tl = stats[DAE_STAT_TRIPLET_SUM] / (nv + 1e-16);
      num = stats[DAE_STAT_NUM];
      frac = num / (nv + 1e-16);
    } else if (strategy == DAE_TRIPLET_BATCH_HARD) {
      const double na = stats[DAE_STAT_N_ACTIVE];
      tl = stats[DAE_STAT_TRIPLET_SUM] / (na + 1e-16);
      num = na;
      frac = na / (double)B;
    } else if (strategy == 3) {  // explicit triplets: mean over the N_ACTIVE = B triples
      tl = stats[DAE_STAT_TRIPLET_SUM] / stats[DAE_STAT_N_ACTIVE];
    }
    stats[DAE_STAT_SUM_LW] = s;
    stats[DAE_STAT_AE_LOSS] = ae;
    stats[DAE_STAT_TRIPLET_LOSS] = tl;
    stats[DAE_STAT_FRACTION] = frac;
    stats[DAE_STAT_NUM] = num;
    stats[DAE_STAT_COST] = (strategy == DAE_TRIPLET_NONE) ? ae : ae + (double)alpha * tl;
  }
  __syncthreads();
  if (stats_log && threadIdx.x < DAE_STAT_SLOTS) stats_log[threadIdx.x] = stats[threadIdx.x];
}

}  // namespace dae

extern "C" int dae_decode_loss_bwd(const int64_t* indptr, const int32_t* indices, const float* values, const int32_t* rows,
                                   int32_t n_rows, int32_t F, const float* bv, int32_t dec_act, int32_t loss_func,
                                   const float* weight, const double* stats, float* Z, int64_t ldz, float* row_loss, void* stream) {
  using namespace dae;
  DAE_REQUIRE(indptr && indices && values && bv && stats && Z && row_loss, "dae_decode_loss_bwd: null pointer");
  DAE_REQUIRE(n_rows >= 0 && F > 0 && ldz >= F, "dae_decode_loss_bwd: bad shape");
  DAE_REQUIRE(loss_func >= 0 && loss_func <= 2, "dae_decode_loss_bwd: unknown loss %d", loss_func);
  if (n_rows == 0) return DAE_OK;
  cudaStream_t st = (cudaStream_t)stream;
#define DAE_LAUNCH_LOSS(ACT, LOSS) \
  decode_loss_kernel<ACT, LOSS><<<n_rows, kLossThreads, 0, st>>>(indptr, indices, values, rows, F, bv, weight, stats, Z, ldz, row_loss)
  DAE_DISPATCH_ACT(dec_act, ACT, {
    if (loss_func == DAE_LOSS_CE) DAE_LAUNCH_LOSS(ACT, DAE_LOSS_CE);
    else if (loss_func == DAE_LOSS_MSE) DAE_LAUNCH_LOSS(ACT, DAE_LOSS_MSE);
    else DAE_LAUNCH_LOSS(ACT, DAE_LOSS_COSINE);
  });
#undef DAE_LAUNCH_LOSS
  DAE_CHECK_LAUNCH("dae_decode_loss_bwd");
  return DAE_OK;
}

extern "C" int dae_colsum(const float* M, int32_t n_rows, int32_t n_cols, int64_t ld, float* out, void* stream) {
  using namespace dae;
  DAE_REQUIRE(M && out && n_rows >= 0 && n_cols > 0 && ld >= n_cols, "dae_colsum: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DAE_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * n_cols, st));
  if (n_rows == 0) return DAE_OK;
  const int rows_per_slab = 64;
  dim3 grid((n_cols + 127) / 128, (n_rows + rows_per_slab - 1) / rows_per_slab);
  colsum_kernel<<<grid, 128, 0, st>>>(M, n_rows, n_cols, ld, rows_per_slab, out);
  DAE_CHECK_LAUNCH("dae_colsum");
  return DAE_OK;
}

extern "C" int dae_step_finalize(const float* row_loss, const float* parts, int32_t n_parts, const float* weight, int32_t B,
                                 int32_t strategy, float alpha, double* stats, double* stats_log, const int64_t* ctl, void* stream) {
  using namespace dae;
  DAE_REQUIRE((row_loss || (parts && n_parts >= 1)) && stats && B >= 1, "dae_step_finalize: bad arguments");
  step_finalize_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(row_loss, parts, n_parts, weight, B, strategy, alpha, stats, stats_log, ctl);
  DAE_CHECK_LAUNCH("dae_step_finalize");
  return DAE_OK;
}

extern "C" int dae_reduce_parts(const float* parts, int32_t n_parts, int32_t n, float* out, void* stream) {
  using namespace dae;
  DAE_REQUIRE(parts && out && n_parts >= 1 && n >= 1, "dae_reduce_parts: bad arguments");
  reduce_parts_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(parts, n_parts, n, out);
  DAE_CHECK_LAUNCH("dae_reduce_parts");
  return DAE_OK;
}
