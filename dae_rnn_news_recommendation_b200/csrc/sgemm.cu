// fp32 CUDA-core GEMM with generic strides.  Two jobs:
//  * the validation path for the large dense contractions (the production path is the tcgen05 kernel in gemm_tc.cu), so the
//    tensor-core kernels can be checked on the GPU at full size;
//  * the production path of the two SMALL contractions of the mining branch (S = E.E^T and dE2 = alpha (G + G^T) E, 0.64 GFLOP each):
//    a persistent tcgen05 CTA owns its SM's shared memory, so a small tensor-core GEMM cannot start while a large one runs; this
//    kernel needs 17 KB and the FFMA pipe, and runs NEXT TO the tensor-core decode chain.
#include "common.cuh"

namespace dae {

constexpr int BM = 128, BN = 128, BK = 16, TPB = 256;

// C[m,n] = alpha * sum_k A(m,k) B(n,k) + beta * C[m,n];   A(m,k) = A[m*sam + k*sak], B(n,k) = B[n*sbn + k*sbk]
__global__ void __launch_bounds__(TPB) sgemm_kernel(int M, int N, int K, float alpha, const float* __restrict__ A, int64_t sam,
                                                    int64_t sak, const float* __restrict__ B, int64_t sbn, int64_t sbk,
                                                    float beta, float* __restrict__ C, int64_t ldc, int kchunk) {
  __shared__ float As[BK][BM + 1];
  __shared__ float Bs[BK][BN + 1];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int kbeg = blockIdx.z * kchunk, kend = min(K, kbeg + kchunk);
  const bool a_kfast = (sak == 1), b_kfast = (sbk == 1);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  for (int k0 = kbeg; k0 < kend; k0 += BK) {
#pragma unroll
    for (int e = tid; e < BM * BK; e += TPB) {
      int m, k;
      if (a_kfast) { k = e % BK; m = e / BK; } else { m = e % BM; k = e / BM; }
      const int gm = m0 + m, gk = k0 + k;
      As[k][m] = (gm < M && gk < kend) ? __ldg(A + (int64_t)gm * sam + (int64_t)gk * sak) : 0.0f;
    }
#pragma unroll
    for (int e = tid; e < BN * BK; e += TPB) {
      int n, k;
      if (b_kfast) { k = e % BK; n = e / BK; } else { n = e % BN; k = e / BN; }
      const int gn = n0 + n, gk = k0 + k;
      Bs[k][n] = (gn < N && gk < kend) ? __ldg(B + (int64_t)gn * sbn + (int64_t)gk * sbk) : 0.0f;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float a[8], b[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) a[i] = As[k][ty + 16 * i];
#pragma unroll
      for (int j = 0; j < 8; ++j) b[j] = Bs[k][tx + 16 * j];
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int gm = m0 + ty + 16 * i;
    if (gm >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int gn = n0 + tx + 16 * j;
      if (gn >= N) continue;
      float* c = C + (int64_t)gm * ldc + gn;
      if (gridDim.z == 1) {
        *c = alpha * acc[i][j] + (beta != 0.0f ? beta * *c : 0.0f);
      } else {
        atomicAdd(c, alpha * acc[i][j]);  // split-K: C pre-scaled by beta in the launcher
      }
    }
  }
}

__global__ void scale_matrix_kernel(float* C, int M, int N, int64_t ldc, float beta) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n < N && m < M) C[(int64_t)m * ldc + n] = (beta == 0.0f) ? 0.0f : beta * C[(int64_t)m * ldc + n];
}

}  // namespace dae

extern "C" int dae_sgemm(int32_t M, int32_t N, int32_t K, float alpha, const float* A, int64_t sam, int64_t sak, const float* B,
                         int64_t sbn, int64_t sbk, float beta, float* C, int64_t ldc, void* stream) {
  using namespace dae;
  DAE_REQUIRE(A && B && C && M > 0 && N > 0 && K > 0 && ldc >= N, "dae_sgemm: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM, 1);
  // split-K when the tile grid cannot fill the 148 SMs: ~2 CTAs per SM, k chunks of at least 64
  int splits = 1;
  const int tiles = grid.x * grid.y;
  if (tiles < 148 && K >= 256) {
    splits = min((148 * 2 + tiles - 1) / tiles, K / 64);
    if (splits < 1) splits = 1;
  }
  int kchunk = (K + splits - 1) / splits;
  kchunk = (kchunk + BK - 1) / BK * BK;
  splits = (K + kchunk - 1) / kchunk;
  grid.z = splits;
  if (splits > 1) {
    dim3 g2((N + 255) / 256, M);
    scale_matrix_kernel<<<g2, 256, 0, st>>>(C, M, N, ldc, beta);
  }
  sgemm_kernel<<<grid, TPB, 0, st>>>(M, N, K, alpha, A, sam, sak, B, sbn, sbk, beta, C, ldc, kchunk);
  DAE_CHECK_LAUNCH("dae_sgemm");
  return DAE_OK;
}
