// "Next" row of the scope table: related-vs-unrelated AUROC of a pairwise similarity matrix.
//
// Reference ops replaced (helpers.visualize_pairwise_similarity, helpers.py:88-100):
//   not_nan / same-label masks -> np.tril(., -1) -> fancy-index gather of the two groups        -> pair_partition_kernel
//   sklearn roc_curve + auc over the concatenated groups ('Related' positive)                   -> auroc_count_kernel
// The area is evaluated as the Mann-Whitney statistic in exact integer arithmetic,
//   2 * AUROC * R * U = sum over related r of ( 2 * #{u < r} + #{u == r} ),
// which equals the trapezoid area of the ROC curve including its treatment of ties.  One of the groups is sorted in between
// (device radix sort driven by the host side); both kernels are HBM/L2-bandwidth work, no tensor cores involved.
#include "common.cuh"

namespace dae {

// One CTA per row i of the strict lower triangle.  Pass 1 counts the row's related / unrelated partners (labels only) and
// reserves two contiguous output ranges with ONE global atomic each; pass 2 streams the row of S and compacts it with warp
// ballots + shared-memory cursors.  Order inside a group is irrelevant (the consumer sorts or only counts).
__global__ void __launch_bounds__(256) pair_partition_kernel(const float* __restrict__ S, int64_t lds, int n, const int32_t* __restrict__ labels,
                                                             float* __restrict__ rel, float* __restrict__ unrel,
                                                             unsigned long long* __restrict__ cursors) {
  const int i = blockIdx.x + 1;
  if (i >= n) return;
  const int li = labels[i];
  if (li < 0) return;   // label -1 = missing: the whole row is dropped (helpers.py:91)
  __shared__ int s_cnt[2];
  __shared__ unsigned long long s_base[2];
  __shared__ int s_run[2];
  if (threadIdx.x < 2) { s_cnt[threadIdx.x] = 0; s_run[threadIdx.x] = 0; }
  __syncthreads();
  int cr = 0, cu = 0;
  for (int j = threadIdx.x; j < i; j += blockDim.x) {
    const int lj = labels[j];
    cr += (lj >= 0 && lj == li);
    cu += (lj >= 0 && lj != li);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    cr += __shfl_xor_sync(0xffffffffu, cr, o);
    cu += __shfl_xor_sync(0xffffffffu, cu, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (cr) atomicAdd(&s_cnt[0], cr);
    if (cu) atomicAdd(&s_cnt[1], cu);
  }
  __syncthreads();
  if (threadIdx.x < 2) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&cursors[threadIdx.x], (unsigned long long)s_cnt[threadIdx.x]) : 0ull;
  __syncthreads();
  const unsigned long long base_r = s_base[0], base_u = s_base[1];
  const float* row = S + (int64_t)i * lds;
  const int lane = threadIdx.x & 31;
  const unsigned lt = (1u << lane) - 1u;
  for (int j0 = (threadIdx.x >> 5) * 32; j0 < i; j0 += blockDim.x) {   // warp-uniform trip count
    const int j = j0 + lane;
    int lj = -1;
    float v = 0.0f;
    if (j < i) { lj = labels[j]; v = row[j]; }
    const bool is_r = (lj >= 0) && (lj == li);
    const bool is_u = (lj >= 0) && (lj != li);
    const unsigned mr = __ballot_sync(0xffffffffu, is_r), mu = __ballot_sync(0xffffffffu, is_u);
    int off_r = 0, off_u = 0;
    if (lane == 0) {
      if (mr) off_r = atomicAdd(&s_run[0], __popc(mr));
      if (mu) off_u = atomicAdd(&s_run[1], __popc(mu));
    }
    off_r = __shfl_sync(0xffffffffu, off_r, 0);
    off_u = __shfl_sync(0xffffffffu, off_u, 0);
    if (is_r) rel[base_r + off_r + __popc(mr & lt)] = v;
    if (is_u) unrel[base_u + off_u + __popc(mu & lt)] = v;
  }
}

// acc += sum over queries q of 2 * #{t "beaten" by the positive side} + #{t == q}, T sorted ascending.
//   query_is_positive = 1: Q = related scores,   T = unrelated (sorted): 2 * #{t < q}        + #{t == q}
//   query_is_positive = 0: Q = unrelated scores, T = related (sorted):   2 * #{t > q}        + #{t == q}
// Both bounds by branch-free binary search; consecutive threads take consecutive queries, so with a sorted (or merely
// clustered) Q neighbouring lanes walk the same cache lines of T.
__global__ void __launch_bounds__(256) auroc_count_kernel(const float* __restrict__ Q, int64_t nq, const float* __restrict__ T, int64_t nt,
                                                          int query_is_positive, unsigned long long* __restrict__ acc) {
  unsigned long long local = 0;
  for (int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; p < nq; p += (int64_t)gridDim.x * blockDim.x) {
    const float q = Q[p];
    int64_t lo = 0, len = nt;      // lower bound: first index with T[idx] >= q
    while (len > 0) {
      const int64_t half = len >> 1;
      const bool right = T[lo + half] < q;
      lo = right ? lo + half + 1 : lo;
      len = right ? len - half - 1 : half;
    }
    int64_t hi = lo;               // upper bound: first index with T[idx] > q (starts at the lower bound)
    len = nt - lo;
    while (len > 0) {
      const int64_t half = len >> 1;
      const bool right = !(q < T[hi + half]);
      hi = right ? hi + half + 1 : hi;
      len = right ? len - half - 1 : half;
    }
    const unsigned long long eq = (unsigned long long)(hi - lo);
    local += query_is_positive ? 2ull * (unsigned long long)lo + eq : 2ull * (unsigned long long)(nt - hi) + eq;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
  if ((threadIdx.x & 31) == 0 && local) atomicAdd(acc, local);
}

}  // namespace dae

extern "C" int dae_pair_partition(const float* S, int64_t lds, int32_t n, const int32_t* labels, float* related, float* unrelated,
                                  uint64_t* cursors, void* stream) {
  using namespace dae;
  DAE_REQUIRE(S && labels && related && unrelated && cursors && n > 0 && lds >= n, "dae_pair_partition: bad arguments");
  if (n < 2) return DAE_OK;
  pair_partition_kernel<<<n - 1, 256, 0, (cudaStream_t)stream>>>(S, lds, n, labels, related, unrelated, (unsigned long long*)cursors);
  DAE_CHECK_LAUNCH("dae_pair_partition");
  return DAE_OK;
}

extern "C" int dae_auroc_count(const float* queries, int64_t n_queries, const float* sorted_targets, int64_t n_targets,
                               int32_t query_is_positive, uint64_t* twice_u, void* stream) {
  using namespace dae;
  DAE_REQUIRE(twice_u && n_queries >= 0 && n_targets >= 0, "dae_auroc_count: bad arguments");
  if (n_queries == 0 || n_targets == 0) return DAE_OK;
  DAE_REQUIRE(queries && sorted_targets, "dae_auroc_count: null input");
  const int64_t blocks = (n_queries + 255) / 256;
  const int grid = (int)(blocks < 148 * 16 ? blocks : 148 * 16);   // 16 resident CTAs of 256 threads per SM, grid-stride beyond
  auroc_count_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(queries, n_queries, sorted_targets, n_targets, query_is_positive,
                                                            (unsigned long long*)twice_u);
  DAE_CHECK_LAUNCH("dae_auroc_count");
  return DAE_OK;
}
