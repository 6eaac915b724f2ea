// Error plumbing + batch preparation kernel of libdae_sm100.so.
#include <cstdarg>
#include "common.cuh"

namespace dae {
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace dae

extern "C" int dae_version(void) { return 100; }

extern "C" int dae_last_error(char* buf, size_t len) {
  if (!buf || len == 0) return 0;
  strncpy(buf, dae::g_err, len - 1);
  buf[len - 1] = 0;
  return (int)strlen(buf);
}

namespace dae {

// One CTA. Orders the batch rows by label (bitonic sort in smem), derives class segments and the
// closed-form batch_all data weights:
//   w_i = 2(n-1)(B-n) + sum_{c != c_i} n_c(n_c-1),   N_valid = sum_c n_c(n_c-1)(B-n_c)
// which equal the three axis reductions of the B^3 mask in triplet_loss_utils.py:129 / :111.
constexpr int kMaxB = 4096;

__global__ void __launch_bounds__(1024) batch_prepare_kernel(
    const int32_t* __restrict__ perm, int64_t offset, const int64_t* __restrict__ ctl, int B, const float* __restrict__ labels_all,
    int strategy, int32_t* __restrict__ rows_out, float* __restrict__ labels_out, int32_t* seg_lo,
    int32_t* seg_hi, float* __restrict__ weight_out, double* __restrict__ stats, int64_t n_perm) {
  if (ctl) offset += ctl[0];  // device-resident batch cursor (CUDA-graph replay)
  if (n_perm > 0 && offset + B > n_perm) return;  // staging the batch AFTER the epoch's last one: nothing to prepare
  __shared__ float keys[kMaxB];
  __shared__ int vals[kMaxB];
  __shared__ double red[32];
  int32_t* lo = seg_lo;
  int32_t* hi = seg_hi;
  const int tid = threadIdx.x, nt = blockDim.x;
  int P = 1;
  while (P < B) P <<= 1;
  for (int i = tid; i < P; i += nt) {
    if (i < B) {
      const int r = perm ? perm[offset + i] : (int)(offset + i);
      vals[i] = r;
      keys[i] = (strategy != DAE_TRIPLET_NONE && labels_all) ? labels_all[r] : 0.0f;
    } else {
      vals[i] = 0x7fffffff;
      keys[i] = __int_as_float(0x7f800000);  // +inf sorts last
    }
  }
  __syncthreads();
  if (strategy != DAE_TRIPLET_NONE && P <= 1024) {
    // bitonic sort with one element per thread held in registers: partner exchange by warp shuffle for strides < 32
    // (40 of the 55 stages at P = 1024), through shared memory otherwise
    float key = keys[tid < P ? tid : 0];
    int val = vals[tid < P ? tid : 0];
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        float okey; int oval;
        if (j < 32) {
          okey = __shfl_xor_sync(0xffffffffu, key, j);
          oval = __shfl_xor_sync(0xffffffffu, val, j);
        } else {
          __syncthreads();
          if (tid < P) { keys[tid] = key; vals[tid] = val; }
          __syncthreads();
          okey = keys[(tid ^ j) & (P - 1)];
          oval = vals[(tid ^ j) & (P - 1)];
        }
        const bool up = ((tid & k) == 0), is_lower = ((tid & j) == 0);
        const bool other_less = (okey < key) || (okey == key && oval < val);
        const bool take_other = (up == is_lower) ? other_less : !other_less && !(okey == key && oval == val);
        if (take_other) { key = okey; val = oval; }
      }
    }
    __syncthreads();
    if (tid < P) { keys[tid] = key; vals[tid] = val; }
    __syncthreads();
  } else if (strategy != DAE_TRIPLET_NONE) {
    for (int k = 2; k <= P; k <<= 1) {
      for (int j = k >> 1; j > 0; j >>= 1) {
        for (int i = tid; i < P; i += nt) {
          const int ixj = i ^ j;
          if (ixj > i) {
            const float ka = keys[i], kb = keys[ixj];
            const int va = vals[i], vb = vals[ixj];
            const bool gt = (ka > kb) || (ka == kb && va > vb);
            const bool up = ((i & k) == 0);
            if (gt == up) {
              keys[i] = kb; keys[ixj] = ka;
              vals[i] = vb; vals[ixj] = va;
            }
          }
        }
        __syncthreads();
      }
    }
  }
  // class segment of every row: [first index with this label, one past the last) -- binary searches in the sorted keys
  for (int i = tid; i < B; i += nt) {
    const float k = keys[i];
    int a = 0, b = i;               // lower bound in [0, i]
    while (a < b) { const int m = (a + b) >> 1; if (keys[m] < k) a = m + 1; else b = m; }
    lo[i] = a;
    a = i + 1; b = B;               // upper bound in (i, B]
    while (a < b) { const int m = (a + b) >> 1; if (keys[m] <= k) a = m + 1; else b = m; }
    hi[i] = a;
  }
  __syncthreads();
  double t_part = 0.0, nv_part = 0.0;
  for (int i = tid; i < B; i += nt) {
    const double n = (double)(hi[i] - lo[i]);
    t_part += n - 1.0;
    nv_part += (n - 1.0) * ((double)B - n);
  }
  const double T = block_sum(t_part, red);
  const double NV = block_sum(nv_part, red);
  for (int i = tid; i < B; i += nt) {
    const double n = (double)(hi[i] - lo[i]);
    rows_out[i] = vals[i];
    if (labels_out) labels_out[i] = keys[i];
    if (weight_out) {
      float w = 1.0f;
      if (strategy == DAE_TRIPLET_BATCH_ALL) w = (float)(2.0 * (n - 1.0) * ((double)B - n) + T - n * (n - 1.0));
      if (strategy == DAE_TRIPLET_BATCH_HARD) w = 0.0f;  // filled by dae_triplet_batch_hard
      weight_out[i] = w;
    }
  }
  if (tid < DAE_STAT_SLOTS) {
    double v = 0.0;
    if (tid == DAE_STAT_SUM_W) v = (strategy == DAE_TRIPLET_BATCH_ALL) ? 3.0 * NV : (strategy == DAE_TRIPLET_NONE ? (double)B : 0.0);
    if (tid == DAE_STAT_N_VALID) v = (strategy == DAE_TRIPLET_BATCH_ALL) ? NV : 0.0;
    stats[tid] = v;
  }
}

// strategy none: keep the permutation order, w = 1, one segment; any B.
__global__ void batch_rows_kernel(const int32_t* __restrict__ perm, int64_t offset, const int64_t* __restrict__ ctl, int B,
                                  int32_t* __restrict__ rows_out, float* __restrict__ labels_out, int32_t* __restrict__ seg_lo,
                                  int32_t* __restrict__ seg_hi, float* __restrict__ weight_out, double* __restrict__ stats) {
  if (ctl) offset += ctl[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    rows_out[i] = perm ? perm[offset + i] : (int)(offset + i);
    if (labels_out) labels_out[i] = 0.0f;
    if (seg_lo) seg_lo[i] = 0;
    if (seg_hi) seg_hi[i] = B;
    if (weight_out) weight_out[i] = 1.0f;
  }
  if (i < DAE_STAT_SLOTS) stats[i] = (i == DAE_STAT_SUM_W) ? (double)B : 0.0;
}

// staged batch (prepared on a side branch during the previous step) -> the live per-batch buffers
__global__ void batch_commit_kernel(int B, const int32_t* __restrict__ rows_s, const float* __restrict__ labels_s,
                                    const int32_t* __restrict__ seg_lo_s, const int32_t* __restrict__ seg_hi_s,
                                    const float* __restrict__ weight_s, const double* __restrict__ stats_s, int32_t* __restrict__ rows,
                                    float* __restrict__ labels_b, int32_t* __restrict__ seg_lo, int32_t* __restrict__ seg_hi,
                                    float* __restrict__ weight, double* __restrict__ stats) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) { rows[i] = rows_s[i]; labels_b[i] = labels_s[i]; seg_lo[i] = seg_lo_s[i]; seg_hi[i] = seg_hi_s[i]; weight[i] = weight_s[i]; }
  if (i < DAE_STAT_SLOTS) stats[i] = stats_s[i];
}

// explicit (org, pos, neg) triplets: the three row blocks of the stacked [org; pos; neg] matrix for batch perm[offset : offset+B]
// (autoencoder/utils.py:73-91 gen_batches_triplet); each of the three reconstruction terms is a mean over B rows.
__global__ void batch_rows_explicit_kernel(const int32_t* __restrict__ perm, int64_t offset, const int64_t* __restrict__ ctl, int B,
                                           int64_t n_each, int32_t* __restrict__ rows_out, double* __restrict__ stats) {
  if (ctl) offset += ctl[0];
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < B) {
    const int64_t r = perm ? (int64_t)perm[offset + i] : offset + i;
    rows_out[i] = (int32_t)r;
    rows_out[B + i] = (int32_t)(r + n_each);
    rows_out[2 * B + i] = (int32_t)(r + 2 * n_each);
  }
  if (i < DAE_STAT_SLOTS) stats[i] = (i == DAE_STAT_SUM_W) ? (double)B : 0.0;
}

__global__ void step_advance_kernel(int64_t* ctl, int64_t row_stride) {
  ctl[0] += row_stride;  // batch cursor into the epoch permutation
  ctl[1] += 1;           // row of the per-epoch stats log
  ctl[2] += 1;           // optimizer step (Adam bias correction)
}

}  // namespace dae

extern "C" int dae_step_advance(int64_t* ctl, int64_t row_stride, void* stream) {
  DAE_REQUIRE(ctl, "dae_step_advance: null ctl");
  dae::step_advance_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(ctl, row_stride);
  DAE_CHECK_LAUNCH("dae_step_advance");
  return DAE_OK;
}

extern "C" int dae_batch_prepare(const int32_t* perm, int64_t offset, const int64_t* ctl, int32_t B, const float* labels_all,
                                 int32_t strategy, int32_t* rows_out, float* labels_out, int32_t* seg_lo,
                                 int32_t* seg_hi, float* weight_out, double* stats, void* stream) {
  DAE_REQUIRE(B >= 1 && rows_out && stats, "dae_batch_prepare: bad B or null output");
  if (strategy == DAE_TRIPLET_NONE) {
    dae::batch_rows_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(perm, offset, ctl, B, rows_out, labels_out, seg_lo,
                                                                            seg_hi, weight_out, stats);
    DAE_CHECK_LAUNCH("dae_batch_prepare(none)");
    return DAE_OK;
  }
  DAE_REQUIRE(B <= dae::kMaxB, "dae_batch_prepare: triplet strategies need B <= %d (got %d)", dae::kMaxB, B);
  DAE_REQUIRE(seg_lo && seg_hi, "dae_batch_prepare: null segment outputs");
  DAE_REQUIRE(strategy == DAE_TRIPLET_NONE || labels_all, "dae_batch_prepare: labels required for triplet strategies");
  dae::batch_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(perm, offset, ctl, B, labels_all, strategy, rows_out,
                                                                 labels_out, seg_lo, seg_hi, weight_out, stats, 0);
  DAE_CHECK_LAUNCH("dae_batch_prepare");
  return DAE_OK;
}

extern "C" int dae_batch_prepare_next(const int32_t* perm, int64_t n_perm, int64_t stride, const int64_t* ctl, int32_t B,
                                      const float* labels_all, int32_t strategy, int32_t* rows_s, float* labels_s, int32_t* seg_lo_s,
                                      int32_t* seg_hi_s, float* weight_s, double* stats_s, void* stream) {
  DAE_REQUIRE(ctl && n_perm > 0 && B >= 1 && B <= dae::kMaxB && rows_s && seg_lo_s && seg_hi_s && stats_s && labels_all,
              "dae_batch_prepare_next: bad arguments");
  DAE_REQUIRE(strategy == DAE_TRIPLET_BATCH_ALL || strategy == DAE_TRIPLET_BATCH_HARD, "dae_batch_prepare_next: triplet strategies only");
  dae::batch_prepare_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(perm, stride, ctl, B, labels_all, strategy, rows_s, labels_s, seg_lo_s,
                                                                 seg_hi_s, weight_s, stats_s, n_perm);
  DAE_CHECK_LAUNCH("dae_batch_prepare_next");
  return DAE_OK;
}

extern "C" int dae_batch_commit(int32_t B, const int32_t* rows_s, const float* labels_s, const int32_t* seg_lo_s, const int32_t* seg_hi_s,
                                const float* weight_s, const double* stats_s, int32_t* rows, float* labels_b, int32_t* seg_lo,
                                int32_t* seg_hi, float* weight, double* stats, void* stream) {
  DAE_REQUIRE(B >= 1 && rows_s && labels_s && seg_lo_s && seg_hi_s && weight_s && stats_s && rows && labels_b && seg_lo && seg_hi && weight && stats,
              "dae_batch_commit: null pointer");
  dae::batch_commit_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(B, rows_s, labels_s, seg_lo_s, seg_hi_s, weight_s, stats_s, rows,
                                                                           labels_b, seg_lo, seg_hi, weight, stats);
  DAE_CHECK_LAUNCH("dae_batch_commit");
  return DAE_OK;
}

extern "C" int dae_batch_prepare_explicit(const int32_t* perm, int64_t offset, const int64_t* ctl, int32_t B, int64_t n_each,
                                          int32_t* rows_out, double* stats, void* stream) {
  DAE_REQUIRE(B >= 1 && n_each >= 1 && rows_out && stats, "dae_batch_prepare_explicit: bad arguments");
  dae::batch_rows_explicit_kernel<<<(B + 255) / 256, 256, 0, (cudaStream_t)stream>>>(perm, offset, ctl, B, n_each, rows_out, stats);
  DAE_CHECK_LAUNCH("dae_batch_prepare_explicit");
  return DAE_OK;
}
