// tcgen05 / TMA / TMEM GEMM for the dense contractions of the DAE step, fp32-accurate through a 3-pass bf16 split.
//
// Reference ops replaced: tf.matmul(encode, tf.transpose(W)) (autoencoder/autoencoder.py:411) and its autodiff
// (dW = dZ^T.E, dE = dZ.W), and tf.matmul(encode, tf.transpose(encode)) (autoencoder/triplet_loss_utils.py:93,219).
//
// Numerics: every fp32 operand x is carried as two bf16 arrays, hi = bf16(x) and lo = bf16(x - hi) (|x - hi - lo| <= 2^-17 |x|).
//   D += A_hi.B_hi + A_lo.B_hi + A_hi.B_lo      (three kind::f16 MMAs per k-step, fp32 accumulation in TMEM)
// The dropped lo.lo term is <= 2^-18 relative, so the product is fp32-grade (1e-5 rel. worst case vs the 1e-4 budget).
//
// Structure (one CTA per SM, persistent over output tiles; 256 threads):
//   warp 0   : TMA producer   -- cp.async.bulk.tensor.2d, 128B swizzle, into a STAGES-deep smem ring (hi and lo tiles)
//   warp 1   : MMA issuer     -- one elected lane issues tcgen05.mma (UMMA 128 x BLOCK_N x 16), commits to mbarriers
//   warp 2   : TMEM allocator -- 512 columns = two BLOCK_N(<=256)-column fp32 accumulator stages
//   warps 4-7: epilogue       -- tcgen05.ld (lane = output row), fused epilogue, global stores; overlaps the next tile's MMAs
// Operands may be K-major (K contiguous) or MN-major (M/N contiguous) -- both straight from row-major arrays, so no
// transposed copies of dZ / E / W are ever made.
#include <cuda.h>
#include <cuda_bf16.h>
#include "common.cuh"

namespace dae {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;    // 64 bf16 = 128 bytes = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kAccStages = 2;

// ---------------------------------------------------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.u32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Spin, then back off with nanosleep; a wait that lasts longer than ~4 s (a lost TMA / commit: a programming error, not a
// slow peer) traps so the failure surfaces on the host at the next synchronisation instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t spins = 0;
  while (!mbar_try_wait(bar, parity)) {
    if (++spins > 4096u) {
      __nanosleep(64);
      if (spins > (1u << 26)) __trap();
    }
  }
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA pair (cta_group::2): two CTAs of a cluster, on the two SMs of one TPC, work on ONE 256-row UMMA tile.  Each CTA
// loads its own 128 rows of A and HALF of the B tile (the tensor cores read the other half from the peer's shared memory), so a
// k-block costs each SM 64 KB of L2 -> SM traffic instead of 96 KB.  Only the even ("leader") CTA issues MMAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;   // clears the CTA-rank bit of a shared::cluster address: "the same offset in the leader CTA"
__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load of a CTA pair: lands in THIS CTA's shared memory, completes its bytes on the LEADER's mbarrier
__device__ __forceinline__ void tma_load_2d_2sm(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void umma_bf16_2sm(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// commit that arrives on the mbarrier at this shared-memory offset in BOTH CTAs of the pair
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)0x3)
               : "memory");
}
// arrive on the mbarrier at this offset in CTA `rank` of the cluster
__device__ __forceinline__ void mbar_arrive_cluster(uint64_t* bar, uint32_t rank) {
  uint32_t raddr;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(raddr) : "r"(smem_u32(bar)), "r"(rank));
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(raddr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {   // acquire at cluster scope (remote arrivals)
  uint32_t spins = 0;
  for (;;) {
    uint32_t ok;
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n"
        "selp.u32 %0, 1, 0, p;\n"
        "}\n"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
    if (ok) return;
    if (++spins > 4096u) { __nanosleep(64); if (spins > (1u << 26)) __trap(); }
  }
}

__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
        "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor (sm_100 "version 1"), 128-byte swizzle.
//   K-major : rows of 128 B, 8-row groups 1024 B apart (SBO); LBO unused (1).
//   MN-major: 64-element (128 B) column slabs, 8 k-rows per 1024 B group (SBO), next slab LBO bytes further.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}

__host__ __device__ constexpr uint32_t make_idesc(int block_n, bool a_mn, bool b_mn, int umma_m = BLOCK_M) {
  return (1u << 4)                       // D format f32
         | (1u << 7) | (1u << 10)        // A, B = bf16
         | ((a_mn ? 1u : 0u) << 15)      // A major: 0 = K, 1 = MN
         | ((b_mn ? 1u : 0u) << 16)      // B major
         | ((uint32_t)(block_n >> 3) << 17) | ((uint32_t)(umma_m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------------------------------
// kernel parameters
// ---------------------------------------------------------------------------------------------------------------------
struct GemmParams {
  int M, N, K;           // logical sizes (TMA zero-fills out-of-range rows / columns)
  int k_splits;          // > 1: uniform split-K
  int stream_k;          // 1: stream-K -- the CTAs share the (tile, k-block) units evenly, partial tiles are added with atomics
  int atomic;            // accumulate into C with fp32 atomics (split-K / stream-K, or C += ...)
  int a_mn, b_mn;        // operand majorness
  int a_sym_kb;          // > 0: C = (A + A^T).B -- k-blocks [0, a_sym_kb) read the square A K-major, k-blocks [a_sym_kb, 2 a_sym_kb) read it
                         // MN-major (= A^T) through the second pair of tensor maps, and B's k index wraps (both halves multiply B)
  float alpha;
  float* C;              // EPI_STORE: fp32 output [M x ldc]
  int64_t ldc;
  int n_store;           // columns < n_store are stored
  int special_col;       // EPI_STORE: this column (the all-ones column of [E | 1]) goes to special_out[m] instead; -1 = none
  float* special_out;
  // EPI_DECODE (fused decode loss; M = batch rows, N = features)
  const int64_t* indptr; const int32_t* indices; const float* values; const int32_t* rows;
  const float* bv; const float* weight; const double* stats;
  __nv_bfloat16* dz_hi; __nv_bfloat16* dz_lo; int64_t ld_dz;
  float* row_loss_part;  // [M] row losses, accumulated with fp32 atomics (zeroed by the launcher)
  const int32_t* tile_ptr; // [M x (2 * n_tiles_n + 1)]: first CSR entry of every half tile, relative to the row start
};

enum { EPI_STORE = 0, EPI_DECODE = 1 };

// epilogue warps: EW / 4 warps per TMEM lane quarter, each takes 1/(EW/4) of the tile's columns
constexpr int kEwStore = 8, kEwDecode = 8;
constexpr int tc_threads(int ew) { return 128 + 32 * ew; }  // warps 0-3: TMA / MMA / TMEM-alloc / spare; then the epilogue warps

// Work distribution, walked identically by the TMA producer, the MMA issuer and the epilogue warps of a CTA.
//   classic : work item w = (tile, k split), items blockIdx.x, blockIdx.x + gridDim.x, ...
//   stream-K: the tiles x k-blocks units are cut into gridDim.x equal contiguous ranges; a CTA's range covers the tail of one
//             tile, whole tiles, and the head of another -- every segment is one accumulator pass + one (atomic) epilogue.
//             Balances shapes whose tile count does not fill the 148 SMs evenly (dW: 158 tiles) without shrinking the tiles.
struct Sched {
  int tiles_m, tiles, kb_total, kb_per_split, n_work, stream, w, n_cta;
  long long u, u_end;
  // pair != 0: the two CTAs of a cluster walk the SAME list; an m index then names a pair of 128-row tiles
  __device__ __forceinline__ void init(const GemmParams& p, int block_n, int pair = 0) {
    const int cta = pair ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
    n_cta = pair ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
    if (pair) tiles_m = (tiles_m + 1) / 2;
    tiles = tiles_m * ((p.N + block_n - 1) / block_n);
    kb_total = (p.K + BLOCK_K - 1) / BLOCK_K;
    stream = p.stream_k;
    kb_per_split = (kb_total + p.k_splits - 1) / p.k_splits;
    n_work = tiles * p.k_splits;
    w = cta;
    const long long U = (long long)tiles * kb_total;
    u = (long long)cta * U / n_cta;
    u_end = (long long)(cta + 1) * U / n_cta;
  }
  __device__ __forceinline__ bool next(int& mb, int& nb, int& kb0, int& kb1) {
    int tile;
    if (stream) {
      if (u >= u_end) return false;
      tile = (int)(u / kb_total);
      kb0 = (int)(u - (long long)tile * kb_total);
      const long long left = u_end - u;
      kb1 = (left < (long long)(kb_total - kb0)) ? kb0 + (int)left : kb_total;
      u += kb1 - kb0;
    } else {
      if (w >= n_work) return false;
      tile = w % tiles;
      kb0 = (w / tiles) * kb_per_split;
      kb1 = min(kb_total, kb0 + kb_per_split);
      w += n_cta;
    }
    mb = tile % tiles_m;
    nb = tile / tiles_m;
    return true;
  }
};

__device__ __forceinline__ uint32_t pack_bf16(float a, float b) {
  const __nv_bfloat162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<const uint32_t*>(&v);
}
__device__ __forceinline__ float f_rcp(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_lg2(float x) { float r; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
__device__ __forceinline__ float f_ex2(float x) { float r; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
constexpr float kLog2e = 1.4426950408889634f;

// decoder activation on the MUFU pipe (relative error ~1e-6, inside the 1e-4 parity budget)
template <int ACT>
__device__ __forceinline__ float act_fast(float z) {
  if (ACT == DAE_ACT_SIGMOID) return f_rcp(1.0f + f_ex2(-kLog2e * z));
  if (ACT == DAE_ACT_TANH) return 1.0f - 2.0f * f_rcp(1.0f + f_ex2(2.0f * kLog2e * z));
  return z;
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
        "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}

// r[j] for a runtime j in [0,16) without dynamic register indexing
__device__ __forceinline__ float select16(const uint32_t (&r)[16], int j) {
  uint32_t v = r[0];
#pragma unroll
  for (int k = 1; k < 16; ++k) v = (j == k) ? r[k] : v;
  return __uint_as_float(v);
}

// One 16-column chunk of the fused decode epilogue, evaluated as if every target x were 0 (99 % of a bag-of-words row is):
// z = acc + bv -> D = g(z) -> loss term -> dZ = sc * dloss/dz, packed as bf16 hi / lo pairs.  `lsum`: CE in log2 units, MSE plain.
template <int ACT, int LOSS>
__device__ __forceinline__ void decode_chunk_generic(const uint32_t (&r)[16], const float* __restrict__ bias, float sc, bool edge,
                                                     int n_lim, int nc, uint32_t (&hpk)[8], uint32_t (&lpk)[8], float& lsum) {
#pragma unroll
  for (int j2 = 0; j2 < 8; ++j2) {
    float dzp[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const int j = 2 * j2 + e;
      const float z = __uint_as_float(r[j]) + bias[j];
      const float d = act_fast<ACT>(z);
      float dz, lt;
      if (LOSS == DAE_LOSS_CE) {
        const float omd = 1.0f - d;
        const float b = omd + kEps;                 // 1. - decode + 1e-16, left to right (:269)
        lt = -f_lg2(b);
        // dl * g' = d * omd / b; omd / b == 1 exactly in fp32 unless omd == 0 (then the product is 0)
        dz = (ACT == DAE_ACT_SIGMOID) ? ((omd != 0.0f) ? sc * d : 0.0f) : sc * act_grad_from_y<ACT>(d) * f_rcp(b);
      } else {
        lt = d * d;
        dz = 2.0f * sc * d * act_grad_from_y<ACT>(d);
      }
      if (edge) { const bool in = (nc + j < n_lim); lt = in ? lt : 0.0f; dz = in ? dz : 0.0f; }  // uniform branch
      lsum += lt;
      dzp[e] = dz;
    }
    const uint32_t hp = pack_bf16(dzp[0], dzp[1]);
    hpk[j2] = hp;
    lpk[j2] = pack_bf16(dzp[0] - __uint_as_float(hp << 16), dzp[1] - __uint_as_float(hp & 0xffff0000u));
  }
}

// sigmoid + cross-entropy fast path (x = 0): with t = e^z,  1 - D = 1/(1+t),  -log(1 - D) = log(1+t),  dZ = sc * t/(1+t).
// Four columns share ONE reciprocal (Montgomery batch inversion of P = prod(1+t)) and ONE logarithm (log2 P): 1.5 MUFU per
// element instead of 3.  Valid while every z <= 10 (P <= e^10 bounds each factor, all factors being >= 1): there the fp32
// reference's own `1 - decode` rounding (2^-25 / (1 - D) relative) stays far below the parity budget; the caller re-evaluates the
// chunk with decode_chunk_generic when this returns false.
__device__ __forceinline__ bool decode_chunk_sigmoid_ce(const uint32_t (&r)[16], const float* __restrict__ biasc, float sc,
                                                        uint32_t (&hpk)[8], uint32_t (&lpk)[8], float& lsum) {
  bool ok = true;
#pragma unroll
  for (int g = 0; g < 4; ++g) {
    float t[4], a[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      t[j] = f_ex2(fmaf(__uint_as_float(r[4 * g + j]), kLog2e, biasc[4 * g + j]));   // biasc = bv * log2(e)
      a[j] = 1.0f + t[j];
    }
    const float p01 = a[0] * a[1], p23 = a[2] * a[3], P = p01 * p23;
    ok = ok && (P <= 22026.0f);                         // false for NaN / inf as well
    const float Rs = f_rcp(P) * sc;
    lsum += f_lg2(P);
    const float r23 = Rs * p23, r01 = Rs * p01;
    const float dz0 = t[0] * (r23 * a[1]), dz1 = t[1] * (r23 * a[0]), dz2 = t[2] * (r01 * a[3]), dz3 = t[3] * (r01 * a[2]);
    const uint32_t h0 = pack_bf16(dz0, dz1), h1 = pack_bf16(dz2, dz3);
    hpk[2 * g] = h0; hpk[2 * g + 1] = h1;
    lpk[2 * g] = pack_bf16(dz0 - __uint_as_float(h0 << 16), dz1 - __uint_as_float(h0 & 0xffff0000u));
    lpk[2 * g + 1] = pack_bf16(dz2 - __uint_as_float(h1 << 16), dz3 - __uint_as_float(h1 & 0xffff0000u));
  }
  return ok;
}

// PAIR = 1: cta_group::2 (see the PTX wrappers above); BLOCK_N is then the pair tile's N, of which each CTA stages BLOCK_N / 2.
template <int BLOCK_N, int STAGES, int EPI, int ACT, int LOSS, int PAIR>
__global__ void __launch_bounds__(tc_threads(EPI == EPI_DECODE ? kEwDecode : kEwStore), 1) gemm_bf16x3_kernel(const __grid_constant__ CUtensorMap tm_a_hi,
                                                                      const __grid_constant__ CUtensorMap tm_a_lo,
                                                                      const __grid_constant__ CUtensorMap tm_b_hi,
                                                                      const __grid_constant__ CUtensorMap tm_b_lo,
                                                                      const __grid_constant__ CUtensorMap tm_at_hi,   // A^T views (a_sym_kb > 0)
                                                                      const __grid_constant__ CUtensorMap tm_at_lo,
                                                                      const GemmParams p) {
  constexpr int A_TILE = BLOCK_M * BLOCK_K * 2;   // bytes of one bf16 A tile (16 KB)
  constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;   // n rows of the B tile staged by this CTA
  constexpr int B_TILE = B_ROWS * BLOCK_K * 2;
  constexpr int STAGE_BYTES = 2 * A_TILE + 2 * B_TILE;
  constexpr int kEpiWarps = (EPI == EPI_DECODE) ? kEwDecode : kEwStore;
  constexpr int kParts = kEpiWarps / 4;          // column parts per tile (one per epilogue warp of a lane quarter)
  constexpr int HALF_N = BLOCK_N / kParts;       // columns handled by one epilogue warp
  constexpr bool kFast = (EPI == EPI_DECODE) && (ACT == DAE_ACT_SIGMOID) && (LOSS == DAE_LOSS_CE);
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  __shared__ __align__(8) uint64_t full_bar[STAGES], empty_bar[STAGES], tmem_full_bar[kAccStages], tmem_empty_bar[kAccStages];
  __shared__ uint32_t tmem_base_smem;
  __shared__ float s_bias[EPI == EPI_DECODE ? kAccStages : 1][EPI == EPI_DECODE ? BLOCK_N : 1];
  __shared__ float s_biasc[kFast ? kAccStages : 1][kFast ? BLOCK_N : 1];   // bv * log2(e) for the sigmoid/CE fast path
  __shared__ __align__(16) float s_tr[EPI == EPI_STORE ? kEpiWarps : 1][32][20];  // per-warp transpose staging for coalesced stores
  __shared__ __align__(16) uint8_t s_stage[EPI == EPI_DECODE ? kEpiWarps : 1][2][32][48];  // bf16 hi / lo dZ blocks [32 rows x 16 cols], rows padded to 48 B

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  Sched sched;
  sched.init(p, BLOCK_N, PAIR);
  const uint32_t crank = PAIR ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs), 1 = peer

  if (warp == 0 && lane == 0) {
    prefetch_tmap(&tm_a_hi); prefetch_tmap(&tm_a_lo); prefetch_tmap(&tm_b_hi); prefetch_tmap(&tm_b_lo);
    if (p.a_sym_kb > 0) { prefetch_tmap(&tm_at_hi); prefetch_tmap(&tm_at_lo); }
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    // PAIR: the leader's accumulator-free barrier collects the epilogue warps of BOTH CTAs
    for (int s = 0; s < kAccStages; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], PAIR ? 2 * kEpiWarps : kEpiWarps); }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_smem)), "n"(512));
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // the peer's barriers must be initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = tmem_base_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0; uint32_t phase = 0;
      int mb, nb, kb0, kb1;
      while (sched.next(mb, nb, kb0, kb1)) {
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa_hi = smem + stage * STAGE_BYTES;
          uint8_t* sa_lo = sa_hi + A_TILE;
          uint8_t* sb_hi = sa_lo + A_TILE;
          uint8_t* sb_lo = sb_hi + B_TILE;
          const int mt = PAIR ? mb * 2 + (int)crank : mb;            // this CTA's 128-row m tile
          const int n_base = nb * BLOCK_N + (PAIR ? (int)crank * B_ROWS : 0);
          if (!PAIR) {
            mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
            if (p.a_sym_kb > 0 && kb >= p.a_sym_kb) {      // second half of (A + A^T).B: the same square array read M-contiguous
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) {
                tma_load_2d(&tm_at_hi, &full_bar[stage], sa_hi + j * 8192, mt * BLOCK_M + j * 64, (kb - p.a_sym_kb) * BLOCK_K);
                tma_load_2d(&tm_at_lo, &full_bar[stage], sa_lo + j * 8192, mt * BLOCK_M + j * 64, (kb - p.a_sym_kb) * BLOCK_K);
              }
            } else if (!p.a_mn) {
              tma_load_2d(&tm_a_hi, &full_bar[stage], sa_hi, kb * BLOCK_K, mt * BLOCK_M);
              tma_load_2d(&tm_a_lo, &full_bar[stage], sa_lo, kb * BLOCK_K, mt * BLOCK_M);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) {
                tma_load_2d(&tm_a_hi, &full_bar[stage], sa_hi + j * 8192, mt * BLOCK_M + j * 64, kb * BLOCK_K);
                tma_load_2d(&tm_a_lo, &full_bar[stage], sa_lo + j * 8192, mt * BLOCK_M + j * 64, kb * BLOCK_K);
              }
            }
            const int kbb = (p.a_sym_kb > 0 && kb >= p.a_sym_kb) ? kb - p.a_sym_kb : kb;   // B's k block
            if (!p.b_mn) {
              tma_load_2d(&tm_b_hi, &full_bar[stage], sb_hi, kbb * BLOCK_K, n_base);
              tma_load_2d(&tm_b_lo, &full_bar[stage], sb_lo, kbb * BLOCK_K, n_base);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j) {
                tma_load_2d(&tm_b_hi, &full_bar[stage], sb_hi + j * 8192, n_base + j * 64, kbb * BLOCK_K);
                tma_load_2d(&tm_b_lo, &full_bar[stage], sb_lo + j * 8192, n_base + j * 64, kbb * BLOCK_K);
              }
            }
          } else {
            // the LEADER's full barrier counts the bytes of both CTAs; the peer only issues its loads (which complete there)
            if (crank == 0) mbar_expect_tx(&full_bar[stage], 2 * STAGE_BYTES);
            if (!p.a_mn) {
              tma_load_2d_2sm(&tm_a_hi, &full_bar[stage], sa_hi, kb * BLOCK_K, mt * BLOCK_M);
              tma_load_2d_2sm(&tm_a_lo, &full_bar[stage], sa_lo, kb * BLOCK_K, mt * BLOCK_M);
            } else {
#pragma unroll
              for (int j = 0; j < BLOCK_M / 64; ++j) {
                tma_load_2d_2sm(&tm_a_hi, &full_bar[stage], sa_hi + j * 8192, mt * BLOCK_M + j * 64, kb * BLOCK_K);
                tma_load_2d_2sm(&tm_a_lo, &full_bar[stage], sa_lo + j * 8192, mt * BLOCK_M + j * 64, kb * BLOCK_K);
              }
            }
            if (!p.b_mn) {
              tma_load_2d_2sm(&tm_b_hi, &full_bar[stage], sb_hi, kb * BLOCK_K, n_base);
              tma_load_2d_2sm(&tm_b_lo, &full_bar[stage], sb_lo, kb * BLOCK_K, n_base);
            } else {
#pragma unroll
              for (int j = 0; j < B_ROWS / 64; ++j) {
                tma_load_2d_2sm(&tm_b_hi, &full_bar[stage], sb_hi + j * 8192, n_base + j * 64, kb * BLOCK_K);
                tma_load_2d_2sm(&tm_b_lo, &full_bar[stage], sb_lo + j * 8192, n_base + j * 64, kb * BLOCK_K);
              }
            }
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0 && crank == 0) {
      const uint32_t idesc_n = make_idesc(BLOCK_N, p.a_mn != 0, p.b_mn != 0, PAIR ? 2 * BLOCK_M : BLOCK_M);
      const uint32_t idesc_t = make_idesc(BLOCK_N, true, p.b_mn != 0, PAIR ? 2 * BLOCK_M : BLOCK_M);   // A read MN-major (A^T half)
      int stage = 0; uint32_t phase = 0;
      int acc = 0; uint32_t acc_phase = 0;
      int mb, nb, kb0, kb1;
      while (sched.next(mb, nb, kb0, kb1)) {
        if (PAIR) mbar_wait_cluster(&tmem_empty_bar[acc], acc_phase ^ 1);   // the epilogues of both CTAs have drained this stage
        else mbar_wait(&tmem_empty_bar[acc], acc_phase ^ 1);                // epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t tmem_d = tmem_base + acc * BLOCK_N;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t sa_hi = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t sa_lo = sa_hi + A_TILE, sb_hi = sa_lo + A_TILE, sb_lo = sb_hi + B_TILE;
          const bool a_t = (p.a_sym_kb > 0 && kb >= p.a_sym_kb) || p.a_mn;                 // this k-block's A tile is MN-major
          const uint32_t idesc = (p.a_sym_kb > 0 && kb >= p.a_sym_kb) ? idesc_t : idesc_n;
          const uint32_t a_lbo = a_t ? 8192u : 16u, b_lbo = p.b_mn ? 8192u : 16u;
          const uint32_t a_step = a_t ? 2048u : 32u, b_step = p.b_mn ? 2048u : 32u;  // bytes per UMMA_K
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            const uint64_t da_hi = make_desc(sa_hi + k * a_step, a_lbo, 1024);
            const uint64_t da_lo = make_desc(sa_lo + k * a_step, a_lbo, 1024);
            const uint64_t db_hi = make_desc(sb_hi + k * b_step, b_lbo, 1024);
            const uint64_t db_lo = make_desc(sb_lo + k * b_step, b_lbo, 1024);
            const uint32_t first = (kb == kb0 && k == 0) ? 0u : 1u;
            if (PAIR) {
              umma_bf16_2sm(tmem_d, da_lo, db_hi, idesc, first);
              umma_bf16_2sm(tmem_d, da_hi, db_lo, idesc, 1u);
              umma_bf16_2sm(tmem_d, da_hi, db_hi, idesc, 1u);
            } else {
              umma_bf16(tmem_d, da_lo, db_hi, idesc, first);   // small terms first
              umma_bf16(tmem_d, da_hi, db_lo, idesc, 1u);
              umma_bf16(tmem_d, da_hi, db_hi, idesc, 1u);
            }
          }
          if (PAIR) {                            // both CTAs' producers / epilogues are released
            umma_commit_2sm(&empty_bar[stage]);
            if (kb == kb1 - 1) umma_commit_2sm(&tmem_full_bar[acc]);
          } else {
            umma_commit(&empty_bar[stage]);      // frees this smem stage when the MMAs retire
            if (kb == kb1 - 1) umma_commit(&tmem_full_bar[acc]);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
      }
    }
  } else if (warp >= 4) {
    // ===================== epilogue: TMEM -> registers -> global =====================
    const int ew = warp - 4;
    const int quarter = ew & 3;              // TMEM lane quarter this warp may access (warp id % 4)
    const int half = ew >> 2;                // which column part of the tile
    const int row_in_tile = quarter * 32 + lane;
    const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
    int acc = 0; uint32_t acc_phase = 0;
    int mb, nb, kb0, kb1;
    while (sched.next(mb, nb, kb0, kb1)) {
      if (PAIR) mb = mb * 2 + (int)crank;     // this CTA's 128-row m tile of the pair
      const int m = mb * BLOCK_M + row_in_tile;
      const int n0 = nb * BLOCK_N + half * HALF_N;
      if (EPI == EPI_DECODE) {  // stage this tile's visible-bias slice (named barrier 1: the epilogue threads)
        for (int j = threadIdx.x - 128; j < BLOCK_N; j += 32 * kEpiWarps) {
          const float b = (nb * BLOCK_N + j < p.N) ? p.bv[nb * BLOCK_N + j] : 0.0f;
          s_bias[acc][j] = b;
          if (kFast) s_biasc[acc][j] = b * kLog2e;
        }
        asm volatile("bar.sync 1, %0;" ::"n"(32 * kEpiWarps) : "memory");
      }
      mbar_wait(&tmem_full_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N + half * HALF_N;

      if (EPI == EPI_STORE) {
        float (*tr)[20] = s_tr[ew];
        const int m_base = mb * BLOCK_M + quarter * 32;
        const bool vec_ok = ((p.ldc & 3) == 0) && ((reinterpret_cast<uintptr_t>(p.C) & 15) == 0);
#pragma unroll 1
        for (int c = 0; c < HALF_N / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(taddr + c * 16, r);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<float4*>(&tr[lane][4 * q]) =
                make_float4(p.alpha * __uint_as_float(r[4 * q]), p.alpha * __uint_as_float(r[4 * q + 1]),
                            p.alpha * __uint_as_float(r[4 * q + 2]), p.alpha * __uint_as_float(r[4 * q + 3]));
          __syncwarp();
          const int nc = n0 + c * 16;   // first column of this 16-wide chunk
          const bool interior = vec_ok && (m_base + 32 <= p.M) && (nc + 16 <= p.n_store) && (p.special_col < nc || p.special_col >= nc + 16);
          if (interior) {               // fast path: 8 rows x 64 B per store instruction, 128-bit accesses
            const int rsub = lane >> 2, c4 = (lane & 3) * 4;
            float* dst = p.C + (int64_t)(m_base + rsub) * p.ldc + nc + c4;
            const int64_t step = 8 * p.ldc;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float4 v = *reinterpret_cast<const float4*>(&tr[rsub + 8 * i][c4]);
              if (p.atomic) atomicAdd(reinterpret_cast<float4*>(dst), v); else *reinterpret_cast<float4*>(dst) = v;
              dst += step;
            }
          } else {                      // edge tiles / the [dW | dbv] column / unaligned C
            const int rsub = lane >> 4, csub = lane & 15;
            const int n = nc + csub;
            for (int i = 0; i < 16; ++i) {
              const int rr = 2 * i + rsub;
              const int mm = m_base + rr;
              const float v = tr[rr][csub];
              if (mm < p.M) {
                if (n == p.special_col) {
                  if (p.atomic) atomicAdd(p.special_out + mm, v); else p.special_out[mm] = v;
                } else if (n < p.n_store) {
                  float* dst = p.C + (int64_t)mm * p.ldc + n;
                  if (p.atomic) atomicAdd(dst, v); else *dst = v;
                }
              }
            }
          }
          __syncwarp();
        }
      } else {
        // ---- fused decode epilogue: D = g(Z + bv); row loss; dZ -> bf16 hi/lo (autoencoder.py:411, triplet_loss_utils.py:269-275)
        // 99 % of a bag-of-words target row is zero, so each 16-column chunk is first evaluated branch-free as if x == 0,
        // stored, and then the row's few stored entries inside the chunk are re-evaluated exactly and patched in place.  The clean
        // CSR row is walked with a cursor (columns are sorted) that keeps the next THREE entries (c0,v0),(c1,v1),(c2,v2) in
        // registers, loaded well before they are needed, so the L2 latency of the CSR stream never sits on the column loop.
        int64_t pc = 0, pe = 0;
        int c0 = 0x7fffffff, c1 = 0x7fffffff, c2 = 0x7fffffff;
        float v0 = 0.0f, v1 = 0.0f, v2 = 0.0f;
        float sc = 0.0f;
        if (m < p.M) {
          const int64_t row = p.rows ? (int64_t)p.rows[m] : (int64_t)m;
          const int64_t rbeg = p.indptr[row];
          const int32_t* tp = p.tile_ptr + (int64_t)m * (kParts * tiles_n + 1) + nb * kParts + half;
          pc = rbeg + tp[0]; pe = rbeg + tp[1];     // entries of this row that fall into this half tile
          if (pc < pe) { c0 = p.indices[pc]; v0 = p.values[pc]; }
          if (pc + 1 < pe) { c1 = p.indices[pc + 1]; v1 = p.values[pc + 1]; }
          if (pc + 2 < pe) { c2 = p.indices[pc + 2]; v2 = p.values[pc + 2]; }
          sc = (p.weight ? p.weight[m] : 1.0f) / ((float)p.stats[DAE_STAT_SUM_W] + kEps);
        }
        const float inv_sc = (sc != 0.0f) ? 1.0f / sc : 0.0f;
        const bool edge = (n0 + HALF_N > p.N);   // only the last column tile has out-of-range columns
        float lsum = 0.0f;   // CE: accumulated in log2 units, scaled by ln2 at the end
        // dZ leaves through a per-warp shared-memory transpose: each thread (= batch row) drops the bf16 hi and lo parts of its
        // 16 values into its two staging rows (patching the row's stored entries there), then the warp writes both 32 x 16
        // blocks with full 32-byte sectors (16 rows per store instruction).
        uint8_t* stg_hi = reinterpret_cast<uint8_t*>(&s_stage[ew][0][0][0]);
        uint8_t* stg_lo = reinterpret_cast<uint8_t*>(&s_stage[ew][1][0][0]);
        const int rsub = lane >> 1, csub = lane & 1;          // write-out mapping: 16 rows x 2 x 16 B per instruction
        const int m_base = mb * BLOCK_M + quarter * 32;
        const int n_lim = edge ? p.N : 0x7fffffff;
#pragma unroll 1
        for (int c = 0; c < HALF_N / 16; ++c) {
          uint32_t r[16];
          tmem_ld16(taddr + c * 16, r);
          tmem_ld_wait();
          const int nc = n0 + c * 16;
          if (m < p.M) {
            const float* bias = &s_bias[acc][half * HALF_N + c * 16];
            uint32_t hpk[8], lpk[8];
            bool fast_ok = false;   // this chunk went through the sigmoid/CE fast path: staged dZ = sc * D exactly
            if (kFast && !edge) {
              const float l0 = lsum;
              fast_ok = decode_chunk_sigmoid_ce(r, &s_biasc[acc][half * HALF_N + c * 16], sc, hpk, lpk, lsum);
              if (!fast_ok) { lsum = l0; decode_chunk_generic<ACT, LOSS>(r, bias, sc, edge, n_lim, nc, hpk, lpk, lsum); }
            } else {
              decode_chunk_generic<ACT, LOSS>(r, bias, sc, edge, n_lim, nc, hpk, lpk, lsum);
            }
            uint4* sh = reinterpret_cast<uint4*>(stg_hi + lane * 48);
            uint4* sl = reinterpret_cast<uint4*>(stg_lo + lane * 48);
            sh[0] = make_uint4(hpk[0], hpk[1], hpk[2], hpk[3]); sh[1] = make_uint4(hpk[4], hpk[5], hpk[6], hpk[7]);
            sl[0] = make_uint4(lpk[0], lpk[1], lpk[2], lpk[3]); sl[1] = make_uint4(lpk[4], lpk[5], lpk[6], lpk[7]);
            // exact re-evaluation of the stored entries of this row inside the group (densified target, :264)
            while (c0 < nc + 16) {
              const float x = v0;
              const int j = c0 - nc;
              __nv_bfloat16* ph = reinterpret_cast<__nv_bfloat16*>(stg_hi + lane * 48) + j;
              __nv_bfloat16* pl = reinterpret_cast<__nv_bfloat16*>(stg_lo + lane * 48) + j;
              float d;
              if (kFast && fast_ok && sc != 0.0f) d = (__bfloat162float(*ph) + __bfloat162float(*pl)) * inv_sc;   // D back from the staged sc * D
              else d = act_fast<ACT>(select16(r, j) + bias[j]);
              const float gp = act_grad_from_y<ACT>(d);
              float dz;
              if (LOSS == DAE_LOSS_CE) {
                const float a = d + kEps, b = (1.0f - d) + kEps;
                const float la = f_lg2(a), lb = f_lg2(b);
                lsum += lb - (x * la + (1.0f - x) * lb);      // replace the x == 0 term by the exact one
                dz = sc * gp * ((1.0f - x) * f_rcp(b) - x * f_rcp(a));
              } else {
                const float e2 = x - d;
                lsum += e2 * e2 - d * d;
                dz = -2.0f * sc * e2 * gp;
              }
              const __nv_bfloat16 hb = __float2bfloat16_rn(dz);
              *ph = hb;
              *pl = __float2bfloat16_rn(dz - __bfloat162float(hb));
              ++pc;
              c0 = c1; v0 = v1; c1 = c2; v1 = v2;
              c2 = 0x7fffffff;
              if (pc + 2 < pe) { c2 = p.indices[pc + 2]; v2 = p.values[pc + 2]; }
            }
          }
          __syncwarp();
          if (nc < p.ld_dz) {  // ld_dz is a multiple of 32 (launcher), so whole 16-column groups are in range
#pragma unroll
            for (int i = 0; i < 2; ++i) {
              const int rr = rsub + 16 * i;
              if (m_base + rr < p.M) {
                const int64_t off = (int64_t)(m_base + rr) * p.ld_dz + nc + csub * 8;
                *reinterpret_cast<uint4*>(p.dz_hi + off) = *reinterpret_cast<const uint4*>(stg_hi + rr * 48 + csub * 16);
                *reinterpret_cast<uint4*>(p.dz_lo + off) = *reinterpret_cast<const uint4*>(stg_lo + rr * 48 + csub * 16);
              }
            }
          }
          __syncwarp();
        }
        if (m < p.M) atomicAdd(p.row_loss_part + m, (LOSS == DAE_LOSS_CE) ? lsum * 0.6931471805599453f : lsum);  // 2 * tiles_n partials per row
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) { if (PAIR) mbar_arrive_cluster(&tmem_empty_bar[acc], 0); else mbar_arrive(&tmem_empty_bar[acc]); }
      if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();   // nobody leaves while the peer may still read this CTA's tiles / signal its barriers
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(512));
  }
}

// tile_ptr[m][t] = number of stored entries of batch row m with column < t * half_n (t = 0 .. n_half_tiles): where each
// half tile of the fused decode epilogue starts in the (sorted) clean CSR row.  One thread per (row, t).
__global__ void decode_tile_ptr_kernel(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                                       const int32_t* __restrict__ rows, int M, int n_half_tiles, int half_n,
                                       int32_t* __restrict__ tile_ptr) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (t > n_half_tiles || m >= M) return;
  const int64_t row = rows ? (int64_t)rows[m] : (int64_t)m;
  const int64_t b = indptr[row], e = indptr[row + 1];
  const int target = t * half_n;
  int64_t lo = b, hi = e;
  while (lo < hi) { const int64_t mid = (lo + hi) >> 1; if (indices[mid] < target) lo = mid + 1; else hi = mid; }
  tile_ptr[(int64_t)m * (n_half_tiles + 1) + t] = (int32_t)(lo - b);
}

// fp32 -> (bf16 hi, bf16 lo) split, row by row, zero padding up to ld_dst columns; optional 1.0 in column `ones_col`.
__global__ void split_bf16_kernel(const float* __restrict__ src, int rows, int cols, int64_t ld_src, __nv_bfloat16* __restrict__ hi,
                                  __nv_bfloat16* __restrict__ lo, int64_t ld_dst, int ones_col, float scale) {
  const int64_t total = (int64_t)rows * ld_dst;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / ld_dst;
    const int c = (int)(i - r * ld_dst);
    float v = (c < cols) ? src[r * ld_src + c] * scale : 0.0f;
    if (c == ones_col) v = 1.0f;
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    hi[i] = h;
    lo[i] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
}

// GG = alpha * (G + G^T) split to bf16 hi/lo (B x ld), for dE_tri = GG . E
__global__ void sym_split_kernel(const float* __restrict__ G, int B, int64_t ldg, float alpha, __nv_bfloat16* __restrict__ hi,
                                 __nv_bfloat16* __restrict__ lo, int64_t ld) {
  __shared__ float t[32][33];
  const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
  const int tx = threadIdx.x, ty = threadIdx.y;  // 32 x 8
  for (int j = ty; j < 32; j += 8) {
    const int r = bx + j, c = by + tx;            // read G^T tile: G[bx + j][by + tx]
    t[j][tx] = (r < B && c < B) ? G[(int64_t)r * ldg + c] : 0.0f;
  }
  __syncthreads();
  for (int j = ty; j < 32; j += 8) {
    const int r = by + j, c = bx + tx;            // output element (r, c) = G[r][c] + G[c][r]
    if (r < B && c < ld) {
      float v = 0.0f;
      if (c < B) v = alpha * (G[(int64_t)r * ldg + c] + t[tx][j]);
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      hi[(int64_t)r * ld + c] = h;
      lo[(int64_t)r * ld + c] = __float2bfloat16_rn(v - __bfloat162float(h));
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------------
// host side: tensor maps + launch
// ---------------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                    const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                    CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_encodeTiled>(sym);
  }
  return fn;
}

// 2-D bf16 tensor [outer x inner] (inner contiguous), row stride ld elements; box = {64, box_outer}, 128B swizzle.
static int make_map(CUtensorMap* m, const void* base, uint64_t inner, uint64_t outer, uint64_t ld, uint32_t box_outer) {
  PFN_encodeTiled enc = get_encode();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not found"); return DAE_ERR_CUDA; }
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { set_error("cuTensorMapEncodeTiled failed (%d): inner=%llu outer=%llu ld=%llu", (int)r,
                                     (unsigned long long)inner, (unsigned long long)outer, (unsigned long long)ld); return DAE_ERR_CUDA; }
  return DAE_OK;
}

struct Operand { const void* hi; const void* lo; int64_t ld; int mn_major; };

// cudaFuncSetAttribute is per device: remember which devices have been configured for this instantiation
template <typename Kern>
static int ensure_smem_attr(Kern kern, int smem, bool (&done)[64]) {
  int dev = 0;
  DAE_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || !done[dev]) {
    DAE_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    if (dev >= 0 && dev < 64) done[dev] = true;
  }
  return DAE_OK;
}

static int sm_count() {
  static int n[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (!n[dev]) {
    int v = 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) v = 148;
    n[dev] = v;
  }
  return n[dev];
}

static int g_pair_mode = -1;   // -1: by size (dae_gemm_config); 0: never; 1: whenever the shape allows
static int g_lean = 0;         // 1: leave ~70 KB of each SM's shared memory to concurrently running kernels (2-stage pipelines)

template <int BLOCK_N, int STAGES, int EPI, int ACT, int LOSS, int PAIR>
static int launch_gemm(const Operand& A, const Operand& B, GemmParams p, cudaStream_t st) {
  CUtensorMap ta_hi, ta_lo, tb_hi, tb_lo;
  int rc;
  constexpr int B_ROWS = PAIR ? BLOCK_N / 2 : BLOCK_N;
  // K-major: tensor [rows=MN x cols=K], box {64 k, tile rows};  MN-major: tensor [rows=K x cols=MN], box {64 mn, 64 k}
  if (!A.mn_major) {
    const uint64_t a_cols = p.a_sym_kb > 0 ? (uint64_t)p.M : (uint64_t)p.K;   // (A + A^T).B: A is M x M, K = 2 x (padded M)
    if ((rc = make_map(&ta_hi, A.hi, a_cols, p.M, A.ld, BLOCK_M))) return rc;
    if ((rc = make_map(&ta_lo, A.lo, a_cols, p.M, A.ld, BLOCK_M))) return rc;
  } else {
    if ((rc = make_map(&ta_hi, A.hi, p.M, p.K, A.ld, 64))) return rc;
    if ((rc = make_map(&ta_lo, A.lo, p.M, p.K, A.ld, 64))) return rc;
  }
  if (!B.mn_major) {
    if ((rc = make_map(&tb_hi, B.hi, p.K, p.N, B.ld, B_ROWS))) return rc;
    if ((rc = make_map(&tb_lo, B.lo, p.K, p.N, B.ld, B_ROWS))) return rc;
  } else {
    const uint64_t b_rows = p.a_sym_kb > 0 ? (uint64_t)p.M : (uint64_t)p.K;
    if ((rc = make_map(&tb_hi, B.hi, p.N, b_rows, B.ld, 64))) return rc;
    if ((rc = make_map(&tb_lo, B.lo, p.N, b_rows, B.ld, 64))) return rc;
  }
  p.a_mn = A.mn_major; p.b_mn = B.mn_major;
  CUtensorMap tat_hi = ta_hi, tat_lo = ta_lo;
  if (p.a_sym_kb > 0) {   // A is a square [M x M] array read K-major above; these are its M-contiguous (transposed) views
    if (PAIR || A.mn_major) { set_error("dae_gemm_sym_bf16x3: unsupported configuration"); return DAE_ERR_UNSUPPORTED; }
    if ((rc = make_map(&tat_hi, A.hi, p.M, p.M, A.ld, 64))) return rc;
    if ((rc = make_map(&tat_lo, A.lo, p.M, p.M, A.ld, 64))) return rc;
  }
  constexpr int smem = STAGES * (2 * BLOCK_M * BLOCK_K * 2 + 2 * B_ROWS * BLOCK_K * 2) + 1024;
  constexpr int threads = tc_threads(EPI == EPI_DECODE ? kEwDecode : kEwStore);
  int tiles_m = (p.M + BLOCK_M - 1) / BLOCK_M;
  if (PAIR) tiles_m = (tiles_m + 1) / 2;
  const int tiles_n = (p.N + BLOCK_N - 1) / BLOCK_N;
  const int kblocks = (p.K + BLOCK_K - 1) / BLOCK_K;
  auto kern = gemm_bf16x3_kernel<BLOCK_N, STAGES, EPI, ACT, LOSS, PAIR>;
  static bool attr_done[64] = {false};
  if ((rc = ensure_smem_attr(kern, smem, attr_done))) return rc;
  const int slots = PAIR ? sm_count() / 2 : sm_count();   // CTAs, or CTA pairs, resident at once
  int n;
  if (p.stream_k) {   // segments of at least ~6 k-blocks: a shorter main loop does not amortise its (atomic) epilogue
    const long long units = (long long)tiles_m * tiles_n * kblocks;
    long long g = units / 6;
    if (g < 1) g = 1;
    n = (int)(g < slots ? g : slots);
  } else {
    const int items = tiles_m * tiles_n * p.k_splits;
    n = items < slots ? items : slots;
  }
  if (!PAIR) {
    kern<<<n, threads, smem, st>>>(ta_hi, ta_lo, tb_hi, tb_lo, tat_hi, tat_lo, p);
  } else {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * n); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    DAE_CUDA(cudaLaunchKernelEx(&cfg, kern, ta_hi, ta_lo, tb_hi, tb_lo, tat_hi, tat_lo, p));
  }
  return DAE_OK;
}

// column-tile width of the fused decode kernel under the current configuration (dae_decode_prepare lays tile_ptr out for it)
static int decode_tile_width() { return (g_pair_mode == 1 || !g_lean) ? 256 : 128; }

// CTA pairs pay off once the GEMM is large enough to be bound by L2 -> SM operand traffic (>= ~2 GFLOP here: decode, dW, dE)
static bool use_pair(int M, int N, int K) {
  if (g_pair_mode == 0) return false;
  if (g_pair_mode == 1) return true;
  return (double)M * (double)N * (double)K >= 1.0e9 && M > 128;
}

}  // namespace dae

using namespace dae;

extern "C" int dae_split_bf16(const float* src, int32_t rows, int32_t cols, int64_t ld_src, void* hi, void* lo, int64_t ld_dst,
                              int32_t ones_col, float scale, void* stream) {
  DAE_REQUIRE(src && hi && lo && rows > 0 && cols > 0 && ld_src >= cols && ld_dst >= cols, "dae_split_bf16: bad arguments");
  const int64_t total = (int64_t)rows * ld_dst;
  const int blocks = (int)((total + 255) / 256 < 148 * 16 ? (total + 255) / 256 : 148 * 16);
  split_bf16_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(src, rows, cols, ld_src, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld_dst,
                                                              ones_col, scale);
  DAE_CHECK_LAUNCH("dae_split_bf16");
  return DAE_OK;
}

extern "C" int dae_sym_split_bf16(const float* G, int32_t B, int64_t ldg, float alpha, void* hi, void* lo, int64_t ld, void* stream) {
  DAE_REQUIRE(G && hi && lo && B > 0 && ldg >= B && ld >= B, "dae_sym_split_bf16: bad arguments");
  dim3 grid((unsigned)((ld + 31) / 32), (B + 31) / 32), block(32, 8);
  sym_split_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(G, B, ldg, alpha, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, ld);
  DAE_CHECK_LAUNCH("dae_sym_split_bf16");
  return DAE_OK;
}

// test hook: -1 = CTA pairs (cta_group::2) for the large GEMMs only (default), 0 = never, 1 = always
extern "C" int dae_gemm_config(int32_t pair_mode, int32_t lean) {
  dae::g_pair_mode = pair_mode < 0 ? -1 : (pair_mode ? 1 : 0);
  dae::g_lean = lean ? 1 : 0;
  return DAE_OK;
}

extern "C" int dae_gemm_bf16x3(int32_t M, int32_t N, int32_t K, float alpha, const void* a_hi, const void* a_lo, int64_t lda,
                               int32_t a_mn_major, const void* b_hi, const void* b_lo, int64_t ldb, int32_t b_mn_major, float* C,
                               int64_t ldc, int32_t n_store, int32_t special_col, float* special_out, int32_t k_splits,
                               int32_t accumulate, void* stream) {
  DAE_REQUIRE(a_hi && a_lo && b_hi && b_lo && C && M > 0 && N > 0 && K > 0, "dae_gemm_bf16x3: bad arguments");
  DAE_REQUIRE(lda % 8 == 0 && ldb % 8 == 0, "dae_gemm_bf16x3: operand leading dimensions must be multiples of 8 (TMA 16-byte strides)");
  DAE_REQUIRE(((uintptr_t)a_hi | (uintptr_t)a_lo | (uintptr_t)b_hi | (uintptr_t)b_lo) % 16 == 0, "dae_gemm_bf16x3: operands must be 16-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  if (n_store <= 0 || n_store > N) n_store = N;
  const int kblocks = (K + BLOCK_K - 1) / BLOCK_K;
  const int tm = (M + 127) / 128, tn256 = (N + 255) / 256, tn128 = (N + 127) / 128;
  const int sms = sm_count();
  int stream_k = 0;
  if (k_splits < 0) {   // auto: stream-K unless the 128 x 256 tiling already fills the SMs in whole waves
    const int t = tm * tn256;
    stream_k = (t % sms == 0) ? 0 : 1;
    k_splits = 1;
  }
  if (k_splits < 1) k_splits = 1;
  if (k_splits > kblocks) k_splits = kblocks;
  {  // no empty splits
    const int per = (kblocks + k_splits - 1) / k_splits;
    k_splits = (kblocks + per - 1) / per;
  }
  const bool partial = (k_splits > 1) || stream_k;
  if (partial && !accumulate) {
    DAE_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)n_store * sizeof(float), M, st));
    if (special_col >= 0 && special_out) DAE_CUDA(cudaMemsetAsync(special_out, 0, sizeof(float) * M, st));
  }
  GemmParams p{};
  p.M = M; p.N = N; p.K = K; p.k_splits = k_splits; p.stream_k = stream_k; p.atomic = (partial || accumulate) ? 1 : 0; p.alpha = alpha;
  p.C = C; p.ldc = ldc; p.n_store = n_store;
  p.special_col = (special_out ? special_col : -1); p.special_out = special_out;
  Operand A{a_hi, a_lo, lda, a_mn_major}, B{b_hi, b_lo, ldb, b_mn_major};
  // 128 x 256 tiles move fewer operand bytes per output, 128 x 128 tiles quantise better onto the SMs: pick the variant with the
  // smaller (waves x relative tile cost).  Stream-K balances by construction, so it always takes the 128 x 256 tiles.
  int rc;
  const int tiles256 = tm * tn256 * k_splits, tiles128 = tm * tn128 * k_splits;
  const float cost256 = 2.0f * (float)((tiles256 + sms - 1) / sms), cost128 = 1.1f * (float)((tiles128 + sms - 1) / sms);
  if (use_pair(M, N, K)) rc = g_lean ? launch_gemm<256, 2, EPI_STORE, 0, 0, 1>(A, B, p, st) : launch_gemm<256, 3, EPI_STORE, 0, 0, 1>(A, B, p, st);
  else if (!stream_k && cost128 < cost256) rc = launch_gemm<128, 3, EPI_STORE, 0, 0, 0>(A, B, p, st);
  else rc = launch_gemm<256, 2, EPI_STORE, 0, 0, 0>(A, B, p, st);
  if (rc) return rc;
  DAE_CHECK_LAUNCH("dae_gemm_bf16x3");
  return DAE_OK;
}

// C[m, n] (+)= alpha * sum_k (G[m, k] + G[k, m]) * B[k, n] for a square G [M x M] (bf16 hi / lo, row-major, ld = ldg) and B stored
// [M x ldb] row-major (n contiguous).  ONE launch: the k loop runs over G's columns and then over G's rows (the same array through
// an M-contiguous tensor map), so G + G^T is never formed.  dE2 = alpha (G + G^T) E of the batch_all / batch_hard backward.
extern "C" int dae_gemm_sym_bf16x3(int32_t M, int32_t N, float alpha, const void* g_hi, const void* g_lo, int64_t ldg, const void* b_hi,
                                   const void* b_lo, int64_t ldb, float* C, int64_t ldc, int32_t accumulate, void* stream) {
  DAE_REQUIRE(g_hi && g_lo && b_hi && b_lo && C && M > 0 && N > 0, "dae_gemm_sym_bf16x3: bad arguments");
  DAE_REQUIRE(ldg % 8 == 0 && ldb % 8 == 0 && ldg >= M && ldb >= N, "dae_gemm_sym_bf16x3: leading dimensions must be multiples of 8 and cover the matrices");
  DAE_REQUIRE(((uintptr_t)g_hi | (uintptr_t)g_lo | (uintptr_t)b_hi | (uintptr_t)b_lo) % 16 == 0, "dae_gemm_sym_bf16x3: operands must be 16-byte aligned");
  GemmParams p{};
  const int kb_half = (M + BLOCK_K - 1) / BLOCK_K;
  // stream-K: 28 tiles of 128 x 128 at B = 800 would leave 120 SMs idle for a 26-k-block main loop; ~6 k-blocks per CTA instead
  p.M = M; p.N = N; p.K = 2 * kb_half * BLOCK_K; p.k_splits = 1; p.stream_k = 1; p.atomic = 1; p.alpha = alpha;
  p.C = C; p.ldc = ldc; p.n_store = N; p.special_col = -1; p.special_out = nullptr; p.a_sym_kb = kb_half;
  if (!accumulate) DAE_CUDA(cudaMemset2DAsync(C, ldc * sizeof(float), 0, (size_t)N * sizeof(float), M, (cudaStream_t)stream));
  Operand A{g_hi, g_lo, ldg, 0}, B{b_hi, b_lo, ldb, 1};
  int rc = launch_gemm<128, 3, EPI_STORE, 0, 0, 0>(A, B, p, (cudaStream_t)stream);
  if (rc) return rc;
  DAE_CHECK_LAUNCH("dae_gemm_sym_bf16x3");
  return DAE_OK;
}

extern "C" int dae_decode_prepare(int32_t Brows, int32_t F, const int64_t* indptr, const int32_t* indices, const int32_t* rows,
                                  float* row_loss_part, int32_t* tile_ptr, void* stream) {
  DAE_REQUIRE(Brows > 0 && F > 0 && indptr && indices && row_loss_part && tile_ptr, "dae_decode_prepare: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  DAE_CUDA(cudaMemsetAsync(row_loss_part, 0, sizeof(float) * Brows, st));
  const int width = decode_tile_width();                              // column tile of the fused decode kernel that will consume this
  const int n_half = (kEwDecode / 4) * ((F + width - 1) / width);
  dim3 grid((n_half + 1 + 127) / 128, Brows);
  decode_tile_ptr_kernel<<<grid, 128, 0, st>>>(indptr, indices, rows, Brows, n_half, width / (kEwDecode / 4), tile_ptr);
  DAE_CHECK_LAUNCH("dae_decode_prepare");
  return DAE_OK;
}

extern "C" int dae_decode_fused_bf16x3(int32_t Brows, int32_t F, int32_t K, const void* e_hi, const void* e_lo, int64_t lde,
                                       const void* w_hi, const void* w_lo, int64_t ldw, const int64_t* indptr, const int32_t* indices,
                                       const float* values, const int32_t* rows, const float* bv, int32_t dec_act, int32_t loss_func,
                                       const float* weight, const double* stats, void* dz_hi, void* dz_lo, int64_t ld_dz,
                                       float* row_loss_part, int32_t* tile_ptr, int32_t prepared, void* stream) {
  DAE_REQUIRE(e_hi && e_lo && w_hi && w_lo && indptr && indices && values && bv && stats && dz_hi && dz_lo && row_loss_part && tile_ptr,
              "dae_decode_fused_bf16x3: null pointer");
  DAE_REQUIRE(loss_func == DAE_LOSS_CE || loss_func == DAE_LOSS_MSE, "dae_decode_fused_bf16x3: cosine loss uses the unfused path");
  DAE_REQUIRE(lde % 8 == 0 && ldw % 8 == 0 && ld_dz % 32 == 0 && ld_dz >= F, "dae_decode_fused_bf16x3: bad leading dimensions");
  cudaStream_t st = (cudaStream_t)stream;
  GemmParams p{};
  p.M = Brows; p.N = F; p.K = K; p.k_splits = 1; p.alpha = 1.0f; p.special_col = -1;
  p.indptr = indptr; p.indices = indices; p.values = values; p.rows = rows; p.bv = bv; p.weight = weight; p.stats = stats;
  p.dz_hi = (__nv_bfloat16*)dz_hi; p.dz_lo = (__nv_bfloat16*)dz_lo; p.ld_dz = ld_dz; p.row_loss_part = row_loss_part;
  p.tile_ptr = tile_ptr;
  if (!prepared) {   // dae_decode_prepare not issued by the caller (e.g. on a parallel graph branch): do it in line
    int rc0 = dae_decode_prepare(Brows, F, indptr, indices, rows, row_loss_part, tile_ptr, stream);
    if (rc0) return rc0;
  }
  Operand A{e_hi, e_lo, lde, 0}, B{w_hi, w_lo, ldw, 0};
  int rc = 0;
  // measured at C2 (800 x 10000 x 500): the pair kernel's 3 rounds of 160 pair tiles on 74 SM pairs lose to 2 rounds of 280
  // single-CTA tiles (56 vs 47 us) -- the fused epilogue, not operand traffic, bounds this kernel; pairs only when forced (tests)
  const bool pair = (g_pair_mode == 1);
  // lean: 128 x 128 tiles, 2 stages = 128 KB of operand ring instead of 192 KB, so that other kernels can share the SM.  Measured at
  // C2 (profiles/README.md): 51 vs 47.6 us standalone, and co-residency does not pay -- a sweep sharing SMs with the GEMMs slows both
  // (shared-memory bandwidth is the tensor cores' operand path) -- so the deep rings are the default
#define DAE_DEC(ACT, LOSS) rc = pair ? launch_gemm<256, 3, EPI_DECODE, ACT, LOSS, 1>(A, B, p, st) \
                              : (g_lean ? launch_gemm<128, 2, EPI_DECODE, ACT, LOSS, 0>(A, B, p, st) : launch_gemm<256, 2, EPI_DECODE, ACT, LOSS, 0>(A, B, p, st))
  if (loss_func == DAE_LOSS_CE) {
    if (dec_act == DAE_ACT_SIGMOID) DAE_DEC(DAE_ACT_SIGMOID, DAE_LOSS_CE);
    else if (dec_act == DAE_ACT_TANH) DAE_DEC(DAE_ACT_TANH, DAE_LOSS_CE);
    else DAE_DEC(DAE_ACT_NONE, DAE_LOSS_CE);
  } else {
    if (dec_act == DAE_ACT_SIGMOID) DAE_DEC(DAE_ACT_SIGMOID, DAE_LOSS_MSE);
    else if (dec_act == DAE_ACT_TANH) DAE_DEC(DAE_ACT_TANH, DAE_LOSS_MSE);
    else DAE_DEC(DAE_ACT_NONE, DAE_LOSS_MSE);
  }
#undef DAE_DEC
  if (rc) return rc;
  DAE_CHECK_LAUNCH("dae_decode_fused_bf16x3");
  return DAE_OK;
}
