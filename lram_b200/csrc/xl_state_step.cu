// mLSTM recurrent state step — the HBM-bound kernel of the path.
//
// Replaces, per block and per env step, what the reference runs as ~17 ATen kernels per token
// ([ext-xlstm] mLSTMCell.step -> recurrent_step_stabilized_simple + MultiHeadLayerNorm, reached from
// src/algos/models/decision_xlstm.py:163) plus the output gating of mLSTMLayer.step:
//
//   for t in 0..T-1 (the T tokens of one env step, in order):
//     lf = logsigmoid(f~_t);  m' = max(lf + m, i~_t);  f = exp(lf + m - m');  i = exp(i~_t - m')
//     C  = f*C + i*(k_t/sqrt(DH)) (x) v_t          C[dk, dv], fp32, resident in HBM
//     n  = f*n + i*(k_t/sqrt(DH))
//     h_t = (q_t^T C) / (max(|q_t^T n|, exp(-m')) + eps)
//     out_t = (GroupNorm_head(h_t)*(1+w) + skip*a_t) * silu(z_t)
//
// Kernel 1 (stream): touches every C element exactly once (one 128-bit read, T fused multiply-adds + T
// dot-product FMAs in registers, one 128-bit write): algorithmic bytes = 8*NH*DH^2 per env regardless of T.
//   CTA = (env b, head h, row chunk rs, column slab cs). A thread owns 4 consecutive columns (dv) and walks
//   rows (dk); the q^T C reduction over rows is thread-local in the loop, crosses warps once through shared
//   memory, and leaves the CTA as a partial numerator [T, slab]. Nothing else: no atomics, no fences, so a CTA
//   retires while its stores are still draining.
//     impl 1 (default): a producer warp keeps a ring of [32 rows x 128 cols] fp32 tiles of C in flight with
//       TMA (cp.async.bulk.tensor.2d + mbarrier, L2 evict-first) and bulk-copies the step's (q,k) pairs;
//       8 consumer warps pull their rows from shared memory into registers, release the slot at once, do the
//       FMAs and write C back with streaming 128-bit stores.
//     impl 0: plain 128-bit global loads with register batching (comparison baseline).
// Kernel 2 (finalize): one CTA per (env, head): sums the row-chunk partials in fixed order (deterministic),
//   updates n and m, divides by the stabilised denominator, applies the multi-head GroupNorm, the learnable
//   skip and the output gate, and emits the bf16 hi/lo planes the proj_down tensor-core GEMM consumes.
#include <cuda.h>
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

constexpr int kThreads = 256;        // consumer threads
constexpr int kUnroll = 8;
constexpr int kMaxNCH = 16;

// Gate recurrence of one (env, head) for the T tokens of the step. Called by one full warp; results in shared
// memory. Streaming CTAs and the finalize CTA run this same code on the same inputs -> identical f, i, m.
template <int T>
__device__ __forceinline__ void compute_gates(const StateStepParams& p, int b, int hd, int bh, float* s_f,
                                              float* s_i, float* s_m, float* s_pre) {
  const int lane = threadIdx.x & 31;
  const int NH = p.NH;
  if (lane < 2 * T) {
    const int t = lane >> 1, is_f = lane & 1;
    const float* gp = p.gate_part + ((int64_t)b * T + t) * p.NCH * 2 * NH + (is_f ? NH : 0) + hd;
    float s = 0.f;
    for (int c = 0; c < p.NCH; ++c) s += gp[c * 2 * NH];       // fixed order
    const float* bias = is_f ? p.fgate_b : p.igate_b;
    if (bias) s += bias[hd];
    s_pre[lane] = s;
  }
  __syncwarp();
  if (lane == 0) {
    float mprev = p.m[bh];
    s_m[0] = mprev;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float ig = s_pre[2 * t], fg = s_pre[2 * t + 1];
      const float lf = log_sigmoid(fg);
      const float mnew = fmaxf(lf + mprev, ig);
      s_f[t] = expf(lf + mprev - mnew);
      s_i[t] = expf(ig - mnew);
      s_m[t + 1] = mnew;
      mprev = mnew;
    }
  }
  __syncwarp();
}

// One row of 4 columns through the T tokens: c <- f_t*c + k_t[r] * vi_t ; acc_t += q_t[r] * c
// with vi_t = v_t * i_t / sqrt(DH) pre-scaled per thread. qk = (q,k) pair of token 0 for this row; token t's
// pair is tstride floats further.
template <int T>
__device__ __forceinline__ void row_update(float (&c)[4], const float* __restrict__ qk, int tstride,
                                           const float (&f)[T], const float (&vi)[T][4], float (&acc)[T][4]) {
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float2 q2 = *reinterpret_cast<const float2*>(qk + t * tstride);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      c[j] = fmaf(f[t], c[j], q2.y * vi[t][j]);
      acc[t][j] = fmaf(q2.x, c[j], acc[t][j]);
    }
  }
}

struct TileCoord {
  int bh, b, hd, rs, cs, r0, nrows, c0, rows_per;
};
__device__ __forceinline__ TileCoord tile_coord(const StateStepParams& p) {
  TileCoord tc;
  const int CS = p.DH / p.cols_per_cta;
  const int tiles_per_head = CS * p.rows_split;
  tc.bh = blockIdx.x / tiles_per_head;
  const int tile = blockIdx.x - tc.bh * tiles_per_head;
  tc.rs = tile / CS;
  tc.cs = tile - tc.rs * CS;
  tc.b = tc.bh / p.NH;
  tc.hd = tc.bh - tc.b * p.NH;
  tc.rows_per = (p.DH + p.rows_split - 1) / p.rows_split;
  tc.r0 = tc.rs * tc.rows_per;
  tc.nrows = max(0, min(p.DH, tc.r0 + tc.rows_per) - tc.r0);
  tc.c0 = tc.cs * p.cols_per_cta;
  return tc;
}

// cross-row-lane reduction of the per-thread numerators -> partial[bh][rs][t][c0 + c]
template <int T>
__device__ __forceinline__ void write_partials(const StateStepParams& p, const TileCoord& tc, float* sacc,
                                               const float (&acc)[T][4], int tx, int ty, int TX, int TY,
                                               int tid) {
#pragma unroll
  for (int t = 0; t < T; ++t)
    *reinterpret_cast<float4*>(sacc + ((ty * T + t) * TX + tx) * 4) =
        make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
  const int W = p.cols_per_cta;
  for (int idx = tid; idx < T * W; idx += kThreads) {
    const int t = idx / W, c = idx - t * W;
    float s = 0.f;
    for (int y = 0; y < TY; ++y) s += sacc[(y * T + t) * W + c];
    p.partial[(((int64_t)tc.bh * p.rows_split + tc.rs) * T + t) * p.DH + tc.c0 + c] = s;
  }
}

// ------------------------------------------------------------------------------------------------
// impl 0: register-batched global loads
// ------------------------------------------------------------------------------------------------
template <int T>
__global__ void __launch_bounds__(kThreads, 3) mlstm_state_stream_ldg_kernel(StateStepParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_f[T], s_i[T], s_m[T + 1], s_pre[2 * T];
  const TileCoord tc = tile_coord(p);
  const int DH = p.DH;
  const int TX = p.cols_per_cta >> 2, TY = kThreads / TX;
  const int tid = threadIdx.x, tx = tid % TX, ty = tid / TX;
  const int tstride = 2 * tc.rows_per;
  float* sqk = smem;                         // [T][rows_per][2]
  float* sacc = smem + T * tstride;          // [TY][T][cols_per_cta]

  if (tid < 32) compute_gates<T>(p, tc.b, tc.hd, tc.bh, s_f, s_i, s_m, s_pre);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float2* src = reinterpret_cast<const float2*>(p.qk) + (((int64_t)tc.b * T + t) * p.NH + tc.hd) * DH + tc.r0;
    for (int r = tid; r < tc.nrows; r += kThreads) reinterpret_cast<float2*>(sqk + t * tstride)[r] = src[r];
  }
  float vi[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float4 vv = *reinterpret_cast<const float4*>(p.v + ((int64_t)tc.b * T + t) * p.inner + tc.hd * DH +
                                                       tc.c0 + 4 * tx);
    vi[t][0] = vv.x; vi[t][1] = vv.y; vi[t][2] = vv.z; vi[t][3] = vv.w;
  }
  __syncthreads();
  const float kscale = rsqrtf((float)DH);
  float f[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    f[t] = s_f[t];
    const float it = s_i[t] * kscale;
#pragma unroll
    for (int j = 0; j < 4; ++j) vi[t][j] *= it;
  }
  float acc[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  float4* Cp = reinterpret_cast<float4*>(p.C + ((int64_t)tc.bh * DH + tc.r0) * DH + tc.c0) + tx;
  const int64_t rs4 = DH >> 2;               // row stride in float4
  int rbase = ty;
  for (; rbase + (kUnroll - 1) * TY < tc.nrows; rbase += TY * kUnroll) {   // full batches: no predicates
    float4 cv[kUnroll];
    float4* rp = Cp + (int64_t)rbase * rs4;
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) cv[u] = ld_stream(rp + (int64_t)u * TY * rs4);
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      float c[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w};
      row_update<T>(c, sqk + 2 * (rbase + u * TY), tstride, f, vi, acc);
      st_stream(rp + (int64_t)u * TY * rs4, make_float4(c[0], c[1], c[2], c[3]));
    }
  }
  for (; rbase < tc.nrows; rbase += TY) {
    float4* rp = Cp + (int64_t)rbase * rs4;
    const float4 cv = ld_stream(rp);
    float c[4] = {cv.x, cv.y, cv.z, cv.w};
    row_update<T>(c, sqk + 2 * rbase, tstride, f, vi, acc);
    st_stream(rp, make_float4(c[0], c[1], c[2], c[3]));
  }
  write_partials<T>(p, tc, sacc, acc, tx, ty, TX, TY, tid);
}

// ------------------------------------------------------------------------------------------------
// impl 1: TMA-fed ring
// ------------------------------------------------------------------------------------------------
namespace tma {

constexpr int kStageRows = 32;
constexpr int kStages = 3;
constexpr int kConsumerWarps = kThreads / 32;
constexpr int kBlock = kThreads + 32;   // + producer warp

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!done);
}
__device__ __forceinline__ void bulk_copy_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(dst), "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}

template <int T, bool kStream>
__global__ void __launch_bounds__(kBlock, 3)
mlstm_state_stream_tma_kernel(const __grid_constant__ CUtensorMap mapC, StateStepParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ float s_f[T], s_i[T], s_m[T + 1], s_pre[2 * T];
  __shared__ __align__(8) uint64_t s_bar[2 * kStages + 1];

  const TileCoord tc = tile_coord(p);
  const int DH = p.DH;
  const int W = p.cols_per_cta;
  const int TX = W >> 2, TY = kThreads / TX;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int tstride = 2 * tc.rows_per;
  const int nblk = (tc.nrows + kStageRows - 1) / kStageRows;
  const uint32_t stage_bytes = (uint32_t)(kStageRows * W * sizeof(float));

  // dynamic smem: [ring: kStages x (32 x W) fp32 | 128-B aligned] [sqk: T x rows_per x 2]; the reduction
  // scratch aliases the ring once the loop is over
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  float* stage0 = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)));
  float* sqk = stage0 + (size_t)kStages * kStageRows * W;
  float* sacc = stage0;
  const uint32_t bar0 = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  const uint32_t qk_bar = bar0 + 8u * (2 * kStages);

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kConsumerWarps);
    }
    mbar_init(qk_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();

  if (warp == kConsumerWarps) {
    // ===== producer warp =====
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&mapC) : "memory");
      // the step's (q, k) pairs of this row chunk: T contiguous runs of nrows*8 bytes
      mbar_expect_tx(qk_bar, (uint32_t)(T * tc.nrows * 8));
#pragma unroll
      for (int t = 0; t < T; ++t)
        bulk_copy_g2s(smem_u32(sqk + t * tstride),
                      p.qk + ((((int64_t)tc.b * T + t) * p.NH + tc.hd) * DH + tc.r0) * 2,
                      (uint32_t)(tc.nrows * 8), qk_bar);
      uint64_t policy = 0;
      if (kStream) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
      const int grow0 = tc.bh * DH + tc.r0;
      for (int i = 0; i < nblk; ++i) {
        const int s = i % kStages;
        mbar_wait(empty_bar(s), ((i / kStages) & 1) ^ 1);
        mbar_expect_tx(full_bar(s), stage_bytes);
        const uint32_t dst = sbase + s * stage_bytes;
        if (kStream) {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
              "[%0], [%1, {%3, %4}], [%2], %5;"
              ::"r"(dst), "l"(&mapC), "r"(full_bar(s)), "r"(tc.c0), "r"(grow0 + i * kStageRows), "l"(policy)
              : "memory");
        } else {
          asm volatile(
              "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
              ::"r"(dst), "l"(&mapC), "r"(full_bar(s)), "r"(tc.c0), "r"(grow0 + i * kStageRows)
              : "memory");
        }
      }
    }
    return;   // consumers only use the named barrier 1 from here on
  }

  // ===== consumers =====
  const int tx = tid % TX, ty = tid / TX;
  if (warp == 0) compute_gates<T>(p, tc.b, tc.hd, tc.bh, s_f, s_i, s_m, s_pre);
  float vi[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float4 vv = *reinterpret_cast<const float4*>(p.v + ((int64_t)tc.b * T + t) * p.inner + tc.hd * DH +
                                                       tc.c0 + 4 * tx);
    vi[t][0] = vv.x; vi[t][1] = vv.y; vi[t][2] = vv.z; vi[t][3] = vv.w;
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");
  const float kscale = rsqrtf((float)DH);
  float f[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    f[t] = s_f[t];
    const float it = s_i[t] * kscale;
#pragma unroll
    for (int j = 0; j < 4; ++j) vi[t][j] *= it;
  }
  float acc[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  float4* Cp = reinterpret_cast<float4*>(p.C + ((int64_t)tc.bh * DH + tc.r0) * DH + tc.c0) + tx;
  const int64_t rs4 = DH >> 2;
  constexpr int kRowsPerThread = 4;                 // fast path: TY == 8 -> 32 rows / 8 row lanes
  const bool fast = (TY * kRowsPerThread == kStageRows);
  mbar_wait(qk_bar, 0);
  for (int i = 0; i < nblk; ++i) {
    const int s = i % kStages;
    const float* st = stage0 + (size_t)s * kStageRows * W;
    const int rblk = i * kStageRows;
    const int rcur = min(kStageRows, tc.nrows - rblk);
    mbar_wait(full_bar(s), (i / kStages) & 1);
    if (fast && rcur == kStageRows) {
      float4 cv[kRowsPerThread];
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j)
        cv[j] = *reinterpret_cast<const float4*>(st + (ty + j * 8) * W + 4 * tx);
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));     // the slot is free as soon as it sits in registers
      float4* rp = Cp + (int64_t)(rblk + ty) * rs4;
#pragma unroll
      for (int j = 0; j < kRowsPerThread; ++j) {
        float c[4] = {cv[j].x, cv[j].y, cv[j].z, cv[j].w};
        row_update<T>(c, sqk + 2 * (rblk + ty + j * 8), tstride, f, vi, acc);
        const float4 o = make_float4(c[0], c[1], c[2], c[3]);
        if (kStream) st_stream(rp + (int64_t)j * 8 * rs4, o);
        else rp[(int64_t)j * 8 * rs4] = o;
      }
    } else {
      // generic path (small heads / ragged last tile)
      for (int r = ty; r < rcur; r += TY) {
        const float4 cv = *reinterpret_cast<const float4*>(st + r * W + 4 * tx);
        float c[4] = {cv.x, cv.y, cv.z, cv.w};
        row_update<T>(c, sqk + 2 * (rblk + r), tstride, f, vi, acc);
        Cp[(int64_t)(rblk + r) * rs4] = make_float4(c[0], c[1], c[2], c[3]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(empty_bar(s));
    }
  }
  asm volatile("bar.sync 1, %0;" ::"n"(kThreads) : "memory");   // every warp is done with the ring
  write_partials<T>(p, tc, sacc, acc, tx, ty, TX, TY, tid);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* q = nullptr;
    cudaDriverEntryPointQueryResult r;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &q, cudaEnableDefault, &r) == cudaSuccess &&
        r == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)q;
  }
  return fn;
}

template <int T, bool kStream>
static cudaError_t launch(const StateStepParams& p, cudaStream_t s) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return cudaErrorUnknown;
  CUtensorMap map;
  cuuint64_t gdim[2] = {(cuuint64_t)p.DH, (cuuint64_t)p.B * p.NH * p.DH};
  cuuint64_t gstride[1] = {(cuuint64_t)p.DH * sizeof(float)};
  cuuint32_t box[2] = {(cuuint32_t)p.cols_per_cta, (cuuint32_t)kStageRows};
  cuuint32_t estr[2] = {1, 1};
  if (enc(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, p.C, gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
          CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) !=
      CUDA_SUCCESS)
    return cudaErrorInvalidValue;
  const int rows_per = (p.DH + p.rows_split - 1) / p.rows_split;
  const size_t ring = sizeof(float) * (size_t)kStages * kStageRows * p.cols_per_cta;
  const int TY = kThreads / (p.cols_per_cta / 4);
  const size_t red = sizeof(float) * (size_t)TY * T * p.cols_per_cta;       // aliases the ring
  const size_t smem = 128 + (ring > red ? ring : red) + sizeof(float) * (size_t)2 * T * rows_per;
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(mlstm_state_stream_tma_kernel<T, kStream>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  const int CS = p.DH / p.cols_per_cta;
  const int64_t grid = (int64_t)p.B * p.NH * CS * p.rows_split;
  mlstm_state_stream_tma_kernel<T, kStream><<<(unsigned)grid, kBlock, smem, s>>>(map, p);
  return cudaGetLastError();
}

}  // namespace tma

// ------------------------------------------------------------------------------------------------
// kernel 2: finalize one (env, head)
// Latency-bound (a few KB per CTA), so: every global load is issued up front, and the T tokens share each
// block reduction (3 reductions in total: q.n, mean, variance) instead of doing 3 per token.
// ------------------------------------------------------------------------------------------------
template <int N>
__device__ __forceinline__ void block_sum_n(float (&v)[N], float* red /* [N][32] */) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();                     // protect `red` from the previous use
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float r = (lane < nw) ? red[k * 32 + lane] : 0.f;
    v[k] = warp_sum(r);
  }
}

// blockDim = DH rounded up to a warp multiple (<= 1024): one head channel per thread, all T tokens. Straight
// and short on purpose: every CTA runs this code once, so its size is what the instruction cache sees.
template <int T>
__global__ void __launch_bounds__(1024) mlstm_state_finalize_kernel(StateStepParams p) {
  __shared__ float s_f[T], s_i[T], s_m[T + 1], s_pre[2 * T];
  __shared__ float s_red[T * 32];
  const int bh = blockIdx.x;
  const int b = bh / p.NH, hd = bh - b * p.NH;
  const int DH = p.DH, inner = p.inner, RS = p.rows_split;
  const int a = threadIdx.x;
  const bool ok = a < DH;
  const int ch = hd * DH + (ok ? a : 0);

  // ---- all global loads first ---------------------------------------------------------------------
  float nreg = ok ? p.n[(int64_t)bh * DH + a] : 0.f;
  const float wn = ok ? p.outnorm_w[ch] : 0.f;
  const float wskip = (ok && p.skip) ? p.skip[ch] : 0.f;
  float2 qk[T];
  float num[T], act[T], zz[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const int64_t row = (int64_t)b * T + t;
    qk[t] = ok ? reinterpret_cast<const float2*>(p.qk)[(row * p.NH + hd) * DH + a] : make_float2(0.f, 0.f);
    float s = 0.f;
    if (ok)
      for (int r = 0; r < RS; ++r) s += p.partial[(((int64_t)bh * RS + r) * T + t) * DH + a];   // fixed order
    num[t] = s;
    act[t] = (ok && p.skip) ? p.act[row * inner + ch] : 0.f;
    zz[t] = (ok && p.skip) ? p.u[row * 2 * inner + inner + ch] : 0.f;
  }
  if (threadIdx.x < 32) compute_gates<T>(p, b, hd, bh, s_f, s_i, s_m, s_pre);
  __syncthreads();

  // ---- n recurrence and q.n for the T tokens --------------------------------------------------------
  const float kscale = rsqrtf((float)DH);
  float qn[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    nreg = fmaf(s_f[t], nreg, s_i[t] * kscale * qk[t].y);
    qn[t] = qk[t].x * nreg;
  }
  block_sum_n<T>(qn, s_red);
  if (ok) p.n[(int64_t)bh * DH + a] = nreg;
  // ---- h = num / den, GroupNorm over the head, skip + output gate -----------------------------------
  float mean[T], var[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float den = fmaxf(fabsf(qn[t]), expf(-s_m[t + 1])) + p.cell_eps;
    num[t] = num[t] / den;                   // h (zero for the padded threads)
    mean[t] = num[t];
  }
  block_sum_n<T>(mean, s_red);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    mean[t] /= (float)DH;
    const float dlt = ok ? num[t] - mean[t] : 0.f;
    var[t] = dlt * dlt;
  }
  block_sum_n<T>(var, s_red);
  if (ok) {
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float rstd = rsqrtf(var[t] / (float)DH + p.ln_eps);
      const int64_t row = (int64_t)b * T + t;
      float o = (num[t] - mean[t]) * rstd * (1.f + wn);
      if (p.h_raw) p.h_raw[row * inner + ch] = num[t];
      if (p.skip) o = (o + wskip * act[t]) * silu_fast(zz[t]);
      if (p.out) p.out[row * inner + ch] = o;
      if (p.out_hi) {
        // A = hi + lo with hi = bf16(A), lo = bf16(A - hi): operand planes of the tcgen05 proj_down
        const __nv_bfloat16 hi = __float2bfloat16_rn(o);
        reinterpret_cast<__nv_bfloat16*>(p.out_hi)[row * inner + ch] = hi;
        reinterpret_cast<__nv_bfloat16*>(p.out_lo)[row * inner + ch] = __float2bfloat16_rn(o - __bfloat162float(hi));
      }
    }
  }
  if (threadIdx.x == 0) p.m[bh] = s_m[T];
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
void state_step_auto_tiling(int B, int NH, int DH, int num_sms, int* rows_split, int* cols_per_cta) {
  // column slab: widest of {128, 64, 32, 16, 8, 4} that divides DH (a warp then reads 512 contiguous B)
  int cols = 128;
  while (cols > 4 && (DH % cols) != 0) cols >>= 1;
  if (cols > DH) cols = DH;
  const int CS = DH / cols;
  // split rows (in multiples of the 32-row TMA tile) until there are enough CTAs to fill the machine
  const int64_t target = (int64_t)num_sms * 3;
  int rs = 1;
  while ((int64_t)B * NH * CS * rs < target && (DH % (rs * 2 * 32)) == 0 && rs < 32) rs *= 2;
  *rows_split = rs;
  *cols_per_cta = cols;
}

template <int T>
static cudaError_t launch_T(const StateStepParams& p, cudaStream_t s) {
  const int rows_per = (p.DH + p.rows_split - 1) / p.rows_split;
  cudaError_t e;
  if (p.impl == 1 && rows_per % 4 == 0) {
    // stream (evict-first) when the stack's C does not fit comfortably in L2; keep it cached when it does
    const size_t c_bytes = sizeof(float) * (size_t)p.B * p.NH * p.DH * p.DH;
    const bool stream = c_bytes * (size_t)(p.num_layers > 0 ? p.num_layers : 1) > ((size_t)64 << 20);
    e = stream ? tma::launch<T, true>(p, s) : tma::launch<T, false>(p, s);
  } else {
    const int TX = p.cols_per_cta / 4, TY = kThreads / TX;
    const size_t smem = sizeof(float) * ((size_t)2 * T * rows_per + (size_t)TY * T * p.cols_per_cta);
    static size_t attr_smem = 48 * 1024;
    if (smem > attr_smem) {
      e = cudaFuncSetAttribute(mlstm_state_stream_ldg_kernel<T>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)smem);
      if (e != cudaSuccess) return e;
      attr_smem = smem;
    }
    const int CS = p.DH / p.cols_per_cta;
    const int64_t grid = (int64_t)p.B * p.NH * CS * p.rows_split;
    mlstm_state_stream_ldg_kernel<T><<<(unsigned)grid, kThreads, smem, s>>>(p);
    e = cudaGetLastError();
  }
  return e;
}

static void resolve_tiling(StateStepParams& p, int num_sms) {
  if (p.rows_split <= 0 || p.cols_per_cta <= 0) {
    int rs, cols;
    state_step_auto_tiling(p.B, p.NH, p.DH, num_sms, &rs, &cols);
    if (p.rows_split <= 0) p.rows_split = rs;
    if (p.cols_per_cta <= 0) p.cols_per_cta = cols;
  }
}

// kernel 2; p must carry the same rows_split the stream kernel ran with (resolved here the same way)
cudaError_t launch_state_finalize(StateStepParams p, int num_sms, cudaStream_t s) {
  resolve_tiling(p, num_sms);
  const int thr = ((p.DH + 31) / 32) * 32;
  switch (p.T) {
    case 1: mlstm_state_finalize_kernel<1><<<p.B * p.NH, thr, 0, s>>>(p); break;
    case 2: mlstm_state_finalize_kernel<2><<<p.B * p.NH, thr, 0, s>>>(p); break;
    case 3: mlstm_state_finalize_kernel<3><<<p.B * p.NH, thr, 0, s>>>(p); break;
    case 4: mlstm_state_finalize_kernel<4><<<p.B * p.NH, thr, 0, s>>>(p); break;
    default: return cudaErrorInvalidValue;
  }
  return cudaGetLastError();
}

// kernel 1
cudaError_t launch_state_step(StateStepParams p, int num_sms, cudaStream_t s) {
  if (p.rows_split <= 0 || p.cols_per_cta <= 0) {
    int rs, cols;
    state_step_auto_tiling(p.B, p.NH, p.DH, num_sms, &rs, &cols);
    if (p.rows_split <= 0) p.rows_split = rs;
    if (p.cols_per_cta <= 0) p.cols_per_cta = cols;
  }
  if (p.cols_per_cta % 4 || p.DH % p.cols_per_cta || p.cols_per_cta > 256 ||
      kThreads % (p.cols_per_cta / 4) || p.DH > 1024 || p.NCH > kMaxNCH || p.rows_split > p.DH)
    return cudaErrorInvalidValue;
  switch (p.T) {
    case 1: return launch_T<1>(p, s);
    case 2: return launch_T<2>(p, s);
    case 3: return launch_T<3>(p, s);
    case 4: return launch_T<4>(p, s);
    default: return cudaErrorInvalidValue;
  }
}

// [M, 3, inner] (q | k | v) -> qk [M, NH, DH, 2] interleaved + v [M, inner]   (unit-test entry point only)
__global__ void repack_qkv_kernel(const float* __restrict__ qkv, float* __restrict__ qk, float* __restrict__ v,
                                  int M, int inner) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)M * inner) return;
  const int64_t m = i / inner;
  const int c = (int)(i - m * inner);
  const float* row = qkv + m * 3 * inner;
  reinterpret_cast<float2*>(qk)[i] = make_float2(row[c], row[inner + c]);
  v[i] = row[2 * inner + c];
}
void launch_repack_qkv(const float* qkv, float* qk, float* v, int M, int inner, cudaStream_t s) {
  const int64_t n = (int64_t)M * inner;
  repack_qkv_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(qkv, qk, v, M, inner);
}

}  // namespace xl
