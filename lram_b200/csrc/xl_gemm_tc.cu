// Tensor-core Linear for sm_100a: out[M,N] = A[M,K] * W[N,K]^T (+bias) (+residual)
//
//   * W is bf16 [N,K] (K contiguous) exactly as nn.Linear stores it -> UMMA "B" operand, K-major.
//   * A is fp32 in the model; it arrives here split into two bf16 planes, A = A_hi + A_lo
//     (A_hi = bf16(A), A_lo = bf16(A - A_hi)), and the kernel accumulates A_hi*W^T + A_lo*W^T into the same
//     fp32 TMEM accumulator. bf16 x bf16 products are exact in fp32, so the result carries ~16 mantissa bits
//     of the activations (vs 8 for a plain bf16 cast) — what keeps hidden states within 1e-3 and action
//     tokens bit-exact against the fp32 oracle while still running on tcgen05.
//   * One CTA computes a 128 x BLOCK_N tile: warp 0 = TMA producer (cp.async.bulk.tensor, 128B swizzle),
//     warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue
//     (tcgen05.ld -> registers -> bias/residual -> global). kStages-deep smem ring with mbarriers.
//
// Descriptor formats follow CUTLASS cute/arch/mma_sm100_desc.hpp (SmemDescriptor / InstrDescriptor).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

namespace tc {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = 128 B = one swizzle row
constexpr int UMMA_K = 16;
constexpr int kThreads = 192;

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
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
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major, 128B-swizzled operand tile: rows of 128 B, 8-row atoms of 1024 B (SBO = 1024 B), LBO unused (=1),
// descriptor version 1 (Blackwell), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);   // start address, bits [0,14)
  d |= (uint64_t)1 << 16;                        // leading byte offset (ignored for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;              // stride byte offset, bits [32,46)
  d |= (uint64_t)1 << 46;                        // version = 1
  d |= (uint64_t)2 << 61;                        // SWIZZLE_128B
  return d;
}

// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M = 128, N = BLOCK_N.
__host__ __device__ constexpr uint32_t make_idesc(int n) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                          uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}

template <int BLOCK_N, int kStages>
struct Smem {
  static constexpr int kABytes = BLOCK_M * BLOCK_K * 2;
  static constexpr int kBBytes = BLOCK_N * BLOCK_K * 2;
  static constexpr int kStageBytes = 2 * kABytes + kBBytes;
  static constexpr int kTotal = kStages * kStageBytes + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int kStages>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_a_hi, const __grid_constant__ CUtensorMap map_a_lo,
               const __grid_constant__ CUtensorMap map_w, const float* __restrict__ bias,
               const float* __restrict__ residual, float* __restrict__ out, int M, int N, int K,
               long long split_stride) {
  using S = Smem<BLOCK_N, kStages>;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * S::kStageBytes;     // full[kStages], empty[kStages], tmem_full
  const uint32_t tmem_slot = bar_base + 8 * (2 * kStages + 1);
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (kStages + s); };
  const uint32_t tmem_full_bar = bar_base + 8 * (2 * kStages);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BLOCK_N, m0 = blockIdx.y * BLOCK_M;
  // split-K: blockIdx.z owns k-blocks [kb0, kb0 + num_kb) and writes its raw partial tile to the plane
  // out + blockIdx.z * split_stride; the CONSUMER kernels (LayerNorm, conv/qkv, finalize) add the planes in
  // plane order, so there is no reduction step, no atomics and no inter-CTA synchronisation here.
  const int splits = gridDim.z;
  const int kb_total = K / BLOCK_K;
  const int kb_per = (kb_total + splits - 1) / splits;
  const int kb0 = blockIdx.z * kb_per;
  const int num_kb = max(0, min(kb_total, kb0 + kb_per) - kb0);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_hi) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a_lo) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    mbar_init(tmem_full_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    // allocate BLOCK_N (power of two >= 32) TMEM columns; address is written to shared memory
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(BLOCK_N)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");

  // The weight tiles of the first ring round are requested BEFORE the dependency wait (weights are never
  // written by a kernel): they stream in while the kernel that produces the activations is still running.
  const int pre_kb = num_kb < kStages ? num_kb : kStages;
  if (warp == 0 && lane == 0) {
    for (int kb = 0; kb < pre_kb; ++kb) {
      mbar_expect_tx(full_bar(kb), S::kStageBytes);
      tma_load_2d(base + kb * S::kStageBytes + 2 * S::kABytes, &map_w, full_bar(kb), (kb0 + kb) * BLOCK_K, n0);
    }
  }
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      for (int kb = 0; kb < pre_kb; ++kb) {
        const uint32_t st = base + kb * S::kStageBytes;
        tma_load_2d(st, &map_a_hi, full_bar(kb), (kb0 + kb) * BLOCK_K, m0);
        tma_load_2d(st + S::kABytes, &map_a_lo, full_bar(kb), (kb0 + kb) * BLOCK_K, m0);
      }
      for (int kb = pre_kb; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t st = base + s * S::kStageBytes;
        mbar_expect_tx(full_bar(s), S::kStageBytes);
        tma_load_2d(st, &map_a_hi, full_bar(s), (kb0 + kb) * BLOCK_K, m0);
        tma_load_2d(st + S::kABytes, &map_a_lo, full_bar(s), (kb0 + kb) * BLOCK_K, m0);
        tma_load_2d(st + 2 * S::kABytes, &map_w, full_bar(s), (kb0 + kb) * BLOCK_K, n0);
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc(BLOCK_N);
      for (int kb = 0; kb < num_kb; ++kb) {
        const int s = kb % kStages;
        const uint32_t ph = (kb / kStages) & 1;
        mbar_wait(full_bar(s), ph);
        tcgen05_fence_after();
        const uint32_t st = base + s * S::kStageBytes;
        const uint64_t a_hi = make_smem_desc(st);
        const uint64_t a_lo = make_smem_desc(st + S::kABytes);
        const uint64_t bw = make_smem_desc(st + 2 * S::kABytes);
#pragma unroll
        for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
          const uint64_t kofs = (uint64_t)((k * UMMA_K * 2) >> 4);   // advance inside the 128-B swizzle row
          umma_bf16(tmem_base, a_hi + kofs, bw + kofs, idesc, (kb | k) != 0);
          umma_bf16(tmem_base, a_lo + kofs, bw + kofs, idesc, 1u);
        }
        tcgen05_commit(empty_bar(s));            // frees the smem stage when these MMAs retire
      }
      tcgen05_commit(tmem_full_bar);             // accumulator complete
    }
  }

  // ===== epilogue: warps 2..5; a warp may only touch TMEM lanes 32*(warp%4) .. +31 =====================
  // TMEM -> registers -> (bias, residual: only when splits == 1, host-enforced) -> global, 32 columns at a time.
  const bool is_epi = warp >= 2;
  const int q = warp & 3;
  const int row = m0 + q * 32 + lane;
  out += (long long)blockIdx.z * split_stride;

  auto store_chunk = [&](const float (&v)[32], int c) {       // bias + residual + store of 32 columns
    if (row >= M) return;
    const int col0 = n0 + c;
    float* orow = out + (int64_t)row * N + col0;
    const float* rrow = residual ? residual + (int64_t)row * N + col0 : nullptr;
    if (col0 + 32 <= N && (N & 3) == 0) {
#pragma unroll
      for (int j = 0; j < 32; j += 4) {
        float4 o = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
        if (bias) {
          const float4 b4 = *reinterpret_cast<const float4*>(bias + col0 + j);
          o.x += b4.x; o.y += b4.y; o.z += b4.z; o.w += b4.w;
        }
        if (rrow) {
          const float4 r4 = *reinterpret_cast<const float4*>(rrow + j);
          o.x += r4.x; o.y += r4.y; o.z += r4.z; o.w += r4.w;
        }
        *reinterpret_cast<float4*>(orow + j) = o;
      }
    } else {
#pragma unroll
      for (int j = 0; j < 32; ++j) {
        if (col0 + j < N) {
          float o = v[j];
          if (bias) o += bias[col0 + j];
          if (rrow) o += rrow[j];
          orow[j] = o;
        }
      }
    }
  };
  auto tmem_load_chunk = [&](float (&v)[32], int c) {
    uint32_t r[32];
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]),
          "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]),
          "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  };

  if (is_epi) {
    mbar_wait(tmem_full_bar, 0);
    tcgen05_fence_after();
    if (num_kb > 0) {
#pragma unroll
      for (int c = 0; c < BLOCK_N; c += 32) {
        float v[32];
        tmem_load_chunk(v, c);
        store_chunk(v, c);
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(BLOCK_N) : "memory");
  }
}

// fp32 rows -> bf16 hi/lo planes (dense [rows, K])
__global__ void __launch_bounds__(256) split_bf16_kernel(const float* __restrict__ in, int64_t in_stride,
                                                         __nv_bfloat16* __restrict__ hi,
                                                         __nv_bfloat16* __restrict__ lo, int rows, int K) {
  const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
  pdl_wait();
  pdl_trigger();
  if (i >= (int64_t)rows * K) return;
  const int r = (int)(i / K), c = (int)(i - (int64_t)r * K);
  const float4 v = *reinterpret_cast<const float4*>(in + (int64_t)r * in_stride + c);
  const float x[4] = {v.x, v.y, v.z, v.w};
  __nv_bfloat16 h[4], l[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    h[j] = __float2bfloat16_rn(x[j]);
    l[j] = __float2bfloat16_rn(x[j] - __bfloat162float(h[j]));
  }
  *reinterpret_cast<uint2*>(hi + i) = *reinterpret_cast<uint2*>(h);
  *reinterpret_cast<uint2*>(lo + i) = *reinterpret_cast<uint2*>(l);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  // function-local static: initialised once, thread-safe (C++11); the driver entry point is process-wide
  static const EncodeTiledFn fn = [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      return (EncodeTiledFn)p;
    return (EncodeTiledFn) nullptr;
  }();
  return fn;
}

// bf16 row-major [rows, K] -> 2D map, box = {BLOCK_K, box_rows}, 128B swizzle, OOB reads give zeros
static bool make_map(CUtensorMap* m, const void* ptr, int rows, int K, int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[2] = {(cuuint64_t)K, (cuuint64_t)rows};
  cuuint64_t gstride[1] = {(cuuint64_t)K * 2};
  cuuint32_t box[2] = {(cuuint32_t)BLOCK_K, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <int BLOCK_N, int kStages>
static cudaError_t launch(const CUtensorMap& ma, const CUtensorMap& ml, const CUtensorMap& mw, const float* bias,
                          const float* residual, float* out, int M, int N, int K, int splits, long long split_stride,
                          cudaStream_t s) {
  using S = Smem<BLOCK_N, kStages>;
  if (cudaError_t e = ensure_dyn_smem<&gemm_tc_kernel<BLOCK_N, kStages>>(S::kTotal); e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((N + BLOCK_N - 1) / BLOCK_N, (M + BLOCK_M - 1) / BLOCK_M, splits);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = S::kTotal;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  int na = 0;
  if (g_use_pdl) {
    attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[na].val.programmaticStreamSerializationAllowed = 1;
    ++na;
  }
  cfg.attrs = attr;
  cfg.numAttrs = na;
  return cudaLaunchKernelEx(&cfg, gemm_tc_kernel<BLOCK_N, kStages>, ma, ml, mw, bias, residual, out, M, N, K,
                            split_stride);
}

}  // namespace tc

bool gemm_tc_supported(int M, int N, int K) { return K % tc::BLOCK_K == 0 && K >= tc::BLOCK_K && M >= 1 && N >= 8; }

void launch_split_bf16(const float* in, int64_t in_stride, void* hi, void* lo, int rows, int K, cudaStream_t s) {
  const int64_t n4 = (int64_t)rows * K / 4;
  launch_k(tc::split_bf16_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, in, in_stride,
           (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, rows, K);
}

// Tile width and split-K factor from a small cost model. These skinny GEMMs (M = envs x tokens, a few hundred
// rows) are bound by the rate at which ONE SM can pull operand tiles out of L2 (~80 GB/s per SM through TMA,
// ~6.3 KB/clk chip-wide), not by the tensor pipe and not by L2 bandwidth: every CTA re-reads the A planes of
// its row tile. So the goal is to put ~one CTA on every SM with as few operand bytes per CTA as possible:
// wide tiles (fewer A re-reads) x split-K (more CTAs). A split-K CTA writes its partial tile to its own
// plane; the consumer kernels add the planes (costed below at L2 speed).
void gemm_tc_plan(int M, int N, int K, int num_sms, int max_splits, int* bn_out, int* splits_out) {
  const int m_tiles = (M + tc::BLOCK_M - 1) / tc::BLOCK_M;
  const int kb = K / tc::BLOCK_K;
  double best = 1e30;
  int best_bn = 64, best_sp = 1;
  const int bns[3] = {128, 64, 32};
  for (int bi = 0; bi < 3; ++bi) {
    const int bn = bns[bi];
    const int n_tiles = (N + bn - 1) / bn;
    for (int sp = 1; sp <= (max_splits < 1 ? 1 : max_splits); ++sp) {
      if (kb % sp) continue;
      const int kbp = kb / sp;
      const int ctas = m_tiles * n_tiles * sp;
      const int waves = (ctas + num_sms - 1) / num_sms;
      const double fill_kb = kbp * (2.0 * 16.0 + bn * 0.125);                // operand KB through TMA per CTA
      const double store_kb = 128.0 * bn * 4.0 / 1024.0;                     // output tile written per CTA
      double t = waves * (2.5 + fill_kb / 80.0 + store_kb / 80.0);           // us
      // consumers re-read sp planes; they are latency-bound kernels, so the extra loads cost as if at ~1.5 TB/s
      // (fitted to the B200 sweep in profiles/r01_gemm_splitk.md)
      if (sp > 1) t += 0.6 + (double)(sp - 1) * M * N * 4.0 / 1.5e6;
      if (t < best) { best = t; best_bn = bn; best_sp = sp; }
    }
  }
  *bn_out = best_bn;
  *splits_out = best_sp;
}

// bn: 128 / 64 / 32 (0 = plan it, without split-K). splits > 1: raw partial tiles go to out + z * split_stride
// (bias and residual must be null; the consumer adds the planes).
cudaError_t launch_gemm_tc(const void* a_hi, const void* a_lo, const __nv_bfloat16* W, const float* bias,
                           const float* residual, float* out, int M, int N, int K, int num_sms, int bn, int splits,
                           long long split_stride, int low_smem, cudaStream_t s) {
  if (!gemm_tc_supported(M, N, K)) return cudaErrorInvalidValue;
  if (splits < 1) splits = 1;
  if (splits > 1 && (bias || residual || (K / tc::BLOCK_K) % splits)) return cudaErrorInvalidValue;
  if (bn != 128 && bn != 64 && bn != 32) {
    int sp;
    gemm_tc_plan(M, N, K, num_sms, 1, &bn, &sp);
  }
  CUtensorMap ma, ml, mw;
  if (!tc::make_map(&ma, a_hi, M, K, tc::BLOCK_M) || !tc::make_map(&ml, a_lo, M, K, tc::BLOCK_M) ||
      !tc::make_map(&mw, W, N, K, bn))
    return cudaErrorUnknown;
  if (low_smem) {
    // shallow rings (<= 110 KB): the CTA must fit beside a resident state-stream CTA of another micro-batch
    switch (bn) {
      case 128: return tc::launch<128, 2>(ma, ml, mw, bias, residual, out, M, N, K, splits, split_stride, s);
      case 64: return tc::launch<64, 2>(ma, ml, mw, bias, residual, out, M, N, K, splits, split_stride, s);
      default: return tc::launch<32, 3>(ma, ml, mw, bias, residual, out, M, N, K, splits, split_stride, s);
    }
  }
  switch (bn) {
    case 128: return tc::launch<128, 4>(ma, ml, mw, bias, residual, out, M, N, K, splits, split_stride, s);
    case 64: return tc::launch<64, 4>(ma, ml, mw, bias, residual, out, M, N, K, splits, split_stride, s);
    default: return tc::launch<32, 6>(ma, ml, mw, bias, residual, out, M, N, K, splits, split_stride, s);
  }
}

}  // namespace xl
