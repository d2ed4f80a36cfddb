// Context prefill, chunkwise sequence cell on tcgen05 (SURVEY.md §8 row a11 / BASELINE.json configs[3]).
//
// Same contract as launch_cell_seq / launch_cell_mma (xl_prefill.cu, xl_prefill_mma.cu): given the per-token stabilised
// gates f_t, i_t of a run of S tokens and q, k, v, advance  C <- f_t C + (i_t k_t / sqrt(DH)) v_t^T  (n likewise) through
// the run and emit  num_t = q_t^T C_t,  qn_t = q_t . n_t  per token -- what S calls of the reference's recurrent step
// leave (src/algos/models/decision_xlstm.py:161-165 -> [ext-xlstm] recurrent_step_stabilized_simple).
//
// The run is cut into chunks of L = 128 tokens. With a_t = sum_{s<=t} log f_s (chunk-local, accumulated in fp64) and C_0
// the memory at the start of the chunk,
//     num_t = e^{a_t} q_t^T C_0 + sum_{j<=t} P~_tj v_j ,     P~_tj = e^{a_t - a_j} i_j (q_t . k_j) / sqrt(DH)
//     C_L   = e^{a_L} C_0 + sum_j K~_j v_j^T ,               K~_j  = e^{a_L - a_j} i_j k_j / sqrt(DH)
// (every exponent <= 0). Everything that is a contraction runs as a BATCHED tcgen05 GEMM over all
// (env, head, chunk) triples of the run at once -- these are real GEMMs (128-row tiles, K = DH or DH + L):
//     G2  dC^T[dv, dk]  = V^T K~            A = V^T [DH x L],  W = K~^T [DH x L]          (chunk update)
//         + scan        : C_{c+1} = e^{a_L} C_c + dC_c. One CTA owns a 128 x 128 tile of C^T of one (env, head) IN
//                         REGISTERS and walks the chunks: MMA of chunk c+1 (second TMEM accumulator) overlaps the
//                         epilogue of chunk c; leaves every chunk-start memory C_c^T as bf16 hi/lo planes (the
//                         inter-chunk operand of G3) and the run's final C in the state; dC never touches HBM
//     G1  S[t, j]       = q_t . k_j         A = Q [L x DH],    W = K [L x DH]
//     P~                : decay/gate matrix applied to S (causal), row sums for q.n
//     G3  num[t, dv]    = [e^{a_t} q_t | P~_t] . [C_c^T | V^T]^T     K = DH + L: inter- and intra-chunk part in ONE
//                         accumulation, written straight into the [B*S, inner] numerator rows
//     qn_t              = e^{a_t} (q_t . n_c) + rowsum_t
// fp32 operands are split into bf16 hi + lo planes and every product is hi*hi + hi*lo + lo*hi in the fp32 TMEM
// accumulator (~2^-17 relative, the trick of the tcgen05 Linear in xl_gemm_tc.cu), so the state left agrees with
// token-by-token stepping to ~1e-6 relative.
// GEMM kernel: one CTA = one 128 x 128 output tile of one batch; warp 0 = TMA producer (3-D tensor maps
// {K, rows, batch}, 128-B swizzle), warp 1 = TMEM allocator + single-thread tcgen05.mma issuer, warps 2..5 = epilogue
// (tcgen05.ld -> registers -> global); 3-stage ring of {A_hi, A_lo, W_hi, W_lo} k-blocks (64 KB per stage).
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

extern int g_prefill_tc_fused;
extern int g_prefill_prep;

namespace ptc {

constexpr int L = 128;          // tokens per chunk
constexpr int BM = 128, BN = 128, BK = 64, UK = 16;
constexpr int kStages = 3;
constexpr int kGemmThreads = 192;
constexpr int kABytes = BM * BK * 2, kWBytes = BN * BK * 2;
constexpr int kStageBytes = 2 * kABytes + 2 * kWBytes;
constexpr int kSmemTotal = kStages * kStageBytes + 1024 + 256;
constexpr int kScanStage = 4 * 32 * 256;                       // update_scan_kernel: per-warp [32 rows][256 B] store staging
constexpr int kSmemScan = kSmemTotal + kScanStage;

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
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
      ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void tcgen05_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// K-major operand tile with 128-B swizzle: rows of 128 B, 8-row atoms of 1024 B (stride byte offset), descriptor
// version 1, layout type 2 (the encoding of CUTLASS cute/arch/mma_sm100_desc.hpp)
__device__ __forceinline__ uint64_t smem_desc(uint32_t addr) {
  uint64_t d = 0;
  d |= (uint64_t)((addr & 0x3FFFF) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)(1024 >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
// kind::f16: D = F32, A = B = BF16, both K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);

__device__ __forceinline__ void umma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
      ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}

// Where batch z = (env*NH + head)*nchunk + chunk puts its [M x N] result: out + env*s_b + head*s_h + chunk*s_c, row
// stride ldo. rows_valid > 0: chunk c only owns rows < rows_valid - c*BM of its tile (ragged last chunk of a run).
struct BatchOut {
  float* out;
  long long s_b, s_h, s_c;
  int ldo, NH, nchunk, rows_valid;
};

// out_z[M, N] = (A_hi + A_lo)_z[M, K] (W_hi + W_lo)_z[N, K]^T  without the lo*lo term.
// PERSISTENT: grid = min(tiles, SMs) CTAs walk the (batch, row tile, column tile) list; the accumulator is double-buffered
// in TMEM (2 x 128 columns), so the epilogue of a tile (tcgen05.ld -> swizzled shared-memory transpose -> coalesced
// stores) overlaps the MMAs of the next one, and the operand ring runs on across tiles.
constexpr int kEpiStage = 4 * 4096;                    // per-warp [32 rows][128 B] store staging
constexpr int kSmemGemm = kStages * kStageBytes + 1024 + 256 + kEpiStage;

__global__ void __launch_bounds__(kGemmThreads, 1)
bgemm_kernel(const __grid_constant__ CUtensorMap map_ah, const __grid_constant__ CUtensorMap map_al,
             const __grid_constant__ CUtensorMap map_wh, const __grid_constant__ CUtensorMap map_wl, BatchOut o, int M,
             int N, int K, int nb) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStageBytes;       // full[S], empty[S], tfull[2], tempty[2], slot
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (kStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8 * (2 * kStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8 * (2 * kStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8 * (2 * kStages + 4);
  const uint32_t epi_base = bar_base + 256;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int num_kb = K / BK;
  const int tn = N / BN, tm = M / BM;
  const int ntiles = nb * tm * tn;
  auto tile_of = [&](int t, int& z, int& m0, int& n0) {         // column tiles of one (batch, row tile) are neighbours
    z = t / (tm * tn);
    const int r = t - z * tm * tn;
    m0 = (r / tn) * BM;
    n0 = (r % tn) * BN;
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_ah) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_al) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_wl) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), 4);                                // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();            // every operand plane is written by the kernels right before this one
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      int it = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        int z, m0, n0;
        tile_of(t, z, m0, n0);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
          const uint32_t st = base + s * kStageBytes;
          mbar_expect_tx(full_bar(s), kStageBytes);
          tma_load_3d(st, &map_ah, full_bar(s), kb * BK, m0, z);
          tma_load_3d(st + kABytes, &map_al, full_bar(s), kb * BK, m0, z);
          tma_load_3d(st + 2 * kABytes, &map_wh, full_bar(s), kb * BK, n0, z);
          tma_load_3d(st + 2 * kABytes + kWBytes, &map_wl, full_bar(s), kb * BK, n0, z);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      int it = 0, j = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
        const int b = j & 1, use = j >> 1;
        if (use > 0) {                                            // the epilogue has drained this accumulator
          mbar_wait(tempty_bar(b), (use - 1) & 1);
          tcgen05_fence_after();
        }
        const uint32_t tacc = tmem_base + (uint32_t)(b * BN);
        for (int kb = 0; kb < num_kb; ++kb, ++it) {
          const int s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1);
          tcgen05_fence_after();
          const uint32_t st = base + s * kStageBytes;
          const uint64_t ah = smem_desc(st), al = smem_desc(st + kABytes);
          const uint64_t wh = smem_desc(st + 2 * kABytes), wl = smem_desc(st + 2 * kABytes + kWBytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t kofs = (uint64_t)((k * UK * 2) >> 4);     // advance inside the 128-B swizzle row
            umma(tacc, al + kofs, wh + kofs, (kb | k) != 0);         // small terms first
            umma(tacc, ah + kofs, wl + kofs, 1u);
            umma(tacc, ah + kofs, wh + kofs, 1u);
          }
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(tfull_bar(b));
      }
    }
  } else {
    // ===== epilogue: warp w may only touch TMEM lanes 32*(w%4) .. +31; accumulator row r lives in lane r =====
    // A lane holds one accumulator row; written as is, a warp store would touch 32 different lines. Each 32-column piece
    // goes through a per-warp [32 rows][128 B] buffer (16-byte chunks XOR-swizzled by row) and leaves as 4 rows x 128
    // contiguous bytes per warp store.
    const int q = warp & 3;
    const uint32_t wst = epi_base + (uint32_t)q * 4096u;
    int j = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++j) {
      const int b = j & 1, use = j >> 1;
      int z, m0, n0;
      tile_of(t, z, m0, n0);
      const int c = z % o.nchunk, bh = z / o.nchunk, be = bh / o.NH, hd = bh - be * o.NH;
      const int mvalid = o.rows_valid > 0 ? min(M, o.rows_valid - c * BM) : M;
      const int rbase = m0 + q * 32;
      float* otile = o.out + (long long)be * o.s_b + (long long)hd * o.s_h + (long long)c * o.s_c + (long long)rbase * o.ldo + n0;
      mbar_wait(tfull_bar(b), use & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < BN; cc += 32) {
        uint32_t r[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + cc);
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
        for (int k = 0; k < 8; ++k)
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wst + (uint32_t)lane * 128u + (uint32_t)((k ^ (lane & 7)) << 4)),
                       "r"(r[4 * k]), "r"(r[4 * k + 1]), "r"(r[4 * k + 2]), "r"(r[4 * k + 3]) : "memory");
        __syncwarp();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const int rr = 4 * i + (lane >> 3), kk = lane & 7;
          uint32_t v0, v1, v2, v3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                       : "r"(wst + (uint32_t)rr * 128u + (uint32_t)((kk ^ (rr & 7)) << 4)) : "memory");
          if (rbase + rr < mvalid)
            *reinterpret_cast<uint4*>(otile + (long long)rr * o.ldo + cc + kk * 4) = make_uint4(v0, v1, v2, v3);
        }
        __syncwarp();
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(b)) : "memory");
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------------------------
// workspace of one run: nb = B * NH * nchunk batches
// ------------------------------------------------------------------------------------------------------------------
struct Ws {
  __nv_bfloat16 *q_hi, *q_lo;       // [nb][L][DH]     Q
  __nv_bfloat16 *k_hi, *k_lo;       // [nb][L][DH]     K
  __nv_bfloat16 *a3_hi, *a3_lo;     // [nb][L][DH + L]  [ e^{a_t} q_t | P~_t ]
  __nv_bfloat16 *w3_hi, *w3_lo;     // [nb][DH][DH + L] [ C_c^T | V^T ]
  __nv_bfloat16 *kt_hi, *kt_lo;     // [nb][DH][L]     K~^T
  float* dC;                        // [nb][DH][DH]    chunk update, transposed (dv, dk)
  float* Sm;                        // [nb][L][L]      q_t . k_j
  double* acum;                     // [nb][L]         a_t
  float* ig;                        // [nb][L]         i_j
  float* FL;                        // [nb]            e^{a_L}
  float* dn;                        // [nb][DH]        chunk update of n
  float* nc;                        // [nb][DH]        n at the start of the chunk
  float* rowsum;                    // [nb][L]         sum_j P~_tj
};

static size_t carve_ws(Ws* w, char* base, int nb, int DH) {
  size_t off = 0;
  auto take = [&](size_t bytes) {
    char* p = base ? base + off : nullptr;
    off = (off + bytes + 255) & ~(size_t)255;
    return p;
  };
  const size_t NB = (size_t)nb, K3 = (size_t)DH + L;
  Ws t;
  t.q_hi = (__nv_bfloat16*)take(NB * L * DH * 2); t.q_lo = (__nv_bfloat16*)take(NB * L * DH * 2);
  t.k_hi = (__nv_bfloat16*)take(NB * L * DH * 2); t.k_lo = (__nv_bfloat16*)take(NB * L * DH * 2);
  t.a3_hi = (__nv_bfloat16*)take(NB * L * K3 * 2); t.a3_lo = (__nv_bfloat16*)take(NB * L * K3 * 2);
  t.w3_hi = (__nv_bfloat16*)take(NB * DH * K3 * 2); t.w3_lo = (__nv_bfloat16*)take(NB * DH * K3 * 2);
  t.kt_hi = (__nv_bfloat16*)take(NB * DH * L * 2); t.kt_lo = (__nv_bfloat16*)take(NB * DH * L * 2);
  t.dC = (float*)take(NB * DH * DH * 4);
  t.Sm = (float*)take(NB * L * L * 4);
  t.acum = (double*)take(NB * L * 8);
  t.ig = (float*)take(NB * L * 4);
  t.FL = (float*)take(NB * 4);
  t.dn = (float*)take(NB * DH * 4);
  t.nc = (float*)take(NB * DH * 4);
  t.rowsum = (float*)take(NB * L * 4);
  if (w) *w = t;
  return off;
}

struct CellParams {
  float* C;                 // [B, NH, DH/128, DH, 128] slab-major state (xl_state_step.cu layout)
  float* n;                 // [B, NH, DH]
  const float *q, *k, *v;   // [B*S, inner]
  const float *fseq, *iseq; // [B*NH, S]
  float* num;               // [B*S, inner]
  float* qn;                // [B*S, NH]
  int B, S, NH, DH, inner, nchunk;
  Ws w;
};

__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

// ------------------------------------------------------------------------------------------------------------------
// Chunk update + scan (G2 + scan above). grid = (DH/128 dk tiles, DH/128 dv tiles, B*NH), 320 threads, 1 CTA / SM.
//   warp 0: TMA producer over the flattened (chunk, k-block) sequence of the run (3-stage ring, runs up to 1.5 chunks ahead)
//   warp 1: tcgen05.mma issuer; chunk c accumulates into TMEM buffer c & 1 (2 x 128 columns)
//   warps 2..9: thread = one dv row of the tile with 64 of its 128 dk values of C^T in registers (two warps share a TMEM
//               lane quarter and split the columns: the per-chunk epilogue is what bounds this kernel). Per chunk:
//               chunk-start values -> bf16 hi/lo planes into W3[:, 0:DH] (coalesced through shared memory); then tcgen05.ld of dC and
//               C <- e^{a_L} C + dC. First / last: the state tile (slab-major [dk][dv]: a warp reads 32 consecutive dv).
// ------------------------------------------------------------------------------------------------------------------
// NSPLIT = epilogue warps per TMEM lane quarter (column splits of the tile): 2 (8 epilogue warps, 64 columns per thread) or
// 4 (16 warps, 32 columns per thread). Same arithmetic per element either way (bit-identical).
template <int NSPLIT>
__global__ void __launch_bounds__(64 + 4 * NSPLIT * 32, 1)
update_scan_kernel(const __grid_constant__ CUtensorMap map_vh, const __grid_constant__ CUtensorMap map_vl,
                   const __grid_constant__ CUtensorMap map_kh, const __grid_constant__ CUtensorMap map_kl, CellParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar_base = base + kStages * kStageBytes;   // full[kStages], empty[kStages], tfull[2], tempty[2], slot
  auto full_bar = [&](int s) { return bar_base + 8 * s; };
  auto empty_bar = [&](int s) { return bar_base + 8 * (kStages + s); };
  auto tfull_bar = [&](int b) { return bar_base + 8 * (2 * kStages + b); };
  auto tempty_bar = [&](int b) { return bar_base + 8 * (2 * kStages + 2 + b); };
  const uint32_t tmem_slot = bar_base + 8 * (2 * kStages + 4);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, bh = blockIdx.z;
  const int DH = p.DH, K3 = DH + L, nchunk = p.nchunk;
  constexpr int kKb = L / BK;                                 // k-blocks per chunk
  const int niter = nchunk * kKb;
  constexpr int kScanThreads = 64 + 4 * NSPLIT * 32;          // TMA warp + MMA warp + 4 * NSPLIT epilogue warps

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_vl) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kh) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&map_kl) : "memory");
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(tfull_bar(b), 1);
      mbar_init(tempty_bar(b), kScanThreads / 32 - 2);        // one arrival per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(tmem_slot), "n"(2 * BN) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot) : "memory");
  pdl_wait();
  pdl_trigger();

  if (warp == 0) {
    if (lane == 0) {
      for (int it = 0; it < niter; ++it) {
        const int s = it % kStages;
        const int c = it / kKb, kb = it - c * kKb;
        const int z = bh * nchunk + c;
        if (it >= kStages) mbar_wait(empty_bar(s), ((it / kStages) & 1) ^ 1);
        const uint32_t st = base + s * kStageBytes;
        mbar_expect_tx(full_bar(s), kStageBytes);
        tma_load_3d(st, &map_vh, full_bar(s), kb * BK, m0, z);
        tma_load_3d(st + kABytes, &map_vl, full_bar(s), kb * BK, m0, z);
        tma_load_3d(st + 2 * kABytes, &map_kh, full_bar(s), kb * BK, n0, z);
        tma_load_3d(st + 2 * kABytes + kWBytes, &map_kl, full_bar(s), kb * BK, n0, z);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int c = 0; c < nchunk; ++c) {
        const int b = c & 1, use = c >> 1;
        if (use > 0) {                                         // the epilogue has drained this accumulator
          mbar_wait(tempty_bar(b), (use - 1) & 1);
          tcgen05_fence_after();
        }
        const uint32_t tacc = tmem_base + (uint32_t)(b * BN);
        for (int kb = 0; kb < kKb; ++kb) {
          const int it = c * kKb + kb, s = it % kStages;
          mbar_wait(full_bar(s), (it / kStages) & 1);
          tcgen05_fence_after();
          const uint32_t st = base + s * kStageBytes;
          const uint64_t ah = smem_desc(st), al = smem_desc(st + kABytes);
          const uint64_t wh = smem_desc(st + 2 * kABytes), wl = smem_desc(st + 2 * kABytes + kWBytes);
#pragma unroll
          for (int k = 0; k < BK / UK; ++k) {
            const uint64_t kofs = (uint64_t)((k * UK * 2) >> 4);
            umma(tacc, al + kofs, wh + kofs, (kb | k) != 0);
            umma(tacc, ah + kofs, wl + kofs, 1u);
            umma(tacc, ah + kofs, wh + kofs, 1u);
          }
          tcgen05_commit(empty_bar(s));
        }
        tcgen05_commit(tfull_bar(b));
      }
    }
  } else {
    // 4 * NSPLIT epilogue warps: warp w reads TMEM lanes 32*(w%4).. (the hardware rule) and owns column part (w-2)/4 of
    // the tile, i.e. a thread carries HC = 128 / NSPLIT dk values of one dv row
    constexpr int HC = BN / NSPLIT;
    constexpr int KC = HC / 8;                                 // 16-byte chunks (8 bf16) per row and plane
    constexpr int kPitch = HC * 2;                             // staging row pitch in bytes
    const int q = warp & 3, half = (warp - 2) >> 2;
    const int r = q * 32 + lane;                               // dv row of the tile == TMEM lane
    const int nc0 = n0 + half * HC;                            // first dk column of this thread
    const uint32_t wst = base + kStages * kStageBytes + 256 + (uint32_t)(warp - 2) * (32u * kPitch);
    // XOR key of a staging row: 8 consecutive lanes (a quarter warp of 128-bit accesses) must hit 8 different bank groups
    auto skey = [](int row) { return KC == 8 ? (row & 7) : ((row >> 1) & 3); };
    // state tile: C[dk = nc0 + j][dv] at slab (dv / 128) = blockIdx.y, column r
    float* cs = p.C + (((int64_t)bh * (DH >> 7) + blockIdx.y) * DH + nc0) * 128 + r;
    float cv[HC];
#pragma unroll
    for (int j = 0; j < HC; ++j) cv[j] = cs[(int64_t)j * 128];
    for (int c = 0; c < nchunk; ++c) {
      const int b = c & 1, use = c >> 1;
      const int64_t z = (int64_t)bh * nchunk + c;
      const float F = p.w.FL[z];
      // chunk-start tile -> W3 planes. A lane holds one row (HC * 2 B per plane): staged through a per-warp
      // [32 rows][HC * 2 B] buffer (16-byte chunks XOR-swizzled by row), so a warp store writes 32 / KC rows of HC * 2
      // contiguous bytes instead of 32 different lines
#pragma unroll
      for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
        for (int k = 0; k < KC; ++k) {
          uint32_t wv[4];
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            uint32_t hh, ll;
            split2(cv[8 * k + 2 * e], cv[8 * k + 2 * e + 1], hh, ll);
            wv[e] = pl ? ll : hh;
          }
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(wst + (uint32_t)lane * kPitch + (uint32_t)((k ^ skey(lane)) << 4)),
                       "r"(wv[0]), "r"(wv[1]), "r"(wv[2]), "r"(wv[3]) : "memory");
        }
        __syncwarp();
        __nv_bfloat16* gt = (pl ? p.w.w3_lo : p.w.w3_hi) + (z * DH + m0 + q * 32) * K3 + nc0;
#pragma unroll
        for (int i = 0; i < KC; ++i) {
          const int rr = (32 / KC) * i + lane / KC, kk = lane % KC;
          uint32_t v0, v1, v2, v3;
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(v0), "=r"(v1), "=r"(v2), "=r"(v3)
                       : "r"(wst + (uint32_t)rr * kPitch + (uint32_t)((kk ^ skey(rr)) << 4)) : "memory");
          *reinterpret_cast<uint4*>(gt + (int64_t)rr * K3 + kk * 8) = make_uint4(v0, v1, v2, v3);
        }
        __syncwarp();
      }
      mbar_wait(tfull_bar(b), use & 1);
      tcgen05_fence_after();
#pragma unroll
      for (int cc = 0; cc < HC; cc += 32) {
        uint32_t t[32];
        const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(b * BN + half * HC + cc);
        asm volatile(
            "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
            "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
            "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
            : "=r"(t[0]), "=r"(t[1]), "=r"(t[2]), "=r"(t[3]), "=r"(t[4]), "=r"(t[5]), "=r"(t[6]), "=r"(t[7]),
              "=r"(t[8]), "=r"(t[9]), "=r"(t[10]), "=r"(t[11]), "=r"(t[12]), "=r"(t[13]), "=r"(t[14]), "=r"(t[15]),
              "=r"(t[16]), "=r"(t[17]), "=r"(t[18]), "=r"(t[19]), "=r"(t[20]), "=r"(t[21]), "=r"(t[22]),
              "=r"(t[23]), "=r"(t[24]), "=r"(t[25]), "=r"(t[26]), "=r"(t[27]), "=r"(t[28]), "=r"(t[29]),
              "=r"(t[30]), "=r"(t[31])
            : "r"(taddr)
            : "memory");
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
        for (int j = 0; j < 32; ++j) cv[cc + j] = fmaf(F, cv[cc + j], __uint_as_float(t[j]));
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(tempty_bar(b)) : "memory");
    }
#pragma unroll
    for (int j = 0; j < HC; ++j) cs[(int64_t)j * 128] = cv[j];
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(2 * BN) : "memory");
  }
}

// n over the chunks of a run: n_c (chunk start, fp32) for q.n, final n into the state. grid = (B*NH, DH/128), 128 threads
__global__ void __launch_bounds__(128) nscan_kernel(CellParams p) {
  const int DH = p.DH, nchunk = p.nchunk, bh = blockIdx.x;
  const int r = blockIdx.y * 128 + threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (r >= DH) return;
  float nv = p.n[(int64_t)bh * DH + r];
  for (int c0 = 0; c0 < nchunk; c0 += 8) {
    float F[8], d[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {                         // the loads of 8 chunks are in flight together
      const int64_t z = (int64_t)bh * nchunk + min(c0 + j, nchunk - 1);
      F[j] = p.w.FL[z];
      d[j] = p.w.dn[z * DH + r];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (c0 + j < nchunk) {
        p.w.nc[((int64_t)bh * nchunk + c0 + j) * DH + r] = nv;
        nv = fmaf(F[j], nv, d[j]);
      }
    }
  }
  p.n[(int64_t)bh * DH + r] = nv;
}

constexpr int kPrepParts = 5;     // channel ranges per (chunk, env*head): more CTAs for the plane writes

// ------------------------------------------------------------------------------------------------------------------
// Chunk preparation. grid = (nchunk, B*NH, kPrepParts), 256 threads. Decay tables of the chunk (every CTA), then this
// CTA's channel range of: Q, K and e^{a_t} q_t planes (token-major rows), K~^T and V^T planes (channel-major rows,
// 16 tokens = one 32-byte piece per thread and step), and dn = sum_j K~_j.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) prep_kernel(CellParams p) {
  __shared__ float s_lf[L], s_i[L], s_F[L], s_ks[L];
  __shared__ double s_a[L];
  const int c = blockIdx.x, bh = blockIdx.y, part = blockIdx.z;
  const int NH = p.NH, DH = p.DH, S = p.S, inner = p.inner, K3 = DH + L;
  const int b = bh / NH, hd = bh - b * NH;
  const int tid = threadIdx.x;
  const int64_t z = (int64_t)bh * p.nchunk + c;
  const int t0 = c * L;
  const int nvalid = min(L, S - t0);
  const int64_t row0 = (int64_t)b * S + t0;
  const int hoff = hd * DH;
  const float kscale = rsqrtf((float)DH);
  pdl_wait();
  pdl_trigger();

  if (tid < L) {
    float lf = 0.f, ii = 0.f;
    if (tid < nvalid) {
      const float f = p.fseq[(int64_t)bh * S + t0 + tid];
      lf = f > 0.f ? fmaxf(logf(f), -200.f) : -200.f;
      ii = p.iseq[(int64_t)bh * S + t0 + tid];
    }
    s_lf[tid] = lf;
    s_i[tid] = ii;
  }
  __syncthreads();
  if (tid < 32) {
    // inclusive prefix sum of 128 doubles: 4 per lane, then a warp scan of the lane totals
    double v[4];
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      run += (double)s_lf[tid * 4 + j];
      v[j] = run;
    }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += up;
    }
    const double excl = incl - run;
#pragma unroll
    for (int j = 0; j < 4; ++j) s_a[tid * 4 + j] = excl + v[j];
  }
  __syncthreads();
  if (tid < L) {
    const double aL = s_a[L - 1];
    s_F[tid] = expf((float)s_a[tid]);
    s_ks[tid] = expf((float)(aL - s_a[tid])) * s_i[tid] * kscale;
    if (part == 0) {
      p.w.acum[z * L + tid] = s_a[tid];
      p.w.ig[z * L + tid] = s_i[tid];
      if (tid == 0) p.w.FL[z] = expf((float)aL);
    }
  }
  __syncthreads();

  // channel range of this CTA (multiples of 4)
  const int cw = ((DH / 4 + kPrepParts - 1) / kPrepParts) * 4;
  const int ch0 = part * cw, ch1 = min(DH, ch0 + cw);
  if (ch0 >= ch1) return;
  const int w4 = (ch1 - ch0) >> 2;

  // token-major planes: Q, K [L][DH]; e^{a_t} q_t into A3[:, 0:DH]. The loads of 4 items are issued together (the plane
  // stores in between would otherwise order them one item at a time: the kernel is bound by load latency)
  const int nit = L * w4;
  for (int i0 = tid; i0 < nit; i0 += 4 * 256) {
    float4 xq[4], xk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = i0 + u * 256;
      const int t = idx / w4, c4 = idx - t * w4;
      xq[u] = xk[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (idx < nit && t < nvalid) {
        xq[u] = *reinterpret_cast<const float4*>(p.q + (row0 + t) * inner + hoff + ch0 + 4 * c4);
        xk[u] = *reinterpret_cast<const float4*>(p.k + (row0 + t) * inner + hoff + ch0 + 4 * c4);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = i0 + u * 256;
      if (idx >= nit) break;
      const int t = idx / w4, c4 = idx - t * w4;
      const int ch = ch0 + 4 * c4;
      uint32_t h0, l0, h1, l1;
      const int64_t o = (z * L + t) * DH + ch;
      split2(xq[u].x, xq[u].y, h0, l0);
      split2(xq[u].z, xq[u].w, h1, l1);
      *reinterpret_cast<uint2*>(p.w.q_hi + o) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p.w.q_lo + o) = make_uint2(l0, l1);
      split2(xk[u].x, xk[u].y, h0, l0);
      split2(xk[u].z, xk[u].w, h1, l1);
      *reinterpret_cast<uint2*>(p.w.k_hi + o) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p.w.k_lo + o) = make_uint2(l0, l1);
      const float F = s_F[t];
      const int64_t o3 = (z * L + t) * K3 + ch;
      split2(F * xq[u].x, F * xq[u].y, h0, l0);
      split2(F * xq[u].z, F * xq[u].w, h1, l1);
      *reinterpret_cast<uint2*>(p.w.a3_hi + o3) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p.w.a3_lo + o3) = make_uint2(l0, l1);
    }
  }
  // channel-major planes: K~^T [DH][L], V^T into W3[:, DH:DH+L]; item = (token group of 16, channel). k and v of the item
  // are loaded together, before any store
  const int nchs = ch1 - ch0;
  for (int idx = tid; idx < (L / 16) * nchs; idx += 256) {
    const int g = idx / nchs, r = ch0 + (idx - g * nchs);
    float kv[16], vv[16];
#pragma unroll
    for (int jj = 0; jj < 16; ++jj) {
      const int j = g * 16 + jj;
      kv[jj] = j < nvalid ? p.k[(row0 + j) * inner + hoff + r] : 0.f;
      vv[jj] = j < nvalid ? p.v[(row0 + j) * inner + hoff + r] : 0.f;
    }
    uint32_t h[8], l[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j0 = g * 16 + 2 * jj;
      split2(kv[2 * jj] * s_ks[j0], kv[2 * jj + 1] * s_ks[j0 + 1], h[jj], l[jj]);
    }
    const int64_t ok = (z * DH + r) * L + g * 16;
    *reinterpret_cast<uint4*>(p.w.kt_hi + ok) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(p.w.kt_hi + ok + 8) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(p.w.kt_lo + ok) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(p.w.kt_lo + ok + 8) = make_uint4(l[4], l[5], l[6], l[7]);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) split2(vv[2 * jj], vv[2 * jj + 1], h[jj], l[jj]);
    const int64_t ov = (z * DH + r) * K3 + DH + g * 16;
    *reinterpret_cast<uint4*>(p.w.w3_hi + ov) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(p.w.w3_hi + ov + 8) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(p.w.w3_lo + ov) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(p.w.w3_lo + ov + 8) = make_uint4(l[4], l[5], l[6], l[7]);
  }
  // dn[r] = sum_j K~_j[r]: two threads per channel (token halves), 8 loads in flight each, halves added in a fixed order
  {
    __shared__ float s_part[256];
    const int rr = tid & 127, half = tid >> 7;
    for (int rb = 0; rb < nchs; rb += 128) {             // (CTA-uniform trip count)
      const int r = ch0 + rb + rr;
      float acc = 0.f;
      if (r < ch1) {
        const int j0 = half * (L / 2), j1 = min(nvalid, j0 + L / 2);
        for (int jb = j0; jb < j1; jb += 8) {
          float kk[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) kk[u] = jb + u < j1 ? p.k[(row0 + jb + u) * inner + hoff + r] : 0.f;
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (jb + u < j1) acc = fmaf(kk[u], s_ks[jb + u], acc);
        }
      }
      __syncthreads();
      s_part[tid] = acc;
      __syncthreads();
      if (half == 0 && r < ch1) p.w.dn[z * DH + r] = s_part[rr] + s_part[rr + 128];
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Chunk preparation, single-read version (default; xl_set_option("prefill_prep", 0) selects prep_kernel above). Same
// outputs, bit for bit. prep_kernel reads k three times (token-major planes, channel-major planes, dn) and was measured at
// 1.55 GB of DRAM traffic per launch for 1.30 GB of operands (ncu, 206M, 16k tokens), load-latency bound in short phases.
// Here a CTA owns a [128 tokens x 64 channels] tile: q and k are read once with coalesced 128-bit loads, the token-major
// planes leave straight from registers, the raw k (then v) tile is parked in shared memory ([token][64 + 4] floats: 128-bit
// row writes and stride-1 column reads are both conflict-free) and the channel-major planes and dn are formed from there.
// grid = (nchunk, B*NH, DH/64), 256 threads, ~36 KB of shared memory.
// ------------------------------------------------------------------------------------------------------------------
constexpr int kPrepCW = 64;                   // channels per CTA
constexpr int kPrepLd = kPrepCW + 4;          // tile row stride in floats

__global__ void __launch_bounds__(256) prep2_kernel(CellParams p) {
  __shared__ float s_lf[L], s_i[L], s_F[L], s_ks[L];
  __shared__ double s_a[L];
  __shared__ __align__(16) float s_tile[L * kPrepLd];
  __shared__ float s_part[2 * kPrepCW];
  const int c = blockIdx.x, bh = blockIdx.y, part = blockIdx.z;
  const int NH = p.NH, DH = p.DH, S = p.S, inner = p.inner, K3 = DH + L;
  const int b = bh / NH, hd = bh - b * NH;
  const int tid = threadIdx.x;
  const int64_t z = (int64_t)bh * p.nchunk + c;
  const int t0 = c * L;
  const int nvalid = min(L, S - t0);
  const int64_t row0 = (int64_t)b * S + t0;
  const int ch0 = part * kPrepCW;
  const int hoff = hd * DH + ch0;
  const float kscale = rsqrtf((float)DH);
  pdl_wait();
  pdl_trigger();

  // decay tables of the chunk (as prep_kernel)
  if (tid < L) {
    float lf = 0.f, ii = 0.f;
    if (tid < nvalid) {
      const float f = p.fseq[(int64_t)bh * S + t0 + tid];
      lf = f > 0.f ? fmaxf(logf(f), -200.f) : -200.f;
      ii = p.iseq[(int64_t)bh * S + t0 + tid];
    }
    s_lf[tid] = lf;
    s_i[tid] = ii;
  }
  __syncthreads();
  if (tid < 32) {
    double v[4];
    double run = 0.0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      run += (double)s_lf[tid * 4 + j];
      v[j] = run;
    }
    double incl = run;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const double up = __shfl_up_sync(0xffffffffu, incl, o);
      if (tid >= o) incl += up;
    }
    const double excl = incl - run;
#pragma unroll
    for (int j = 0; j < 4; ++j) s_a[tid * 4 + j] = excl + v[j];
  }
  __syncthreads();
  if (tid < L) {
    const double aL = s_a[L - 1];
    s_F[tid] = expf((float)s_a[tid]);
    s_ks[tid] = expf((float)(aL - s_a[tid])) * s_i[tid] * kscale;
    if (part == 0) {
      p.w.acum[z * L + tid] = s_a[tid];
      p.w.ig[z * L + tid] = s_i[tid];
      if (tid == 0) p.w.FL[z] = expf((float)aL);
    }
  }
  __syncthreads();

  // token-major pass over q and k: thread = (float4 column c4, row tr of a 16-row band); 4 rows per thread and step
  const int c4 = tid & 15, tr = tid >> 4;
  const float* qg = p.q + row0 * inner + hoff + 4 * c4;
  const float* kg = p.k + row0 * inner + hoff + 4 * c4;
  const float* vg = p.v + row0 * inner + hoff + 4 * c4;
#pragma unroll 1
  for (int tb = 0; tb < L; tb += 64) {
    float4 xq[4], xk[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = tb + 16 * u + tr;
      xq[u] = xk[u] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (t < nvalid) {
        xq[u] = *reinterpret_cast<const float4*>(qg + (int64_t)t * inner);
        xk[u] = *reinterpret_cast<const float4*>(kg + (int64_t)t * inner);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int t = tb + 16 * u + tr;
      const int ch = ch0 + 4 * c4;
      uint32_t h0, l0, h1, l1;
      const int64_t o = (z * L + t) * DH + ch;
      split2(xq[u].x, xq[u].y, h0, l0);
      split2(xq[u].z, xq[u].w, h1, l1);
      *reinterpret_cast<uint2*>(p.w.q_hi + o) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p.w.q_lo + o) = make_uint2(l0, l1);
      split2(xk[u].x, xk[u].y, h0, l0);
      split2(xk[u].z, xk[u].w, h1, l1);
      *reinterpret_cast<uint2*>(p.w.k_hi + o) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p.w.k_lo + o) = make_uint2(l0, l1);
      const float F = s_F[t];
      const int64_t o3 = (z * L + t) * K3 + ch;
      split2(F * xq[u].x, F * xq[u].y, h0, l0);
      split2(F * xq[u].z, F * xq[u].w, h1, l1);
      *reinterpret_cast<uint2*>(p.w.a3_hi + o3) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(p.w.a3_lo + o3) = make_uint2(l0, l1);
      *reinterpret_cast<float4*>(s_tile + t * kPrepLd + 4 * c4) = xk[u];
    }
  }
  // v of the first half of the tile is requested now: in flight under the channel-major K~^T work
  float4 xv[4];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int t = 16 * u + tr;
    xv[u] = t < nvalid ? *reinterpret_cast<const float4*>(vg + (int64_t)t * inner) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  __syncthreads();
  // channel-major K~^T [DH][L]: item = (token group of 16, channel); a warp = 32 consecutive channels of one group
#pragma unroll 1
  for (int idx = tid; idx < (L / 16) * kPrepCW; idx += 256) {
    const int g = idx / kPrepCW, rl = idx - g * kPrepCW;
    uint32_t h[8], l[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j0 = g * 16 + 2 * jj;
      split2(s_tile[j0 * kPrepLd + rl] * s_ks[j0], s_tile[(j0 + 1) * kPrepLd + rl] * s_ks[j0 + 1], h[jj], l[jj]);
    }
    const int64_t ok = (z * DH + ch0 + rl) * L + g * 16;
    *reinterpret_cast<uint4*>(p.w.kt_hi + ok) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(p.w.kt_hi + ok + 8) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(p.w.kt_lo + ok) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(p.w.kt_lo + ok + 8) = make_uint4(l[4], l[5], l[6], l[7]);
  }
  // dn[r] = sum_j K~_j[r]: two threads per channel (token halves, each in token order), halves added in a fixed order
  if (tid < 2 * kPrepCW) {
    const int rl = tid & (kPrepCW - 1), half = tid / kPrepCW;
    const int j0 = half * (L / 2), j1 = min(nvalid, j0 + L / 2);
    float acc = 0.f;
    for (int j = j0; j < j1; ++j) acc = fmaf(s_tile[j * kPrepLd + rl], s_ks[j], acc);
    s_part[tid] = acc;
  }
  __syncthreads();                                         // s_part written, tile reads done
  if (tid < kPrepCW) p.w.dn[z * DH + ch0 + tid] = s_part[tid] + s_part[tid + kPrepCW];
  // v tile
#pragma unroll
  for (int u = 0; u < 4; ++u) *reinterpret_cast<float4*>(s_tile + (16 * u + tr) * kPrepLd + 4 * c4) = xv[u];
#pragma unroll
  for (int u = 0; u < 4; ++u) {
    const int t = 64 + 16 * u + tr;
    xv[u] = t < nvalid ? *reinterpret_cast<const float4*>(vg + (int64_t)t * inner) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int u = 0; u < 4; ++u) *reinterpret_cast<float4*>(s_tile + (64 + 16 * u + tr) * kPrepLd + 4 * c4) = xv[u];
  __syncthreads();
  // channel-major V^T into W3[:, DH:DH+L]
#pragma unroll 1
  for (int idx = tid; idx < (L / 16) * kPrepCW; idx += 256) {
    const int g = idx / kPrepCW, rl = idx - g * kPrepCW;
    uint32_t h[8], l[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const int j0 = g * 16 + 2 * jj;
      split2(s_tile[j0 * kPrepLd + rl], s_tile[(j0 + 1) * kPrepLd + rl], h[jj], l[jj]);
    }
    const int64_t ov = (z * DH + ch0 + rl) * K3 + DH + g * 16;
    *reinterpret_cast<uint4*>(p.w.w3_hi + ov) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(p.w.w3_hi + ov + 8) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(p.w.w3_lo + ov) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(p.w.w3_lo + ov + 8) = make_uint4(l[4], l[5], l[6], l[7]);
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Chunk scan of one run. grid = (DH*DH/1024 + 1, B*NH), 256 threads. A thread owns 4 consecutive dk of one dv of C^T and
// walks the chunks: chunk-start memory -> bf16 hi/lo planes W3[:, 0:DH] (the inter-chunk operand of G3), then
// C <- e^{a_L} C + dC_c. The last CTA row does the same for n (fp32 copies of every chunk-start n).
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) scan_kernel(CellParams p) {
  const int DH = p.DH, K3 = DH + L, nchunk = p.nchunk;
  const int bh = blockIdx.y;
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  if (blockIdx.x == gridDim.x - 1) {
    for (int r = tid; r < DH; r += 256) {
      float nv = p.n[(int64_t)bh * DH + r];
      for (int c = 0; c < nchunk; ++c) {
        const int64_t z = (int64_t)bh * nchunk + c;
        p.w.nc[z * DH + r] = nv;
        nv = fmaf(p.w.FL[z], nv, p.w.dn[z * DH + r]);
      }
      p.n[(int64_t)bh * DH + r] = nv;
    }
    return;
  }
  const int e = (blockIdx.x * 256 + tid) * 4;          // element of C^T: dv = e / DH, dk = e % DH .. +3
  if (e >= DH * DH) return;
  const int dv = e / DH, dk = e - dv * DH;
  // state element C[dk][dv] (slab-major: [DH/128 slabs][DH rows dk][128 cols dv])
  float* cs = p.C + (((int64_t)bh * (DH >> 7) + (dv >> 7)) * DH + dk) * 128 + (dv & 127);
  float cv[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cv[j] = cs[j * 128];
#pragma unroll 4
  for (int c = 0; c < nchunk; ++c) {
    const int64_t z = (int64_t)bh * nchunk + c;
    const float4 d = *reinterpret_cast<const float4*>(p.w.dC + (z * DH + dv) * DH + dk);
    const float F = p.w.FL[z];
    uint32_t h0, l0, h1, l1;
    split2(cv[0], cv[1], h0, l0);
    split2(cv[2], cv[3], h1, l1);
    const int64_t o = (z * DH + dv) * K3 + dk;
    *reinterpret_cast<uint2*>(p.w.w3_hi + o) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(p.w.w3_lo + o) = make_uint2(l0, l1);
    cv[0] = fmaf(F, cv[0], d.x);
    cv[1] = fmaf(F, cv[1], d.y);
    cv[2] = fmaf(F, cv[2], d.z);
    cv[3] = fmaf(F, cv[3], d.w);
  }
#pragma unroll
  for (int j = 0; j < 4; ++j) cs[j * 128] = cv[j];
}

// ------------------------------------------------------------------------------------------------------------------
// P~ = decay/gate matrix applied to S (causal) -> A3[:, DH:DH+L] planes, and its row sums. grid = batches, 256 threads:
// warp w owns rows w, w+8, ...; a lane owns 4 consecutive j.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) pmat_kernel(CellParams p) {
  __shared__ double s_a[L];
  __shared__ float s_g[L];
  const int64_t z = blockIdx.x;
  const int DH = p.DH, K3 = DH + L;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const float kscale = rsqrtf((float)DH);
  pdl_wait();
  pdl_trigger();
  if (tid < L) {
    s_a[tid] = p.w.acum[z * L + tid];
    s_g[tid] = p.w.ig[z * L + tid] * kscale;
  }
  __syncthreads();
  for (int t = warp; t < L; t += 8) {
    const float4 sv = *reinterpret_cast<const float4*>(p.w.Sm + (z * L + t) * L + 4 * lane);
    const float sx[4] = {sv.x, sv.y, sv.z, sv.w};
    const double at = s_a[t];
    float pv[4];
    float rs = 0.f;
#pragma unroll
    for (int jj = 0; jj < 4; ++jj) {
      const int j = 4 * lane + jj;
      pv[jj] = j <= t ? expf((float)(at - s_a[j])) * s_g[j] * sx[jj] : 0.f;
      rs += pv[jj];
    }
    rs = warp_sum(rs);
    uint32_t h0, l0, h1, l1;
    split2(pv[0], pv[1], h0, l0);
    split2(pv[2], pv[3], h1, l1);
    const int64_t o = (z * L + t) * K3 + DH + 4 * lane;
    *reinterpret_cast<uint2*>(p.w.a3_hi + o) = make_uint2(h0, h1);
    *reinterpret_cast<uint2*>(p.w.a3_lo + o) = make_uint2(l0, l1);
    if (lane == 0) p.w.rowsum[z * L + t] = rs;
  }
}

// qn_t = e^{a_t} (q_t . n_c) + rowsum_t.   grid = (batches, L/8), 256 threads: one warp per token
__global__ void __launch_bounds__(256) qn_kernel(CellParams p) {
  const int64_t z = blockIdx.x;
  const int DH = p.DH, NH = p.NH, S = p.S;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c = (int)(z % p.nchunk), bh = (int)(z / p.nchunk), b = bh / NH, hd = bh - b * NH;
  const int t = blockIdx.y * 8 + warp, tok = c * L + t;
  pdl_wait();
  pdl_trigger();
  if (tok >= S) return;
  const float* qp = p.q + ((int64_t)b * S + tok) * p.inner + hd * DH;
  const float* np = p.w.nc + z * DH;
  float a = 0.f;
  for (int e = lane; e < DH; e += 32) a = fmaf(qp[e], np[e], a);
  a = warp_sum(a);
  if (lane == 0)
    p.qn[((int64_t)b * S + tok) * NH + hd] = fmaf(expf((float)p.w.acum[z * L + t]), a, p.w.rowsum[z * L + t]);
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode() {
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

// bf16 [batches][rows][K] with row stride `ld` and batch stride `bs` (elements): box = {64, 128, 1}, 128-B swizzle
static bool make_map3(CUtensorMap* m, const void* ptr, int K, int rows, int nb, int64_t ld, int64_t bs) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return false;
  cuuint64_t gdim[3] = {(cuuint64_t)K, (cuuint64_t)rows, (cuuint64_t)nb};
  cuuint64_t gstride[2] = {(cuuint64_t)ld * 2, (cuuint64_t)bs * 2};
  cuuint32_t box[3] = {(cuuint32_t)BK, (cuuint32_t)BM, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  return enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(ptr), gdim, gstride, box, estr,
             CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
             CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

struct Operand {
  const __nv_bfloat16 *hi, *lo;
  int64_t ld, bs;
};

static int device_sms() {                          // SMs of the current device (every device of a box is the same part)
  static const int sms = [] {
    int dev = 0, n = 148;
    if (cudaGetDevice(&dev) == cudaSuccess) cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    return n > 0 ? n : 148;
  }();
  return sms;
}

// max_ctas > 0 caps the persistent grid (the GEMM then shares the GPU with a kernel of another stream)
static cudaError_t launch_bgemm(const Operand& a, const Operand& w, const BatchOut& o, int M, int N, int K, int nb,
                                cudaStream_t s, int max_ctas = 0) {
  if (M % BM || N % BN || K % BK || K < BK) return cudaErrorInvalidValue;
  CUtensorMap mah, mal, mwh, mwl;
  if (!make_map3(&mah, a.hi, K, M, nb, a.ld, a.bs) || !make_map3(&mal, a.lo, K, M, nb, a.ld, a.bs) ||
      !make_map3(&mwh, w.hi, K, N, nb, w.ld, w.bs) || !make_map3(&mwl, w.lo, K, N, nb, w.ld, w.bs))
    return cudaErrorUnknown;
  if (cudaError_t e = ensure_dyn_smem<&bgemm_kernel>(kSmemGemm); e != cudaSuccess) return e;
  const int tiles = nb * (M / BM) * (N / BN);
  int grid = tiles < device_sms() ? tiles : device_sms();
  if (max_ctas > 0 && grid > max_ctas) grid = max_ctas;
  return launch_k(bgemm_kernel, dim3(grid), dim3(kGemmThreads), kSmemGemm, s, mah, mal, mwh, mwl, o, M, N, K, nb);
}

}  // namespace ptc

int g_prefill_tc_fused = 1;   // xl_set_option("prefill_tc_fused"): 0 = chunk updates through HBM + element-parallel scan (A/B)
int g_prefill_prep = 1;       // xl_set_option("prefill_prep"): 0 = prep_kernel (k read three times), 1 = prep2_kernel

bool prefill_cell_tc_supported(int DH) { return DH % 128 == 0 && DH >= 128 && DH <= 1024; }
int prefill_cell_tc_chunk() { return ptc::L; }

size_t prefill_cell_tc_ws_bytes(int B, int S, int NH, int DH) {
  const int nchunk = (S + ptc::L - 1) / ptc::L;
  return ptc::carve_ws(nullptr, nullptr, B * NH * nchunk, DH);
}

int g_prefill_scan_split = 2;    // xl_set_option("prefill_scan_split"): epilogue warps per TMEM lane quarter of the chunk scan (2 or 4)
int g_prefill_tc_overlap = 1;   // xl_set_option("prefill_tc_overlap"): S = QK^T, P~ and the n scan beside the chunk update + scan

cudaError_t launch_cell_tc(float* C, float* n, const float* q, const float* k, const float* v, const float* fseq,
                           const float* iseq, float* num, float* qn, void* ws, int B, int S, int NH, int DH, int inner,
                           cudaStream_t s, const CellSideStream* side) {
  using namespace ptc;
  if (!prefill_cell_tc_supported(DH) || S <= 0 || inner != NH * DH) return cudaErrorInvalidValue;
  CellParams p;
  p.C = C; p.n = n; p.q = q; p.k = k; p.v = v; p.fseq = fseq; p.iseq = iseq; p.num = num; p.qn = qn;
  p.B = B; p.S = S; p.NH = NH; p.DH = DH; p.inner = inner;
  p.nchunk = (S + L - 1) / L;
  const int BH = B * NH, nb = BH * p.nchunk, K3 = DH + L;
  carve_ws(&p.w, (char*)ws, nb, DH);
  cudaError_t e;
  // chunk operands
  if (g_prefill_prep)
    e = launch_k(prep2_kernel, dim3(p.nchunk, BH, DH / kPrepCW), dim3(256), 0, s, p);
  else
    e = launch_k(prep_kernel, dim3(p.nchunk, BH, kPrepParts), dim3(256), 0, s, p);
  if (e != cudaSuccess) return e;
  // The chunk update + scan runs (DH/128)^2 * B*NH CTAs, one per SM, each bound by what one SM ingests through TMA: with
  // few envs it leaves SMs idle (206M x 1 env: 100 CTAs on 148 SMs for 285 us). S = QK^T, P~ and the n scan depend on the
  // chunk operands only, so in that case they run on a side stream BESIDE it -- the S GEMM as a persistent grid capped at
  // the SMs the scan leaves free (whatever the block scheduler does first, the scan's CTAs find their SMs) -- and the
  // numerator GEMM joins both.
  const int scan_ctas = (DH / BN) * (DH / BM) * BH;
  const int free_sms = device_sms() - scan_ctas;
  const bool overlap = g_prefill_tc_fused && g_prefill_tc_overlap && side && side->stream && free_sms >= 32;
  cudaStream_t s2 = overlap ? side->stream : s;
  if (overlap) {
    if ((e = cudaEventRecord(side->fork, s)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(s2, side->fork, 0)) != cudaSuccess) return e;
  }
  if (g_prefill_tc_fused) {
    // G2 + scan in one kernel: a CTA keeps its 128 x 128 tile of C^T in registers across the chunks of the run
    CUtensorMap mvh, mvl, mkh, mkl;
    if (!make_map3(&mvh, p.w.w3_hi + DH, L, DH, nb, K3, (int64_t)DH * K3) ||
        !make_map3(&mvl, p.w.w3_lo + DH, L, DH, nb, K3, (int64_t)DH * K3) ||
        !make_map3(&mkh, p.w.kt_hi, L, DH, nb, L, (int64_t)DH * L) ||
        !make_map3(&mkl, p.w.kt_lo, L, DH, nb, L, (int64_t)DH * L))
      return cudaErrorUnknown;
    if (g_prefill_scan_split == 4) {
      if ((e = ensure_dyn_smem<&update_scan_kernel<4>>(kSmemScan)) != cudaSuccess) return e;
      e = launch_k(update_scan_kernel<4>, dim3(DH / BN, DH / BM, BH), dim3(64 + 16 * 32), kSmemScan, s, mvh, mvl, mkh, mkl, p);
    } else {
      if ((e = ensure_dyn_smem<&update_scan_kernel<2>>(kSmemScan)) != cudaSuccess) return e;
      e = launch_k(update_scan_kernel<2>, dim3(DH / BN, DH / BM, BH), dim3(64 + 8 * 32), kSmemScan, s, mvh, mvl, mkh, mkl, p);
    }
    if (e != cudaSuccess) return e;
    if (!overlap && (e = launch_k(nscan_kernel, dim3(BH, (DH + 127) / 128), dim3(128), 0, s, p)) != cudaSuccess) return e;
  } else {
    // G2: dC^T = V^T K~ for every chunk, then the element-parallel scan over dC in HBM
    const Operand a = {p.w.w3_hi + DH, p.w.w3_lo + DH, K3, (int64_t)DH * K3};
    const Operand w = {p.w.kt_hi, p.w.kt_lo, L, (int64_t)DH * L};
    const long long sz = (long long)DH * DH;
    const BatchOut o = {p.w.dC, sz * p.nchunk * NH, sz * p.nchunk, sz, DH, NH, p.nchunk, 0};
    if ((e = launch_bgemm(a, w, o, DH, DH, L, nb, s)) != cudaSuccess) return e;
    if ((e = launch_k(scan_kernel, dim3(DH * DH / 1024 + 1, BH), dim3(256), 0, s, p)) != cudaSuccess) return e;
  }
  // G1: S = Q K^T
  {
    const Operand a = {p.w.q_hi, p.w.q_lo, DH, (int64_t)L * DH};
    const Operand w = {p.w.k_hi, p.w.k_lo, DH, (int64_t)L * DH};
    const long long sz = (long long)L * L;
    const BatchOut o = {p.w.Sm, sz * p.nchunk * NH, sz * p.nchunk, sz, L, NH, p.nchunk, 0};
    if ((e = launch_bgemm(a, w, o, L, L, DH, nb, s2, overlap ? free_sms : 0)) != cudaSuccess) return e;
  }
  if ((e = launch_k(pmat_kernel, dim3(nb), dim3(256), 0, s2, p)) != cudaSuccess) return e;
  if (overlap) {
    // q.n needs the chunk-start n (n scan), the row sums of P~ and q: all on this side -- it runs beside the scan as well
    if ((e = launch_k(nscan_kernel, dim3(BH, (DH + 127) / 128), dim3(128), 0, s2, p)) != cudaSuccess) return e;
    if ((e = launch_k(qn_kernel, dim3(nb, L / 8), dim3(256), 0, s2, p)) != cudaSuccess) return e;
    if ((e = cudaEventRecord(side->join, s2)) != cudaSuccess) return e;
    if ((e = cudaStreamWaitEvent(s, side->join, 0)) != cudaSuccess) return e;
  }
  // G3: num = [e^{a_t} q_t | P~_t] [C_c^T | V^T]^T, straight into the numerator rows of the run
  {
    const Operand a = {p.w.a3_hi, p.w.a3_lo, K3, (int64_t)L * K3};
    const Operand w = {p.w.w3_hi, p.w.w3_lo, K3, (int64_t)DH * K3};
    const BatchOut o = {num, (long long)S * inner, DH, (long long)L * inner, inner, NH, p.nchunk, S};
    if ((e = launch_bgemm(a, w, o, L, DH, K3, nb, s)) != cudaSuccess) return e;
  }
  return overlap ? cudaSuccess : launch_k(qn_kernel, dim3(nb, L / 8), dim3(256), 0, s, p);
}

}  // namespace xl
