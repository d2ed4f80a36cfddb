// Context prefill, chunkwise tensor-core sequence cell (SURVEY.md §8 row a11 / BASELINE.json configs[3]).
//
// Same contract as mlstm_cell_seq_kernel (xl_prefill.cu): given the per-token stabilised gates f_t, i_t of a run of S
// tokens and q, k, v, advance the matrix memory  C <- f_t C + (i_t k_t / sqrt(DH)) v_t^T  (and n likewise) through
// the run and emit  num_t = q_t^T C_t,  qn_t = q_t . n_t  for every token — what S calls of the reference's
// recurrent step leave (src/algos/models/decision_xlstm.py:161-165 -> [ext-xlstm] recurrent_step_stabilized_simple).
// Here the recurrence is evaluated chunkwise over chunks of LC = 16 tokens. With F_t = f_1..f_t (chunk-local),
//     num_t  = F_t (q_t^T C_0) + sum_{j<=t} D_tj (q_t . k_j) v_j ,   D_tj = (F_t / F_j) i_j / sqrt(DH)   (<= 1)
//     C_16   = F_16 C_0 + sum_j K~_j v_j^T ,                          K~_j = (F_16 / F_j) i_j k_j / sqrt(DH)
// i.e. three small GEMMs per chunk, which run on the tensor cores. The tiles are far below tcgen05's 64/128-row
// instruction shape (a CTA owns a 32-column slab of one head's C and the contraction over tokens is 16 deep), and
// the chain over chunks is serial per (env, head), so the kernel uses warp-level mma.sync m16n8k16 (bf16 inputs,
// fp32 accumulate) with the accumulators — the C slab itself — resident in registers for the whole run.
// fp32 operands are split into bf16 hi + lo planes and every product is hi*hi + hi*lo + lo*hi (~2^-17 relative),
// the same trick as the tcgen05 Linear (xl_gemm_tc.cu), so the state left agrees with token-by-token stepping to
// ~1e-6 relative. Decay ratios are formed from sums of log f over (j, t], never as F_t / F_j, so a forget gate
// that underflows to 0 is handled without the fallback path the fp32 kernel needs.
//
//   cell_prep_kernel   (parallel over (env*head, chunk)): decay tables, P~ = D (.) (Q K^T) in fp32, and the bf16
//                       hi/lo planes of Q, K~^T, V^T, P~ laid out as the shared-memory image the cell kernel wants
//                       (ldmatrix-friendly, bank-conflict free: padded rows for Q, XOR-swizzled 32-byte rows else).
//   mlstm_cell_mma_kernel  CTA = (env*head, 32-column slab of C | the "n slab" whose v is the unit vector e0).
//                       Warp w owns DKW rows (dk) of the slab as mma accumulators C^T[dv][dk]; a producer warp streams
//                       one chunk image per stage (cp.async.bulk + mbarrier, 2 stages). Per chunk and warp:
//                         G^T  = C^T_w Q_w^T            accumulators -> A fragments in registers (no smem round trip)
//                         G^T *= F_t ; G^T += V^T P~^T  (the intra-chunk part, 4 tiles spread over the warps)
//                         C^T_w = F_16 C^T_w + V^T K~_w
//                       and the per-warp partial numerators meet in shared memory, summed in fixed warp order.
#include <cuda.h>
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

namespace pfm {

constexpr int LC = 16;        // tokens per chunk (the k extent of one m16n8k16)
constexpr int WS = 32;        // dv columns per slab: two m16 tiles
constexpr int kStg = 2;       // chunk images in flight
constexpr int VB = 2 * WS * LC * 2;   // bytes of one slab's V^T block (hi + lo planes): 2048

// byte offsets inside the per-(env*head, chunk) common block
__host__ __device__ constexpr int off_qh(int) { return 0; }
__host__ __device__ constexpr int off_ql(int DH) { return LC * (DH + 8) * 2; }
__host__ __device__ constexpr int off_kh(int DH) { return 2 * LC * (DH + 8) * 2; }
__host__ __device__ constexpr int off_kl(int DH) { return off_kh(DH) + DH * LC * 2; }
__host__ __device__ constexpr int off_ph(int DH) { return off_kl(DH) + DH * LC * 2; }
__host__ __device__ constexpr int off_pl(int DH) { return off_ph(DH) + LC * LC * 2; }
__host__ __device__ constexpr int off_f(int DH) { return off_pl(DH) + LC * LC * 2; }
__host__ __device__ constexpr int common_bytes(int DH) { return off_f(DH) + LC * 4; }   // 128 DH + 1600

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

// fp32 pair -> packed bf16 hi pair + packed bf16 lo pair (x = hi + lo to ~2^-17 relative); low half = first value
__device__ __forceinline__ void split2(float a, float b, uint32_t& hi, uint32_t& lo) {
  const __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  const float2 hf = __bfloat1622float2(h);
  const __nv_bfloat162 l = __floats2bfloat162_rn(a - hf.x, b - hf.y);
  hi = *reinterpret_cast<const uint32_t*>(&h);
  lo = *reinterpret_cast<const uint32_t*>(&l);
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// D (16x8, fp32) += A (16x16, bf16, row) * B (16x8, bf16, col)
__device__ __forceinline__ void mma16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
// hi/lo product: small terms first
__device__ __forceinline__ void mma3(float (&d)[4], const uint32_t (&ah)[4], const uint32_t (&al)[4], uint32_t bh0,
                                     uint32_t bh1, uint32_t bl0, uint32_t bl1) {
  mma16816(d, al, bh0, bh1);
  mma16816(d, ah, bl0, bl1);
  mma16816(d, ah, bh0, bh1);
}

// 16-byte slot of a 32-byte row (16 bf16): half h of row r lives at slot h ^ ((r >> 2) & 1) -> any 8 consecutive
// rows of one half hit 8 different 16-byte bank groups
__host__ __device__ __forceinline__ int swz32(int row, int half) { return row * 32 + ((half ^ ((row >> 2) & 1)) << 4); }

struct PrepParams {
  const float* q;       // [B*S, inner]
  const float* k;
  const float* v;
  const float* fseq;    // [B*NH, S]
  const float* iseq;
  uint8_t* common;      // [B*NH, S/LC] blocks of common_bytes(DH)
  uint8_t* vblk;        // [B*NH, S/LC, DH/WS] blocks of VB bytes
  int B, S, NH, DH, inner;
};

// ------------------------------------------------------------------------------------------------
// Chunk preparation: grid = (S / LC, B * NH), 256 threads.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) cell_prep_kernel(PrepParams p) {
  __shared__ float s_lf[LC], s_i[LC], s_F[LC], s_ks[LC];
  __shared__ float s_D[LC][LC + 1], s_S[LC][LC + 1];
  const int c = blockIdx.x, bh = blockIdx.y;
  const int NH = p.NH, DH = p.DH, S = p.S, inner = p.inner;
  const int b = bh / NH, hd = bh - b * NH;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int nch = S / LC;
  const int64_t row0 = (int64_t)b * S + (int64_t)c * LC;
  const int hoff = hd * DH;
  const float kscale = rsqrtf((float)DH);
  uint8_t* cb = p.common + ((int64_t)bh * nch + c) * common_bytes(DH);
  uint8_t* vb = p.vblk + ((int64_t)bh * nch + c) * (int64_t)(DH / WS) * VB;
  pdl_wait();
  pdl_trigger();

  if (tid < LC) {
    const float f = p.fseq[(int64_t)bh * S + (int64_t)c * LC + tid];
    s_lf[tid] = f > 0.f ? fmaxf(logf(f), -200.f) : -200.f;
    s_i[tid] = p.iseq[(int64_t)bh * S + (int64_t)c * LC + tid];
  }
  __syncthreads();
  if (tid < LC) {
    const int t = tid;
    float a = 0.f;
    for (int s = 0; s <= t; ++s) a += s_lf[s];
    s_F[t] = expf(a);                                   // F_t = f_0 .. f_t
    float lr = 0.f;                                     // log(f_{j+1} .. f_t), built from the t end: no cancellation
    for (int j = t; j >= 0; --j) {
      s_D[t][j] = expf(lr) * s_i[j] * kscale;
      lr += s_lf[j];
    }
    for (int j = t + 1; j < LC; ++j) s_D[t][j] = 0.f;
  } else if (tid >= 32 && tid < 32 + LC) {
    const int j = tid - 32;
    float lr = 0.f;
    for (int s = j + 1; s < LC; ++s) lr += s_lf[s];
    s_ks[j] = expf(lr) * s_i[j] * kscale;               // (F_16 / F_j) i_j / sqrt(DH)
  }
  __syncthreads();

  // S[t][j] = q_t . k_j (j <= t), fp32; warp w takes rows w and LC-1-w (17 dot products each)
#pragma unroll 1
  for (int rr = 0; rr < 2; ++rr) {
    const int t = rr == 0 ? warp : LC - 1 - warp;
    const float* qp = p.q + (row0 + t) * inner + hoff;
#pragma unroll 1
    for (int j = 0; j <= t; ++j) {
      const float* kp = p.k + (row0 + j) * inner + hoff;
      float a = 0.f;
      for (int e = lane; e < DH; e += 32) a = fmaf(qp[e], kp[e], a);
      a = warp_sum(a);
      if (lane == 0) s_S[t][j] = a;
    }
  }

  // Q planes [LC][DH + 8]
  {
    __nv_bfloat16* qh = reinterpret_cast<__nv_bfloat16*>(cb + off_qh(DH));
    __nv_bfloat16* ql = reinterpret_cast<__nv_bfloat16*>(cb + off_ql(DH));
    const int rs = DH + 8, q4 = DH >> 2;
    for (int idx = tid; idx < LC * q4; idx += 256) {
      const int t = idx / q4, c4 = idx - t * q4;
      const float4 x = *reinterpret_cast<const float4*>(p.q + (row0 + t) * inner + hoff + 4 * c4);
      uint32_t h0, l0, h1, l1;
      split2(x.x, x.y, h0, l0);
      split2(x.z, x.w, h1, l1);
      *reinterpret_cast<uint2*>(qh + t * rs + 4 * c4) = make_uint2(h0, h1);
      *reinterpret_cast<uint2*>(ql + t * rs + 4 * c4) = make_uint2(l0, l1);
    }
    if (tid < LC) {                                     // the 8 pad columns are never read; keep them defined
      *reinterpret_cast<uint4*>(qh + tid * rs + DH) = make_uint4(0, 0, 0, 0);
      *reinterpret_cast<uint4*>(ql + tid * rs + DH) = make_uint4(0, 0, 0, 0);
    }
  }
  // K~^T planes [DH][LC] and V^T planes [DH/WS][WS][LC]: thread = one dk (dv) row, 16 tokens -> one 32-byte row
  for (int r = tid; r < DH; r += 256) {
    uint32_t h[8], l[8];
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float x0 = p.k[(row0 + 2 * jj) * inner + hoff + r] * s_ks[2 * jj];
      const float x1 = p.k[(row0 + 2 * jj + 1) * inner + hoff + r] * s_ks[2 * jj + 1];
      split2(x0, x1, h[jj], l[jj]);
    }
    uint8_t* kh = cb + off_kh(DH);
    uint8_t* kl = cb + off_kl(DH);
    *reinterpret_cast<uint4*>(kh + swz32(r, 0)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(kh + swz32(r, 1)) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(kl + swz32(r, 0)) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(kl + swz32(r, 1)) = make_uint4(l[4], l[5], l[6], l[7]);
#pragma unroll
    for (int jj = 0; jj < 8; ++jj) {
      const float x0 = p.v[(row0 + 2 * jj) * inner + hoff + r];
      const float x1 = p.v[(row0 + 2 * jj + 1) * inner + hoff + r];
      split2(x0, x1, h[jj], l[jj]);
    }
    uint8_t* vh = vb + (r / WS) * VB;
    uint8_t* vl = vh + VB / 2;
    const int rv = r % WS;
    *reinterpret_cast<uint4*>(vh + swz32(rv, 0)) = make_uint4(h[0], h[1], h[2], h[3]);
    *reinterpret_cast<uint4*>(vh + swz32(rv, 1)) = make_uint4(h[4], h[5], h[6], h[7]);
    *reinterpret_cast<uint4*>(vl + swz32(rv, 0)) = make_uint4(l[0], l[1], l[2], l[3]);
    *reinterpret_cast<uint4*>(vl + swz32(rv, 1)) = make_uint4(l[4], l[5], l[6], l[7]);
  }
  __syncthreads();
  // P~[t][j] = D_tj S_tj (0 above the diagonal), [LC][LC] planes
  {
    const int t = tid >> 4, j = tid & 15;
    const float x = j <= t ? s_D[t][j] * s_S[t][j] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(x);
    const __nv_bfloat16 l = __float2bfloat16_rn(x - __bfloat162float(h));
    const int o = swz32(t, j >> 3) + (j & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(cb + off_ph(DH) + o) = h;
    *reinterpret_cast<__nv_bfloat16*>(cb + off_pl(DH) + o) = l;
  }
  if (tid < LC) reinterpret_cast<float*>(cb + off_f(DH))[tid] = s_F[tid];
}

struct CellMmaParams {
  float* C;                 // [B, NH, DH/wc, DH, wc] slab-major state (xl_state_step.cu layout)
  float* n;                 // [B, NH, DH]
  const uint8_t* common;
  const uint8_t* vblk;
  float* num;               // [B*S, inner]
  float* qn;                // [B*S, NH]
  int B, S, NH, inner;
};

// ------------------------------------------------------------------------------------------------
// Sequence cell on mma.sync. grid = B * NH * (DH / WS + 1), block = (NW + 1) warps, 1 CTA / SM.
// ------------------------------------------------------------------------------------------------
template <int DKW, int NW>
__global__ void __launch_bounds__((NW + 1) * 32, 1) mlstm_cell_mma_kernel(CellMmaParams p) {
  constexpr int DH = DKW * NW;
  constexpr int NT = DKW / 8;               // accumulator n-tiles (8 dk each) per warp
  constexpr int KS = DKW / 16;              // k-steps of the readout = n-tile pairs of the update
  constexpr int QRS = (DH + 8) * 2;         // Q row stride, bytes
  constexpr int CB = common_bytes(DH);
  constexpr int STAGE = CB + VB;
  constexpr int REDW = LC * 33;             // floats per warp in the reduction buffer, rows padded to 33
  constexpr int OQH = off_qh(DH), OQL = off_ql(DH), OKH = off_kh(DH), OKL = off_kl(DH), OPH = off_ph(DH),
                OPL = off_pl(DH), OF = off_f(DH);
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2 * kStg];
  const int NH = p.NH, S = p.S;
  constexpr int nsl = DH / WS;
  const int bh = blockIdx.x / (nsl + 1);
  const int slab = blockIdx.x - bh * (nsl + 1);
  const bool n_slab = slab == nsl;
  const int b = bh / NH, hd = bh - b * NH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const uint32_t sraw = smem_u32(smem_raw);
  const uint32_t sbase = (sraw + 127u) & ~127u;
  uint8_t* sm = smem_raw + (sbase - sraw);
  uint8_t* e0t = sm + kStg * STAGE;                               // the n slab's V^T block (v = e0)
  float* red = reinterpret_cast<float*>(e0t + VB);                // [2][NW][REDW]
  const uint32_t bar0 = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStg + s); };
  const int nch = S / LC;                                         // S % LC == 0 (host-checked)

  if (tid == 0) {
    for (int s = 0; s < kStg; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), NW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (n_slab) {
    // hi plane: row 0 (dv = 0) = 1.0 for all 16 tokens, everything else (and the lo plane) 0
    for (int i = tid; i < VB / 4; i += (NW + 1) * 32) reinterpret_cast<uint32_t*>(e0t)[i] = i < 8 ? 0x3F803F80u : 0u;
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (warp == NW) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      const uint8_t* src = p.common + (int64_t)bh * nch * CB;
      const uint8_t* vsrc = p.vblk + ((int64_t)bh * nch * nsl + slab) * VB;
      for (int c = 0; c < nch; ++c) {
        mbar_wait(empty_bar(s), ph ^ 1);
        const uint32_t dst = sbase + (uint32_t)s * STAGE;
        mbar_expect_tx(full_bar(s), (uint32_t)(CB + (n_slab ? 0 : VB)));
        for (int o = 0; o < CB; o += 16384) {
          const int sz = CB - o < 16384 ? CB - o : 16384;
          bulk_copy_g2s(dst + o, src + (int64_t)c * CB + o, (uint32_t)sz, full_bar(s));
        }
        if (!n_slab) bulk_copy_g2s(dst + CB, vsrc + (int64_t)c * nsl * VB, (uint32_t)VB, full_bar(s));
        if (++s == kStg) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ===== consumers =====
  const int g = lane >> 2, tq = lane & 3;
  const int mi = lane >> 3, r8 = lane & 7;                        // ldmatrix: this lane addresses row r8 of matrix mi
  const int dk0 = warp * DKW;
  constexpr int wc = (DH % 128 == 0) ? 128 : DH;                  // slab width of the state layout in HBM
  constexpr int CSl = DH / wc;
  auto c_ptr = [&](int row, int col) {
    const int hs = col / wc;
    return p.C + ((int64_t)(bh * CSl + hs) * DH + row) * wc + (col - hs * wc);
  };
  // C^T accumulators: acc[mt][nt] = tile (dv 16 mt .. +15, dk dk0 + 8 nt .. +7); element (g | g+8, 2 tq | 2 tq + 1)
  float acc[2][NT][4];
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int dv = slab * WS + 16 * mt + g, dk = dk0 + 8 * nt + 2 * tq;
      if (!n_slab) {
        acc[mt][nt][0] = *c_ptr(dk, dv);
        acc[mt][nt][1] = *c_ptr(dk + 1, dv);
        acc[mt][nt][2] = *c_ptr(dk, dv + 8);
        acc[mt][nt][3] = *c_ptr(dk + 1, dv + 8);
      } else {
        const bool row0 = mt == 0 && g == 0;                      // n rides in row dv = 0
        acc[mt][nt][0] = row0 ? p.n[(int64_t)bh * DH + dk] : 0.f;
        acc[mt][nt][1] = row0 ? p.n[(int64_t)bh * DH + dk + 1] : 0.f;
        acc[mt][nt][2] = 0.f;
        acc[mt][nt][3] = 0.f;
      }
    }

  int s = 0;
  uint32_t ph = 0;
  for (int c = 0; c < nch; ++c) {
    const uint32_t st = sbase + (uint32_t)s * STAGE;
    const float* Fv = reinterpret_cast<const float*>(sm + (size_t)s * STAGE + OF);
    mbar_wait(full_bar(s), ph);

    // ---- (a) G^T[dv][t] = sum_dk C^T[dv][dk] Q[t][dk] over this warp's dk range
    float N[2][2][4];
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt)
#pragma unroll
        for (int e = 0; e < 4; ++e) N[mt][nt][e] = 0.f;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      uint32_t ah[2][4], al[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        split2(acc[mt][2 * ks][0], acc[mt][2 * ks][1], ah[mt][0], al[mt][0]);
        split2(acc[mt][2 * ks][2], acc[mt][2 * ks][3], ah[mt][1], al[mt][1]);
        split2(acc[mt][2 * ks + 1][0], acc[mt][2 * ks + 1][1], ah[mt][2], al[mt][2]);
        split2(acc[mt][2 * ks + 1][2], acc[mt][2 * ks + 1][3], ah[mt][3], al[mt][3]);
      }
      uint32_t qh[4], ql[4];
      const uint32_t qa = st + OQH + (uint32_t)(((mi >> 1) * 8 + r8) * QRS + (dk0 + 16 * ks + (mi & 1) * 8) * 2);
      ldmatrix_x4(qh, qa);
      ldmatrix_x4(ql, qa + (OQL - OQH));
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          mma3(N[mt][nt], ah[mt], al[mt], qh[2 * nt], qh[2 * nt + 1], ql[2 * nt], ql[2 * nt + 1]);
    }
    // ---- (b) decay from the chunk start: column t scaled by F_t
#pragma unroll
    for (int nt = 0; nt < 2; ++nt) {
      const float f0 = Fv[8 * nt + 2 * tq], f1 = Fv[8 * nt + 2 * tq + 1];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        N[mt][nt][0] *= f0; N[mt][nt][1] *= f1; N[mt][nt][2] *= f0; N[mt][nt][3] *= f1;
      }
    }
    // V^T fragments (A operand of the intra-chunk product and of the update)
    uint32_t vh[2][4], vl[2][4];
    {
      const uint32_t vt = n_slab ? sbase + (uint32_t)(kStg * STAGE) : st + CB;
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        const int row = 16 * mt + (mi & 1) * 8 + r8;
        const uint32_t va = vt + (uint32_t)swz32(row, mi >> 1);
        ldmatrix_x4(vh[mt], va);
        ldmatrix_x4(vl[mt], va + VB / 2);
      }
    }
    // ---- (c) intra-chunk part: G^T[dv][t] += sum_j V^T[dv][j] P~[t][j]; the 4 tiles go to warps 0..3 (mod NW)
    {
      bool mine = false;
#pragma unroll
      for (int i = 0; i < 4; ++i) mine |= (i % NW) == warp;
      if (mine) {
        uint32_t pH[4], pL[4];
        const int trow = (mi >> 1) * 8 + r8;
        const uint32_t pa = st + OPH + (uint32_t)swz32(trow, mi & 1);
        ldmatrix_x4(pH, pa);
        ldmatrix_x4(pL, pa + (OPL - OPH));
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if ((i % NW) == warp) {
            const int mt = i >> 1, nt = i & 1;
            mma3(N[mt][nt], vh[mt], vl[mt], pH[2 * nt], pH[2 * nt + 1], pL[2 * nt], pL[2 * nt + 1]);
          }
        }
      }
    }
    // ---- (d) partial numerators -> shared memory, [t][dv] rows of 33
    float* rb = red + (size_t)(c & 1) * (NW * REDW) + (size_t)warp * REDW;
#pragma unroll
    for (int mt = 0; mt < 2; ++mt)
#pragma unroll
      for (int nt = 0; nt < 2; ++nt) {
        const int t0 = 8 * nt + 2 * tq, dv = 16 * mt + g;
        rb[t0 * 33 + dv] = N[mt][nt][0];
        rb[(t0 + 1) * 33 + dv] = N[mt][nt][1];
        rb[t0 * 33 + dv + 8] = N[mt][nt][2];
        rb[(t0 + 1) * 33 + dv + 8] = N[mt][nt][3];
      }
    // ---- (e) C^T <- F_16 C^T + V^T K~
    {
      const float FL = Fv[LC - 1];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int nt = 0; nt < NT; ++nt)
#pragma unroll
          for (int e = 0; e < 4; ++e) acc[mt][nt][e] *= FL;
#pragma unroll
      for (int pz = 0; pz < KS; ++pz) {
        uint32_t kh[4], kl[4];
        const int row = dk0 + 16 * pz + (mi >> 1) * 8 + r8;
        const uint32_t ka = st + OKH + (uint32_t)swz32(row, mi & 1);
        ldmatrix_x4(kh, ka);
        ldmatrix_x4(kl, ka + (OKL - OKH));
#pragma unroll
        for (int mt = 0; mt < 2; ++mt)
#pragma unroll
          for (int q2 = 0; q2 < 2; ++q2)
            mma3(acc[mt][2 * pz + q2], vh[mt], vl[mt], kh[2 * q2], kh[2 * q2 + 1], kl[2 * q2], kl[2 * q2 + 1]);
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));
    if (++s == kStg) { s = 0; ph ^= 1; }
    asm volatile("bar.sync 1, %0;" ::"n"(NW * 32) : "memory");
    // ---- (f) sum the warps' partials in fixed order; dv fastest -> 128-byte rows of num
    {
      const float* rs = red + (size_t)(c & 1) * (NW * REDW);
      for (int o = tid; o < LC * WS; o += NW * 32) {
        const int t = o >> 5, dv = o & 31;
        float sum = 0.f;
#pragma unroll
        for (int y = 0; y < NW; ++y) sum += rs[y * REDW + t * 33 + dv];
        const int64_t row = (int64_t)b * S + (int64_t)c * LC + t;
        if (!n_slab) p.num[row * p.inner + hd * DH + slab * WS + dv] = sum;
        else if (dv == 0) p.qn[row * NH + hd] = sum;
      }
    }
  }
  // write the slab back
#pragma unroll
  for (int mt = 0; mt < 2; ++mt)
#pragma unroll
    for (int nt = 0; nt < NT; ++nt) {
      const int dv = slab * WS + 16 * mt + g, dk = dk0 + 8 * nt + 2 * tq;
      if (!n_slab) {
        *c_ptr(dk, dv) = acc[mt][nt][0];
        *c_ptr(dk + 1, dv) = acc[mt][nt][1];
        *c_ptr(dk, dv + 8) = acc[mt][nt][2];
        *c_ptr(dk + 1, dv + 8) = acc[mt][nt][3];
      } else if (mt == 0 && g == 0) {
        p.n[(int64_t)bh * DH + dk] = acc[mt][nt][0];
        p.n[(int64_t)bh * DH + dk + 1] = acc[mt][nt][1];
      }
    }
}

}  // namespace pfm

// (DKW, NW) instantiations: head dims 640 / 512 / 384 / 256 (the presets) and the small test heads 128 / 64 / 32
#define XL_CELL_MMA_CASES(X) X(64, 10) X(64, 8) X(32, 12) X(32, 8) X(32, 4) X(32, 2) X(32, 1)

static bool cell_mma_plan(int DH, int* DKW, int* NW) {
#define XL_MMA_MATCH(dkw, nw) \
  if (DH == dkw * nw) { *DKW = dkw; *NW = nw; return true; }
  XL_CELL_MMA_CASES(XL_MMA_MATCH)
#undef XL_MMA_MATCH
  return false;
}

bool prefill_cell_mma_supported(int DH) {
  int a, b;
  return cell_mma_plan(DH, &a, &b);
}

int prefill_cell_mma_chunk() { return pfm::LC; }

// bytes of the two prepared-operand buffers for `rows` = B * S token rows (S % 16 == 0)
void prefill_cell_mma_ws(int rows, int NH, int DH, size_t* common_bytes, size_t* vblk_bytes) {
  const size_t blocks = (size_t)(rows / pfm::LC + 1) * NH;
  *common_bytes = blocks * (size_t)pfm::common_bytes(DH);
  *vblk_bytes = blocks * (size_t)(DH / pfm::WS) * pfm::VB;
}

template <int DKW, int NW>
static cudaError_t launch_cell_mma_T(const pfm::CellMmaParams& p, cudaStream_t s) {
  constexpr int DH = DKW * NW;
  const size_t smem = 128 + (size_t)pfm::kStg * (pfm::common_bytes(DH) + pfm::VB) + pfm::VB +
                      sizeof(float) * 2 * (size_t)NW * pfm::LC * 33;
  if (cudaError_t e = ensure_dyn_smem<&pfm::mlstm_cell_mma_kernel<DKW, NW>>(smem); e != cudaSuccess) return e;
  const int grid = p.B * p.NH * (DH / pfm::WS + 1);
  return launch_k(pfm::mlstm_cell_mma_kernel<DKW, NW>, dim3(grid), dim3((NW + 1) * 32), smem, s, p);
}

cudaError_t launch_cell_mma(float* C, float* n, const float* q, const float* k, const float* v, const float* fseq,
                            const float* iseq, float* num, float* qn, void* common, void* vblk, int B, int S, int NH,
                            int DH, int inner, cudaStream_t s) {
  int DKW, NW;
  if (!cell_mma_plan(DH, &DKW, &NW) || S % pfm::LC != 0 || S <= 0) return cudaErrorInvalidValue;
  pfm::PrepParams pp;
  pp.q = q; pp.k = k; pp.v = v; pp.fseq = fseq; pp.iseq = iseq;
  pp.common = (uint8_t*)common; pp.vblk = (uint8_t*)vblk;
  pp.B = B; pp.S = S; pp.NH = NH; pp.DH = DH; pp.inner = inner;
  cudaError_t e = launch_k(pfm::cell_prep_kernel, dim3(S / pfm::LC, B * NH), dim3(256), 0, s, pp);
  if (e != cudaSuccess) return e;
  pfm::CellMmaParams cp;
  cp.C = C; cp.n = n; cp.common = (const uint8_t*)common; cp.vblk = (const uint8_t*)vblk; cp.num = num; cp.qn = qn;
  cp.B = B; cp.S = S; cp.NH = NH; cp.inner = inner;
#define XL_MMA_LAUNCH(dkw, nw) \
  if (DKW == dkw && NW == nw) return launch_cell_mma_T<dkw, nw>(cp, s);
  XL_CELL_MMA_CASES(XL_MMA_LAUNCH)
#undef XL_MMA_LAUNCH
  return cudaErrorInvalidValue;
}

}  // namespace xl
