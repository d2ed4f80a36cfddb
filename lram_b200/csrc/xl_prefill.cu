// Context prefill: the encoder over S tokens per env, layer by layer, leaving exactly the recurrent state that
// S calls of xLSTMBlockStack.step would leave (the reference reaches long contexts only by stepping:
// src/algos/models/decision_xlstm.py:161-165; its `chunkwise_step` hook :158-159 calls layers.step with S > 1,
// which pip xlstm 1.0.x does not implement). SURVEY.md §8 row a11 / BASELINE.json configs[3].
//
// Because every token of the context is known up front, the stack runs LAYER-major over chunks of tokens:
//   LayerNorm rows -> proj_up as ONE tcgen05 GEMM over all rows of the chunk -> sequence conv/qkv/gates ->
//   gate scan (m, f, i per token) -> sequence cell -> sequence finalize -> proj_down GEMM (+ residual).
// The projections are real GEMMs here (M = envs x chunk tokens rows) and run on the tensor cores. The cell
// keeps the matrix memory ON CHIP for the whole chunk: a CTA owns 16 columns (dv) of one (env, head) C in
// registers, streams the chunk's (q,k) pairs / v / gates through a TMA-fed shared-memory ring, and touches HBM
// for C once per chunk instead of once per token. The arithmetic per element is the recurrent step's
// (c <- f c + k v i, num += q c), in token order, so the state it leaves is the stepping state to rounding.
// The normaliser n rides along as one more column slab whose "v" is the unit vector e0 (n = f n + i k is the
// same recurrence), which also yields q.n per token.
#include <cuda.h>
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

namespace pf {

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

// ------------------------------------------------------------------------------------------------
// Sequence pre-cell kernel: CausalConv1d over the chunk (window seeded from conv_state) + SiLU + headwise
// q/k/v + partial gate pre-activations. Same arithmetic as conv_qkv_gates_kernel (xl_elementwise.cu), token by
// token. grid = (NCH channel chunks, B envs, token runs); a thread owns one 4-channel block for a run of
// tokens and keeps the conv window in registers; gate partials are reduced 4 tokens at a time.
// ------------------------------------------------------------------------------------------------
template <int KS, int NH>
__global__ void __launch_bounds__(128, 2) conv_qkv_gates_seq_kernel(ConvQkvParams p, int S, int run) {
  constexpr int TG = 4;                                  // tokens per gate-reduction group
  __shared__ float red[2 * NH * TG * 4];
  const int b = blockIdx.y, chunk = blockIdx.x;
  const int s_begin = blockIdx.z * run;
  const int s_end = min(S, s_begin + run);
  const int inner = p.inner;
  const int nblk = inner >> 2;
  const int blk_per_chunk = (nblk + p.NCH - 1) / p.NCH;
  const int j = chunk * blk_per_chunk + threadIdx.x;
  const bool active = threadIdx.x < blk_per_chunk && j < nblk;
  const int c = 4 * (active ? j : 0);
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  const int nw = (blockDim.x + 31) >> 5;

  float win[KS][4];
  float cw[4][KS];
  float cbv[4] = {0.f, 0.f, 0.f, 0.f};
  float wq[16], wk[16], wv[16];
  if (active) {
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
      for (int r = 0; r < KS; ++r) cw[ch][r] = p.conv_w[(int64_t)(c + ch) * KS + r];
    const float4 cb = *reinterpret_cast<const float4*>(p.conv_b + c);
    cbv[0] = cb.x; cbv[1] = cb.y; cbv[2] = cb.z; cbv[3] = cb.w;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const float4 a4 = reinterpret_cast<const float4*>(p.wq + (int64_t)j * 16)[i];
      const float4 k4 = reinterpret_cast<const float4*>(p.wk + (int64_t)j * 16)[i];
      const float4 v4 = reinterpret_cast<const float4*>(p.wv + (int64_t)j * 16)[i];
      wq[4 * i] = a4.x; wq[4 * i + 1] = a4.y; wq[4 * i + 2] = a4.z; wq[4 * i + 3] = a4.w;
      wk[4 * i] = k4.x; wk[4 * i + 1] = k4.y; wk[4 * i + 2] = k4.z; wk[4 * i + 3] = k4.w;
      wv[4 * i] = v4.x; wv[4 * i + 1] = v4.y; wv[4 * i + 2] = v4.z; wv[4 * i + 3] = v4.w;
    }
  }
  pdl_wait();
  pdl_trigger();
  if (active) {
    // window rows 1..KS-1 = the KS-1 inputs before token s_begin: earlier rows of this chunk, or the carried
    // conv_state (rows = last KS inputs, oldest first) for tokens before the chunk
#pragma unroll
    for (int r = 1; r < KS; ++r) {
      const int s = s_begin - (KS - r);
      float4 w4;
      if (s >= 0) w4 = *reinterpret_cast<const float4*>(p.u + ((int64_t)b * S + s) * 2 * inner + c);
      else w4 = *reinterpret_cast<const float4*>(p.conv_state + ((int64_t)b * KS + (KS + s)) * inner + c);
      win[r][0] = w4.x; win[r][1] = w4.y; win[r][2] = w4.z; win[r][3] = w4.w;
    }
#pragma unroll
    for (int ch = 0; ch < 4; ++ch) win[0][ch] = 0.f;
  }
  for (int s0 = s_begin; s0 < s_end; s0 += TG) {
    float gi[NH][TG], gf[NH][TG];
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int t = 0; t < TG; ++t) gi[h][t] = gf[h][t] = 0.f;
    if (active) {
#pragma unroll
      for (int t = 0; t < TG; ++t) {
        const int s = s0 + t;
        if (s < s_end) {
          const int64_t row = (int64_t)b * S + s;
          const float4 x4 = *reinterpret_cast<const float4*>(p.u + row * 2 * inner + c);
          const float xm[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
          for (int r = 0; r < KS - 1; ++r)
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) win[r][ch] = win[r + 1][ch];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) win[KS - 1][ch] = xm[ch];
          float a[4];
#pragma unroll
          for (int ch = 0; ch < 4; ++ch) {
            float acc = 0.f;
#pragma unroll
            for (int r = 0; r < KS; ++r) acc = fmaf(win[r][ch], cw[ch][r], acc);
            a[ch] = silu(acc + cbv[ch]);
          }
          float q[4], k[4], v[4];
#pragma unroll
          for (int o = 0; o < 4; ++o) {
            float sq = 0.f, sk = 0.f, sv = 0.f;
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
              sq = fmaf(a[dd], wq[4 * o + dd], sq);
              sk = fmaf(a[dd], wk[4 * o + dd], sk);
              sv = fmaf(xm[dd], wv[4 * o + dd], sv);
            }
            q[o] = sq; k[o] = sk; v[o] = sv;
          }
          float* qk = p.qk + (row * inner + c) * 2;
          *reinterpret_cast<float4*>(qk) = make_float4(q[0], k[0], q[1], k[1]);
          *reinterpret_cast<float4*>(qk + 4) = make_float4(q[2], k[2], q[3], k[3]);
          *reinterpret_cast<float4*>(p.v + row * inner + c) = make_float4(v[0], v[1], v[2], v[3]);
          *reinterpret_cast<float4*>(p.act + row * inner + c) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
          for (int h = 0; h < NH; ++h) {
            // gate weights of this 4-channel block: L1-resident after the first token of the run
            const float* wi = p.wi + (int64_t)h * 3 * inner + c;
            const float* wf = p.wf + (int64_t)h * 3 * inner + c;
            const float4 iq = *reinterpret_cast<const float4*>(wi);
            const float4 ik = *reinterpret_cast<const float4*>(wi + inner);
            const float4 iv = *reinterpret_cast<const float4*>(wi + 2 * inner);
            const float4 fq = *reinterpret_cast<const float4*>(wf);
            const float4 fk = *reinterpret_cast<const float4*>(wf + inner);
            const float4 fv = *reinterpret_cast<const float4*>(wf + 2 * inner);
            float si = q[0] * iq.x + q[1] * iq.y + q[2] * iq.z + q[3] * iq.w;
            si += k[0] * ik.x + k[1] * ik.y + k[2] * ik.z + k[3] * ik.w;
            si += v[0] * iv.x + v[1] * iv.y + v[2] * iv.z + v[3] * iv.w;
            float sf = q[0] * fq.x + q[1] * fq.y + q[2] * fq.z + q[3] * fq.w;
            sf += k[0] * fk.x + k[1] * fk.y + k[2] * fk.z + k[3] * fk.w;
            sf += v[0] * fv.x + v[1] * fv.y + v[2] * fv.z + v[3] * fv.w;
            gi[h][t] = si;
            gf[h][t] = sf;
          }
        }
      }
    }
    // chunk-level partial sums of the 4 tokens: warp tree, then fixed-order sum over the (<= 4) warps
#pragma unroll
    for (int h = 0; h < NH; ++h)
#pragma unroll
      for (int t = 0; t < TG; ++t) {
        const float si = warp_sum(gi[h][t]);
        const float sf = warp_sum(gf[h][t]);
        if (lane == 0) {
          red[((h * 2 + 0) * TG + t) * 4 + wid] = si;
          red[((h * 2 + 1) * TG + t) * 4 + wid] = sf;
        }
      }
    __syncthreads();
    if ((int)threadIdx.x < NH * 2 * TG) {
      const int h = threadIdx.x / (2 * TG);
      const int rem = threadIdx.x - h * 2 * TG;
      const int g = rem / TG, t = rem - g * TG;
      if (s0 + t < s_end) {
        float s = 0.f;
        for (int w = 0; w < nw; ++w) s += red[((h * 2 + g) * TG + t) * 4 + w];
        p.gate_part[(((int64_t)b * S + s0 + t) * p.NCH + chunk) * 2 * NH + g * NH + h] = s;
      }
    }
    __syncthreads();
  }
}

// new conv_state = the last KS x_m rows of the chunk (S >= KS): separate kernel, because the first token run of
// the conv kernel still reads the old window while its last run would overwrite it
__global__ void __launch_bounds__(256) conv_state_seq_kernel(const float* __restrict__ u, float* __restrict__ conv_state,
                                                             int B, int S, int KS, int inner) {
  pdl_wait();
  pdl_trigger();
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t n4 = (int64_t)B * KS * (inner >> 2);
  if (i >= n4) return;
  const int c4 = (int)(i % (inner >> 2));
  const int r = (int)((i / (inner >> 2)) % KS);
  const int b = (int)(i / ((int64_t)KS * (inner >> 2)));
  const float4 v = reinterpret_cast<const float4*>(u + ((int64_t)b * S + (S - KS + r)) * 2 * inner)[c4];
  reinterpret_cast<float4*>(conv_state + ((int64_t)b * KS + r) * inner)[c4] = v;
}

// ------------------------------------------------------------------------------------------------
// Gate scan: per (env, head) the stabiliser recurrence over the chunk's tokens,
//   lf = logsigmoid(f~);  m' = max(lf + m, i~);  f = exp(lf + m - m');  i = exp(i~ - m')
// with exactly the arithmetic of compute_gates() (xl_state_step.cu). One warp per (env, head): 32 tokens'
// pre-activations are summed / log-sigmoided in parallel, the max-plus chain runs over warp shuffles, f and i
// are again computed in parallel. Outputs are [B*NH][S] so that a (env, head)'s tokens are contiguous.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) gate_scan_seq_kernel(const float* __restrict__ gate_part,
                                                            const float* __restrict__ igate_b,
                                                            const float* __restrict__ fgate_b, float* __restrict__ m_state,
                                                            float* __restrict__ fseq, float* __restrict__ iseq,
                                                            float* __restrict__ mseq, int B, int S, int NH, int NCH) {
  pdl_wait();
  pdl_trigger();
  const int bh = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (bh >= B * NH) return;
  const int b = bh / NH, hd = bh - b * NH;
  float m = m_state[bh];
  const float bi = igate_b ? igate_b[hd] : 0.f, bf = fgate_b ? fgate_b[hd] : 0.f;
  for (int s0 = 0; s0 < S; s0 += 32) {
    const int s = s0 + lane;
    float ig = -INFINITY, lf = 0.f;
    if (s < S) {
      const float* gp = gate_part + ((int64_t)b * S + s) * NCH * 2 * NH + hd;
      float si = 0.f, sf = 0.f;
      for (int c = 0; c < NCH; ++c) {      // fixed order, as compute_gates()
        si += gp[c * 2 * NH];
        sf += gp[c * 2 * NH + NH];
      }
      ig = si + bi;
      lf = log_sigmoid(sf + bf);
    }
    float mprev_mine = 0.f, mnew_mine = 0.f;
    const int cnt = min(32, S - s0);
    for (int jj = 0; jj < cnt; ++jj) {
      const float lfj = __shfl_sync(0xffffffffu, lf, jj);
      const float igj = __shfl_sync(0xffffffffu, ig, jj);
      const float mn = fmaxf(lfj + m, igj);
      if (jj == lane) { mprev_mine = m; mnew_mine = mn; }
      m = mn;
    }
    if (s < S) {
      const int64_t o = (int64_t)bh * S + s;
      fseq[o] = expf(lf + mprev_mine - mnew_mine);
      iseq[o] = expf(ig - mnew_mine);
      mseq[o] = mnew_mine;
    }
  }
  if (lane == 0) m_state[bh] = m;
}

// ------------------------------------------------------------------------------------------------
// Sequence cell: CTA = (env*head, 16-column slab of C, or the extra "n slab"); C slab [DH x 16] lives in
// registers for the whole chunk. 8 consumer warps: thread = (row lane rl in 0..63, column group cg in 0..3),
// rows rl + 64 r. A producer warp streams the chunk's tokens through a 3-deep ring of 4-token stages
// ((q,k) pairs of the whole head, the slab's v, f and i) with bulk copies + mbarriers.
// ------------------------------------------------------------------------------------------------
constexpr int kSlabW = 16;
constexpr int kTB = 4;          // tokens per stage
constexpr int kStages = 3;
constexpr int kConsumers = 256;
constexpr int kBlock = kConsumers + 32;

struct CellSeqParams {
  float* C;               // [B, NH, DH/Wc, DH, Wc] slab-major state
  float* n;               // [B, NH, DH]
  const float* qk;        // [B*S, NH, DH, 2]
  const float* v;         // [B*S, inner]
  const float* fseq;      // [B*NH, S]
  const float* iseq;      // [B*NH, S]
  float* num;             // [B*S, inner]   q^T C per token (un-normalised)
  float* qn;              // [B*S, NH]      q . n per token
  int B, S, NH, DH, inner;
};

__host__ __device__ inline int cell_stage_floats(int DH) { return kTB * DH * 2 + kTB * kSlabW + 2 * kTB; }

template <int R>
__global__ void __launch_bounds__(kBlock, (R <= 8 ? 2 : 1)) mlstm_cell_seq_kernel(CellSeqParams p) {
  extern __shared__ uint8_t smem_raw[];
  __shared__ __align__(8) uint64_t s_bar[2 * kStages];
  const int DH = p.DH, NH = p.NH, S = p.S;
  const int nsl = DH / kSlabW;
  const int bh = blockIdx.x / (nsl + 1);
  const int slab = blockIdx.x - bh * (nsl + 1);
  const bool n_slab = slab == nsl;
  const int b = bh / NH, hd = bh - b * NH;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int sfl = cell_stage_floats(DH);
  const uint32_t sbase = (smem_u32(smem_raw) + 127u) & ~127u;
  float* stage0 = reinterpret_cast<float*>(smem_raw + (sbase - smem_u32(smem_raw)));
  float* red = stage0 + (size_t)kStages * sfl;                  // 2 x [8 warps][kTB][16]
  const uint32_t bar0 = smem_u32(s_bar);
  auto full_bar = [&](int s) { return bar0 + 8u * s; };
  auto empty_bar = [&](int s) { return bar0 + 8u * (kStages + s); };
  const int nbatch = (S + kTB - 1) / kTB;                       // S % kTB == 0 (host-checked)

  if (tid == 0) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), kConsumers / 32);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  pdl_wait();
  pdl_trigger();

  if (warp == kConsumers / 32) {
    if (lane == 0) {
      int s = 0;
      uint32_t ph = 0;
      for (int bt = 0; bt < nbatch; ++bt) {
        mbar_wait(empty_bar(s), ph ^ 1);
        float* st = stage0 + (size_t)s * sfl;
        const uint32_t bytes = (uint32_t)(kTB * DH * 8 + (n_slab ? 0 : kTB * kSlabW * 4) + 2 * kTB * 4);
        mbar_expect_tx(full_bar(s), bytes);
        const int64_t row0 = (int64_t)b * S + (int64_t)bt * kTB;
#pragma unroll
        for (int t = 0; t < kTB; ++t) {
          bulk_copy_g2s(smem_u32(st + t * DH * 2), p.qk + (((row0 + t) * NH + hd) * DH) * 2, (uint32_t)(DH * 8),
                        full_bar(s));
          if (!n_slab)
            bulk_copy_g2s(smem_u32(st + kTB * DH * 2 + t * kSlabW),
                          p.v + (row0 + t) * p.inner + hd * DH + slab * kSlabW, (uint32_t)(kSlabW * 4), full_bar(s));
        }
        bulk_copy_g2s(smem_u32(st + kTB * DH * 2 + kTB * kSlabW), p.fseq + (int64_t)bh * S + (int64_t)bt * kTB,
                      (uint32_t)(kTB * 4), full_bar(s));
        bulk_copy_g2s(smem_u32(st + kTB * DH * 2 + kTB * kSlabW + kTB), p.iseq + (int64_t)bh * S + (int64_t)bt * kTB,
                      (uint32_t)(kTB * 4), full_bar(s));
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
    return;
  }

  // ===== consumers =====
  const int cg = lane & 3;                      // column group: columns 4 cg .. 4 cg + 3 of the slab
  const int rl = (warp << 3) + (lane >> 2);     // row lane 0..63
  const int wc = (DH % 128 == 0) ? 128 : DH;    // slab width of the state layout in HBM
  const int CSl = DH / wc;
  const float kscale = rsqrtf((float)DH);
  float c[R][4];
  // load the slab (or n in column 0 of the n slab)
  if (!n_slab) {
    const int col0 = slab * kSlabW + 4 * cg;
    const int hs = col0 / wc, cin = col0 - hs * wc;
    const float* base = p.C + ((int64_t)(bh * CSl + hs) * DH) * wc + cin;
#pragma unroll
    for (int r = 0; r < R; ++r) {
      const float4 v4 = *reinterpret_cast<const float4*>(base + (int64_t)(rl + 64 * r) * wc);
      c[r][0] = v4.x; c[r][1] = v4.y; c[r][2] = v4.z; c[r][3] = v4.w;
    }
  } else {
#pragma unroll
    for (int r = 0; r < R; ++r) {
      c[r][0] = (cg == 0) ? p.n[(int64_t)bh * DH + rl + 64 * r] : 0.f;
      c[r][1] = c[r][2] = c[r][3] = 0.f;
    }
  }
  int s = 0;
  uint32_t ph = 0;
  for (int bt = 0; bt < nbatch; ++bt) {
    const float* st = stage0 + (size_t)s * sfl;
    const float* sqk = st;
    const float* sv = st + kTB * DH * 2;
    const float* sf = sv + kTB * kSlabW;
    const float* si = sf + kTB;
    mbar_wait(full_bar(s), ph);
    float acc[kTB][4];
#pragma unroll
    for (int t = 0; t < kTB; ++t) {
      const float f = sf[t];
      const float it = si[t] * kscale;
      float vi[4];
      if (!n_slab) {
        const float4 v4 = *reinterpret_cast<const float4*>(sv + t * kSlabW + 4 * cg);
        vi[0] = v4.x * it; vi[1] = v4.y * it; vi[2] = v4.z * it; vi[3] = v4.w * it;
      } else {
        vi[0] = (cg == 0) ? it : 0.f;
        vi[1] = vi[2] = vi[3] = 0.f;
      }
      acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        const float2 q2 = *reinterpret_cast<const float2*>(sqk + (t * DH + rl + 64 * r) * 2);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          c[r][j] = fmaf(f, c[r][j], q2.y * vi[j]);
          acc[t][j] = fmaf(q2.x, c[r][j], acc[t][j]);
        }
      }
    }
    __syncwarp();
    if (lane == 0) mbar_arrive(empty_bar(s));
    if (++s == kStages) { s = 0; ph ^= 1; }
    // reduce over the 8 row lanes of the warp (lane bits 2..4), then over the 8 warps through shared memory
#pragma unroll
    for (int t = 0; t < kTB; ++t)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        float v = acc[t][j];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 8);
        v += __shfl_xor_sync(0xffffffffu, v, 16);
        acc[t][j] = v;
      }
    float* rb = red + (bt & 1) * (8 * kTB * kSlabW);
    if (lane < 4) {
#pragma unroll
      for (int t = 0; t < kTB; ++t)
        *reinterpret_cast<float4*>(rb + (warp * kTB + t) * kSlabW + 4 * cg) =
            make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
    }
    asm volatile("bar.sync 1, %0;" ::"n"(kConsumers) : "memory");
    if (tid < kTB * kSlabW) {
      const int t = tid / kSlabW, col = tid - t * kSlabW;
      float sum = 0.f;
#pragma unroll
      for (int w = 0; w < 8; ++w) sum += rb[(w * kTB + t) * kSlabW + col];   // fixed order
      const int64_t row = (int64_t)b * S + (int64_t)bt * kTB + t;
      if (!n_slab) p.num[row * p.inner + hd * DH + slab * kSlabW + col] = sum;
      else if (col == 0) p.qn[row * NH + hd] = sum;
    }
  }
  // write the slab back
  if (!n_slab) {
    const int col0 = slab * kSlabW + 4 * cg;
    const int hs = col0 / wc, cin = col0 - hs * wc;
    float* base = p.C + ((int64_t)(bh * CSl + hs) * DH) * wc + cin;
#pragma unroll
    for (int r = 0; r < R; ++r)
      *reinterpret_cast<float4*>(base + (int64_t)(rl + 64 * r) * wc) = make_float4(c[r][0], c[r][1], c[r][2], c[r][3]);
  } else if (cg == 0) {
#pragma unroll
    for (int r = 0; r < R; ++r) p.n[(int64_t)bh * DH + rl + 64 * r] = c[r][0];
  }
}

// ------------------------------------------------------------------------------------------------
// Sequence finalize (token-parallel): h = num / (max(|q.n|, exp(-m)) + eps), MultiHeadLayerNorm over the head,
// learnable skip and output gate; emits the bf16 hi/lo planes of proj_down's A operand (or fp32).
// grid = (B*S rows) x NH, blockDim = DH rounded up to a warp multiple.
// ------------------------------------------------------------------------------------------------
struct FinalizeSeqParams {
  const float* num;       // [M, inner]
  const float* qn;        // [M, NH]
  const float* mseq;      // [B*NH, S]
  const float* outnorm_w; // [inner]
  const float* skip;      // [inner]
  const float* act;       // [M, inner]
  const float* u;         // [M, 2*inner]
  float* out;             // [M, inner] or nullptr
  __nv_bfloat16 *out_hi, *out_lo;
  int B, S, NH, DH, inner;
  float ln_eps, cell_eps;
};

__global__ void __launch_bounds__(1024) mlstm_finalize_seq_kernel(FinalizeSeqParams p) {
  __shared__ float red[32];
  const int row = blockIdx.x, hd = blockIdx.y;
  const int DH = p.DH, inner = p.inner;
  const int a = threadIdx.x;
  const bool ok = a < DH;
  const int ch = hd * DH + (ok ? a : 0);
  const float wn = ok ? p.outnorm_w[ch] : 0.f;
  const float wskip = ok ? p.skip[ch] : 0.f;
  pdl_wait();
  pdl_trigger();
  const int b = row / p.S, s = row - b * p.S;
  const float qn = p.qn[(int64_t)row * p.NH + hd];
  const float m = p.mseq[((int64_t)b * p.NH + hd) * p.S + s];
  const float num = ok ? p.num[(int64_t)row * inner + ch] : 0.f;
  const float act = ok ? p.act[(int64_t)row * inner + ch] : 0.f;
  const float zz = ok ? p.u[(int64_t)row * 2 * inner + inner + ch] : 0.f;
  const float den = fmaxf(fabsf(qn), expf(-m)) + p.cell_eps;
  const float h = num / den;
  const float mean = block_sum(h, red) / (float)DH;
  const float dlt = ok ? h - mean : 0.f;
  const float var = block_sum(dlt * dlt, red) / (float)DH;
  if (!ok) return;
  const float rstd = rsqrtf(var + p.ln_eps);
  float o = (h - mean) * rstd * (1.f + wn);
  o = (o + wskip * act) * silu_fast(zz);
  if (p.out) p.out[(int64_t)row * inner + ch] = o;
  if (p.out_hi) {
    const __nv_bfloat16 hi = __float2bfloat16_rn(o);
    p.out_hi[(int64_t)row * inner + ch] = hi;
    p.out_lo[(int64_t)row * inner + ch] = __float2bfloat16_rn(o - __bfloat162float(hi));
  }
}

}  // namespace pf

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
bool prefill_cell_supported(int DH) {
  const int R = DH / 64;
  return DH % 64 == 0 && (R == 1 || R == 2 || R == 4 || R == 6 || R == 8 || R == 10 || R == 12 || R == 16);
}

bool launch_conv_qkv_gates_seq(const ConvQkvParams& p, int S, cudaStream_t s) {
  const int nblk = p.inner / 4;
  const int per_chunk = (nblk + p.NCH - 1) / p.NCH;
  const int threads = ((per_chunk + 31) / 32) * 32;
  const int run = 32;                                   // tokens per CTA
  dim3 grid(p.NCH, p.B, (S + run - 1) / run);
  if (p.KS == 4 && p.NH == 4) {
    launch_k(pf::conv_qkv_gates_seq_kernel<4, 4>, grid, dim3(threads), 0, s, p, S, run);
  } else if (p.KS == 4 && p.NH == 8) {
    launch_k(pf::conv_qkv_gates_seq_kernel<4, 8>, grid, dim3(threads), 0, s, p, S, run);
  } else if (p.KS == 4 && p.NH == 2) {
    launch_k(pf::conv_qkv_gates_seq_kernel<4, 2>, grid, dim3(threads), 0, s, p, S, run);
  } else if (p.KS == 4 && p.NH == 1) {
    launch_k(pf::conv_qkv_gates_seq_kernel<4, 1>, grid, dim3(threads), 0, s, p, S, run);
  } else {
    return false;
  }
  const int64_t n4 = (int64_t)p.B * p.KS * (p.inner / 4);
  launch_k(pf::conv_state_seq_kernel, dim3((unsigned)((n4 + 255) / 256)), dim3(256), 0, s, p.u, p.conv_state, p.B, S,
           p.KS, p.inner);
  return true;
}

void launch_gate_scan_seq(const float* gate_part, const float* igate_b, const float* fgate_b, float* m_state,
                          float* fseq, float* iseq, float* mseq, int B, int S, int NH, int NCH, cudaStream_t s) {
  const int warps = B * NH;
  launch_k(pf::gate_scan_seq_kernel, dim3((warps + 3) / 4), dim3(128), 0, s, gate_part, igate_b, fgate_b, m_state, fseq,
           iseq, mseq, B, S, NH, NCH);
}

template <int R>
static cudaError_t launch_cell_R(const pf::CellSeqParams& p, cudaStream_t s) {
  const size_t smem = 128 + sizeof(float) * ((size_t)pf::kStages * pf::cell_stage_floats(p.DH) +
                                             2 * (size_t)8 * pf::kTB * pf::kSlabW);
  static size_t attr_smem = 0;
  if (smem > attr_smem) {
    cudaError_t e = cudaFuncSetAttribute(pf::mlstm_cell_seq_kernel<R>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    attr_smem = smem;
  }
  const int grid = p.B * p.NH * (p.DH / pf::kSlabW + 1);
  return launch_k(pf::mlstm_cell_seq_kernel<R>, dim3(grid), dim3(pf::kBlock), smem, s, p);
}

cudaError_t launch_cell_seq(float* C, float* n, const float* qk, const float* v, const float* fseq, const float* iseq,
                            float* num, float* qn, int B, int S, int NH, int DH, int inner, cudaStream_t s) {
  if (!prefill_cell_supported(DH) || S % pf::kTB != 0) return cudaErrorInvalidValue;
  pf::CellSeqParams p;
  p.C = C; p.n = n; p.qk = qk; p.v = v; p.fseq = fseq; p.iseq = iseq; p.num = num; p.qn = qn;
  p.B = B; p.S = S; p.NH = NH; p.DH = DH; p.inner = inner;
  switch (DH / 64) {
    case 1: return launch_cell_R<1>(p, s);
    case 2: return launch_cell_R<2>(p, s);
    case 4: return launch_cell_R<4>(p, s);
    case 6: return launch_cell_R<6>(p, s);
    case 8: return launch_cell_R<8>(p, s);
    case 10: return launch_cell_R<10>(p, s);
    case 12: return launch_cell_R<12>(p, s);
    case 16: return launch_cell_R<16>(p, s);
    default: return cudaErrorInvalidValue;
  }
}

cudaError_t launch_finalize_seq(const float* num, const float* qn, const float* mseq, const float* outnorm_w,
                                const float* skip, const float* act, const float* u, float* out, void* out_hi,
                                void* out_lo, int B, int S, int NH, int DH, int inner, float ln_eps, float cell_eps,
                                cudaStream_t s) {
  pf::FinalizeSeqParams p;
  p.num = num; p.qn = qn; p.mseq = mseq; p.outnorm_w = outnorm_w; p.skip = skip; p.act = act; p.u = u; p.out = out;
  p.out_hi = (__nv_bfloat16*)out_hi; p.out_lo = (__nv_bfloat16*)out_lo;
  p.B = B; p.S = S; p.NH = NH; p.DH = DH; p.inner = inner; p.ln_eps = ln_eps; p.cell_eps = cell_eps;
  const int thr = ((DH + 31) / 32) * 32;
  return launch_k(pf::mlstm_finalize_seq_kernel, dim3(B * S, NH), dim3(thr), 0, s, p);
}

}  // namespace xl
