// Small-batch (latency) path of the recurrent step: the whole mLSTM block stack as ONE persistent kernel.
//
// The reference evaluates ONE environment at a time (src/callbacks/evaluation.py:80 asserts num_envs == 1): per env
// step it runs T = 3 tokens through L blocks of [ext-xlstm] xLSTMBlockStack.step (decision_xlstm.py:161-165). At
// such batch sizes nothing is bandwidth- or tensor-bound: the multi-kernel path (xl_api.cu run_blocks) spends the
// step in ~6 dependent launches per block (205 us for 16M x 1 env against ~7 us of HBM time). This kernel keeps
// one CTA resident per SM for the whole stack and replaces launch boundaries by grid barriers, four per block:
//
//   A  x_n = LN(x) (every CTA, redundantly) ; u = x_n W_up^T as warp GEMVs over 4-column groups (bf16 weights
//      streamed once, fp32 FMA, activations exact) ; a warp that owns an x_m group runs conv + SiLU + headwise
//      q/k/v for those 4 channels right away and accumulates its share of the gate pre-activations      | barrier
//   C  gates: sum of the per-CTA partials (fixed order), m/f/i recurrence ; C <- f C + (i k/sqrt(DH)) v^T and
//      partial numerators q^T C over (strip, row chunk) units spread over all SMs                        | barrier
//   D1 per (env, head): n update, q.n, h = num / (max(|q.n|, e^-m) + eps), MultiHeadLayerNorm, skip, SiLU(z) gate
//                                                                                                      | barrier
//   D2 x += g W_down^T (warp GEMVs)                                                                    | barrier
//
// and ends with post_blocks_norm. Arithmetic per element is the multi-kernel path's (same formulas, fp32), only
// the GEMV summation order differs; results agree with the oracle within the same 1e-3 bound (tests).
// Rows are ordered [env][token] (row = b*T + t), M = B*T <= 16. Requirements: d % 256 == 0, DH % 128 == 0 (slab-
// major C), KS == 4, NH <= 8, mLSTM blocks only — the four shipped presets qualify; anything else stays on the
// multi-kernel path.
#include <cuda.h>
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"
#include "xl_gemv.cuh"

#ifndef XL_LL_KBA4
#define XL_LL_KBA4 2
#endif
#ifndef XL_LL_KBD4
#define XL_LL_KBD4 4
#endif

namespace xl {

namespace ll {

using namespace gv;

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kGateSlots = 16;      // 2 * NH <= 16
constexpr int kSmallFloats = 512;   // per-warp staging area of phase A (small weights + conv window)
constexpr unsigned kSpinLimit = 1u << 22;

__device__ __forceinline__ unsigned ld_acquire(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Grid barrier, self re-arming across launches: bar[0] = arrivals, bar[1] = generation, bar[2] = abort flag.
// Every CTA is resident (cooperative launch, one CTA per SM). A CTA that spins for seconds raises the abort
// flag, after which every barrier falls through: a bug ends as wrong numbers + a flag, never as a hung GPU.
__device__ __forceinline__ void grid_barrier(unsigned* bar, unsigned& gen, int G) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned arrived = atomicAdd(bar, 1u);
    if (arrived == (unsigned)G - 1u) {
      bar[0] = 0u;
      __threadfence();
      atomicAdd(bar + 1, 1u);
    } else {
      unsigned spins = 0;
      while (ld_acquire(bar + 1) == gen) {
        if (++spins > kSpinLimit) {
          atomicExch(bar + 2, 1u);
          break;
        }
        if ((spins & 255u) == 0u && ld_acquire(bar + 2)) break;
      }
    }
  }
  gen += 1u;
  __syncthreads();
}

// per-phase clock stamps of CTA 0 (measurement aid, p.dbg nullable): [layer][9]
#define XL_LL_STAMP(k) \
  do { if (p.dbg && blockIdx.x == 0 && tid == 0) p.dbg[layer * 9 + (k)] = clock64(); } while (0)

// sum over the CTA of N values per thread; result in every thread. red: N*32 floats of shared memory.
template <int N>
__device__ __forceinline__ void cta_sum_n(float (&v)[N], float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) v[k] = warp_sum(v[k]);
  __syncthreads();
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[k * 32 + wid] = v[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    const float r = (lane < kWarps) ? red[k * 32 + lane] : 0.f;
    v[k] = warp_sum(r);
  }
}

template <int MR, int T>
__global__ void __launch_bounds__(kThreads, 1) lowlat_stack_kernel(const LowLatParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ unsigned s_gen;
  constexpr int CGA = (MR == 16) ? 2 : 4;    // columns per GEMV pass in phase A (a group is always 4 columns)
  constexpr int KBA = (MR == 8) ? 2 : ((MR == 4) ? XL_LL_KBA4 : 4);
  constexpr int CGD = (MR == 4) ? 1 : 2;     // columns per warp item in phase D2
  constexpr int KBD = (MR == 4) ? XL_LL_KBD4 : 4;

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int G = gridDim.x;
  const int gw = blockIdx.x * kWarps + warp, GW = G * kWarps;
  const int M = p.M, B = p.B, d = p.d, inner = p.inner, NH = p.NH, DH = p.DH;
  const int NSL = DH >> 7;
  float* xs = smem;                          // [M][d] (phase A) / [M][inner] (phase D2), permuted rows
  float* scr = smem + (size_t)MR * inner;    // phase scratch

  if (tid == 0) s_gen = ld_acquire(p.bar + 1);
  __syncthreads();
  unsigned gen = s_gen;

  for (int layer = 0; layer < p.L; ++layer) {
    const LowLatLayer* lw = p.layers + layer;
    char* sbase = p.state + (size_t)layer * p.layer_bytes;
    float* Cst = reinterpret_cast<float*>(sbase + p.c_off);
    float* nst = reinterpret_cast<float*>(sbase + p.n_off);
    float* mst = reinterpret_cast<float*>(sbase + p.m_off);
    float* cvst = reinterpret_cast<float*>(sbase + p.conv_off);

    XL_LL_STAMP(0);
    // =========================== phase A: LN, proj_up, conv / q k v / gate partials =====================
    {
      if (warp < M) {                         // warp w normalises row w (xlstm LayerNorm: gamma = 1 + w, no bias)
        const float* xr = p.x + (size_t)warp * d;
        float* xo = xs + warp * d;
        const float* nw = lw->norm_w;
        float sum = 0.f;
        for (int k = lane * 4; k < d; k += 128) {
          const float4 v = __ldcg(reinterpret_cast<const float4*>(xr + k));
          *reinterpret_cast<float4*>(xo + perm4(k)) = v;
          sum += (v.x + v.y) + (v.z + v.w);
        }
        const float mean = warp_sum(sum) / (float)d;
        float q = 0.f;
        for (int k = lane * 4; k < d; k += 128) {
          const float4 v = *reinterpret_cast<const float4*>(xo + perm4(k));
          const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
          q += (a * a + b * b) + (c * c + e * e);
        }
        const float rstd = rsqrtf(warp_sum(q) / (float)d + p.ln_eps);
        for (int k = lane * 4; k < d; k += 128) {
          float4 v = *reinterpret_cast<const float4*>(xo + perm4(k));
          const float4 g = __ldg(reinterpret_cast<const float4*>(nw + k));
          v.x = (v.x - mean) * rstd * (g.x + 1.f);
          v.y = (v.y - mean) * rstd * (g.y + 1.f);
          v.z = (v.z - mean) * rstd * (g.z + 1.f);
          v.w = (v.w - mean) * rstd * (g.w + 1.f);
          *reinterpret_cast<float4*>(xo + perm4(k)) = v;
        }
      }
      float* wsm = scr + warp * kSmallFloats;
      float* wsc = scr + kWarps * kSmallFloats + warp * (MR * 4);
      float* gs_all = scr + kWarps * kSmallFloats + kWarps * MR * 4;
      float* wgs = gs_all + warp * (MR * kGateSlots);
      for (int i = lane; i < MR * kGateSlots; i += 32) wgs[i] = 0.f;
      __syncthreads();

      const int nblk = inner >> 2;
      const __nv_bfloat16* w_up = lw->w_up;
      const int gofs = 68, cofs = 68 + 24 * NH;
      for (int item = gw; item < 2 * nblk; item += GW) {
        const bool is_xm = item < nblk;
        const int col0 = item * 4;
        if (is_xm) {
          // stage what the conv / headwise / gate epilogue of these 4 channels needs (overlaps the GEMV)
          const int c = col0;
          const int nchunks = 17 + 6 * NH + 3 * B;
          for (int ci = lane; ci < nchunks; ci += 32) {
            const float* src;
            if (ci < 4) src = lw->conv_w + (size_t)c * 4 + ci * 4;
            else if (ci == 4) src = lw->conv_b + c;
            else if (ci < 9) src = lw->wq + (size_t)item * 16 + (ci - 5) * 4;
            else if (ci < 13) src = lw->wk + (size_t)item * 16 + (ci - 9) * 4;
            else if (ci < 17) src = lw->wv + (size_t)item * 16 + (ci - 13) * 4;
            else if (ci < 17 + 6 * NH) {
              const int r = ci - 17, hh = r / 6, part = r - hh * 6;
              const float* gwt = (part >= 3) ? lw->wf : lw->wi;
              src = gwt + (size_t)hh * 3 * inner + (size_t)(part % 3) * inner + c;
            } else {
              const int r = ci - 17 - 6 * NH, bb = r / 3, row = 1 + (r - bb * 3);
              src = cvst + ((size_t)bb * 4 + row) * inner + c;
            }
            cp_async16(wsm + ci * 4, src);
          }
        }
#pragma unroll
        for (int half = 0; half < 4 / CGA; ++half) {
          float acc[CGA][MR];
          gemv_cols<CGA, MR, KBA>(w_up, d, col0 + half * CGA, xs, d, M, acc, lane);
#pragma unroll
          for (int c = 0; c < CGA; ++c)
#pragma unroll
            for (int m = 0; m < MR; ++m) {
              const float s = warp_sum(acc[c][m]);
              if (lane == ((m * 4 + half * CGA + c) & 31)) wsc[m * 4 + half * CGA + c] = s;
            }
        }
        if (is_xm) cp_async_wait_all();
        __syncwarp();
        if (!is_xm) {
          // z half of u: rows of 4 consecutive floats
          for (int idx = lane; idx < M * 4; idx += 32)
            p.z[(size_t)(idx >> 2) * inner + (col0 - inner) + (idx & 3)] = wsc[idx];
        } else {
          // conv + SiLU + headwise q/k/v + gate partials: lane = (row m, output channel o) of this 4-channel block.
          // Token t's conv window is the last 4 of [old rows 1..3, x_0 .. x_t]: no dependency between tokens.
          const int c = col0;
          const float* cw = wsm;          // [ch][r]
          const float* cb = wsm + 16;
          const float* wq = wsm + 20;
          const float* wk = wsm + 36;
          const float* wv = wsm + 52;
          for (int base = 0; base < M * 4; base += 32) {
            const int idx = base + lane;
            const bool on = idx < M * 4;
            const int m = on ? (idx >> 2) : 0, o = idx & 3;
            const int b = m / T, t = m - b * T;
            float a[4], xm[4];
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) {
              float acc = 0.f;
#pragma unroll
              for (int r = 0; r < 4; ++r) {
                const int j = t + r;
                const float wv_ = (j < 3) ? wsm[cofs + (b * 3 + j) * 4 + ch] : wsc[(b * T + j - 3) * 4 + ch];
                acc = fmaf(wv_, cw[ch * 4 + r], acc);
              }
              a[ch] = silu(acc + cb[ch]);
              xm[ch] = wsc[m * 4 + ch];
            }
            float sq = 0.f, sk = 0.f, sv = 0.f;
#pragma unroll
            for (int dd = 0; dd < 4; ++dd) {
              sq = fmaf(a[dd], wq[4 * o + dd], sq);
              sk = fmaf(a[dd], wk[4 * o + dd], sk);
              sv = fmaf(xm[dd], wv[4 * o + dd], sv);
            }
            if (on) {
              const size_t o1 = (size_t)m * inner + c + o;
              p.q[o1] = sq; p.k[o1] = sk; p.v[o1] = sv;
              p.act[o1] = (o == 0) ? a[0] : ((o == 1) ? a[1] : ((o == 2) ? a[2] : a[3]));
            }
            for (int hh = 0; hh < NH; ++hh) {
              const float* g6 = wsm + gofs + hh * 24;
              float si = on ? (sq * g6[o] + sk * g6[4 + o] + sv * g6[8 + o]) : 0.f;
              float sf = on ? (sq * g6[12 + o] + sk * g6[16 + o] + sv * g6[20 + o]) : 0.f;
              si += __shfl_xor_sync(0xffffffffu, si, 1);
              sf += __shfl_xor_sync(0xffffffffu, sf, 1);
              si += __shfl_xor_sync(0xffffffffu, si, 2);
              sf += __shfl_xor_sync(0xffffffffu, sf, 2);
              if (on && o == 0) {
                wgs[m * kGateSlots + hh] += si;
                wgs[m * kGateSlots + NH + hh] += sf;
              }
            }
          }
          // conv window after the T tokens: last 4 of [old rows 1..3, x_0 .. x_{T-1}], oldest first
          for (int idx = lane; idx < B * 4; idx += 32) {
            const int b = idx >> 2, r = idx & 3, j = T - 1 + r;
            const float* src = (j < 3) ? wsm + cofs + (b * 3 + j) * 4 : wsc + (b * T + j - 3) * 4;
            *reinterpret_cast<float4*>(cvst + ((size_t)b * 4 + r) * inner + c) = make_float4(src[0], src[1], src[2], src[3]);
          }
        }
        __syncwarp();
      }
      __syncthreads();
      for (int idx = tid; idx < M * 2 * NH; idx += kThreads) {
        const int m = idx / (2 * NH), g = idx - m * 2 * NH;
        float s = 0.f;
        for (int w = 0; w < kWarps; ++w) s += gs_all[(w * MR + m) * kGateSlots + g];
        p.gate_part[((size_t)blockIdx.x * MR + m) * kGateSlots + g] = s;
      }
    }
    XL_LL_STAMP(1);
    grid_barrier(p.bar, gen, G);
    XL_LL_STAMP(2);

    // =========================== phase C: gates, matrix-memory update, partial numerators ===============
    {
      float* s_pre = scr;                 // [M][16]
      float* s_f = scr + 256;             // [B*NH][T]
      float* s_i = s_f + 512;
      float* s_m = s_i + 512;             // [B*NH][T+1]
      float* sq = s_m + 640;              // [T][DH]
      float* sk = sq + T * DH;
      float* sacc = sk + T * DH;          // [kWarps][T][128]
      for (int idx = warp; idx < M * 2 * NH; idx += kWarps) {
        const int m = idx / (2 * NH), g = idx - m * 2 * NH;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const int c = lane + 32 * j;
          v[j] = (c < G) ? __ldcg(p.gate_part + ((size_t)c * MR + m) * kGateSlots + g) : 0.f;
        }
        float s = ((v[0] + v[1]) + (v[2] + v[3])) + ((v[4] + v[5]) + (v[6] + v[7]));
        s = warp_sum(s);
        if (lane == 0) s_pre[m * kGateSlots + g] = s + ((g < NH) ? __ldg(lw->bi + g) : __ldg(lw->bf + g - NH));
      }
      __syncthreads();
      if (tid < B * NH) {
        const int bh = tid, b = bh / NH, hd = bh - b * NH;
        float mprev = mst[bh];
        s_m[bh * (T + 1)] = mprev;
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float ig = s_pre[(b * T + t) * kGateSlots + hd], fg = s_pre[(b * T + t) * kGateSlots + NH + hd];
          const float lf = log_sigmoid(fg);
          const float mnew = fmaxf(lf + mprev, ig);
          s_f[bh * T + t] = expf(lf + mprev - mnew);
          s_i[bh * T + t] = expf(ig - mnew);
          s_m[bh * (T + 1) + t + 1] = mnew;
          mprev = mnew;
        }
        if (blockIdx.x == 0) {
          float* gg = p.gates + (size_t)bh * (3 * T + 1);
#pragma unroll
          for (int t = 0; t < T; ++t) {
            gg[t] = s_f[bh * T + t];
            gg[T + t] = s_i[bh * T + t];
            gg[2 * T + 1 + t] = s_m[bh * (T + 1) + t + 1];
          }
          gg[2 * T] = s_m[bh * (T + 1)];
        }
      }
      __syncthreads();
      const float kscale = rsqrtf((float)DH);
      const int nunits = B * NH * NSL * p.RS;
      for (int unit = blockIdx.x; unit < nunits; unit += G) {
        const int strip = unit / p.RS, rs = unit - strip * p.RS;
        const int bh = strip / NSL, slab = strip - bh * NSL;
        const int b = bh / NH, hd = bh - b * NH;
        const int r0 = rs * p.rpu;
        const int nrows = min(p.rpu, DH - r0);
        for (int idx = tid; idx < T * nrows; idx += kThreads) {
          const int t = idx / nrows, r = idx - t * nrows;
          const size_t off = (size_t)(b * T + t) * inner + hd * DH + r0 + r;
          sq[t * DH + r] = __ldcg(p.q + off);
          sk[t * DH + r] = __ldcg(p.k + off);
        }
        float f[T], vi[T][4], acc[T][4];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          f[t] = s_f[bh * T + t];
          const float sc = s_i[bh * T + t] * kscale;
          const float4 v4 = __ldcg(reinterpret_cast<const float4*>(p.v + (size_t)(b * T + t) * inner + hd * DH +
                                                                   slab * 128 + lane * 4));
          vi[t][0] = v4.x * sc; vi[t][1] = v4.y * sc; vi[t][2] = v4.z * sc; vi[t][3] = v4.w * sc;
#pragma unroll
          for (int j = 0; j < 4; ++j) acc[t][j] = 0.f;
        }
        __syncthreads();
        float* Cb = Cst + ((size_t)(bh * NSL + slab) * DH + r0) * 128 + lane * 4;
        for (int rb = warp; rb < nrows; rb += kWarps * 8) {
          float4 c4[8];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = rb + kWarps * j;
            if (r < nrows) c4[j] = __ldcg(reinterpret_cast<const float4*>(Cb + (size_t)r * 128));
          }
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = rb + kWarps * j;
            if (r < nrows) {
              float c[4] = {c4[j].x, c4[j].y, c4[j].z, c4[j].w};
#pragma unroll
              for (int t = 0; t < T; ++t) {
                const float qr = sq[t * DH + r], kr = sk[t * DH + r];
#pragma unroll
                for (int jj = 0; jj < 4; ++jj) {
                  c[jj] = fmaf(f[t], c[jj], kr * vi[t][jj]);
                  acc[t][jj] = fmaf(qr, c[jj], acc[t][jj]);
                }
              }
              __stcg(reinterpret_cast<float4*>(Cb + (size_t)r * 128), make_float4(c[0], c[1], c[2], c[3]));
            }
          }
        }
#pragma unroll
        for (int t = 0; t < T; ++t)
          *reinterpret_cast<float4*>(sacc + (warp * T + t) * 128 + lane * 4) =
              make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
        __syncthreads();
        for (int idx = tid; idx < T * 128; idx += kThreads) {
          const int t = idx >> 7, col = idx & 127;
          float s = 0.f;
          for (int w = 0; w < kWarps; ++w) s += sacc[(w * T + t) * 128 + col];
          p.partial[((size_t)(strip * p.RS + rs) * T + t) * 128 + col] = s;
        }
        __syncthreads();
      }
    }
    XL_LL_STAMP(3);
    grid_barrier(p.bar, gen, G);
    XL_LL_STAMP(4);

    // =========================== phase D1: n, q.n, h, MultiHeadLayerNorm, skip, output gate =============
    {
      float* s_red = scr;                  // [T][32]
      for (int bh = blockIdx.x; bh < B * NH; bh += G) {
        const int b = bh / NH, hd = bh - b * NH;
        const float* gg = p.gates + (size_t)bh * (3 * T + 1);
        float gf[T], gi[T], gm[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          gf[t] = __ldcg(gg + t);
          gi[t] = __ldcg(gg + T + t);
          gm[t] = __ldcg(gg + 2 * T + 1 + t);
        }
        const float kscale = rsqrtf((float)DH);
        float nreg[2], wn[2], wsk[2], qv[2][T], kv[2][T], num[2][T], av[2][T], zv[2][T];
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = tid + e * kThreads;
          const bool ok = a < DH;
          const int ch = hd * DH + (ok ? a : 0);
          nreg[e] = ok ? nst[(size_t)bh * DH + a] : 0.f;
          wn[e] = ok ? __ldg(lw->outnorm + ch) : 0.f;
          wsk[e] = ok ? __ldg(lw->skip + ch) : 0.f;
          const int strip = bh * NSL + ((ok ? a : 0) >> 7), col = a & 127;
#pragma unroll
          for (int t = 0; t < T; ++t) {
            const size_t off = (size_t)(b * T + t) * inner + ch;
            qv[e][t] = ok ? __ldcg(p.q + off) : 0.f;
            kv[e][t] = ok ? __ldcg(p.k + off) : 0.f;
            av[e][t] = ok ? __ldcg(p.act + off) : 0.f;
            zv[e][t] = ok ? __ldcg(p.z + off) : 0.f;
            float s = 0.f;
            if (ok)
#pragma unroll 8
              for (int rs = 0; rs < p.RS; ++rs)
                s += __ldcg(p.partial + ((size_t)(strip * p.RS + rs) * T + t) * 128 + col);
            num[e][t] = s;
          }
        }
        float qn[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          qn[t] = 0.f;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            nreg[e] = fmaf(gf[t], nreg[e], gi[t] * kscale * kv[e][t]);
            qn[t] += qv[e][t] * nreg[e];
          }
        }
        cta_sum_n<T>(qn, s_red);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = tid + e * kThreads;
          if (a < DH) nst[(size_t)bh * DH + a] = nreg[e];
        }
        float mean[T], var[T];
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float den = fmaxf(fabsf(qn[t]), expf(-gm[t])) + p.cell_eps;
          num[0][t] = num[0][t] / den;
          num[1][t] = num[1][t] / den;
          mean[t] = num[0][t] + num[1][t];
        }
        cta_sum_n<T>(mean, s_red);
#pragma unroll
        for (int t = 0; t < T; ++t) {
          mean[t] /= (float)DH;
          var[t] = 0.f;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            const float dl = (tid + e * kThreads < DH) ? num[e][t] - mean[t] : 0.f;
            var[t] += dl * dl;
          }
        }
        cta_sum_n<T>(var, s_red);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const int a = tid + e * kThreads;
          if (a < DH) {
            const int ch = hd * DH + a;
#pragma unroll
            for (int t = 0; t < T; ++t) {
              const float rstd = rsqrtf(var[t] / (float)DH + p.ln_eps);
              float o = (num[e][t] - mean[t]) * rstd * (1.f + wn[e]);
              o = (o + wsk[e] * av[e][t]) * silu_fast(zv[e][t]);
              p.g[(size_t)(b * T + t) * inner + ch] = o;
            }
          }
        }
        if (tid == 0) mst[bh] = gm[T - 1];
        __syncthreads();
      }
    }
    XL_LL_STAMP(5);
    grid_barrier(p.bar, gen, G);
    XL_LL_STAMP(6);

    // =========================== phase D2: x += g W_down^T ==============================================
    {
      const int nq = inner >> 2;
      for (int idx = tid; idx < M * nq; idx += kThreads) {
        const int m = idx / nq, k = (idx - m * nq) * 4;
        *reinterpret_cast<float4*>(xs + m * inner + perm4(k)) =
            __ldcg(reinterpret_cast<const float4*>(p.g + (size_t)m * inner + k));
      }
      __syncthreads();
      float* wsc = scr + warp * (MR * 4);
      const __nv_bfloat16* w_down = lw->w_down;
      for (int item = gw; item < d / CGD; item += GW) {
        const int col0 = item * CGD;
        float acc[CGD][MR];
        gemv_cols<CGD, MR, KBD>(w_down, inner, col0, xs, inner, M, acc, lane);
#pragma unroll
        for (int c = 0; c < CGD; ++c)
#pragma unroll
          for (int m = 0; m < MR; ++m) {
            const float s = warp_sum(acc[c][m]);
            if (lane == ((m * CGD + c) & 31)) wsc[m * CGD + c] = s;
          }
        __syncwarp();
        for (int idx = lane; idx < M * CGD; idx += 32) {
          float* xp = p.x + (size_t)(idx / CGD) * d + col0 + (idx % CGD);
          *xp = __ldcg(xp) + wsc[idx];
        }
        __syncwarp();
      }
    }
    XL_LL_STAMP(7);
    grid_barrier(p.bar, gen, G);
    XL_LL_STAMP(8);
  }

  // =========================== post_blocks_norm ==========================================================
  if ((int)blockIdx.x < M) {
    float* s_red = scr;
    const int row = blockIdx.x;
    const bool ok = tid < (d >> 2);
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f), g = v;
    if (ok) {
      v = __ldcg(reinterpret_cast<const float4*>(p.x + (size_t)row * d) + tid);
      g = __ldg(reinterpret_cast<const float4*>(p.post_w) + tid);
    }
    float s1[1] = {(v.x + v.y) + (v.z + v.w)};
    cta_sum_n<1>(s1, s_red);
    const float mean = s1[0] / (float)d;
    const float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
    float s2[1] = {ok ? (a * a + b * b) + (c * c + e * e) : 0.f};
    cta_sum_n<1>(s2, s_red);
    const float rstd = rsqrtf(s2[0] / (float)d + p.ln_eps);
    if (ok)
      reinterpret_cast<float4*>(p.out + (size_t)row * p.out_stride)[tid] =
          make_float4(a * rstd * (g.x + 1.f), b * rstd * (g.y + 1.f), c * rstd * (g.z + 1.f), e * rstd * (g.w + 1.f));
  }
}

size_t scratch_floats(int MR, int T, int DH) {
  const size_t a = (size_t)kWarps * kSmallFloats + (size_t)kWarps * MR * 4 + (size_t)kWarps * MR * kGateSlots;
  const size_t c = 256 + 512 + 512 + 640 + (size_t)2 * T * DH + (size_t)kWarps * T * 128;
  return a > c ? a : c;
}

template <int MR, int T>
cudaError_t launch(const LowLatParams& p, size_t smem, int coop, cudaStream_t s) {
  if (cudaError_t e = ensure_dyn_smem<&lowlat_stack_kernel<MR, T>>(smem); e != cudaSuccess) return e;
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(p.G);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = coop ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, lowlat_stack_kernel<MR, T>, p);
}

}  // namespace ll

int lowlat_row_bucket(int M) { return M <= 4 ? 4 : (M <= 8 ? 8 : (M <= 16 ? 16 : 0)); }

bool lowlat_supported(int B, int T, int d, int inner, int NH, int DH, int KS, size_t smem_limit, size_t* smem_bytes) {
  const int M = B * T;
  const int MR = lowlat_row_bucket(M);
  if (!MR || B < 1 || B > 16 || T < 1 || T > 4) return false;
  if (d % 256 || inner % 256 || DH % 128 || DH > 2 * ll::kThreads || KS != 4 || NH < 1 || NH > 8) return false;
  if (inner != NH * DH || d > 4 * ll::kThreads || 17 + 6 * NH + 3 * B > ll::kSmallFloats / 4) return false;
  const size_t smem = sizeof(float) * ((size_t)MR * inner + ll::scratch_floats(MR, T, DH));
  if (smem_bytes) *smem_bytes = smem;
  return smem <= smem_limit;
}

// rows per (strip, row chunk) unit of phase C: fewest rounds x rows over G CTAs, then fewest partials
void lowlat_plan_state(int B, int NH, int DH, int G, int* rpu_out, int* rs_out) {
  const int nstrips = B * NH * (DH / 128);
  int best = 16;
  double best_cost = 1e30;
  for (int rpu = 16; rpu <= DH; rpu += 16) {
    const int RS = (DH + rpu - 1) / rpu;
    const int rounds = (nstrips * RS + G - 1) / G;
    const double cost = (double)rounds * (rpu + 8) + 0.5 * RS;
    if (cost < best_cost - 1e-9) {
      best_cost = cost;
      best = rpu;
    }
  }
  *rpu_out = best;
  *rs_out = (DH + best - 1) / best;
}

cudaError_t launch_lowlat_stack(const LowLatParams& p, size_t smem, int coop, cudaStream_t s) {
  const int MR = lowlat_row_bucket(p.M);
#define XL_LL_CASE(MRV, TV) \
  if (MR == MRV && p.T == TV) return ll::launch<MRV, TV>(p, smem, coop, s);
#ifdef XL_LL_ONLY_43
  XL_LL_CASE(4, 3)
#else
  XL_LL_CASE(4, 1) XL_LL_CASE(4, 2) XL_LL_CASE(4, 3) XL_LL_CASE(4, 4)
  XL_LL_CASE(8, 1) XL_LL_CASE(8, 2) XL_LL_CASE(8, 3) XL_LL_CASE(8, 4)
  XL_LL_CASE(16, 1) XL_LL_CASE(16, 2) XL_LL_CASE(16, 3) XL_LL_CASE(16, 4)
#endif
#undef XL_LL_CASE
  return cudaErrorInvalidValue;
}

}  // namespace xl
