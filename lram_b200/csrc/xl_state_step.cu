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
// One launch touches every C element exactly once (one 128-bit load, T fused multiply-adds + T dot-product
// FMAs in registers, one 128-bit store): algorithmic bytes = 8*NH*DH^2 per env (+ n, m) regardless of T.
//
// Work decomposition: CTA = (env b, head h, row chunk rs, column slab cs). A thread owns 4 consecutive
// columns (dv) and walks the rows (dk) of its chunk with kUnroll independent 128-bit loads in flight; the
// q^T C reduction over rows is therefore thread-local in the streaming loop and only crosses warps once
// at the end (shared memory), then crosses CTAs through a small partial buffer. The last CTA of a head to
// finish (atomic ticket) sums the partials in fixed order (deterministic), updates n and m, divides, applies
// the multi-head GroupNorm and the output gate.
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

constexpr int kThreads = 256;
constexpr int kUnroll = 8;
constexpr int kMaxNCH = 16;

template <int T>
__global__ void __launch_bounds__(kThreads, 3) mlstm_state_step_kernel(StateStepParams p) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float s_f[T], s_i[T], s_m[T + 1];
  __shared__ float s_red[32];
  __shared__ int s_last;

  const int DH = p.DH, NH = p.NH, inner = p.inner;
  const int CS = DH / p.cols_per_cta;
  const int RS = p.rows_split;
  const int tiles_per_head = CS * RS;
  const int bh = blockIdx.x / tiles_per_head;
  const int tile = blockIdx.x - bh * tiles_per_head;
  const int rs = tile / CS, cs = tile - rs * CS;
  const int b = bh / NH, hd = bh - b * NH;
  const int rows_per = (DH + RS - 1) / RS;
  const int r0 = rs * rows_per;
  const int nrows = max(0, min(DH, r0 + rows_per) - r0);
  const int c0 = cs * p.cols_per_cta;
  const int TX = p.cols_per_cta >> 2;      // threads along columns
  const int TY = kThreads / TX;            // row lanes
  const int tx = threadIdx.x % TX, ty = threadIdx.x / TX;
  const int tid = threadIdx.x;

  float* sq = smem;                        // [T][rows_per]
  float* sk = smem + T * rows_per;         // [T][rows_per]  (k/sqrt(DH)) * i_t
  float* sacc = smem + 2 * T * rows_per;   // [TY][T][cols_per_cta]

  // ---- prologue: gates (one thread, sequential in t), q/k chunk -> smem --------------------------
  if (tid == 0) {
    float mprev = p.m[bh];
    s_m[0] = mprev;
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float* gp = p.gate_part + ((int64_t)b * T + t) * p.NCH * 2 * NH;
      float ig = 0.f, fg = 0.f;
      for (int c = 0; c < p.NCH; ++c) {
        ig += gp[c * 2 * NH + hd];
        fg += gp[c * 2 * NH + NH + hd];
      }
      if (p.igate_b) ig += p.igate_b[hd];
      if (p.fgate_b) fg += p.fgate_b[hd];
      const float lf = log_sigmoid(fg);
      const float mnew = fmaxf(lf + mprev, ig);
      s_f[t] = expf(lf + mprev - mnew);
      s_i[t] = expf(ig - mnew);
      s_m[t + 1] = mnew;
      mprev = mnew;
    }
  }
  const float kscale = rsqrtf((float)DH);
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float* qrow = p.qkv + ((int64_t)b * T + t) * 3 * inner + hd * DH + r0;
    for (int r = tid; r < nrows; r += kThreads) {
      sq[t * rows_per + r] = qrow[r];
      sk[t * rows_per + r] = qrow[inner + r] * kscale;
    }
  }
  float v[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float4 vv = *reinterpret_cast<const float4*>(p.qkv + ((int64_t)b * T + t) * 3 * inner +
                                                       2 * inner + hd * DH + c0 + 4 * tx);
    v[t][0] = vv.x; v[t][1] = vv.y; v[t][2] = vv.z; v[t][3] = vv.w;
  }
  __syncthreads();
  float f[T];
#pragma unroll
  for (int t = 0; t < T; ++t) {
    f[t] = s_f[t];
    const float it = s_i[t];
    for (int r = tid; r < nrows; r += kThreads) sk[t * rows_per + r] *= it;
  }
  __syncthreads();

  // ---- streaming loop over the rows of this chunk -------------------------------------------------
  float acc[T][4];
#pragma unroll
  for (int t = 0; t < T; ++t) acc[t][0] = acc[t][1] = acc[t][2] = acc[t][3] = 0.f;

  float* Cbase = p.C + ((int64_t)bh * DH + r0) * DH + c0 + 4 * tx;
  for (int rbase = ty; rbase < nrows; rbase += TY * kUnroll) {
    float4 cv[kUnroll];
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int r = rbase + u * TY;
      if (r < nrows) cv[u] = ld_stream(reinterpret_cast<const float4*>(Cbase + (int64_t)r * DH));
    }
#pragma unroll
    for (int u = 0; u < kUnroll; ++u) {
      const int r = rbase + u * TY;
      if (r < nrows) {
        float c[4] = {cv[u].x, cv[u].y, cv[u].z, cv[u].w};
#pragma unroll
        for (int t = 0; t < T; ++t) {
          const float qr = sq[t * rows_per + r];
          const float kr = sk[t * rows_per + r];
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            c[j] = fmaf(f[t], c[j], kr * v[t][j]);
            acc[t][j] = fmaf(qr, c[j], acc[t][j]);
          }
        }
        st_stream(reinterpret_cast<float4*>(Cbase + (int64_t)r * DH), make_float4(c[0], c[1], c[2], c[3]));
      }
    }
  }

  // ---- reduce the per-thread partial numerators over the TY row lanes ------------------------------
#pragma unroll
  for (int t = 0; t < T; ++t)
    *reinterpret_cast<float4*>(sacc + ((ty * T + t) * TX + tx) * 4) =
        make_float4(acc[t][0], acc[t][1], acc[t][2], acc[t][3]);
  __syncthreads();
  const int W = p.cols_per_cta;
  for (int idx = tid; idx < T * W; idx += kThreads) {
    const int t = idx / W, c = idx - t * W;
    float s = 0.f;
    for (int y = 0; y < TY; ++y) s += sacc[(y * T + t) * W + c];
    p.partial[(((int64_t)bh * RS + rs) * T + t) * DH + c0 + c] = s;
  }

  // ---- ticket: the last CTA of this (env, head) finalises ------------------------------------------
  __threadfence();
  __syncthreads();
  if (tid == 0) {
    const unsigned int prev = atomicAdd(p.counters + bh, 1u);
    s_last = (prev == (unsigned)(tiles_per_head - 1));
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();

  // n update and q.n per token (all DH rows; q/k re-read from global, L2 resident)
  float qn[T];
  {
    constexpr int kMaxRowsPerThread = 4;     // DH <= 4*256
    float nreg[kMaxRowsPerThread];
#pragma unroll
    for (int i = 0; i < kMaxRowsPerThread; ++i) {
      const int a = tid + i * kThreads;
      nreg[i] = (a < DH) ? p.n[(int64_t)bh * DH + a] : 0.f;
    }
#pragma unroll
    for (int t = 0; t < T; ++t) {
      const float* qrow = p.qkv + ((int64_t)b * T + t) * 3 * inner + hd * DH;
      const float ft = s_f[t], it = s_i[t];
      float part = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxRowsPerThread; ++i) {
        const int a = tid + i * kThreads;
        if (a < DH) {
          nreg[i] = fmaf(ft, nreg[i], it * (qrow[inner + a] * kscale));
          part = fmaf(qrow[a], nreg[i], part);
        }
      }
      qn[t] = block_sum(part, s_red);
    }
#pragma unroll
    for (int i = 0; i < kMaxRowsPerThread; ++i) {
      const int a = tid + i * kThreads;
      if (a < DH) p.n[(int64_t)bh * DH + a] = nreg[i];
    }
  }
  // h = num / den, GroupNorm over the head, output gate
  constexpr int kMaxColsPerThread = 4;
#pragma unroll
  for (int t = 0; t < T; ++t) {
    const float den = fmaxf(fabsf(qn[t]), expf(-s_m[t + 1])) + p.cell_eps;
    float hreg[kMaxColsPerThread];
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxColsPerThread; ++i) {
      const int c = tid + i * kThreads;
      hreg[i] = 0.f;
      if (c < DH) {
        float s = 0.f;
        for (int r = 0; r < RS; ++r) s += ld_cg(p.partial + (((int64_t)bh * RS + r) * T + t) * DH + c);
        hreg[i] = s / den;
        part += hreg[i];
      }
    }
    const float mean = block_sum(part, s_red) / (float)DH;
    float vpart = 0.f;
#pragma unroll
    for (int i = 0; i < kMaxColsPerThread; ++i) {
      const int c = tid + i * kThreads;
      if (c < DH) {
        const float dlt = hreg[i] - mean;
        vpart = fmaf(dlt, dlt, vpart);
      }
    }
    const float rstd = rsqrtf(block_sum(vpart, s_red) / (float)DH + p.ln_eps);
    const int64_t row = (int64_t)b * T + t;
#pragma unroll
    for (int i = 0; i < kMaxColsPerThread; ++i) {
      const int c = tid + i * kThreads;
      if (c < DH) {
        const int ch = hd * DH + c;
        float o = (hreg[i] - mean) * rstd * (1.f + p.outnorm_w[ch]);
        if (p.h_raw) p.h_raw[row * inner + ch] = hreg[i];
        if (p.skip) {
          const float a = p.act[row * inner + ch];
          const float z = p.u[row * 2 * inner + inner + ch];
          o = (o + p.skip[ch] * a) * silu(z);
        }
        p.out[row * inner + ch] = o;
        if (p.out_hi) {
          const __nv_bfloat16 hi = __float2bfloat16_rn(o);
          reinterpret_cast<__nv_bfloat16*>(p.out_hi)[row * inner + ch] = hi;
          reinterpret_cast<__nv_bfloat16*>(p.out_lo)[row * inner + ch] =
              __float2bfloat16_rn(o - __bfloat162float(hi));
        }
      }
    }
  }
  if (tid == 0) {
    p.m[bh] = s_m[T];
    p.counters[bh] = 0u;   // ready for the next launch
  }
}

void state_step_auto_tiling(int B, int NH, int DH, int num_sms, int* rows_split, int* cols_per_cta) {
  // column slab: widest of {128, 64, 32, 16, 8, 4} that divides DH (a warp then reads 512 contiguous B)
  int cols = 128;
  while (cols > 4 && (DH % cols) != 0) cols >>= 1;
  if (cols > DH) cols = DH;
  const int CS = DH / cols;
  // split rows until there are enough CTAs to fill the machine a few times over
  const int64_t target = (int64_t)num_sms * 3 * 2;
  int rs = 1;
  const int TY = kThreads / (cols / 4);
  while ((int64_t)B * NH * CS * rs < target && (DH / (rs * 2)) >= TY * 2 && rs < 32) rs *= 2;
  *rows_split = rs;
  *cols_per_cta = cols;
}

template <int T>
static cudaError_t launch_T(const StateStepParams& p, cudaStream_t s) {
  const int rows_per = (p.DH + p.rows_split - 1) / p.rows_split;
  const int TX = p.cols_per_cta / 4, TY = kThreads / TX;
  const size_t smem = sizeof(float) * ((size_t)2 * T * rows_per + (size_t)TY * T * p.cols_per_cta);
  if (smem > 48 * 1024) {
    cudaError_t e = cudaFuncSetAttribute(mlstm_state_step_kernel<T>,
                                         cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
  }
  const int CS = p.DH / p.cols_per_cta;
  const int64_t grid = (int64_t)p.B * p.NH * CS * p.rows_split;
  mlstm_state_step_kernel<T><<<(unsigned)grid, kThreads, smem, s>>>(p);
  return cudaGetLastError();
}

cudaError_t launch_state_step(StateStepParams p, int num_sms, cudaStream_t s) {
  if (p.rows_split <= 0 || p.cols_per_cta <= 0) {
    int rs, cols;
    state_step_auto_tiling(p.B, p.NH, p.DH, num_sms, &rs, &cols);
    if (p.rows_split <= 0) p.rows_split = rs;
    if (p.cols_per_cta <= 0) p.cols_per_cta = cols;
  }
  if (p.cols_per_cta % 4 || p.DH % p.cols_per_cta || p.cols_per_cta > 1024 ||
      kThreads % (p.cols_per_cta / 4) || p.DH > 4 * kThreads || p.NCH > kMaxNCH || p.rows_split > p.DH)
    return cudaErrorInvalidValue;
  switch (p.T) {
    case 1: return launch_T<1>(p, s);
    case 2: return launch_T<2>(p, s);
    case 3: return launch_T<3>(p, s);
    case 4: return launch_T<4>(p, s);
    default: return cudaErrorInvalidValue;
  }
}

}  // namespace xl
