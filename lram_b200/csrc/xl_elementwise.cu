// Small fused kernels around the GEMMs and the state step: LayerNorms, token embedding, the pre-cell
// conv / SiLU / headwise-qkv / gate pre-activation kernel, argmax + inv_tokenize, state reset.
// Reference arithmetic: SURVEY.md Appendix A (xlstm v1.0.x mLSTMLayer.step) and the LRAM files cited at
// each kernel.
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

// ------------------------------------------------------------------------------------------------
// LayerNorm rows: one warp per row, three passes over an L1-resident row (two-pass variance like ATen).
// xlstm LayerNorm: F.layer_norm(x, weight = 1 + w, bias = None, eps)  (residual_weight = 1)
// HF embed_ln:     nn.LayerNorm(d) affine                              (residual_weight = 0, bias)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ln_rows_kernel(const float* __restrict__ in, int64_t in_stride,
                                                      float* __restrict__ out, int64_t out_stride,
                                                      const float* __restrict__ w,
                                                      const float* __restrict__ bias, int residual_weight,
                                                      float eps, int rows, int d,
                                                      __nv_bfloat16* __restrict__ a_hi,
                                                      __nv_bfloat16* __restrict__ a_lo) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const float* x = in + (int64_t)warp * in_stride;
  const int d4 = d >> 2;
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float s = 0.f;
  for (int i = lane; i < d4; i += 32) {
    float4 v = x4[i];
    s += (v.x + v.y) + (v.z + v.w);
  }
  s = warp_sum(s);
  const float mean = s / (float)d;
  float q = 0.f;
  for (int i = lane; i < d4; i += 32) {
    float4 v = x4[i];
    float a = v.x - mean, b = v.y - mean, c = v.z - mean, e = v.w - mean;
    q += (a * a + b * b) + (c * c + e * e);
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)d + eps);
  float* o = out ? out + (int64_t)warp * out_stride : nullptr;
  const float4* w4 = reinterpret_cast<const float4*>(w);
  const float4* b4 = reinterpret_cast<const float4*>(bias);
  const float wofs = residual_weight ? 1.f : 0.f;
  for (int i = lane; i < d4; i += 32) {
    float4 v = x4[i];
    float4 g = w4[i];
    float4 r;
    r.x = (v.x - mean) * rstd * (g.x + wofs);
    r.y = (v.y - mean) * rstd * (g.y + wofs);
    r.z = (v.z - mean) * rstd * (g.z + wofs);
    r.w = (v.w - mean) * rstd * (g.w + wofs);
    if (bias) {
      float4 bb = b4[i];
      r.x += bb.x; r.y += bb.y; r.z += bb.z; r.w += bb.w;
    }
    if (o) reinterpret_cast<float4*>(o)[i] = r;
    if (a_hi) {
      // x = hi + lo with hi = bf16(x), lo = bf16(x - hi): two bf16 tensor-core passes recover ~16 bits
      float rr[4] = {r.x, r.y, r.z, r.w};
      __nv_bfloat16 hi[4], lo[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        hi[j] = __float2bfloat16_rn(rr[j]);
        lo[j] = __float2bfloat16_rn(rr[j] - __bfloat162float(hi[j]));
      }
      const int64_t base = (int64_t)warp * d + 4 * i;
      *reinterpret_cast<uint2*>(a_hi + base) = *reinterpret_cast<uint2*>(hi);
      *reinterpret_cast<uint2*>(a_lo + base) = *reinterpret_cast<uint2*>(lo);
    }
  }
}

void launch_ln_rows(const float* in, int64_t in_stride, float* out, int64_t out_stride, const float* w,
                    const float* bias, int residual_weight, float eps, int rows, int d, void* a_hi,
                    void* a_lo, cudaStream_t s) {
  if (rows <= 0) return;
  const int warps_per_block = 8;
  dim3 grid((rows + warps_per_block - 1) / warps_per_block);
  ln_rows_kernel<<<grid, warps_per_block * 32, 0, s>>>(in, in_stride, out, out_stride, w, bias,
                                                       residual_weight, eps, rows, d,
                                                       (__nv_bfloat16*)a_hi, (__nv_bfloat16*)a_lo);
}

// ------------------------------------------------------------------------------------------------
// Token embedding: embed_inputs + stack (s, rtg, r) + embed_ln
// (online_decision_transformer_model.py:522-530,588-612; discrete_decision_transformer_model.py:266-275)
// one warp per (b, tok) row.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) embed_tokens_kernel(
    const float* __restrict__ s_emb, const float* __restrict__ rtg, const float* __restrict__ rew,
    const float* __restrict__ w_ret, const float* __restrict__ b_ret, const float* __restrict__ w_rew,
    const float* __restrict__ b_rew, const float* __restrict__ ln_w, const float* __restrict__ ln_b,
    float eps, float* __restrict__ x, int B, int d) {
  const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (row >= 3 * B) return;
  const int b = row / 3, tok = row - 3 * b;
  const float scal = (tok == 1) ? rtg[b] : ((tok == 2 && rew) ? rew[b] : 0.f);
  const float* wv = (tok == 1) ? w_ret : w_rew;
  const float* bv = (tok == 1) ? b_ret : b_rew;
  const float* se = s_emb + (int64_t)b * d;
  auto val = [&](int i) -> float { return tok == 0 ? se[i] : fmaf(scal, wv[i], bv[i]); };
  float s = 0.f;
  for (int i = lane; i < d; i += 32) s += val(i);
  s = warp_sum(s);
  const float mean = s / (float)d;
  float q = 0.f;
  for (int i = lane; i < d; i += 32) {
    float a = val(i) - mean;
    q += a * a;
  }
  q = warp_sum(q);
  const float rstd = rsqrtf(q / (float)d + eps);
  float* o = x + (int64_t)row * d;
  for (int i = lane; i < d; i += 32) o[i] = (val(i) - mean) * rstd * ln_w[i] + ln_b[i];
}

void launch_embed_tokens(const float* s_emb, const float* rtg, const float* rew, const float* w_ret,
                         const float* b_ret, const float* w_rew, const float* b_rew, const float* ln_w,
                         const float* ln_b, float eps, float* x, int B, int d, cudaStream_t s) {
  const int rows = 3 * B;
  dim3 grid((rows + 7) / 8);
  embed_tokens_kernel<<<grid, 256, 0, s>>>(s_emb, rtg, rew, w_ret, b_ret, w_rew, b_rew, ln_w, ln_b, eps,
                                           x, B, d);
}

__global__ void pad_rows_kernel(const float* __restrict__ in, int K, float* __restrict__ out, int Kpad,
                                int rows) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * Kpad) return;
  const int r = (int)(i / Kpad), c = (int)(i - (int64_t)r * Kpad);
  out[i] = c < K ? in[(int64_t)r * K + c] : 0.f;
}
void launch_pad_rows(const float* in, int K, float* out, int Kpad, int rows, cudaStream_t s) {
  const int64_t n = (int64_t)rows * Kpad;
  pad_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, K, out, Kpad, rows);
}

__global__ void copy_rows_kernel(const float* __restrict__ in, int64_t in_stride, float* __restrict__ out,
                                 int64_t out_stride, int rows, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * d) return;
  const int r = (int)(i / d), c = (int)(i - (int64_t)r * d);
  out[(int64_t)r * out_stride + c] = in[(int64_t)r * in_stride + c];
}
void launch_copy_rows(const float* in, int64_t in_stride, float* out, int64_t out_stride, int rows, int d,
                      cudaStream_t s) {
  const int64_t n = (int64_t)rows * d;
  copy_rows_kernel<<<(unsigned)((n + 255) / 256), 256, 0, s>>>(in, in_stride, out, out_stride, rows, d);
}

// ------------------------------------------------------------------------------------------------
// Pre-cell kernel: CausalConv1d.step (ring window) + bias -> SiLU -> block-diagonal q/k/v
// (LinearHeadwiseExpand, 4x4 blocks) -> partial igate/fgate pre-activations.
// [ext-xlstm] mLSTMLayer.step / conv1d_step / mLSTMCell.step gate Linear(cat[q,k,v]) — SURVEY App. A.
// grid = (NCH channel chunks, B envs); one thread owns one 4-channel block for all T tokens, so the conv
// window lives in registers across the tokens of the step. Gate partial sums per chunk are written out
// and summed in fixed order by the state kernel (deterministic, no float atomics).
// ------------------------------------------------------------------------------------------------
constexpr int kMaxNH = 8;
constexpr int kMaxT = 4;

template <int KS>
__global__ void __launch_bounds__(256) conv_qkv_gates_kernel(ConvQkvParams p) {
  constexpr int kMaxKS = KS;
  __shared__ float red[32];
  const int b = blockIdx.y;
  const int chunk = blockIdx.x;
  const int inner = p.inner, NH = p.NH, T = p.T;
  const int nblk = inner >> 2;                       // 4-channel blocks
  const int blk_per_chunk = (nblk + p.NCH - 1) / p.NCH;
  const int blk0 = chunk * blk_per_chunk;
  const int blk1 = min(nblk, blk0 + blk_per_chunk);

  float gi[kMaxT][kMaxNH], gf[kMaxT][kMaxNH];
#pragma unroll
  for (int t = 0; t < kMaxT; ++t)
#pragma unroll
    for (int h = 0; h < kMaxNH; ++h) gi[t][h] = gf[t][h] = 0.f;

  for (int j = blk0 + threadIdx.x; j < blk1; j += blockDim.x) {
    const int c = 4 * j;
    // conv window: rows 1..KS-1 of the state are the KS-1 most recent inputs (oldest first)
    float win[kMaxKS][4];
    float* cs = p.conv_state + (int64_t)b * KS * inner + c;
#pragma unroll
    for (int r = 0; r < kMaxKS; ++r) {
      if (r < KS) {
        float4 v = *reinterpret_cast<const float4*>(cs + (int64_t)r * inner);
        win[r][0] = v.x; win[r][1] = v.y; win[r][2] = v.z; win[r][3] = v.w;
      }
    }
    float cw[4][kMaxKS];
#pragma unroll
    for (int ch = 0; ch < 4; ++ch)
#pragma unroll
      for (int r = 0; r < kMaxKS; ++r)
        if (r < KS) cw[ch][r] = p.conv_w[(int64_t)(c + ch) * KS + r];
    const float4 cb = *reinterpret_cast<const float4*>(p.conv_b + c);
    float wq[16], wk[16], wv[16];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float4 a = reinterpret_cast<const float4*>(p.wq + (int64_t)j * 16)[i];
      float4 k = reinterpret_cast<const float4*>(p.wk + (int64_t)j * 16)[i];
      float4 v = reinterpret_cast<const float4*>(p.wv + (int64_t)j * 16)[i];
      wq[4 * i] = a.x; wq[4 * i + 1] = a.y; wq[4 * i + 2] = a.z; wq[4 * i + 3] = a.w;
      wk[4 * i] = k.x; wk[4 * i + 1] = k.y; wk[4 * i + 2] = k.z; wk[4 * i + 3] = k.w;
      wv[4 * i] = v.x; wv[4 * i + 1] = v.y; wv[4 * i + 2] = v.z; wv[4 * i + 3] = v.w;
    }
#pragma unroll
    for (int t = 0; t < kMaxT; ++t) {
      if (t < T) {
        const int64_t row = (int64_t)b * T + t;
        const float4 xm4 = *reinterpret_cast<const float4*>(p.u + row * 2 * inner + c);
        const float xm[4] = {xm4.x, xm4.y, xm4.z, xm4.w};
        // roll(-1); state[-1] = x
#pragma unroll
        for (int r = 0; r < kMaxKS - 1; ++r)
          if (r < KS - 1) {
#pragma unroll
            for (int ch = 0; ch < 4; ++ch) win[r][ch] = win[r + 1][ch];
          }
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) win[KS - 1][ch] = xm[ch];
        float a[4];
        const float cbv[4] = {cb.x, cb.y, cb.z, cb.w};
#pragma unroll
        for (int ch = 0; ch < 4; ++ch) {
          float acc = 0.f;
#pragma unroll
          for (int r = 0; r < kMaxKS; ++r)
            if (r < KS) acc = fmaf(win[r][ch], cw[ch][r], acc);
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
        // channel c of head h sits at (row*NH + h)*DH + (c - h*DH) == row*inner + c: (q,k) pairs interleaved
        float* qk = p.qk + (row * inner + c) * 2;
        *reinterpret_cast<float4*>(qk) = make_float4(q[0], k[0], q[1], k[1]);
        *reinterpret_cast<float4*>(qk + 4) = make_float4(q[2], k[2], q[3], k[3]);
        *reinterpret_cast<float4*>(p.v + row * inner + c) = make_float4(v[0], v[1], v[2], v[3]);
        *reinterpret_cast<float4*>(p.act + row * inner + c) = make_float4(a[0], a[1], a[2], a[3]);
#pragma unroll
        for (int h = 0; h < kMaxNH; ++h) {
          if (h < NH) {
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
            gi[t][h] += si;
            gf[t][h] += sf;
          }
        }
      }
    }
    // write the window back: rows = last KS inputs, oldest first (reference conv_state layout)
#pragma unroll
    for (int r = 0; r < kMaxKS; ++r)
      if (r < KS)
        *reinterpret_cast<float4*>(cs + (int64_t)r * inner) =
            make_float4(win[r][0], win[r][1], win[r][2], win[r][3]);
  }
  // block-reduce the gate partials of this chunk
#pragma unroll
  for (int t = 0; t < kMaxT; ++t) {
    if (t < T) {
#pragma unroll
      for (int h = 0; h < kMaxNH; ++h) {
        if (h < NH) {
          const float si = block_sum(gi[t][h], red);
          const float sf = block_sum(gf[t][h], red);
          if (threadIdx.x == 0) {
            float* gp = p.gate_part + (((int64_t)b * T + t) * p.NCH + chunk) * 2 * NH;
            gp[h] = si;
            gp[NH + h] = sf;
          }
        }
      }
    }
  }
}

void launch_conv_qkv_gates(const ConvQkvParams& p, cudaStream_t s) {
  const int nblk = p.inner / 4;
  const int per_chunk = (nblk + p.NCH - 1) / p.NCH;
  int threads = ((per_chunk + 31) / 32) * 32;
  if (threads > 256) threads = 256;
  if (threads < 32) threads = 32;
  dim3 grid(p.NCH, p.B);
  switch (p.KS) {
    case 1: conv_qkv_gates_kernel<1><<<grid, threads, 0, s>>>(p); break;
    case 2: conv_qkv_gates_kernel<2><<<grid, threads, 0, s>>>(p); break;
    case 3: conv_qkv_gates_kernel<3><<<grid, threads, 0, s>>>(p); break;
    case 4: conv_qkv_gates_kernel<4><<<grid, threads, 0, s>>>(p); break;
    case 5: conv_qkv_gates_kernel<5><<<grid, threads, 0, s>>>(p); break;
    case 6: conv_qkv_gates_kernel<6><<<grid, threads, 0, s>>>(p); break;
    case 7: conv_qkv_gates_kernel<7><<<grid, threads, 0, s>>>(p); break;
    case 8: conv_qkv_gates_kernel<8><<<grid, threads, 0, s>>>(p); break;
    default: break;
  }
}

// ------------------------------------------------------------------------------------------------
// argmax over action logits + MinMaxTokenizer.inv_tokenize
// (multi_domain_discrete_dt_model.py:83-94; src/tokenizers_custom/minmax_tokenizer.py:31-47).
// torch.argmax semantics: first index of the maximum; NaN counts as the maximum.
// one warp per (b, action dim).
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) argmax_tokens_kernel(const float* __restrict__ logits,
                                                            int64_t row_pitch, int B, int act_dim,
                                                            int num_actions, int discrete_actions,
                                                            int discrete, float bin_width, float min_val,
                                                            int32_t* __restrict__ tokens,
                                                            float* __restrict__ actions) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  const int nrows = discrete ? B : B * act_dim;
  if (w >= nrows) return;
  const int b = discrete ? w : w / act_dim;
  const int j = discrete ? 0 : w - b * act_dim;
  const float* lg = logits + (int64_t)b * row_pitch + (int64_t)j * num_actions;
  const int n = discrete ? discrete_actions : num_actions;
  float best = -INFINITY;
  int bi = 0x7fffffff;
  bool best_nan = false;
  for (int i = lane; i < n; i += 32) {
    const float v = lg[i];
    const bool vn = (v != v);
    bool take;
    if (best_nan) take = false;                       // earlier NaN in this lane wins (smaller index)
    else if (vn) take = true;
    else take = (bi == 0x7fffffff) || (v > best);
    if (take) { best = v; bi = i; best_nan = vn; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    const int on = __shfl_xor_sync(0xffffffffu, (int)best_nan, o);
    bool take;
    if (oi == 0x7fffffff) take = false;
    else if (bi == 0x7fffffff) take = true;
    else if (best_nan && on) take = oi < bi;
    else if (on) take = true;
    else if (best_nan) take = false;
    else take = (ov > best) || (ov == best && oi < bi);
    if (take) { best = ov; bi = oi; best_nan = (on != 0); }
  }
  if (lane == 0) {
    if (discrete) {
      tokens[(int64_t)b * act_dim] = bi;
      actions[(int64_t)b * act_dim] = (float)bi;
    } else {
      tokens[(int64_t)b * act_dim + j] = bi;
      int t = bi - discrete_actions;
      if (t < 0) t = 0;
      actions[(int64_t)b * act_dim + j] = (float)t * bin_width + min_val;
    }
  }
}

void launch_argmax_tokens(const float* logits, int64_t row_pitch, int B, int act_dim, int num_actions,
                          int discrete_actions, int discrete, float bin_width, float min_val,
                          int32_t* tokens, float* actions, cudaStream_t s) {
  const int rows = discrete ? B : B * act_dim;
  dim3 grid((rows + 7) / 8);
  argmax_tokens_kernel<<<grid, 256, 0, s>>>(logits, row_pitch, B, act_dim, num_actions, discrete_actions,
                                            discrete, bin_width, min_val, tokens, actions);
}

// ------------------------------------------------------------------------------------------------
// per-env state reset (past_key_values = None for the masked envs; evaluation.py:124,251,261)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) state_reset_kernel(float* C, float* n, float* m, float* conv,
                                                          const uint8_t* __restrict__ mask, int B,
                                                          int64_t c_per_env, int64_t n_per_env,
                                                          int64_t m_per_env, int64_t conv_per_env) {
  const int b = blockIdx.y;
  if (mask && mask[b] == 0) return;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t t0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float4* c4 = reinterpret_cast<float4*>(C + (int64_t)b * c_per_env);
  const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
  for (int64_t i = t0; i < c_per_env / 4; i += stride) c4[i] = z;
  for (int64_t i = t0; i < n_per_env; i += stride) n[(int64_t)b * n_per_env + i] = 0.f;
  for (int64_t i = t0; i < m_per_env; i += stride) m[(int64_t)b * m_per_env + i] = 0.f;
  for (int64_t i = t0; i < conv_per_env; i += stride) conv[(int64_t)b * conv_per_env + i] = 0.f;
}

void launch_state_reset(float* C, float* n, float* m, float* conv, const uint8_t* mask, int B,
                        int64_t c_per_env, int64_t n_per_env, int64_t m_per_env, int64_t conv_per_env,
                        cudaStream_t s) {
  dim3 grid(32, B);
  state_reset_kernel<<<grid, 256, 0, s>>>(C, n, m, conv, mask, B, c_per_env, n_per_env, m_per_env,
                                          conv_per_env);
}

}  // namespace xl
