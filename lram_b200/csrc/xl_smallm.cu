// Small-batch front half of an mLSTM block as ONE kernel (M = B*T <= 16 rows):
//   x_n = LN(x) ; u = x_n W_up^T ; conv + SiLU ; headwise q / k / v ; partial igate / fgate pre-activations
// i.e. what ln_rows_cta_kernel + the tcgen05 proj_up + conv_qkv_gates_kernel do as three launches
// ([ext-xlstm] xLSTMBlock.step -> LayerNorm, mLSTMLayer.step: proj_up, conv1d.step, q/k/v_proj, gate Linear — reached
// from src/algos/models/decision_xlstm.py:163). The reference evaluates ONE env at a time (evaluation.py:80); at such M a
// 128-row tensor-core tile is >= 87 % padding and the step is bound by the number of dependent kernels (a kernel
// boundary inside the CUDA graph costs ~1.3 us on B200, profiles/r01_lowlat_persistent.md), so the three kernels become
// one GEMV-style kernel: each warp owns one 4-column group of u, streams its bf16 weight rows once (fp32 FMA, the
// activations stay fp32 — no hi/lo planes needed), and a warp that owns an x_m group finishes conv / q / k / v / gate
// partials for those 4 channels in its epilogue. Outputs are exactly what the state-stream and finalize kernels read:
// (q,k) pairs [M, inner, 2], v [M, inner], a [M, inner], z in u[:, inner:], the conv window, and gate partials
// [M, NCH, 2 NH], one chunk per cluster of 4 or 8 CTAs (shares added in rank order over distributed shared memory).
#include <cooperative_groups.h>
#include <cuda.h>
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_gemv.cuh"
#include "xl_internal.h"

namespace xl {

namespace sm {

using namespace gv;
namespace cg = cooperative_groups;

constexpr int kMaxChunks = 16;     // gate-partial chunks the state-stream / finalize kernels add up (xl_state_step.cu)
constexpr int kWarps = 8;
constexpr int kThreads = kWarps * 32;
constexpr int kSmallFloats = 512;   // per-warp staging: conv taps, headwise blocks, gate columns, conv window

template <int MR, int T>
__global__ void __launch_bounds__(kThreads) smallm_pre_kernel(const SmallPreParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CGA = (MR == 16) ? 2 : 4;
  constexpr int KBA = (MR == 8) ? 2 : 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.B * T, B = p.B, d = p.d, inner = p.inner, NH = p.NH;
  float* xs = smem;                                   // [M][d], permuted rows (xl_gemv.cuh)
  float* wsm = smem + (size_t)MR * d + warp * kSmallFloats;
  float* wsc = smem + (size_t)MR * d + kWarps * kSmallFloats + warp * (MR * 4);
  float* gs_all = smem + (size_t)MR * d + kWarps * kSmallFloats + kWarps * MR * 4;   // [kWarps][MR][2 NH]
  float* wgs = gs_all + warp * (MR * 2 * NH);

  const int nblk = inner >> 2;
  const int item = blockIdx.x * kWarps + warp;        // 4-column group of u: [0, nblk) = x_m, [nblk, 2 nblk) = z
  const bool live = item < 2 * nblk;
  const bool is_xm = live && item < nblk;
  const int col0 = item * 4;
  const int gofs = 68, cofs = 68 + 24 * NH;

  // ---- prologue: everything no kernel of this step writes (weights; the conv window, which only this kernel
  //      touches) is staged before the dependency wait, and the warp's weight rows are pulled towards L2 ----------
  if (is_xm) {
    const int c = col0;
    const int nchunks = 17 + 6 * NH + 3 * B;
    for (int ci = lane; ci < nchunks; ci += 32) {
      const float* src;
      if (ci < 4) src = p.conv_w + (size_t)c * 4 + ci * 4;
      else if (ci == 4) src = p.conv_b + c;
      else if (ci < 9) src = p.wq + (size_t)item * 16 + (ci - 5) * 4;
      else if (ci < 13) src = p.wk + (size_t)item * 16 + (ci - 9) * 4;
      else if (ci < 17) src = p.wv + (size_t)item * 16 + (ci - 13) * 4;
      else if (ci < 17 + 6 * NH) {
        const int r = ci - 17, hh = r / 6, part = r - hh * 6;
        const float* gwt = (part >= 3) ? p.wf : p.wi;
        src = gwt + (size_t)hh * 3 * inner + (size_t)(part % 3) * inner + c;
      } else {
        const int r = ci - 17 - 6 * NH, bb = r / 3, row = 1 + (r - bb * 3);
        src = p.conv_state + ((size_t)bb * 4 + row) * inner + c;
      }
      cp_async16(wsm + ci * 4, src);
    }
  }
  if (live && lane < 4) {
    const char* wrow = reinterpret_cast<const char*>(p.w_up + (size_t)(col0 + lane) * d);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wrow), "r"((unsigned)(d * 2)) : "memory");
  }
  for (int i = lane; i < MR * 2 * NH; i += 32) wgs[i] = 0.f;
  // the first KBA k-blocks of this warp's first CGA weight rows go to registers before the dependency wait
  uint4 w0[KBA][CGA];
  if (live) gemv_load<CGA, KBA>(p.w_up, d, col0, 0, w0, lane);
  pdl_wait();
  pdl_trigger();

  // ---- LayerNorm of the M rows (every CTA, redundantly: M*d floats from L2), gamma = 1 + w -------------------
  for (int row = warp; row < M; row += kWarps) {
    const float* xr = p.x + (size_t)row * d;
    float* xo = xs + row * d;
    float sum = 0.f;
    for (int k0 = lane * 4; k0 < d; k0 += 4 * 128) {       // 4 loads in flight (not load -> store -> load); same sum order
      float4 v[4];
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k0 + u * 128 < d) v[u] = *reinterpret_cast<const float4*>(xr + k0 + u * 128);
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (k0 + u * 128 < d) {
          *reinterpret_cast<float4*>(xo + perm4(k0 + u * 128)) = v[u];
          sum += (v[u].x + v[u].y) + (v[u].z + v[u].w);
        }
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
      const float4 g = __ldg(reinterpret_cast<const float4*>(p.norm_w + k));
      v.x = (v.x - mean) * rstd * (g.x + 1.f);
      v.y = (v.y - mean) * rstd * (g.y + 1.f);
      v.z = (v.z - mean) * rstd * (g.z + 1.f);
      v.w = (v.w - mean) * rstd * (g.w + 1.f);
      *reinterpret_cast<float4*>(xo + perm4(k)) = v;
    }
  }
  __syncthreads();

  if (live) {
#pragma unroll
    for (int half = 0; half < 4 / CGA; ++half) {
      float acc[CGA][MR];
      if (half == 0) gemv_cols_pre<CGA, MR, KBA>(p.w_up, d, col0, xs, d, M, acc, lane, w0);
      else gemv_cols<CGA, MR, KBA>(p.w_up, d, col0 + half * CGA, xs, d, M, acc, lane);
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
      // z half of u (the finalize kernel reads u[:, inner:])
      for (int idx = lane; idx < M * 4; idx += 32)
        p.u[(size_t)(idx >> 2) * 2 * inner + col0 + (idx & 3)] = wsc[idx];
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
            const float w_ = (j < 3) ? wsm[cofs + (b * 3 + j) * 4 + ch] : wsc[(b * T + j - 3) * 4 + ch];
            acc = fmaf(w_, cw[ch * 4 + r], acc);
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
          *reinterpret_cast<float2*>(p.qk + o1 * 2) = make_float2(sq, sk);       // (q,k) pairs interleaved
          p.v[o1] = sv;
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
            wgs[m * 2 * NH + hh] += si;
            wgs[m * 2 * NH + NH + hh] += sf;
          }
        }
      }
      // conv window after the T tokens: last 4 of [old rows 1..3, x_0 .. x_{T-1}], oldest first
      for (int idx = lane; idx < B * 4; idx += 32) {
        const int b = idx >> 2, r = idx & 3, j = T - 1 + r;
        const float* src = (j < 3) ? wsm + cofs + (b * 3 + j) * 4 : wsc + (b * T + j - 3) * 4;
        *reinterpret_cast<float4*>(p.conv_state + ((size_t)b * 4 + r) * inner + c) =
            make_float4(src[0], src[1], src[2], src[3]);
      }
    }
  }
  __syncthreads();
  // Gate partials: the CTAs run in clusters of CL = 4 or 8; each CTA leaves its share in its own shared memory and the
  // cluster's rank-0 CTA adds the shares in rank order over distributed shared memory (deterministic) into chunk
  // blockIdx.x / CL of [M, NCH, 2 NH] — NCH <= 16 chunks for the state-stream / finalize kernels to add up.
  cg::cluster_group cluster = cg::this_cluster();
  const int CL = (int)cluster.num_blocks();
  float* share = gs_all;                      // [M][2 NH], reuses the start of the per-warp slots after the sum below
  const bool xm_cta = (int)blockIdx.x < p.NCH * CL;
  float mine = 0.f;
  if (xm_cta && tid < M * 2 * NH) {
    const int m = tid / (2 * NH), g = tid - m * 2 * NH;
    for (int w = 0; w < kWarps; ++w) mine += gs_all[(w * MR + m) * 2 * NH + g];
  }
  __syncthreads();
  if (xm_cta && tid < M * 2 * NH) share[tid] = mine;
  cluster.sync();
  if (xm_cta && cluster.block_rank() == 0 && tid < M * 2 * NH) {
    const int m = tid / (2 * NH), g = tid - m * 2 * NH;
    float s = 0.f;
    for (int r = 0; r < CL; ++r) s += cluster.map_shared_rank(share, r)[tid];
    p.gate_part[((size_t)m * p.NCH + blockIdx.x / CL) * 2 * NH + g] = s;
  }
  cluster.sync();                             // remote shared memory stays valid until rank 0 has read it
}

size_t smem_floats(int MR, int d, int NH) {
  return (size_t)MR * d + (size_t)kWarps * kSmallFloats + (size_t)kWarps * MR * 4 + (size_t)kWarps * MR * 2 * NH;
}

template <int MR, int T>
cudaError_t launch(const SmallPreParams& p, size_t smem, cudaStream_t s) {
  if (cudaError_t e = ensure_dyn_smem<&smallm_pre_kernel<MR, T>>(smem); e != cudaSuccess) return e;
  const int nblk = p.inner >> 2;
  const int grid = (2 * nblk + kWarps - 1) / kWarps;          // multiple of the cluster size (host-checked)
  const int cl = (nblk / kWarps) / p.NCH;                      // CTAs per cluster: 4 or 8
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cl;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = g_use_pdl ? 2 : 1;
  return cudaLaunchKernelEx(&cfg, smallm_pre_kernel<MR, T>, p);
}

// ---- back half: x += g W_down^T for M <= 16 rows ------------------------------------------------------------
// g = (h~ + skip*a) * silu(z) [M, inner] fp32 from the finalize kernel; one warp per CGD output columns streams its
// weight rows once (K = inner), every CTA holds g in shared memory. Replaces the split-K tcgen05 proj_down + the
// plane-reducing LayerNorm of the next block at these sizes: the residual stream x is complete when the kernel ends.
template <int MR>
__global__ void __launch_bounds__(kThreads) smallm_down_kernel(const SmallDownParams p) {
  extern __shared__ __align__(16) float smem[];
  constexpr int CGD = (MR == 4) ? 1 : 2;
  constexpr int KBD = (MR == 4) ? 4 : 4;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int M = p.M, d = p.d, inner = p.inner;
  float* xs = smem;                                            // [M][inner], permuted rows
  float* wsc = smem + (size_t)MR * inner + warp * (MR * 2);
  const int item = blockIdx.x * kWarps + warp;
  const bool live = item * CGD < d;
  const int col0 = item * CGD;
  if (live && lane < CGD) {
    const char* wrow = reinterpret_cast<const char*>(p.w_down + (size_t)(col0 + lane) * inner);
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(wrow), "r"((unsigned)(inner * 2)) : "memory");
  }
  // the first KBD k-blocks of this warp's weight rows go to registers before the dependency wait
  uint4 w0[KBD][CGD];
  if (live) gemv_load<CGD, KBD>(p.w_down, inner, col0, 0, w0, lane);
  pdl_wait();
  pdl_trigger();
  const int nq = inner >> 2;
  for (int i0 = tid; i0 < M * nq; i0 += 4 * kThreads) {       // 4 loads in flight per thread (not load -> store -> load)
    float4 v[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = i0 + u * kThreads;
      if (idx < M * nq) {
        const int m = idx / nq, k = (idx - m * nq) * 4;
        v[u] = *reinterpret_cast<const float4*>(p.g + (size_t)m * inner + k);
      }
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int idx = i0 + u * kThreads;
      if (idx < M * nq) {
        const int m = idx / nq, k = (idx - m * nq) * 4;
        *reinterpret_cast<float4*>(xs + m * inner + perm4(k)) = v[u];
      }
    }
  }
  __syncthreads();
  if (!live) return;
  float acc[CGD][MR];
  gemv_cols_pre<CGD, MR, KBD>(p.w_down, inner, col0, xs, inner, M, acc, lane, w0);
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
    *xp = *xp + wsc[idx];
  }
}

template <int MR>
cudaError_t launch_down(const SmallDownParams& p, cudaStream_t s) {
  constexpr int CGD = (MR == 4) ? 1 : 2;
  const size_t smem = sizeof(float) * ((size_t)MR * p.inner + (size_t)kWarps * MR * 2);
  if (cudaError_t e = ensure_dyn_smem<&smallm_down_kernel<MR>>(smem); e != cudaSuccess) return e;
  const int items = p.d / CGD;
  return launch_k(smallm_down_kernel<MR>, dim3((items + kWarps - 1) / kWarps), dim3(kThreads), smem, s, p);
}

}  // namespace sm

int smallm_row_bucket(int M) { return M <= 4 ? 4 : (M <= 8 ? 8 : (M <= 16 ? 16 : 0)); }

// chunks of gate partials the kernel writes (= CTAs that own x_m groups); 0 when the shape is not supported
int smallm_pre_chunks(int B, int T, int d, int inner, int NH, int KS) {
  const int M = B * T;
  if (!smallm_row_bucket(M) || B < 1 || B > 16 || T < 1 || T > 4) return 0;
  if (d % 256 || inner % 256 || KS != 4 || NH < 1 || NH > 8 || d > 4096) return 0;
  if (17 + 6 * NH + 3 * B > sm::kSmallFloats / 4) return 0;
  const int nblk = inner >> 2;
  for (int cl = 4; cl <= 8; cl *= 2) {                       // one chunk per cluster of CTAs that own x_m groups
    if (nblk % (sm::kWarps * cl)) continue;
    const int nch = nblk / (sm::kWarps * cl);
    if (nch <= sm::kMaxChunks) return nch;
  }
  return 0;
}

cudaError_t launch_smallm_pre(const SmallPreParams& p, cudaStream_t s) {
  const int MR = smallm_row_bucket(p.B * p.T);
  const size_t smem = sizeof(float) * sm::smem_floats(MR, p.d, p.NH);
#define XL_SM_CASE(MRV, TV) \
  if (MR == MRV && p.T == TV) return sm::launch<MRV, TV>(p, smem, s);
  XL_SM_CASE(4, 1) XL_SM_CASE(4, 2) XL_SM_CASE(4, 3) XL_SM_CASE(4, 4)
  XL_SM_CASE(8, 1) XL_SM_CASE(8, 2) XL_SM_CASE(8, 3) XL_SM_CASE(8, 4)
  XL_SM_CASE(16, 1) XL_SM_CASE(16, 2) XL_SM_CASE(16, 3) XL_SM_CASE(16, 4)
#undef XL_SM_CASE
  return cudaErrorInvalidValue;
}

cudaError_t launch_smallm_down(const SmallDownParams& p, cudaStream_t s) {
  switch (smallm_row_bucket(p.M)) {
    case 4: return sm::launch_down<4>(p, s);
    case 8: return sm::launch_down<8>(p, s);
    case 16: return sm::launch_down<16>(p, s);
  }
  return cudaErrorInvalidValue;
}

}  // namespace xl
