// out[M,N] = A[M,K] * W[N,K]^T (+ bias[N]) (+ residual[M,N])
//
// The nn.Linear layers of the path (proj_up, proj_down: [ext-xlstm] mLSTMLayer.step; embed_state:
// multi_domain_discrete_dt_model.py:47-49; action_net: :67-81). W is stored in bf16 ("bf16 weights"),
// activations stay fp32, products are exact in fp32 and accumulated in fp32.
//
// This file holds the CUDA-core kernel: used for small M (few envs: weight-bandwidth bound, tensor cores
// cannot be filled) and as the in-library reference the tcgen05 kernel (xl_gemm_tc.cu) is tested against.
#include <cuda_bf16.h>

#include "xl_common.cuh"
#include "xl_internal.h"

namespace xl {

constexpr int BM = 64, BN = 64, BK = 16;
constexpr int PAD = 4;

__global__ void __launch_bounds__(256) gemm_simple_kernel(const float* __restrict__ A,
                                                          const __nv_bfloat16* __restrict__ W,
                                                          const float* __restrict__ bias,
                                                          const float* __restrict__ residual,
                                                          float* __restrict__ out, int M, int N, int K) {
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Ws[2][BK][BN + PAD];
  const int tid = threadIdx.x;
  pdl_wait();
  pdl_trigger();
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int ty = tid >> 4, tx = tid & 15;

  // global->register staging
  const int a_row = tid >> 2, a_k = (tid & 3) * 4;        // 64 rows x 16 k, float4 along k
  const int w_row = (tid & 127) >> 1, w_k = (tid & 1) * 8;  // 64 rows x 16 k, 8 bf16 along k
  const bool w_loader = tid < 128;

  float4 a_reg;
  uint4 w_reg;
  auto load_tiles = [&](int k0) {
    a_reg = make_float4(0.f, 0.f, 0.f, 0.f);
    const int gr = m0 + a_row, gk = k0 + a_k;
    if (gr < M && gk < K) a_reg = *reinterpret_cast<const float4*>(A + (int64_t)gr * K + gk);
    w_reg = make_uint4(0u, 0u, 0u, 0u);
    if (w_loader) {
      const int gn = n0 + w_row, gkw = k0 + w_k;
      if (gn < N && gkw < K) w_reg = *reinterpret_cast<const uint4*>(W + (int64_t)gn * K + gkw);
    }
  };
  auto store_tiles = [&](int buf) {
    As[buf][a_k + 0][a_row] = a_reg.x;
    As[buf][a_k + 1][a_row] = a_reg.y;
    As[buf][a_k + 2][a_row] = a_reg.z;
    As[buf][a_k + 3][a_row] = a_reg.w;
    if (w_loader) {
      const uint32_t ww[4] = {w_reg.x, w_reg.y, w_reg.z, w_reg.w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        Ws[buf][w_k + 2 * j][w_row] = __uint_as_float(ww[j] << 16);
        Ws[buf][w_k + 2 * j + 1][w_row] = __uint_as_float(ww[j] & 0xffff0000u);
      }
    }
  };

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int nk = (K + BK - 1) / BK;
  load_tiles(0);
  store_tiles(0);
  __syncthreads();
  for (int kb = 0; kb < nk; ++kb) {
    const int buf = kb & 1;
    if (kb + 1 < nk) load_tiles((kb + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      const float4 a4 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 w4 = *reinterpret_cast<const float4*>(&Ws[buf][k][tx * 4]);
      const float a[4] = {a4.x, a4.y, a4.z, a4.w};
      const float w[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], w[j], acc[i][j]);
    }
    if (kb + 1 < nk) store_tiles(buf ^ 1);
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int r = m0 + ty * 4 + i;
    if (r >= M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int c = n0 + tx * 4 + j;
      if (c >= N) continue;
      float v = acc[i][j];
      if (bias) v += bias[c];
      if (residual) v += residual[(int64_t)r * N + c];
      out[(int64_t)r * N + c] = v;
    }
  }
}

void launch_gemm_simple(const float* A, const __nv_bfloat16* W, const float* bias, const float* residual,
                        float* out, int M, int N, int K, cudaStream_t s) {
  dim3 grid((N + BN - 1) / BN, (M + BM - 1) / BM);
  launch_k(gemm_simple_kernel, grid, dim3(256), 0, s, A, W, bias, residual, out, M, N, K);
}

}  // namespace xl
