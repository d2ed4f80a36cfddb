// Warp-GEMV building blocks for the small-batch kernels (M <= 16 rows): bf16 weights streamed once with 128-bit
// loads, fp32 activations held in shared memory, fp32 FMA accumulation (activations are NOT rounded to bf16).
// Used by xl_lowlat.cu (persistent stack kernel) and xl_smallm.cu (fused LN + proj_up + conv/qkv kernel).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace xl {
namespace gv {

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned sa = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(sa), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// xs stores a row of K floats so that the 8 consecutive k a lane multiplies with one 16-byte weight load sit as
// two float4 that are bank-conflict free across the warp: float4 group at k (k % 4 == 0) lives at perm4(k).
__device__ __forceinline__ int perm4(int k) { return (k & ~255) + (((k >> 2) & 1) << 7) + (((k & 255) >> 3) << 2); }

__device__ __forceinline__ float bf_lo(unsigned w) { return __uint_as_float(w << 16); }
__device__ __forceinline__ float bf_hi(unsigned w) { return __uint_as_float(w & 0xffff0000u); }

// one batch of KB k-blocks (256 k each) of CG weight rows against the M activation rows in shared memory
template <int CG, int MR, int KB>
__device__ __forceinline__ void gemv_batch(const uint4 (&w)[KB][CG], int j0, int iters, const float* xs, int ld, int M,
                                           float (&acc)[CG][MR], int lane) {
#pragma unroll
  for (int jj = 0; jj < KB; ++jj)
    if (j0 + jj < iters) {
      const float* xb = xs + (j0 + jj) * 256 + lane * 4;
#pragma unroll
      for (int m = 0; m < MR; ++m)
        if (m < M) {
          const float4 xa = *reinterpret_cast<const float4*>(xb + m * ld);
          const float4 xc = *reinterpret_cast<const float4*>(xb + m * ld + 128);
#pragma unroll
          for (int c = 0; c < CG; ++c) {
            float a = acc[c][m];
            a = fmaf(bf_lo(w[jj][c].x), xa.x, a);
            a = fmaf(bf_hi(w[jj][c].x), xa.y, a);
            a = fmaf(bf_lo(w[jj][c].y), xa.z, a);
            a = fmaf(bf_hi(w[jj][c].y), xa.w, a);
            a = fmaf(bf_lo(w[jj][c].z), xc.x, a);
            a = fmaf(bf_hi(w[jj][c].z), xc.y, a);
            a = fmaf(bf_lo(w[jj][c].w), xc.z, a);
            a = fmaf(bf_hi(w[jj][c].w), xc.w, a);
            acc[c][m] = a;
          }
        }
    }
}

template <int CG, int KB>
__device__ __forceinline__ void gemv_load(const __nv_bfloat16* __restrict__ W, int K, int col0, int j0, uint4 (&w)[KB][CG],
                                          int lane) {
  const int iters = K >> 8;
  const int rowq = K >> 3;
  const uint4* Wp = reinterpret_cast<const uint4*>(W + (size_t)col0 * K) + lane;
#pragma unroll
  for (int jj = 0; jj < KB; ++jj)
    if (j0 + jj < iters) {
#pragma unroll
      for (int c = 0; c < CG; ++c) w[jj][c] = __ldg(Wp + (size_t)c * rowq + (j0 + jj) * 32);
    }
}

// acc[c][m] = sum over this lane's k of W[col0 + c][k] * xs[m][k]  (lane owns k = 256 j + 8 lane .. + 7)
template <int CG, int MR, int KB>
__device__ __forceinline__ void gemv_cols(const __nv_bfloat16* __restrict__ W, int K, int col0, const float* xs,
                                          int ld, int M, float (&acc)[CG][MR], int lane) {
#pragma unroll
  for (int c = 0; c < CG; ++c)
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[c][m] = 0.f;
  const int iters = K >> 8;
  for (int j0 = 0; j0 < iters; j0 += KB) {
    uint4 w[KB][CG];
    gemv_load<CG, KB>(W, K, col0, j0, w, lane);
    gemv_batch<CG, MR, KB>(w, j0, iters, xs, ld, M, acc, lane);
  }
}

// Same sums in the same order, with the first batch of weights already in registers: the caller requested it with
// gemv_load(..., j0 = 0, ...) BEFORE its dependency wait (weights are never written by a kernel), which takes one L2
// round trip off the critical path of the one-env step.
template <int CG, int MR, int KB>
__device__ __forceinline__ void gemv_cols_pre(const __nv_bfloat16* __restrict__ W, int K, int col0, const float* xs,
                                              int ld, int M, float (&acc)[CG][MR], int lane, const uint4 (&w0)[KB][CG]) {
#pragma unroll
  for (int c = 0; c < CG; ++c)
#pragma unroll
    for (int m = 0; m < MR; ++m) acc[c][m] = 0.f;
  const int iters = K >> 8;
  gemv_batch<CG, MR, KB>(w0, 0, iters, xs, ld, M, acc, lane);
  for (int j0 = KB; j0 < iters; j0 += KB) {
    uint4 w[KB][CG];
    gemv_load<CG, KB>(W, K, col0, j0, w, lane);
    gemv_batch<CG, MR, KB>(w, j0, iters, xs, ld, M, acc, lane);
  }
}

}  // namespace gv
}  // namespace xl
