// Cost of a software grid barrier on B200 (148 CTAs x 512 threads, one CTA per SM, cooperative launch), the
// building block of the persistent small-batch kernel (lram_b200/csrc/xl_lowlat.cu). Measurement aid only.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o gpurun_out/gridbar tools/microbench_grid_barrier.cu
//   gpurun_out/gridbar
//
// Variants: 0 = two-word counter + generation, ld.acquire polling (what xl_lowlat.cu shipped first)
//           1 = one monotonic 64-bit counter, fence + relaxed atomic arrive, relaxed polling + one fence at exit
//           2 = variant 1 with __nanosleep(32) back-off in the poll loop
//           3 = variant 1, but only ONE thread per CTA polls L2 and arrival uses red.release (no separate fence)
//           4 = cooperative_groups grid.sync()
// Each barrier is preceded by one dependent global store + followed by one dependent global load of a value another
// CTA wrote (the pattern a phase boundary has), so the fence has real work to order.
#include <cooperative_groups.h>
#include <cstdio>
#include <cuda_runtime.h>
namespace cg = cooperative_groups;

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_relaxed_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.relaxed.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_u64(unsigned long long* p) {
  asm volatile("red.release.gpu.global.add.u64 [%0], 1;" ::"l"(p) : "memory");
}

template <int V>
__global__ void __launch_bounds__(512, 1) bar_kernel(unsigned* bar32, unsigned long long* bar64, float* data,
                                                     long long* out, int iters) {
  cg::grid_group grid = cg::this_grid();
  const int G = gridDim.x, tid = threadIdx.x;
  __shared__ unsigned s_gen;
  __shared__ unsigned long long s_base;
  if (tid == 0) {
    s_gen = ld_acquire_u32(bar32 + 1);
    const unsigned long long per = (unsigned long long)G * iters;
    s_base = ld_relaxed_u64(bar64) / per * per;
  }
  __syncthreads();
  unsigned gen = s_gen;
  unsigned long long target = s_base;
  float acc = 0.f;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    data[(size_t)blockIdx.x * 512 + tid] = acc + it;                 // phase output
    if (V == 0) {
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        const unsigned a = atomicAdd(bar32, 1u);
        if (a == (unsigned)G - 1u) {
          bar32[0] = 0u;
          __threadfence();
          atomicAdd(bar32 + 1, 1u);
        } else {
          while (ld_acquire_u32(bar32 + 1) == gen) {}
        }
      }
      gen += 1u;
      __syncthreads();
    } else if (V == 1 || V == 2) {
      target += G;
      __syncthreads();
      if (tid == 0) {
        __threadfence();
        atomicAdd(bar64, 1ull);
        while (ld_relaxed_u64(bar64) < target) {
          if (V == 2) __nanosleep(32);
        }
        __threadfence();
      }
      __syncthreads();
    } else if (V == 3) {
      target += G;
      __syncthreads();
      if (tid == 0) {
        red_release_u64(bar64);
        while (ld_relaxed_u64(bar64) < target) {}
        asm volatile("fence.acq_rel.gpu;" ::: "memory");
      }
      __syncthreads();
    } else {
      grid.sync();
    }
    // consume what the neighbour CTA wrote before the barrier
    acc += __ldcg(data + (size_t)((blockIdx.x + 1) % G) * 512 + tid);
  }
  const long long t1 = clock64();
  if (tid == 0) out[blockIdx.x] = t1 - t0;
  if (acc == -1.f) out[0] = 0;
}

template <int V>
static void run(const char* name, int G, int iters, unsigned* bar32, unsigned long long* bar64, float* data,
                long long* out) {
  void* args[] = {&bar32, &bar64, &data, &out, &iters};
  cudaMemset(bar32, 0, 64);
  cudaMemset(bar64, 0, 64);
  for (int rep = 0; rep < 3; ++rep) {
    cudaError_t e = cudaLaunchCooperativeKernel((void*)bar_kernel<V>, dim3(G), dim3(512), args, 0, 0);
    if (e != cudaSuccess) { printf("%s: launch failed: %s\n", name, cudaGetErrorString(e)); return; }
    e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); return; }
  }
  long long h[256];
  cudaMemcpy(h, out, sizeof(long long) * G, cudaMemcpyDeviceToHost);
  long long mx = 0;
  for (int i = 0; i < G; ++i) mx = h[i] > mx ? h[i] : mx;
  printf("%-58s %8.0f SM clocks per (store + barrier + dependent load)\n", name, (double)mx / iters);
}

// The alternative to a barrier: a kernel boundary. Same store -> dependent load pattern as 2000 tiny kernels replayed
// from a CUDA graph, with and without programmatic dependent launch.
__global__ void __launch_bounds__(512, 1) step_kernel(float* data, int it) {
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int G = gridDim.x, tid = threadIdx.x;
  const float v = __ldcg(data + (size_t)(it & 1) * 512 * G + (size_t)((blockIdx.x + 1) % G) * 512 + tid);
  data[(size_t)((it + 1) & 1) * 512 * G + (size_t)blockIdx.x * 512 + tid] = v + it;
}

static void run_graph(int G, int iters, float* data, int pdl) {
  cudaStream_t s;
  cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaGraph_t graph;
  cudaGraphExec_t exec;
  cudaStreamBeginCapture(s, cudaStreamCaptureModeThreadLocal);
  for (int it = 0; it < iters; ++it) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(G); cfg.blockDim = dim3(512); cfg.stream = s;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = pdl ? 1 : 0;
    cudaLaunchKernelEx(&cfg, step_kernel, data, it);
  }
  cudaStreamEndCapture(s, &graph);
  cudaGraphInstantiate(&exec, graph, 0);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e30f;
  for (int rep = 0; rep < 4; ++rep) {
    cudaEventRecord(e0, s);
    cudaGraphLaunch(exec, s);
    cudaEventRecord(e1, s);
    cudaStreamSynchronize(s);
    float ms = 0.f;
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep && ms < best) best = ms;
  }
  printf("kernel boundary in a CUDA graph, PDL %d: %40.3f us per (kernel: dependent load + store)\n", pdl, best * 1e3 / iters);
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  unsigned* bar32; unsigned long long* bar64; float* data; long long* out;
  cudaMalloc(&bar32, 64); cudaMalloc(&bar64, 64); cudaMalloc(&data, sizeof(float) * 512 * sms); cudaMalloc(&out, 8 * 256);
  cudaMemset(data, 0, sizeof(float) * 512 * sms);
  const int iters = 2000;
  printf("grid barrier on %d SMs, %d iterations\n", sms, iters);
  run<0>("0 counter + generation, fence, ld.acquire poll", sms, iters, bar32, bar64, data, out);
  run<1>("1 monotonic u64, fence + atomic, relaxed poll", sms, iters, bar32, bar64, data, out);
  run<2>("2 = 1 + nanosleep(32) back-off", sms, iters, bar32, bar64, data, out);
  run<3>("3 monotonic u64, red.release arrive, relaxed poll, fence", sms, iters, bar32, bar64, data, out);
  run<4>("4 cooperative_groups grid.sync()", sms, iters, bar32, bar64, data, out);
  float* d2; cudaMalloc(&d2, sizeof(float) * 2 * 512 * sms); cudaMemset(d2, 0, sizeof(float) * 2 * 512 * sms);
  run_graph(sms, iters, d2, 0);
  run_graph(sms, iters, d2, 1);
  int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  printf("SM clock (max) %d kHz: 1000 clocks = %.3f us\n", khz, 1e6 / khz);
  return 0;
}
