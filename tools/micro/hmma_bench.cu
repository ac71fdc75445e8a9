// Micro-benchmark: throughput of legacy warp-level mma.sync (HMMA) on sm_100a, bf16 m16n8k16, fp32 accumulate.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_bench hmma_bench.cu ; ./hmma_bench
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

template <int ILP>
__global__ void k(int reps, long long* cyc, float* sink) {
  uint32_t a[4] = {threadIdx.x, 2u, 3u, 4u}, b[2] = {5u, threadIdx.x};
  float d[ILP][4];
#pragma unroll
  for (int i = 0; i < ILP; ++i) d[i][0] = d[i][1] = d[i][2] = d[i][3] = 0.f;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < ILP; ++i) mma16816(d[i], a, b);
  }
  long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < ILP; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 123.456f) sink[0] = s;
}

template <int ILP>
void run(int warps, int ctas_per_sm) {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 8 * 1024); cudaMalloc(&sink, 4);
  const int reps = 2000;
  k<ILP><<<148 * ctas_per_sm, warps * 32>>>(reps, cyc, sink);
  k<ILP><<<148 * ctas_per_sm, warps * 32>>>(reps, cyc, sink);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[8];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  const double mmas = (double)reps * ILP * warps * ctas_per_sm;   // per SM
  printf("ILP %d warps/CTA %d CTAs/SM %d: %lld cycles, %.2f cycles per mma.m16n8k16 per SM = %.0f FMA/clk/SM (%s)\n", ILP, warps,
         ctas_per_sm, h[0], h[0] / mmas, 2048.0 * mmas / h[0], cudaGetErrorString(e));
  cudaFree(cyc); cudaFree(sink);
}

int main() {
  run<1>(4, 1); run<4>(4, 1); run<8>(4, 1); run<4>(8, 1); run<8>(8, 1); run<4>(16, 1); run<8>(16, 1); run<4>(8, 2); run<8>(8, 4);
  return 0;
}
