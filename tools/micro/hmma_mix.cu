// Micro-benchmark: does a legacy HMMA (mma.sync.m16n8k16 bf16) share the scheduler's issue bandwidth with ALU work?  Each warp runs
// reps x (4 HMMA + NF FFMA); if the two overlap the time is max(4 x 8, NF) cycles per scheduler and round (4 warps per scheduler:
// x 4), if they serialise it is the sum.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hmma_mix_probe hmma_mix.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
__device__ __forceinline__ void mma16816(float* d, const uint32_t* a, const uint32_t* b) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
               : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
template <int NH, int NF>
__global__ void __launch_bounds__(512) k(int reps, long long* cyc, float* sink) {
  uint32_t a[4] = {threadIdx.x, 2u, 3u, 4u}, b[2] = {5u, threadIdx.x};
  float d[4][4] = {};
  float f[8];
  for (int i = 0; i < 8; ++i) f[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
#pragma unroll
    for (int i = 0; i < NH; ++i) mma16816(d[i & 3], a, b);
#pragma unroll
    for (int i = 0; i < NF; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(f[i & 7]) : "f"(1.0001f), "f"(0.5f));
  }
  long long t1 = clock64();
  float s = 0.f;
  for (int i = 0; i < 4; ++i) s += d[i][0] + d[i][1] + d[i][2] + d[i][3];
  for (int i = 0; i < 8; ++i) s += f[i];
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
  if (s == 123.456f) sink[0] = s;
}
template <int NH, int NF>
void run() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 8 * 1024); cudaMalloc(&sink, 4);
  const int reps = 2000;
  k<NH, NF><<<148, 512>>>(reps, cyc, sink);
  k<NH, NF><<<148, 512>>>(reps, cyc, sink);
  cudaDeviceSynchronize();
  long long h[2];
  cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
  // 16 warps per SM = 4 per scheduler
  printf("per round and scheduler (4 warps): %d HMMA + %3d FFMA per warp -> %.1f cycles  (HMMA alone: %d, FFMA alone: %d)\n", NH, NF,
         (double)h[0] / reps, NH * 4 * 8, NF * 4);
  cudaFree(cyc); cudaFree(sink);
}
int main() {
  run<4, 0>(); run<0, 32>(); run<4, 32>(); run<4, 64>(); run<4, 16>(); run<2, 32>(); run<8, 32>();
  return 0;
}
