// Micro-benchmark: global-store throughput of a lightly occupied SM (2 CTAs x 256 threads per SM, as the tcgen05 kernels run):
// st.global.v4 (16 B / thread), st.global.v8 (32 B / thread), and cp.async.bulk shared -> global (TMA engine, issued by one thread).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o store_bench store_bench.cu
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// each CTA writes `bytes_per_cta` bytes (its own contiguous region) in rounds of 64 KB
template <int MODE>
__global__ void __launch_bounds__(256, 2) k(uint8_t* out, size_t bytes_per_cta, int work) {
  extern __shared__ __align__(128) uint8_t sm[];
  uint8_t* base = out + (size_t)blockIdx.x * bytes_per_cta;
  const int tid = threadIdx.x;
  float acc = tid;
  for (size_t off = 0; off < bytes_per_cta; off += 65536) {
    for (int i = 0; i < work; ++i) acc = fmaf(acc, 1.0001f, 0.5f);   // stand-in for the work between bursts
    if (MODE == 0) {          // 16 x st.global.v4 per thread: 256 threads x 16 B = 4 KB per instruction slot
#pragma unroll
      for (int j = 0; j < 16; ++j) {
        float4 v = make_float4(acc, acc, acc, acc);
        asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(base + off + (size_t)j * 4096 + tid * 16), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
      }
    } else if (MODE == 1) {   // 8 x st.global.v8
#pragma unroll
      for (int j = 0; j < 8; ++j)
        asm volatile("st.global.v8.f32 [%0], {%1,%1,%1,%1,%1,%1,%1,%1};" ::"l"(base + off + (size_t)j * 8192 + tid * 32), "f"(acc) : "memory");
    } else if (MODE >= 3) {   // smem staging + many small bulk stores (MODE 3: 128 x 512 B, MODE 4: 32 x 2 KB) issued by the lanes of warp 0
      const int piece = MODE == 3 ? 512 : 2048, per_lane = 65536 / piece / 32;
      if (off) {
        if (tid < 32) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) *reinterpret_cast<float4*>(sm + j * 4096 + tid * 16) = make_float4(acc, acc, acc, acc);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid < 32) {
        for (int j = 0; j < per_lane; ++j) {
          const int o = (j * 32 + tid) * piece;
          asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off + o), "r"(smem_u32(sm) + o), "r"(piece) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    } else {                  // smem staging + one bulk store of 64 KB
      if (off) {
        if (tid == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
      }
#pragma unroll
      for (int j = 0; j < 16; ++j) *reinterpret_cast<float4*>(sm + j * 4096 + tid * 16) = make_float4(acc, acc, acc, acc);
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      __syncthreads();
      if (tid == 0) {
        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(base + off), "r"(smem_u32(sm)), "r"(65536) : "memory");
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      }
    }
  }
  if (MODE == 2 && tid == 0) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (MODE >= 3 && tid < 32) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  if (acc == 12345.678f) out[0] = 1;
}

template <int MODE>
void run(const char* name, int work) {
  const size_t per = 64ull << 20 >> 2;  // 16 MB per CTA
  const int grid = 296;
  uint8_t* out;
  cudaMalloc(&out, per * grid);
  cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, 65536);
  cudaEvent_t a, b;
  cudaEventCreate(&a); cudaEventCreate(&b);
  k<MODE><<<grid, 256, MODE >= 2 ? 65536 : 0>>>(out, per, work);
  cudaEventRecord(a);
  k<MODE><<<grid, 256, MODE >= 2 ? 65536 : 0>>>(out, per, work);
  cudaEventRecord(b);
  cudaError_t e = cudaDeviceSynchronize();
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  printf("%-28s work %5d: %.3f ms, %.2f TB/s  (%s)\n", name, work, ms, per * grid / ms / 1e9, cudaGetErrorString(e));
  cudaFree(out);
}

int main() {
  for (int work : {0, 2000, 8000}) {
    run<0>("st.global.v4 (no_allocate)", work);
    run<1>("st.global.v8", work);
    run<2>("smem + cp.async.bulk store", work);
    run<3>("smem + 128 x 512 B bulk", work);
    run<4>("smem + 32 x 2 KB bulk", work);
  }
  return 0;
}
