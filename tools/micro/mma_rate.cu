// Micro-benchmark: issue rate of mma.sync m16n8k8 tf32 (legacy tensor path) on sm_100a, 16 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ void mma_tf32(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_bf16(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void mma_f16_k8(float (&c)[4], const unsigned (&a)[4], unsigned b0, unsigned b1) {
  asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
               : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
               : "r"(a[0]), "r"(a[1]), "r"(b0));
}
template <int KIND>
__global__ void __launch_bounds__(512, 1) k2(float* out, long long* cyc, int iters) {
  unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  float c[3][4] = {};
  unsigned b0 = threadIdx.x, b1 = threadIdx.x + 1;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 24; ++j) {
      if (KIND == 0) mma_f16(c[j % 3], a, b0 + j, b1);
      if (KIND == 1) mma_bf16(c[j % 3], a, b0 + j, b1);
      if (KIND == 2) mma_f16_k8(c[j % 3], a, b0 + j, b1);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < 3; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int CHAINS>
__global__ void __launch_bounds__(512, 1) k(float* out, long long* cyc, int iters) {
  unsigned a[4] = {threadIdx.x, threadIdx.x * 3u, threadIdx.x * 5u, threadIdx.x * 7u};
  float c[CHAINS][4] = {};
  unsigned b0 = threadIdx.x, b1 = threadIdx.x + 1;
  __syncthreads();
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int j = 0; j < 24; ++j) mma_tf32(c[j % CHAINS], a, b0 + j, b1);
  }
  __syncthreads();
  long long t1 = clock64();
  float s = 0;
  for (int j = 0; j < CHAINS; ++j) s += c[j][0] + c[j][1] + c[j][2] + c[j][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 512 * 4); cudaMalloc(&cyc, 148 * 8);
  const int iters = 1000;
  for (int rep = 0; rep < 2; ++rep) {
    k<3><<<148, 512>>>(out, cyc, iters);
    long long h[148]; cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("chains=3 16 warps: %.2f cycles per MMA per SM (warp-level %.1f)\n", (double)h[0] / (iters * 24.0 * 16), (double)h[0] / (iters * 24.0));
    k<1><<<148, 512>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("chains=1 16 warps: %.2f cycles per MMA per SM (dependent latency/warp %.1f)\n", (double)h[0] / (iters * 24.0 * 16), (double)h[0] / (iters * 24.0));
    k<3><<<148, 128>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("chains=3 4 warps: %.2f cycles per MMA per SM\n", (double)h[0] / (iters * 24.0 * 4));
    k<1><<<148, 32>>>(out, cyc, iters);
    cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("chains=1 1 warp: latency %.1f cycles\n", (double)h[0] / (iters * 24.0));
  }
  {
    long long h[148];
    k2<0><<<148, 512>>>(out, cyc, iters); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("f16 m16n8k16, 16 warps: %.2f cycles per MMA per SM\n", (double)h[0] / (iters * 24.0 * 16));
    k2<1><<<148, 512>>>(out, cyc, iters); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("bf16 m16n8k16, 16 warps: %.2f cycles per MMA per SM\n", (double)h[0] / (iters * 24.0 * 16));
    k2<2><<<148, 512>>>(out, cyc, iters); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("f16 m16n8k8, 16 warps: %.2f cycles per MMA per SM\n", (double)h[0] / (iters * 24.0 * 16));
    k2<0><<<148, 128>>>(out, cyc, iters); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("f16 m16n8k16, 4 warps: %.2f cycles per MMA per SM\n", (double)h[0] / (iters * 24.0 * 4));
    k2<0><<<148, 32>>>(out, cyc, iters); cudaMemcpy(h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    printf("f16 m16n8k16, 1 warp (3 chains): %.2f cycles per MMA\n", (double)h[0] / (iters * 24.0));
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
