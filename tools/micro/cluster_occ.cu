// How many clusters of each size are co-resident on this GPU when a CTA owns its SM (512 threads, 128 registers)?
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) big(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; p[threadIdx.x] = s[(threadIdx.x * 7) % 512]; }
__global__ void __launch_bounds__(256, 2) half(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; p[threadIdx.x] = s[(threadIdx.x * 7) % 256]; }
__global__ void __launch_bounds__(384, 1) mid(float* p) { extern __shared__ float s[]; s[threadIdx.x] = 1.f; p[threadIdx.x] = s[(threadIdx.x * 7) % 384]; }
template <typename K>
void query(const char* name, K kern, int threads, int smem) {
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(threads); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
    printf("%s threads=%d smem=%d cluster=%d: max active clusters %d (%d CTAs) %s\n", name, threads, smem, cs, n, n * cs,
           e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
}
int main() {
  cudaDeviceProp p; cudaGetDeviceProperties(&p, 0);
  printf("%s SMs=%d\n", p.name, p.multiProcessorCount);
  query("big", big, 512, 120 * 1024);
  query("mid", mid, 384, 120 * 1024);
  query("half", half, 256, 60 * 1024);
  return 0;
}
