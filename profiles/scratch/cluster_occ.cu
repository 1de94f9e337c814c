// How many clusters of 2 / 4 / 8 CTAs (one CTA per SM: 384 threads, 185 KB of dynamic shared memory, the GEMM's footprint) can be
// co-resident on this GPU?  nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occ cluster_occ.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 185 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(sms / c * c);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute a[1];
    a[0].id = cudaLaunchAttributeClusterDimension;
    a[0].val.clusterDim.x = c; a[0].val.clusterDim.y = 1; a[0].val.clusterDim.z = 1;
    cfg.attrs = a; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("SMs %d cluster %2d: max active clusters %d (%d CTAs) %s\n", sms, c, n, n * c, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
