// Prints how many clusters of 1/2/4/8 CTAs (512 threads, ~210 KB dynamic smem: one CTA per SM) can be co-resident.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(512, 1) dummy(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 215000;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cl : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cl * cl); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at; at.id = cudaLaunchAttributeClusterDimension; at.val.clusterDim.x = cl; at.val.clusterDim.y = 1; at.val.clusterDim.z = 1;
    cfg.attrs = &at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %d (%d CTAs)  %s\n", cl, n, n * cl, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
