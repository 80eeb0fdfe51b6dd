// How many clusters of a 1-CTA-per-SM kernel (232 KB of shared memory, 384 threads) can be resident at once, by cluster size?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 scripts/cluster_occupancy.cu -o scripts/bin/cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(384, 1) k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 232448;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs); cfg.blockDim = dim3(384); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d (= %d CTAs)  [%s]\n", cs, n, n * cs, cudaGetErrorString(e));
  }
  return 0;
}
