// How many clusters of C one-CTA-per-SM blocks (200 KB dynamic shared memory each) stay resident on this GPU.
// nvcc -arch=sm_100a scripts/probes/cluster_occupancy.cu -o /tmp/cluster_occupancy && /tmp/cluster_occupancy
#include <cstdio>
#include <cuda_runtime.h>
__global__ void __launch_bounds__(320, 1) k(int* p) { extern __shared__ char s[]; if (p) p[0] = s[0]; }
int main() {
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int c = 1; c <= 16; ++c) {
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = c; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.blockDim = dim3(320); cfg.gridDim = dim3(c); cfg.dynamicSmemBytes = 200 * 1024; cfg.attrs = attr; cfg.numAttrs = 1;
    int n = 0;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster %2d: max active clusters %3d -> %3d SMs (%s)\n", c, n, n * c, cudaGetErrorString(e));
  }
  return 0;
}
