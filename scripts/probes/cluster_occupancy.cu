// Probe: how many thread-block clusters of size 2 / 4 / 8 (one 227 KB-smem CTA per SM) can be co-resident on this GPU?
// nvcc -gencode arch=compute_100a,code=sm_100a -o cluster_occupancy cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void dummy_kernel(int* p) { if (p) *p = 1; }
int main() {
  cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(cs * 200); cfg.blockDim = dim3(256); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy_kernel, &cfg);
    printf("cluster size %2d: max active clusters %d (%d SMs) %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
