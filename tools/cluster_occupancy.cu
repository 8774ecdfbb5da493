// How many thread-block clusters of 1 / 2 / 4 / 8 CTAs can be co-resident on this GPU when a CTA takes a whole SM
// (200 KB of dynamic shared memory, as the persistent GEMM kernels do)? Answers whether a 4-CTA cluster (two CTA pairs
// sharing an operand panel by TMA multicast) can still cover all SMs.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -o tools/cluster_occupancy tools/cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe_kernel(int* out) {
  extern __shared__ unsigned char smem[];
  if (out != nullptr && threadIdx.x == 0 && blockIdx.x == 0) out[0] = smem[0];
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int smem = 200 * 1024;
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  printf("%s: %d SMs\n", prop.name, prop.multiProcessorCount);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(cs * 64);
    cfg.blockDim = dim3(256);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = cs;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
    printf("cluster size %2d: max co-resident clusters %3d = %3d CTAs (one per SM) = %5.1f %% of the SMs  [%s]\n", cs, n,
           n * cs, 100.0 * n * cs / prop.multiProcessorCount, cudaGetErrorString(e));
  }
  return 0;
}
