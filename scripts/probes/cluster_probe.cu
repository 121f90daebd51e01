// How many clusters of 2 / 4 / 8 CTAs (one CTA per SM: ~220 KB of dynamic shared memory each) can be co-resident on
// this GPU?  Decides whether a 4-CTA-cluster GEMM (TMA multicast of the shared operand) can use all 148 SMs.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fbk-fairseq-st_b200/build/cluster_probe scripts/probes/cluster_probe.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int* p) { extern __shared__ int s[]; if (p) p[0] = s[0]; }
int main() {
  const int smem = 220 * 1024;
  cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(k, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(148 / cs * cs);
    cfg.blockDim = dim3(384);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, k, &cfg);
    printf("cluster size %2d: max active clusters %d (= %d SMs)  %s\n", cs, n, n * cs, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
