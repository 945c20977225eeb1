// How many thread-block clusters of size 2 / 4 / 8 with the GEMM kernel's footprint (one CTA per SM: ~198 KB of dynamic
// shared memory, 320 threads) can a B200 hold at once?  148 SMs = 74 pairs; clusters must sit inside one GPC, so a
// 4-CTA cluster (W-tile multicast across two CTA pairs) loses every SM a GPC has beyond a multiple of 4.
//   nvcc -gencode arch=compute_100a,code=sm_100a -o build/cluster_occupancy scripts/probes/cluster_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>

__global__ void dummy(int* p) {
  extern __shared__ char sm[];
  if (p) p[0] = sm[threadIdx.x];
}

int main() {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int smem = 198 * 1024;
  cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  for (int cs : {1, 2, 4, 8, 16}) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((sms / cs) * cs);
    cfg.blockDim = dim3(320);
    cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at;
    at.id = cudaLaunchAttributeClusterDimension;
    at.val.clusterDim.x = cs;
    at.val.clusterDim.y = 1;
    at.val.clusterDim.z = 1;
    cfg.attrs = &at;
    cfg.numAttrs = 1;
    int n = -1;
    cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
    printf("cluster size %2d: max active clusters %d (%d of %d SMs) %s\n", cs, n, n * cs, sms, e == cudaSuccess ? "" : cudaGetErrorString(e));
  }
  return 0;
}
