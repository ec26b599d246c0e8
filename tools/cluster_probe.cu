// How many clusters of 1/2/4/8 CTAs (one CTA per SM: 200 KB of dynamic shared memory, 224 threads -- the footprint of the
// WN pair kernels) can be resident on this GPU at once?  Answers whether a 4-CTA cluster (two CTA pairs sharing one weight
// tile by TMA multicast) could still use every SM.      nvcc -arch=sm_100a tools/cluster_probe.cu -o /tmp/cluster_probe
#include <cstdio>
#include <cuda_runtime.h>

__global__ void probe_kernel(int* out) {
    extern __shared__ unsigned char smem[];
    if (out && threadIdx.x == 0) out[blockIdx.x] = smem[0];
}

int main() {
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    const int smem = 200 * 1024;
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"smem_per_cta\": %d, \"max_active_clusters\": {", prop.name, prop.multiProcessorCount, smem);
    const int sizes[] = {1, 2, 4, 8, 16};
    for (int i = 0; i < 5; ++i) {
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(sizes[i] * 148, 1, 1);
        cfg.blockDim = dim3(224, 1, 1);
        cfg.dynamicSmemBytes = smem;
        cudaLaunchAttribute attr;
        attr.id = cudaLaunchAttributeClusterDimension;
        attr.val.clusterDim.x = sizes[i];
        attr.val.clusterDim.y = 1;
        attr.val.clusterDim.z = 1;
        cfg.attrs = &attr;
        cfg.numAttrs = 1;
        int n = -1;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&n, probe_kernel, &cfg);
        printf("%s\"%d\": {\"clusters\": %d, \"ctas\": %d, \"err\": \"%s\"}", i ? ", " : "", sizes[i], n, n * sizes[i],
               e == cudaSuccess ? "" : cudaGetErrorString(e));
    }
    printf("}}\n");
    return 0;
}
