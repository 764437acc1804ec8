// Probe: how many thread-block clusters of a given size can be co-resident on this GPU when every CTA takes
// (almost) a whole SM's shared memory - decides the cluster size of the fused scoring kernel.
#include <cuda_runtime.h>
#include <cstdio>
__global__ void dummy(int* out) { extern __shared__ char s[]; if (out) out[0] = s[0]; }
int main() {
    int dev = 0, sms = 0; cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    printf("{\"sms\": %d, \"clusters\": [", sms);
    const int sizes[] = {1, 2, 3, 4, 5, 6, 7, 8, 10, 12, 16};
    bool first = true;
    for (int smem : {232448, 200 * 1024, 100 * 1024}) {
        cudaFuncSetAttribute(dummy, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        cudaFuncSetAttribute(dummy, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
        for (int cs : sizes) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = dim3(cs * 64); cfg.blockDim = dim3(640); cfg.dynamicSmemBytes = smem;
            cudaLaunchAttribute at[1]; at[0].id = cudaLaunchAttributeClusterDimension;
            at[0].val.clusterDim.x = cs; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
            cfg.attrs = at; cfg.numAttrs = 1;
            int n = -1; cudaError_t e = cudaOccupancyMaxActiveClusters(&n, dummy, &cfg);
            printf("%s{\"smem\": %d, \"cs\": %d, \"max_active\": %d, \"sms_used\": %d, \"err\": %d}", first ? "" : ", ", smem, cs, n, n * cs, (int)e);
            first = false; cudaGetLastError();
        }
    }
    printf("]}\n");
    return 0;
}
