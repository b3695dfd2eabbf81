// dev: occupancy of a 128-register kernel with dynamic shared memory for small CTAs (why are only 4 single-problem CTAs resident?)
#include <cstdio>
#include <cuda_runtime.h>
extern __shared__ double sm[];
__global__ void __launch_bounds__(512) k(double *o, int n)
{
    double a[40];
    for (int i = 0; i < 40; ++i) a[i] = sm[(threadIdx.x + i) % n];
    for (int j = 0; j < n; ++j) for (int i = 0; i < 40; ++i) a[i] = a[i] * a[(i + 1) % 40] + sm[j];
    double s = 0; for (int i = 0; i < 40; ++i) s += a[i];
    o[threadIdx.x] = s;
}
int main()
{
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    cudaFuncAttributes fa; cudaFuncGetAttributes(&fa, k);
    printf("regs %d static smem %zu\n", fa.numRegs, fa.sharedSizeBytes);
    for (int carve = 0; carve < 2; ++carve) {
        if (carve) cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        for (int t : {32, 64, 128, 256, 512})
            for (int smem : {11184, 21000, 31520, 84000}) {
                int occ = 0; cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k, t, smem);
                printf("carveout %d threads %d smem %d -> %d blocks/SM\n", carve, t, smem, occ);
            }
    }
    return 0;
}
