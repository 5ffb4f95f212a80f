#include <cstdio>
#include <cuda_runtime.h>
#include <cooperative_groups.h>
__global__ void __launch_bounds__(256, 1) k(int* out) {
    extern __shared__ char sm[];
    unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if (threadIdx.x == 0) out[blockIdx.x] = smid;
    cooperative_groups::this_grid().sync();
}
int main() {
    int* d; cudaMalloc(&d, 4096);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 120000);
    int grids[] = {16, 32, 48, 96, 128, 144};
    for (int g : grids) {
        void* args[] = {&d};
        cudaLaunchCooperativeKernel((void*)k, dim3(g), dim3(256), args, 120000, 0);
        cudaDeviceSynchronize();
        int h[160]; cudaMemcpy(h, d, g * 4, cudaMemcpyDeviceToHost);
        printf("grid %d:", g);
        for (int i = 0; i < g; ++i) printf(" %d", h[i]);
        printf("\n");
    }
    return 0;
}
