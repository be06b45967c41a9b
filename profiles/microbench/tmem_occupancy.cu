// Hardware check: do two CTAs that each allocate 256 TMEM columns co-reside on one SM, and what does the occupancy
// API report for kernels that use tcgen05.alloc?  Each CTA spins for a fixed number of clock cycles; 296 CTAs of
// 256 threads with 100 KB shared memory finish in one spin period if two fit per SM, in two if not.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../raw2logit_b200/csrc -o tmem_occupancy tmem_occupancy.cu
#include <cstdio>
#include <cuda_runtime.h>
#include "isp_tmem.cuh"
using namespace r2l;

template <int COLS>
__global__ void __launch_bounds__(256, 2) spin(long long cycles, int* sink) {
    extern __shared__ float smem[];
    __shared__ uint32_t slot;
    if (COLS > 0) {
        if (threadIdx.x < 32) tmem::alloc<(COLS > 0 ? COLS : 32)>(&slot);
        tmem::fence_before_sync();
        __syncthreads();
        tmem::fence_after_sync();
    }
    const long long t0 = clock64();
    while (clock64() - t0 < cycles) { smem[threadIdx.x] += 1.f; }
    if (smem[threadIdx.x] == -1.f) *sink = 1;
    __syncthreads();
    if (COLS > 0 && threadIdx.x < 32) tmem::dealloc<(COLS > 0 ? COLS : 32)>(slot);
}

template <int COLS> void run(const char* name) {
    int* sink; cudaMalloc(&sink, 4);
    cudaFuncSetAttribute(spin<COLS>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, spin<COLS>, 256, 100 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms148 = 0, ms296 = 0;
    for (int rep = 0; rep < 2; ++rep) {
        cudaEventRecord(e0); spin<COLS><<<148, 256, 100 * 1024>>>(200000, sink); cudaEventRecord(e1);
        cudaDeviceSynchronize(); cudaEventElapsedTime(&ms148, e0, e1);
        cudaEventRecord(e0); spin<COLS><<<296, 256, 100 * 1024>>>(200000, sink); cudaEventRecord(e1);
        cudaError_t e = cudaDeviceSynchronize(); cudaEventElapsedTime(&ms296, e0, e1);
        if (e != cudaSuccess) printf("error %s\n", cudaGetErrorString(e));
    }
    printf("%-22s occupancy API %d CTAs/SM; 148 CTAs %.3f ms, 296 CTAs %.3f ms -> %s\n", name, occ, ms148, ms296,
           ms296 < 1.5f * ms148 ? "two CTAs co-reside" : "serialised");
}

int main() {
    run<0>("no TMEM");
    run<32>("32 columns / CTA");
    run<256>("256 columns / CTA");
    run<512>("512 columns / CTA");
    return 0;
}
