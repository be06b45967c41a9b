// Microbenchmark / hardware check: TMEM as a per-thread accumulator file (raw2logit_b200/csrc/isp_tmem.cuh).
// 296 CTAs x 256 threads, two CTAs per SM (100 KB dynamic shared memory each), 256 TMEM columns per CTA;
// every thread keeps 96 running sums in "its" 96 columns and updates them in four groups per round, the way the
// fused backward's phases do.  Verifies the sums on the host and times one load-update-store cycle.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../raw2logit_b200/csrc -o tmem_park tmem_park.cu ; ./tmem_park
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>
#include "isp_tmem.cuh"

using namespace r2l;

template <int N> __device__ __forceinline__ void bump(uint32_t a, int round, int col0) {
    float v[N];
    tmem::load<N>(a, v);
    tmem::wait_ld();
#pragma unroll
    for (int i = 0; i < N; ++i) v[i] += (float)((threadIdx.x + (col0 + i) * 3 + round) & 7);
    tmem::store<N>(a, v);
    tmem::wait_st();
}

template <bool PARK>
__global__ void __launch_bounds__(256, 2) k(float* out, int rounds) {
    extern __shared__ float smem[];
    __shared__ uint32_t slot;
    if (threadIdx.x < 32) tmem::alloc<256>(&slot);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    const uint32_t base = slot;
    const int col0 = (threadIdx.x >> 7) * 96;
    const uint32_t a0 = tmem::addr(base, col0);
    {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
#pragma unroll
        for (int c = 0; c < 96; c += 16) tmem::store<16>(a0 + c, z);
        tmem::wait_st();
    }
    float keep = 0.f;
    for (int r = 0; r < rounds; ++r) {
        if (PARK) {
            bump<2>(a0, r, 0);
            bump<25>(a0 + 2, r, 2);
            bump<9>(a0 + 27, r, 27);
            bump<20>(a0 + 36, r, 36);
            bump<20>(a0 + 56, r, 56);
            bump<20>(a0 + 76, r, 76);
        } else {
            keep += (float)((threadIdx.x + r) & 7);
        }
        smem[threadIdx.x] = keep;          // keep the dynamic allocation alive
        __syncthreads();
    }
    float v[96];
#pragma unroll
    for (int c = 0; c < 96; c += 16) tmem::load<16>(a0 + c, v + c);
    tmem::wait_ld();
    float* o = out + ((size_t)blockIdx.x * 256 + threadIdx.x) * 96;
#pragma unroll
    for (int i = 0; i < 96; ++i) o[i] = v[i] + (PARK ? 0.f : keep + smem[(threadIdx.x + 1) & 255]);
    tmem::fence_before_sync();
    __syncthreads();
    if (threadIdx.x < 32) tmem::dealloc<256>(base);
}

int main() {
    const int grid = 296, rounds = 64;
    float* d;
    cudaMalloc(&d, (size_t)grid * 256 * 96 * 4);
    cudaFuncSetAttribute(k<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    cudaFuncSetAttribute(k<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k<true>, 256, 100 * 1024);
    printf("occupancy %d CTAs/SM\n", occ);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms[2] = {0, 0};
    for (int mode = 0; mode < 2; ++mode) {
        for (int rep = 0; rep < 3; ++rep) {
            cudaEventRecord(e0);
            if (mode == 0) k<true><<<grid, 256, 100 * 1024>>>(d, rounds); else k<false><<<grid, 256, 100 * 1024>>>(d, rounds);
            cudaEventRecord(e1);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("CUDA error: %s\n", cudaGetErrorString(e)); return 1; }
            cudaEventElapsedTime(&ms[mode], e0, e1);
        }
        if (mode == 0) {
            std::vector<float> h((size_t)grid * 256 * 96);
            cudaMemcpy(h.data(), d, h.size() * 4, cudaMemcpyDeviceToHost);
            size_t bad = 0;
            for (int b = 0; b < grid; ++b)
                for (int t = 0; t < 256; ++t)
                    for (int c = 0; c < 96; ++c) {
                        float want = 0.f;
                        for (int r = 0; r < rounds; ++r) want += (float)((t + c * 3 + r) & 7);
                        if (h[((size_t)b * 256 + t) * 96 + c] != want) ++bad;
                    }
            printf("TMEM park check: %zu mismatches of %zu\n", bad, h.size());
        }
    }
    printf("park kernel %.3f ms, empty kernel %.3f ms -> %.1f ns per round of 6 load-update-store groups (96 columns)\n",
           ms[0], ms[1], (ms[0] - ms[1]) * 1e6 / rounds);
    return 0;
}
