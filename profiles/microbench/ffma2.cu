// Microbenchmark: scalar FFMA vs packed fma.rn.f32x2 issue throughput per SM, and FFMA + LDS mixes.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu ; ./ffma2
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void __launch_bounds__(256) k(float* out, int iters, float s) {
    __shared__ float sm[1024];
    for (int i = threadIdx.x; i < 1024; i += 256) sm[i] = s * i;
    __syncthreads();
    float a[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) a[i] = threadIdx.x * 0.001f + i;
    float2 b[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) b[i] = make_float2(a[2 * i], a[2 * i + 1]);
    const float w0 = s, w1 = s * 0.5f;
    const float2 w2 = make_float2(w0, w1);
    for (int it = 0; it < iters; ++it) {
        if (MODE == 0) {          // 16 independent scalar FFMA chains
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, w1);
        } else if (MODE == 1) {   // 8 independent FFMA2 chains (same flop count)
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 8; ++i) b[i] = __ffma2_rn(b[i], w2, w2);
        } else if (MODE == 2) {   // scalar FFMA with one LDS.32 per 4 FFMA
#pragma unroll
            for (int r = 0; r < 8; ++r) {
#pragma unroll
                for (int i = 0; i < 16; i += 4) {
                    const float v = sm[(threadIdx.x + it + r * 16 + i) & 1023];
                    a[i] = fmaf(a[i], w0, v); a[i + 1] = fmaf(a[i + 1], w0, v);
                    a[i + 2] = fmaf(a[i + 2], w0, v); a[i + 3] = fmaf(a[i + 3], w0, v);
                }
            }
        } else if (MODE == 3) {   // scalar FFMA with one LDS.32 per FFMA (the v1 kernels' pattern)
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int i = 0; i < 16; ++i) a[i] = fmaf(sm[(threadIdx.x + it + r * 16 + i) & 1023], w0, a[i]);
        }
    }
    float acc = 0.f;
#pragma unroll
    for (int i = 0; i < 16; ++i) acc += a[i];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc += b[i].x + b[i].y;
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int MODE> void run(const char* name, float* out) {
    const int iters = 4096, blocks = 148 * 8;
    k<MODE><<<blocks, 256>>>(out, 16, 1.0001f);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<MODE><<<blocks, 256>>>(out, iters, 1.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double fma = (double)blocks * 256 * iters * 128.0;
    printf("%-28s %8.3f ms  %8.2f TFMA/s  (%.1f fma/clk/SM at 1.9 GHz)\n", name, ms, fma / ms * 1e-9,
           fma / (ms * 1e-3) / 148 / 1.9e9);
}

int main() {
    float* out; cudaMalloc(&out, 148 * 8 * 256 * 4);
    run<0>("scalar FFMA", out);
    run<1>("packed FFMA2", out);
    run<2>("FFMA + 1 LDS per 4", out);
    run<3>("FFMA + 1 LDS per 1", out);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0); printf("clock attr %d kHz\n", clk);
    return 0;
}
