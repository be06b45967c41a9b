// isp_stages.cu -- the staged mode of the parametrized ISP: one kernel per stage, every stage its own tensor.
//
// Reference: ParametrizedProcessing.forward with track_stages=True (pipeline_torch.py:183-221): `stages['demosaic']`,
// `['color_correct']`, `['sharpening']`, `['gaussian']`, `['clipped']`, `['gamma_correct']` are kept, and -- when the raw
// batch requires a gradient -- each retains its `.grad` for model.track_images (model.py:229-254).  That is an inspection
// path run on a few images at epoch ends; the fused kernels (isp_fwd3.cuh / isp_bwd5.cuh) never materialise the stages.
//
// Every linear stage of the chain is a 3 -> 3 channel K x K correlation of the stage before it:
//   color_correct = CCM . diag(wb) . Debayer(reflect-1 pad)                      (:187-191)   K = 3, reflect
//   sharpening    = M_yuv2rgb . [sharpen(Y) | U | V] . M_rgb2yuv                 (:194-198)   K = 3, zero pad
//   gaussian      = M_yuv2rgb . [gauss(Y, reflect-2 pad) | U | V] . M_rgb2yuv    (:199-203)   K = 5, reflect
// (the YUV -> RGB -> YUV round trip between the last two is part of the chain, as in the reference), so the host forms
// the combined [3][3][K][K] weight with tiny differentiable torch ops (raw2logit_b200/staged.py) and this file provides
// the ONE differentiable operator they need -- forward, input gradient, weight gradient -- plus the two pointwise stages
// (clip :206, gamma :209 with its d/dgamma).  Simple one-thread-per-pixel kernels: clarity over speed here.
// Parameter-gradient reductions are two-stage (per-CTA partial sums, then one CTA adds them in double in CTA order):
// bit-reproducible run to run.
#include <cuda_runtime.h>
#include <stdint.h>

#include "isp_launch.h"

namespace r2l {

constexpr int kStageNT = 256;
constexpr int kStageMaxCtas = 1024;          // rows of the partial-sum workspace

// source row / column of a padded index: reflect (no edge repeat, torch 'reflect') or -1 for a zero-padded site
__device__ __forceinline__ int pad_src(int i, int n, int reflect) {
    if (i >= 0 && i < n) return i;
    if (!reflect) return -1;
    return i < 0 ? -i : 2 * (n - 1) - i;
}

// y[b][co][p] = sum_ci sum_ab w[co][ci][a][b] x_pad[b][ci][p + (a - r, b - r)]
template <int K>
__global__ void __launch_bounds__(kStageNT) stage_conv_kernel(const float* __restrict__ x, const float* __restrict__ w, int B,
                                                              int H, int W, int reflect, float* __restrict__ y) {
    constexpr int R = K / 2;
    __shared__ float sw[9 * K * K];
    for (int i = threadIdx.x; i < 9 * K * K; i += kStageNT) sw[i] = w[i];
    __syncthreads();
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    for (size_t idx = (size_t)blockIdx.x * kStageNT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kStageNT) {
        const int b = (int)(idx / plane);
        const int p = (int)(idx - (size_t)b * plane);
        const int py = p / W, px = p - py * W;
        const float* xb = x + (size_t)b * 3 * plane;
        float acc[3] = {0.f, 0.f, 0.f};
        for (int a = 0; a < K; ++a) {
            const int sy = pad_src(py + a - R, H, reflect);
            if (sy < 0) continue;
            for (int c = 0; c < K; ++c) {
                const int sx = pad_src(px + c - R, W, reflect);
                if (sx < 0) continue;
                const size_t off = (size_t)sy * W + sx;
#pragma unroll
                for (int ci = 0; ci < 3; ++ci) {
                    const float v = __ldg(xb + ci * plane + off);
#pragma unroll
                    for (int co = 0; co < 3; ++co) acc[co] = fmaf(sw[((co * 3 + ci) * K + a) * K + c], v, acc[co]);
                }
            }
        }
        float* yb = y + (size_t)b * 3 * plane + p;
        yb[0] = acc[0]; yb[plane] = acc[1]; yb[2 * plane] = acc[2];
    }
}

// gx[b][ci][q] = sum_co sum_ab w[co][ci][a][b] sum_{(i, j) padded sites whose source is q} gy[b][co][(i, j) - (a - r, b - r)]
// The padded sites that read q: q itself and, under reflect padding, -q (1 <= q <= r) and 2(n-1) - q (n-1-r <= q <= n-2).
template <int K>
__global__ void __launch_bounds__(kStageNT) stage_conv_bwd_input_kernel(const float* __restrict__ gy, const float* __restrict__ w,
                                                                        int B, int H, int W, int reflect, float* __restrict__ gx) {
    constexpr int R = K / 2;
    __shared__ float sw[9 * K * K];
    for (int i = threadIdx.x; i < 9 * K * K; i += kStageNT) sw[i] = w[i];
    __syncthreads();
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    for (size_t idx = (size_t)blockIdx.x * kStageNT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kStageNT) {
        const int b = (int)(idx / plane);
        const int q = (int)(idx - (size_t)b * plane);
        const int qy = q / W, qx = q - qy * W;
        int iy[3], ix[3], ny = 0, nx = 0;
        iy[ny++] = qy;
        ix[nx++] = qx;
        if (reflect) {
            if (qy >= 1 && qy <= R) iy[ny++] = -qy;
            if (qy >= H - 1 - R && qy <= H - 2) iy[ny++] = 2 * (H - 1) - qy;
            if (qx >= 1 && qx <= R) ix[nx++] = -qx;
            if (qx >= W - 1 - R && qx <= W - 2) ix[nx++] = 2 * (W - 1) - qx;
        }
        const float* gb = gy + (size_t)b * 3 * plane;
        float acc[3] = {0.f, 0.f, 0.f};
        for (int u = 0; u < ny; ++u)
            for (int a = 0; a < K; ++a) {
                const int py = iy[u] - (a - R);
                if (py < 0 || py >= H) continue;
                for (int v = 0; v < nx; ++v)
                    for (int c = 0; c < K; ++c) {
                        const int px = ix[v] - (c - R);
                        if (px < 0 || px >= W) continue;
                        const size_t off = (size_t)py * W + px;
#pragma unroll
                        for (int co = 0; co < 3; ++co) {
                            const float g = __ldg(gb + co * plane + off);
#pragma unroll
                            for (int ci = 0; ci < 3; ++ci) acc[ci] = fmaf(sw[((co * 3 + ci) * K + a) * K + c], g, acc[ci]);
                        }
                    }
            }
        float* xb = gx + (size_t)b * 3 * plane + q;
        xb[0] = acc[0]; xb[plane] = acc[1]; xb[2 * plane] = acc[2];
    }
}

// CTA sum of `v` in a fixed order (warp shuffle tree, then warp 0 over the warps); valid in thread 0
__device__ __forceinline__ double block_sum(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    __syncthreads();                                                  // scratch free again
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double s = 0.0;
    if (threadIdx.x == 0)
        for (int i = 0; i < kStageNT / 32; ++i) s += scratch[i];
    return s;
}

// partial[cta][(co*3 + ci)*K*K + a*K + b] = sum over the CTA's pixels of gy[co][p] x_pad[ci][p + (a - r, b - r)]
template <int K>
__global__ void __launch_bounds__(kStageNT) stage_conv_bwd_weight_kernel(const float* __restrict__ x, const float* __restrict__ gy,
                                                                         int B, int H, int W, int reflect,
                                                                         double* __restrict__ partial) {
    constexpr int R = K / 2;
    __shared__ double scratch[kStageNT / 32];
    const size_t plane = (size_t)H * W, total = (size_t)B * plane;
    for (int a = 0; a < K; ++a)
        for (int c = 0; c < K; ++c) {
            double acc[9];
#pragma unroll
            for (int i = 0; i < 9; ++i) acc[i] = 0.0;
            for (size_t idx = (size_t)blockIdx.x * kStageNT + threadIdx.x; idx < total; idx += (size_t)gridDim.x * kStageNT) {
                const int b = (int)(idx / plane);
                const int p = (int)(idx - (size_t)b * plane);
                const int py = p / W, px = p - py * W;
                const int sy = pad_src(py + a - R, H, reflect), sx = pad_src(px + c - R, W, reflect);
                if (sy < 0 || sx < 0) continue;
                const float* xb = x + (size_t)b * 3 * plane + (size_t)sy * W + sx;
                const float* gb = gy + (size_t)b * 3 * plane + p;
                float xv[3], gv[3];
#pragma unroll
                for (int i = 0; i < 3; ++i) { xv[i] = __ldg(xb + i * plane); gv[i] = __ldg(gb + i * plane); }
#pragma unroll
                for (int co = 0; co < 3; ++co)
#pragma unroll
                    for (int ci = 0; ci < 3; ++ci) acc[co * 3 + ci] += (double)gv[co] * (double)xv[ci];
            }
#pragma unroll
            for (int i = 0; i < 9; ++i) {
                const double s = block_sum(acc[i], scratch);
                if (threadIdx.x == 0) partial[(size_t)blockIdx.x * (9 * K * K) + (size_t)i * K * K + a * K + c] = s;
            }
        }
}

// out[e] = sum over the CTAs, in CTA order, of partial[cta][e] * scale
__global__ void stage_finish_kernel(const double* __restrict__ partial, int n_cta, int n, double scale, float* __restrict__ out) {
    for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        double s = 0.0;
        for (int c = 0; c < n_cta; ++c) s += partial[(size_t)c * n + e];
        out[e] = (float)(s * scale);
    }
}

// clip (:206): y = min(max(x, lo), hi); the gradient passes where lo <= x <= hi (torch.clip)
__global__ void stage_clip_kernel(const float* __restrict__ x, size_t n, float lo, float hi, float* __restrict__ y) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = fminf(fmaxf(x[i], lo), hi);
}
__global__ void stage_clip_bwd_kernel(const float* __restrict__ x, const float* __restrict__ gy, size_t n, float lo, float hi,
                                      float* __restrict__ gx) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const float v = x[i];
        gx[i] = (v >= lo && v <= hi) ? gy[i] : 0.f;
    }
}

// gamma (:209): y = exp((1 / gamma) * log(x)) with the reference's own operation order
__global__ void stage_gamma_kernel(const float* __restrict__ x, const float* __restrict__ gamma, size_t n, float* __restrict__ y) {
    const float invg = 1.0f / gamma[0];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
        y[i] = expf(invg * logf(x[i]));
}
// gx = gy * y * (1/gamma) / x;  partial[cta] = sum gy * y * log(x)   (d/dgamma = -1/gamma^2 times the total)
__global__ void __launch_bounds__(kStageNT) stage_gamma_bwd_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                                                   const float* __restrict__ gy, const float* __restrict__ gamma,
                                                                   size_t n, float* __restrict__ gx, double* __restrict__ partial) {
    __shared__ double scratch[kStageNT / 32];
    const float invg = 1.0f / gamma[0];
    double acc = 0.0;
    for (size_t i = (size_t)blockIdx.x * kStageNT + threadIdx.x; i < n; i += (size_t)gridDim.x * kStageNT) {
        const float xv = x[i], yv = y[i], g = gy[i];
        if (gx) gx[i] = g * yv * invg / xv;
        acc += (double)(g * yv) * (double)logf(xv);
    }
    const double s = block_sum(acc, scratch);
    if (threadIdx.x == 0) partial[blockIdx.x] = s;
}

// grad_gamma = -(sum over the CTAs, in CTA order) / gamma^2
__global__ void stage_gamma_finish_kernel(const double* __restrict__ partial, int n_cta, const float* __restrict__ gamma,
                                          float* __restrict__ out) {
    if (blockIdx.x == 0 && threadIdx.x == 0) {
        double s = 0.0;
        for (int c = 0; c < n_cta; ++c) s += partial[c];
        const double g = (double)gamma[0];
        out[0] = (float)(-s / (g * g));
    }
}

static int stage_grid(size_t work_items) {
    size_t g = (work_items + kStageNT - 1) / kStageNT;
    if (g < 1) g = 1;
    if (g > (size_t)kStageMaxCtas) g = kStageMaxCtas;
    return (int)g;
}

static int check_launch() {
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

}  // namespace r2l

using namespace r2l;

extern "C" {

size_t r2l_isp_stage_workspace_bytes(int K) {
    if (K != 3 && K != 5) K = 5;
    return (size_t)kStageMaxCtas * 9 * K * K * sizeof(double);
}

int r2l_isp_stage_conv(const float* x, const float* weight, int B, int H, int W, int K, int pad_mode, float* y, void* stream) {
    if (B < 0 || H < 1 || W < 1) return R2L_ERR_BAD_SHAPE;
    if ((K != 3 && K != 5) || (pad_mode != 0 && pad_mode != 1)) return R2L_ERR_BAD_ARGUMENT;
    if (pad_mode == 1 && (H <= K / 2 || W <= K / 2)) return R2L_ERR_BAD_SHAPE;          // torch's reflect pad raises
    if (B == 0) return R2L_OK;
    if (!x || !weight || !y) return R2L_ERR_NULL_POINTER;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int g = stage_grid((size_t)B * H * W);
    if (K == 3) stage_conv_kernel<3><<<g, kStageNT, 0, st>>>(x, weight, B, H, W, pad_mode, y);
    else stage_conv_kernel<5><<<g, kStageNT, 0, st>>>(x, weight, B, H, W, pad_mode, y);
    return check_launch();
}

int r2l_isp_stage_conv_backward(const float* x, const float* weight, const float* grad_y, int B, int H, int W, int K,
                                int pad_mode, float* grad_x, float* grad_weight, void* workspace, size_t workspace_bytes,
                                void* stream) {
    if (B < 0 || H < 1 || W < 1) return R2L_ERR_BAD_SHAPE;
    if ((K != 3 && K != 5) || (pad_mode != 0 && pad_mode != 1)) return R2L_ERR_BAD_ARGUMENT;
    if (pad_mode == 1 && (H <= K / 2 || W <= K / 2)) return R2L_ERR_BAD_SHAPE;
    if (!grad_y || (grad_x && !weight) || (grad_weight && (!x || !workspace))) return R2L_ERR_NULL_POINTER;
    if (grad_weight && workspace_bytes < r2l_isp_stage_workspace_bytes(K)) return R2L_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0) {
        if (grad_weight) {
            cudaError_t e = cudaMemsetAsync(grad_weight, 0, sizeof(float) * 9 * K * K, st);
            if (e != cudaSuccess) return cuda_fail(e);
        }
        return R2L_OK;
    }
    const int g = stage_grid((size_t)B * H * W);
    if (grad_x) {
        if (K == 3) stage_conv_bwd_input_kernel<3><<<g, kStageNT, 0, st>>>(grad_y, weight, B, H, W, pad_mode, grad_x);
        else stage_conv_bwd_input_kernel<5><<<g, kStageNT, 0, st>>>(grad_y, weight, B, H, W, pad_mode, grad_x);
        int rc = check_launch();
        if (rc != R2L_OK) return rc;
    }
    if (grad_weight) {
        double* partial = static_cast<double*>(workspace);
        if (K == 3) stage_conv_bwd_weight_kernel<3><<<g, kStageNT, 0, st>>>(x, grad_y, B, H, W, pad_mode, partial);
        else stage_conv_bwd_weight_kernel<5><<<g, kStageNT, 0, st>>>(x, grad_y, B, H, W, pad_mode, partial);
        int rc = check_launch();
        if (rc != R2L_OK) return rc;
        stage_finish_kernel<<<1, 256, 0, st>>>(partial, g, 9 * K * K, 1.0, grad_weight);
        rc = check_launch();
        if (rc != R2L_OK) return rc;
    }
    return R2L_OK;
}

int r2l_isp_stage_clip(const float* x, long long n, float lo, float hi, float* y, void* stream) {
    if (n < 0) return R2L_ERR_BAD_SHAPE;
    if (n == 0) return R2L_OK;
    if (!x || !y) return R2L_ERR_NULL_POINTER;
    stage_clip_kernel<<<stage_grid((size_t)n), kStageNT, 0, static_cast<cudaStream_t>(stream)>>>(x, (size_t)n, lo, hi, y);
    return check_launch();
}

int r2l_isp_stage_clip_backward(const float* x, const float* grad_y, long long n, float lo, float hi, float* grad_x,
                                void* stream) {
    if (n < 0) return R2L_ERR_BAD_SHAPE;
    if (n == 0) return R2L_OK;
    if (!x || !grad_y || !grad_x) return R2L_ERR_NULL_POINTER;
    stage_clip_bwd_kernel<<<stage_grid((size_t)n), kStageNT, 0, static_cast<cudaStream_t>(stream)>>>(x, grad_y, (size_t)n, lo, hi,
                                                                                                       grad_x);
    return check_launch();
}

int r2l_isp_stage_gamma(const float* x, const float* gamma, long long n, float* y, void* stream) {
    if (n < 0) return R2L_ERR_BAD_SHAPE;
    if (n == 0) return R2L_OK;
    if (!x || !gamma || !y) return R2L_ERR_NULL_POINTER;
    stage_gamma_kernel<<<stage_grid((size_t)n), kStageNT, 0, static_cast<cudaStream_t>(stream)>>>(x, gamma, (size_t)n, y);
    return check_launch();
}

int r2l_isp_stage_gamma_backward(const float* x, const float* y, const float* grad_y, const float* gamma, long long n,
                                 float* grad_x, float* grad_gamma, void* workspace, size_t workspace_bytes, void* stream) {
    if (n < 0) return R2L_ERR_BAD_SHAPE;
    if (!x || !y || !grad_y || !gamma || !grad_gamma || !workspace) return R2L_ERR_NULL_POINTER;
    if (workspace_bytes < (size_t)kStageMaxCtas * sizeof(double)) return R2L_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (n == 0) {
        cudaError_t e = cudaMemsetAsync(grad_gamma, 0, sizeof(float), st);
        return e == cudaSuccess ? R2L_OK : cuda_fail(e);
    }
    double* partial = static_cast<double*>(workspace);
    const int g = stage_grid((size_t)n);
    stage_gamma_bwd_kernel<<<g, kStageNT, 0, st>>>(x, y, grad_y, gamma, (size_t)n, grad_x, partial);
    int rc = check_launch();
    if (rc != R2L_OK) return rc;
    stage_gamma_finish_kernel<<<1, 32, 0, st>>>(partial, g, gamma, grad_gamma);   // d/dgamma exp(log(x)/gamma) = -y log(x)/gamma^2
    return check_launch();
}

}  // extern "C"
