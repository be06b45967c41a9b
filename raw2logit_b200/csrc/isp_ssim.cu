// isp_ssim.cu -- fused SSIM (forward value + backward image gradients) for the adversarial regulariser.
//
// Reference: utils/ssim.py:19-39 (`_ssim`; window from `create_window` :13-17, 1-D Gaussian :9-11), used as
// `SSIM(window_size=11)` by train.py:261-262 through AuxLoss (utils/base.py:346-358).  The reference runs five grouped
// 11x11 convolutions (img1, img2, img1^2, img2^2, img1*img2; zero padding 5) plus ~15 elementwise kernels and keeps every
// intermediate for autograd.  Here:
//   forward  -- one kernel: a CTA stages a 42 x 42 window of both images (one (b, c) plane, 32 x 32 outputs) in shared
//               memory, runs the separable Gaussian (the 2-D window is the outer product of the 1-D one, :14-15) over
//               the five moment planes, forms the SSIM map and reduces it in double; per-CTA partial sums -> the mean;
//   backward -- one kernel: recomputes the moments on the 42 x 42 positions around the tile (52 x 52 inputs), forms the
//               derivative maps of the SSIM map with respect to (mu, E[x^2], E[x1 x2]), pushes them back through the
//               (symmetric, zero-padded) window with a second separable pass and combines them with the pixel values:
//                 d/dx2(q) = conv(df/dmu2)(q) + 2 x2(q) conv(df/dE22)(q) + x1(q) conv(df/dE12)(q)      (x1 alike).
// Nothing is saved between the two; algorithmic traffic 8 B/px per plane forward, 8 + 4 (or 8) B/px backward.
// HBM-bound stencil, fp32 arithmetic; no tensor cores (a K = 11 contraction at fp32 accuracy).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "isp_launch.h"

namespace r2l {

constexpr int kSsimWin = 11, kSsimR = 5;
constexpr int kSsimT = 32;                          // outputs per tile side
constexpr int kSsimNT = 256;
constexpr float kSsimC1 = 0.01f * 0.01f, kSsimC2 = 0.03f * 0.03f;     // ssim.py:31-32

struct SsimWindow { float w[kSsimWin]; };

// 1-D window exactly as the reference builds it: exp(-(x - 5)^2 / (2 sigma^2)) in double (Python floats), stored as
// fp32 (torch.Tensor), normalised by the fp32 sum (ssim.py:9-11)
static SsimWindow make_window() {
    SsimWindow win;
    float g[kSsimWin];
    float sum = 0.f;
    for (int x = 0; x < kSsimWin; ++x) {
        g[x] = (float)exp(-(double)((x - kSsimWin / 2) * (x - kSsimWin / 2)) / (2.0 * 1.5 * 1.5));
        sum += g[x];
    }
    for (int x = 0; x < kSsimWin; ++x) win.w[x] = g[x] / sum;
    return win;
}

// ---- shared helpers -------------------------------------------------------------------------------------------------
// loads an (N x N) window of plane `p` (H x W) whose top-left corner is (y0, x0) into s[N][N + 1]; zero outside the image
template <int N>
__device__ __forceinline__ void load_window(const float* __restrict__ p, int H, int W, int y0, int x0, float (*s)[N + 1]) {
    for (int i = threadIdx.x; i < N * N; i += kSsimNT) {
        const int r = i / N, c = i - r * N;
        const int gy = y0 + r, gx = x0 + c;
        s[r][c] = (gy >= 0 && gy < H && gx >= 0 && gx < W) ? __ldg(p + (size_t)gy * W + gx) : 0.f;
    }
}

// moments at one position from the horizontally filtered planes h[5][rows][cols]: vertical 11-tap pass
struct Moments { float mu1, mu2, e11, e22, e12; };

// SSIM value and its partial derivatives at one position (ssim.py:23-34)
struct SsimPoint { float f, d_mu1, d_mu2, d_e, d_e12; };      // d_e = df/dE11 = df/dE22
__device__ __forceinline__ SsimPoint ssim_point(const Moments& m) {
    const float mu1s = m.mu1 * m.mu1, mu2s = m.mu2 * m.mu2, mu12 = m.mu1 * m.mu2;
    const float s1 = m.e11 - mu1s, s2 = m.e22 - mu2s, s12 = m.e12 - mu12;
    const float a = 2.f * mu12 + kSsimC1, b = 2.f * s12 + kSsimC2;
    const float c = mu1s + mu2s + kSsimC1, d = s1 + s2 + kSsimC2;
    const float icd = 1.f / (c * d);
    SsimPoint r;
    r.f = a * b * icd;
    // independent variables (mu1, mu2, E11, E22, E12): s1 = E11 - mu1^2, s2 = E22 - mu2^2, s12 = E12 - mu1 mu2
    r.d_e12 = 2.f * a * icd;
    r.d_e = -r.f / d;
    const float t = 2.f * (b - a) * icd, u = 2.f * r.f * (1.f / c - 1.f / d);
    r.d_mu2 = m.mu1 * t - m.mu2 * u;
    r.d_mu1 = m.mu2 * t - m.mu1 * u;
    return r;
}

// ---- forward --------------------------------------------------------------------------------------------------------
// grid: (tiles_x * tiles_y, B * C); partial[plane * tiles + tile] = sum of the SSIM map over the tile (double)
__global__ void __launch_bounds__(kSsimNT) ssim_forward_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                               int H, int W, int tiles_x, SsimWindow win,
                                                               double* __restrict__ partial) {
    constexpr int N = kSsimT + 2 * kSsimR;                       // 42
    __shared__ float x1[N][N + 1], x2[N][N + 1];
    __shared__ float h[5][N][kSsimT + 1];                        // horizontally filtered moments
    __shared__ double red[kSsimNT / 32];
    const int tile = blockIdx.x, plane = blockIdx.y;
    const int ty0 = (tile / tiles_x) * kSsimT, tx0 = (tile % tiles_x) * kSsimT;
    const float* p1 = img1 + (size_t)plane * H * W;
    const float* p2 = img2 + (size_t)plane * H * W;
    load_window<N>(p1, H, W, ty0 - kSsimR, tx0 - kSsimR, x1);
    load_window<N>(p2, H, W, ty0 - kSsimR, tx0 - kSsimR, x2);
    __syncthreads();
    for (int i = threadIdx.x; i < N * kSsimT; i += kSsimNT) {
        const int r = i / kSsimT, c = i - r * kSsimT;
        float a1 = 0.f, a2 = 0.f, a11 = 0.f, a22 = 0.f, a12 = 0.f;
#pragma unroll
        for (int t = 0; t < kSsimWin; ++t) {
            const float u = x1[r][c + t], v = x2[r][c + t], w = win.w[t];
            a1 = fmaf(w, u, a1); a2 = fmaf(w, v, a2);
            a11 = fmaf(w, u * u, a11); a22 = fmaf(w, v * v, a22); a12 = fmaf(w, u * v, a12);
        }
        h[0][r][c] = a1; h[1][r][c] = a2; h[2][r][c] = a11; h[3][r][c] = a22; h[4][r][c] = a12;
    }
    __syncthreads();
    double local = 0.0;
    for (int i = threadIdx.x; i < kSsimT * kSsimT; i += kSsimNT) {
        const int r = i / kSsimT, c = i - r * kSsimT;
        if (ty0 + r >= H || tx0 + c >= W) continue;
        float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < kSsimWin; ++t) {
            const float w = win.w[t];
#pragma unroll
            for (int k = 0; k < 5; ++k) v[k] = fmaf(w, h[k][r + t][c], v[k]);
        }
        Moments m{v[0], v[1], v[2], v[3], v[4]};
        local += (double)ssim_point(m).f;
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) local += __shfl_xor_sync(0xffffffffu, local, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int w = 0; w < kSsimNT / 32; ++w) s += red[w];        // fixed order: bit-reproducible
        partial[(size_t)plane * gridDim.x + tile] = s;
    }
}

// ---- backward -------------------------------------------------------------------------------------------------------
// scale[b]: upstream gradient of image b's mean divided by the number of averaged elements (host tensor op);
// grad1 / grad2 may be null (only the requested gradients are formed)
template <bool G1, bool G2>
__global__ void __launch_bounds__(kSsimNT) ssim_backward_kernel(const float* __restrict__ img1, const float* __restrict__ img2,
                                                                const float* __restrict__ scale, int C, int H, int W,
                                                                int tiles_x, SsimWindow win, float* __restrict__ grad1,
                                                                float* __restrict__ grad2) {
    constexpr int N = kSsimT + 4 * kSsimR;                       // 52: inputs
    constexpr int M = kSsimT + 2 * kSsimR;                       // 42: positions of the SSIM map that reach the tile
    constexpr int NMAP = 2 + (G1 ? 1 : 0) + (G2 ? 1 : 0);        // d_e, d_e12, [d_mu1], [d_mu2]
    extern __shared__ float smem[];
    float (*x1)[N + 1] = reinterpret_cast<float (*)[N + 1]>(smem);
    float (*x2)[N + 1] = x1 + N;
    float* hbuf = reinterpret_cast<float*>(x2 + N);              // [5][N][M + 1], later [NMAP][M][T + 1]
    float* maps = hbuf + 5 * N * (M + 1);                        // [NMAP][M][M + 1]
    auto H5 = [&](int k, int r, int c) -> float& { return hbuf[(k * N + r) * (M + 1) + c]; };
    auto MAP = [&](int k, int r, int c) -> float& { return maps[(k * M + r) * (M + 1) + c]; };
    auto HM = [&](int k, int r, int c) -> float& { return hbuf[(k * M + r) * (kSsimT + 1) + c]; };
    const int tile = blockIdx.x, plane = blockIdx.y;
    const int ty0 = (tile / tiles_x) * kSsimT, tx0 = (tile % tiles_x) * kSsimT;
    const float* p1 = img1 + (size_t)plane * H * W;
    const float* p2 = img2 + (size_t)plane * H * W;
    load_window<N>(p1, H, W, ty0 - 2 * kSsimR, tx0 - 2 * kSsimR, x1);
    load_window<N>(p2, H, W, ty0 - 2 * kSsimR, tx0 - 2 * kSsimR, x2);
    __syncthreads();
    // moments, horizontal pass: rows 0..N-1 of the input window, map columns 0..M-1 (input columns c .. c+10)
    for (int i = threadIdx.x; i < N * M; i += kSsimNT) {
        const int r = i / M, c = i - r * M;
        float a1 = 0.f, a2 = 0.f, a11 = 0.f, a22 = 0.f, a12 = 0.f;
#pragma unroll
        for (int t = 0; t < kSsimWin; ++t) {
            const float u = x1[r][c + t], v = x2[r][c + t], w = win.w[t];
            a1 = fmaf(w, u, a1); a2 = fmaf(w, v, a2);
            a11 = fmaf(w, u * u, a11); a22 = fmaf(w, v * v, a22); a12 = fmaf(w, u * v, a12);
        }
        H5(0, r, c) = a1; H5(1, r, c) = a2; H5(2, r, c) = a11; H5(3, r, c) = a22; H5(4, r, c) = a12;
    }
    __syncthreads();
    // vertical pass + derivative maps at the M x M positions (ty0 - 5 + r, tx0 - 5 + c); zero outside the image: the
    // SSIM map only exists there (the adjoint of a zero-padded convolution reads zeros beyond the border)
    for (int i = threadIdx.x; i < M * M; i += kSsimNT) {
        const int r = i / M, c = i - r * M;
        const int gy = ty0 - kSsimR + r, gx = tx0 - kSsimR + c;
        SsimPoint pt{0.f, 0.f, 0.f, 0.f, 0.f};
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            float v[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
            for (int t = 0; t < kSsimWin; ++t) {
                const float w = win.w[t];
#pragma unroll
                for (int k = 0; k < 5; ++k) v[k] = fmaf(w, H5(k, r + t, c), v[k]);
            }
            Moments m{v[0], v[1], v[2], v[3], v[4]};
            pt = ssim_point(m);
        }
        MAP(0, r, c) = pt.d_e; MAP(1, r, c) = pt.d_e12;
        if (G1) MAP(2, r, c) = pt.d_mu1;
        if (G2) MAP(2 + (G1 ? 1 : 0), r, c) = pt.d_mu2;
    }
    __syncthreads();
    // the maps through the window: horizontal pass (rows 0..M-1, tile columns 0..T-1) ...
    for (int i = threadIdx.x; i < M * kSsimT; i += kSsimNT) {
        const int r = i / kSsimT, c = i - r * kSsimT;
        float acc[NMAP];
#pragma unroll
        for (int k = 0; k < NMAP; ++k) acc[k] = 0.f;
#pragma unroll
        for (int t = 0; t < kSsimWin; ++t) {
            const float w = win.w[t];
#pragma unroll
            for (int k = 0; k < NMAP; ++k) acc[k] = fmaf(w, MAP(k, r, c + t), acc[k]);
        }
#pragma unroll
        for (int k = 0; k < NMAP; ++k) HM(k, r, c) = acc[k];
    }
    __syncthreads();
    // ... vertical pass and the combination with the pixel values
    const float sc = scale[plane / C];
    for (int i = threadIdx.x; i < kSsimT * kSsimT; i += kSsimNT) {
        const int r = i / kSsimT, c = i - r * kSsimT;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy >= H || gx >= W) continue;
        float acc[NMAP];
#pragma unroll
        for (int k = 0; k < NMAP; ++k) acc[k] = 0.f;
#pragma unroll
        for (int t = 0; t < kSsimWin; ++t) {
            const float w = win.w[t];
#pragma unroll
            for (int k = 0; k < NMAP; ++k) acc[k] = fmaf(w, HM(k, r + t, c), acc[k]);
        }
        const float u = x1[r + 2 * kSsimR][c + 2 * kSsimR], v = x2[r + 2 * kSsimR][c + 2 * kSsimR];
        const size_t o = (size_t)plane * H * W + (size_t)gy * W + gx;
        if (G1) grad1[o] = sc * (acc[2] + 2.f * u * acc[0] + v * acc[1]);
        if (G2) grad2[o] = sc * (acc[2 + (G1 ? 1 : 0)] + 2.f * v * acc[0] + u * acc[1]);
    }
}

template <bool G1, bool G2> constexpr size_t ssim_bwd_smem() {
    constexpr int N = kSsimT + 4 * kSsimR, M = kSsimT + 2 * kSsimR, NMAP = 2 + (G1 ? 1 : 0) + (G2 ? 1 : 0);
    return sizeof(float) * ((size_t)2 * N * (N + 1) + (size_t)5 * N * (M + 1) + (size_t)NMAP * M * (M + 1));
}

template <bool G1, bool G2>
static int launch_ssim_backward(const float* img1, const float* img2, const float* scale, int B, int C, int H, int W,
                                float* grad1, float* grad2, cudaStream_t st) {
    const int tiles_x = (W + kSsimT - 1) / kSsimT, tiles_y = (H + kSsimT - 1) / kSsimT;
    constexpr size_t smem = ssim_bwd_smem<G1, G2>();
    static bool configured = false;                            // (idempotent; a race sets the same value twice)
    if (!configured) {
        cudaError_t e = cudaFuncSetAttribute(ssim_backward_kernel<G1, G2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return cuda_fail(e);
        configured = true;
    }
    ssim_backward_kernel<G1, G2><<<dim3(tiles_x * tiles_y, B * C), kSsimNT, smem, st>>>(img1, img2, scale, C, H, W, tiles_x,
                                                                                    make_window(), grad1, grad2);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

}  // namespace r2l

using namespace r2l;

extern "C" {

size_t r2l_isp_ssim_partial_count(int B, int C, int H, int W) {
    if (B < 0 || C < 0 || H <= 0 || W <= 0) return 0;
    return (size_t)B * C * ((H + kSsimT - 1) / kSsimT) * ((W + kSsimT - 1) / kSsimT);
}

int r2l_isp_ssim_forward(const float* img1, const float* img2, int B, int C, int H, int W, int window_size, double* partial,
                     void* stream) {
    if (window_size != kSsimWin) return R2L_ERR_BAD_ARGUMENT;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0) return R2L_ERR_BAD_SHAPE;
    if (B == 0) return R2L_OK;
    if (!img1 || !img2 || !partial) return R2L_ERR_NULL_POINTER;
    if ((long long)B * C > 65535) return R2L_ERR_BAD_SHAPE;
    const int tiles_x = (W + kSsimT - 1) / kSsimT, tiles_y = (H + kSsimT - 1) / kSsimT;
    ssim_forward_kernel<<<dim3(tiles_x * tiles_y, B * C), kSsimNT, 0, static_cast<cudaStream_t>(stream)>>>(
        img1, img2, H, W, tiles_x, make_window(), partial);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

int r2l_isp_ssim_backward(const float* img1, const float* img2, const float* scale, int B, int C, int H, int W,
                      int window_size, float* grad1, float* grad2, void* stream) {
    if (window_size != kSsimWin) return R2L_ERR_BAD_ARGUMENT;
    if (B < 0 || C <= 0 || H <= 0 || W <= 0) return R2L_ERR_BAD_SHAPE;
    if (B == 0 || (!grad1 && !grad2)) return R2L_OK;
    if (!img1 || !img2 || !scale) return R2L_ERR_NULL_POINTER;
    if ((long long)B * C > 65535) return R2L_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (grad1 && grad2) return launch_ssim_backward<true, true>(img1, img2, scale, B, C, H, W, grad1, grad2, st);
    if (grad1) return launch_ssim_backward<true, false>(img1, img2, scale, B, C, H, W, grad1, nullptr, st);
    return launch_ssim_backward<false, true>(img1, img2, scale, B, C, H, W, nullptr, grad2, st);
}

}  // extern "C"
