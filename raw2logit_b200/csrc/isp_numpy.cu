// isp_numpy.cu -- the reference's STATIC numpy pipeline as one fused kernel (numpy-compatible boundary mode).
//
// Reference: processing/pipeline_numpy.py:70-141 `processing(...)` with the train.py defaults debayer='bilinear',
// sharpening='sharpening_filter', denoising='gaussian_denoising' (train.py:95-100), the chain 16 DataLoader workers run
// per image in --processing_mode static (train.py:316-320).  It differs from the torch chain of the fused ISP only at
// the borders and at the clip (pipeline_torch.py:233 notes the mismatch), so it cannot be served by that kernel:
//   * bilinear demosaic = per-channel masked CFA planes convolved with H_RB / H_G by scipy.ndimage.convolve, whose
//     default boundary is the HALF-sample reflection (-1 -> 0, H -> H-1): the mirrored site brings its own CFA phase
//     (colour-demosaicing 0.1.6 `demosaicing_CFA_Bayer_bilinear`, call site :93);
//   * sharpening: scipy.signal.convolve2d(..., 'same', boundary='fill', fillvalue=0) on Y (:180-191);
//   * Gaussian: scipy.ndimage.gaussian_filter(Y, 0.5): separable, radius int(4 * 0.5 + 0.5) = 2, half-sample reflection of
//     the SHARPENED plane (:203-209);
//   * clip to [0, 1] (not [1e-5, 1]) and x ** (1 / gamma) (:138-139, :241-244).
// One CTA = one 32 x 32 tile of one image: raw window (40 x 40, mirrored on the way in) -> Y0 / U / V -> Y1 -> output,
// all in shared memory; 4 (or 2) B/px read, 12 B/px written.  Forward only (the numpy chain has no gradient).
#include <cuda_runtime.h>
#include <math.h>
#include <stdint.h>

#include "isp_launch.h"

namespace r2l {

constexpr int kNpT = 32, kNpNT = 256;

struct NumpyParams {
    float bl[4], wb[3], ccm[9];
    float m1[9], m2[9];           // rgb -> yuv (skimage yuv_from_rgb) and its inverse
    float g5[5];                  // gaussian_filter weights (sigma, radius 2)
    float inv_gamma;
    int sharpen, blur;
};

__device__ __forceinline__ int mirror_half(int i, int n) {      // scipy 'reflect': (d c b a | a b c d | d c b a)
    if (n == 1) return 0;
    const int p = 2 * n;
    i %= p;
    if (i < 0) i += p;
    return i < n ? i : p - 1 - i;
}

template <typename RawT> __device__ __forceinline__ float raw_value(const RawT* p, float denom);
template <> __device__ __forceinline__ float raw_value<float>(const float* p, float) { return __ldg(p); }
template <> __device__ __forceinline__ float raw_value<uint16_t>(const uint16_t* p, float denom) {
    return __fdiv_rn((float)__ldg(p), denom);                       // dataset.py:87: img / (2**bits - 1)
}

template <typename RawT>
__global__ void __launch_bounds__(kNpNT) isp_numpy_forward_kernel(const RawT* __restrict__ raw, float denom, int H, int W,
                                                                  int tiles_x, int tiles_per_image, NumpyParams P,
                                                                  float* __restrict__ out) {
    constexpr int RW = kNpT + 8, YW = kNpT + 6, SW = kNpT + 4;      // raw window, Y0 window, Y1 window
    __shared__ float s_raw[RW][RW + 1];
    __shared__ unsigned char s_par[RW][RW + 3];                    // CFA phase of the (mirrored) site each value came from
    __shared__ float s_y0[YW][YW + 1];
    __shared__ float s_u[kNpT][kNpT + 1], s_v[kNpT][kNpT + 1];
    __shared__ float s_y1[SW][SW + 1];
    const int b = blockIdx.x / tiles_per_image, tile = blockIdx.x - b * tiles_per_image;
    const int ty0 = (tile / tiles_x) * kNpT, tx0 = (tile % tiles_x) * kNpT;
    const RawT* img = raw + (size_t)b * H * W;
    // raw window, black level removed (:152-158), half-sample mirrored: the masked CFA planes are mirrored, so a pad site
    // carries the value AND the colour of the site it mirrors
    for (int i = threadIdx.x; i < RW * RW; i += kNpNT) {
        const int r = i / RW, c = i - r * RW;
        const int gy = mirror_half(ty0 - 4 + r, H), gx = mirror_half(tx0 - 4 + c, W);
        const int par = 2 * (gy & 1) + (gx & 1);
        s_raw[r][c] = raw_value<RawT>(img + (size_t)gy * W + gx, denom) - P.bl[par];
        s_par[r][c] = (unsigned char)par;
    }
    __syncthreads();
    // demosaic (H_RB = [1 2 1; 2 4 2; 1 2 1] / 4 on the R / B planes, H_G = [0 1 0; 1 4 1; 0 1 0] / 4 on G), white
    // balance (:161-162), colour matrix (:165-167), RGB -> YUV (:184): Y0 on the tile +-3 (zero outside the image: the
    // sharpening filter zero-fills), U / V on the tile
    for (int i = threadIdx.x; i < YW * YW; i += kNpNT) {
        const int r = i / YW, c = i - r * YW;
        const int gy = ty0 - 3 + r, gx = tx0 - 3 + c;
        float y0 = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            float rgb[3] = {0.f, 0.f, 0.f};
#pragma unroll
            for (int dy = -1; dy <= 1; ++dy)
#pragma unroll
                for (int dx = -1; dx <= 1; ++dx) {
                    const float v = s_raw[r + 1 + dy][c + 1 + dx];
                    const int par = s_par[r + 1 + dy][c + 1 + dx];
                    const float wrb = (dy == 0 ? 2.f : 1.f) * (dx == 0 ? 2.f : 1.f) * 0.25f;
                    const float wg = (dy == 0 && dx == 0) ? 1.f : ((dy == 0 || dx == 0) ? 0.25f : 0.f);
                    if (par == 0) rgb[0] = fmaf(wrb, v, rgb[0]);
                    else if (par == 3) rgb[2] = fmaf(wrb, v, rgb[2]);
                    else rgb[1] = fmaf(wg, v, rgb[1]);
                }
            const float w0 = rgb[0] * P.wb[0], w1 = rgb[1] * P.wb[1], w2 = rgb[2] * P.wb[2];
            float cc[3], yuv[3];
#pragma unroll
            for (int k = 0; k < 3; ++k) cc[k] = P.ccm[k * 3] * w0 + P.ccm[k * 3 + 1] * w1 + P.ccm[k * 3 + 2] * w2;
#pragma unroll
            for (int k = 0; k < 3; ++k) yuv[k] = P.m1[k * 3] * cc[0] + P.m1[k * 3 + 1] * cc[1] + P.m1[k * 3 + 2] * cc[2];
            y0 = yuv[0];
            const int tr = r - 3, tc = c - 3;
            if (tr >= 0 && tr < kNpT && tc >= 0 && tc < kNpT) { s_u[tr][tc] = yuv[1]; s_v[tr][tc] = yuv[2]; }
        }
        s_y0[r][c] = y0;
    }
    __syncthreads();
    // sharpening filter [0 -1 0; -1 5 -1; 0 -1 0] with zero fill (:178-191) on the tile +-2, in-image sites only
    for (int i = threadIdx.x; i < SW * SW; i += kNpNT) {
        const int r = i / SW, c = i - r * SW;
        const int gy = ty0 - 2 + r, gx = tx0 - 2 + c;
        float y1 = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
            const float ctr = s_y0[r + 1][c + 1];
            y1 = P.sharpen ? 5.f * ctr - s_y0[r][c + 1] - s_y0[r + 2][c + 1] - s_y0[r + 1][c] - s_y0[r + 1][c + 2] : ctr;
        }
        s_y1[r][c] = y1;
    }
    __syncthreads();
    // Gaussian (half-sample reflection of the sharpened plane), YUV -> RGB, clip, gamma
    for (int i = threadIdx.x; i < kNpT * kNpT; i += kNpNT) {
        const int r = i / kNpT, c = i - r * kNpT;
        const int gy = ty0 + r, gx = tx0 + c;
        if (gy >= H || gx >= W) continue;
        float y2;
        if (P.blur) {
            // scipy filters the rows first (axis 0), then the columns; both passes in fp32 here
            float acc = 0.f;
#pragma unroll
            for (int dx = -2; dx <= 2; ++dx) {
                const int cx = mirror_half(gx + dx, W) - (tx0 - 2);
                float col = 0.f;
#pragma unroll
                for (int dy = -2; dy <= 2; ++dy) {
                    const int cy = mirror_half(gy + dy, H) - (ty0 - 2);
                    col = fmaf(P.g5[dy + 2], s_y1[cy][cx], col);
                }
                acc = fmaf(P.g5[dx + 2], col, acc);
            }
            y2 = acc;
        } else {
            y2 = s_y1[r + 2][c + 2];
        }
        const float u = s_u[r][c], v = s_v[r][c];
        const size_t plane = (size_t)H * W;
        float* o = out + (size_t)b * 3 * plane + (size_t)gy * W + gx;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            float x = P.m2[k * 3] * y2 + P.m2[k * 3 + 1] * u + P.m2[k * 3 + 2] * v;
            x = fminf(fmaxf(x, 0.f), 1.f);                           // np.clip(img, 0, 1)  :138
            o[k * plane] = x > 0.f ? exp2f(P.inv_gamma * log2f(x)) : 0.f;   // img ** (1 / gamma)  :241-244
        }
    }
}

}  // namespace r2l

using namespace r2l;

extern "C" int r2l_isp_numpy_forward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                                     const float* black_level, const float* white_balance, const float* colour_matrix,
                                     int sharpening_filter, int gaussian_denoising, float gaussian_sigma, float gamma,
                                     float* out, void* stream) {
    if (raw_dtype != R2L_F32 && raw_dtype != R2L_U16) return R2L_ERR_BAD_DTYPE;
    if (B < 0 || H < 1 || W < 1) return R2L_ERR_BAD_SHAPE;
    if (B == 0) return R2L_OK;
    if (!raw || !out || !black_level || !white_balance || !colour_matrix) return R2L_ERR_NULL_POINTER;   // HOST arrays
    if (!(gamma > 0.f) || (gaussian_denoising && !(gaussian_sigma > 0.f))) return R2L_ERR_BAD_ARGUMENT;
    NumpyParams P;
    for (int i = 0; i < 4; ++i) P.bl[i] = black_level[i];
    for (int i = 0; i < 3; ++i) P.wb[i] = white_balance[i];
    for (int i = 0; i < 9; ++i) P.ccm[i] = colour_matrix[i];
    // skimage.color.colorconv.yuv_from_rgb and its inverse (call sites :184,189,203,207), inverted in double
    const double m1[9] = {0.299, 0.587, 0.114, -0.14714119, -0.28886916, 0.43601035, 0.61497538, -0.51496512, -0.10001026};
    const double det = m1[0] * (m1[4] * m1[8] - m1[5] * m1[7]) - m1[1] * (m1[3] * m1[8] - m1[5] * m1[6]) +
                       m1[2] * (m1[3] * m1[7] - m1[4] * m1[6]);
    const double inv[9] = {(m1[4] * m1[8] - m1[5] * m1[7]) / det, (m1[2] * m1[7] - m1[1] * m1[8]) / det, (m1[1] * m1[5] - m1[2] * m1[4]) / det,
                           (m1[5] * m1[6] - m1[3] * m1[8]) / det, (m1[0] * m1[8] - m1[2] * m1[6]) / det, (m1[2] * m1[3] - m1[0] * m1[5]) / det,
                           (m1[3] * m1[7] - m1[4] * m1[6]) / det, (m1[1] * m1[6] - m1[0] * m1[7]) / det, (m1[0] * m1[4] - m1[1] * m1[3]) / det};
    for (int i = 0; i < 9; ++i) { P.m1[i] = (float)m1[i]; P.m2[i] = (float)inv[i]; }
    // scipy.ndimage.gaussian_filter: radius = int(truncate * sigma + 0.5) with truncate = 4; this kernel holds radius 2
    const int radius = gaussian_denoising ? (int)(4.0 * (double)gaussian_sigma + 0.5) : 0;
    if (radius > 2) return R2L_ERR_BAD_ARGUMENT;
    double g[5] = {0, 0, 1, 0, 0}, sum = 0.0;
    for (int x = -2; x <= 2; ++x) g[x + 2] = (x < -radius || x > radius) ? 0.0 : exp(-0.5 * x * x / ((double)gaussian_sigma * gaussian_sigma));
    for (int i = 0; i < 5; ++i) sum += g[i];
    for (int i = 0; i < 5; ++i) P.g5[i] = (float)(g[i] / sum);
    P.inv_gamma = 1.0f / gamma;
    P.sharpen = sharpening_filter ? 1 : 0;
    P.blur = gaussian_denoising ? 1 : 0;
    const int tiles_x = (W + kNpT - 1) / kNpT, tiles_y = (H + kNpT - 1) / kNpT;
    const long long blocks = (long long)B * tiles_x * tiles_y;
    if (blocks > 0x7fffffffLL) return R2L_ERR_BAD_SHAPE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (raw_dtype == R2L_F32)
        isp_numpy_forward_kernel<float><<<(unsigned)blocks, kNpNT, 0, st>>>(static_cast<const float*>(raw), raw_denominator, H, W,
                                                                         tiles_x, tiles_x * tiles_y, P, out);
    else
        isp_numpy_forward_kernel<uint16_t><<<(unsigned)blocks, kNpNT, 0, st>>>(static_cast<const uint16_t*>(raw), raw_denominator,
                                                                            H, W, tiles_x, tiles_x * tiles_y, P, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}
