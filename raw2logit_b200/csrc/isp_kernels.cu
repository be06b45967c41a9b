// isp_kernels.cu -- sm_100a kernels and the C ABI (include/r2l_isp.h) of the fused differentiable ISP.
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC  (see _build.py)
// No torch headers, no host-side state: every call validates its arguments, enqueues kernels on the caller's
// stream and returns.
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/r2l_isp.h"
#include "isp_config.h"

namespace r2l {

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
template <class Cfg, typename RawT>
__global__ void __launch_bounds__(Cfg::NT) isp_forward_kernel(FwdArgs a, TileGrid grid) {
    extern __shared__ __align__(16) float smem[];
    fwd_cta<Cfg, RawT>(blockIdx.x, gridDim.x, a, grid, smem);
}

template <class Cfg, typename RawT>
__global__ void __launch_bounds__(Cfg::NT) isp_backward_kernel(BwdArgs a, TileGrid grid) {
    extern __shared__ __align__(16) float smem[];
    bwd_cta<Cfg, RawT>(blockIdx.x, gridDim.x, a, grid, smem);
}

// statistics of all CTAs -> 132 parameter gradients.  One CTA; sums over CTAs in double, in a fixed order.
constexpr int kFinishThreads = 256;
__global__ void __launch_bounds__(kFinishThreads) isp_backward_finish_kernel(Params P, const float* partials,
                                                                             int n_cta, float* grads) {
    __shared__ Tables T;
    __shared__ double S[kNumStats];
    R2L_BUILD_TABLES(kFinishThreads, P, &T)
    for (int s = threadIdx.x; s < kNumStats; s += kFinishThreads) {
        double sum = 0.0;
        for (int c = 0; c < n_cta; ++c) sum += (double)partials[(size_t)c * kStatPitch + s];
        S[s] = sum;
    }
    __syncthreads();
    for (int e = threadIdx.x; e < R2L_NUM_PARAM_GRADS; e += kFinishThreads) grads[e] = finish_grad(e, S, &T);
}

template <typename RawT>
__global__ void mosaic_kernel(const RawT* raw, float denom, int B, int H, int W, const float* black_level,
                              int reduce_size, int C, float* out) {
    float bl[4] = {0.f, 0.f, 0.f, 0.f};
    const bool has_bl = black_level != nullptr;
    if (has_bl) for (int i = 0; i < 4; ++i) bl[i] = black_level[i];
    const size_t plane = (size_t)H * W;
    if (!reduce_size) {
        const size_t n = (size_t)B * plane;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const size_t b = i / plane, pix = i - b * plane;
            const int y = (int)(pix / W), x = (int)(pix - (size_t)y * W);
            const int par = par_of(y, x);
            float v = RawLoad<RawT>::get(raw + i, denom);
            if (has_bl) v = v - bl[par];
            const int c_on = mosaic_channel(par, C);
            for (int c = 0; c < C; ++c) out[(b * C + c) * plane + pix] = (c == c_on) ? v : 0.f;
        }
    } else {
        const int h2 = H / 2, w2 = W / 2;
        const size_t plane2 = (size_t)h2 * w2, n = (size_t)B * plane2;
        for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
            const size_t b = i / plane2, q = i - b * plane2;
            const int qy = (int)(q / w2), qx = (int)(q - (size_t)qy * w2);
            const RawT* p = raw + b * plane + (size_t)(2 * qy) * W + 2 * qx;
            float v[4] = {RawLoad<RawT>::get(p, denom), RawLoad<RawT>::get(p + 1, denom),
                          RawLoad<RawT>::get(p + W, denom), RawLoad<RawT>::get(p + W + 1, denom)};
            if (has_bl) for (int k = 0; k < 4; ++k) v[k] = v[k] - bl[k];
            float* o = out + b * C * plane2 + q;
            if (C == 3) { o[0] = v[0]; o[plane2] = (v[1] + v[2]) / 2.f; o[2 * plane2] = v[3]; }
            else { o[0] = v[0]; o[plane2] = v[1]; o[2 * plane2] = v[2]; o[3 * plane2] = v[3]; }
        }
    }
}

__global__ void batch_sum_kernel(const float* x, const float* scale, int B, int C, int HW, float* out) {
    const size_t n = (size_t)C * HW;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float s = 0.f;
        for (int b = 0; b < B; ++b) s += x[(size_t)b * n + i];
        if (scale) s *= scale[i / HW];
        out[i] = s;
    }
}

__global__ void mosaic_backward_kernel(const float* gout, int B, int H, int W, int reduce_size, int C, float* graw) {
    const size_t plane = (size_t)H * W, n = (size_t)B * plane;
    const int h2 = H / 2, w2 = W / 2;
    const size_t plane2 = (size_t)h2 * w2;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        const size_t b = i / plane, pix = i - b * plane;
        const int y = (int)(pix / W), x = (int)(pix - (size_t)y * W);
        const int par = par_of(y, x);
        const int c = mosaic_channel(par, C);
        float g;
        if (!reduce_size) g = gout[(b * C + c) * plane + pix];
        else {
            g = gout[(b * C + c) * plane2 + (size_t)(y >> 1) * w2 + (x >> 1)];
            if (C == 3 && c == 1) g = g / 2.f;
        }
        graw[i] = g;
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
static thread_local int g_last_cuda_error = 0;

static int cuda_fail(cudaError_t e) { g_last_cuda_error = (int)e; return R2L_ERR_CUDA; }

static Params to_params(const r2l_isp_params* p) {
    Params q;
    q.black_level = p->black_level; q.white_balance = p->white_balance; q.colour_correction = p->colour_correction;
    q.gamma_correct = p->gamma_correct; q.debayer_weight = p->debayer_weight; q.sharpen_weight = p->sharpen_weight;
    q.gauss_weight = p->gauss_weight; q.rgb2yuv = p->rgb2yuv; q.yuv2rgb = p->yuv2rgb;
    return q;
}
static bool params_ok(const r2l_isp_params* p) {
    return p && p->black_level && p->white_balance && p->colour_correction && p->gamma_correct &&
           p->debayer_weight && p->sharpen_weight && p->gauss_weight && p->rgb2yuv && p->yuv2rgb;
}
static bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

static int check_common(const void* raw, int raw_dtype, int B, int H, int W, const r2l_isp_params* params) {
    if (raw_dtype != R2L_F32 && raw_dtype != R2L_U16) return R2L_ERR_BAD_DTYPE;
    if (B < 0 || H < 3 || W < 3) return R2L_ERR_BAD_SHAPE;
    if (B > 0 && !raw) return R2L_ERR_NULL_POINTER;
    if (!params_ok(params)) return R2L_ERR_NULL_POINTER;
    if (!aligned(raw, raw_dtype == R2L_F32 ? 4 : 2)) return R2L_ERR_MISALIGNED;
    return R2L_OK;
}

// persistent grid: one wave of resident CTAs (or fewer when the job is small)
template <typename K>
static int persistent_grid(K kernel, int threads, size_t smem, int n_tiles, int* grid_out) {
    int dev = 0, sms = 0, per_sm = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return cuda_fail(e);
    if (per_sm < 1) return R2L_ERR_BAD_ARGUMENT;
    int g = sms * per_sm;
    if (g > n_tiles) g = n_tiles;
    if (g > kMaxCtas) g = kMaxCtas;
    *grid_out = g;
    return R2L_OK;
}

template <class Cfg, typename RawT>
static int launch_forward(const FwdArgs& a, cudaStream_t st) {
    const TileGrid grid = make_grid(a.B, a.H, a.W, Cfg::TH, Cfg::TW);
    int g = 0;
    int rc = persistent_grid(isp_forward_kernel<Cfg, RawT>, Cfg::NT, Cfg::kSmemBytes, grid.n, &g);
    if (rc != R2L_OK) return rc;
    isp_forward_kernel<Cfg, RawT><<<g, Cfg::NT, Cfg::kSmemBytes, st>>>(a, grid);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

template <class Cfg, typename RawT>
static int launch_backward(const BwdArgs& a, float* grads, cudaStream_t st) {
    const TileGrid grid = make_grid(a.B, a.H, a.W, Cfg::TH, Cfg::TW);
    int g = 0;
    int rc = persistent_grid(isp_backward_kernel<Cfg, RawT>, Cfg::NT, Cfg::kSmemBytes, grid.n, &g);
    if (rc != R2L_OK) return rc;
    isp_backward_kernel<Cfg, RawT><<<g, Cfg::NT, Cfg::kSmemBytes, st>>>(a, grid);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    isp_backward_finish_kernel<<<1, kFinishThreads, 0, st>>>(a.P, a.partials, g, grads);
    e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

}  // namespace r2l

using namespace r2l;

extern "C" {

int r2l_isp_abi_version(void) { return R2L_ABI_VERSION; }

const char* r2l_isp_error_string(int code) {
    switch (code) {
        case R2L_OK: return "ok";
        case R2L_ERR_BAD_SHAPE: return "bad shape: need B >= 0 and H, W >= 3 (reflect padding of 2)";
        case R2L_ERR_BAD_DTYPE: return "bad raw dtype: expected R2L_F32 or R2L_U16";
        case R2L_ERR_NULL_POINTER: return "a required pointer is NULL";
        case R2L_ERR_MISALIGNED: return "pointer not aligned to its element size";
        case R2L_ERR_WORKSPACE: return "workspace too small";
        case R2L_ERR_CUDA: return "CUDA runtime error (see r2l_isp_last_cuda_error)";
        case R2L_ERR_BAD_ARGUMENT: return "bad argument";
        default: return "unknown error";
    }
}

int r2l_isp_last_cuda_error(void) { return g_last_cuda_error; }

int r2l_isp_forward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                    const r2l_isp_params* params, const r2l_isp_tail* tail, float* out, void* stream) {
    int rc = check_common(raw, raw_dtype, B, H, W, params);
    if (rc != R2L_OK) return rc;
    if (B == 0) return R2L_OK;
    if (!out) return R2L_ERR_NULL_POINTER;
    if (!aligned(out, 4)) return R2L_ERR_MISALIGNED;
    FwdArgs a;
    a.raw = raw; a.denom = raw_denominator; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.additive = tail ? tail->additive : nullptr;
    a.affine = tail ? tail->affine : nullptr;
    a.out = out;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return raw_dtype == R2L_F32 ? launch_forward<FwdDefault, float>(a, st)
                                : launch_forward<FwdDefault, uint16_t>(a, st);
}

size_t r2l_isp_backward_workspace_bytes(int B, int H, int W) {
    (void)B; (void)H; (void)W;
    return (size_t)kMaxCtas * kStatPitch * sizeof(float);
}

int r2l_isp_backward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                     const r2l_isp_params* params, const float* grad_out, const float* grad_out_scale,
                     float* grad_raw, float* grad_params, void* workspace, size_t workspace_bytes, void* stream) {
    int rc = check_common(raw, raw_dtype, B, H, W, params);
    if (rc != R2L_OK) return rc;
    if (!grad_params || !workspace || (B > 0 && !grad_out)) return R2L_ERR_NULL_POINTER;
    if (!aligned(grad_out, 4) || !aligned(grad_raw, 4) || !aligned(grad_params, 4) || !aligned(workspace, 8))
        return R2L_ERR_MISALIGNED;
    if (workspace_bytes < r2l_isp_backward_workspace_bytes(B, H, W)) return R2L_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0) {
        cudaError_t e = cudaMemsetAsync(grad_params, 0, R2L_NUM_PARAM_GRADS * sizeof(float), st);
        return e == cudaSuccess ? R2L_OK : cuda_fail(e);
    }
    BwdArgs a;
    a.raw = raw; a.denom = raw_denominator; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.gout = grad_out; a.gscale = grad_out_scale; a.graw = grad_raw; a.partials = static_cast<float*>(workspace);
    if (grad_raw) {
        return raw_dtype == R2L_F32 ? launch_backward<BwdWithRaw, float>(a, grad_params, st)
                                    : launch_backward<BwdWithRaw, uint16_t>(a, grad_params, st);
    }
    return raw_dtype == R2L_F32 ? launch_backward<BwdNoRaw, float>(a, grad_params, st)
                                : launch_backward<BwdNoRaw, uint16_t>(a, grad_params, st);
}

int r2l_isp_mosaic(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                   const float* black_level, int reduce_size, int out_channels, float* out, void* stream) {
    if (raw_dtype != R2L_F32 && raw_dtype != R2L_U16) return R2L_ERR_BAD_DTYPE;
    if (out_channels != 3 && out_channels != 4) return R2L_ERR_BAD_ARGUMENT;
    if (B < 0 || H < 0 || W < 0) return R2L_ERR_BAD_SHAPE;
    if (reduce_size && ((H & 1) || (W & 1))) return R2L_ERR_BAD_SHAPE;   // the reference raises for odd sizes
    const size_t n = (size_t)B * H * W;
    if (n == 0) return R2L_OK;
    if (!raw || !out) return R2L_ERR_NULL_POINTER;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t work = reduce_size ? n / 4 : n;
    const int threads = 256;
    const int blocks = (int)((work + threads - 1) / threads < 148 * 16 ? (work + threads - 1) / threads : 148 * 16);
    if (raw_dtype == R2L_F32)
        mosaic_kernel<float><<<blocks, threads, 0, st>>>(static_cast<const float*>(raw), raw_denominator, B, H, W,
                                                         black_level, reduce_size, out_channels, out);
    else
        mosaic_kernel<uint16_t><<<blocks, threads, 0, st>>>(static_cast<const uint16_t*>(raw), raw_denominator, B,
                                                            H, W, black_level, reduce_size, out_channels, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

int r2l_isp_mosaic_backward(const float* grad_out, int B, int H, int W, int reduce_size, int out_channels,
                            float* grad_raw, void* stream) {
    if (out_channels != 3 && out_channels != 4) return R2L_ERR_BAD_ARGUMENT;
    if (B < 0 || H < 0 || W < 0) return R2L_ERR_BAD_SHAPE;
    if (reduce_size && ((H & 1) || (W & 1))) return R2L_ERR_BAD_SHAPE;
    const size_t n = (size_t)B * H * W;
    if (n == 0) return R2L_OK;
    if (!grad_out || !grad_raw) return R2L_ERR_NULL_POINTER;
    const int threads = 256;
    const int blocks = (int)((n + threads - 1) / threads < 148 * 16 ? (n + threads - 1) / threads : 148 * 16);
    mosaic_backward_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(grad_out, B, H, W, reduce_size,
                                                                                      out_channels, grad_raw);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

int r2l_isp_batch_sum(const float* x, const float* scale, int B, int C, int HW, float* out, void* stream) {
    if (B < 0 || C < 0 || HW < 0) return R2L_ERR_BAD_SHAPE;
    const size_t n = (size_t)C * HW;
    if (n == 0) return R2L_OK;
    if (!out || (B > 0 && !x)) return R2L_ERR_NULL_POINTER;
    const int threads = 256;
    const int blocks = (int)((n + threads - 1) / threads < 148 * 16 ? (n + threads - 1) / threads : 148 * 16);
    batch_sum_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(x, scale, B, C, HW, out);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

}  // extern "C"
