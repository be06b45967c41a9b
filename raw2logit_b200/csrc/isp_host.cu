// isp_host.cu -- the C ABI (include/r2l_isp.h) of the fused differentiable ISP, its argument checks and dispatch,
// and the small auxiliary kernels (BatchNorm tail, statistics finish, CFA split).  The fused forward / backward
// kernels live in their own translation units (isp_fwd_*.cu, isp_bwd_*.cu, isp_generic.cu; see isp_launch.h).
//
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -shared -Xcompiler -fPIC  (see _build.py)
// No torch headers, no host-side state: every call validates its arguments, enqueues kernels on the caller's
// stream and returns.
#include <atomic>
#include <mutex>
#include <vector>

#include "isp_launch.h"

namespace r2l {

// ---------------------------------------------------------------------------------------------------------
// kernels
// ---------------------------------------------------------------------------------------------------------
// ---- train-mode BatchNorm2d(3, affine=False) tail (pipeline_torch.py:168, 216-217) -------------------------
// per-CTA channel sums -> batch mean / biased variance -> {scale, shift}; running statistics updated in place
// exactly like torch (momentum update, unbiased variance for the running estimate).
__global__ void bn_finish_kernel(const float* partials, int n_cta, double count, float momentum, float eps,
                                 float* running_mean, float* running_var, float* affine, long long* num_batches) {
    // one warp per channel; lanes stride over the per-CTA partial sums (independent loads), fixed-order fp64 reduction
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= 3) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < n_cta; i += 32) {
        s1 += (double)partials[(size_t)i * kChanPitch + c];
        s2 += (double)partials[(size_t)i * kChanPitch + 3 + c];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane != 0) return;
    const double mean = s1 / count;
    double var = s2 / count - mean * mean;
    if (var < 0.0) var = 0.0;
    const double inv = 1.0 / sqrt(var + (double)eps);
    affine[c] = (float)inv;
    affine[3 + c] = (float)(-mean * inv);
    if (c == 0 && num_batches) *num_batches += 1;
    if (running_mean) running_mean[c] = (float)((1.0 - momentum) * (double)running_mean[c] + momentum * mean);
    if (running_var) {
        const double unbiased = count > 1.0 ? var * count / (count - 1.0) : var;
        running_var[c] = (float)((1.0 - momentum) * (double)running_var[c] + momentum * unbiased);
    }
}

// x[b][c][i] = x*scale[c] + shift[c], in place
__global__ void affine_inplace_kernel(float* x, const float* affine, int BC, int HW) {
    const int bc = blockIdx.y;
    if (bc >= BC) return;
    const int c = bc % 3;
    const float sc = affine[c], sh = affine[3 + c];
    float* p = x + (size_t)bc * HW;
    if ((HW & 3) == 0 && (reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        float4* p4 = reinterpret_cast<float4*>(p);
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW / 4; i += gridDim.x * blockDim.x) {
            float4 v = p4[i];
            v.x = fmaf(v.x, sc, sh); v.y = fmaf(v.y, sc, sh); v.z = fmaf(v.z, sc, sh); v.w = fmaf(v.w, sc, sh);
            p4[i] = v;
        }
    } else {
        for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x)
            p[i] = fmaf(p[i], sc, sh);
    }
}

// backward of the batch statistics: per channel sum(gy), sum(gy*y) with y the normalised output.
// A pure streaming reduction over 24 B/px: 128-bit loads, four independent accumulator pairs per thread (eight loads in
// flight), two CTAs per SM and channel.  (The first version -- scalar loads into one dependent accumulator chain, 128
// CTAs per channel -- ran at 1.4 TB/s: 69.6 us for 64 x 256 x 256, more than the fused forward kernel.)
// head (may be null): deferred mode -- CTA 0 of every channel also writes the channel's entries of the 15-float tail that
// need no reduction (gs, ysc, ysh from the forward's saved affine) and tags c1 / c2 as "finish me from the partials"
template <bool VEC>
__global__ void __launch_bounds__(256) bn_backward_stats_kernel(const float* __restrict__ gy, const float* __restrict__ y,
                                                                int B, int HW, float* partials /* [3][kBnBwdBlocks][2] */,
                                                                const float* affine, float* head) {
    const int c = blockIdx.y;
    float s1[4] = {0.f, 0.f, 0.f, 0.f}, s2[4] = {0.f, 0.f, 0.f, 0.f};
    if (VEC) {                                          // HW % 4 == 0, 16-byte aligned tensors, B * HW / 4 < 2^31
        // the (image, position) space of this channel, flattened, is cut into one contiguous chunk per CTA, so every
        // CTA has work whatever the plane size (a grid-stride loop inside each plane left 3/4 of the CTAs idle at 256^2)
        const int n4 = HW >> 2, total = B * n4;
        const int per = (total + (int)gridDim.x - 1) / (int)gridDim.x;
        const int l0 = blockIdx.x * per, l1 = min(total, l0 + per);
        const float4* g4 = reinterpret_cast<const float4*>(gy);
        const float4* y4 = reinterpret_cast<const float4*>(y);
        for (int l = l0 + threadIdx.x; l < l1; l += 4 * 256) {
            float4 g[4], v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int lu = l + u * 256;
                g[u] = make_float4(0.f, 0.f, 0.f, 0.f); v[u] = g[u];
                if (lu < l1) {
                    const int b = lu / n4, i = lu - b * n4;
                    const size_t at = ((size_t)b * 3 + c) * n4 + i;
                    g[u] = g4[at]; v[u] = y4[at];              // default cache policy: the fused backward re-reads both from L2
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                s1[u] += (g[u].x + g[u].y) + (g[u].z + g[u].w);
                s2[u] = fmaf(g[u].x, v[u].x, fmaf(g[u].y, v[u].y, fmaf(g[u].z, v[u].z, fmaf(g[u].w, v[u].w, s2[u]))));
            }
        }
    } else {
        for (int b = 0; b < B; ++b) {
            const size_t base = ((size_t)b * 3 + c) * HW;
            for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < HW; i += gridDim.x * blockDim.x) {
                const float g = gy[base + i];
                s1[0] += g;
                s2[0] = fmaf(g, y[base + i], s2[0]);
            }
        }
    }
    float t1 = (s1[0] + s1[1]) + (s1[2] + s1[3]), t2 = (s2[0] + s2[1]) + (s2[2] + s2[3]);
    __shared__ float red[2][8];
    t1 = warp_sum_all(t1); t2 = warp_sum_all(t2);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) { red[0][warp] = t1; red[1][warp] = t2; }
    __syncthreads();
    if (threadIdx.x < 2) {
        float t = 0.f;
        for (int w = 0; w < 8; ++w) t += red[threadIdx.x][w];
        partials[((size_t)c * kBnBwdBlocks + blockIdx.x) * 2 + threadIdx.x] = t;
    }
    if (head && blockIdx.x == 0 && threadIdx.x == 0) {
        head[c] = affine[c];                                    // gs  = 1/sqrt(var+eps)
        head[3 + c] = __uint_as_float(kTailDeferredTag);        // c1, c2: from the partials, by the consumer
        head[6 + c] = __uint_as_float(kTailDeferredTag);
        head[9 + c] = affine[c];                                // ysc
        head[12 + c] = affine[3 + c];                           // ysh
    }
}
// one CTA of >= 96 threads: gtail_in (deferred or complete) -> a complete 15-float tail in `out` (for the kernels that
// do not finish a deferred tail themselves: third generation, generic)
__global__ void bn_tail_resolve_kernel(const float* gtail_in, const float* partials, double count, float* out) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= 3) return;
    float c1 = gtail_in[3 + c], c2 = gtail_in[6 + c];
    if (partials && __float_as_uint(gtail_in[3]) == kTailDeferredTag) {
        double s1 = 0.0, s2 = 0.0;
        for (int i = lane; i < kBnBwdBlocks; i += 32) {
            s1 += (double)partials[((size_t)c * kBnBwdBlocks + i) * 2];
            s2 += (double)partials[((size_t)c * kBnBwdBlocks + i) * 2 + 1];
        }
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) {
            s1 += __shfl_xor_sync(0xffffffffu, s1, o);
            s2 += __shfl_xor_sync(0xffffffffu, s2, o);
        }
        c1 = (float)(s1 / count);
        c2 = (float)(s2 / count);
    }
    if (lane != 0) return;
    out[c] = gtail_in[c]; out[3 + c] = c1; out[6 + c] = c2; out[9 + c] = gtail_in[9 + c]; out[12 + c] = gtail_in[12 + c];
}
__global__ void bn_backward_finish_kernel(const float* partials, const float* affine, double count, float* gtail) {
    const int c = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (c >= 3) return;
    double s1 = 0.0, s2 = 0.0;
    for (int i = lane; i < kBnBwdBlocks; i += 32) {
        s1 += (double)partials[((size_t)c * kBnBwdBlocks + i) * 2];
        s2 += (double)partials[((size_t)c * kBnBwdBlocks + i) * 2 + 1];
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        s1 += __shfl_xor_sync(0xffffffffu, s1, o);
        s2 += __shfl_xor_sync(0xffffffffu, s2, o);
    }
    if (lane != 0) return;
    gtail[c] = affine[c];                       // gs  = 1/sqrt(var+eps)
    gtail[3 + c] = (float)(s1 / count);         // c1  = mean(gy)
    gtail[6 + c] = (float)(s2 / count);         // c2  = mean(gy * yhat)
    gtail[9 + c] = affine[c];                   // ysc
    gtail[12 + c] = affine[3 + c];              // ysh
}

// statistics of all CTAs -> 132 parameter gradients.  One CTA of 1024 threads.  The per-CTA partial sums are read with
// every load independent and in flight at once, summed in double in a fixed order (bit-reproducible); the chain
// rule from the collapsed-table statistics back to black level / white balance / colour matrix / demosaic taps
// (finish_grad_sc, isp_core.cuh) is spread over warps: its long sums (36 and 27 terms) are one term per lane.
constexpr int kFinishThreads = 1024;
constexpr int kFinishSegs = kFinishThreads / kStatPitch;        // 6 segments of CTAs
__device__ __forceinline__ double warp_sum_all_f64(double v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(kFinishThreads) isp_backward_finish_kernel(Params P, const float* partials,
                                                                             int n_cta, float* grads) {
    __shared__ Tables T;
    __shared__ double Sseg[kFinishSegs][kStatPitch];
    __shared__ double S[kNumStats];
    __shared__ double Qr[108];          // [k][par][t]
    __shared__ double Tkc[9];           // [k][c] = sum_{par,t} wd[c][ch(par_tap)][t] * Qr[k][par][t]
    __shared__ double Gbl[4];
    __shared__ double Sc9[9];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int s = tid % kStatPitch, seg = tid / kStatPitch;
    double sum = 0.0;
    if (seg < kFinishSegs && s < kNumStats) {
        // up to 32 CTAs per thread and round, every load issued before the first add (one L2 round trip per round)
        for (int c0 = seg; c0 < n_cta; c0 += 32 * kFinishSegs) {
            float v[32];
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                const int c = c0 + u * kFinishSegs;
                v[u] = c < n_cta ? __ldcg(partials + (size_t)c * kStatPitch + s) : 0.f;
            }
#pragma unroll
            for (int u = 0; u < 32; ++u) sum += (double)v[u];
        }
    }
    if (seg < kFinishSegs) Sseg[seg][s] = sum;
    R2L_BUILD_TABLES(kFinishThreads, P, &T)                         // ends with a barrier
    if (tid < kNumStats) {
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < kFinishSegs; ++g) t += Sseg[g][tid];
        S[tid] = t;
    }
    __syncthreads();
    if (tid < 108) { const int k = tid / 36, r = tid - 36 * k; Qr[tid] = finish_qr(S, &T, k, r / 9, r % 9); }
    __syncthreads();
    if (warp < 9) {                                                  // Tkc[k][c]: 36 terms, lanes take (par, t)
        const int k = warp / 3, c = warp - 3 * k;
        double v = 0.0;
        for (int i = lane; i < 36; i += 32) {
            const int par = i / 9, t = i - 9 * par;
            v += (double)T.wd[(c * 3 + ch_of(par_tap(par, t))) * 9 + t] * Qr[k * 36 + i];
        }
        v = warp_sum_all_f64(v);
        if (lane == 0) Tkc[warp] = v;
    } else if (warp < 13) {                                          // black_level[e]: 27 terms, lanes take (t, k)
        const int e = warp - 9;
        double v = 0.0;
        if (lane < 27) {
            const int t = lane / 3, k = lane - 3 * t, par = par_tap(e, t);      // par_tap(par, t) == e (an involution)
            v = (double)T.AW[par][k][t] * S[stat_p_index(k, par)];
        }
        v = warp_sum_all_f64(v);
        if (lane == 0) Gbl[e] = -v;
    }
    __syncthreads();
    if (tid < 9) {                                                   // Sc[m][c] = sum_k M1[k][m] * Tkc[k][c]
        const int m = tid / 3, c = tid - 3 * m;
        double v = 0.0;
        for (int k = 0; k < 3; ++k) v += (double)T.m1[k * 3 + m] * Tkc[k * 3 + c];
        Sc9[tid] = v;
    }
    __syncthreads();
    if (tid < 4) grads[tid] = (float)Gbl[tid];
    else if (tid < R2L_NUM_PARAM_GRADS) grads[tid] = finish_grad_sc(tid, S, &T, Sc9);
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

int cuda_fail(cudaError_t e) { g_last_cuda_error = (int)e; return R2L_ERR_CUDA; }

// resident CTAs of `kernel` on the current device (SM count x CTAs per SM), remembered per (kernel, device)
int cached_ctas_per_device(const void* kernel, int threads, size_t smem, int* out) {
    struct Entry { const void* kernel; int dev; int ctas; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Entry& en : cache)
            if (en.kernel == kernel && en.dev == dev) { *out = en.ctas; return R2L_OK; }
    }
    int sms = 0, per_sm = 0;
    e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e);
    e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem);
    if (e != cudaSuccess) return cuda_fail(e);
    if (per_sm < 1) return R2L_ERR_BAD_ARGUMENT;
    std::lock_guard<std::mutex> lock(mu);
    cache.push_back({kernel, dev, sms * per_sm});
    *out = sms * per_sm;
    return R2L_OK;
}

int tmem_ctas_per_device(const void* kernel, int threads, size_t smem, int tmem_cols, int want_per_sm, int* out) {
    struct Entry { const void* kernel; int dev; int ctas; };
    static std::mutex mu;
    static std::vector<Entry> cache;
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return cuda_fail(e);
    {
        std::lock_guard<std::mutex> lock(mu);
        for (const Entry& en : cache)
            if (en.kernel == kernel && en.dev == dev) { *out = en.ctas; return R2L_OK; }
    }
    e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return cuda_fail(e);
    cudaFuncAttributes fa;
    e = cudaFuncGetAttributes(&fa, kernel);
    if (e != cudaSuccess) return cuda_fail(e);
    int sms = 0, regs_sm = 0, smem_sm = 0, smem_reserved = 0;
    if ((e = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev)) != cudaSuccess ||
        (e = cudaDeviceGetAttribute(&smem_reserved, cudaDevAttrReservedSharedMemoryPerBlock, dev)) != cudaSuccess)
        return cuda_fail(e);
    const int warps = (threads + 31) / 32;
    const int regs_cta = warps * ((fa.numRegs * 32 + 255) / 256 * 256);                // per-warp allocation unit: 256
    const size_t smem_cta = (smem + fa.sharedSizeBytes + (size_t)smem_reserved + 127) / 128 * 128;
    int per_sm = want_per_sm;
    if (regs_cta > 0 && regs_sm / regs_cta < per_sm) per_sm = regs_sm / regs_cta;
    if ((int)((size_t)smem_sm / smem_cta) < per_sm) per_sm = (int)((size_t)smem_sm / smem_cta);
    if (tmem_cols > 0 && 512 / tmem_cols < per_sm) per_sm = 512 / tmem_cols;
    if (per_sm < 1) return R2L_ERR_BAD_ARGUMENT;
    std::lock_guard<std::mutex> lock(mu);
    cache.push_back({kernel, dev, sms * per_sm});
    *out = sms * per_sm;
    return R2L_OK;
}

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

static int check_common(const void* raw, int raw_dtype, int B, int H, int W, const r2l_isp_params* params) {
    if (raw_dtype != R2L_F32 && raw_dtype != R2L_U16) return R2L_ERR_BAD_DTYPE;
    if (B < 0 || H < 3 || W < 3) return R2L_ERR_BAD_SHAPE;
    if (B > 0 && !raw) return R2L_ERR_NULL_POINTER;
    if (!params_ok(params)) return R2L_ERR_NULL_POINTER;
    if (!aligned(raw, raw_dtype == R2L_F32 ? 4 : 2)) return R2L_ERR_MISALIGNED;
    return R2L_OK;
}

// ---- tensor map over the raw batch: dims (W, H, B), box (box_w, box_h, 2), zero fill outside ---------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            return nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}
// false when the shape/pointer does not meet TMA's 16-byte rules (then a non-TMA kernel runs)
bool make_raw_tensor_map(CUtensorMap* map, const void* raw, int elem_bytes, int B, int H, int W, int box_w,
                                int box_h) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    if (const char* off = getenv("R2L_ISP_NO_TMA")) {          // debugging knob: force the generic loader
        if (off[0] == '1') return false;
    }
    if ((reinterpret_cast<uintptr_t>(raw) & 15) || ((size_t)W * elem_bytes) % 16 != 0) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
    const cuuint64_t gstride[2] = {(cuuint64_t)W * elem_bytes, (cuuint64_t)W * H * elem_bytes};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 2u};
    const cuuint32_t estr[3] = {1u, 1u, 1u};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    return fn(map, dt, 3, const_cast<void*>(raw), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// generic fp32 / uint16 tiled map: dims / strides (bytes, dims 1..rank-1) / box, zero fill, no swizzle
static bool encode_map(CUtensorMap* map, const void* base, int elem_bytes, int rank, const cuuint64_t* gdim,
                       const cuuint64_t* gstride, const cuuint32_t* box) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) return false;
    if (reinterpret_cast<uintptr_t>(base) & 15) return false;
    for (int i = 0; i + 1 < rank; ++i) if (gstride[i] % 16 != 0) return false;
    for (int i = 0; i < rank; ++i) if (box[i] == 0 || box[i] > 256) return false;
    const cuuint32_t estr[5] = {1u, 1u, 1u, 1u, 1u};
    const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_UINT16;
    return fn(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

bool make_bwd5_prefetch_maps(Bwd5Maps* m, const BwdArgs& a, int raw_elem_bytes, int th, int tw) {
    if (!a.out || !a.luma) return false;
    const cuuint64_t W = (cuuint64_t)a.W, H = (cuuint64_t)a.H, B = (cuuint64_t)a.B, P = (B + 1) / 2;
    {
        const cuuint64_t gdim[4] = {W, H, 3, B};
        const cuuint64_t gstr[3] = {W * 4, W * H * 4, W * H * 12};
        const cuuint32_t box[4] = {(cuuint32_t)tw + 8, (cuuint32_t)th + 8, 3, 2};
        if (!encode_map(&m->gout, a.gout, 4, 4, gdim, gstr, box)) return false;
        if (!encode_map(&m->out, a.out, 4, 4, gdim, gstr, box)) return false;
    }
    {
        const cuuint64_t gdim[4] = {2 * W, H, P, 2};
        const cuuint64_t gstr[3] = {W * 8, W * H * 8, W * H * P * 8};
        const cuuint32_t box[4] = {(cuuint32_t)(2 * tw), (cuuint32_t)th, 1, 2};
        if (!encode_map(&m->luma, a.luma, 4, 4, gdim, gstr, box)) return false;
    }
    {
        const cuuint64_t eb = (cuuint64_t)raw_elem_bytes;
        const cuuint64_t gdim[3] = {W, H, B};
        const cuuint64_t gstr[2] = {W * eb, W * H * eb};
        const cuuint32_t box[3] = {(cuuint32_t)tw, (cuuint32_t)th, 2};
        if (!encode_map(&m->raw, a.raw, raw_elem_bytes, 3, gdim, gstr, box)) return false;
    }
    return true;
}

// the ticket words live behind the per-CTA statistics rows of the workspace: bytes 0..7 the backward's final ticket,
// 64..79 the fused BatchNorm forward's barrier words, 128..255 the 16 group tickets of the backward's two-level finish
// (isp_bwd4.cuh), whose 16 fp64 group rows occupy the last 32 statistics rows (kFinishMaxCtas caps that kernel's grid)
constexpr size_t kTicketOffset = (size_t)kMaxCtas * kStatPitch * sizeof(float);
// behind the 256 bytes of ticket words: the BatchNorm backward's per-CTA sums (deferred tail) and a resolved tail
constexpr size_t kBnPartialsOffset = kTicketOffset + 256;
constexpr size_t kBnPartialsBytes = ((size_t)3 * kBnBwdBlocks * 2 * sizeof(float) + 127) / 128 * 128;
constexpr size_t kBnTailOffset = kBnPartialsOffset + kBnPartialsBytes;
constexpr size_t kWorkspaceBytes = kBnTailOffset + 64;

// does a forward call with these arguments run the kernel that can save the luma planes?  (one rule, used by
// r2l_isp_forward, r2l_isp_forward_bn_train and r2l_isp_luma_supported)
static bool luma_path_ok(const void* raw, int raw_dtype, int H, int W, const float* out, const float* additive) {
    const char* force = getenv("R2L_ISP_FORCE_GENERIC");
    if (force && force[0] == '1') return false;
    if (const char* off = getenv("R2L_ISP_NO_TMA")) { if (off[0] == '1') return false; }
    const int eb = raw_dtype == R2L_F32 ? 4 : 2;
    return fwd3_shape_ok(H, W) && aligned(out, 16) && aligned(additive, 16) && aligned(raw, 16) &&
           ((size_t)W * eb) % 16 == 0;
}

unsigned next_ticket_generation() {
    static std::atomic<unsigned> gen{0};
    unsigned g = gen.fetch_add(1u) + 1u;
    if (g == 0u) g = gen.fetch_add(1u) + 1u;
    return g;
}

bool pdl_enabled() {
    static const bool on = [] { const char* v = getenv("R2L_ISP_NO_PDL"); return !(v && v[0] == '1'); }();
    return on;
}

bool pdl_enabled_backward() {
    static const bool on = [] { const char* v = getenv("R2L_ISP_BWD_PDL"); return v && v[0] == '1'; }();
    return on && pdl_enabled();
}

// statistics of the backward CTAs -> 132 gradients
static int launch_finish(const BwdArgs& a, int n_cta, float* grads, cudaStream_t st) {
    isp_backward_finish_kernel<<<1, kFinishThreads, 0, st>>>(a.P, a.partials, n_cta, grads);
    cudaError_t e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

static int launch_forward_any(const FwdArgs& a, int raw_dtype, bool stats, cudaStream_t st, int* grid_used,
                              bool* fused_tail = nullptr) {
    return raw_dtype == R2L_F32 ? launch_forward_f32(a, stats, st, grid_used, fused_tail)
                                : launch_forward_u16(a, stats, st, grid_used, fused_tail);
}

// vectorised third-generation kernel when the shape / alignment allows, generic scalar kernel otherwise
static int launch_backward_any(const BwdArgs& a, int raw_dtype, float* grads, cudaStream_t st) {
    int g = 0;
    int rc = kNotServed;
    const char* force = getenv("R2L_ISP_FORCE_GENERIC");        // debugging knob
    if (a.world > 1) {                                          // the fused exchange lives in the fifth generation only
        if ((force && force[0] == '1') || !(a.out && a.luma)) return R2L_ERR_BAD_ARGUMENT;
    }
    if (!(force && force[0] == '1')) {
        if (a.out && a.luma) {                                  // fourth / fifth generation: nothing recomputed, fused finish
            BwdArgs a4 = a;
            a4.grads = grads;
            a4.ticket = reinterpret_cast<unsigned*>(reinterpret_cast<char*>(a.partials) + kTicketOffset);
            rc = raw_dtype == R2L_F32 ? launch_backward5_f32(a4, st, &g) : launch_backward5_u16(a4, st, &g);
            if (rc == R2L_OK) return rc;
        }
        if (a.world > 1) return rc == kNotServed ? (int)R2L_ERR_BAD_ARGUMENT : rc;   // the exchange lives in that kernel only
    }
    // the older generations read a complete tail: resolve a (possibly) deferred one into the workspace first
    BwdArgs b = a;
    if (rc == kNotServed && a.gtail && a.tail_ws) {
        bn_tail_resolve_kernel<<<1, 96, 0, st>>>(a.gtail, a.bn_partials, a.bn_count, a.tail_ws);
        cudaError_t e = cudaGetLastError();
        if (e != cudaSuccess) return cuda_fail(e);
        b.gtail = a.tail_ws;
    }
    if (rc == kNotServed && !(force && force[0] == '1'))
        rc = raw_dtype == R2L_F32 ? launch_backward3_f32(b, st, &g) : launch_backward3_u16(b, st, &g);
    if (rc == kNotServed) rc = launch_backward_generic(b, raw_dtype, st, &g);
    if (rc != R2L_OK) return rc;
    return launch_finish(b, g, grads, st);
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
                    const r2l_isp_params* params, const r2l_isp_tail* tail, float* out, float* saved_luma,
                    void* stream) {
    int rc = check_common(raw, raw_dtype, B, H, W, params);
    if (rc != R2L_OK) return rc;
    if (B == 0) return R2L_OK;
    if (!out) return R2L_ERR_NULL_POINTER;
    if (!aligned(out, 4)) return R2L_ERR_MISALIGNED;
    if (saved_luma && (!aligned(saved_luma, 32) || !luma_path_ok(raw, raw_dtype, H, W, out, tail ? tail->additive : nullptr)))
        return R2L_ERR_BAD_ARGUMENT;                           // ask r2l_isp_luma_supported first
    FwdArgs a;
    a.luma = saved_luma;
    a.raw = raw; a.denom = raw_denominator; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.additive = tail ? tail->additive : nullptr;
    a.affine = tail ? tail->affine : nullptr;
    a.out = out; a.chan_partials = nullptr;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    return launch_forward_any(a, raw_dtype, false, st, nullptr);
}

size_t r2l_isp_workspace_bytes(int B, int H, int W) {
    (void)B; (void)H; (void)W;
    return kWorkspaceBytes;
}

size_t r2l_isp_saved_luma_floats(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return (size_t)4 * ((B + 1) / 2) * H * W;
}

int r2l_isp_luma_supported(const void* raw, int raw_dtype, int B, int H, int W, const float* out, const float* additive) {
    if (B <= 0 || H < 3 || W < 3 || (raw_dtype != R2L_F32 && raw_dtype != R2L_U16)) return 0;
    return luma_path_ok(raw, raw_dtype, H, W, out, additive) ? 1 : 0;
}

int r2l_isp_forward_bn_train(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                             const r2l_isp_params* params, const float* additive, float* out,
                             float* running_mean, float* running_var, long long* num_batches_tracked, float momentum,
                             float eps, float* saved_affine, float* saved_luma, void* workspace, size_t workspace_bytes,
                             void* stream) {
    int rc = check_common(raw, raw_dtype, B, H, W, params);
    if (rc != R2L_OK) return rc;
    if (!out || !saved_affine || !workspace) return R2L_ERR_NULL_POINTER;
    if (saved_luma && (!aligned(saved_luma, 32) || !luma_path_ok(raw, raw_dtype, H, W, out, additive)))
        return R2L_ERR_BAD_ARGUMENT;
    if (!aligned(out, 4) || !aligned(workspace, 8)) return R2L_ERR_MISALIGNED;
    if (workspace_bytes < r2l_isp_workspace_bytes(B, H, W)) return R2L_ERR_WORKSPACE;
    if ((size_t)B * H * W < 2) return R2L_ERR_BAD_SHAPE;     // torch: "Expected more than 1 value per channel"
    FwdArgs a;
    a.raw = raw; a.denom = raw_denominator; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.additive = additive; a.affine = nullptr; a.out = out; a.chan_partials = static_cast<float*>(workspace);
    a.luma = saved_luma;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // the third-generation kernel finishes the statistics and normalises its tiles itself (one launch); the ticket words
    // of its grid-wide barrier sit behind the backward's ticket in the workspace
    const char* split = getenv("R2L_ISP_BN_SPLIT");             // debugging knob: the three-launch path
    if (!(split && split[0] == '1')) {
        a.bn_sync = reinterpret_cast<unsigned*>(static_cast<char*>(workspace) + kTicketOffset + 64);
        a.bn_gen = next_ticket_generation();
        a.bn_count = (double)B * H * W; a.bn_momentum = momentum; a.bn_eps = eps;
        a.bn_running_mean = running_mean; a.bn_running_var = running_var; a.bn_saved_affine = saved_affine;
        a.bn_num_batches = num_batches_tracked;
    }
    int g = 0;
    bool fused = false;
    rc = launch_forward_any(a, raw_dtype, true, st, &g, &fused);
    if (rc != R2L_OK) return rc;
    if (fused) return R2L_OK;
    bn_finish_kernel<<<1, 96, 0, st>>>(a.chan_partials, g, (double)B * H * W, momentum, eps, running_mean,
                                       running_var, saved_affine, num_batches_tracked);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    const int hw = H * W;
    dim3 grid((unsigned)((hw / 4 + 255) / 256 < 64 ? (hw / 4 + 255) / 256 + 1 : 64), (unsigned)(B * 3));
    affine_inplace_kernel<<<grid, 256, 0, st>>>(out, saved_affine, B * 3, hw);
    e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

int r2l_isp_bn_backward_prepare(const float* grad_out, const float* out, const float* saved_affine, int B, int H,
                                int W, float* grad_tail, void* workspace, size_t workspace_bytes, void* stream) {
    if (B < 1 || H < 1 || W < 1) return R2L_ERR_BAD_SHAPE;
    if (!grad_out || !out || !saved_affine || !grad_tail || !workspace) return R2L_ERR_NULL_POINTER;
    if (workspace_bytes < (size_t)3 * kBnBwdBlocks * 2 * sizeof(float)) return R2L_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // With the full workspace the tail is DEFERRED: the statistics kernel leaves its per-CTA sums in the workspace and tags
    // c1 / c2 of grad_tail; r2l_isp_backward (same workspace, same stream, next) finishes them in its kernel's prologue.
    // A caller with the minimal workspace gets the separate one-CTA finish kernel and a complete grad_tail.
    const bool deferred = workspace_bytes >= kWorkspaceBytes && aligned(workspace, 8);
    float* partials = deferred ? reinterpret_cast<float*>(static_cast<char*>(workspace) + kBnPartialsOffset)
                               : static_cast<float*>(workspace);
    float* head = deferred ? grad_tail : nullptr;
    if (((H * W) & 3) == 0 && aligned(grad_out, 16) && aligned(out, 16) && (size_t)B * H * W / 4 < ((size_t)1 << 31) - 1024)
        bn_backward_stats_kernel<true><<<dim3(kBnBwdBlocks, 3), 256, 0, st>>>(grad_out, out, B, H * W, partials, saved_affine, head);
    else
        bn_backward_stats_kernel<false><<<dim3(kBnBwdBlocks, 3), 256, 0, st>>>(grad_out, out, B, H * W, partials, saved_affine, head);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail(e);
    if (deferred) return R2L_OK;
    bn_backward_finish_kernel<<<1, 96, 0, st>>>(partials, saved_affine, (double)B * H * W, grad_tail);
    e = cudaGetLastError();
    return e == cudaSuccess ? R2L_OK : cuda_fail(e);
}

size_t r2l_isp_exchange_bytes(int world) {
    if (world < 1 || world > kMaxWorld) return 0;
    return (size_t)2 * world * kSlotPitch * 8 + 16;            // + the device-side epoch counter (R2L_EPOCH_DEVICE)
}

static int backward_impl(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                         const r2l_isp_params* params, const float* grad_out, const float* grad_tail,
                         const float* additive, const float* out, const float* saved_luma, float* grad_raw,
                         float* grad_params, void* workspace, size_t workspace_bytes, const r2l_isp_allreduce* dp,
                         void* stream);

int r2l_isp_backward(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                     const r2l_isp_params* params, const float* grad_out, const float* grad_tail,
                     const float* additive, const float* out, const float* saved_luma, float* grad_raw,
                     float* grad_params, void* workspace, size_t workspace_bytes, void* stream) {
    return backward_impl(raw, raw_dtype, raw_denominator, B, H, W, params, grad_out, grad_tail, additive, out, saved_luma,
                         grad_raw, grad_params, workspace, workspace_bytes, nullptr, stream);
}

int r2l_isp_backward_dp(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                        const r2l_isp_params* params, const float* grad_out, const float* grad_tail,
                        const float* additive, const float* out, const float* saved_luma, float* grad_raw,
                        float* grad_params, void* workspace, size_t workspace_bytes, const r2l_isp_allreduce* dp,
                        void* stream) {
    if (!dp) return R2L_ERR_NULL_POINTER;
    if (dp->world < 1 || dp->world > kMaxWorld || dp->rank < 0 || dp->rank >= dp->world || dp->epoch == 0)
        return R2L_ERR_BAD_ARGUMENT;
    if (dp->world > 1 && !dp->peers) return R2L_ERR_NULL_POINTER;
    if (B <= 0) return R2L_ERR_BAD_ARGUMENT;                  // an empty shard cannot take part in the exchange
    return backward_impl(raw, raw_dtype, raw_denominator, B, H, W, params, grad_out, grad_tail, additive, out, saved_luma,
                         grad_raw, grad_params, workspace, workspace_bytes, dp, stream);
}

static int backward_impl(const void* raw, int raw_dtype, float raw_denominator, int B, int H, int W,
                         const r2l_isp_params* params, const float* grad_out, const float* grad_tail,
                         const float* additive, const float* out, const float* saved_luma, float* grad_raw,
                         float* grad_params, void* workspace, size_t workspace_bytes, const r2l_isp_allreduce* dp,
                         void* stream) {
    int rc = check_common(raw, raw_dtype, B, H, W, params);
    if (rc != R2L_OK) return rc;
    if (!grad_params || !workspace || (B > 0 && !grad_out)) return R2L_ERR_NULL_POINTER;
    if (!aligned(grad_out, 4) || !aligned(grad_raw, 4) || !aligned(grad_params, 4) || !aligned(workspace, 8))
        return R2L_ERR_MISALIGNED;
    if (workspace_bytes < r2l_isp_workspace_bytes(B, H, W)) return R2L_ERR_WORKSPACE;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    if (B == 0) {
        cudaError_t e = cudaMemsetAsync(grad_params, 0, R2L_NUM_PARAM_GRADS * sizeof(float), st);
        return e == cudaSuccess ? R2L_OK : cuda_fail(e);
    }
    BwdArgs a;
    a.raw = raw; a.denom = raw_denominator; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.gout = grad_out; a.gtail = grad_tail; a.additive = additive; a.graw = grad_raw; a.partials = static_cast<float*>(workspace);
    a.out = out;
    a.luma = out ? saved_luma : nullptr;
    if (grad_tail) {                                            // a deferred tail is finished from the workspace
        a.bn_partials = reinterpret_cast<const float*>(static_cast<const char*>(workspace) + kBnPartialsOffset);
        a.bn_count = (double)B * H * W;
        a.tail_ws = reinterpret_cast<float*>(static_cast<char*>(workspace) + kBnTailOffset);
    }
    if (out && additive && !grad_tail) return R2L_ERR_BAD_ARGUMENT;   // an additive tail needs grad_tail to invert it
    if (dp && dp->world > 1) {
        a.peers = dp->peers; a.world = dp->world; a.rank = dp->rank; a.epoch = dp->epoch; a.dp_scale = dp->scale;
    }
    return launch_backward_any(a, raw_dtype, grad_params, st);
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
