// isp_core.cuh -- the arithmetic of the fused ISP forward / backward, written once for two compilers.
//
// nvcc (sm_100a) compiles this into the product kernels (isp_kernels.cu).  With R2L_HOST_EMU defined, g++
// compiles the very same CTA-level functions into a sequential emulation (tests/emu) that the CPU test-suite
// checks against the oracle -- a logic check of indexing / borders / adjoints in a container without a GPU.
// The emulation is test infrastructure only: the package never loads it.
//
// Notation: par(y,x) = 2*(y&1)+(x&1) -> R,G1,G2,B; ch(par) = {0,1,1,2}; tap t = 3*i+j of a
// 3x3 stencil reaches the neighbour p + (i-1, j-1).  Reference: processing/pipeline_torch.py:175-225.
#pragma once
#include <stdint.h>

#ifdef R2L_HOST_EMU
#include <cmath>
#include <algorithm>
#include <vector>
#include <cstring>
#define R2L_HD inline
#define R2L_FOR_THREADS(NT) for (int tid = 0; tid < (NT); ++tid)
#define R2L_SYNC()
#define R2L_ACC(arr, tid) arr[tid]
#else
#define R2L_HD __device__ __forceinline__
#define R2L_FOR_THREADS(NT) const int tid = threadIdx.x;
#define R2L_SYNC() __syncthreads()
#define R2L_ACC(arr, tid) arr
#endif

namespace r2l {

constexpr float kClipLo = 1e-5f;   // pipeline_torch.py:206
constexpr float kClipHi = 1.0f;

// ---------------------------------------------------------------------------------------------------------
// scalar helpers
// ---------------------------------------------------------------------------------------------------------
R2L_HD float fast_log2(float x) {
#ifdef R2L_HOST_EMU
    return std::log2(x);
#else
    float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#endif
}
R2L_HD float fast_exp2(float x) {
#ifdef R2L_HOST_EMU
    return std::exp2(x);
#else
    float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#endif
}
R2L_HD float fast_rcp(float x) {
#ifdef R2L_HOST_EMU
    return 1.0f / x;
#else
    float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y;
#endif
}
R2L_HD float fmaf_(float a, float b, float c) {
#ifdef R2L_HOST_EMU
    return std::fma(a, b, c);
#else
    return __fmaf_rn(a, b, c);
#endif
}
#ifndef R2L_HOST_EMU
__device__ __forceinline__ float warp_sum_all(float v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// sums lanes of equal (lane & 1): lanes 0 and 1 end up holding the even / odd totals
__device__ __forceinline__ float warp_sum_same_parity(float v) {
#pragma unroll
    for (int o = 16; o >= 2; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Sums 32 per-lane values across the warp at once: on return lane L holds the warp total of v[L].  31 shuffles
// instead of 160 (each step exchanges one half of the still-live values with the partner lane), fixed order.
__device__ __forceinline__ float warp_transpose_sum32(float (&v)[32]) {
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
        const bool up = (lane & o) != 0;
#pragma unroll
        for (int i = 0; i < o; ++i) {
            const float send = up ? v[i] : v[i + o];
            const float keep = up ? v[i + o] : v[i];
            v[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
        }
    }
    return v[0];
}
#endif
R2L_HD int imin(int a, int b) { return a < b ? a : b; }
R2L_HD int imax(int a, int b) { return a > b ? a : b; }

// whole-sample reflection (torch 'reflect'): -1 -> 1, n -> n-2.  Preserves the CFA phase of an index.
R2L_HD int mirror(int i, int n) {
    i = i < 0 ? -i : i;
    return i >= n ? 2 * (n - 1) - i : i;
}
R2L_HD int mirror_clamped(int i, int n) { return imin(imax(mirror(i, n), 0), n - 1); }

R2L_HD int par_of(int y, int x) { return ((y & 1) << 1) | (x & 1); }
R2L_HD int ch_of(int par) { return (par + 1) >> 1; }
// phase of the neighbour reached by tap t from a site of phase par
R2L_HD int par_tap(int par, int t) {
    const int i = t / 3, j = t - 3 * i;
    return par ^ (((i != 1) << 1) | (j != 1));
}

template <typename RawT> struct RawLoad;
template <> struct RawLoad<float> {
    static R2L_HD float get(const float* p, float) {
#ifdef R2L_HOST_EMU
        return *p;
#else
        return __ldg(p);
#endif
    }
};
template <> struct RawLoad<uint16_t> {
    // dataset.py:87 -- img / (2**bits - 1); a correctly rounded fp32 divide of an exactly representable integer
    static R2L_HD float get(const uint16_t* p, float denom) {
#ifdef R2L_HOST_EMU
        return (float)(*p) / denom;
#else
        return __fdiv_rn((float)__ldg(p), denom);
#endif
    }
};

// ---------------------------------------------------------------------------------------------------------
// parameter-derived tables (shared memory, rebuilt by every CTA once per launch)
// ---------------------------------------------------------------------------------------------------------
struct Params {                 // mirrors r2l_isp_params (include/r2l_isp.h)
    const float* black_level; const float* white_balance; const float* colour_correction;
    const float* gamma_correct; const float* debayer_weight; const float* sharpen_weight;
    const float* gauss_weight; const float* rgb2yuv; const float* yuv2rgb;
};

struct Tables {
    float bl[4];
    float wb[3];
    float ccm[9];
    float m1[9];
    float wd[81];
    float A[9];            // M1 * CCM * diag(wb): raw-demosaic RGB -> YUV in one 3x3
    float AW[4][3][9];     // [par][k][t]: YUV channel k at a site of phase par directly from the raw 3x3 window
    float Cb[4][3];        // black-level part of the same sum: yuv = sum AW*raw - Cb
    float AWq[4][3][9];    // adjoint gather: weight with which g_yuv[k](q - tap t) reaches a site q of phase par
    float Ws[9];
    float Wg[25];
    float M2[9];
    float gamma;
    float invg;            // 1/gamma, an fp32 reciprocal like the reference's `1 / self.gamma_correct` (:209)
};

// Four steps with a CTA barrier between them (caller provides the barriers).
R2L_HD void build_tables_step(int step, int tid, int nt, const Params& P, Tables* T) {
    if (step == 0) {
        for (int i = tid; i < 81; i += nt) T->wd[i] = P.debayer_weight[i];
        for (int i = tid; i < 25; i += nt) T->Wg[i] = P.gauss_weight[i];
        for (int i = tid; i < 9; i += nt) {
            T->Ws[i] = P.sharpen_weight[i];
            T->M2[i] = P.yuv2rgb[i];
            T->m1[i] = P.rgb2yuv[i];
            T->ccm[i] = P.colour_correction[i];
        }
        for (int i = tid; i < 4; i += nt) T->bl[i] = P.black_level[i];
        for (int i = tid; i < 3; i += nt) T->wb[i] = P.white_balance[i];
        if (tid == 0) {
            const float g = P.gamma_correct[0];
            T->gamma = g;
            T->invg = 1.0f / g;
        }
    } else if (step == 1) {
        for (int e = tid; e < 9; e += nt) {
            const int k = e / 3, c = e - 3 * k;
            float a = 0.f;
            for (int m = 0; m < 3; ++m) a = fmaf_(T->m1[k * 3 + m], T->ccm[m * 3 + c], a);
            T->A[e] = a * T->wb[c];
        }
    } else if (step == 2) {
        for (int e = tid; e < 108; e += nt) {
            const int par = e / 27, r = e - 27 * par, k = r / 9, t = r - 9 * k;
            const int cin = ch_of(par_tap(par, t));
            float a = 0.f;
            for (int c = 0; c < 3; ++c) a = fmaf_(T->A[k * 3 + c], T->wd[(c * 3 + cin) * 9 + t], a);
            T->AW[par][k][t] = a;
        }
    } else {
        for (int e = tid; e < 12; e += nt) {
            const int par = e / 3, k = e - 3 * par;
            float a = 0.f;
            for (int t = 0; t < 9; ++t) a = fmaf_(T->AW[par][k][t], T->bl[par_tap(par, t)], a);
            T->Cb[par][k] = a;
        }
        for (int e = tid; e < 108; e += nt) {
            const int par = e / 27, r = e - 27 * par, k = r / 9, t = r - 9 * k;
            T->AWq[par][k][t] = T->AW[par_tap(par, t)][k][t];
        }
    }
}

#define R2L_BUILD_TABLES(NT, P, T)                                                         \
    for (int step_ = 0; step_ < 4; ++step_) {                                              \
        { R2L_FOR_THREADS(NT) { r2l::build_tables_step(step_, tid, (NT), (P), (T)); } }    \
        R2L_SYNC();                                                                        \
    }

// ---------------------------------------------------------------------------------------------------------
// shared-memory regions: a dense h x w window of the image plane whose top-left site is (oy, ox)
// ---------------------------------------------------------------------------------------------------------
template <int H_, int W_> struct Reg {
    static constexpr int h = H_, w = W_, n = H_ * W_;
    float* s; int oy, ox;
    R2L_HD float& at(int gy, int gx) const { return s[(gy - oy) * W_ + (gx - ox)]; }
    R2L_HD bool holds(int gy, int gx) const { return gy >= oy && gy < oy + H_ && gx >= ox && gx < ox + W_; }
};

// raw window, mirrored at the image border (only the +-1 ring outside the image is ever consumed)
template <typename RawT, int NT, class RegR>
R2L_HD void phase_load_raw(int tid, const RawT* img, int H, int W, float denom, const RegR& R) {
    for (int i = tid; i < RegR::n; i += NT) {
        const int ly = i / RegR::w, lx = i - ly * RegR::w;
        const int sy = mirror_clamped(R.oy + ly, H), sx = mirror_clamped(R.ox + lx, W);
        R.s[i] = RawLoad<RawT>::get(img + (size_t)sy * W + sx, denom);
    }
}

// YUV channel k at an in-image site straight from the raw window (a1..a6 collapsed)
template <class RegR>
R2L_HD float yuv_at(const Tables* T, const RegR& R, int gy, int gx, int par, int k) {
    float a = -T->Cb[par][k];
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) a = fmaf_(T->AW[par][k][i * 3 + j], R.at(gy + i - 1, gx + j - 1), a);
    return a;
}

// Y0 = luma before sharpening; zero outside the image (the sharpen conv zero-pads, :162/:195)
template <int NT, class RegR, class RegY0>
R2L_HD void phase_y0(int tid, const Tables* T, int H, int W, const RegR& R, const RegY0& Y0) {
    for (int i = tid; i < RegY0::n; i += NT) {
        const int ly = i / RegY0::w, lx = i - ly * RegY0::w;
        const int gy = Y0.oy + ly, gx = Y0.ox + lx;
        float v = 0.f;
        if (gy >= 0 && gy < H && gx >= 0 && gx < W) v = yuv_at(T, R, gy, gx, par_of(gy, gx), 0);
        Y0.s[i] = v;
    }
}

// Y1 = sharpened luma; outside the image it holds the value at the reflected site (Gaussian reflect-pads the
// *sharpened* plane, :165/:202) -- i.e. the sharpen stencil evaluated at the mirrored coordinate.
template <int NT, class RegY0, class RegY1>
R2L_HD void phase_y1(int tid, const Tables* T, int H, int W, const RegY0& Y0, const RegY1& Y1) {
    for (int i = tid; i < RegY1::n; i += NT) {
        const int ly = i / RegY1::w, lx = i - ly * RegY1::w;
        const int gy = Y1.oy + ly, gx = Y1.ox + lx;
        float v = 0.f;
        if (gy >= -2 && gy <= H + 1 && gx >= -2 && gx <= W + 1) {
            const int my = mirror(gy, H), mx = mirror(gx, W);
#pragma unroll
            for (int a = 0; a < 3; ++a)
#pragma unroll
                for (int b = 0; b < 3; ++b) v = fmaf_(T->Ws[a * 3 + b], Y0.at(my + a - 1, mx + b - 1), v);
        }
        Y1.s[i] = v;
    }
}

template <class RegY1>
R2L_HD float gauss_at(const Tables* T, const RegY1& Y1, int gy, int gx) {
    float v = 0.f;
#pragma unroll
    for (int a = 0; a < 5; ++a)
#pragma unroll
        for (int b = 0; b < 5; ++b) v = fmaf_(T->Wg[a * 5 + b], Y1.at(gy + a - 2, gx + b - 2), v);
    return v;
}

// colour tail at one site: YUV->RGB (:203), clip (:206), gamma (:209)
struct Tone { float r[3], cl[3], l2[3], o[3]; };
R2L_HD void tone_at(const Tables* T, float y2, float u, float v, Tone& t) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float r = fmaf_(T->M2[k * 3 + 2], v, fmaf_(T->M2[k * 3 + 1], u, T->M2[k * 3] * y2));
        const float cl = fminf(fmaxf(r, kClipLo), kClipHi);
        const float l2 = fast_log2(cl);
        t.r[k] = r; t.cl[k] = cl; t.l2[k] = l2;
        t.o[k] = fast_exp2(T->invg * l2);
    }
}

// ---------------------------------------------------------------------------------------------------------
// forward CTA
// ---------------------------------------------------------------------------------------------------------
template <int TH_, int TW_, int NT_> struct FwdCfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    using RegR = Reg<TH + 8, TW + 8>;
    using RegY0 = Reg<TH + 6, TW + 6>;
    using RegY1 = Reg<TH + 4, TW + 4>;
    static constexpr int kTableFloats = (sizeof(Tables) + 3) / 4;
    static constexpr int kSmemFloats = kTableFloats + RegR::n + RegY0::n + RegY1::n;
    static constexpr size_t kSmemBytes = (size_t)kSmemFloats * 4;
};

struct FwdArgs {
    const void* raw; float denom; int B, H, W;
    Params P;
    const float* additive;     // (3,H,W) or null
    const float* affine;       // {scale[3], shift[3]} or null
    float* out;
    float* chan_partials;      // STATS only: [n_cta][kChanPitch] per-CTA sums of o and o*o per channel
    float* luma = nullptr;     // null, or [2][ceil(B/2)][H][W][2]: Y0 (plane 0) and Y1 (plane 1) of image pairs
                               // (2p, 2p+1) interleaved per site, saved for the fourth-generation backward
    // Fused train-mode BatchNorm tail (STATS kernels of the third generation, device only).  With bn_sync set the launch
    // finishes the batch statistics itself behind a grid-wide barrier -- every CTA reduces the per-CTA channel sums in
    // the same fixed order -- writes saved_affine / the running statistics (CTA 0) and normalises the tiles it wrote
    // while their lines are still in L2: one launch instead of three (forward, finish, in-place normalisation).
    unsigned* bn_sync = nullptr;           // two 8-byte ticket words {count, launch tag}: arrive, depart
    unsigned bn_gen = 0;                   // launch tag (never 0)
    double bn_count = 0.0;                 // B * H * W
    float bn_momentum = 0.f, bn_eps = 0.f;
    float* bn_running_mean = nullptr;      // may be null
    float* bn_running_var = nullptr;       // may be null
    long long* bn_num_batches = nullptr;   // may be null: nn.BatchNorm2d.num_batches_tracked, incremented by CTA 0
    float* bn_saved_affine = nullptr;      // {1/sqrt(var+eps)[3], -mean/sqrt(var+eps)[3]}
};
constexpr int kChanPitch = 8;

struct TileGrid {
    int tiles_x, tiles_y, n;
    R2L_HD void decode(int id, int TH, int TW, int& b, int& ty0, int& tx0) const {
        const int per = tiles_x * tiles_y;
        b = id / per;
        const int r = id - b * per;
        const int ty = r / tiles_x;
        ty0 = ty * TH; tx0 = (r - ty * tiles_x) * TW;
    }
};
inline TileGrid make_grid(int B, int H, int W, int TH, int TW) {
    TileGrid g; g.tiles_x = (W + TW - 1) / TW; g.tiles_y = (H + TH - 1) / TH; g.n = B * g.tiles_x * g.tiles_y;
    return g;
}

struct ChanAcc { float s[6]; };     // sum o[k], sum o[k]^2

template <class Cfg, typename RawT, bool STATS>
R2L_HD void fwd_cta(int cta, int n_cta, const FwdArgs& a, const TileGrid& grid, float* smem) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT;
    Tables* T = reinterpret_cast<Tables*>(smem);
    typename Cfg::RegR R; typename Cfg::RegY0 Y0; typename Cfg::RegY1 Y1;
    R.s = smem + Cfg::kTableFloats; Y0.s = R.s + Cfg::RegR::n; Y1.s = Y0.s + Cfg::RegY0::n;
#ifdef R2L_HOST_EMU
    std::vector<ChanAcc> cacc(NT);
    for (int i = 0; i < NT; ++i) for (int k = 0; k < 6; ++k) cacc[i].s[k] = 0.f;
#else
    ChanAcc cacc;
#pragma unroll
    for (int k = 0; k < 6; ++k) cacc.s[k] = 0.f;
#endif
    R2L_BUILD_TABLES(NT, a.P, T)
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b, ty0, tx0;
        grid.decode(tile, TH, TW, b, ty0, tx0);
        R.oy = ty0 - 4; R.ox = tx0 - 4; Y0.oy = ty0 - 3; Y0.ox = tx0 - 3; Y1.oy = ty0 - 2; Y1.ox = tx0 - 2;
        const RawT* img = static_cast<const RawT*>(a.raw) + (size_t)b * plane;
        { R2L_FOR_THREADS(NT) { phase_load_raw<RawT, NT>(tid, img, H, W, a.denom, R); } }
        R2L_SYNC();
        { R2L_FOR_THREADS(NT) { phase_y0<NT>(tid, T, H, W, R, Y0); } }
        R2L_SYNC();
        { R2L_FOR_THREADS(NT) { phase_y1<NT>(tid, T, H, W, Y0, Y1); } }
        R2L_SYNC();
        { R2L_FOR_THREADS(NT) {
            for (int i = tid; i < TH * TW; i += NT) {
                const int ly = i / TW, lx = i - ly * TW;
                const int gy = ty0 + ly, gx = tx0 + lx;
                if (gy >= H || gx >= W) continue;
                const int par = par_of(gy, gx);
                const float y2 = gauss_at(T, Y1, gy, gx);
                const float u = yuv_at(T, R, gy, gx, par, 1);
                const float v = yuv_at(T, R, gy, gx, par, 2);
                Tone t; tone_at(T, y2, u, v, t);
                const size_t pix = (size_t)gy * W + gx;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float o = t.o[k];
                    if (a.additive) o += a.additive[(size_t)k * plane + pix];
                    if (STATS) {
                        ChanAcc& c = R2L_ACC(cacc, tid);
                        c.s[k] += o;
                        c.s[3 + k] = fmaf_(o, o, c.s[3 + k]);
                    }
                    if (a.affine) o = fmaf_(o, a.affine[k], a.affine[3 + k]);
                    a.out[((size_t)b * 3 + k) * plane + pix] = o;
                }
            }
        } }
        R2L_SYNC();   // smem is rewritten by the next tile
    }
    if (STATS) {
        float* part = a.chan_partials + (size_t)cta * kChanPitch;
#ifdef R2L_HOST_EMU
        for (int k = 0; k < 6; ++k) {
            double sum = 0.0;
            for (int i = 0; i < NT; ++i) sum += cacc[i].s[k];
            part[k] = (float)sum;
        }
#else
        constexpr int NW = NT / 32;
        float* red = smem + Cfg::kTableFloats;             // regions are dead
        const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
#pragma unroll
        for (int k = 0; k < 6; ++k) {
            const float v = warp_sum_all(cacc.s[k]);
            if (lane == 0) red[warp * 6 + k] = v;
        }
        __syncthreads();
        if (tid < 6) {
            float sum = 0.f;
            for (int w = 0; w < NW; ++w) sum += red[w * 6 + tid];
            part[tid] = sum;
        }
#endif
    }
}

// ---------------------------------------------------------------------------------------------------------
// backward CTA
// ---------------------------------------------------------------------------------------------------------
// per-thread partial sums, kept in registers across all tiles of a persistent CTA
struct BwdAcc {
    float sg;        // sum G * o * log2(cl)                       -> gamma
    float wg[25];    // sum gY2(p) * Y1(p + tap)                   -> gaussian_blur.weight
    float ws[9];     // sum gY1(p) * Y0(p + tap)                   -> sharpening_filter.weight
    float q[27];     // [k][t] sum g_yuv[k](p) * raw(p + tap), sites of this thread's CFA phase only
    float p[3];      // [k]    sum g_yuv[k](p),                    same sites
};
constexpr int kStatGamma = 0, kStatWg = 1, kStatWs = 26, kStatQ = 35, kStatP = 143, kNumStats = 155;
constexpr int kStatPitch = 160;
constexpr int kMaxCtas = 2048;          // upper bound on persistent CTAs == rows of the statistics workspace
// exchange buffer of the fused all-reduce: [2 epoch parities][world][kSlotPitch] 8-byte words {value, epoch tag}
constexpr int kSlotPitch = 136, kMaxWorld = 16;
R2L_HD int stat_q_index(int k, int par, int t) { return kStatQ + (k * 4 + par) * 9 + t; }
R2L_HD int stat_p_index(int k, int par) { return kStatP + k * 4 + par; }

template <int TH_, int TW_, int NT_, bool GRAW_> struct BwdCfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    static constexpr bool GRAW = GRAW_;
    static constexpr int E = GRAW_ ? 1 : 0;
    // halo radii: g_raw(q) needs g_yuv on +-1, gY1 on +-2, gY2 / forward recompute on +-4, Y1 +-6, Y0 +-7, raw +-8
    static constexpr int rR = 7 + E, rY0 = 6 + E, rY1 = 5 + E, rF = 3 + E, rG1 = 1 + E, rG0 = E;
    using RegR = Reg<TH + 2 * rR, TW + 2 * rR>;
    using RegY0 = Reg<TH + 2 * rY0, TW + 2 * rY0>;
    using RegY1 = Reg<TH + 2 * rY1, TW + 2 * rY1>;
    using RegF = Reg<TH + 2 * rF, TW + 2 * rF>;       // gY2, gU, gV
    using RegG1 = Reg<TH + 2 * rG1, TW + 2 * rG1>;    // gY1
    using RegG0 = Reg<TH + 2 * rG0, TW + 2 * rG0>;    // gY0 (only materialised when GRAW)
    static constexpr int kTableFloats = (sizeof(Tables) + 3) / 4;
    static constexpr int kSmemFloats = kTableFloats + RegR::n + RegY0::n + RegY1::n + 3 * RegF::n + RegG1::n +
                                       (GRAW_ ? RegG0::n : 0);
    static constexpr size_t kSmemBytes = (size_t)kSmemFloats * 4;
    // every thread must always meet sites of one CFA phase in the owned-pixel loops
    static_assert(NT_ % TW_ == 0 && ((NT_ / TW_) % 2) == 0 && TW_ % 32 == 0 && TH_ % 2 == 0, "phase-stable mapping");
};

constexpr int kBnBwdBlocks = 296;                  // CTAs per channel of the BatchNorm backward statistics kernel
constexpr unsigned kTailDeferredTag = 0x7fc0b200u; // quiet NaN with a payload no arithmetic produces

struct BwdArgs {
    const void* raw; float denom; int B, H, W;
    Params P;
    const float* gout;     // (B,3,H,W)
    const float* gtail;    // null or 15 floats {gs[3], c1[3], c2[3], ysc[3], ysh[3]}: the BatchNorm tail's backward,
                           // dL/do = gs*(G - c1 - c2*yhat) with yhat = (o + additive)*ysc + ysh (eval mode: c1=c2=0)
    const float* additive; // (3,H,W) or null, only read when gtail is given
    // Deferred tail (r2l_isp_bn_backward_prepare with the full workspace): gtail's c1 / c2 entries carry kTailDeferredTag
    // and the per-CTA sums of the statistics kernel are still in the workspace; the fifth-generation kernel finishes
    // them in its prologue (no separate finish launch), older generations get a resolved copy (tail_ws) from the host.
    const float* bn_partials = nullptr;    // [3][kBnBwdBlocks][2]: sum(gy), sum(gy * yhat) per statistics CTA
    double bn_count = 0.0;                 // B * H * W
    float* tail_ws = nullptr;              // 15 floats of workspace for the resolved copy
    float* graw;           // (B,H,W) or null
    float* partials;       // [n_cta][kStatPitch]
    const float* out;      // null, or the forward's output (B,3,H,W): lets the vectorised backward skip the Gaussian /
                           // colour-tail recompute (the generic kernel ignores it)
    const float* luma = nullptr;   // null, or the Y0 / Y1 planes the forward saved (FwdArgs::luma): with `out`, the
                                   // fourth-generation backward recomputes nothing
    unsigned ticket_gen = 0;       // launch tag of the ticket word (never 0): {gen, count} -- see take_ticket()
    unsigned* ticket = nullptr;    // 8-byte aligned device word {count, gen} for the fused finish: the last CTA to publish its
                                   // partial sums turns them into the 132 gradients (null: separate finish kernel)
    float* grads = nullptr;        // destination of the fused finish
    // data-parallel exchange fused into the finish (r2l_isp_backward_dp): peers[r] = rank r's exchange buffer
    float* const* peers = nullptr;
    int world = 1, rank = 0;
    unsigned epoch = 0;
    float dp_scale = 1.f;
};

// gather of the transposed 5x5 at a (possibly padded) site q': sum_ij Wg[ij] * gY2(q' - (i-2, j-2)), in-image only
template <class RegF>
R2L_HD float gy1_gather_checked(const Tables* T, const RegF& GY2, int H, int W, int qy, int qx) {
    float v = 0.f;
    for (int a = 0; a < 5; ++a) {
        const int py = qy + 2 - a;
        if (py < 0 || py >= H) continue;
        for (int b = 0; b < 5; ++b) {
            const int px = qx + 2 - b;
            if (px < 0 || px >= W) continue;
            v = fmaf_(T->Wg[a * 5 + b], GY2.at(py, px), v);
        }
    }
    return v;
}

// reflect-pad-2 pre-images of an in-image index q: q itself, -q if q in {1,2}, 2(n-1)-q if q in {n-2,n-3}
R2L_HD int preimages2(int q, int n, int out[3]) {
    int c = 0;
    out[c++] = q;
    if (q >= 1 && q <= 2) out[c++] = -q;
    if (q <= n - 2 && q >= n - 3) out[c++] = 2 * (n - 1) - q;
    return c;
}
// reflect-pad-1 pre-images: q, -1 if q == 1, n if q == n-2
R2L_HD int preimages1(int q, int n, int out[3]) {
    int c = 0;
    out[c++] = q;
    if (q == 1) out[c++] = -1;
    if (q == n - 2) out[c++] = n;
    return c;
}

template <class Cfg, typename RawT>
R2L_HD void bwd_tile(int tid, int phase, const BwdArgs& a, const Tables* T, int b, int ty0, int tx0,
                     const typename Cfg::RegR& R, const typename Cfg::RegY0& Y0, const typename Cfg::RegY1& Y1,
                     const typename Cfg::RegF& GY2, const typename Cfg::RegF& GU, const typename Cfg::RegF& GV,
                     const typename Cfg::RegG1& GY1, const typename Cfg::RegG0& GY0, BwdAcc& acc) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    if (phase == 0) {
        const RawT* img = static_cast<const RawT*>(a.raw) + (size_t)b * plane;
        phase_load_raw<RawT, NT>(tid, img, H, W, a.denom, R);
    } else if (phase == 1) {
        phase_y0<NT>(tid, T, H, W, R, Y0);
    } else if (phase == 2) {
        phase_y1<NT>(tid, T, H, W, Y0, Y1);
    } else if (phase == 3) {
        // forward recompute of the colour tail + pull-back of grad_out to (gY2, gU, gV); gamma statistic
        using RegF = typename Cfg::RegF;
        for (int i = tid; i < RegF::n; i += NT) {
            const int ly = i / RegF::w, lx = i - ly * RegF::w;
            const int gy = GY2.oy + ly, gx = GY2.ox + lx;
            float gy2 = 0.f, gu = 0.f, gv = 0.f;
            if (gy >= 0 && gy < H && gx >= 0 && gx < W) {
                const int par = par_of(gy, gx);
                const float y2 = gauss_at(T, Y1, gy, gx);
                const float u = yuv_at(T, R, gy, gx, par, 1);
                const float v = yuv_at(T, R, gy, gx, par, 2);
                Tone t; tone_at(T, y2, u, v, t);
                const size_t pix = (size_t)gy * W + gx;
                const bool owned = gy >= ty0 && gy < ty0 + TH && gx >= tx0 && gx < tx0 + TW;
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float G = a.gout[((size_t)b * 3 + k) * plane + pix];
                    if (a.gtail) {
                        float yhat = t.o[k];
                        if (a.additive) yhat += a.additive[(size_t)k * plane + pix];
                        yhat = fmaf_(yhat, a.gtail[9 + k], a.gtail[12 + k]);
                        G = a.gtail[k] * (G - a.gtail[3 + k] - a.gtail[6 + k] * yhat);
                    }
                    const float go = G * t.o[k];
                    if (owned) acc.sg = fmaf_(go, t.l2[k], acc.sg);
                    const bool pass = (t.r[k] >= kClipLo) && (t.r[k] <= kClipHi);   // clamp backward mask, inclusive
                    const float gr = pass ? go * T->invg * fast_rcp(t.cl[k]) : 0.f;
                    gy2 = fmaf_(T->M2[k * 3 + 0], gr, gy2);
                    gu = fmaf_(T->M2[k * 3 + 1], gr, gu);
                    gv = fmaf_(T->M2[k * 3 + 2], gr, gv);
                }
            }
            GY2.s[i] = gy2; GU.s[i] = gu; GV.s[i] = gv;
        }
    } else if (phase == 4) {
        // gY1 = fold_reflect2(corr^T(gY2, Wg)) on the tile +- rG1; Wg statistic on owned sites
        using RegG1 = typename Cfg::RegG1;
        for (int i = tid; i < RegG1::n; i += NT) {
            const int ly = i / RegG1::w, lx = i - ly * RegG1::w;
            const int qy = GY1.oy + ly, qx = GY1.ox + lx;
            float v = 0.f;
            if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
                if (qy >= 3 && qy <= H - 4 && qx >= 3 && qx <= W - 4) {
#pragma unroll
                    for (int aa = 0; aa < 5; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb)
                            v = fmaf_(T->Wg[aa * 5 + bb], GY2.at(qy + 2 - aa, qx + 2 - bb), v);
                } else {
                    int ys[3], xs[3];
                    const int ny = preimages2(qy, H, ys), nx = preimages2(qx, W, xs);
                    for (int iy = 0; iy < ny; ++iy)
                        for (int ix = 0; ix < nx; ++ix) v += gy1_gather_checked(T, GY2, H, W, ys[iy], xs[ix]);
                }
                const bool owned = qy >= ty0 && qy < ty0 + TH && qx >= tx0 && qx < tx0 + TW;
                if (owned) {
                    const float g2 = GY2.at(qy, qx);
#pragma unroll
                    for (int aa = 0; aa < 5; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb)
                            acc.wg[aa * 5 + bb] = fmaf_(g2, Y1.at(qy + aa - 2, qx + bb - 2), acc.wg[aa * 5 + bb]);
                }
            }
            GY1.s[i] = v;
        }
    } else if (phase == 5) {
        // owned sites: gY0 = corr^T(gY1, Ws) (zero pad), statistics for Ws, and Q / P (everything upstream of YUV)
        for (int i = tid; i < TH * TW; i += NT) {
            const int ly = i / TW, lx = i - ly * TW;
            const int py = ty0 + ly, px = tx0 + lx;
            float g0 = 0.f;
            if (py < H && px < W) {
#pragma unroll
                for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) g0 = fmaf_(T->Ws[aa * 3 + bb], GY1.at(py + 1 - aa, px + 1 - bb), g0);
                const float g1 = GY1.at(py, px);
                const float g[3] = {g0, GU.at(py, px), GV.at(py, px)};
#pragma unroll
                for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        const int t = aa * 3 + bb;
                        acc.ws[t] = fmaf_(g1, Y0.at(py + aa - 1, px + bb - 1), acc.ws[t]);
                        const float rv = R.at(py + aa - 1, px + bb - 1);
#pragma unroll
                        for (int k = 0; k < 3; ++k) acc.q[k * 9 + t] = fmaf_(g[k], rv, acc.q[k * 9 + t]);
                    }
#pragma unroll
                for (int k = 0; k < 3; ++k) acc.p[k] += g[k];
            }
            if (Cfg::GRAW) GY0.at(py, px) = g0;
        }
        if (Cfg::GRAW) {
            // ring of width 1 around the tile: gY0 only
            constexpr int ring = 2 * (TW + 2) + 2 * TH;
            for (int i = tid; i < ring; i += NT) {
                int qy, qx;
                if (i < TW + 2) { qy = ty0 - 1; qx = tx0 - 1 + i; }
                else if (i < 2 * (TW + 2)) { qy = ty0 + TH; qx = tx0 - 1 + (i - (TW + 2)); }
                else if (i < 2 * (TW + 2) + TH) { qy = ty0 + (i - 2 * (TW + 2)); qx = tx0 - 1; }
                else { qy = ty0 + (i - 2 * (TW + 2) - TH); qx = tx0 + TW; }
                float g0 = 0.f;
                if (qy >= 0 && qy < H && qx >= 0 && qx < W) {
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb)
                            g0 = fmaf_(T->Ws[aa * 3 + bb], GY1.at(qy + 1 - aa, qx + 1 - bb), g0);
                }
                GY0.at(qy, qx) = g0;
            }
        }
    } else if (phase == 6) {
        // g_raw = fold_reflect1(corr^T(g_d, Wd)) summed over the CFA-masked channels, in YUV space via AWq
        if (Cfg::GRAW) {
            for (int i = tid; i < TH * TW; i += NT) {
                const int ly = i / TW, lx = i - ly * TW;
                const int qy = ty0 + ly, qx = tx0 + lx;
                if (qy >= H || qx >= W) continue;
                const int par = par_of(qy, qx);
                float v = 0.f;
                if (qy >= 2 && qy <= H - 3 && qx >= 2 && qx <= W - 3) {
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) {
                            const int t = aa * 3 + bb;
                            const int py = qy + 1 - aa, px = qx + 1 - bb;
                            v = fmaf_(T->AWq[par][0][t], GY0.at(py, px), v);
                            v = fmaf_(T->AWq[par][1][t], GU.at(py, px), v);
                            v = fmaf_(T->AWq[par][2][t], GV.at(py, px), v);
                        }
                } else {
                    int ys[3], xs[3];
                    const int ny = preimages1(qy, H, ys), nx = preimages1(qx, W, xs);
                    for (int iy = 0; iy < ny; ++iy)
                        for (int ix = 0; ix < nx; ++ix)
                            for (int aa = 0; aa < 3; ++aa) {
                                const int py = ys[iy] + 1 - aa;
                                if (py < 0 || py >= H) continue;
                                for (int bb = 0; bb < 3; ++bb) {
                                    const int px = xs[ix] + 1 - bb;
                                    if (px < 0 || px >= W) continue;
                                    const int t = aa * 3 + bb;
                                    v = fmaf_(T->AWq[par][0][t], GY0.at(py, px), v);
                                    v = fmaf_(T->AWq[par][1][t], GU.at(py, px), v);
                                    v = fmaf_(T->AWq[par][2][t], GV.at(py, px), v);
                                }
                            }
                }
                a.graw[(size_t)b * plane + (size_t)qy * W + qx] = v;
            }
        }
    }
}

// thread -> CFA phase of the sites it meets in the owned-pixel loops (tile origins are even)
template <class Cfg> R2L_HD int thread_par(int tid) { return (((tid / Cfg::TW) & 1) << 1) | (tid & 1); }

// ---------------------------------------------------------------------------------------------------------
// finish: statistics -> the 132 parameter gradients (runs in one small CTA, double precision)
// ---------------------------------------------------------------------------------------------------------
// S: the kNumStats sums over all CTAs.  T: tables of the same launch.
// Qr[k][par][t] = sum g_yuv[k](p) * (raw - black)(p + tap t) over sites p of phase par
R2L_HD double finish_qr(const double* S, const Tables* T, int k, int par, int t) {
    return S[stat_q_index(k, par, t)] - (double)T->bl[par_tap(par, t)] * S[stat_p_index(k, par)];
}
// Sc[m][c] = sum_p g_c[m](p) * d[c](p), g_c = M1^T g_yuv, d = demosaiced RGB before white balance
R2L_HD double finish_sc(const double* S, const Tables* T, int m, int c) {
    double s = 0.0;
    for (int k = 0; k < 3; ++k) {
        double tkc = 0.0;
        for (int par = 0; par < 4; ++par)
            for (int t = 0; t < 9; ++t)
                tkc += (double)T->wd[(c * 3 + ch_of(par_tap(par, t))) * 9 + t] * finish_qr(S, T, k, par, t);
        s += (double)T->m1[k * 3 + m] * tkc;
    }
    return s;
}
// grad e of the flat vector; Sc9 = the nine finish_sc values [m][c] (computed once by the caller)
R2L_HD float finish_grad_sc(int e, const double* S, const Tables* T, const double* Sc9) {
    const double ln2 = 0.693147180559945309417;
    if (e < 4) {                      // black_level[par']
        double s = 0.0;
        for (int par = 0; par < 4; ++par)
            for (int t = 0; t < 9; ++t)
                if (par_tap(par, t) == e)
                    for (int k = 0; k < 3; ++k) s += (double)T->AW[par][k][t] * S[stat_p_index(k, par)];
        return (float)(-s);
    }
    if (e < 7) {                      // white_balance[c] = sum_m CCM[m][c] * Sc[m][c]
        const int c = e - 4;
        double s = 0.0;
        for (int m = 0; m < 3; ++m) s += (double)T->ccm[m * 3 + c] * Sc9[m * 3 + c];
        return (float)s;
    }
    if (e < 16) {                     // colour_correction[m][c] = Sc[m][c] * wb[c]
        const int m = (e - 7) / 3, c = (e - 7) - 3 * m;
        return (float)(Sc9[m * 3 + c] * (double)T->wb[c]);
    }
    if (e < 17) {                     // gamma: -(1/g^2) * ln2 * sum G o log2(cl)
        const double ig = (double)T->invg;
        return (float)(-ig * ig * ln2 * S[kStatGamma]);
    }
    if (e < 98) {                     // debayer.weight[kd][c][t] = sum_{par: ch(par_tap)=c} sum_k A[k][kd] Qr[k][par][t]
        const int r = e - 17, kd = r / 27, c = (r - 27 * kd) / 9, t = r - 27 * kd - 9 * c;
        double s = 0.0;
        for (int par = 0; par < 4; ++par)
            if (ch_of(par_tap(par, t)) == c)
                for (int k = 0; k < 3; ++k) s += (double)T->A[k * 3 + kd] * finish_qr(S, T, k, par, t);
        return (float)s;
    }
    if (e < 107) return (float)S[kStatWs + (e - 98)];
    return (float)S[kStatWg + (e - 107)];
}
R2L_HD float finish_grad(int e, const double* S, const Tables* T) {
    double Sc9[9];
    for (int i = 0; i < 9; ++i) Sc9[i] = (e >= 4 && e < 16) ? finish_sc(S, T, i / 3, i % 3) : 0.0;
    return finish_grad_sc(e, S, T, Sc9);
}

// ---------------------------------------------------------------------------------------------------------
// CFA split (raw2rgb, pipeline_torch.py:240-283) and its adjoint, one output element per call
// ---------------------------------------------------------------------------------------------------------
// channel of a phase in the C-channel layouts: C=3 -> {0,1,1,2}; C=4 -> {0,1,2,3}
R2L_HD int mosaic_channel(int par, int C) { return C == 3 ? ch_of(par) : par; }

}  // namespace r2l

// ---------------------------------------------------------------------------------------------------------
// backward CTA driver (persistent: tiles cta, cta+n_cta, ...; statistics stay in registers until the end)
// ---------------------------------------------------------------------------------------------------------
#ifdef R2L_HOST_EMU
#include <vector>
#include <cstring>
#endif

namespace r2l {


template <class Cfg, typename RawT>
R2L_HD void bwd_cta(int cta, int n_cta, const BwdArgs& a, const TileGrid& grid, float* smem) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT;
    Tables* T = reinterpret_cast<Tables*>(smem);
    typename Cfg::RegR R; typename Cfg::RegY0 Y0; typename Cfg::RegY1 Y1;
    typename Cfg::RegF GY2, GU, GV; typename Cfg::RegG1 GY1; typename Cfg::RegG0 GY0;
    R.s = smem + Cfg::kTableFloats; Y0.s = R.s + Cfg::RegR::n; Y1.s = Y0.s + Cfg::RegY0::n;
    GY2.s = Y1.s + Cfg::RegY1::n; GU.s = GY2.s + Cfg::RegF::n; GV.s = GU.s + Cfg::RegF::n;
    GY1.s = GV.s + Cfg::RegF::n; GY0.s = GY1.s + Cfg::RegG1::n;
#ifdef R2L_HOST_EMU
    std::vector<BwdAcc> accs(NT);
    std::memset(accs.data(), 0, sizeof(BwdAcc) * NT);
#else
    BwdAcc accs;
    accs.sg = 0.f;
#pragma unroll
    for (int i = 0; i < 25; ++i) accs.wg[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 9; ++i) accs.ws[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 27; ++i) accs.q[i] = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i) accs.p[i] = 0.f;
#endif
    R2L_BUILD_TABLES(NT, a.P, T)
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b, ty0, tx0;
        grid.decode(tile, TH, TW, b, ty0, tx0);
        R.oy = ty0 - Cfg::rR; R.ox = tx0 - Cfg::rR;
        Y0.oy = ty0 - Cfg::rY0; Y0.ox = tx0 - Cfg::rY0;
        Y1.oy = ty0 - Cfg::rY1; Y1.ox = tx0 - Cfg::rY1;
        GY2.oy = GU.oy = GV.oy = ty0 - Cfg::rF; GY2.ox = GU.ox = GV.ox = tx0 - Cfg::rF;
        GY1.oy = ty0 - Cfg::rG1; GY1.ox = tx0 - Cfg::rG1;
        GY0.oy = ty0 - Cfg::rG0; GY0.ox = tx0 - Cfg::rG0;
#pragma unroll
        for (int phase = 0; phase < 7; ++phase) {
            { R2L_FOR_THREADS(NT) {
                bwd_tile<Cfg, RawT>(tid, phase, a, T, b, ty0, tx0, R, Y0, Y1, GY2, GU, GV, GY1, GY0,
                                    R2L_ACC(accs, tid));
            } }
            R2L_SYNC();
        }
    }
    float* part = a.partials + (size_t)cta * kStatPitch;
#ifdef R2L_HOST_EMU
    double S[kNumStats];
    for (int s = 0; s < kNumStats; ++s) S[s] = 0.0;
    for (int tid = 0; tid < NT; ++tid) {
        const BwdAcc& c = accs[tid];
        const int par = thread_par<Cfg>(tid);
        S[kStatGamma] += c.sg;
        for (int i = 0; i < 25; ++i) S[kStatWg + i] += c.wg[i];
        for (int i = 0; i < 9; ++i) S[kStatWs + i] += c.ws[i];
        for (int k = 0; k < 3; ++k) {
            for (int t = 0; t < 9; ++t) S[stat_q_index(k, par, t)] += c.q[k * 9 + t];
            S[stat_p_index(k, par)] += c.p[k];
        }
    }
    for (int s = 0; s < kNumStats; ++s) part[s] = (float)S[s];
#else
    // deterministic CTA reduction: warp shuffles, then a fixed-order sum over warps
    constexpr int NW = NT / 32;
    constexpr int kInd = 35, kPar = 30;                    // phase-independent / phase-bound values per thread
    float* red = smem + Cfg::kTableFloats;                 // [NW][kInd + 2*kPar], regions are dead by now
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    float* mine = red + warp * (kInd + 2 * kPar);
    {
        float v = warp_sum_all(accs.sg);
        if (lane == 0) mine[0] = v;
#pragma unroll
        for (int i = 0; i < 25; ++i) { v = warp_sum_all(accs.wg[i]); if (lane == 0) mine[1 + i] = v; }
#pragma unroll
        for (int i = 0; i < 9; ++i) { v = warp_sum_all(accs.ws[i]); if (lane == 0) mine[26 + i] = v; }
#pragma unroll
        for (int i = 0; i < 27; ++i) { v = warp_sum_same_parity(accs.q[i]); if (lane < 2) mine[kInd + lane * kPar + i] = v; }
#pragma unroll
        for (int i = 0; i < 3; ++i) { v = warp_sum_same_parity(accs.p[i]); if (lane < 2) mine[kInd + lane * kPar + 27 + i] = v; }
    }
    __syncthreads();
    for (int s = tid; s < kNumStats; s += NT) {
        float sum = 0.f;
        if (s < kStatQ) {
            for (int w = 0; w < NW; ++w) sum += red[w * (kInd + 2 * kPar) + s];
        } else {
            int k, par, slot;
            if (s < kStatP) { const int r = s - kStatQ; k = r / 36; par = (r - 36 * k) / 9; slot = k * 9 + (r - 36 * k - 9 * par); }
            else { const int r = s - kStatP; k = r / 4; par = r - 4 * k; slot = 27 + k; }
            for (int w = 0; w < NW; ++w) {
                const int wpar_row = (((w * 32) / Cfg::TW) & 1);          // row phase of every thread in warp w
                if (wpar_row == (par >> 1)) sum += red[w * (kInd + 2 * kPar) + kInd + (par & 1) * kPar + slot];
            }
        }
        part[s] = sum;
    }
#endif
}

}  // namespace r2l
