// isp_bwd3.cuh -- third-generation fused backward: branch-free phases over padded domains.
//
// Reference: the autograd graph of pipeline_torch.py:183-217.  Same adjoint algebra and the same
// float2 (image A, image B) planes / FFMA2 arithmetic as the second generation, but the schedule is rebuilt around
// what its profile showed (profiles/r02_v2_summary.md: FFMA2 was 19 % of the issued instructions, the rest was
// border slow paths, index arithmetic and instruction-cache misses of a 29 k-instruction kernel):
//   * every stencil loop is branch-free.  Border rules are realised on the DATA, not in the loops:
//       - the raw window is mirrored while it is de-interleaved into the plane (reflect-1 of the mosaic);
//       - Y0 is stored as exact zeros outside the image (the sharpen conv zero-pads, and every statistic that
//         multiplies Y0 or a gradient plane by a neighbour is then automatically restricted to the image);
//       - Y1 pad rows are computed at the mirrored row, pad columns are written by the run that owns the mirror
//         source (Gaussian reflect-2 of the sharpened plane);
//       - grad_out is read as zero outside the image, so gY2/gU/gV vanish there;
//       - adjoints are evaluated on the padded domain and the pad sites are folded onto their mirror targets by one
//         small gather pass (only tiles that touch the image border run it);
//   * work items are 1x4 site runs that are entirely inside or entirely outside the image (needs W % 4 == 0; other
//     shapes take the generic scalar kernel), so "inside the image" is one predicate per item;
//   * threads are split by warp parity into the two CFA row phases, so the phase-bound weights (B2) and
//     statistics (B7) stay in registers for every tile of the persistent CTA, with a uniform item count per warp;
//   * every correlation statistic is taken in its "flipped" form, a sum over the *stencil centre* q:
//         dWg[t] = sum_q Y1pad(q) * gY2(q - t),   dWs[t] = sum_q Y0(q) * gY1(q - t),
//         Q'[par(q)][k][t] = sum_q rawpad(q) * g_yuv[k](q - t)
//     so the window the adjoint gather loads around q is reused for the statistic and only the centre value is read
//     in addition; pad sites q outside the image belong to the nearest border tile; predicates are per run only.
//     Q' is re-indexed to the p-based layout of kStatQ when the CTA writes its partial sums (finish_grad).
#pragma once
#include <type_traits>
#include "isp_fwd2.cuh"

namespace r2l {

template <int P> R2L_HD f2& site3(f2* pl, int row, int col) { return pl[row * P + phys<P>(col)]; }
R2L_HD f2 add2v(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
R2L_HD f2 fma2vv(f2 a, f2 b, f2 c) {
#ifdef R2L_HOST_EMU
    return mk2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y));
#else
    return __ffma2_rn(a, b, c);
#endif
}
R2L_HD f2 mul2vv(f2 a, f2 b) {
#ifdef R2L_HOST_EMU
    return mk2(a.x * b.x, a.y * b.y);
#else
    return __fmul2_rn(a, b);
#endif
}


// read-once global data (grad_out, forward output, luma planes, raw centres): 128-bit load
R2L_HD f4 ld_stream4(const float* p) {
#ifdef R2L_HOST_EMU
    f4 v; v.x = p[0]; v.y = p[1]; v.z = p[2]; v.w = p[3]; return v;
#else
    // L2-only (ld.global.cg): the data is used once per CTA, the halo is re-read by neighbouring CTAs from L2.  Measured
    // on the fifth-generation backward (in-call A/B, 64 x 256 x 256): .cg 89.4 us, .cs (evict-first) 90.7 us, default 90.5 us
    return __ldcg(reinterpret_cast<const float4*>(p));
#endif
}

// Items of a region = owned rectangle (TH x G runs) first, then its halo ring: HR rows above and below (runs -1..G)
// and the runs -1 and G beside the owned rows.  The owned items carry the statistics, so putting them first gives
// every thread the same number of expensive items.
template <int TH, int G, int HR> R2L_HD void region_item(int i, int& r, int& g) {
    constexpr int GG = G + 2, TB = HR * GG;
    if (i < TH * G) { r = i / G; g = i - r * G; return; }
    int h = i - TH * G;
    if (h < TB) { const int q = h / GG; r = q - HR; g = h - q * GG - 1; }
    else if (h < 2 * TB) { h -= TB; const int q = h / GG; r = TH + q; g = h - q * GG - 1; }
    else { h -= 2 * TB; r = h >> 1; g = (h & 1) ? G : -1; }
}

struct Bwd3Acc {
    f2 sg;               // sum G o log2(cl), per stream
    float wg[25];        // flipped Gaussian-weight statistic
    float ws[9];         // flipped sharpen-weight statistic
    float q[2][3][9];    // Q'[col phase of q][k][t] for this thread's row phase
    float p[2][3];       // P[col phase][k] = sum g_yuv[k](q) over owned sites
};
constexpr int kBwd3AccFloats = 2 + 25 + 9 + 54 + 6;

// reflect-pad pre-images (isp_core.cuh: preimages1 / preimages2) are reused for the fold passes

template <int TH_, int TW_, int NT_, bool GRAW_, bool TAIL_, bool OUT_ = false> struct Bwd3Cfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    static constexpr bool GRAW = GRAW_, TAIL = TAIL_;
    static constexpr bool OUT = OUT_;     // the forward output is available: no Gaussian / colour-tail recompute
    static constexpr int G = TW / 4;
    static constexpr int PW = TW + 24;        // wide planes (raw, Y0, Y1): column index = gx - x0 + 12, run q = g + 3
    static constexpr int PN = TW + 16;        // narrow planes (F, gY1):    column index = gx - x0 + 8,  run q = g + 2
    static constexpr int RH = TH + 16, Y0H = TH + 14, Y1H = TH + 12, FH = TH + 8, G1H = TH + 4;
    static constexpr int kTableFloats = (sizeof(Tables2) + 15) / 16 * 4;
    static constexpr int kXR = RH * PW, kY0 = Y0H * PW, kF = FH * PN, kG1 = G1H * PN;
    static constexpr int kSites = kXR + kY0 + 3 * kF + kG1;
    static constexpr size_t kPlaneBytes = (size_t)kTableFloats * 4 + (size_t)kSites * 8;
    static constexpr size_t kStageOffset = (kPlaneBytes + 127) / 128 * 128;
    static constexpr size_t kStageBytes = (size_t)2 * RH * PW * 4;
    static constexpr size_t kSmemBytes = kPlaneBytes;
    static constexpr size_t kSmemBytesTma = kStageOffset + kStageBytes + 16;
    // TMA staging [image][row][pitch]: the box must start on a 16-byte boundary of the global row, so 2-byte raw
    // starts 16 elements left of the tile (and is 8 wider) where 4-byte raw starts 12 left
    // (stage x origin = x0 - (elem 2 bytes ? 16 : 12), stage pitch = PW + (elem 2 bytes ? 8 : 0))
    static constexpr int HALF = NT / 2;       // threads per CFA row phase
    static_assert(NT % 64 == 0 && TH % 2 == 0 && TW % 8 == 0, "warp-parity mapping");
    static_assert(Y1H * PW <= kXR, "Y1 aliases the raw window");
};

// shapes the third generation serves (the rest goes to the generic scalar kernel)
inline bool bwd3_shape_ok(int H, int W, int TH, int TW) {
    (void)TH; (void)TW;
    return (W % 4) == 0 && H >= 8 && W >= 8;
}

template <class Cfg, typename RawT, bool TMA = false>
R2L_HD void bwd3_cta(int cta, int n_cta, const BwdArgs& a, const TileGrid& grid, float* smem, const void* tmap = nullptr) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT, PW = Cfg::PW, PN = Cfg::PN, G = Cfg::G, HALF = Cfg::HALF;
    constexpr int GG = G + 2;                                     // runs -1 .. G
    Tables2* T2 = reinterpret_cast<Tables2*>(smem);
    Tables* T = &T2->base;
    f2* XR = reinterpret_cast<f2*>(smem + Cfg::kTableFloats);     // raw window, later Y1
    f2* Y0 = XR + Cfg::kXR;
    f2* PU = Y0 + Cfg::kY0;                                       // U, then gU
    f2* PV = PU + Cfg::kF;                                        // V, then gV
    f2* PG = PV + Cfg::kF;                                        // gY2, then gY0
    f2* GY1 = PG + Cfg::kF;                                       // gY1, then the padded g_raw
    f2* Y1 = XR;
#ifdef R2L_HOST_EMU
    std::vector<Bwd3Acc> accs(NT);
    std::memset(accs.data(), 0, sizeof(Bwd3Acc) * NT);
#else
    Bwd3Acc accs;
    {
        float* z = reinterpret_cast<float*>(&accs);
#pragma unroll
        for (int i = 0; i < kBwd3AccFloats; ++i) z[i] = 0.f;
    }
#endif
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
#ifndef R2L_HOST_EMU
    // the first tile's raw window is requested before anything else so the copy overlaps the CTA prologue
    RawT* stage = reinterpret_cast<RawT*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset + Cfg::kStageBytes);
    uint32_t tma_phase = 0;
    constexpr int SX = sizeof(RawT) == 2 ? 16 : 12, SP = sizeof(RawT) == 2 ? PW + 8 : PW;
    constexpr uint32_t kTmaBytes = 2u * Cfg::RH * SP * sizeof(RawT);
    if (TMA && threadIdx.x == 0) {
        mbar_init(mbar, 1);
        if (cta < grid.n) {
            int pb0, pb1, py0, px0;
            decode_pair_tile(grid, cta, TH, TW, a.B, pb0, pb1, py0, px0);
            tma_load_3d(stage, tmap, px0 - SX, py0 - 8, pb0, mbar, kTmaBytes);
        }
    }
#endif
    // planes start finite: never-written pad columns and out-of-image sites are read by don't-care items
    { R2L_FOR_THREADS(NT) {
        for (int i = tid; i < Cfg::kSites; i += NT) XR[i] = mk2(0.f, 0.f);
    } }
    R2L_BUILD_TABLES(NT, a.P, T)
    { R2L_FOR_THREADS(NT) { build_tables2_extra(tid, NT, T2); } }
    R2L_SYNC();                                  // also publishes the mbarrier initialisation to every thread
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b0, b1, ty0, tx0;
        decode_pair_tile(grid, tile, TH, TW, a.B, b0, b1, ty0, tx0);
        const bool dup = b1 == b0;
        const RawT* imgA = static_cast<const RawT*>(a.raw) + (size_t)b0 * plane;
        const RawT* imgB = static_cast<const RawT*>(a.raw) + (size_t)b1 * plane;
        const int e_top = ty0 == 0, e_bot = ty0 + TH >= H, e_lft = tx0 == 0, e_rgt = tx0 + TW >= W;

#ifndef R2L_HOST_EMU
        if (TMA) {
            mbar_wait(mbar, tma_phase);      // this tile's raw window has landed in the staging buffer
            tma_phase ^= 1u;
        }
#endif
        // ---- B1: raw window (rows -8..TH+7, runs -3..G+2) de-interleaved into float2 sites, mirrored --------------
        // Runs outside the image are never stored; the run that starts at column 0 also writes pad column -1
        // (= column 1) and the run that ends at column W-1 writes pad column W (= column W-2).  Pad rows read the
        // mirrored source row.
        { R2L_FOR_THREADS(NT) {
#ifndef R2L_HOST_EMU
            // B4 reads the grad_out window (rows -4..TH+3 of 3 channels x 2 images) long after this point: start
            // pulling its 128-byte lines into L2 now so those loads are L2 hits
            {
                constexpr int LPR = TW * 4 / 128;                      // lines per owned row
                constexpr int NP = Cfg::OUT ? 12 : 6;                  // planes: grad_out (and the forward output) x 3 x 2
                const int rows = Cfg::FH, nl = rows * NP * LPR;
                for (int i = tid; i < nl; i += NT) {
                    const int l = i % LPR, pr = i / LPR, pk = pr % NP, rr = pr / NP;
                    const int gy = ty0 - 4 + rr, gx = tx0 + l * 32;
                    const int pk6 = pk % 6, img = pk6 < 3 ? b0 : b1, k = pk6 < 3 ? pk6 : pk6 - 3;
                    if (gy >= 0 && gy < H && gx < W) {
                        const float* base = pk < 6 ? a.gout : a.out;
                        const float* ptr = base + ((size_t)img * 3 + k) * plane + (size_t)gy * W + gx;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr));
                    }
                }
            }
#endif
#ifdef R2L_HOST_EMU
            const RawT* stage = nullptr;
            constexpr int SX = 12, SP = PW;
#endif
            // with the forward output at hand only Y0 / Y1 around the tile are recomputed: raw rows -4..TH+3 suffice
            phase_deinterleave<PW, Cfg::RH, 8, 12, SP, SX - 12, NT, RawT, TMA>(tid, XR, stage, imgA, imgB, a.denom, ty0, tx0, H, W,
                                                                                Cfg::OUT ? 4 : 0, Cfg::OUT ? Cfg::RH - 4 : Cfg::RH);
        } }
        R2L_SYNC();
#ifndef R2L_HOST_EMU
        if (TMA && threadIdx.x == 0) {
            const int next = tile + n_cta;
            if (next < grid.n) {
                int nb0, nb1, ny0, nx0;
                decode_pair_tile(grid, next, TH, TW, a.B, nb0, nb1, ny0, nx0);
                tma_load_3d(stage, tmap, nx0 - SX, ny0 - 8, nb0, mbar, kTmaBytes);
            }
        }
#endif

        if (Cfg::OUT) {
            // ---- B2 (forward output available): luma only, rows -3..TH+2 x runs -1..G; exact zero outside the image ----
            { R2L_FOR_THREADS(NT) {
                const int rp = (tid >> 5) & 1, slot = ((tid >> 6) << 5) | (tid & 31);
                float wy[2][9];
                {
                    const f4* src = reinterpret_cast<const f4*>(T2->awy[rp]);
                    float tmp[20];
#pragma unroll
                    for (int q = 0; q < 5; ++q) { const f4 v = src[q]; tmp[4 * q] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w; }
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                        for (int t = 0; t < 9; ++t) wy[cp][t] = tmp[cp * 9 + t];
                }
                const float cb0 = T2->cbrow[rp][0], cb1 = T2->cbrow[rp][3];
                // rows -3..TH+2 of this thread's CFA row phase: (TH+6)/2 rows x GG runs
                const int r_first = -3 + ((-3 ^ rp) & 1);
                for (int i = slot; i < ((TH + 6) / 2) * GG; i += HALF) {
                    const int ri = i / GG, g = i - ri * GG - 1;
                    const int ry = r_first + 2 * ri;
                    f2 acc[4] = {mk2(-cb0, -cb0), mk2(-cb1, -cb1), mk2(-cb0, -cb0), mk2(-cb1, -cb1)};
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa) {
                        f2 in[6];
                        ld6<PW>(XR, (ry + 7 + aa) * PW + 2 * (g + 3), in);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) acc[j] = fma2s(in[j + bb], wy[j & 1][aa * 3 + bb], acc[j]);
                    }
                    const int gy = ty0 + ry, gx = tx0 + 4 * g;
                    if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[j] = mk2(0.f, 0.f);
                    }
                    st4<PW>(Y0, (ry + 7) * PW + 2 * (g + 3), acc[0], acc[1], acc[2], acc[3]);
                }
            } }
        } else {
        // ---- B2: Y0 (exact zero outside the image) on rows -7..TH+6; U,V on the F region (rows -4..TH+3, runs -1..G)
        { R2L_FOR_THREADS(NT) {
            const int rp = (tid >> 5) & 1, slot = ((tid >> 6) << 5) | (tid & 31);
            float w[2][3][9], cb[2][3];
            {
                const f4* src = reinterpret_cast<const f4*>(T2->awrow[rp]);
                float tmp[56];
#pragma unroll
                for (int q = 0; q < 14; ++q) { const f4 v = src[q]; tmp[4 * q] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w; }
#pragma unroll
                for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int t = 0; t < 9; ++t) w[cp][k][t] = tmp[cp * 27 + k * 9 + t];
#pragma unroll
                for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                    for (int k = 0; k < 3; ++k) cb[cp][k] = T2->cbrow[rp][cp * 3 + k];
            }
            // rows -4+rp, -2+rp, ... of this thread's phase: (TH+8)/2 rows x GG runs
            for (int i = slot; i < (Cfg::FH / 2) * GG; i += HALF) {
                const int ri = i / GG, g = i - ri * GG - 1;
                const int r = -4 + rp + 2 * ri;
                f2 acc[4][3];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc[j][k] = mk2(-cb[j & 1][k], -cb[j & 1][k]);
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<PW>(XR, (r + 7 + aa) * PW + 2 * (g + 3), in);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                acc[j][k] = fma2s(in[j + bb], w[j & 1][k][aa * 3 + bb], acc[j][k]);
                }
                const int gy = ty0 + r, gx = tx0 + 4 * g;
                if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][0] = mk2(0.f, 0.f);
                }
                st4<PW>(Y0, (r + 7) * PW + 2 * (g + 3), acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
                st4<PN>(PU, (r + 4) * PN + 2 * (g + 2), acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
                st4<PN>(PV, (r + 4) * PN + 2 * (g + 2), acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
            }
        } }
        // luma-only ring: rows -7..-5 and TH+4..TH+6 x runs -2..G+1, plus runs -2 and G+1 of the F rows
        { R2L_FOR_THREADS(NT) {
            constexpr int kTopBot = 6 * (G + 4), kSide = 2 * (TH + 8);
            for (int item = tid; item < kTopBot + kSide; item += NT) {
                int ry, g;
                if (item < kTopBot) {
                    const int rr = item / (G + 4);
                    g = item - rr * (G + 4) - 2;
                    ry = rr < 3 ? rr - 7 : TH + 4 + (rr - 3);
                } else {
                    const int s = item - kTopBot;
                    ry = (s >> 1) - 4;
                    g = (s & 1) ? G + 1 : -2;
                }
                const int hp = ry & 1;
                float wy[2][9];
                {
                    const f4* src = reinterpret_cast<const f4*>(T2->awy[hp]);
                    float tmp[20];
#pragma unroll
                    for (int q = 0; q < 5; ++q) { const f4 v = src[q]; tmp[4 * q] = v.x; tmp[4 * q + 1] = v.y; tmp[4 * q + 2] = v.z; tmp[4 * q + 3] = v.w; }
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                        for (int t = 0; t < 9; ++t) wy[cp][t] = tmp[cp * 9 + t];
                }
                const float cb0 = T2->cbrow[hp][0], cb1 = T2->cbrow[hp][3];
                f2 acc[4] = {mk2(-cb0, -cb0), mk2(-cb1, -cb1), mk2(-cb0, -cb0), mk2(-cb1, -cb1)};
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<PW>(XR, (ry + 7 + aa) * PW + 2 * (g + 3), in);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) acc[j] = fma2s(in[j + bb], wy[j & 1][aa * 3 + bb], acc[j]);
                }
                const int gy = ty0 + ry, gx = tx0 + 4 * g;
                if (gy < 0 || gy >= H || gx < 0 || gx >= W) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j] = mk2(0.f, 0.f);
                }
                st4<PW>(Y0, (ry + 7) * PW + 2 * (g + 3), acc[0], acc[1], acc[2], acc[3]);
            }
        } }
        }
        R2L_SYNC();

        // ---- B3: Y1 = sharpen(Y0) on rows -6..TH+5, runs -2..G+1 (overwrites the raw window) ----------------------
        // Pad rows (-2,-1,H,H+1) evaluate the stencil at the mirrored row; runs outside the image are not stored,
        // their pad columns are written by the first / last run inside the image.
        { R2L_FOR_THREADS(NT) {
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            const f2* __restrict__ Y0r = Y0;                            // B3 only reads Y0 and only writes Y1
            f2* __restrict__ Y1w = Y1;
            // rows -6..TH+5 x runs -2..G+1; with the forward output only the statistic centres are needed:
            // rows -2..TH+1 x runs -1..G (the pad ring matters at image borders only)
            constexpr int kRows = Cfg::OUT ? TH + 4 : Cfg::Y1H, kRow0 = Cfg::OUT ? 4 : 0;
            constexpr int kRuns = Cfg::OUT ? GG : G + 4, kRun0 = Cfg::OUT ? -1 : -2;
#pragma unroll 2
            for (int item = tid; item < kRows * kRuns; item += NT) {
                const int rq = item / kRuns, g = item - rq * kRuns + kRun0;
                const int rr = rq + kRow0;
                const int gy = ty0 - 6 + rr, gx = tx0 + 4 * g;
                if (gx < 0 || gx >= W) continue;
                int sr = rr;                                           // Y0 rows sr .. sr+2 <-> image rows gy-1 .. gy+1
                if ((gy < 0 && gy >= -2) || (gy >= H && gy <= H + 1)) sr = mirror(gy, H) - (ty0 - 6);
                f2 acc[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<PW>(Y0r, (sr + aa) * PW + 2 * (g + 3), in);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) acc[j] = fma2s(in[j + bb], ws[aa * 3 + bb], acc[j]);
                }
                st4<PW>(Y1w, rr * PW + 2 * (g + 3), acc[0], acc[1], acc[2], acc[3]);
                if (gx == 0) { site3<PW>(Y1w, rr, 4 * (g + 3) - 1) = acc[1]; site3<PW>(Y1w, rr, 4 * (g + 3) - 2) = acc[2]; }
                if (gx + 4 == W) { site3<PW>(Y1w, rr, 4 * (g + 3) + 4) = acc[2]; site3<PW>(Y1w, rr, 4 * (g + 3) + 5) = acc[1]; }
            }
        } }
        R2L_SYNC();

        if (Cfg::OUT) {
            // ---- B4 (forward output y available): no Gaussian / colour-tail recompute.  o = y (or (y - shift)/scale -
            // additive behind a tail); with lo = log2(o): e = cl^(1/g - 1) = 2^((1 - g) lo), log2(cl) = g lo, and the
            // clamp passed iff o lies strictly between its two clipped values (exact compare without a tail, where o is
            // bit-identical to the forward's; a 1e-4 / 1e-6 relative margin behind a tail, where o is recovered by an
            // affine inverse).  All twelve 128-bit loads of an item are issued before the first use (the lines were
            // prefetched into L2 during B1); deeper software pipelining across items / phases measured slower
            // (load-queue throttling), profiles/r01_v4_summary.md.
            { R2L_FOR_THREADS(NT) {
                float m2g[9];
                const float invg = T->invg, gam = T->gamma, one_m_g = 1.0f - T->gamma;
#pragma unroll
                for (int t = 0; t < 9; ++t) m2g[t] = T->M2[t] * invg;
                const float o_lo_exact = fast_exp2(invg * fast_log2(kClipLo));      // the forward's value of a low clip
                const float o_lo = Cfg::TAIL ? o_lo_exact * (1.0f + 1e-4f) : o_lo_exact;
                const float o_hi = Cfg::TAIL ? 1.0f - 1e-6f : 1.0f;
                Bwd3Acc& acc = R2L_ACC(accs, tid);
                for (int item = tid; item < Cfg::FH * GG; item += NT) {
                    const int rr = item / GG, g = item - rr * GG - 1;
                    const int r = rr - 4;
                    const int gy = ty0 + r, gx = tx0 + 4 * g;
                    const bool valid = gy >= 0 && gy < H && gx >= 0 && gx < W;
                    const bool owned = r >= 0 && r < TH && g >= 0 && g < G;
                    f2 gy2[4], gu[4], gv[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) { gy2[j] = mk2(0.f, 0.f); gu[j] = mk2(0.f, 0.f); gv[j] = mk2(0.f, 0.f); }
                    if (valid) {
                        const size_t pix = (size_t)gy * W + gx;
                        f4 ga[3], gb[3], ya[3], yb[3], ad[3];
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const size_t oa = ((size_t)b0 * 3 + k) * plane + pix, ob = ((size_t)b1 * 3 + k) * plane + pix;
                            ga[k] = ld_stream4(a.gout + oa);
                            ya[k] = ld_stream4(a.out + oa);
                            gb[k] = ld_stream4(a.gout + ob);
                            yb[k] = ld_stream4(a.out + ob);
                            ad[k].x = ad[k].y = ad[k].z = ad[k].w = 0.f;
                            if (Cfg::TAIL && a.additive) ad[k] = *reinterpret_cast<const f4*>(a.additive + (size_t)k * plane + pix);
                        }
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const float gak[4] = {ga[k].x, ga[k].y, ga[k].z, ga[k].w}, gbk[4] = {gb[k].x, gb[k].y, gb[k].z, gb[k].w};
                            const float yak[4] = {ya[k].x, ya[k].y, ya[k].z, ya[k].w}, ybk[4] = {yb[k].x, yb[k].y, yb[k].z, yb[k].w};
                            const float adk[4] = {ad[k].x, ad[k].y, ad[k].z, ad[k].w};
                            float t_gs = 1.f, t_c1 = 0.f, t_c2 = 0.f, t_isc = 1.f, t_osh = 0.f;
                            if (Cfg::TAIL) {
                                t_gs = a.gtail[k]; t_c1 = a.gtail[3 + k]; t_c2 = a.gtail[6 + k];
                                t_isc = 1.0f / a.gtail[9 + k]; t_osh = -a.gtail[12 + k] * t_isc;
                            }
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                float Ga = gak[j], Gb = dup ? 0.f : gbk[j];
                                f2 o = mk2(yak[j], ybk[j]);
                                if (Cfg::TAIL) {
                                    Ga = t_gs * (Ga - t_c1 - t_c2 * o.x);
                                    Gb = dup ? 0.f : t_gs * (Gb - t_c1 - t_c2 * o.y);
                                    o = mk2(fmaf_(o.x, t_isc, t_osh) - adk[j], fmaf_(o.y, t_isc, t_osh) - adk[j]);
                                }
                                const f2 lo = mk2(fast_log2(o.x), fast_log2(o.y));
                                const f2 ex = mul2s(lo, one_m_g);
                                const f2 e = mk2(fast_exp2(ex.x), fast_exp2(ex.y));
                                if (owned) acc.sg = fma2vv(mk2(Ga * o.x, Gb * o.y), mul2s(lo, gam), acc.sg);
                                const f2 gr = mk2((o.x > o_lo && o.x < o_hi) ? Ga * e.x : 0.f,
                                                  (o.y > o_lo && o.y < o_hi) ? Gb * e.y : 0.f);
                                gy2[j] = fma2s(gr, m2g[k * 3 + 0], gy2[j]);
                                gu[j] = fma2s(gr, m2g[k * 3 + 1], gu[j]);
                                gv[j] = fma2s(gr, m2g[k * 3 + 2], gv[j]);
                            }
                        }
                    }
                    st4<PN>(PG, (r + 4) * PN + 2 * (g + 2), gy2[0], gy2[1], gy2[2], gy2[3]);
                    st4<PN>(PU, (r + 4) * PN + 2 * (g + 2), gu[0], gu[1], gu[2], gu[3]);
                    st4<PN>(PV, (r + 4) * PN + 2 * (g + 2), gv[0], gv[1], gv[2], gv[3]);
                }
            } }
            R2L_SYNC();
        } else {
        // ---- B4: forward tail recomputed on the F region; grad_out pulled back to (gY2, gU, gV); gamma statistic ----
        // With cl = clip(rgb), l2 = log2(cl), e = cl^(1/g - 1) = 2^((1/g - 1) l2):  o = cl e,  dL/drgb = pass G e / g,
        // and the gamma statistic sum G o l2 = sum (G e)(cl l2) -- two MUFU per value.
        { R2L_FOR_THREADS(NT) {
            float wg[25], m2[9], m2g[9];
            const float invg = T->invg, invg1 = T->invg - 1.0f;
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
#pragma unroll
            for (int t = 0; t < 9; ++t) { m2[t] = T->M2[t]; m2g[t] = m2[t] * invg; }
            Bwd3Acc& acc = R2L_ACC(accs, tid);
            for (int item = tid; item < Cfg::FH * GG; item += NT) {
                const int rr = item / GG, g = item - rr * GG - 1;
                const int r = rr - 4;
                const int gy = ty0 + r, gx = tx0 + 4 * g;
                const bool valid = gy >= 0 && gy < H && gx >= 0 && gx < W;
                const bool owned = r >= 0 && r < TH && g >= 0 && g < G;
                // the grad_out loads have the longest latency of the phase: issue them before the Gaussian
                f4 ga[3], gb[3], ad[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    ga[k].x = ga[k].y = ga[k].z = ga[k].w = 0.f;
                    gb[k] = ga[k]; ad[k] = ga[k];
                }
                if (valid) {
                    const size_t pix = (size_t)gy * W + gx;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        ga[k] = ld_stream4(a.gout + ((size_t)b0 * 3 + k) * plane + pix);
                        if (!dup) gb[k] = ld_stream4(a.gout + ((size_t)b1 * 3 + k) * plane + pix);
                        if (Cfg::TAIL && a.additive) ad[k] = *reinterpret_cast<const f4*>(a.additive + (size_t)k * plane + pix);
                    }
                }
                f2 y2[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
#pragma unroll
                for (int aa = 0; aa < 5; ++aa) {
                    f2 in[8];
                    ld8<PW>(Y1, (r + aa + 4) * PW + 2 * (g + 3), in);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) y2[j] = fma2s(in[j + bb], wg[aa * 5 + bb], y2[j]);
                }
                f2 u[4], v[4];
                ld4<PN>(PU, (r + 4) * PN + 2 * (g + 2), u);
                ld4<PN>(PV, (r + 4) * PN + 2 * (g + 2), v);
                f2 gy2[4], gu[4], gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { gy2[j] = mk2(0.f, 0.f); gu[j] = mk2(0.f, 0.f); gv[j] = mk2(0.f, 0.f); }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    const float gak[4] = {ga[k].x, ga[k].y, ga[k].z, ga[k].w}, gbk[4] = {gb[k].x, gb[k].y, gb[k].z, gb[k].w};
                    const float adk[4] = {ad[k].x, ad[k].y, ad[k].z, ad[k].w};
                    float t_gs = 0.f, t_c1 = 0.f, t_c2 = 0.f, t_sc = 0.f, t_sh = 0.f;
                    if (Cfg::TAIL) { t_gs = a.gtail[k]; t_c1 = a.gtail[3 + k]; t_c2 = a.gtail[6 + k]; t_sc = a.gtail[9 + k]; t_sh = a.gtail[12 + k]; }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const f2 rgb = fma2s(v[j], m2[k * 3 + 2], fma2s(u[j], m2[k * 3 + 1], mul2s(y2[j], m2[k * 3])));
                        const f2 cl = mk2(fminf(fmaxf(rgb.x, kClipLo), kClipHi), fminf(fmaxf(rgb.y, kClipLo), kClipHi));
                        const f2 l2 = mk2(fast_log2(cl.x), fast_log2(cl.y));
                        const f2 ex = mul2s(l2, invg1);
                        const f2 e = mk2(fast_exp2(ex.x), fast_exp2(ex.y));
                        // scalar products: the loaded values are consumed where they are needed, not re-paired early
                        float Ga = gak[j], Gb = gbk[j];
                        if (Cfg::TAIL) {
                            const f2 o = mul2vv(cl, e);
                            const float ya = fmaf_(o.x + adk[j], t_sc, t_sh), yb = fmaf_(o.y + adk[j], t_sc, t_sh);
                            Ga = valid ? t_gs * (Ga - t_c1 - t_c2 * ya) : 0.f;
                            Gb = (valid && !dup) ? t_gs * (Gb - t_c1 - t_c2 * yb) : 0.f;
                        }
                        const f2 ge = mk2(Ga * e.x, Gb * e.y);
                        if (owned) acc.sg = fma2vv(ge, mul2vv(cl, l2), acc.sg);
                        // clamp backward mask (inclusive at both ends): the value passed iff clamping left it unchanged
                        const f2 gr = mk2(rgb.x == cl.x ? ge.x : 0.f, rgb.y == cl.y ? ge.y : 0.f);
                        gy2[j] = fma2s(gr, m2g[k * 3 + 0], gy2[j]);
                        gu[j] = fma2s(gr, m2g[k * 3 + 1], gu[j]);
                        gv[j] = fma2s(gr, m2g[k * 3 + 2], gv[j]);
                    }
                }
                st4<PN>(PG, (r + 4) * PN + 2 * (g + 2), gy2[0], gy2[1], gy2[2], gy2[3]);
                st4<PN>(PU, (r + 4) * PN + 2 * (g + 2), gu[0], gu[1], gu[2], gu[3]);
                st4<PN>(PV, (r + 4) * PN + 2 * (g + 2), gv[0], gv[1], gv[2], gv[3]);
            }
        } }
        R2L_SYNC();
        }

        // ---- B5: gY1 = fold_reflect2(corr^T(gY2, Wg)) on rows -2..TH+1, runs -1..G (zero outside the image); dWg -----------
        // The reflect-2 padding of the sharpened plane (pipeline_torch.py:165/202) folds pad rows -1,-2 / H,H+1 onto rows
        // 1,2 / H-2,H-3 and likewise for columns.  A pad site's adjoint value only involves gY2 entries that lie inside
        // the folded-onto site's own 5x5 window, and its Y1 value is the folded-onto site's own, so both the fold and the
        // pad sites' share of the dWg statistic are a few extra products inside the items of those rows / columns
        // (uniform per item): no pad items, no fold pass, no extra barrier.
        //   pad row of target row ty, window row d (image row ty-2+d)  ->  tap row a' = 4 - 2 ty - d (top), mirrored below
        //   pad column of target column tx: taps b <= 2 - tx (left), b >= 5 - j (right, j = site index in the last run)
        { R2L_FOR_THREADS(NT) {
            float wg[25];
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
            Bwd3Acc& acc = R2L_ACC(accs, tid);
            // one work item = NR runs of one row processed in lockstep (runs g and g + G/2): the statistic chains then
            // run over 4*NR sites before their horizontal add, and two independent windows are in flight per thread
            auto b5 = [&](auto NRc, int r, int g, bool stat) {
                constexpr int NR = decltype(NRc)::value;
                const int qy = ty0 + r;
                f2 c[NR][4], out[NR][4];
                bool inside[NR], lft[NR], rgt[NR];
                bool any = false;
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    const int qx = tx0 + 4 * (g + u * (G / 2));
                    inside[u] = qy >= 0 && qy < H && qx >= 0 && qx < W;
                    lft[u] = qx == 0; rgt[u] = qx + 4 == W;
                    any |= inside[u];
#pragma unroll
                    for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                }
                if (any) {
#pragma unroll
                    for (int u = 0; u < NR; ++u) {
                        ld4<PW>(Y1, (r + 6) * PW + 2 * (g + u * (G / 2) + 3), c[u]);
                        if (!inside[u]) {          // pad sites are accounted for by the folded-onto sites (below)
#pragma unroll
                            for (int j = 0; j < 4; ++j) c[u][j] = mk2(0.f, 0.f);
                        }
                    }
                    // row type of the folded-onto rows: 1 -> row 1, 2 -> row 2, 3 -> row H-2, 4 -> row H-3
                    const int rt = qy == 1 ? 1 : (qy == 2 ? 2 : (qy == H - 2 ? 3 : (qy == H - 3 ? 4 : 0)));
#pragma unroll
                    for (int d = 0; d < 5; ++d) {
                        f2 row[NR][8];                                  // gY2 row q.y - 2 + d, columns q.x - 2 .. q.x + 5
#pragma unroll
                        for (int u = 0; u < NR; ++u) ld8<PN>(PG, (r + 2 + d) * PN + 2 * (g + u * (G / 2) + 2), row[u]);
                        const int aa = 4 - d;                           // tap row whose transpose reaches this row
#pragma unroll
                        for (int u = 0; u < NR; ++u)
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int bb = 0; bb < 5; ++bb) out[u][j] = fma2s(row[u][j + 4 - bb], wg[aa * 5 + bb], out[u][j]);
                        float tsum[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
                        if (stat) {
#pragma unroll
                            for (int bb = 0; bb < 5; ++bb) {
                                f2 t = mul2vv(c[0][0], row[0][4 - bb]);
#pragma unroll
                                for (int u = 0; u < NR; ++u)
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (u | j) t = fma2vv(c[u][j], row[u][j + 4 - bb], t);
                                tsum[bb] = t.x + t.y;
                                acc.wg[aa * 5 + bb] += tsum[bb];
                            }
                        }
                        // pad columns of this window row, for tap row A (the regular one, or a pad row's)
                        auto col_extra = [&](const int A) {      // A is a constant after unrolling / inlining
#pragma unroll
                            for (int u = 0; u < NR; ++u) {
                                if (lft[u]) {        // pad columns -1 (-> site 1, taps b = 0,1) and -2 (-> site 2, tap b = 0)
                                    out[u][1] = fma2s(row[u][3], wg[A * 5 + 0], fma2s(row[u][2], wg[A * 5 + 1], out[u][1]));
                                    out[u][2] = fma2s(row[u][2], wg[A * 5 + 0], out[u][2]);
                                    if (stat) {
                                        const f2 t0 = fma2vv(c[u][1], row[u][3], mul2vv(c[u][2], row[u][2]));
                                        const f2 t1 = mul2vv(c[u][1], row[u][2]);
                                        acc.wg[A * 5 + 0] += t0.x + t0.y;
                                        acc.wg[A * 5 + 1] += t1.x + t1.y;
                                    }
                                }
                                if (rgt[u]) {        // pad columns W (-> site 2, taps b = 3,4) and W+1 (-> site 1, tap b = 4)
                                    out[u][2] = fma2s(row[u][5], wg[A * 5 + 3], fma2s(row[u][4], wg[A * 5 + 4], out[u][2]));
                                    out[u][1] = fma2s(row[u][5], wg[A * 5 + 4], out[u][1]);
                                    if (stat) {
                                        const f2 t3 = mul2vv(c[u][2], row[u][5]);
                                        const f2 t4 = fma2vv(c[u][2], row[u][4], mul2vv(c[u][1], row[u][5]));
                                        acc.wg[A * 5 + 3] += t3.x + t3.y;
                                        acc.wg[A * 5 + 4] += t4.x + t4.y;
                                    }
                                }
                            }
                        };
                        // a pad row reaches this window row through tap row A2
                        auto row_extra = [&](const int A2) {
#pragma unroll
                            for (int u = 0; u < NR; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
#pragma unroll
                                    for (int bb = 0; bb < 5; ++bb) out[u][j] = fma2s(row[u][j + 4 - bb], wg[A2 * 5 + bb], out[u][j]);
                            if (stat) {
#pragma unroll
                                for (int bb = 0; bb < 5; ++bb) acc.wg[A2 * 5 + bb] += tsum[bb];
                            }
                            col_extra(A2);                                   // corner pads
                        };
                        col_extra(aa);
                        if (d == 0 && rt == 2) row_extra(0);
                        if (d == 1 && rt == 1) row_extra(1);
                        if (d == 2 && rt == 1) row_extra(0);
                        if (d == 2 && rt == 3) row_extra(4);
                        if (d == 3 && rt == 3) row_extra(3);
                        if (d == 4 && rt == 4) row_extra(4);
                    }
                }
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    if (!inside[u]) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                    }
                    st4<PN>(GY1, (r + 2) * PN + 2 * (g + u * (G / 2) + 2), out[u][0], out[u][1], out[u][2], out[u][3]);
                }
            };
            for (int item = tid; item < TH * (G / 2); item += NT)       // owned rectangle: one paired item per thread
                b5(std::integral_constant<int, 2>(), item / (G / 2), item % (G / 2), true);
            for (int item = TH * G + tid; item < Cfg::G1H * GG; item += NT) {      // halo ring, single runs, no statistic
                int r, g;
                region_item<TH, G, 2>(item, r, g);
                b5(std::integral_constant<int, 1>(), r, g, false);
            }
        } }
        R2L_SYNC();

        // ---- B6: gY0 = corr^T(gY1, Ws) (zero pad) on rows -1..TH, runs -1..G, zero outside the image; Ws statistic ----
        { R2L_FOR_THREADS(NT) {
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            Bwd3Acc& acc = R2L_ACC(accs, tid);
            auto b6 = [&](auto NRc, int r, int g, bool owned) {
                constexpr int NR = decltype(NRc)::value;
                f2 c[NR][4], out[NR][4];
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    ld4<PW>(Y0, (r + 7) * PW + 2 * (g + u * (G / 2) + 3), c[u]);
#pragma unroll
                    for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                }
#pragma unroll
                for (int d = 0; d < 3; ++d) {
                    f2 row[NR][6];                                      // gY1 row q.y - 1 + d, columns q.x - 1 .. q.x + 4
#pragma unroll
                    for (int u = 0; u < NR; ++u) ld6<PN>(GY1, (r + 1 + d) * PN + 2 * (g + u * (G / 2) + 2), row[u]);
                    const int aa = 2 - d;
#pragma unroll
                    for (int u = 0; u < NR; ++u)
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) out[u][j] = fma2s(row[u][j + 2 - bb], ws[aa * 3 + bb], out[u][j]);
                    if (owned) {
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) {
                            f2 t = mul2vv(c[0][0], row[0][2 - bb]);
#pragma unroll
                            for (int u = 0; u < NR; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    if (u | j) t = fma2vv(c[u][j], row[u][j + 2 - bb], t);
                            acc.ws[aa * 3 + bb] += t.x + t.y;
                        }
                    }
                }
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    const int qy = ty0 + r, qx = tx0 + 4 * (g + u * (G / 2));
                    if (qy < 0 || qy >= H || qx < 0 || qx >= W) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                    }
                    st4<PN>(PG, (r + 4) * PN + 2 * (g + u * (G / 2) + 2), out[u][0], out[u][1], out[u][2], out[u][3]);
                }
            };
            for (int item = tid; item < TH * (G / 2); item += NT)       // owned rectangle: one paired item per thread
                b6(std::integral_constant<int, 2>(), item / (G / 2), item % (G / 2), true);
            for (int item = TH * G + tid; item < (TH + 2) * GG; item += NT) {      // halo ring, single runs
                int r, g;
                region_item<TH, G, 1>(item, r, g);
                b6(std::integral_constant<int, 1>(), r, g, false);
            }
        } }
        R2L_SYNC();

        // ---- B7: Q' / P statistics and g_raw from the (gY0, gU, gV) windows -------------------------------------------
        // The reflect-1 padding of the mosaic (pipeline_torch.py:233) folds the pad ring onto rows 1 / H-2 and columns
        // 1 / W-2.  A pad site's contribution only involves gradient values the folded-onto site already holds in its
        // window (pad row -1 reaches row 0 only, through the tap row a = 0; pad column -1 reaches column 0 through
        // b = 0; ...) and its raw value is the folded-onto site's own, so the fold is a few extra products inside
        // the items of those rows / columns: no pad items, no staging, no second pass, stores go straight to global.
        { R2L_FOR_THREADS(NT) {
            const int rp = (tid >> 5) & 1, slot = ((tid >> 6) << 5) | (tid & 31);
            Bwd3Acc& acc = R2L_ACC(accs, tid);
            float awq[2][3][9];
            if (Cfg::GRAW) {
#pragma unroll
                for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                    for (int k = 0; k < 3; ++k)
#pragma unroll
                        for (int t = 0; t < 9; ++t) awq[cp][k][t] = T->AWq[2 * rp + cp][k][t];
            }
            // owned rows of this thread's row phase: TH/2 rows x G runs, the same count for every thread
            for (int i = slot; i < (TH / 2) * G; i += HALF) {
                const int ri = i / G, g = i - ri * G;
                const int r = rp + 2 * ri;
                const int qy = ty0 + r, qx = tx0 + 4 * g;
                if (qy >= H || qx >= W) continue;                       // partial tiles: nothing there (all gradients zero)
                const bool f_top = qy == 1, f_bot = qy == H - 2, f_lft = qx == 0, f_rgt = qx + 4 == W;
                // raw centres of the 4 sites
                f2 c[4];
                if (TMA && sizeof(RawT) == 4) {
                    const f4 xa = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgA) + (size_t)qy * W + qx);
                    const f4 xb = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgB) + (size_t)qy * W + qx);
                    c[0] = mk2(xa.x, xb.x); c[1] = mk2(xa.y, xb.y); c[2] = mk2(xa.z, xb.z); c[3] = mk2(xa.w, xb.w);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        c[j] = mk2(RawLoad<RawT>::get(imgA + (size_t)qy * W + qx + j, a.denom),
                                   RawLoad<RawT>::get(imgB + (size_t)qy * W + qx + j, a.denom));
                }
                f2 graw[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    f2* pl = k == 0 ? PG : (k == 1 ? PU : PV);
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        f2 row[6];                                      // g_yuv[k] row q.y - 1 + d, columns q.x - 1 .. q.x + 4
                        ld6<PN>(pl, (r + 3 + d) * PN + 2 * (g + 2), row);
                        const int aa = 2 - d;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) {
                                const f2 t = row[j + 2 - bb];                                   // g_yuv[k](q - (a-1, b-1))
                                acc.q[j & 1][k][aa * 3 + bb] = fmaf_(c[j].x, t.x, fmaf_(c[j].y, t.y, acc.q[j & 1][k][aa * 3 + bb]));
                                if (Cfg::GRAW) graw[j] = fma2s(t, awq[j & 1][k][aa * 3 + bb], graw[j]);
                            }
                        if (d == 1) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) acc.p[j & 1][k] += row[j + 1].x + row[j + 1].y;
                        }
                        // ---- folded pad contributions (uniform per item; rows 1 / H-2 and the first / last run only) ----
                        // pad row above (d == 0 is image row 0, reached through tap row 0) / below (d == 2, tap row 2)
                        if ((f_top && d == 0) || (f_bot && d == 2)) {
                            const int ap = d;                            // tap row of the pad site: 0 above, 2 below
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int bb = 0; bb < 3; ++bb) {
                                    const f2 t = row[j + 2 - bb];
                                    acc.q[j & 1][k][ap * 3 + bb] = fmaf_(c[j].x, t.x, fmaf_(c[j].y, t.y, acc.q[j & 1][k][ap * 3 + bb]));
                                    if (Cfg::GRAW) graw[j] = fma2s(t, awq[j & 1][k][ap * 3 + bb], graw[j]);
                                }
                        }
                        // pad column left of the image folds onto site j = 1 (column 1) through b = 0, value at column 0
                        if (f_lft) {
                            const f2 t = row[1];
                            acc.q[1][k][aa * 3 + 0] = fmaf_(c[1].x, t.x, fmaf_(c[1].y, t.y, acc.q[1][k][aa * 3 + 0]));
                            if (Cfg::GRAW) graw[1] = fma2s(t, awq[1][k][aa * 3 + 0], graw[1]);
                            if ((f_top && d == 0) || (f_bot && d == 2)) {            // corner pad
                                acc.q[1][k][d * 3 + 0] = fmaf_(c[1].x, t.x, fmaf_(c[1].y, t.y, acc.q[1][k][d * 3 + 0]));
                                if (Cfg::GRAW) graw[1] = fma2s(t, awq[1][k][d * 3 + 0], graw[1]);
                            }
                        }
                        // pad column right of the image folds onto site j = 2 (column W-2) through b = 2, value at column W-1
                        if (f_rgt) {
                            const f2 t = row[4];
                            acc.q[0][k][aa * 3 + 2] = fmaf_(c[2].x, t.x, fmaf_(c[2].y, t.y, acc.q[0][k][aa * 3 + 2]));
                            if (Cfg::GRAW) graw[2] = fma2s(t, awq[0][k][aa * 3 + 2], graw[2]);
                            if ((f_top && d == 0) || (f_bot && d == 2)) {            // corner pad
                                acc.q[0][k][d * 3 + 2] = fmaf_(c[2].x, t.x, fmaf_(c[2].y, t.y, acc.q[0][k][d * 3 + 2]));
                                if (Cfg::GRAW) graw[2] = fma2s(t, awq[0][k][d * 3 + 2], graw[2]);
                            }
                        }
                    }
                }
                if (Cfg::GRAW) {
                    float* pa = a.graw + (size_t)b0 * plane + (size_t)qy * W + qx;
                    f4 va; va.x = graw[0].x; va.y = graw[1].x; va.z = graw[2].x; va.w = graw[3].x;
                    *reinterpret_cast<f4*>(pa) = va;
                    if (!dup) {
                        float* pb = a.graw + (size_t)b1 * plane + (size_t)qy * W + qx;
                        f4 vb; vb.x = graw[0].y; vb.y = graw[1].y; vb.z = graw[2].y; vb.w = graw[3].y;
                        *reinterpret_cast<f4*>(pb) = vb;
                    }
                }
            }
        } }
        R2L_SYNC();   // planes are rewritten by the next tile
    }

    // ---- CTA reduction of the per-thread statistics into the kStat* layout (deterministic, fixed order) -------------
    // warps reduce their lanes (all lanes of a warp share the CFA row phase), then one thread per statistic adds the
    // warp totals in warp order
    float* part = a.partials + (size_t)cta * kStatPitch;
    constexpr int NW = NT / 32;
    constexpr int RP = kBwd3AccFloats + 1;
    float* red = reinterpret_cast<float*>(XR);                       // [NW][RP]
    R2L_SYNC();                                                      // planes are dead
#ifdef R2L_HOST_EMU
    for (int w = 0; w < NW; ++w)
        for (int i = 0; i < kBwd3AccFloats; ++i) {
            float sum = 0.f;
            for (int l = 0; l < 32; ++l) sum += reinterpret_cast<const float*>(&accs[w * 32 + l])[i];
            red[w * RP + i] = sum;
        }
#else
    {
        static_assert(kBwd3AccFloats == 96, "three groups of 32 running sums");
        const float* src = reinterpret_cast<const float*>(&accs);
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int grp = 0; grp < 3; ++grp) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = src[grp * 32 + i];
            red[warp * RP + grp * 32 + lane] = warp_transpose_sum32(v);
        }
    }
#endif
    R2L_SYNC();
    { R2L_FOR_THREADS(NT) {
        for (int s = tid; s < kNumStats; s += NT) {
            float sum = 0.f;
            if (s == kStatGamma) {
                for (int w = 0; w < NW; ++w) sum += red[w * RP] + red[w * RP + 1];
            } else if (s < kStatQ) {                                 // Wg, Ws: same slot in every thread
                for (int w = 0; w < NW; ++w) sum += red[w * RP + s + 1];
            } else {
                int k, parp, tt = 0;
                bool is_q;
                if (s < kStatP) { const int rI = s - kStatQ; k = rI / 36; parp = (rI - 36 * k) / 9; tt = rI - 36 * k - 9 * parp; is_q = true; }
                else { const int rI = s - kStatP; k = rI / 4; parp = rI - 4 * k; is_q = false; }
                // Q[k][par(p)][t] = Q'[par(q) = par_tap(par(p), t)][k][t];  P is already p-indexed (p = q)
                const int parq = is_q ? par_tap(parp, tt) : parp;
                const int rpq = parq >> 1, cpq = parq & 1;
                const int off = is_q ? 36 + (cpq * 3 + k) * 9 + tt : 36 + 54 + cpq * 3 + k;
                for (int w = rpq; w < NW; w += 2) sum += red[w * RP + off];      // warps of row phase rpq
            }
            part[s] = sum;
        }
    } }
}

}  // namespace r2l
