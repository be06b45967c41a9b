// isp_fwd3.cuh -- third-generation fused forward: the border rules live on the data, four barriers per tile.
//
// Reference: ParametrizedProcessing.forward, pipeline_torch.py:175-225.  Same float2 (image A, image B) planes, FFMA2
// arithmetic and chunk de-interleaved rows as isp_fwd2.cuh; what changes is that no fix-up pass (and none of their
// barriers) is left -- the second generation spent ~15 % of its samples waiting at the barriers of three tiny
// border passes (profiles/r03_summary.md):
//   * F1 de-interleaves the TMA-staged raw window and mirrors it on the way (reflect-1 of the mosaic, :233): pad
//     rows read the mirrored staging row, pad columns -1 / W are written by the first / last run inside the image;
//   * F2 stores Y0 as exact zero outside the image (the sharpen conv zero-pads, :162);
//   * F3 evaluates pad rows of Y1 at the mirrored row and lets the border-adjacent run write the pad columns
//     (the Gaussian reflect-pads the sharpened plane, :165/:202);
//   * F4 = Gaussian on 2x4 register micro-tiles + YUV->RGB + clip + gamma [+ additive] [+ affine | channel sums].
// Work items are 1x4 site runs entirely inside or outside the image: needs W % 4 == 0 (other shapes: isp_fwd2.cuh).
#pragma once
#include "isp_fwd2.cuh"

namespace r2l {

// Row pitch of the haloed planes = LW sites in use + 4 pad sites, for the same reason as the backward's (isp_bwd4.cuh: rows
// of 80 sites all start on one bank and the halo-ring items pay 16-way conflicts): forward 41.5 -> 39.6 us.  The TMA staging
// buffer keeps the unpadded width (its box rows must be multiples of 16 bytes for 2-byte raw as well).
#ifndef R2L_FWD_P_PAD
#define R2L_FWD_P_PAD 20
#endif
// 4 sites x (image A, image B) of a saved luma plane: 32 contiguous, 32-byte aligned bytes written by ONE 256-bit store
// (sm_100: STG.E.256) -- as two 128-bit stores each warp instruction filled only half of every sector it touched
R2L_HD void st_luma4(float* dst, f2 a0, f2 a1, f2 a2, f2 a3) {
#ifdef R2L_HOST_EMU
    st2(reinterpret_cast<f2*>(dst), a0, a1);
    st2(reinterpret_cast<f2*>(dst) + 2, a2, a3);
#else
    asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 :: "l"(dst), "f"(a0.x), "f"(a0.y), "f"(a1.x), "f"(a1.y), "f"(a2.x), "f"(a2.y), "f"(a3.x), "f"(a3.y) : "memory");
#endif
}

template <int TH_, int TW_, int NT_> struct Fwd3Cfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    static constexpr int LW = TW + 16;                // sites of a haloed row: column index = gx - x0 + 8 (also the TMA box width)
    static constexpr int P = TW + R2L_FWD_P_PAD;      // row pitch (sites) of the haloed planes (>= LW, a multiple of 4)
    static constexpr int RH = TH + 8, Y0H = TH + 6, Y1H = TH + 4;
    static constexpr int G = TW / 4;                  // 4-site runs per tile row
    static constexpr int kTableFloats = (sizeof(Tables2) + 15) / 16 * 4;
    static constexpr int kXR = RH * P, kY0 = Y0H * P, kUV = TH * TW;          // sizes in float2 sites
    static constexpr int kSites = kXR + kY0 + 2 * kUV;
    static constexpr size_t kPlaneBytes = (size_t)kTableFloats * 4 + (size_t)kSites * 8;
    static constexpr size_t kStageOffset = (kPlaneBytes + 127) / 128 * 128;       // TMA destination: 128-byte aligned
    static constexpr size_t kStageBytes = (size_t)2 * RH * LW * 4;                // [image][row][LW] of the raw element
    static constexpr size_t kSmemBytes = kPlaneBytes;
    static constexpr size_t kSmemBytesTma = kStageOffset + kStageBytes + 16;      // + staging + mbarrier
    static constexpr int HALF = NT / 2;
    static_assert(TW % 8 == 0 && TH % 2 == 0 && NT % 64 == 0, "warp-parity mapping");
    static_assert(Y1H * P <= kXR, "Y1 aliases the raw window");
};

inline bool fwd3_shape_ok(int H, int W) { return (W % 4) == 0 && H >= 4 && W >= 8; }

// TAIL: additive layer and/or affine epilogue may be present (runtime pointers); STATS: per-channel sums for the
// train-mode BatchNorm tail (additive may be present, affine is applied by a later pass)
template <class Cfg, typename RawT, bool STATS, bool TAIL, bool TMA>
R2L_HD void fwd3_cta(int cta, int n_cta, const FwdArgs& a, const TileGrid& grid, float* smem, const void* tmap = nullptr) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT, P = Cfg::P, G = Cfg::G, HALF = Cfg::HALF;
    constexpr int GG = G + 2;                                     // runs -1 .. G
    Tables2* T2 = reinterpret_cast<Tables2*>(smem);
    Tables* T = &T2->base;
    f2* XR = reinterpret_cast<f2*>(smem + Cfg::kTableFloats);    // raw window, later Y1
    f2* Y0 = XR + Cfg::kXR;
    f2* U = Y0 + Cfg::kY0;
    f2* V = U + Cfg::kUV;
    f2* Y1 = XR;
#ifdef R2L_HOST_EMU
    std::vector<ChanAcc> cacc(NT);
    for (int i = 0; i < NT; ++i) for (int k = 0; k < 6; ++k) cacc[i].s[k] = 0.f;
#else
    ChanAcc cacc;
#pragma unroll
    for (int k = 0; k < 6; ++k) cacc.s[k] = 0.f;
#endif
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
#ifndef R2L_HOST_EMU
    pdl_launch_dependents();                     // the next kernel's CTAs may take the slots this grid frees
    bool pdl_pending = true;                     // until the first tile reaches its first global store
    // the first tile's raw window is requested before anything else so the copy overlaps the CTA prologue
    RawT* stage = reinterpret_cast<RawT*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset + Cfg::kStageBytes);
    uint32_t tma_phase = 0;
    constexpr uint32_t kTmaBytes = 2u * Cfg::RH * Cfg::LW * sizeof(RawT);
    if (TMA && threadIdx.x == 0) {
        mbar_init(mbar, 1);
        if (cta < grid.n) {
            int pb0, pb1, py0, px0;
            decode_pair_tile(grid, cta, TH, TW, a.B, pb0, pb1, py0, px0);
            tma_load_3d(stage, tmap, px0 - 8, py0 - 4, pb0, mbar, kTmaBytes);
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
#ifndef R2L_HOST_EMU
        if (TMA) {
            mbar_wait(mbar, tma_phase);      // this tile's raw window has landed in the staging buffer
            tma_phase ^= 1u;
        }
#endif
        // ---- F1: raw window (rows -4..TH+3, runs -2..G+1) de-interleaved into float2 sites, mirrored ---------------
        { R2L_FOR_THREADS(NT) {
#ifdef R2L_HOST_EMU
            const RawT* stage = nullptr;
#endif
            phase_deinterleave<P, Cfg::RH, 4, 8, Cfg::LW, 0, NT, RawT, TMA, Cfg::LW>(tid, XR, stage, imgA, imgB, a.denom, ty0, tx0, H, W);
        } }
        R2L_SYNC();
#ifndef R2L_HOST_EMU
        // Dependent launch: the prologue, the first raw window (raw is never written by this library's kernels) and its
        // de-interleave may run while the kernel before this one drains; everything this kernel WRITES (luma planes in
        // F2, the output in F4, the channel sums) may still be read by that kernel, so the first store waits for it.
        if (pdl_pending) { pdl_wait(); pdl_pending = false; }
        if (TMA && threadIdx.x == 0) {
            const int next = tile + n_cta;
            if (next < grid.n) {
                int nb0, nb1, ny0, nx0;
                decode_pair_tile(grid, next, TH, TW, a.B, nb0, nb1, ny0, nx0);
                tma_load_3d(stage, tmap, nx0 - 8, ny0 - 4, nb0, mbar, kTmaBytes);
            }
        }
#endif

        // ---- F2: Y0 (exact zero outside the image) on rows -3..TH+2, runs -1..G; U and V on the tile ---------------
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
            // owned rows of this thread's CFA row phase: TH/2 rows x G runs
            for (int i = slot; i < (TH / 2) * G; i += HALF) {
                const int ri = i / G, g = i - ri * G;
                const int r = rp + 2 * ri, q = 2 + g;
                f2 acc[4][3];
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int k = 0; k < 3; ++k) acc[j][k] = mk2(-cb[j & 1][k], -cb[j & 1][k]);
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<P>(XR, (r + 3 + aa) * P + 2 * q, in);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                            for (int k = 0; k < 3; ++k)
                                acc[j][k] = fma2s(in[j + bb], w[j & 1][k][aa * 3 + bb], acc[j][k]);
                }
                if (ty0 + r >= H || tx0 + 4 * g >= W) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j][0] = mk2(0.f, 0.f);
                }
                else if (a.luma) {                     // owned, in the image: keep Y0 for the backward's dWs statistic
                    float* dst = a.luma + (((size_t)(b0 >> 1) * H + (ty0 + r)) * W + (tx0 + 4 * g)) * 2;
                    st_luma4(dst, acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
                }
                st4<P>(Y0, (r + 3) * P + 2 * q, acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
                st4<TW>(U, r * TW + 2 * g, acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
                st4<TW>(V, r * TW + 2 * g, acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
            }
        } }
        // luma-only ring: rows -3..-1 and TH..TH+2 x runs -1..G, plus runs -1 and G of the tile rows
        { R2L_FOR_THREADS(NT) {
            constexpr int kTopBot = 6 * GG, kSide = 2 * TH;
            for (int item = tid; item < kTopBot + kSide; item += NT) {
                int ry, g;
                if (item < kTopBot) {
                    const int rr = item / GG;
                    g = item - rr * GG - 1;
                    ry = rr < 3 ? rr - 3 : TH + rr - 3;
                } else {
                    const int s = item - kTopBot;
                    ry = s >> 1;
                    g = (s & 1) ? G : -1;
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
                const int q = 2 + g;
                f2 acc[4] = {mk2(-cb0, -cb0), mk2(-cb1, -cb1), mk2(-cb0, -cb0), mk2(-cb1, -cb1)};
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<P>(XR, (ry + 3 + aa) * P + 2 * q, in);
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
                st4<P>(Y0, (ry + 3) * P + 2 * q, acc[0], acc[1], acc[2], acc[3]);
            }
        } }
        R2L_SYNC();

        // ---- F3: Y1 = sharpen(Y0) on rows -2..TH+1, runs -1..G; overwrites the raw window --------------------------------
        // Items are 3x4 register micro-tiles (three Y1 rows of one run from five Y0 rows: 20 window loads for 108 FFMA2
        // instead of 36).  A pad row of Y1 (Gaussian reflect-2 of the sharpened plane) is the stencil at the MIRRORED row;
        // a micro-tile that holds one (first / last block of a border tile) takes its rows one by one.
        { R2L_FOR_THREADS(NT) {
            static_assert(Cfg::Y1H % 3 == 0, "F3 micro-tiles of three rows");
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            for (int item = tid; item < (Cfg::Y1H / 3) * GG; item += NT) {
                const int blk = item / GG, g = item - blk * GG - 1;
                const int rr0 = 3 * blk;
                const int gx = tx0 + 4 * g;
                if (gx < 0 || gx >= W) continue;
                const int q = 2 + g;
                const int gy0 = ty0 - 2 + rr0;
                f2 acc[3][4];
#pragma unroll
                for (int o = 0; o < 3; ++o)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[o][j] = mk2(0.f, 0.f);
                const bool plain = gy0 >= 0 && gy0 + 2 < H;            // no pad row among the three
                if (plain) {
#pragma unroll
                    for (int ir = 0; ir < 5; ++ir) {                   // Y0 rows rr0 .. rr0+4 <-> image rows gy0-1 .. gy0+3
                        f2 in[6];
                        ld6<P>(Y0, (rr0 + ir) * P + 2 * q, in);
#pragma unroll
                        for (int o = 0; o < 3; ++o) {
                            const int aa = ir - o;
                            if (aa >= 0 && aa < 3) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
#pragma unroll
                                    for (int bb = 0; bb < 3; ++bb) acc[o][j] = fma2s(in[j + bb], ws[aa * 3 + bb], acc[o][j]);
                            }
                        }
                    }
                } else {
#pragma unroll
                    for (int o = 0; o < 3; ++o) {
                        const int rr = rr0 + o, gy = gy0 + o;
                        int sr = rr;                                   // Y0 rows sr .. sr+2 <-> image rows gy-1 .. gy+1
                        if ((gy < 0 && gy >= -2) || (gy >= H && gy <= H + 1)) sr = mirror(gy, H) - (ty0 - 2);
#pragma unroll
                        for (int aa = 0; aa < 3; ++aa) {
                            f2 in[6];
                            ld6<P>(Y0, (sr + aa) * P + 2 * q, in);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int bb = 0; bb < 3; ++bb) acc[o][j] = fma2s(in[j + bb], ws[aa * 3 + bb], acc[o][j]);
                        }
                    }
                }
#pragma unroll
                for (int o = 0; o < 3; ++o) {
                    const int rr = rr0 + o, gy = gy0 + o;
                    if (a.luma && rr >= 2 && rr < TH + 2 && g >= 0 && g < G && gy < H) {     // owned, in the image: keep Y1
                        float* dst = a.luma + ((((size_t)((a.B + 1) >> 1) + (b0 >> 1)) * H + gy) * W + gx) * 2;
                        st_luma4(dst, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
                    }
                    st4<P>(Y1, rr * P + 2 * q, acc[o][0], acc[o][1], acc[o][2], acc[o][3]);
                    if (gx == 0) { Y1[rr * P + phys<P>(4 * q - 1)] = acc[o][1]; Y1[rr * P + phys<P>(4 * q - 2)] = acc[o][2]; }
                    if (gx + 4 == W) { Y1[rr * P + phys<P>(4 * q + 4)] = acc[o][2]; Y1[rr * P + phys<P>(4 * q + 5)] = acc[o][1]; }
                }
            }
        } }
        R2L_SYNC();

        // ---- F4: Gaussian (2x4 runs) + YUV->RGB + clip + gamma [+ additive] [+ affine | sums] -> global ----------------
        { R2L_FOR_THREADS(NT) {
            float wg[25], m2[9];
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
#pragma unroll
            for (int t = 0; t < 9; ++t) m2[t] = T->M2[t];
            const float invg = T->invg;
            float sc[3] = {1.f, 1.f, 1.f}, sh[3] = {0.f, 0.f, 0.f};
            bool has_aff = false;
            if (TAIL && a.affine) {
                has_aff = true;
#pragma unroll
                for (int k = 0; k < 3; ++k) { sc[k] = a.affine[k]; sh[k] = a.affine[3 + k]; }
            }
            for (int item = tid; item < (TH / 2) * G; item += NT) {
                const int r0 = 2 * (item / G), g = item % G;
                const int q = 2 + g;
                const int gx = tx0 + 4 * g;
                f2 acc[2][4];
#pragma unroll
                for (int o = 0; o < 2; ++o)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[o][j] = mk2(0.f, 0.f);
#pragma unroll
                for (int ir = 0; ir < 6; ++ir) {
                    f2 in[8];
                    ld8<P>(Y1, (r0 + ir) * P + 2 * q, in);        // Y1 row index of image row (ty0 + r0 - 2 + ir)
#pragma unroll
                    for (int o = 0; o < 2; ++o) {
                        const int aa = ir - o;
                        if (aa >= 0 && aa < 5) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
#pragma unroll
                                for (int bb = 0; bb < 5; ++bb) acc[o][j] = fma2s(in[j + bb], wg[aa * 5 + bb], acc[o][j]);
                        }
                    }
                }
                if (gx >= W) continue;
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int gy = ty0 + r0 + o;
                    if (gy >= H) continue;
                    f2 u[4], v[4];
                    ld4<TW>(U, (r0 + o) * TW + 2 * g, u);
                    ld4<TW>(V, (r0 + o) * TW + 2 * g, v);
                    const size_t pix = (size_t)gy * W + gx;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        float oa[4], ob[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const f2 rr = fma2s(v[j], m2[k * 3 + 2], fma2s(u[j], m2[k * 3 + 1], mul2s(acc[o][j], m2[k * 3])));
                            const float ca = fminf(fmaxf(rr.x, kClipLo), kClipHi);
                            const float cbv = fminf(fmaxf(rr.y, kClipLo), kClipHi);
                            oa[j] = fast_exp2(invg * fast_log2(ca));
                            ob[j] = fast_exp2(invg * fast_log2(cbv));
                        }
                        if ((TAIL || STATS) && a.additive) {
                            const f4 ad = *reinterpret_cast<const f4*>(a.additive + (size_t)k * plane + pix);
                            oa[0] += ad.x; oa[1] += ad.y; oa[2] += ad.z; oa[3] += ad.w;
                            ob[0] += ad.x; ob[1] += ad.y; ob[2] += ad.z; ob[3] += ad.w;
                        }
                        if (STATS) {
                            ChanAcc& cs = R2L_ACC(cacc, tid);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                cs.s[k] += oa[j];
                                cs.s[3 + k] = fmaf_(oa[j], oa[j], cs.s[3 + k]);
                                if (!dup) {
                                    cs.s[k] += ob[j];
                                    cs.s[3 + k] = fmaf_(ob[j], ob[j], cs.s[3 + k]);
                                }
                            }
                        }
                        if (TAIL && has_aff) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) { oa[j] = fmaf_(oa[j], sc[k], sh[k]); ob[j] = fmaf_(ob[j], sc[k], sh[k]); }
                        }
                        f4 va; va.x = oa[0]; va.y = oa[1]; va.z = oa[2]; va.w = oa[3];
                        *reinterpret_cast<f4*>(a.out + ((size_t)b0 * 3 + k) * plane + pix) = va;
                        if (!dup) {
                            f4 vb; vb.x = ob[0]; vb.y = ob[1]; vb.z = ob[2]; vb.w = ob[3];
                            *reinterpret_cast<f4*>(a.out + ((size_t)b1 * 3 + k) * plane + pix) = vb;
                        }
                    }
                }
            }
        } }
        R2L_SYNC();   // planes are rewritten by the next tile
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
        float* red = smem + Cfg::kTableFloats;
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
        if (a.bn_sync) {
            // ---- fused BatchNorm tail (FwdArgs::bn_sync): grid-wide barrier, statistics, normalisation of own tiles --------
            // Every CTA of the persistent grid is resident (or becomes so as the kernel before this one drains), so
            // waiting for all of them cannot deadlock.  arrive: the channel sums above are visible device-wide first.
            __shared__ float s_aff[6];
            __threadfence();
            __syncthreads();
            if (tid == 0) {
                take_ticket(a.bn_sync, a.bn_gen);
                const volatile unsigned long long* w = reinterpret_cast<const volatile unsigned long long*>(a.bn_sync);
                for (;;) {
                    const unsigned long long v = *w;
                    if ((unsigned)(v >> 32) == a.bn_gen && (unsigned)v >= (unsigned)n_cta) break;
                    __nanosleep(64);
                }
                __threadfence();
            }
            __syncthreads();
            // batch mean / biased variance per channel: one warp per channel, lanes stride over the per-CTA sums, fp64,
            // fixed order -- the same arithmetic in every CTA (and as bn_finish_kernel), so all CTAs normalise alike
            if (warp < 3) {
                const int c = warp;
                double s1 = 0.0, s2 = 0.0;
                for (int i = lane; i < n_cta; i += 32) {
                    s1 += (double)__ldcg(a.chan_partials + (size_t)i * kChanPitch + c);
                    s2 += (double)__ldcg(a.chan_partials + (size_t)i * kChanPitch + 3 + c);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                if (lane == 0) {
                    const double mean = s1 / a.bn_count;
                    double var = s2 / a.bn_count - mean * mean;
                    if (var < 0.0) var = 0.0;
                    const double inv = 1.0 / sqrt(var + (double)a.bn_eps);
                    s_aff[c] = (float)inv;
                    s_aff[3 + c] = (float)(-mean * inv);
                    if (cta == 0) {
                        a.bn_saved_affine[c] = (float)inv;
                        a.bn_saved_affine[3 + c] = (float)(-mean * inv);
                        const double m = (double)a.bn_momentum;
                        if (c == 0 && a.bn_num_batches) *a.bn_num_batches += 1;      // nn.BatchNorm2d bookkeeping
                        if (a.bn_running_mean) a.bn_running_mean[c] = (float)((1.0 - m) * (double)a.bn_running_mean[c] + m * mean);
                        if (a.bn_running_var) {
                            const double unbiased = a.bn_count > 1.0 ? var * a.bn_count / (a.bn_count - 1.0) : var;
                            a.bn_running_var[c] = (float)((1.0 - m) * (double)a.bn_running_var[c] + m * unbiased);
                        }
                    }
                }
            }
            __syncthreads();
            // normalise, in place, exactly the tiles this CTA produced (128-bit accesses; L2-only loads: the lines were
            // written by this CTA a few microseconds ago)
            for (int tile = cta; tile < grid.n; tile += n_cta) {
                int b0, b1, ty0, tx0;
                decode_pair_tile(grid, tile, TH, TW, a.B, b0, b1, ty0, tx0);
                const int rows = imin(TH, H - ty0), runs = imin(G, (W - tx0) >> 2);
                const int n_img = b1 == b0 ? 1 : 2;
                static_assert((TH * G) % NT == 0 && (G & (G - 1)) == 0, "whole passes over the tile, shift / mask indexing");
                // (image, channel) planes outermost, a tile plane = TH x G runs of four sites: no division in the loop.
                // All six planes of the tile are requested before the first is written back: one L2 round trip per
                // tile instead of six (the pass is latency-bound: 2 x 16 bytes per thread and plane).
                constexpr int NU = TH * G / NT;
                float4 v[6][NU];
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) {
                    if (pl >= 3 * n_img) continue;
                    const int im = pl >= 3 ? 1 : 0, k = pl - 3 * im;
                    const float* base = a.out + ((size_t)(b0 + im) * 3 + k) * plane + (size_t)ty0 * W + tx0;
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        const int i = tid + u * NT, rr = i / G, g = i & (G - 1);
                        if (rr < rows && g < runs) v[pl][u] = __ldcg(reinterpret_cast<const float4*>(base + (size_t)rr * W) + g);
                    }
                }
#pragma unroll
                for (int pl = 0; pl < 6; ++pl) {
                    if (pl >= 3 * n_img) continue;
                    const int im = pl >= 3 ? 1 : 0, k = pl - 3 * im;
                    const float sc = s_aff[k], sh = s_aff[3 + k];
                    float* base = a.out + ((size_t)(b0 + im) * 3 + k) * plane + (size_t)ty0 * W + tx0;
#pragma unroll
                    for (int u = 0; u < NU; ++u) {
                        const int i = tid + u * NT, rr = i / G, g = i & (G - 1);
                        if (rr < rows && g < runs) {
                            float4 t = v[pl][u];
                            t.x = fmaf(t.x, sc, sh); t.y = fmaf(t.y, sc, sh); t.z = fmaf(t.z, sc, sh); t.w = fmaf(t.w, sc, sh);
                            reinterpret_cast<float4*>(base + (size_t)rr * W)[g] = t;
                        }
                    }
                }
            }
            // depart: the last CTA to leave clears both ticket words (a captured launch replays with the same tag)
            __syncthreads();
            if (tid == 0 && take_ticket(a.bn_sync + 2, a.bn_gen) == (unsigned)n_cta - 1u) {
                clear_ticket(a.bn_sync);
                clear_ticket(a.bn_sync + 2);
            }
        }
#endif
    }
}

}  // namespace r2l
