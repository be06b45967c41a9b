// isp_bwd2.cuh -- second-generation fused backward: two images per lane (float2), 1x4 register runs, statistics
// accumulated next to the adjoint gathers that already hold the right windows in registers.
//
// Reference: the autograd graph of pipeline_torch.py:183-217.  Same adjoint algebra as bwd_tile
// (isp_core.cuh); what changes is the schedule:
//   * planes are float2 (image A, image B) in the chunk de-interleaved layout of isp_fwd2.cuh -> FFMA2 everywhere;
//   * every correlation statistic is taken in its "flipped" form, sum over the *stencil centre* q:
//         dWg[t] = sum_q Y1pad(q) * gY2(q - t),   dWs[t] = sum_q Y0(q) * gY1(q - t),
//         Q'[par(q)][k][t] = sum_q rawpad(q) * g_yuv[k](q - t)
//     so the window the adjoint gather loads (gY2 / gY1 / g_yuv around q) is reused for the statistic and only the
//     centre value is read in addition.  Pad sites q outside the image (reflect padding) belong to the nearest
//     border tile.  Q' is re-indexed to the p-based layout of kStatQ when the CTA writes its partial sums, so the
//     finish kernel (finish_grad) is shared with the first generation.
//   * the per-thread partial sums (95 scalars) live in registers across all tiles of the persistent CTA.
#pragma once
#include "isp_fwd2.cuh"

namespace r2l {

template <int P> R2L_HD f2& site(f2* pl, int row, int col) { return pl[row * P + phys<P>(col)]; }
R2L_HD f2 add2(f2 a, f2 b) { return mk2(a.x + b.x, a.y + b.y); }
R2L_HD f2 fma2v(f2 a, f2 b, f2 c) {
#ifdef R2L_HOST_EMU
    return mk2(std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y));
#else
    return __ffma2_rn(a, b, c);
#endif
}

struct Bwd2Acc {
    float sg;            // sum G o log2(cl)
    float wg[25];        // flipped Gaussian-weight statistic
    float ws[9];         // flipped sharpen-weight statistic
    float q[2][3][9];    // Q'[col phase of q][k][t] for this thread's row phase
    float p[2][3];       // P[col phase][k] = sum g_yuv[k](q) over owned sites
};
constexpr int kBwd2AccFloats = 1 + 25 + 9 + 54 + 6;

template <int TH_, int TW_, int NT_, bool GRAW_> struct Bwd2Cfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    static constexpr bool GRAW = GRAW_;
    static constexpr int G = TW / 4;
    static constexpr int PW = TW + 24;        // wide planes (raw, Y0, Y1): column index = gx - x0 + 12, run q = g + 3
    static constexpr int PN = TW + 16;        // narrow planes (F, gY1):    column index = gx - x0 + 8,  run q = g + 2
    static constexpr int RH = TH + 16, Y0H = TH + 14, Y1H = TH + 12, FH = TH + 8, G1H = TH + 4;
    static constexpr int kTableFloats = (sizeof(Tables2) + 15) / 16 * 4;
    static constexpr int kXR = RH * PW, kY0 = Y0H * PW, kF = FH * PN, kG1 = G1H * PN;
    static constexpr size_t kPlaneBytes = (size_t)kTableFloats * 4 + (size_t)(kXR + kY0 + 3 * kF + kG1) * 8;
    static constexpr size_t kStageOffset = (kPlaneBytes + 127) / 128 * 128;
    static constexpr size_t kStageBytes = (size_t)2 * RH * PW * 4;
    static constexpr size_t kSmemBytes = kPlaneBytes;
    static constexpr size_t kSmemBytesTma = kStageOffset + kStageBytes + 16;
    static constexpr int kRowsPerPass = 14;                       // 14 rows x (G+2) groups = 252 threads keep one row phase
    static_assert((G + 2) * kRowsPerPass <= NT && (kRowsPerPass % 2) == 0, "phase-stable mapping");
    static_assert(Y1H * PW <= kXR, "Y1 aliases the raw window");
    static_assert((size_t)NT * (kBwd2AccFloats + 1) * 4 <= (size_t)(kXR + kY0 + 3 * kF + kG1) * 8, "reduction scratch");
};

// reflect-pad pre-images and checked gathers for sites whose adjoint folds at the image border (slow path)
template <int PN>
R2L_HD f2 gather5_checked(const float* wg, f2* GY2, int row0, int col0, int H, int W, int qy, int qx, int ty0, int tx0) {
    // sum_ab Wg[ab] * gY2(q + (2-a, 2-b)), in-image terms only.  F-plane row = gy - (ty0-4), col = gx - (tx0-8)
    f2 v = mk2(0.f, 0.f);
    for (int a = 0; a < 5; ++a) {
        const int py = qy + 2 - a;
        if (py < 0 || py >= H) continue;
        for (int b = 0; b < 5; ++b) {
            const int px = qx + 2 - b;
            if (px < 0 || px >= W) continue;
            v = fma2s(site<PN>(GY2, py - (ty0 - 4), px - (tx0 - 8)), wg[a * 5 + b], v);
        }
    }
    (void)row0; (void)col0;
    return v;
}

template <class Cfg, typename RawT, bool TMA = false>
R2L_HD void bwd2_cta(int cta, int n_cta, const BwdArgs& a, const TileGrid& grid, float* smem, const void* tmap = nullptr) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT, PW = Cfg::PW, PN = Cfg::PN, G = Cfg::G;
    constexpr int GG = G + 2;                                     // groups -1 .. G
    Tables2* T2 = reinterpret_cast<Tables2*>(smem);
    Tables* T = &T2->base;
    f2* XR = reinterpret_cast<f2*>(smem + Cfg::kTableFloats);     // raw window, later Y1
    f2* Y0 = XR + Cfg::kXR;
    f2* PU = Y0 + Cfg::kY0;                                       // U, then gU
    f2* PV = PU + Cfg::kF;                                        // V, then gV
    f2* PG = PV + Cfg::kF;                                        // gY2, then gY0
    f2* GY1 = PG + Cfg::kF;
    f2* Y1 = XR;
#ifdef R2L_HOST_EMU
    std::vector<Bwd2Acc> accs(NT);
    std::memset(accs.data(), 0, sizeof(Bwd2Acc) * NT);
#else
    Bwd2Acc accs;
    {
        float* z = reinterpret_cast<float*>(&accs);
#pragma unroll
        for (int i = 0; i < kBwd2AccFloats; ++i) z[i] = 0.f;
    }
#endif
    // planes start finite (pad columns are read by edge runs but never written)
    { R2L_FOR_THREADS(NT) {
        f2* all = XR;
        for (int i = tid; i < Cfg::kXR + Cfg::kY0 + 3 * Cfg::kF + Cfg::kG1; i += NT) all[i] = mk2(0.f, 0.f);
    } }
    R2L_BUILD_TABLES(NT, a.P, T)
    { R2L_FOR_THREADS(NT) { build_tables2_extra(tid, NT, T2); } }
    R2L_SYNC();

    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const bool vec_ok = (W % 4) == 0;
#ifndef R2L_HOST_EMU
    RawT* stage = reinterpret_cast<RawT*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset + Cfg::kStageBytes);
    uint32_t tma_phase = 0;
    constexpr uint32_t kTmaBytes = 2u * Cfg::RH * PW * sizeof(RawT);
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(mbar, 1);
            if (cta < grid.n) {
                int pb0, pb1, py0, px0;
                decode_pair_tile(grid, cta, TH, TW, a.B, pb0, pb1, py0, px0);
                tma_load_3d(stage, tmap, px0 - 12, py0 - 8, pb0, mbar, kTmaBytes);
            }
        }
        __syncthreads();
    }
#endif
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b0, b1, ty0, tx0;
        decode_pair_tile(grid, tile, TH, TW, a.B, b0, b1, ty0, tx0);
        const bool dup = b1 == b0;
        const RawT* imgA = static_cast<const RawT*>(a.raw) + (size_t)b0 * plane;
        const RawT* imgB = static_cast<const RawT*>(a.raw) + (size_t)b1 * plane;
        const bool interior = ty0 >= 8 && tx0 >= 12 && ty0 + TH + 8 <= H && tx0 + TW + 12 <= W;
        // statistic domains: owned rectangle, extended over the reflect padding on image-border sides
        const int oy1 = imin(ty0 + TH, H), ox1 = imin(tx0 + TW, W);
        const int e_top = ty0 == 0, e_bot = ty0 + TH >= H, e_lft = tx0 == 0, e_rgt = tx0 + TW >= W;

        // ---- B1: raw window (rows -8..TH+7, columns -12..TW+11) ------------------------------------------------
#ifndef R2L_HOST_EMU
        if (TMA) {
            mbar_wait(mbar, tma_phase);
            tma_phase ^= 1u;
            {
                const int tid = threadIdx.x;
                constexpr int Q = PW / 4;
                const RawT* sa = stage;
                const RawT* sb = stage + Cfg::RH * PW;
                for (int i = tid; i < Cfg::RH * Q; i += NT) {
                    const int ly = i / Q, lq = i - ly * Q;
                    float va[4], vb[4];
                    if (sizeof(RawT) == 4) {
                        const f4 xa = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(sa) + ly * PW + 4 * lq);
                        const f4 xb = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(sb) + ly * PW + 4 * lq);
                        va[0] = xa.x; va[1] = xa.y; va[2] = xa.z; va[3] = xa.w;
                        vb[0] = xb.x; vb[1] = xb.y; vb[2] = xb.z; vb[3] = xb.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            va[j] = __fdiv_rn((float)sa[ly * PW + 4 * lq + j], a.denom);
                            vb[j] = __fdiv_rn((float)sb[ly * PW + 4 * lq + j], a.denom);
                        }
                    }
                    st4<PW>(XR, ly * PW + 2 * lq, mk2(va[0], vb[0]), mk2(va[1], vb[1]), mk2(va[2], vb[2]), mk2(va[3], vb[3]));
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const int next = tile + n_cta;
                if (next < grid.n) {
                    int nb0, nb1, ny0, nx0;
                    decode_pair_tile(grid, next, TH, TW, a.B, nb0, nb1, ny0, nx0);
                    tma_load_3d(stage, tmap, nx0 - 12, ny0 - 8, nb0, mbar, kTmaBytes);
                }
            }
            if (!interior) {
                const int tid = threadIdx.x;
                for (int i = tid; i < 2 * PW; i += NT) {
                    const int gy = i < PW ? -1 : H, lx = i < PW ? i : i - PW;
                    const int ly = gy - (ty0 - 8), sy = mirror(gy, H) - (ty0 - 8);
                    if (ly >= 0 && ly < Cfg::RH && sy >= 0 && sy < Cfg::RH) XR[ly * PW + lx] = XR[sy * PW + lx];
                }
                __syncthreads();
                for (int i = tid; i < 2 * Cfg::RH; i += NT) {
                    const int gx = i < Cfg::RH ? -1 : W, ly = i < Cfg::RH ? i : i - Cfg::RH;
                    const int lx = gx - (tx0 - 12), sx = mirror(gx, W) - (tx0 - 12);
                    if (lx >= 0 && lx < PW && sx >= 0 && sx < PW) site<PW>(XR, ly, lx) = site<PW>(XR, ly, sx);
                }
                __syncthreads();
            }
        } else
#endif
        { R2L_FOR_THREADS(NT) {
            constexpr int Q = PW / 4;
            for (int i = tid; i < Cfg::RH * Q; i += NT) {
                const int ly = i / Q, lq = i - ly * Q;
                const int gy = ty0 - 8 + ly, gx = tx0 - 12 + 4 * lq;
                const int eb = ly * PW + 2 * lq;
                if (vec_ok && sizeof(RawT) == 4 && gy >= 0 && gy < H && gx >= 0 && gx + 3 < W) {
                    const f4 va = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgA) + (size_t)gy * W + gx);
                    const f4 vb = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgB) + (size_t)gy * W + gx);
                    st4<PW>(XR, eb, mk2(va.x, vb.x), mk2(va.y, vb.y), mk2(va.z, vb.z), mk2(va.w, vb.w));
                } else {
                    const int sy = mirror_clamped(gy, H);
                    f2 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sx = mirror_clamped(gx + j, W);
                        v[j] = mk2(RawLoad<RawT>::get(imgA + (size_t)sy * W + sx, a.denom),
                                   RawLoad<RawT>::get(imgB + (size_t)sy * W + sx, a.denom));
                    }
                    st4<PW>(XR, eb, v[0], v[1], v[2], v[3]);
                }
            }
        } }
        R2L_SYNC();

        // ---- B2: Y0 on rows -7..TH+6; U,V on the F region (rows -4..TH+3, groups -1..G) -------------------------
        { R2L_FOR_THREADS(NT) {
            if (tid < Cfg::kRowsPerPass * GG) {
                const int row0 = tid / GG, g = tid - row0 * GG - 1;
                const int rp = row0 & 1;                              // rows -4 + row0 + 14k keep this phase
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
                for (int r = -4 + row0; r < TH + 4; r += Cfg::kRowsPerPass) {
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
                    st4<PW>(Y0, (r + 7) * PW + 2 * (g + 3), acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
                    st4<PN>(PU, (r + 4) * PN + 2 * (g + 2), acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
                    st4<PN>(PV, (r + 4) * PN + 2 * (g + 2), acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
                }
            }
            // luma-only ring: rows -7..-5 and TH+4..TH+6 x groups -2..G+1, plus groups -2 and G+1 of the F rows
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
                st4<PW>(Y0, (ry + 7) * PW + 2 * (g + 3), acc[0], acc[1], acc[2], acc[3]);
            }
        } }
        R2L_SYNC();
        if (!interior) {
            { R2L_FOR_THREADS(NT) {        // sharpen zero-pads: Y0 = 0 on the 1-wide ring outside the image
                for (int i = tid; i < 2 * PW + 2 * Cfg::Y0H; i += NT) {
                    int gy, gx;
                    if (i < PW) { gy = -1; gx = tx0 - 12 + i; }
                    else if (i < 2 * PW) { gy = H; gx = tx0 - 12 + (i - PW); }
                    else if (i < 2 * PW + Cfg::Y0H) { gy = ty0 - 7 + (i - 2 * PW); gx = -1; }
                    else { gy = ty0 - 7 + (i - 2 * PW - Cfg::Y0H); gx = W; }
                    const int ly = gy - (ty0 - 7), lx = gx - (tx0 - 12);
                    if (ly >= 0 && ly < Cfg::Y0H && lx >= 0 && lx < PW) site<PW>(Y0, ly, lx) = mk2(0.f, 0.f);
                }
            } }
            R2L_SYNC();
        }

        // ---- B3: Y1 = sharpen(Y0) on rows -6..TH+5, groups -2..G+1 (overwrites the raw window) -------------------
        { R2L_FOR_THREADS(NT) {
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            for (int item = tid; item < Cfg::Y1H * (G + 4); item += NT) {
                const int rr = item / (G + 4), g = item - rr * (G + 4) - 2;
                f2 acc[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<PW>(Y0, (rr + aa) * PW + 2 * (g + 3), in);     // Y1 row rr <-> image row ty0-6+rr; Y0 row of (that-1) = rr
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) acc[j] = fma2s(in[j + bb], ws[aa * 3 + bb], acc[j]);
                }
                st4<PW>(Y1, rr * PW + 2 * (g + 3), acc[0], acc[1], acc[2], acc[3]);
            }
        } }
        R2L_SYNC();
        if (!interior) {
            { R2L_FOR_THREADS(NT) {        // Gaussian reflect-pads the sharpened plane: rows, then columns
                for (int i = tid; i < 4 * PW; i += NT) {
                    const int q = i / PW, lx = i - q * PW;
                    const int gy = q == 0 ? -2 : (q == 1 ? -1 : (q == 2 ? H : H + 1));
                    const int ly = gy - (ty0 - 6), sy = mirror(gy, H) - (ty0 - 6);
                    if (ly >= 0 && ly < Cfg::Y1H && sy >= 0 && sy < Cfg::Y1H) Y1[ly * PW + lx] = Y1[sy * PW + lx];
                }
            } }
            R2L_SYNC();
            { R2L_FOR_THREADS(NT) {
                for (int i = tid; i < 4 * Cfg::Y1H; i += NT) {
                    const int q = i / Cfg::Y1H, ly = i - q * Cfg::Y1H;
                    const int gx = q == 0 ? -2 : (q == 1 ? -1 : (q == 2 ? W : W + 1));
                    const int lx = gx - (tx0 - 12), sx = mirror(gx, W) - (tx0 - 12);
                    if (lx >= 0 && lx < PW && sx >= 0 && sx < PW) site<PW>(Y1, ly, lx) = site<PW>(Y1, ly, sx);
                }
            } }
            R2L_SYNC();
        }

        // ---- B4: forward tail recomputed on the F region; grad_out pulled back to (gY2, gU, gV); gamma statistic ----
        { R2L_FOR_THREADS(NT) {
            float wg[25], m2[9];
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
#pragma unroll
            for (int t = 0; t < 9; ++t) m2[t] = T->M2[t];
            const float invg = T->invg;
            Bwd2Acc& acc = R2L_ACC(accs, tid);
            for (int item = tid; item < Cfg::FH * GG; item += NT) {
                const int rr = item / GG, g = item - rr * GG - 1;
                const int r = rr - 4;
                const int gy = ty0 + r, gx = tx0 + 4 * g;
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
                const bool row_in = gy >= 0 && gy < H;
                const bool row_owned = r >= 0 && r < TH && g >= 0 && g < G;
                const size_t pix = (size_t)(row_in ? gy : 0) * W + gx;
                const bool full = row_in && vec_ok && gx >= 0 && gx + 3 < W;
                f2 gy2[4], gu[4], gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { gy2[j] = mk2(0.f, 0.f); gu[j] = mk2(0.f, 0.f); gv[j] = mk2(0.f, 0.f); }
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    float ga[4] = {0.f, 0.f, 0.f, 0.f}, gb[4] = {0.f, 0.f, 0.f, 0.f};
                    const float* pa = a.gout + ((size_t)b0 * 3 + k) * plane + pix;
                    const float* pb = a.gout + ((size_t)b1 * 3 + k) * plane + pix;
                    if (full) {
                        const f4 xa = *reinterpret_cast<const f4*>(pa);
                        ga[0] = xa.x; ga[1] = xa.y; ga[2] = xa.z; ga[3] = xa.w;
                        if (!dup) { const f4 xb = *reinterpret_cast<const f4*>(pb); gb[0] = xb.x; gb[1] = xb.y; gb[2] = xb.z; gb[3] = xb.w; }
                    } else if (row_in) {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (gx + j >= 0 && gx + j < W) { ga[j] = pa[j]; if (!dup) gb[j] = pb[j]; }
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const bool in_img = row_in && gx + j >= 0 && gx + j < W;
                        const f2 rgb = fma2s(v[j], m2[k * 3 + 2], fma2s(u[j], m2[k * 3 + 1], mul2s(y2[j], m2[k * 3])));
                        const float ca = fminf(fmaxf(rgb.x, kClipLo), kClipHi), cbv = fminf(fmaxf(rgb.y, kClipLo), kClipHi);
                        const float la = fast_log2(ca), lb = fast_log2(cbv);
                        const float oa = fast_exp2(invg * la), ob = fast_exp2(invg * lb);
                        float Ga = ga[j], Gb = gb[j];
                        if (a.gtail) {
                            float ya = oa, yb = ob;
                            if (a.additive && in_img) { const float ad = a.additive[(size_t)k * plane + pix + j]; ya += ad; yb += ad; }
                            ya = fmaf_(ya, a.gtail[9 + k], a.gtail[12 + k]);
                            yb = fmaf_(yb, a.gtail[9 + k], a.gtail[12 + k]);
                            Ga = a.gtail[k] * (Ga - a.gtail[3 + k] - a.gtail[6 + k] * ya);
                            Gb = a.gtail[k] * (Gb - a.gtail[3 + k] - a.gtail[6 + k] * yb);
                            if (!in_img) { Ga = 0.f; Gb = 0.f; }
                            if (dup) Gb = 0.f;
                        }
                        const float goa = Ga * oa, gob = Gb * ob;
                        if (row_owned && in_img) acc.sg = fmaf_(goa, la, fmaf_(gob, lb, acc.sg));
                        const bool pa_ = rgb.x >= kClipLo && rgb.x <= kClipHi, pb_ = rgb.y >= kClipLo && rgb.y <= kClipHi;
                        const f2 gr = mk2(pa_ ? goa * invg * fast_rcp(ca) : 0.f, pb_ ? gob * invg * fast_rcp(cbv) : 0.f);
                        gy2[j] = fma2s(gr, m2[k * 3 + 0], gy2[j]);
                        gu[j] = fma2s(gr, m2[k * 3 + 1], gu[j]);
                        gv[j] = fma2s(gr, m2[k * 3 + 2], gv[j]);
                    }
                }
                st4<PN>(PG, (r + 4) * PN + 2 * (g + 2), gy2[0], gy2[1], gy2[2], gy2[3]);
                st4<PN>(PU, (r + 4) * PN + 2 * (g + 2), gu[0], gu[1], gu[2], gu[3]);
                st4<PN>(PV, (r + 4) * PN + 2 * (g + 2), gv[0], gv[1], gv[2], gv[3]);
            }
        } }
        R2L_SYNC();

        // ---- B5: gY1 = fold_reflect2(corr^T(gY2, Wg)) on rows -2..TH+1, groups -1..G; flipped Wg statistic --------
        { R2L_FOR_THREADS(NT) {
            float wg[25];
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
            Bwd2Acc& acc = R2L_ACC(accs, tid);
            // Wg statistic domain (padded sites q): owned rectangle, +2 on image-border sides
            const int sy0 = e_top ? -2 : ty0, sy1 = e_bot ? H + 2 : oy1, sx0 = e_lft ? -2 : tx0, sx1 = e_rgt ? W + 2 : ox1;
            for (int item = tid; item < Cfg::G1H * GG; item += NT) {
                const int rr = item / GG, g = item - rr * GG - 1;
                const int r = rr - 2;
                const int qy = ty0 + r, qx = tx0 + 4 * g;
                f2 win[5][8];
#pragma unroll
                for (int d = 0; d < 5; ++d) ld8<PN>(PG, (r + 2 + d) * PN + 2 * (g + 2), win[d]);   // rows qy-2 .. qy+2
                f2 out[4];
                const bool regular = qy >= 3 && qy <= H - 4 && qx >= 3 && qx + 3 <= W - 4;
                if (regular) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        f2 s = mk2(0.f, 0.f);
#pragma unroll
                        for (int aa = 0; aa < 5; ++aa)
#pragma unroll
                            for (int bb = 0; bb < 5; ++bb) s = fma2s(win[4 - aa][j + 4 - bb], wg[aa * 5 + bb], s);
                        out[j] = s;
                    }
                } else {
                    for (int j = 0; j < 4; ++j) {
                        f2 s = mk2(0.f, 0.f);
                        const int x = qx + j;
                        if (qy >= 0 && qy < H && x >= 0 && x < W) {
                            int ys[3], xs[3];
                            const int ny = preimages2(qy, H, ys), nx = preimages2(x, W, xs);
                            for (int iy = 0; iy < ny; ++iy)
                                for (int ix = 0; ix < nx; ++ix)
                                    s = add2(s, gather5_checked<PN>(wg, PG, 0, 0, H, W, ys[iy], xs[ix], ty0, tx0));
                        }
                        out[j] = s;
                    }
                }
                st4<PN>(GY1, rr * PN + 2 * (g + 2), out[0], out[1], out[2], out[3]);
                // statistic: dWg[ab] += Y1pad(q) * gY2(q - (a-2, b-2)) for the sites q of this run inside the domain
                if (qy >= sy0 && qy < sy1 && qx + 3 >= sx0 && qx < sx1) {
                    f2 c[4];
                    ld4<PW>(Y1, (r + 6) * PW + 2 * (g + 3), c);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (qx + j < sx0 || qx + j >= sx1) c[j] = mk2(0.f, 0.f);
#pragma unroll
                    for (int aa = 0; aa < 5; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) {
                            f2 t = mk2(0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < 4; ++j) t = fma2v(c[j], win[4 - aa][j + 4 - bb], t);
                            acc.wg[aa * 5 + bb] += t.x + t.y;
                        }
                }
            }
        } }
        R2L_SYNC();

        // ---- B6: gY0 = corr^T(gY1, Ws) (zero pad) on rows -1..TH, groups -1..G; flipped Ws statistic -----------------
        { R2L_FOR_THREADS(NT) {
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            Bwd2Acc& acc = R2L_ACC(accs, tid);
            for (int item = tid; item < (TH + 2) * GG; item += NT) {
                const int rr = item / GG, g = item - rr * GG - 1;
                const int r = rr - 1;
                const int qy = ty0 + r, qx = tx0 + 4 * g;
                f2 win[3][6];
#pragma unroll
                for (int d = 0; d < 3; ++d) ld6<PN>(GY1, (r + 1 + d) * PN + 2 * (g + 2), win[d]);  // rows qy-1 .. qy+1
                f2 out[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    f2 s = mk2(0.f, 0.f);
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) s = fma2s(win[2 - aa][j + 2 - bb], ws[aa * 3 + bb], s);
                    const bool in_img = qy >= 0 && qy < H && qx + j >= 0 && qx + j < W;
                    out[j] = in_img ? s : mk2(0.f, 0.f);
                }
                st4<PN>(PG, (r + 4) * PN + 2 * (g + 2), out[0], out[1], out[2], out[3]);
                if (r >= 0 && r < TH && g >= 0 && g < G && qy < H) {
                    f2 c[4];
                    ld4<PW>(Y0, (r + 7) * PW + 2 * (g + 3), c);
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (qx + j >= W) c[j] = mk2(0.f, 0.f);
#pragma unroll
                    for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) {
                            f2 t = mk2(0.f, 0.f);
#pragma unroll
                            for (int j = 0; j < 4; ++j) t = fma2v(c[j], win[2 - aa][j + 2 - bb], t);
                            acc.ws[aa * 3 + bb] += t.x + t.y;
                        }
                }
            }
        } }
        R2L_SYNC();

        // ---- B7: Q' / P statistics and g_raw from the (gY0, gU, gV) windows; raw centre from global memory ----------
        { R2L_FOR_THREADS(NT) {
            if (tid < Cfg::kRowsPerPass * GG) {
                const int row0 = tid / GG, g = tid - row0 * GG - 1;
                const int rp = (row0 + 1) & 1;                         // rows -1 + row0 + 14k keep this phase
                Bwd2Acc& acc = R2L_ACC(accs, tid);
                float awq[2][3][9];                                    // adjoint demosaic->YUV taps of this row phase
                if (Cfg::GRAW) {
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                        for (int k = 0; k < 3; ++k)
#pragma unroll
                            for (int t = 0; t < 9; ++t) awq[cp][k][t] = T->AWq[2 * rp + cp][k][t];
                }
                // Q' domain: owned rectangle, +1 on image-border sides (reflect-1 padding of the mosaic)
                const int sy0 = e_top ? -1 : ty0, sy1 = e_bot ? H + 1 : oy1, sx0 = e_lft ? -1 : tx0, sx1 = e_rgt ? W + 1 : ox1;
                const int qx = tx0 + 4 * g;
                for (int r = -1 + row0; r < TH + 1; r += Cfg::kRowsPerPass) {
                    const int qy = ty0 + r;
                    const bool in_dom = qy >= sy0 && qy < sy1 && qx + 3 >= sx0 && qx < sx1;
                    const bool owned_row = r >= 0 && r < TH && g >= 0 && g < G && qy < H;
                    if (!in_dom && !owned_row) continue;
                    // raw centres of the 4 sites (reflected for pad sites), zero outside the domain
                    f2 c[4];
                    {
                        const int sy = mirror_clamped(qy, H);
                        if (vec_ok && sizeof(RawT) == 4 && qy >= 0 && qy < H && qx >= 0 && qx + 3 < W) {
                            const f4 xa = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgA) + (size_t)qy * W + qx);
                            const f4 xb = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgB) + (size_t)qy * W + qx);
                            c[0] = mk2(xa.x, xb.x); c[1] = mk2(xa.y, xb.y); c[2] = mk2(xa.z, xb.z); c[3] = mk2(xa.w, xb.w);
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const int sx = mirror_clamped(qx + j, W);
                                c[j] = mk2(RawLoad<RawT>::get(imgA + (size_t)sy * W + sx, a.denom),
                                           RawLoad<RawT>::get(imgB + (size_t)sy * W + sx, a.denom));
                            }
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (!in_dom || qx + j < sx0 || qx + j >= sx1) c[j] = mk2(0.f, 0.f);
                    }
                    f2 graw[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
                    const bool regular = qy >= 2 && qy <= H - 3 && qx >= 2 && qx + 3 <= W - 3;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        f2* pl = k == 0 ? PG : (k == 1 ? PU : PV);
                        f2 win[3][6];
#pragma unroll
                        for (int d = 0; d < 3; ++d) ld6<PN>(pl, (r + 3 + d) * PN + 2 * (g + 2), win[d]);   // rows qy-1 .. qy+1
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
#pragma unroll
                            for (int aa = 0; aa < 3; ++aa)
#pragma unroll
                                for (int bb = 0; bb < 3; ++bb) {
                                    const f2 t = win[2 - aa][j + 2 - bb];                          // g_yuv[k](q - (a-1, b-1))
                                    acc.q[j & 1][k][aa * 3 + bb] = fmaf_(c[j].x, t.x, fmaf_(c[j].y, t.y, acc.q[j & 1][k][aa * 3 + bb]));
                                    if (Cfg::GRAW) graw[j] = fma2s(t, awq[j & 1][k][aa * 3 + bb], graw[j]);
                                }
                            if (owned_row && qx + j < W) acc.p[j & 1][k] += win[1][j + 1].x + win[1][j + 1].y;
                        }
                    }
                    if (Cfg::GRAW && owned_row) {
                        if (!regular) {
                            // adjoint of the reflect-1 padding folds onto rows/columns 1 and n-2: checked gathers
                            for (int j = 0; j < 4; ++j) {
                                const int x = qx + j;
                                if (x >= W) continue;
                                const int par = par_of(qy, x);
                                int ys[3], xs[3];
                                const int ny = preimages1(qy, H, ys), nx = preimages1(x, W, xs);
                                f2 s = mk2(0.f, 0.f);
                                for (int iy = 0; iy < ny; ++iy)
                                    for (int ix = 0; ix < nx; ++ix)
                                        for (int aa = 0; aa < 3; ++aa) {
                                            const int py = ys[iy] + 1 - aa;
                                            if (py < 0 || py >= H) continue;
                                            for (int bb = 0; bb < 3; ++bb) {
                                                const int px = xs[ix] + 1 - bb;
                                                if (px < 0 || px >= W) continue;
                                                const int t = aa * 3 + bb, fr = py - (ty0 - 4), fc = px - (tx0 - 8);
                                                s = fma2s(site<PN>(PG, fr, fc), T->AWq[par][0][t], s);
                                                s = fma2s(site<PN>(PU, fr, fc), T->AWq[par][1][t], s);
                                                s = fma2s(site<PN>(PV, fr, fc), T->AWq[par][2][t], s);
                                            }
                                        }
                                graw[j] = s;
                            }
                        }
                        float* pa = a.graw + (size_t)b0 * plane + (size_t)qy * W + qx;
                        float* pb = a.graw + (size_t)b1 * plane + (size_t)qy * W + qx;
                        if (vec_ok && qx + 3 < W) {
                            f4 va; va.x = graw[0].x; va.y = graw[1].x; va.z = graw[2].x; va.w = graw[3].x;
                            *reinterpret_cast<f4*>(pa) = va;
                            if (!dup) { f4 vb; vb.x = graw[0].y; vb.y = graw[1].y; vb.z = graw[2].y; vb.w = graw[3].y; *reinterpret_cast<f4*>(pb) = vb; }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (qx + j < W) { pa[j] = graw[j].x; if (!dup) pb[j] = graw[j].y; }
                        }
                    }
                }
            }
        } }
        R2L_SYNC();   // planes are rewritten by the next tile
    }

    // ---- CTA reduction of the per-thread statistics into the kStat* layout (deterministic, fixed order) -------------
    float* part = a.partials + (size_t)cta * kStatPitch;
    constexpr int RP = kBwd2AccFloats + 1;                           // odd pitch: conflict-light column reads
    float* red = reinterpret_cast<float*>(XR);
    { R2L_FOR_THREADS(NT) {
        const float* src = reinterpret_cast<const float*>(&R2L_ACC(accs, tid));
        for (int i = 0; i < kBwd2AccFloats; ++i) red[tid * RP + i] = src[i];
    } }
    R2L_SYNC();
    { R2L_FOR_THREADS(NT) {
        for (int s = tid; s < kNumStats; s += NT) {
            float sum = 0.f;
            if (s < kStatQ) {                                        // gamma, Wg, Ws: same slot in every thread
                for (int t = 0; t < NT; ++t) sum += red[t * RP + s];
            } else {
                int k, parp, slot_kind, tt = 0;
                if (s < kStatP) { const int rI = s - kStatQ; k = rI / 36; parp = (rI - 36 * k) / 9; tt = rI - 36 * k - 9 * parp; slot_kind = 0; }
                else { const int rI = s - kStatP; k = rI / 4; parp = rI - 4 * k; slot_kind = 1; }
                // Q[k][par(p)][t] = Q'[par(q) = par_tap(par(p), t)][k][t];  P is already p-indexed (p = q)
                const int parq = slot_kind == 0 ? par_tap(parp, tt) : parp;
                const int rpq = parq >> 1, cpq = parq & 1;
                const int off = slot_kind == 0 ? 35 + (cpq * 3 + k) * 9 + tt : 35 + 54 + cpq * 3 + k;
                for (int t = 0; t < Cfg::kRowsPerPass * GG; ++t) {
                    const int trp = ((t / GG) + 1) & 1;
                    if (trp == rpq) sum += red[t * RP + off];
                }
            }
            part[s] = sum;
        }
    } }
}

}  // namespace r2l
