// isp_bwd4.cuh -- fourth-generation fused backward: nothing is recomputed.
//
// Reference: the autograd graph of pipeline_torch.py:183-217.  Same adjoint algebra, float2
// (image A, image B) planes, FFMA2 arithmetic, padded-domain border rules and flipped statistics as the third
// generation (isp_bwd3.cuh), but the forward quantities come from what the forward kept instead of being rebuilt
// per tile from a +-8 raw window:
//   * the clip mask and the gamma derivative are read off the forward OUTPUT (as in the third generation's OUT mode);
//   * the flipped statistics only need Y1 / Y0 / raw at the stencil CENTRE q (dWg[t] = sum_q Y1(q) gY2(q - t), ...),
//     so the forward saves its Y0 and Y1 planes (8 B/px, pair-interleaved so a centre run is two 128-bit loads that
//     land as FFMA2 operands) and the backward reads centres straight from global memory.
// That removes the raw window, its TMA staging, the de-interleave pass, the Y0 and Y1 recompute phases (21 % of the
// third generation's time, profiles/r01_v4_summary.md) and 100 KB of shared memory: two CTAs fit per SM, so the
// global-load phase of one CTA overlaps the stencil phases of the other (the chip has the HBM bandwidth to spare,
// DESIGN.md section 1; the third generation's single CTA per SM issued all of its loads in one burst per tile).
// Per tile: B4 (grad_out, out) -> (gY2, gU, gV); B5 adjoint Gaussian + dWg; B6 adjoint sharpen + dWs;
// B7 Q'/P statistics + g_raw.  Four barriers.  The last CTA to finish turns the per-CTA partial sums into the 132
// gradients (no separate finish launch).
#pragma once
#include "isp_bwd3.cuh"

namespace r2l {

// Row pitch of the shared-memory planes = TW + 16 sites in use + 4 pad sites.  With 80 sites (640 B) every row starts on the
// same bank, so the halo-ring items -- a warp's lanes on one column of 16 rows -- met 16-way bank conflicts (ncu: 8.7 wavefronts
// per 128-bit load against 3); 84 sites shift consecutive rows by 8 banks: backward 84.5 -> 83.0 us (in-call A/B; 82 sites, one
// bank group per row, measured the same).  The pad sites are never read.
#ifndef R2L_BWD_PN_PAD
#define R2L_BWD_PN_PAD 20
#endif
template <int TH_, int TW_, int NT_, bool GRAW_, bool TAIL_> struct Bwd4Cfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    static constexpr bool GRAW = GRAW_, TAIL = TAIL_;
    static constexpr int G = TW / 4;
    static constexpr int PN = TW + R2L_BWD_PN_PAD;   // plane pitch (sites): column index = gx - x0 + 8, run q = g + 2
    static constexpr int FH = TH + 8, G1H = TH + 4;
    static constexpr int kTableFloats = (sizeof(Tables2) + 15) / 16 * 4;
    static constexpr int kF = FH * PN, kG1 = G1H * PN;
    static constexpr int kSites = 3 * kF + kG1;
    static constexpr size_t kSmemBytes = (size_t)kTableFloats * 4 + (size_t)kSites * 8;
    static constexpr int HALF = NT / 2;       // threads per CFA row phase
    static_assert(NT % 64 == 0 && TH % 2 == 0 && TW % 8 == 0, "warp-parity mapping");
};

// the shapes / pointers the fourth generation serves (everything else: third generation or generic kernel)
inline bool bwd4_shape_ok(int H, int W) { return (W % 4) == 0 && H >= 8 && W >= 8; }

// 4 sites x (image A, image B) of a saved luma plane: 32 contiguous bytes, read once
R2L_HD void ld_luma4(const float* p, f2 c[4]) {
#ifdef R2L_HOST_EMU
    const f4 v0 = ld_stream4(p), v1 = ld_stream4(p + 4);
    c[0] = mk2(v0.x, v0.y); c[1] = mk2(v0.z, v0.w); c[2] = mk2(v1.x, v1.y); c[3] = mk2(v1.z, v1.w);
#else
    // ONE 256-bit load (sm_100: LDG.E.256, L2 only like ld_stream4): as two 128-bit loads every warp instruction touched
    // only half of each 32-byte sector it requested (the lane stride is 32 bytes).  p is 32-byte aligned.
    float v[8];
    asm volatile("ld.global.cg.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7]) : "l"(p));
    c[0] = mk2(v[0], v[1]); c[1] = mk2(v[2], v[3]); c[2] = mk2(v[4], v[5]); c[3] = mk2(v[6], v[7]);
#endif
}

#ifndef R2L_HOST_EMU
__device__ __forceinline__ void prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#endif

#ifndef R2L_HOST_EMU
// ---- fused finish, one level (fourth generation; the fifth uses the two-level finish further down, which shares the chain
// rule): the last CTA to publish its partial sums turns them into the 132 gradients ------------------
// Same arithmetic and summation order as isp_backward_finish_kernel (isp_host.cu): per-CTA partials are summed in
// double in CTA order (bit-reproducible), then the chain rule of finish_grad_sc.  Loads are issued 8 CTA rows deep
// per warp (40 independent 4-byte loads per lane) so the whole read is a handful of L2 round trips.
// scratch: doubles, [NW][kStatPitch] + kNumStats + 108 + 9 + 4 + 9
template <int NT>
__device__ __forceinline__ void finish_chain_rule(const Tables* T, const double* S, float* grads, double* Qr, double* Tkc,
                                                  double* Gbl, double* Sc9);

template <int NT>
__device__ __forceinline__ void fused_finish(const Tables* T, const float* partials, int n_cta, float* grads, double* scratch) {
    constexpr int NW = NT / 32, SPL = kStatPitch / 32;              // 5 statistics per lane
    static_assert(kStatPitch % 32 == 0, "one CTA row = SPL coalesced warp loads");
    double* Sw = scratch;                                            // [NW][kStatPitch]
    double* S = Sw + NW * kStatPitch;                                // [kNumStats] (+ pad to 160)
    double* Qr = S + kStatPitch;                                     // [108]
    double* Tkc = Qr + 108;                                          // [9]
    double* Gbl = Tkc + 9;                                           // [4]
    double* Sc9 = Gbl + 4;                                           // [9]
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    {
        // warp w sums CTA rows c0..c1-1 (a contiguous block, so the total below is in CTA order up to the block split)
        const int per = (n_cta + NW - 1) / NW, c0 = warp * per, c1 = min(n_cta, c0 + per);
        double sum[SPL];
#pragma unroll
        for (int i = 0; i < SPL; ++i) sum[i] = 0.0;
        for (int c = c0; c < c1; c += 8) {
            float v[8][SPL];
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < SPL; ++i)
                    v[u][i] = (c + u < c1) ? __ldcg(partials + (size_t)(c + u) * kStatPitch + i * 32 + lane) : 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u)
#pragma unroll
                for (int i = 0; i < SPL; ++i) sum[i] += (double)v[u][i];
        }
#pragma unroll
        for (int i = 0; i < SPL; ++i) Sw[warp * kStatPitch + i * 32 + lane] = sum[i];
    }
    __syncthreads();
    for (int s = tid; s < kNumStats; s += NT) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += Sw[w * kStatPitch + s];
        S[s] = t;
    }
    __syncthreads();
    finish_chain_rule<NT>(T, S, grads, Qr, Tkc, Gbl, Sc9);
}

// stages 3 - 5 of the finish: the 155 sums S (shared memory, complete and visible to the CTA) -> the 132 gradients
template <int NT>
__device__ __forceinline__ void finish_chain_rule(const Tables* T, const double* S, float* grads, double* Qr, double* Tkc,
                                                  double* Gbl, double* Sc9) {
    constexpr int NW = NT / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 108; i += NT) { const int k = i / 36, r = i - 36 * k; Qr[i] = finish_qr(S, T, k, r / 9, r % 9); }
    __syncthreads();
    for (int job = warp; job < 13; job += NW) {
        if (job < 9) {                                               // Tkc[k][c]: 36 terms, lanes take (par, t)
            const int k = job / 3, c = job - 3 * k;
            double v = 0.0;
            for (int i = lane; i < 36; i += 32) {
                const int par = i / 9, t = i - 9 * par;
                v += (double)T->wd[(c * 3 + ch_of(par_tap(par, t))) * 9 + t] * Qr[k * 36 + i];
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) Tkc[job] = v;
        } else {                                                     // black_level[e]: 27 terms, lanes take (t, k)
            const int e = job - 9;
            double v = 0.0;
            if (lane < 27) {
                const int t = lane / 3, k = lane - 3 * t, par = par_tap(e, t);
                v = (double)T->AW[par][k][t] * S[stat_p_index(k, par)];
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) Gbl[e] = -v;
        }
    }
    __syncthreads();
    if (tid < 9) {                                                   // Sc[m][c] = sum_k M1[k][m] * Tkc[k][c]
        const int m = tid / 3, c = tid - 3 * m;
        double v = 0.0;
        for (int k = 0; k < 3; ++k) v += (double)T->m1[k * 3 + m] * Tkc[k * 3 + c];
        Sc9[tid] = v;
    }
    __syncthreads();
    for (int e = tid; e < R2L_NUM_PARAM_GRADS; e += NT) grads[e] = e < 4 ? (float)Gbl[e] : finish_grad_sc(e, S, T, Sc9);
}

// ---- two-level finish (fifth generation) --------------------------------------------------------------------------
// The one-level finish above is a serial tail: the last CTA reads all n_cta rows while the rest of the GPU is idle
// (7.3 us of an 82 us launch, profiles/r02_experiments.md).  Here the CTAs form kFinishGroups static groups of consecutive
// CTA indices; the last CTA of a GROUP to arrive sums that group's rows (a few rows per warp: one L2 round trip) into one
// fp64 row, and only the last group finisher adds the <= 16 group rows and runs the chain rule.  No CTA ever waits for
// another one.  The order of every addition is fixed by CTA index, so the result is reproducible run to run.
constexpr int kFinishGroups = 16;
// the group rows live behind the per-CTA rows of the workspace: the last 32 of its kMaxCtas rows = 16 x 160 doubles
constexpr int kFinishMaxCtas = kMaxCtas - kFinishGroups * 2;
static_assert((size_t)kFinishGroups * kStatPitch * sizeof(double) <= (size_t)(kMaxCtas - kFinishMaxCtas) * kStatPitch * sizeof(float),
              "the group rows fit the workspace rows the capped grid never uses");
static_assert(128 + kFinishGroups * 8 <= 256, "the group tickets fit the workspace's 256-byte ticket block (isp_host.cu)");
__device__ __forceinline__ double* finish_group_rows(float* partials) {
    return reinterpret_cast<double*>(partials + (size_t)kFinishMaxCtas * kStatPitch);
}
__device__ __forceinline__ int finish_group_size(int n_cta) { return (n_cta + kFinishGroups - 1) / kFinishGroups; }

// rows c0 .. c1-1 of the per-CTA partials -> out[kStatPitch] (global, fp64).  scratch: [NW][kStatPitch] doubles.
template <int NT>
__device__ __forceinline__ void finish_group_sum(const float* partials, int c0, int c1, double* out, double* scratch) {
    constexpr int NW = NT / 32, SPL = kStatPitch / 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    double sum[SPL];
#pragma unroll
    for (int i = 0; i < SPL; ++i) sum[i] = 0.0;
    for (int c = c0 + warp; c < c1; c += 3 * NW) {                   // warp w: rows c0 + w, + NW, + 2 NW, ... (three in flight)
        float v[3][SPL];
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int i = 0; i < SPL; ++i)
                v[u][i] = (c + u * NW < c1) ? __ldcg(partials + (size_t)(c + u * NW) * kStatPitch + i * 32 + lane) : 0.f;
#pragma unroll
        for (int u = 0; u < 3; ++u)
#pragma unroll
            for (int i = 0; i < SPL; ++i) sum[i] += (double)v[u][i];
    }
#pragma unroll
    for (int i = 0; i < SPL; ++i) scratch[warp * kStatPitch + i * 32 + lane] = sum[i];
    __syncthreads();
    for (int s = tid; s < kStatPitch; s += NT) {
        double t = 0.0;
#pragma unroll
        for (int w = 0; w < NW; ++w) t += scratch[w * kStatPitch + s];
        out[s] = t;
    }
}

// the ng group rows -> S (shared) -> the 132 gradients.  scratch: kStatPitch + 108 + 9 + 4 + 9 doubles.
template <int NT>
__device__ __forceinline__ void finish_from_groups(const Tables* T, const double* rows, int ng, float* grads, double* scratch) {
    double* S = scratch;
    double* Qr = S + kStatPitch;
    double* Tkc = Qr + 108;
    double* Gbl = Tkc + 9;
    double* Sc9 = Gbl + 4;
    const int tid = threadIdx.x;
    for (int s = tid; s < kStatPitch; s += NT) {
        double v[kFinishGroups];
#pragma unroll
        for (int g = 0; g < kFinishGroups; ++g) v[g] = g < ng ? __ldcg(rows + (size_t)g * kStatPitch + s) : 0.0;
        double t = 0.0;
#pragma unroll
        for (int g = 0; g < kFinishGroups; ++g) t += v[g];
        S[s] = t;
    }
    __syncthreads();
    finish_chain_rule<NT>(T, S, grads, Qr, Tkc, Gbl, Sc9);
}
#endif

template <class Cfg, typename RawT>
R2L_HD void bwd4_cta(int cta, int n_cta, const BwdArgs& a, const TileGrid& grid, float* smem) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT, PN = Cfg::PN, G = Cfg::G, HALF = Cfg::HALF;
    constexpr int GG = G + 2;                                     // runs -1 .. G
    Tables2* T2 = reinterpret_cast<Tables2*>(smem);
    Tables* T = &T2->base;
    f2* PU = reinterpret_cast<f2*>(smem + Cfg::kTableFloats);     // gU
    f2* PV = PU + Cfg::kF;                                        // gV
    f2* PG = PV + Cfg::kF;                                        // gY2, then gY0
    f2* GY1 = PG + Cfg::kF;                                       // gY1
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
    const size_t luma_plane = (size_t)((a.B + 1) >> 1) * plane * 2;      // floats per saved plane
    // planes start finite: never-written pad columns are read by don't-care items
    { R2L_FOR_THREADS(NT) {
        for (int i = tid; i < Cfg::kSites; i += NT) PU[i] = mk2(0.f, 0.f);
    } }
    R2L_BUILD_TABLES(NT, a.P, T)
    { R2L_FOR_THREADS(NT) { build_tables2_extra(tid, NT, T2); } }
    R2L_SYNC();
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b0, b1, ty0, tx0;
        decode_pair_tile(grid, tile, TH, TW, a.B, b0, b1, ty0, tx0);
        const bool dup = b1 == b0;
        const RawT* imgA = static_cast<const RawT*>(a.raw) + (size_t)b0 * plane;
        const RawT* imgB = static_cast<const RawT*>(a.raw) + (size_t)b1 * plane;
        const float* y0pair = a.luma + (size_t)(b0 >> 1) * plane * 2;   // Y0 of this image pair, [H][W][2]
        const float* y1pair = y0pair + luma_plane;

        // ---- B4: grad_out pulled back through gamma / clip / YUV->RGB to (gY2, gU, gV) on rows -4..TH+3, runs -1..G;
        // gamma statistic.  o = y (or (y - shift)/scale - additive behind a tail); with lo = log2(o):
        // e = cl^(1/g - 1) = 2^((1 - g) lo), log2(cl) = g lo, and the clamp passed iff o lies strictly between its two
        // clipped values (exact compare without a tail, where o is bit-identical to the forward's; a 1e-4 / 1e-6
        // relative margin behind a tail, where o is recovered by an affine inverse).
        { R2L_FOR_THREADS(NT) {
#ifndef R2L_HOST_EMU
            // the centres B5 / B6 / B7 read from global memory (Y1, Y0, raw of the owned rows): pull their lines into L2
            {
                constexpr int LPR = TW * 8 / 128;                      // 128-byte lines per owned row of a luma plane
                for (int i = tid; i < TH * LPR * 2; i += NT) {
                    const int l = i % LPR, rr = (i / LPR) % TH, pl = i / (LPR * TH);
                    const int gy = ty0 + rr, gx = tx0 + l * 16;
                    if (gy < H && gx < W) prefetch_l2((pl ? y1pair : y0pair) + ((size_t)gy * W + gx) * 2);
                }
                constexpr int RPL = 128 / (int)sizeof(RawT);           // raw elements per line
                constexpr int LPRR = (TW + RPL - 1) / RPL;
                for (int i = tid; i < TH * LPRR * 2; i += NT) {
                    const int l = i % LPRR, rr = (i / LPRR) % TH, im = i / (LPRR * TH);
                    const int gy = ty0 + rr, gx = tx0 + l * RPL;
                    if (gy < H && gx < W) prefetch_l2((im ? imgB : imgA) + (size_t)gy * W + gx);
                }
            }
#endif
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

        // ---- B5: gY1 = fold_reflect2(corr^T(gY2, Wg)) on rows -2..TH+1, runs -1..G (zero outside the image); dWg ----
        // Border rules as in isp_bwd3.cuh B5: the reflect-2 fold and the pad sites' share of dWg are a few extra products
        // inside the items of rows 1,2 / H-2,H-3 and of the first / last run of the image.  Y1 centres come from the
        // plane the forward saved.
        { R2L_FOR_THREADS(NT) {
#ifndef R2L_HOST_EMU
            // next tile's grad_out / forward-output windows (rows -4..TH+3 of 3 channels x 2 images x 2 tensors) -> L2
            {
                const int next = tile + n_cta;
                if (next < grid.n) {
                    int nb0, nb1, ny0, nx0;
                    decode_pair_tile(grid, next, TH, TW, a.B, nb0, nb1, ny0, nx0);
                    constexpr int LPR = TW * 4 / 128;
                    const int nl = Cfg::FH * 12 * LPR;
                    for (int i = tid; i < nl; i += NT) {
                        const int l = i % LPR, pr = i / LPR, pk = pr % 12, rr = pr / 12;
                        const int gy = ny0 - 4 + rr, gx = nx0 + l * 32;
                        const int pk6 = pk % 6, img = pk6 < 3 ? nb0 : nb1, k = pk6 < 3 ? pk6 : pk6 - 3;
                        if (gy >= 0 && gy < H && gx < W)
                            prefetch_l2((pk < 6 ? a.gout : a.out) + ((size_t)img * 3 + k) * plane + (size_t)gy * W + gx);
                    }
                }
            }
#endif
            float wg[25];
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
            Bwd3Acc& acc = R2L_ACC(accs, tid);
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
                    for (int j = 0; j < 4; ++j) { out[u][j] = mk2(0.f, 0.f); c[u][j] = mk2(0.f, 0.f); }
                }
                if (any) {
                    if (stat) {
#pragma unroll
                        for (int u = 0; u < NR; ++u)
                            if (inside[u]) ld_luma4(y1pair + ((size_t)qy * W + tx0 + 4 * (g + u * (G / 2))) * 2, c[u]);
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
                        if (stat) {
#pragma unroll
                            for (int bb = 0; bb < 5; ++bb) {
                                f2 t = mul2vv(c[0][0], row[0][4 - bb]);
#pragma unroll
                                for (int u = 0; u < NR; ++u)
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
                                        if (u | j) t = fma2vv(c[u][j], row[u][j + 4 - bb], t);
                                acc.wg[aa * 5 + bb] += t.x + t.y;
                            }
                        }
                    }
                    // ---- folded pad contributions: a separate, rarely taken block (rows 1,2 / H-2,H-3, first / last run of
                    // the image); the window rows it needs are loaded again.  A pad site's Y1 value is the folded-onto
                    // site's own, so the statistic share needs the centre even on the halo ring: it is only taken by
                    // statistic-carrying (owned) items, whose centres are loaded ----
                    bool anycol = false;
#pragma unroll
                    for (int u = 0; u < NR; ++u) anycol |= lft[u] | rgt[u];
                    if (rt != 0 || anycol) {
#pragma unroll
                        for (int d = 0; d < 5; ++d) {
                            const bool rowx = (d == 0 && rt == 2) || (d == 1 && rt == 1) || (d == 2 && (rt == 1 || rt == 3)) ||
                                              (d == 3 && rt == 3) || (d == 4 && rt == 4);
                            if (rowx || anycol) {
                            f2 row[NR][8];
#pragma unroll
                            for (int u = 0; u < NR; ++u) ld8<PN>(PG, (r + 2 + d) * PN + 2 * (g + u * (G / 2) + 2), row[u]);
                            auto col_extra = [&](const int A) {
#pragma unroll
                                for (int u = 0; u < NR; ++u) {
                                    if (lft[u]) {    // pad columns -1 (-> site 1, taps b = 0,1) and -2 (-> site 2, tap b = 0)
                                        out[u][1] = fma2s(row[u][3], wg[A * 5 + 0], fma2s(row[u][2], wg[A * 5 + 1], out[u][1]));
                                        out[u][2] = fma2s(row[u][2], wg[A * 5 + 0], out[u][2]);
                                        if (stat) {
                                            const f2 t0 = fma2vv(c[u][1], row[u][3], mul2vv(c[u][2], row[u][2]));
                                            const f2 t1 = mul2vv(c[u][1], row[u][2]);
                                            acc.wg[A * 5 + 0] += t0.x + t0.y;
                                            acc.wg[A * 5 + 1] += t1.x + t1.y;
                                        }
                                    }
                                    if (rgt[u]) {    // pad columns W (-> site 2, taps b = 3,4) and W+1 (-> site 1, tap b = 4)
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
                            auto row_extra = [&](const int A2) {
#pragma unroll
                                for (int u = 0; u < NR; ++u)
#pragma unroll
                                    for (int j = 0; j < 4; ++j)
#pragma unroll
                                        for (int bb = 0; bb < 5; ++bb) out[u][j] = fma2s(row[u][j + 4 - bb], wg[A2 * 5 + bb], out[u][j]);
                                if (stat) {
#pragma unroll
                                    for (int bb = 0; bb < 5; ++bb) {
                                        f2 t = mul2vv(c[0][0], row[0][4 - bb]);
#pragma unroll
                                        for (int u = 0; u < NR; ++u)
#pragma unroll
                                            for (int j = 0; j < 4; ++j)
                                                if (u | j) t = fma2vv(c[u][j], row[u][j + 4 - bb], t);
                                        acc.wg[A2 * 5 + bb] += t.x + t.y;
                                    }
                                }
                                col_extra(A2);                               // corner pads
                            };
                            if (anycol) col_extra(4 - d);
                            if (d == 0 && rt == 2) row_extra(0);
                            if (d == 1 && rt == 1) row_extra(1);
                            if (d == 2 && rt == 1) row_extra(0);
                            if (d == 2 && rt == 3) row_extra(4);
                            if (d == 3 && rt == 3) row_extra(3);
                            if (d == 4 && rt == 4) row_extra(4);
                            }
                        }
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
            for (int item = tid; item < TH * (G / 2); item += NT)       // owned rectangle: paired items
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
                const int qy = ty0 + r;
                f2 c[NR][4], out[NR][4];
#pragma unroll
                for (int u = 0; u < NR; ++u) {
                    const int qx = tx0 + 4 * (g + u * (G / 2));
#pragma unroll
                    for (int j = 0; j < 4; ++j) { out[u][j] = mk2(0.f, 0.f); c[u][j] = mk2(0.f, 0.f); }
                    if (owned && qy < H && qx < W) ld_luma4(y0pair + ((size_t)qy * W + qx) * 2, c[u]);
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
                    const int qx = tx0 + 4 * (g + u * (G / 2));
                    if (qy < 0 || qy >= H || qx < 0 || qx >= W) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                    }
                    st4<PN>(PG, (r + 4) * PN + 2 * (g + u * (G / 2) + 2), out[u][0], out[u][1], out[u][2], out[u][3]);
                }
            };
            for (int item = tid; item < TH * (G / 2); item += NT)       // owned rectangle: paired items
                b6(std::integral_constant<int, 2>(), item / (G / 2), item % (G / 2), true);
            for (int item = TH * G + tid; item < (TH + 2) * GG; item += NT) {      // halo ring, single runs
                int r, g;
                region_item<TH, G, 1>(item, r, g);
                b6(std::integral_constant<int, 1>(), r, g, false);
            }
        } }
        R2L_SYNC();

        // ---- B7: Q' / P statistics and g_raw from the (gY0, gU, gV) windows; border rules as in isp_bwd3.cuh B7 ----
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
                f2 c[4];                                                // raw centres of the 4 sites
                if (sizeof(RawT) == 4) {
                    const f4 xa = ld_stream4(reinterpret_cast<const float*>(imgA) + (size_t)qy * W + qx);
                    const f4 xb = ld_stream4(reinterpret_cast<const float*>(imgB) + (size_t)qy * W + qx);
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
                    }
                }
                if (f_top | f_bot | f_lft | f_rgt) {
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        f2* pl = k == 0 ? PG : (k == 1 ? PU : PV);
#pragma unroll
                        for (int d = 0; d < 3; ++d) {
                            const bool padrow = (f_top && d == 0) || (f_bot && d == 2);
                            if (!(padrow | f_lft | f_rgt)) continue;
                            f2 row[6];
                            ld6<PN>(pl, (r + 3 + d) * PN + 2 * (g + 2), row);
                            const int aa = 2 - d;
                            if (padrow) {
#pragma unroll
                                for (int j = 0; j < 4; ++j)
#pragma unroll
                                    for (int bb = 0; bb < 3; ++bb) {
                                        const f2 t = row[j + 2 - bb];
                                        acc.q[j & 1][k][d * 3 + bb] = fmaf_(c[j].x, t.x, fmaf_(c[j].y, t.y, acc.q[j & 1][k][d * 3 + bb]));
                                        if (Cfg::GRAW) graw[j] = fma2s(t, awq[j & 1][k][d * 3 + bb], graw[j]);
                                    }
                            }
                            if (f_lft) {
                                const f2 t = row[1];
                                acc.q[1][k][aa * 3 + 0] = fmaf_(c[1].x, t.x, fmaf_(c[1].y, t.y, acc.q[1][k][aa * 3 + 0]));
                                if (Cfg::GRAW) graw[1] = fma2s(t, awq[1][k][aa * 3 + 0], graw[1]);
                                if (padrow) {                                                    // corner pad
                                    acc.q[1][k][d * 3 + 0] = fmaf_(c[1].x, t.x, fmaf_(c[1].y, t.y, acc.q[1][k][d * 3 + 0]));
                                    if (Cfg::GRAW) graw[1] = fma2s(t, awq[1][k][d * 3 + 0], graw[1]);
                                }
                            }
                            if (f_rgt) {
                                const f2 t = row[4];
                                acc.q[0][k][aa * 3 + 2] = fmaf_(c[2].x, t.x, fmaf_(c[2].y, t.y, acc.q[0][k][aa * 3 + 2]));
                                if (Cfg::GRAW) graw[2] = fma2s(t, awq[0][k][aa * 3 + 2], graw[2]);
                                if (padrow) {                                                    // corner pad
                                    acc.q[0][k][d * 3 + 2] = fmaf_(c[2].x, t.x, fmaf_(c[2].y, t.y, acc.q[0][k][d * 3 + 2]));
                                    if (Cfg::GRAW) graw[2] = fma2s(t, awq[0][k][d * 3 + 2], graw[2]);
                                }
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

    // ---- CTA reduction of the per-thread statistics into the kStat* layout (deterministic, fixed order), as in
    // isp_bwd3.cuh ----------------------------------------------------------------------------------------------
    float* part = a.partials + (size_t)cta * kStatPitch;
    constexpr int NW = NT / 32;
    constexpr int RP = kBwd3AccFloats + 1;
    float* red = reinterpret_cast<float*>(PU);                       // [NW][RP]
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
#ifndef R2L_HOST_EMU
    if (a.ticket) {
        __shared__ unsigned last_flag;
        __threadfence();                                             // this CTA's partial sums are visible device-wide ...
        __syncthreads();
        if (threadIdx.x == 0) last_flag = atomicAdd(a.ticket, 1u) == (unsigned)n_cta - 1u;   // ... before its ticket is
        __syncthreads();
        if (last_flag) {
            __threadfence();
            static_assert((size_t)(NT / 32 + 1) * kStatPitch * 8 + 130 * 8 <= (size_t)Cfg::kSites * 8, "finish scratch fits the planes");
            fused_finish<NT>(T, a.partials, n_cta, a.grads, reinterpret_cast<double*>(PU));
        }
    }
#endif
}

}  // namespace r2l
