// isp_bwd6.cuh -- sixth-generation fused backward: a warp-specialised, TMA-fed dataflow pipeline.
//
// Reference: the autograd graph of pipeline_torch.py:183-217.  Same adjoint algebra, float2 (image A,
// image B) planes, FFMA2 arithmetic, flipped statistics and border rules on the data as the fifth generation
// (isp_bwd5.cuh); what changes is the schedule.  The fifth generation walked 32 x 64 tiles through four CTA-wide
// phases: 21 % of the launch the SMs were empty (3.46 tile rounds, prologue / finish tail), the pointwise phase B4
// recomputed a 1.41x halo and stalled on its global loads, and every warp waited at four barriers per tile
// (profiles/r01_v5_summary.md, r02_v5_ncu: issue slots 42 % busy, FMA pipe 37 %).  Here:
//   * ONE CTA of 16 warps per SM; every warp has ONE job for the whole launch:
//       warps [0, N4)   B4  grad_out, out -> (gY2, gU, gV)  (pointwise: gamma / clip / YUV->RGB adjoint, gamma statistic)
//       next N5         B5  gY1 = fold_reflect2(corr^T(gY2, Wg)), dWg statistic (25 packed sums live in registers)
//       next N6         B6  gY0 = corr^T(gY1, Ws), dWs statistic
//       next N7         B7  Q' / P statistics (everything upstream of YUV) and g_raw; sums parked in tensor memory per k
//       last warp       TMA producer: grad_out / out boxes of the next B4 passes into an 8-slot staging ring
//     so the MUFU / ALU-heavy pointwise work, the FMA-heavy stencils and the copy engine overlap inside every SM
//     partition, and no warp carries another stage's accumulators;
//   * the image batch is cut into column STRIPS of 64 sites x all rows of an image pair, and every CTA streams a
//     contiguous, equally long range of strip rows top to bottom (balanced to two rows: no tile-count quantisation, no
//     vertical halo recompute except 8 rows where a CTA's range starts);
//   * the stages are coupled by ring buffers of plane rows in shared memory (gY2 32 rows, gY1 32, gY0 32, gU / gV 64)
//     and by monotone progress words (st.release / ld.acquire, shared memory): a consumer pass waits until the producer
//     passes it reads are complete, a producer pass waits until every consumer has moved past the rows it overwrites.
//     There is no CTA barrier between the prologue and the statistics reduction;
//   * a pass is one warp x 32 work items: two rows x 16 runs of 4 sites (main passes), or 16 rows x the two halo runs
//     beside the strip (halo passes, one per 16 rows): every lane busy in every pass;
//   * B4's main passes read grad_out / out from the TMA staging ring (cp.async.bulk.tensor.4d over (W, H, 3, B), box
//     64 x 2 x 3 x 2, out-of-bounds zero fill realises the image border and an odd batch's missing partner).
// Work items, per-item arithmetic and the statistics layout handed to the finish (kStat*) are those of isp_bwd5.cuh.
#pragma once
#include "isp_bwd5.cuh"

#ifdef R2L_HOST_EMU
#include <atomic>
#include <thread>
#include <chrono>
#endif

namespace r2l {

constexpr int kB6SW = 64, kB6G = 16;              // strip width (sites), runs of 4 sites per strip row
constexpr int kB6P = 84;                          // plane pitch (sites): column index = gx - x0 + 8, run q = g + 2;
                                                  // 84 * 8 B = 672 B rows shift by 8 banks, so the 16-row halo passes
                                                  // do not pile onto one bank group
constexpr int kB6R4 = 32, kB6R5 = 32, kB6R6 = 32, kB6RUV = 64;      // ring rows: gY2, gY1, gY0, gU / gV (powers of two)
constexpr int kB6NS = 8;                          // TMA staging slots (one B4 main pass each)
constexpr int kB6SlotFloats = 2 * 2 * 3 * 2 * 64; // [tensor: grad_out, out][image][plane][row][64]
constexpr int kB6ParkCols = 128;                  // TMEM columns per B7 warp: 3 x (18 + 2) packed sums = 120
constexpr int kB6K = 40;                          // parked floats per gradient plane k: 18 Q' pairs + 2 P pairs

template <bool GRAW_, bool TAIL_, int N4_ = 4, int N5_ = 4, int N6_ = 2, int N7_ = 5> struct Bwd6Cfg {
    static constexpr bool GRAW = GRAW_, TAIL = TAIL_;
    static constexpr int N4 = N4_, N5 = N5_, N6 = N6_, N7 = N7_;
    static constexpr int NW = N4 + N5 + N6 + N7 + 1, NT = NW * 32;
    static constexpr int W5 = N4, W6 = N4 + N5, W7 = N4 + N5 + N6, WT = N4 + N5 + N6 + N7;     // first warp of a role
    static constexpr int kTableFloats = (sizeof(Tables2) + 15) / 16 * 4;
    static constexpr int kSyncInts = 64;          // progress words
    static constexpr int kRingSites = (kB6R4 + kB6R5 + kB6R6 + 2 * kB6RUV) * kB6P;
    static constexpr int kRedFloats = NW * 128;   // per-warp statistics hand-over
    static constexpr size_t kStageOffset = (((size_t)kTableFloats * 4 + kSyncInts * 4 + (size_t)kRingSites * 8 + (size_t)kRedFloats * 4) + 127) / 128 * 128;
    static constexpr size_t kBarOffset = kStageOffset + (size_t)kB6NS * kB6SlotFloats * 4;
    static constexpr size_t kSmemBytes = kBarOffset + 2 * kB6NS * 8 + 16;
    static constexpr int kTmemCols = ((N7 + 3) / 4) * kB6ParkCols <= 128 ? 128 : (((N7 + 3) / 4) * kB6ParkCols <= 256 ? 256 : 512);
    static_assert(N4 <= 8 && N5 <= 8 && N6 <= 8 && N7 <= 8 && NT <= 1024, "progress words: eight per role");
    static_assert(((N7 + 3) / 4) * kB6ParkCols <= 512, "B7 sums must fit the CTA's TMEM columns");
};

inline bool bwd6_shape_ok(int H, int W) { return (W % 4) == 0 && H >= 8 && W >= 8; }

// one range of strip rows handed to a CTA: image pair `pair`, columns [x0, x0 + 64), rows [y0, y1) (y0 even, y1 - y0 even;
// y1 may be H + 1 for an odd H: rows >= H are outside the image).  cum* = pass counts / ring rows of the CTA's earlier
// segments, so pass indices and virtual ring rows keep growing across segments.
struct Seg6 {
    int pair, x0, y0, y1;
    int cum4, cum5, cum6, cum7, cumM4;            // passes of B4 / B5 / B6 / B7 so far; B4 MAIN passes so far (staging slots)
    int base4, base5, base6;                      // virtual row of this segment's first B4 / B5 / B6 row
    R2L_HD int n() const { return y1 - y0; }
    R2L_HD int m4() const { return n() + 8; }     // B4 rows y0-4 .. y1+3
    R2L_HD int m5() const { return n() + 4; }     // B5 rows y0-2 .. y1+1
    R2L_HD int m6() const { return n() + 2; }     // B6 rows y0-1 .. y1
    R2L_HD static int passes(int rows) { return rows / 2 + (rows + 15) / 16; }
    // pass order of a stage inside a segment: [halo block 0, main groups 0..7, halo block 1, main groups 8..15, ...]
    R2L_HD static int idx_main(int g) { return g + (g >> 3) + 1; }
    R2L_HD static int idx_halo(int h) { return 9 * h; }
};

struct SegIter6 {
    int upg, ncols, u, u1;                        // units (2 rows) per strip, strips per pair, this CTA's unit range
    Seg6 s;
    bool first;
    R2L_HD SegIter6(int cta, int n_cta, int B, int H, int W) {
        upg = (H + 1) >> 1;
        ncols = (W + kB6SW - 1) / kB6SW;
        const long long total = (long long)((B + 1) >> 1) * ncols * upg;
        u = (int)(total * cta / n_cta);
        u1 = (int)(total * (cta + 1) / n_cta);
        s.cum4 = s.cum5 = s.cum6 = s.cum7 = s.cumM4 = 0;
        s.base4 = s.base5 = s.base6 = 0;
        s.pair = 0; s.x0 = 0; s.y0 = 0; s.y1 = 0;
        first = true;
    }
    R2L_HD bool next() {
        if (!first) {                             // account for the segment just finished
            s.cum4 += Seg6::passes(s.m4()); s.cum5 += Seg6::passes(s.m5()); s.cum6 += Seg6::passes(s.m6());
            s.cum7 += s.n() / 2; s.cumM4 += s.m4() / 2;
            s.base4 += s.m4(); s.base5 += s.m5(); s.base6 += s.m6();
        }
        first = false;
        if (u >= u1) return false;
        const int strip = u / upg, g0 = u - strip * upg;
        const int g1 = imin(upg, g0 + (u1 - u));
        s.pair = strip / ncols;
        s.x0 = (strip - s.pair * ncols) * kB6SW;
        s.y0 = 2 * g0; s.y1 = 2 * g1;
        u += g1 - g0;
        return true;
    }
};

// ---- synchronisation words ------------------------------------------------------------------------------------------
#ifdef R2L_HOST_EMU
typedef std::atomic<int> sync6_t;
inline int ld_acq(const sync6_t* p) { return p->load(std::memory_order_acquire); }
inline void st_rel(sync6_t* p, int v) { p->store(v, std::memory_order_release); }
struct Emu6Abort {};
inline void backoff6(long long& spins) {
    if (++spins > 200000000LL) throw Emu6Abort();
    if ((spins & 255) == 0) std::this_thread::yield();
}
#else
typedef int sync6_t;
__device__ __forceinline__ int ld_acq(const sync6_t* p) {
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void st_rel(sync6_t* p, int v) {
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const void* tmap, int x, int y, int z, int w, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(w), "r"(smem_u32(bar)) : "memory");
}
#endif

// progress words in shared memory: eight per array
struct Sync6 {
    sync6_t done4[8], done5[8], done6[8];         // passes completed by warp w of B4 / B5 / B6
    sync6_t low5[8], low6[8], low7a[8], low7b[8]; // lowest virtual row a consumer warp may still read: B5 in gY2, B6 in gY1,
                                                  // B7 in gY0, B7 in gU / gV
    sync6_t pad[8];
};
static_assert(sizeof(Sync6) == 64 * 4, "kSyncInts");

// everything a role needs, passed by reference
template <class Cfg> struct Ctx6 {
    const BwdArgs* a;
    Tables2* T2;
    Sync6* sy;
    f2 *gY2, *gY1, *gY0, *gU, *gV;
    float* stage;
    float* red;                                   // [warp][128] statistics hand-over
#ifdef R2L_HOST_EMU
    sync6_t* full_seq; sync6_t* empty_seq;        // emulation of the staging ring's mbarriers
    float* park;                                  // [N7][32 lanes][3 * kB6K]
#else
    uint64_t *full_bar, *empty_bar;
    const void *tmap_g, *tmap_o;
    uint32_t tmem_base;
#endif
    int cta, n_cta;
};

#ifdef R2L_HOST_EMU
#define R2L6_LANES for (int lane = 0; lane < 32; ++lane)
#define R2L6_WARPSYNC()
#define R2L6_L(x) x[lane]
#define R2L6_DECL(type, name) type name[32]
#define R2L6_DECL_ARR(type, name, n) type name[32][n]
#else
#define R2L6_LANES const int lane = threadIdx.x & 31;
#define R2L6_WARPSYNC() __syncwarp()
#define R2L6_L(x) x
#define R2L6_DECL(type, name) type name
#define R2L6_DECL_ARR(type, name, n) type name[n]
#endif

// all passes of a stage with index <= J complete?  (pass i belongs to warp i % n and is that warp's (i / n)-th pass)
R2L_HD bool passes_done(const sync6_t* done, int n, int J) {
    bool ok = true;
    for (int w = 0; w < n; ++w) {
        const int cnt = J >= w ? (J - w) / n + 1 : 0;
        ok = ok && (ld_acq(done + w) >= cnt);
    }
    return ok;
}
// every consumer warp has moved past virtual row `row`?
R2L_HD bool rows_free(const sync6_t* low, int n, int row) {
    bool ok = true;
    for (int w = 0; w < n; ++w) ok = ok && (ld_acq(low + w) > row);
    return ok;
}

#ifdef R2L_HOST_EMU
#define R2L6_WAIT(cond) { long long spins_ = 0; while (!(cond)) backoff6(spins_); }
#else
#define R2L6_WAIT(cond) { while (!(cond)) __nanosleep(40); __syncwarp(); }
#endif

// ---------------------------------------------------------------------------------------------------------------------
// B4: grad_out, out -> (gY2, gU, gV) for one 4-site run of both images; gamma statistic.  Same arithmetic as isp_bwd5.cuh.
// ga/gb/ya/yb[k]: grad_out / out of image A / B, plane k (zeros outside the image / for a missing partner image).
// ---------------------------------------------------------------------------------------------------------------------
struct B4Const {
    float m2g[9], invg, gam, one_m_g;
    unsigned m_base, m_span;
    float t_gs[3], t_c1[3], t_c2[3], t_isc[3], t_osh[3];
};
template <class Cfg> R2L_HD void b4_const(const BwdArgs& a, const Tables* T, B4Const& c) {
    c.invg = T->invg; c.gam = T->gamma; c.one_m_g = 1.0f - T->gamma;
#pragma unroll
    for (int t = 0; t < 9; ++t) c.m2g[t] = T->M2[t] * c.invg;
    const float o_lo_exact = fast_exp2(c.invg * fast_log2(kClipLo));
    const float o_lo = Cfg::TAIL ? o_lo_exact * (1.0f + 1e-4f) : o_lo_exact;
    const float o_hi = Cfg::TAIL ? 1.0f - 1e-6f : 1.0f;
    c.m_base = fbits(o_lo) + 1u; c.m_span = fbits(o_hi) - c.m_base;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        c.t_gs[k] = 1.f; c.t_c1[k] = 0.f; c.t_c2[k] = 0.f; c.t_isc[k] = 1.f; c.t_osh[k] = 0.f;
        if (Cfg::TAIL) {
            c.t_gs[k] = a.gtail[k]; c.t_c1[k] = a.gtail[3 + k]; c.t_c2[k] = a.gtail[6 + k];
            c.t_isc[k] = 1.0f / a.gtail[9 + k]; c.t_osh[k] = -a.gtail[12 + k] * c.t_isc[k];
        }
    }
}
// valid: the run lies inside the image (and the loads were real); stat: it carries the gamma statistic
template <class Cfg>
R2L_HD void b4_item(const B4Const& c, const f4 (&ga)[3], const f4 (&gb)[3], const f4 (&ya)[3], const f4 (&yb)[3],
                    const f4 (&ad)[3], bool stat, f2& sg, f2 (&gy2)[4], f2 (&gu)[4], f2 (&gv)[4]) {
#pragma unroll
    for (int j = 0; j < 4; ++j) { gy2[j] = mk2(0.f, 0.f); gu[j] = mk2(0.f, 0.f); gv[j] = mk2(0.f, 0.f); }
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float gak[4] = {ga[k].x, ga[k].y, ga[k].z, ga[k].w}, gbk[4] = {gb[k].x, gb[k].y, gb[k].z, gb[k].w};
        const float yak[4] = {ya[k].x, ya[k].y, ya[k].z, ya[k].w}, ybk[4] = {yb[k].x, yb[k].y, yb[k].z, yb[k].w};
        const float adk[4] = {ad[k].x, ad[k].y, ad[k].z, ad[k].w};
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float Ga = gak[j], Gb = gbk[j];
            f2 o = mk2(yak[j], ybk[j]);
            if (Cfg::TAIL) {
                Ga = c.t_gs[k] * (Ga - c.t_c1[k] - c.t_c2[k] * o.x);
                Gb = c.t_gs[k] * (Gb - c.t_c1[k] - c.t_c2[k] * o.y);
                o = mk2(fmaf_(o.x, c.t_isc[k], c.t_osh[k]) - adk[j], fmaf_(o.y, c.t_isc[k], c.t_osh[k]) - adk[j]);
            }
            const f2 lo = mk2(fast_log2(o.x), fast_log2(o.y));
            const f2 ex = mul2s(lo, c.one_m_g);
            const f2 e = mk2(fast_exp2(ex.x), fast_exp2(ex.y));
            if (stat) sg = fma2vv(mk2(Ga * o.x, Gb * o.y), lo, sg);
            const f2 gr = mk2((fbits(o.x) - c.m_base < c.m_span) ? Ga * e.x : 0.f,
                              (fbits(o.y) - c.m_base < c.m_span) ? Gb * e.y : 0.f);
            gy2[j] = fma2s(gr, c.m2g[k * 3 + 0], gy2[j]);
            gu[j] = fma2s(gr, c.m2g[k * 3 + 1], gu[j]);
            gv[j] = fma2s(gr, c.m2g[k * 3 + 2], gv[j]);
        }
    }
}

// plane row of virtual row v in a ring of R rows
template <int R> R2L_HD int ring_row(int v) { return (v & (R - 1)) * kB6P; }

// ---------------------------------------------------------------------------------------------------------------------
// role: B4
// ---------------------------------------------------------------------------------------------------------------------
template <class Cfg, typename RawT>
R2L_HD void b6_role_b4(const Ctx6<Cfg>& cx, int w) {
    const BwdArgs& a = *cx.a;
    const Tables* T = &cx.T2->base;
    const int H = a.H, W = a.W;
    const int plane_i = H * W;
    B4Const c;
    b4_const<Cfg>(a, T, c);
    R2L6_DECL(f2, sg);
    { R2L6_LANES { R2L6_L(sg) = mk2(0.f, 0.f); } }
    int mine = 0;                                                    // passes this warp has completed
    SegIter6 it(cx.cta, cx.n_cta, a.B, H, W);
    while (it.next()) {
        const Seg6& s = it.s;
        const int np = Seg6::passes(s.m4()), G4 = s.m4() / 2;
        const int b0 = 2 * s.pair, b1 = b0 + 1;
        const bool dup = b1 >= a.B;                                  // odd batch, last pair: stream B carries nothing
        const float* goA = a.gout + (size_t)b0 * 3 * plane_i;
        const float* yoA = a.out + (size_t)b0 * 3 * plane_i;
        int li = ((w - s.cum4) % Cfg::N4 + Cfg::N4) % Cfg::N4;
        for (; li < np; li += Cfg::N4) {
            const int h = li / 9, rem = li - 9 * h;
            const bool halo = rem == 0;
            const int g = 8 * h + rem - 1;                           // main group
            const int rho0 = halo ? 16 * h : 2 * g;                  // first row of the pass (relative to y0 - 4)
            const int rho1 = halo ? imin(16 * h + 15, s.m4() - 1) : 2 * g + 1;
            // the ring rows this pass overwrites must have been read by every consumer
            R2L6_WAIT(rows_free(cx.sy->low5, Cfg::N5, s.base4 + rho1 - kB6R4) &&
                      rows_free(cx.sy->low7b, Cfg::N7, s.base4 + rho1 - kB6RUV))
            const int m = s.cumM4 + g, slot = m % kB6NS;
            const float* st = cx.stage + (size_t)slot * kB6SlotFloats;
            if (!halo) {
#ifdef R2L_HOST_EMU
                R2L6_WAIT(ld_acq(cx.full_seq + slot) == m + 1)
#else
                mbar_wait(cx.full_bar + slot, (uint32_t)((m / kB6NS) & 1));
#endif
            }
            { R2L6_LANES {
                const int rr = halo ? (lane >> 1) : (lane >> 4);
                const int gr = halo ? ((lane & 1) ? kB6G : -1) : (lane & 15);
                const int rho = rho0 + rr;
                const int gy = s.y0 - 4 + rho, gx = s.x0 + 4 * gr;
                const bool inrange = rho <= rho1;
                const bool valid = inrange && gy >= 0 && gy < H && gx >= 0 && gx < W;
                f4 ga[3], gb[3], ya[3], yb[3], ad[3];
#pragma unroll
                for (int k = 0; k < 3; ++k) {
                    ga[k].x = ga[k].y = ga[k].z = ga[k].w = 0.f;
                    gb[k] = ga[k]; ya[k] = ga[k]; yb[k] = ga[k]; ad[k] = ga[k];
                }
                if (!halo) {
                    // staged by the copy engine: [tensor][image][plane][row][64]; zeros outside the image / batch
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float* p = st + (k * 2 + rr) * 64 + 4 * gr;
                        ga[k] = *reinterpret_cast<const f4*>(p);
                        gb[k] = *reinterpret_cast<const f4*>(p + 384);
                        ya[k] = *reinterpret_cast<const f4*>(p + 768);
                        yb[k] = *reinterpret_cast<const f4*>(p + 768 + 384);
                    }
                } else if (valid) {
                    const int pix = gy * W + gx;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int off = k * plane_i + pix;
                        ga[k] = ld_stream4(goA + off);
                        ya[k] = ld_stream4(yoA + off);
                        if (!dup) {
                            gb[k] = ld_stream4(goA + (off + 3 * plane_i));
                            yb[k] = ld_stream4(yoA + (off + 3 * plane_i));
                        }
                    }
                }
                if (Cfg::TAIL && a.additive && valid) {
                    const int pix = gy * W + gx;
#pragma unroll
                    for (int k = 0; k < 3; ++k) ad[k] = *reinterpret_cast<const f4*>(a.additive + (size_t)k * plane_i + pix);
                }
                const bool stat = !halo && valid && gy >= s.y0 && gy < s.y1;
                f2 gy2[4], gu[4], gv[4];
                f2 sgl = mk2(0.f, 0.f);
                b4_item<Cfg>(c, ga, gb, ya, yb, ad, stat, sgl, gy2, gu, gv);
                if (!valid) {                                        // (behind a tail a zero-filled site does not map to zero)
#pragma unroll
                    for (int j = 0; j < 4; ++j) { gy2[j] = mk2(0.f, 0.f); gu[j] = mk2(0.f, 0.f); gv[j] = mk2(0.f, 0.f); }
                } else if (dup) {                                    // odd batch, last pair: stream B carries no gradient
#pragma unroll
                    for (int j = 0; j < 4; ++j) { gy2[j].y = 0.f; gu[j].y = 0.f; gv[j].y = 0.f; }
                }
                if (stat) {                                          // (the zero-filled partner of an odd batch: 0 * log2(0))
                    R2L6_L(sg).x = fmaf_(sgl.x, c.gam, R2L6_L(sg).x);
                    if (!dup) R2L6_L(sg).y = fmaf_(sgl.y, c.gam, R2L6_L(sg).y);
                }
                if (inrange) {
                    const int v = s.base4 + rho;
                    st4<kB6P>(cx.gY2, ring_row<kB6R4>(v) + 2 * (gr + 2), gy2[0], gy2[1], gy2[2], gy2[3]);
                    st4<kB6P>(cx.gU, ring_row<kB6RUV>(v) + 2 * (gr + 2), gu[0], gu[1], gu[2], gu[3]);
                    st4<kB6P>(cx.gV, ring_row<kB6RUV>(v) + 2 * (gr + 2), gv[0], gv[1], gv[2], gv[3]);
                }
            } }
            R2L6_WARPSYNC();
            ++mine;
#ifdef R2L_HOST_EMU
            if (!halo) st_rel(cx.empty_seq + slot, m + 1);
            st_rel(cx.sy->done4 + w, mine);
#else
            if ((threadIdx.x & 31) == 0) {
                if (!halo) mbar_arrive(cx.empty_bar + slot);
                st_rel(cx.sy->done4 + w, mine);
            }
#endif
        }
    }
    // gamma statistic of this warp -> red[w][0..1]
#ifdef R2L_HOST_EMU
    { float sx = 0.f, sy2 = 0.f; for (int l = 0; l < 32; ++l) { sx += sg[l].x; sy2 += sg[l].y; } cx.red[w * 128 + 0] = sx + sy2; }
#else
    { const float v = warp_sum_all(sg.x + sg.y); if ((threadIdx.x & 31) == 0) cx.red[w * 128] = v; }
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// role: TMA producer (one lane): grad_out / out boxes of B4's main passes, in pass order
// ---------------------------------------------------------------------------------------------------------------------
template <class Cfg>
R2L_HD void b6_role_tma(const Ctx6<Cfg>& cx) {
    const BwdArgs& a = *cx.a;
    SegIter6 it(cx.cta, cx.n_cta, a.B, a.H, a.W);
    while (it.next()) {
        const Seg6& s = it.s;
        const int G4 = s.m4() / 2;
        for (int g = 0; g < G4; ++g) {
            const int m = s.cumM4 + g, slot = m % kB6NS;
            float* dst = cx.stage + (size_t)slot * kB6SlotFloats;
            const int gy = s.y0 - 4 + 2 * g, b0 = 2 * s.pair;
#ifdef R2L_HOST_EMU
            if (m >= kB6NS) R2L6_WAIT(ld_acq(cx.empty_seq + slot) == m + 1 - kB6NS)
            const int H = a.H, W = a.W;
            for (int t = 0; t < 2; ++t)
                for (int im = 0; im < 2; ++im)
                    for (int k = 0; k < 3; ++k)
                        for (int r = 0; r < 2; ++r)
                            for (int x = 0; x < 64; ++x) {
                                const int y = gy + r, xx = s.x0 + x, b = b0 + im;
                                float v = 0.f;
                                if (y >= 0 && y < H && xx < W && b < a.B)
                                    v = (t ? a.out : a.gout)[(((size_t)b * 3 + k) * H + y) * W + xx];
                                dst[(((t * 2 + im) * 3 + k) * 2 + r) * 64 + x] = v;
                            }
            st_rel(cx.full_seq + slot, m + 1);
#else
            mbar_wait(cx.empty_bar + slot, (uint32_t)(((m / kB6NS) & 1) ^ 1));
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;"
                         ::"r"(smem_u32(cx.full_bar + slot)), "r"((uint32_t)(kB6SlotFloats * 4)) : "memory");
            tma_load_4d(dst, cx.tmap_g, s.x0, gy, 0, b0, cx.full_bar + slot);
            tma_load_4d(dst + kB6SlotFloats / 2, cx.tmap_o, s.x0, gy, 0, b0, cx.full_bar + slot);
#endif
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------
// role: B5 -- gY1 = fold_reflect2(corr^T(gY2, Wg)); dWg.  Item = one run (q.y, q.x .. q.x+3); border rules as isp_bwd5.cuh.
// ---------------------------------------------------------------------------------------------------------------------
template <class Cfg, typename RawT>
R2L_HD void b6_role_b5(const Ctx6<Cfg>& cx, int w) {
    const BwdArgs& a = *cx.a;
    const Tables* T = &cx.T2->base;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const size_t luma_plane = (size_t)((a.B + 1) >> 1) * plane * 2;
#ifdef R2L_HOST_EMU
    const float* wg = T->Wg;
#else
    const volatile float* wg = T->Wg;              // read next to their use (uniform-address LDS), not hoisted
#endif
    R2L6_DECL_ARR(f2, acc, 25);
    { R2L6_LANES {
#pragma unroll
        for (int i = 0; i < 25; ++i) R2L6_L(acc)[i] = mk2(0.f, 0.f);
    } }
    int mine = 0;
    SegIter6 it(cx.cta, cx.n_cta, a.B, H, W);
    while (it.next()) {
        const Seg6& s = it.s;
        const int np = Seg6::passes(s.m5()), G4 = s.m4() / 2;
        const float* y1pair = a.luma + (size_t)s.pair * plane * 2 + luma_plane;      // Y1 of this image pair, [H][W][2]
        int li = ((w - s.cum5) % Cfg::N5 + Cfg::N5) % Cfg::N5;
        for (; li < np; li += Cfg::N5) {
            const int h = li / 9, rem = li - 9 * h;
            const bool halo = rem == 0;
            const int g = 8 * h + rem - 1;
            const int rho0 = halo ? 16 * h : 2 * g;                  // relative to y0 - 2
            const int rho1 = halo ? imin(16 * h + 15, s.m5() - 1) : 2 * g + 1;
#ifdef R2L_HOST_EMU
            st_rel(cx.sy->low5 + w, s.base4 + rho0);
#else
            if ((threadIdx.x & 31) == 0) st_rel(cx.sy->low5 + w, s.base4 + rho0);   // gY2 rows rho0 .. (B4 numbering: same offset)
#endif
            // centres first: their global loads run while the warp waits for its window rows
            const int J4 = s.cum4 + Seg6::idx_main(imin(halo ? 8 * h + 10 : g + 2, G4 - 1));
            R2L6_WAIT(passes_done(cx.sy->done4, Cfg::N4, J4) && rows_free(cx.sy->low6, Cfg::N6, s.base5 + rho1 - kB6R5))
            { R2L6_LANES {
                const int rr = halo ? (lane >> 1) : (lane >> 4);
                const int gr = halo ? ((lane & 1) ? kB6G : -1) : (lane & 15);
                const int rho = rho0 + rr;
                const int qy = s.y0 - 2 + rho, qx = s.x0 + 4 * gr;
                const bool inrange = rho <= rho1;
                const bool inside = inrange && qy >= 0 && qy < H && qx >= 0 && qx < W;
                const bool stat = !halo && inside && qy >= s.y0 && qy < s.y1;
                f2 c[4], out[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { c[j] = mk2(0.f, 0.f); out[j] = mk2(0.f, 0.f); }
                if (stat) ld_luma4(y1pair + ((size_t)qy * W + qx) * 2, c);
                const bool lft = qx == 0, rgt = qx + 4 == W;
                int rt = qy == 1 ? 1 : (qy == 2 ? 2 : (qy == H - 2 ? 3 : (qy == H - 3 ? 4 : 0)));
                if (!inside) rt = 0;
                const float Lf = (inside && lft) ? 1.f : 0.f, Rf = (inside && rgt) ? 1.f : 0.f;
                const f2 cl1 = mul2s(c[1], Lf), cl2 = mul2s(c[2], Lf), cr1 = mul2s(c[1], Rf), cr2 = mul2s(c[2], Rf);
                // window row d = gY2 row q.y - 2 + d (B4 numbering: rho + d), columns q.x - 2 .. q.x + 5; tap row A = 4 - d
                auto full = [&](int d, auto Ac, auto statc) {
                    constexpr int A = decltype(Ac)::value;
                    constexpr bool with_stat = decltype(statc)::value;
                    float w5[5];
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) w5[bb] = wg[A * 5 + bb];
                    const float w1_0 = fmaf_(Rf, w5[4], w5[0]), w1_2 = fmaf_(Lf, w5[0], w5[2]), w1_3 = fmaf_(Lf, w5[1], w5[3]);
                    const float w2_1 = fmaf_(Rf, w5[3], w5[1]), w2_2 = fmaf_(Rf, w5[4], w5[2]), w2_4 = fmaf_(Lf, w5[0], w5[4]);
                    f2 row[8];
                    ld8<kB6P>(cx.gY2, ring_row<kB6R4>(s.base4 + rho + d) + 2 * (gr + 2), row);
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) {
                        out[0] = fma2s(row[4 - bb], w5[bb], out[0]);
                        out[3] = fma2s(row[7 - bb], w5[bb], out[3]);
                    }
                    out[1] = fma2s(row[5], w1_0, fma2s(row[4], w5[1], fma2s(row[3], w1_2, fma2s(row[2], w1_3, fma2s(row[1], w5[4], out[1])))));
                    out[2] = fma2s(row[6], w5[0], fma2s(row[5], w2_1, fma2s(row[4], w2_2, fma2s(row[3], w5[3], fma2s(row[2], w2_4, out[2])))));
                    if (with_stat) {
                        f2* ac = R2L6_L(acc) + A * 5;
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb)
#pragma unroll
                            for (int j = 0; j < 4; ++j) ac[bb] = fma2vv(c[j], row[j + 4 - bb], ac[bb]);
                        ac[0] = fma2vv(cl1, row[3], fma2vv(cl2, row[2], ac[0]));
                        ac[1] = fma2vv(cl1, row[2], ac[1]);
                        ac[3] = fma2vv(cr2, row[5], ac[3]);
                        ac[4] = fma2vv(cr2, row[4], fma2vv(cr1, row[5], ac[4]));
                    }
                };
                using std::integral_constant;
                typedef integral_constant<bool, true> Yes;
                typedef integral_constant<bool, false> No;
                if (halo) {
                    if (inside) {
                        full(0, integral_constant<int, 4>(), No()); full(1, integral_constant<int, 3>(), No());
                        full(2, integral_constant<int, 2>(), No()); full(3, integral_constant<int, 1>(), No());
                        full(4, integral_constant<int, 0>(), No());
                    }
                } else {
                    full(0, integral_constant<int, 4>(), Yes()); full(1, integral_constant<int, 3>(), Yes());
                    full(2, integral_constant<int, 2>(), Yes()); full(3, integral_constant<int, 1>(), Yes());
                    full(4, integral_constant<int, 0>(), Yes());
                }
                // folded pad rows: row type rt acts through tap row A on window row d: (rt 1: A 0 / d 2, A 1 / d 1),
                // (rt 2: A 0 / d 0), (rt 3: A 3 / d 3, A 4 / d 2), (rt 4: A 4 / d 4); only warps holding such a row come here
                // (halo items run it with zero centres: their statistic share is nil)
                if (R2L_ANY(rt != 0)) {
                    if (rt == 1) { full(2, integral_constant<int, 0>(), Yes()); full(1, integral_constant<int, 1>(), Yes()); }
                    if (rt == 2) full(0, integral_constant<int, 0>(), Yes());
                    if (rt == 3) { full(3, integral_constant<int, 3>(), Yes()); full(2, integral_constant<int, 4>(), Yes()); }
                    if (rt == 4) full(4, integral_constant<int, 4>(), Yes());
                }
                if (!inside) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) out[j] = mk2(0.f, 0.f);
                }
                if (inrange) st4<kB6P>(cx.gY1, ring_row<kB6R5>(s.base5 + rho) + 2 * (gr + 2), out[0], out[1], out[2], out[3]);
            } }
            R2L6_WARPSYNC();
            ++mine;
#ifdef R2L_HOST_EMU
            st_rel(cx.sy->done5 + w, mine);
#else
            if ((threadIdx.x & 31) == 0) st_rel(cx.sy->done5 + w, mine);
#endif
        }
    }
#ifdef R2L_HOST_EMU
    st_rel(cx.sy->low5 + w, 0x7fffffff);
    for (int i = 0; i < 25; ++i) { float t = 0.f; for (int l = 0; l < 32; ++l) t += acc[l][i].x + acc[l][i].y; cx.red[(Cfg::W5 + w) * 128 + i] = t; }
#else
    if ((threadIdx.x & 31) == 0) st_rel(cx.sy->low5 + w, 0x7fffffff);
    {
        float* red = cx.red + (Cfg::W5 + w) * 128;
#pragma unroll
        for (int i = 0; i < 25; ++i) { const float v = warp_sum_all(acc[i].x + acc[i].y); if ((threadIdx.x & 31) == 0) red[i] = v; }
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// role: B6 -- gY0 = corr^T(gY1, Ws) (zero pad); dWs.  Groups are rows (y0 - 1 + 2j, y0 + 2j).
// ---------------------------------------------------------------------------------------------------------------------
template <class Cfg, typename RawT>
R2L_HD void b6_role_b6(const Ctx6<Cfg>& cx, int w) {
    const BwdArgs& a = *cx.a;
    const Tables* T = &cx.T2->base;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    float ws[9];
#pragma unroll
    for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
    R2L6_DECL_ARR(f2, acc, 9);
    { R2L6_LANES {
#pragma unroll
        for (int i = 0; i < 9; ++i) R2L6_L(acc)[i] = mk2(0.f, 0.f);
    } }
    int mine = 0;
    SegIter6 it(cx.cta, cx.n_cta, a.B, H, W);
    while (it.next()) {
        const Seg6& s = it.s;
        const int np = Seg6::passes(s.m6()), G5 = s.m5() / 2;
        const float* y0pair = a.luma + (size_t)s.pair * plane * 2;                   // Y0 of this image pair
        int li = ((w - s.cum6) % Cfg::N6 + Cfg::N6) % Cfg::N6;
        for (; li < np; li += Cfg::N6) {
            const int h = li / 9, rem = li - 9 * h;
            const bool halo = rem == 0;
            const int g = 8 * h + rem - 1;
            const int rho0 = halo ? 16 * h : 2 * g;                  // relative to y0 - 1
            const int rho1 = halo ? imin(16 * h + 15, s.m6() - 1) : 2 * g + 1;
#ifdef R2L_HOST_EMU
            st_rel(cx.sy->low6 + w, s.base5 + rho0);
#else
            if ((threadIdx.x & 31) == 0) st_rel(cx.sy->low6 + w, s.base5 + rho0);   // gY1 rows rho0 .. (B5 numbering: same offset)
#endif
            const int J5 = s.cum5 + Seg6::idx_main(imin(halo ? 8 * h + 9 : g + 1, G5 - 1));
            R2L6_WAIT(passes_done(cx.sy->done5, Cfg::N5, J5) && rows_free(cx.sy->low7a, Cfg::N7, s.base6 + rho1 - kB6R6))
            { R2L6_LANES {
                const int rr = halo ? (lane >> 1) : (lane >> 4);
                const int gr = halo ? ((lane & 1) ? kB6G : -1) : (lane & 15);
                const int rho = rho0 + rr;
                const int qy = s.y0 - 1 + rho, qx = s.x0 + 4 * gr;
                const bool inrange = rho <= rho1;
                const bool inside = inrange && qy >= 0 && qy < H && qx >= 0 && qx < W;
                const bool stat = !halo && inside && qy >= s.y0 && qy < s.y1;
                f2 c[4], out[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { c[j] = mk2(0.f, 0.f); out[j] = mk2(0.f, 0.f); }
                if (stat) ld_luma4(y0pair + ((size_t)qy * W + qx) * 2, c);
                if (!halo || inside) {
#pragma unroll
                    for (int d = 0; d < 3; ++d) {
                        f2 row[6];                                   // gY1 row q.y - 1 + d (B5 numbering: rho + d), columns q.x - 1 .. q.x + 4
                        ld6<kB6P>(cx.gY1, ring_row<kB6R5>(s.base5 + rho + d) + 2 * (gr + 2), row);
                        const int aa = 2 - d;
#pragma unroll
                        for (int j = 0; j < 4; ++j)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) out[j] = fma2s(row[j + 2 - bb], ws[aa * 3 + bb], out[j]);
                        if (!halo) {
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    R2L6_L(acc)[aa * 3 + bb] = fma2vv(c[j], row[j + 2 - bb], R2L6_L(acc)[aa * 3 + bb]);
                        }
                    }
                }
                if (!inside) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) out[j] = mk2(0.f, 0.f);
                }
                if (inrange) st4<kB6P>(cx.gY0, ring_row<kB6R6>(s.base6 + rho) + 2 * (gr + 2), out[0], out[1], out[2], out[3]);
            } }
            R2L6_WARPSYNC();
            ++mine;
#ifdef R2L_HOST_EMU
            st_rel(cx.sy->done6 + w, mine);
#else
            if ((threadIdx.x & 31) == 0) st_rel(cx.sy->done6 + w, mine);
#endif
        }
    }
#ifdef R2L_HOST_EMU
    st_rel(cx.sy->low6 + w, 0x7fffffff);
    for (int i = 0; i < 9; ++i) { float t = 0.f; for (int l = 0; l < 32; ++l) t += acc[l][i].x + acc[l][i].y; cx.red[(Cfg::W6 + w) * 128 + i] = t; }
#else
    if ((threadIdx.x & 31) == 0) st_rel(cx.sy->low6 + w, 0x7fffffff);
    {
        float* red = cx.red + (Cfg::W6 + w) * 128;
#pragma unroll
        for (int i = 0; i < 9; ++i) { const float v = warp_sum_all(acc[i].x + acc[i].y); if ((threadIdx.x & 31) == 0) red[i] = v; }
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// role: B7 -- Q' / P statistics and g_raw from the (gY0, gU, gV) windows; border rules as isp_bwd5.cuh B7.
// One gradient plane k at a time: its 18 packed Q' sums ([tap row][col phase][b]) and 2 packed P sums come from tensor
// memory, take the run's 36 + 2 products, and go back.  Lanes 0-15 hold the even row of the group, 16-31 the odd row,
// so a lane's CFA row phase never changes.
// ---------------------------------------------------------------------------------------------------------------------
template <class Cfg, typename RawT>
R2L_HD void b6_role_b7(const Ctx6<Cfg>& cx, int w) {
    const BwdArgs& a = *cx.a;
    const Tables* T = &cx.T2->base;
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const float one = a.B > 0 ? 1.f : 2.f;                           // 1.0 the compiler cannot see (pack2)
#ifndef R2L_HOST_EMU
    const int wabs = Cfg::W7 + w;                                    // TMEM lane quadrant = warp % 4; same-quadrant warps take
    int same = 0;                                                    // successive column blocks
    for (int v = 0; v < w; ++v) same += (((Cfg::W7 + v) & 3) == (wabs & 3)) ? 1 : 0;
    const uint32_t tacc = tmem::addr(cx.tmem_base, same * kB6ParkCols);
    {
        float z[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < 3 * kB6K; c0 += 8) tmem::store<8>(tacc + c0, z);
        tmem::wait_st();
    }
#endif
    SegIter6 it(cx.cta, cx.n_cta, a.B, H, W);
    while (it.next()) {
        const Seg6& s = it.s;
        const int np = s.n() / 2, G6 = s.m6() / 2, G4 = s.m4() / 2;
        const int b0 = 2 * s.pair, b1 = b0 + 1 < a.B ? b0 + 1 : b0;
        const bool dup = b1 == b0;
        const RawT* imgA = static_cast<const RawT*>(a.raw) + (size_t)b0 * plane;
        const RawT* imgB = static_cast<const RawT*>(a.raw) + (size_t)b1 * plane;
        int li = ((w - s.cum7) % Cfg::N7 + Cfg::N7) % Cfg::N7;
        for (; li < np; li += Cfg::N7) {
            const int j7 = li;
#ifdef R2L_HOST_EMU
            st_rel(cx.sy->low7a + w, s.base6 + 2 * j7);
            st_rel(cx.sy->low7b + w, s.base4 + 2 * j7 + 3);
#else
            if ((threadIdx.x & 31) == 0) {
                st_rel(cx.sy->low7a + w, s.base6 + 2 * j7);          // gY0 rows y0+2j-1 .. (B6 numbering: 2j ..)
                st_rel(cx.sy->low7b + w, s.base4 + 2 * j7 + 3);      // gU / gV rows (B4 numbering: 2j+3 ..)
            }
#endif
#ifndef R2L_HOST_EMU
            // raw centres: requested before the wait
            f4 xa, xb;
            xa.x = xa.y = xa.z = xa.w = 0.f; xb = xa;
            {
                const int lane = threadIdx.x & 31;
                const int qy = s.y0 + 2 * j7 + (lane >> 4), qx = s.x0 + 4 * (lane & 15);
                if (sizeof(RawT) == 4 && qy < H && qx < W) {
                    xa = ld_stream4(reinterpret_cast<const float*>(imgA) + (size_t)qy * W + qx);
                    xb = ld_stream4(reinterpret_cast<const float*>(imgB) + (size_t)qy * W + qx);
                }
            }
#endif
            const int J6 = s.cum6 + Seg6::idx_main(imin(j7 + 1, G6 - 1));
            const int J4 = s.cum4 + Seg6::idx_main(imin(j7 + 3, G4 - 1));
            R2L6_WAIT(passes_done(cx.sy->done6, Cfg::N6, J6) && passes_done(cx.sy->done4, Cfg::N4, J4))
            { R2L6_LANES {
                const int rp = lane >> 4, g = lane & 15;
                const int qy = s.y0 + 2 * j7 + rp, qx = s.x0 + 4 * g;
                const bool live = qy < H && qx < W;
                const bool f_top = qy == 1, f_bot = qy == H - 2, f_lft = qx == 0, f_rgt = qx + 4 == W;
                f2 c[4], graw[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { c[j] = mk2(0.f, 0.f); graw[j] = mk2(0.f, 0.f); }
                if (live) {
#ifndef R2L_HOST_EMU
                    if (sizeof(RawT) == 4) {
                        c[0] = pack2(xa.x, xb.x, one); c[1] = pack2(xa.y, xb.y, one); c[2] = pack2(xa.z, xb.z, one); c[3] = pack2(xa.w, xb.w, one);
                    } else
#endif
                    {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            c[j] = pack2(RawLoad<RawT>::get(imgA + (size_t)qy * W + qx + j, a.denom),
                                         RawLoad<RawT>::get(imgB + (size_t)qy * W + qx + j, a.denom), one);
                    }
                }
                const float Lf = (live && f_lft) ? 1.f : 0.f, Rf = (live && f_rgt) ? 1.f : 0.f;
                const f2 cl1 = mul2s(c[1], Lf), cr2 = mul2s(c[2], Rf);
                const bool padrow = live && (f_top || f_bot);
                // window row d of plane pl (ring rows `vrow + d`): g_yuv[k] row q.y - 1 + d, columns q.x - 1 .. q.x + 4
                auto full = [&](const f2* pl, int prow, const float (&wt)[2][3], f2* acc) {
                    f2 row[6];
                    ld6<kB6P>(pl, prow + 2 * (g + 2), row);
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[(j & 1) * 3 + bb] = fma2vv(c[j], row[j + 2 - bb], acc[(j & 1) * 3 + bb]);
                    acc[3 + 0] = fma2vv(cl1, row[1], acc[3 + 0]);
                    acc[0 + 2] = fma2vv(cr2, row[4], acc[0 + 2]);
                    if (Cfg::GRAW) {
                        const float w1_2 = fmaf_(Lf, wt[1][0], wt[1][2]), w2_0 = fmaf_(Rf, wt[0][2], wt[0][0]);
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) {
                            graw[0] = fma2s(row[2 - bb], wt[0][bb], graw[0]);
                            graw[3] = fma2s(row[5 - bb], wt[1][bb], graw[3]);
                        }
                        graw[1] = fma2s(row[3], wt[1][0], fma2s(row[2], wt[1][1], fma2s(row[1], w1_2, graw[1])));
                        graw[2] = fma2s(row[4], w2_0, fma2s(row[3], wt[0][1], fma2s(row[2], wt[0][2], graw[2])));
                    }
                };
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    // k = 0: gY0 (ring of kB6R6 rows, B6 numbering: row q.y-1+d <-> 2j + rp + d); k = 1, 2: gU / gV (ring of
                    // kB6RUV rows, B4 numbering: row q.y-1+d <-> 2j + rp + 3 + d)
                    const f2* pl = k == 0 ? cx.gY0 : (k == 1 ? cx.gU : cx.gV);
                    const int v0 = k == 0 ? s.base6 + 2 * j7 + rp : s.base4 + 2 * j7 + rp + 3;
                    const int msk = k == 0 ? kB6R6 - 1 : kB6RUV - 1;
#ifdef R2L_HOST_EMU
                    const float* awq = &T->AWq[2 * rp][k][0];
                    float* qa = cx.park + ((size_t)w * 32 + lane) * (3 * kB6K) + k * kB6K;
#else
                    const volatile float* awq = &T->AWq[2 * rp][k][0];   // [col phase * 27 + tap], read next to their use
                    float qa[kB6K];
                    __syncwarp(); tmem::wait_st(); tmem::load<kB6K>(tacc + k * kB6K, qa);
                    tmem::ready<kB6K>(qa);
#endif
                    f2 acc[3][6];                                    // [tap row A][col phase * 3 + b]
#pragma unroll
                    for (int A = 0; A < 3; ++A)
#pragma unroll
                        for (int i = 0; i < 6; ++i) acc[A][i] = mk2(qa[(A * 6 + i) * 2], qa[(A * 6 + i) * 2 + 1]);
                    f2 p2[2] = {mk2(qa[36], qa[37]), mk2(qa[38], qa[39])};
#pragma unroll
                    for (int A = 0; A < 3; ++A) {
                        float wt[2][3];
#pragma unroll
                        for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) wt[cp][bb] = awq[cp * 27 + A * 3 + bb];
                        full(pl, ((v0 + 2 - A) & msk) * kB6P, wt, acc[A]);     // tap row A reaches window row 2 - A
                    }
                    {                                                // P[col phase] = sum of g_yuv[k] over the owned sites
                        f2 row[4];
                        ld4<kB6P>(pl, ((v0 + 1) & msk) * kB6P + 2 * (g + 2), row);
                        if (live) {
#pragma unroll
                            for (int j = 0; j < 4; ++j) p2[j & 1] = add2v(p2[j & 1], row[j]);
                        }
                    }
                    // reflect-1 pad rows: the row above row 1 acts through tap row 0 on window row 0, the row below row
                    // H-2 through tap row 2 on window row 2
                    if (R2L_ANY(padrow)) {
                        if (live && f_top) {
                            float wt[2][3];
#pragma unroll
                            for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                                for (int bb = 0; bb < 3; ++bb) wt[cp][bb] = awq[cp * 27 + bb];
                            full(pl, (v0 & msk) * kB6P, wt, acc[0]);
                        }
                        if (live && f_bot) {
                            float wt[2][3];
#pragma unroll
                            for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                                for (int bb = 0; bb < 3; ++bb) wt[cp][bb] = awq[cp * 27 + 6 + bb];
                            full(pl, ((v0 + 2) & msk) * kB6P, wt, acc[2]);
                        }
                    }
#pragma unroll
                    for (int A = 0; A < 3; ++A)
#pragma unroll
                        for (int i = 0; i < 6; ++i) { qa[(A * 6 + i) * 2] = acc[A][i].x; qa[(A * 6 + i) * 2 + 1] = acc[A][i].y; }
                    qa[36] = p2[0].x; qa[37] = p2[0].y; qa[38] = p2[1].x; qa[39] = p2[1].y;
#ifndef R2L_HOST_EMU
                    __syncwarp(); tmem::store<kB6K>(tacc + k * kB6K, qa);
#endif
                }
                if (Cfg::GRAW && live) {
                    const size_t off = (size_t)qy * W + qx;
                    float* pa = a.graw + (size_t)b0 * plane + off;
                    f4 va; va.x = graw[0].x; va.y = graw[1].x; va.z = graw[2].x; va.w = graw[3].x;
                    *reinterpret_cast<f4*>(pa) = va;
                    if (!dup) {
                        float* pb = a.graw + (size_t)b1 * plane + off;
                        f4 vb; vb.x = graw[0].y; vb.y = graw[1].y; vb.z = graw[2].y; vb.w = graw[3].y;
                        *reinterpret_cast<f4*>(pb) = vb;
                    }
                }
            } }
            R2L6_WARPSYNC();
        }
    }
#ifdef R2L_HOST_EMU
    st_rel(cx.sy->low7a + w, 0x7fffffff);
    st_rel(cx.sy->low7b + w, 0x7fffffff);
    // red[W7 + w][rp * 60 + i]: sums of the two row phases (lanes 0-15 / 16-31), (x + y) per packed sum
    for (int rp = 0; rp < 2; ++rp)
        for (int i = 0; i < 60; ++i) {
            float t = 0.f;
            for (int l = 16 * rp; l < 16 * rp + 16; ++l) {
                const float* qa = cx.park + ((size_t)w * 32 + l) * (3 * kB6K);
                t += qa[2 * i] + qa[2 * i + 1];
            }
            cx.red[(Cfg::W7 + w) * 128 + rp * 60 + i] = t;
        }
#else
    if ((threadIdx.x & 31) == 0) { st_rel(cx.sy->low7a + w, 0x7fffffff); st_rel(cx.sy->low7b + w, 0x7fffffff); }
    {
        float* red = cx.red + (Cfg::W7 + w) * 128;
        const int lane = threadIdx.x & 31;
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {
            float qa[kB6K];
            __syncwarp(); tmem::wait_st(); tmem::load<kB6K>(tacc + k * kB6K, qa);
            tmem::ready<kB6K>(qa);
#pragma unroll
            for (int i = 0; i < 20; ++i) {
                float v = qa[2 * i] + qa[2 * i + 1];
#pragma unroll
                for (int o = 8; o >= 1; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);    // within each half-warp
                if ((lane & 15) == 0) red[(lane >> 4) * 60 + k * 20 + i] = v;
            }
        }
        tmem::fence_before_sync();
    }
#endif
}

// ---------------------------------------------------------------------------------------------------------------------
// CTA: prologue (tables, rings, barriers, TMEM), roles, statistics hand-over
// ---------------------------------------------------------------------------------------------------------------------
// red[warp][128] (aliasing the gY2 ring once every role is done) -> the kStat* layout of this CTA's partial row:
//   B4 warps: [0] gamma | B5 warps: [0..24] dWg | B6 warps: [0..8] dWs | B7 warps: [rp * 60 + k * 20 + (A * 6 + cp * 3 + b | 18 + cp)]
template <class Cfg> R2L_HD float b6_stat(const float* red, int s) {
    float sum = 0.f;
    if (s == kStatGamma) {
        for (int w = 0; w < Cfg::N4; ++w) sum += red[w * 128];
    } else if (s < kStatWs) {
        for (int w = 0; w < Cfg::N5; ++w) sum += red[(Cfg::W5 + w) * 128 + (s - kStatWg)];
    } else if (s < kStatQ) {
        for (int w = 0; w < Cfg::N6; ++w) sum += red[(Cfg::W6 + w) * 128 + (s - kStatWs)];
    } else {
        int k, parp, tt = 0;
        bool is_q;
        if (s < kStatP) { const int rI = s - kStatQ; k = rI / 36; parp = (rI - 36 * k) / 9; tt = rI - 36 * k - 9 * parp; is_q = true; }
        else { const int rI = s - kStatP; k = rI / 4; parp = rI - 4 * k; is_q = false; }
        // Q[k][par(p)][t] = Q'[par(q) = par_tap(par(p), t)][k][t];  P is already p-indexed (p = q)
        const int parq = is_q ? par_tap(parp, tt) : parp;
        const int rpq = parq >> 1, cpq = parq & 1;
        const int off = rpq * 60 + k * 20 + (is_q ? 6 * (tt / 3) + 3 * cpq + tt % 3 : 18 + cpq);
        for (int w = 0; w < Cfg::N7; ++w) sum += red[(Cfg::W7 + w) * 128 + off];
    }
    return sum;
}

#ifndef R2L_HOST_EMU
template <class Cfg, typename RawT>
__device__ __forceinline__ void bwd6_cta(const BwdArgs& a, float* smem, const void* tmap_g, const void* tmap_o) {
    constexpr int NT = Cfg::NT;
    Ctx6<Cfg> cx;
    cx.a = &a;
    cx.T2 = reinterpret_cast<Tables2*>(smem);
    cx.sy = reinterpret_cast<Sync6*>(smem + Cfg::kTableFloats);
    cx.gY2 = reinterpret_cast<f2*>(smem + Cfg::kTableFloats + Cfg::kSyncInts);
    cx.gY1 = cx.gY2 + kB6R4 * kB6P;
    cx.gY0 = cx.gY1 + kB6R5 * kB6P;
    cx.gU = cx.gY0 + kB6R6 * kB6P;
    cx.gV = cx.gU + kB6RUV * kB6P;
    cx.red = reinterpret_cast<float*>(cx.gV + kB6RUV * kB6P);
    cx.stage = reinterpret_cast<float*>(reinterpret_cast<char*>(smem) + Cfg::kStageOffset);
    cx.full_bar = reinterpret_cast<uint64_t*>(reinterpret_cast<char*>(smem) + Cfg::kBarOffset);
    cx.empty_bar = cx.full_bar + kB6NS;
    cx.tmap_g = tmap_g; cx.tmap_o = tmap_o;
    cx.cta = blockIdx.x; cx.n_cta = gridDim.x;
    Tables* T = &cx.T2->base;
    const int tid = threadIdx.x, warp = tid >> 5;

    __shared__ uint32_t tmem_slot;
    if (warp == 0) tmem::alloc<Cfg::kTmemCols>(&tmem_slot);
    if (tid == 32) {
        for (int i = 0; i < kB6NS; ++i) { mbar_init(cx.full_bar + i, 1); mbar_init(cx.empty_bar + i, 1); }
    }
    for (int i = tid; i < Cfg::kSyncInts; i += NT) reinterpret_cast<int*>(cx.sy)[i] = 0;
    for (int i = tid; i < Cfg::kRingSites; i += NT) cx.gY2[i] = mk2(0.f, 0.f);      // rings start finite
    R2L_BUILD_TABLES(NT, a.P, T)
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    cx.tmem_base = tmem_slot;

    if (warp < Cfg::W5) b6_role_b4<Cfg, RawT>(cx, warp);
    else if (warp < Cfg::W6) b6_role_b5<Cfg, RawT>(cx, warp - Cfg::W5);
    else if (warp < Cfg::W7) b6_role_b6<Cfg, RawT>(cx, warp - Cfg::W6);
    else if (warp < Cfg::WT) b6_role_b7<Cfg, RawT>(cx, warp - Cfg::W7);
    else if ((tid & 31) == 0) b6_role_tma<Cfg>(cx);
    __syncthreads();                                                 // every role is done: rings are dead, red is complete
    if (warp == 0) tmem::dealloc<Cfg::kTmemCols>(cx.tmem_base);
    float* part = a.partials + (size_t)blockIdx.x * kStatPitch;
    for (int s = tid; s < kNumStats; s += NT) part[s] = b6_stat<Cfg>(cx.red, s);
    if (a.ticket) {
        __shared__ unsigned last_flag;
        __threadfence();                                             // this CTA's partial sums are visible device-wide ...
        __syncthreads();
        if (tid == 0) last_flag = atomicAdd(a.ticket, 1u) == (unsigned)gridDim.x - 1u;   // ... before its ticket is
        __syncthreads();
        if (last_flag) {
            __threadfence();
            static_assert((size_t)(NT / 32 + 1) * kStatPitch * 8 + 130 * 8 <= (size_t)Cfg::kRingSites * 8, "finish scratch fits the rings");
            fused_finish<NT>(T, a.partials, (int)gridDim.x, a.grads, reinterpret_cast<double*>(cx.gY2));
            if (a.world > 1) peer_allreduce<NT>(a);
        }
    }
}
#endif

}  // namespace r2l
