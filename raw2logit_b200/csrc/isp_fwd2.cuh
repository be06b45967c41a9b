// isp_fwd2.cuh -- second-generation fused forward: two images per lane (float2), register micro-tiles.
//
// What changed against fwd_cta (isp_core.cuh), and why (profiles/r01_v1_summary.md):
//   * a CTA processes the SAME tile of TWO images; every shared-memory plane holds float2 = (image A, image B)
//     at one site, so each stencil FMA is one FFMA2 whose multiplier is a scalar broadcast register
//     (SASS: FFMA2 Rd, Ra.F32x2.HI_LO, Rb.F32, Rc.F32x2.HI_LO) -- half the issue slots, no shuffles;
//   * weights live in registers, loaded once per phase with LDS.128 broadcasts;
//   * every thread computes 4-site runs (1x4 for the 3x3 stencils, 2x4 for the 5x5 Gaussian + colour tail) from
//     LDS.128 / LDS.64 window loads, so shared-memory traffic is ~100 B per site instead of ~520 B;
//   * U,V are produced together with Y0 (one pass over the raw window) and parked in shared memory.
// Border rules are applied as small fix-up passes on tiles that touch the image border, so the stencil loops
// carry no per-site branches: Y0 is zeroed on the 1-wide ring outside the image (sharpen zero-pads), Y1 is
// mirrored onto the 2-wide ring (Gaussian reflect-pads the sharpened plane).
#pragma once
#include "isp_core.cuh"

namespace r2l {

#ifdef R2L_HOST_EMU
struct f2 { float x, y; };
struct f4 { float x, y, z, w; };
R2L_HD f2 mk2(float x, float y) { f2 r; r.x = x; r.y = y; return r; }
R2L_HD f2 fma2s(f2 a, float w, f2 c) { return mk2(std::fma(a.x, w, c.x), std::fma(a.y, w, c.y)); }
R2L_HD f2 mul2s(f2 a, float w) { return mk2(a.x * w, a.y * w); }
#else
typedef float2 f2;
typedef float4 f4;
R2L_HD f2 mk2(float x, float y) { return make_float2(x, y); }
R2L_HD f2 fma2s(f2 a, float w, f2 c) { return __ffma2_rn(a, make_float2(w, w), c); }
R2L_HD f2 mul2s(f2 a, float w) { return __fmul2_rn(a, make_float2(w, w)); }
#endif

#ifndef R2L_HOST_EMU
// ---- programmatic dependent launch (PDL): the launcher sets cudaLaunchAttributeProgrammaticStreamSerialization, so this
// kernel's CTAs may become resident while the kernel before it on the stream is still draining.  launch_dependents():
// "the next kernel may start launching"; grid_dependency_wait(): blocks until the kernel(s) before this one have
// completed and their memory is visible -- everything this kernel reads that a predecessor may have written, and
// everything it writes, comes after it.  Both are no-ops when the launch carries no such attribute.
#ifdef R2L_NO_PDL_ASM
__device__ __forceinline__ void pdl_launch_dependents() {}
__device__ __forceinline__ void pdl_wait() {}
#else
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif

// ---- ticket of the fused finish: one 64-bit word {count (low), generation (high)} in the workspace.  A launch owns the
// word once its generation tag is in it: whatever the word held before (an uninitialised workspace, the previous
// launch's final count) is replaced by {gen, 0} by the first CTA to arrive -- no memset in front of the kernel (which
// would also sit between this kernel and the one before it on the stream), no zero-fill contract.  The last CTA clears
// the word, so a captured launch (same tag at every replay) starts clean as well.
// Returns this CTA's ticket 0 .. n_cta-1.  The count itself is ONE atomic add per CTA (a compare-and-swap loop here made
// the 160 CTAs that finish their third tile together retry against each other: +86 us at 64 x 256 x 256,
// profiles/r02_summary.md); only a CTA that finds a foreign tag makes one compare-and-swap attempt to install this
// launch's tag first -- if that fails, another CTA of this launch has installed it.
__device__ __forceinline__ unsigned take_ticket(unsigned* ticket, unsigned gen) {
    unsigned long long* w = reinterpret_cast<unsigned long long*>(ticket);
    volatile unsigned long long* vw = reinterpret_cast<volatile unsigned long long*>(w);
    unsigned long long cur = *vw;
    if ((unsigned)(cur >> 32) != gen) {
        atomicCAS(w, cur, (unsigned long long)gen << 32);
        do { cur = *vw; } while ((unsigned)(cur >> 32) != gen);
    }
    return (unsigned)atomicAdd(w, 1ull);
}
__device__ __forceinline__ void clear_ticket(unsigned* ticket) {
    *reinterpret_cast<volatile unsigned long long*>(ticket) = 0ull;
}

#endif

// two adjacent float2 sites with one 16-byte access (p must be 16-byte aligned: even site index)
R2L_HD void ld2(const f2* p, f2& a, f2& b) {
    const f4 v = *reinterpret_cast<const f4*>(p);
    a = mk2(v.x, v.y); b = mk2(v.z, v.w);
}
R2L_HD void st2(f2* p, f2 a, f2 b) {
    f4 v; v.x = a.x; v.y = a.y; v.z = b.x; v.w = b.y;
    *reinterpret_cast<f4*>(p) = v;
}
// Plane rows are stored "chunk de-interleaved": a row of pitch P sites keeps its even 16-byte chunks (site pairs
// 4q, 4q+1) in the first half and its odd chunks (4q+2, 4q+3) in the second half.  A thread owns the 4-site run
// 4q..4q+3, so lane-consecutive runs read/write lane-consecutive 16-byte chunks in each half: every LDS.128 /
// STS.128 of a warp is bank-conflict free (the natural layout gave 2-way conflicts: 32 B lane stride).
template <int P> R2L_HD int phys(int s) { return ((s >> 1) & 1) * (P / 2) + ((s >> 2) << 1) + (s & 1); }
// sites c-1 .. c+4 of a row, c = 4q;  eb = row*P + 2q
template <int P> R2L_HD void ld6(const f2* pl, int eb, f2 v[6]) {
    const int ob = eb + P / 2;
    v[0] = pl[ob - 1];
    ld2(pl + eb, v[1], v[2]);
    ld2(pl + ob, v[3], v[4]);
    v[5] = pl[eb + 2];
}
// sites c-2 .. c+5
template <int P> R2L_HD void ld8(const f2* pl, int eb, f2 v[8]) {
    const int ob = eb + P / 2;
    ld2(pl + ob - 2, v[0], v[1]); ld2(pl + eb, v[2], v[3]); ld2(pl + ob, v[4], v[5]); ld2(pl + eb + 2, v[6], v[7]);
}
// sites c .. c+3
template <int P> R2L_HD void st4(f2* pl, int eb, f2 a0, f2 a1, f2 a2, f2 a3) {
    st2(pl + eb, a0, a1); st2(pl + eb + P / 2, a2, a3);
}
template <int P> R2L_HD void ld4(const f2* pl, int eb, f2 v[4]) {
    ld2(pl + eb, v[0], v[1]); ld2(pl + eb + P / 2, v[2], v[3]);
}


// ---- raw window -> float2 plane (image A, image B), mirrored on the way (third-generation kernels) ----------------
// The plane has RH rows x P sites, its site (0,0) is image site (ty0 - ROFF, tx0 - COFF).  The source is the TMA
// staging buffer [image][RH][SP] whose column 0 is image column tx0 - COFF - SOFF (TMA = true) or global memory.
// Runs (4 sites) outside the image are not stored; pad rows read the mirrored source row (reflect-1 of the mosaic,
// pipeline_torch.py:233: -1 -> 1, H -> H-2); pad columns -1 / W are written by the first / last run of the image.
// The outermost run on each side (beyond the halo any consumer reads) is skipped.
template <int P, int RH, int ROFF, int COFF, int SP, int SOFF, int NT, typename RawT, bool TMA, int LW = P>
R2L_HD void phase_deinterleave(int tid, f2* __restrict__ XR, const RawT* __restrict__ stage, const RawT* imgA,
                               const RawT* imgB, float denom, int ty0, int tx0, int H, int W, int ly0 = 0, int ly1 = RH) {
    constexpr int Q = LW / 4, QI = Q - 2;                               // LW: sites of a row in use (the pitch P may be padded)
#pragma unroll 2
    for (int i = tid; i < (ly1 - ly0) * QI; i += NT) {                 // plane rows ly0 .. ly1-1
        const int lr = i / QI, lq = i - lr * QI + 1, ly = ly0 + lr;
        const int gy = ty0 - ROFF + ly, gx = tx0 - COFF + 4 * lq;
        if ((unsigned)gx >= (unsigned)W) continue;
        int sy = gy;
        if ((unsigned)gy >= (unsigned)H) sy = mirror_clamped(gy, H);
        float va[4], vb[4];
#ifndef R2L_HOST_EMU
        if (TMA) {
            int sl = ly;
            if (sy != gy) sl = imin(imax(sy - (ty0 - ROFF), 0), RH - 1);
            const RawT* sa = stage + sl * SP + 4 * lq + SOFF;
            const RawT* sb = sa + RH * SP;
            if (sizeof(RawT) == 4) {
                const f4 xa = *reinterpret_cast<const f4*>(sa);
                const f4 xb = *reinterpret_cast<const f4*>(sb);
                va[0] = xa.x; va[1] = xa.y; va[2] = xa.z; va[3] = xa.w;
                vb[0] = xb.x; vb[1] = xb.y; vb[2] = xb.z; vb[3] = xb.w;
            } else {
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    va[j] = __fdiv_rn((float)sa[j], denom);
                    vb[j] = __fdiv_rn((float)sb[j], denom);
                }
            }
        } else
#endif
        {
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                va[j] = RawLoad<RawT>::get(imgA + (size_t)sy * W + gx + j, denom);
                vb[j] = RawLoad<RawT>::get(imgB + (size_t)sy * W + gx + j, denom);
            }
        }
        const int eb = ly * P + 2 * lq, ob = eb + P / 2;
        st2(XR + eb, mk2(va[0], vb[0]), mk2(va[1], vb[1]));
        st2(XR + ob, mk2(va[2], vb[2]), mk2(va[3], vb[3]));
        if (gx == 0) XR[ob - 1] = mk2(va[1], vb[1]);                  // site 4*lq - 1 (pad column -1 = column 1)
        if (gx + 4 == W) XR[eb + 2] = mk2(va[2], vb[2]);              // site 4*lq + 4 (pad column W = column W-2)
    }
}

#ifndef R2L_HOST_EMU
// ---- TMA (cp.async.bulk.tensor) + mbarrier: the raw window of both images is fetched by the copy engine into a
// staging buffer while the previous tile is still being computed ------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "R2L_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@!p bra R2L_WAIT_%=;\n\t}"
        ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// one thread: arm the barrier with the byte count and launch the 3-D box copy {x, y, image}
__device__ __forceinline__ void tma_load_3d(void* dst, const void* tmap, int x, int y, int z, uint64_t* bar, uint32_t bytes) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");     // earlier generic-proxy reads of dst are done
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
        ::"r"(smem_u32(dst)), "l"(tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar)) : "memory");
}
#endif

// parameter tables in the layout the v2 phases read with vector loads
struct Tables2 {
    Tables base;                 // bl, A, AW, Cb, Ws, Wg, M2, invg (filled by build_tables_step)
    float awrow[2][56];          // [row phase][col phase 0: 27 | col phase 1: 27 | pad 2]: AW[2rp+cp][k][t]
    float cbrow[2][8];           // [row phase][cp*3 + k], pad 2
    float awy[2][20];            // luma only: [row phase][cp*9 + t], pad 2
};

R2L_HD void build_tables2_extra(int tid, int nt, Tables2* T) {
    for (int e = tid; e < 2 * 56; e += nt) {
        const int rp = e / 56, r = e - 56 * rp;
        float v = 0.f;
        if (r < 54) { const int cp = r / 27, q = r - 27 * cp; v = T->base.AW[2 * rp + cp][q / 9][q % 9]; }
        T->awrow[rp][r] = v;
    }
    for (int e = tid; e < 2 * 8; e += nt) {
        const int rp = e / 8, r = e - 8 * rp;
        T->cbrow[rp][r] = r < 6 ? T->base.Cb[2 * rp + r / 3][r % 3] : 0.f;
    }
    for (int e = tid; e < 2 * 20; e += nt) {
        const int rp = e / 20, r = e - 20 * rp;
        T->awy[rp][r] = r < 18 ? T->base.AW[2 * rp + r / 9][0][r % 9] : 0.f;
    }
}

template <int TH_, int TW_, int NT_> struct Fwd2Cfg {
    static constexpr int TH = TH_, TW = TW_, NT = NT_;
    static constexpr int P = TW + 16;                 // row pitch (sites) of the haloed planes; column index = gx - x0 + 8
    static constexpr int RH = TH + 8, Y0H = TH + 6, Y1H = TH + 4;
    static constexpr int G = TW / 4;                  // 4-site groups per tile row
    static constexpr int kTableFloats = (sizeof(Tables2) + 15) / 16 * 4;
    static constexpr int kXR = RH * P, kY0 = Y0H * P, kUV = TH * TW;          // sizes in float2 sites
    static constexpr size_t kPlaneBytes = (size_t)kTableFloats * 4 + (size_t)(kXR + kY0 + 2 * kUV) * 8;
    static constexpr size_t kStageOffset = (kPlaneBytes + 127) / 128 * 128;       // TMA destination: 128-byte aligned
    static constexpr size_t kStageBytes = (size_t)2 * RH * P * 4;                 // [image][row][P] of the raw element
    static constexpr size_t kSmemBytes = kPlaneBytes;                             // generic loader
    static constexpr size_t kSmemBytesTma = kStageOffset + kStageBytes + 16;      // + staging + mbarrier
    static_assert(TW % 8 == 0 && TH % 2 == 0 && NT % G == 0 && ((NT / G) % 2) == 0, "row phase must be per-thread");
    static_assert(Y1H * P <= kXR, "Y1 aliases the raw window");
};

// tiles of image PAIRS: pair p = images (2p, 2p+1); an odd last image is paired with itself and lane .y ignored
R2L_HD void decode_pair_tile(const TileGrid& grid, int id, int TH, int TW, int B, int& b0, int& b1, int& ty0, int& tx0) {
    int pair;
    grid.decode(id, TH, TW, pair, ty0, tx0);
    b0 = 2 * pair;
    b1 = b0 + 1 < B ? b0 + 1 : b0;
}

// TMA = true (device only): the raw window arrives through a tensor map (box P x RH x 2, out-of-bounds zero filled)
template <class Cfg, typename RawT, bool STATS, bool TMA = false>
R2L_HD void fwd2_cta(int cta, int n_cta, const FwdArgs& a, const TileGrid& grid, float* smem, const void* tmap = nullptr) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT, P = Cfg::P, G = Cfg::G;
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
    constexpr uint32_t kTmaBytes = 2u * Cfg::RH * P * sizeof(RawT);
    if (TMA) {
        if (threadIdx.x == 0) {
            mbar_init(mbar, 1);
            if (cta < grid.n) {
                int pb0, pb1, py0, px0;
                decode_pair_tile(grid, cta, TH, TW, a.B, pb0, pb1, py0, px0);
                tma_load_3d(stage, tmap, px0 - 8, py0 - 4, pb0, mbar, kTmaBytes);
            }
        }
        __syncthreads();
    }
#endif
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b0, b1, ty0, tx0;
        decode_pair_tile(grid, tile, TH, TW, a.B, b0, b1, ty0, tx0);
        const RawT* imgA = static_cast<const RawT*>(a.raw) + (size_t)b0 * plane;
        const RawT* imgB = static_cast<const RawT*>(a.raw) + (size_t)b1 * plane;
        const bool interior = ty0 >= 4 && tx0 >= 8 && ty0 + TH + 4 <= H && tx0 + TW + 8 <= W;

#ifndef R2L_HOST_EMU
        if (TMA) {
            // ---- P1 (TMA): staged [image][row][P] -> XR site pairs; the next tile's window is requested right after ----
            mbar_wait(mbar, tma_phase);
            tma_phase ^= 1u;
            {
                const int tid = threadIdx.x;
                constexpr int Q = P / 4;
                const RawT* sa = stage;
                const RawT* sb = stage + Cfg::RH * P;
                for (int i = tid; i < Cfg::RH * Q; i += NT) {
                    const int ly = i / Q, lq = i - ly * Q;
                    float va[4], vb[4];
                    if (sizeof(RawT) == 4) {
                        const f4 xa = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(sa) + ly * P + 4 * lq);
                        const f4 xb = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(sb) + ly * P + 4 * lq);
                        va[0] = xa.x; va[1] = xa.y; va[2] = xa.z; va[3] = xa.w;
                        vb[0] = xb.x; vb[1] = xb.y; vb[2] = xb.z; vb[3] = xb.w;
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            va[j] = __fdiv_rn((float)sa[ly * P + 4 * lq + j], a.denom);
                            vb[j] = __fdiv_rn((float)sb[ly * P + 4 * lq + j], a.denom);
                        }
                    }
                    st4<P>(XR, ly * P + 2 * lq, mk2(va[0], vb[0]), mk2(va[1], vb[1]), mk2(va[2], vb[2]), mk2(va[3], vb[3]));
                }
            }
            __syncthreads();
            if (threadIdx.x == 0) {
                const int next = tile + n_cta;
                if (next < grid.n) {
                    int nb0, nb1, ny0, nx0;
                    decode_pair_tile(grid, next, TH, TW, a.B, nb0, nb1, ny0, nx0);
                    tma_load_3d(stage, tmap, nx0 - 8, ny0 - 4, nb0, mbar, kTmaBytes);
                }
            }
            if (!interior) {
                // the copy engine zero-fills outside the image; the demosaic needs the reflected 1-wide ring
                const int tid = threadIdx.x;
                for (int i = tid; i < 2 * P; i += NT) {
                    const int gy = i < P ? -1 : H, lx = i < P ? i : i - P;
                    const int ly = gy - (ty0 - 4), sy = mirror(gy, H) - (ty0 - 4);
                    if (ly >= 0 && ly < Cfg::RH && sy >= 0 && sy < Cfg::RH) XR[ly * P + lx] = XR[sy * P + lx];
                }
                __syncthreads();
                for (int i = tid; i < 2 * Cfg::RH; i += NT) {
                    const int gx = i < Cfg::RH ? -1 : W, ly = i < Cfg::RH ? i : i - Cfg::RH;
                    const int lx = gx - (tx0 - 8), sx = mirror(gx, W) - (tx0 - 8);
                    if (lx >= 0 && lx < P && sx >= 0 && sx < P) XR[ly * P + phys<P>(lx)] = XR[ly * P + phys<P>(sx)];
                }
                __syncthreads();
            }
        } else
#endif
        // ---- P1: raw window of both images -> XR (mirror-clamped indices; 4-site runs when fully inside) ----
        { R2L_FOR_THREADS(NT) {
            constexpr int Q = P / 4;                               // 4-site runs per row
            for (int i = tid; i < Cfg::RH * Q; i += NT) {
                const int ly = i / Q, lq = i - ly * Q;
                const int gy = ty0 - 4 + ly, gx = tx0 - 8 + 4 * lq;
                const int eb = ly * P + 2 * lq;
                if (vec_ok && sizeof(RawT) == 4 && gy >= 0 && gy < H && gx >= 0 && gx + 3 < W) {
                    const f4 va = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgA) + (size_t)gy * W + gx);
                    const f4 vb = *reinterpret_cast<const f4*>(reinterpret_cast<const float*>(imgB) + (size_t)gy * W + gx);
                    st4<P>(XR, eb, mk2(va.x, vb.x), mk2(va.y, vb.y), mk2(va.z, vb.z), mk2(va.w, vb.w));
                } else {
                    const int sy = mirror_clamped(gy, H);
                    f2 v[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const int sx = mirror_clamped(gx + j, W);
                        v[j] = mk2(RawLoad<RawT>::get(imgA + (size_t)sy * W + sx, a.denom),
                                   RawLoad<RawT>::get(imgB + (size_t)sy * W + sx, a.denom));
                    }
                    st4<P>(XR, eb, v[0], v[1], v[2], v[3]);
                }
            }
        } }
        R2L_SYNC();

        // ---- P2: Y0 on the haloed region, U and V on the tile, straight from the raw window -----------------
        { R2L_FOR_THREADS(NT) {
            // interior items: thread meets rows of ONE phase -> 54 weights stay in registers for both items
            const int rp = (tid / G) & 1;
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
            for (int item = tid; item < TH * G; item += NT) {
                const int r = item / G, g = item - r * G;
                const int q = 2 + g;                                  // run index: column index c = 8 + 4g = 4q
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
                st4<P>(Y0, (r + 3) * P + 2 * q, acc[0][0], acc[1][0], acc[2][0], acc[3][0]);
                st4<TW>(U, r * TW + 2 * g, acc[0][1], acc[1][1], acc[2][1], acc[3][1]);
                st4<TW>(V, r * TW + 2 * g, acc[0][2], acc[1][2], acc[2][2], acc[3][2]);
            }
            // halo items (luma only): 3 rows above/below x (G+2) groups, and the two side groups of every tile row
            constexpr int kTopBot = 6 * (G + 2), kSide = 2 * TH;
            for (int item = tid; item < kTopBot + kSide; item += NT) {
                int ry, g;
                if (item < kTopBot) {
                    const int rr = item / (G + 2);
                    g = item - rr * (G + 2) - 1;
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
                st4<P>(Y0, (ry + 3) * P + 2 * q, acc[0], acc[1], acc[2], acc[3]);
            }
        } }
        R2L_SYNC();
        if (!interior) {
            // sharpen zero-pads: Y0 must vanish on the 1-wide ring just outside the image
            { R2L_FOR_THREADS(NT) {
                for (int i = tid; i < 2 * P + 2 * Cfg::Y0H; i += NT) {
                    int gy, gx;
                    if (i < P) { gy = -1; gx = tx0 - 8 + i; }
                    else if (i < 2 * P) { gy = H; gx = tx0 - 8 + (i - P); }
                    else if (i < 2 * P + Cfg::Y0H) { gy = ty0 - 3 + (i - 2 * P); gx = -1; }
                    else { gy = ty0 - 3 + (i - 2 * P - Cfg::Y0H); gx = W; }
                    const int ly = gy - (ty0 - 3), lx = gx - (tx0 - 8);
                    if (ly >= 0 && ly < Cfg::Y0H && lx >= 0 && lx < P) Y0[ly * P + phys<P>(lx)] = mk2(0.f, 0.f);
                }
            } }
            R2L_SYNC();
        }

        // ---- P3: Y1 = sharpen(Y0) on rows -2..TH+1, groups -1..G (1x4 runs); overwrites the raw window ---------
        { R2L_FOR_THREADS(NT) {
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            for (int item = tid; item < Cfg::Y1H * (G + 2); item += NT) {
                const int rr = item / (G + 2), g = item - rr * (G + 2) - 1;
                const int q = 2 + g;
                f2 acc[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
#pragma unroll
                for (int aa = 0; aa < 3; ++aa) {
                    f2 in[6];
                    ld6<P>(Y0, (rr + aa) * P + 2 * q, in);        // Y1 row rr = image row ty0-2+rr; Y0 row of (that-1) is rr
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) acc[j] = fma2s(in[j + bb], ws[aa * 3 + bb], acc[j]);
                }
                st4<P>(Y1, rr * P + 2 * q, acc[0], acc[1], acc[2], acc[3]);
            }
        } }
        R2L_SYNC();
        if (!interior) {
            // Gaussian reflect-pads the sharpened plane: rows first, then columns (corners come out right)
            { R2L_FOR_THREADS(NT) {
                for (int i = tid; i < 4 * P; i += NT) {
                    const int q = i / P, lx = i - q * P;
                    const int gy = q == 0 ? -2 : (q == 1 ? -1 : (q == 2 ? H : H + 1));
                    const int ly = gy - (ty0 - 2), sy = mirror(gy, H) - (ty0 - 2);
                    if (ly >= 0 && ly < Cfg::Y1H && sy >= 0 && sy < Cfg::Y1H) Y1[ly * P + lx] = Y1[sy * P + lx];   // same physical column
                }
            } }
            R2L_SYNC();
            { R2L_FOR_THREADS(NT) {
                for (int i = tid; i < 4 * Cfg::Y1H; i += NT) {
                    const int q = i / Cfg::Y1H, ly = i - q * Cfg::Y1H;
                    const int gx = q == 0 ? -2 : (q == 1 ? -1 : (q == 2 ? W : W + 1));
                    const int lx = gx - (tx0 - 8), sx = mirror(gx, W) - (tx0 - 8);
                    if (lx >= 0 && lx < P && sx >= 0 && sx < P) Y1[ly * P + phys<P>(lx)] = Y1[ly * P + phys<P>(sx)];
                }
            } }
            R2L_SYNC();
        }

        // ---- P4: Gaussian (2x4 runs) + YUV->RGB + clip + gamma [+ additive] [+ affine] -> global ----------------
        { R2L_FOR_THREADS(NT) {
            float wg[25], m2[9];
#pragma unroll
            for (int t = 0; t < 25; ++t) wg[t] = T->Wg[t];
#pragma unroll
            for (int t = 0; t < 9; ++t) m2[t] = T->M2[t];
            const float invg = T->invg;
            for (int item = tid; item < (TH / 2) * G; item += NT) {
                const int r0 = 2 * (item / G), g = item % G;
                const int q = 2 + g;
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
#pragma unroll
                for (int o = 0; o < 2; ++o) {
                    const int gy = ty0 + r0 + o, gx = tx0 + 4 * g;
                    if (gy >= H || gx >= W) continue;
                    f2 u[4], v[4];
                    ld4<TW>(U, (r0 + o) * TW + 2 * g, u);
                    ld4<TW>(V, (r0 + o) * TW + 2 * g, v);
                    float oa[3][4], ob[3][4];
#pragma unroll
                    for (int j = 0; j < 4; ++j)
#pragma unroll
                        for (int k = 0; k < 3; ++k) {
                            const f2 rr = fma2s(v[j], m2[k * 3 + 2], fma2s(u[j], m2[k * 3 + 1], mul2s(acc[o][j], m2[k * 3])));
                            const float ca = fminf(fmaxf(rr.x, kClipLo), kClipHi);
                            const float cbv = fminf(fmaxf(rr.y, kClipLo), kClipHi);
                            oa[k][j] = fast_exp2(invg * fast_log2(ca));
                            ob[k][j] = fast_exp2(invg * fast_log2(cbv));
                        }
                    const size_t pix = (size_t)gy * W + gx;
                    const bool full = vec_ok && gx + 3 < W;
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        if (a.additive) {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (gx + j < W) {
                                    const float ad = a.additive[(size_t)k * plane + pix + j];
                                    oa[k][j] += ad; ob[k][j] += ad;
                                }
                        }
                        if (STATS) {
                            ChanAcc& cs = R2L_ACC(cacc, tid);
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (gx + j < W) {
                                    cs.s[k] += oa[k][j];
                                    cs.s[3 + k] = fmaf_(oa[k][j], oa[k][j], cs.s[3 + k]);
                                    if (b1 != b0) {
                                        cs.s[k] += ob[k][j];
                                        cs.s[3 + k] = fmaf_(ob[k][j], ob[k][j], cs.s[3 + k]);
                                    }
                                }
                        }
                        if (a.affine) {
                            const float sc = a.affine[k], sh = a.affine[3 + k];
#pragma unroll
                            for (int j = 0; j < 4; ++j) { oa[k][j] = fmaf_(oa[k][j], sc, sh); ob[k][j] = fmaf_(ob[k][j], sc, sh); }
                        }
                        float* pa = a.out + ((size_t)b0 * 3 + k) * plane + pix;
                        float* pb = a.out + ((size_t)b1 * 3 + k) * plane + pix;
                        if (full) {
                            f4 va; va.x = oa[k][0]; va.y = oa[k][1]; va.z = oa[k][2]; va.w = oa[k][3];
                            *reinterpret_cast<f4*>(pa) = va;
                            if (b1 != b0) {
                                f4 vb; vb.x = ob[k][0]; vb.y = ob[k][1]; vb.z = ob[k][2]; vb.w = ob[k][3];
                                *reinterpret_cast<f4*>(pb) = vb;
                            }
                        } else {
#pragma unroll
                            for (int j = 0; j < 4; ++j)
                                if (gx + j < W) {
                                    pa[j] = oa[k][j];
                                    if (b1 != b0) pb[j] = ob[k][j];
                                }
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
#endif
    }
}

}  // namespace r2l
