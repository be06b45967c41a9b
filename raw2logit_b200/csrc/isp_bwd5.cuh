// isp_bwd5.cuh -- fifth-generation fused backward: the statistics live in tensor memory.
//
// Reference: the autograd graph of pipeline_torch.py:183-217.  Same phases, planes, border rules and
// adjoint algebra as the fourth generation (isp_bwd4.cuh: nothing recomputed, forward output + saved luma planes), but
// the 96 per-thread running sums behind the 132 parameter gradients no longer occupy registers for the whole launch:
//   * they are parked in TMEM (isp_tmem.cuh: one private 32-bit cell per thread and column) and a phase loads only the
//     group it updates -- B4 the gamma sum, B5 five sums per Gaussian tap row, B6 the 9 sharpening taps, B7 six sums per
//     gradient plane and tap row -- and stores it back when the pass ends; the load is requested when the pass starts
//     and awaited at its end, where the sums of the tile are added;
//   * with at most a dozen packed sums live, the kernel fits 128 registers: 256 threads x 2 CTAs per SM = 16 warps (was
//     8 at 255 registers with spills, profiles/r01_v4_summary.md), so the global loads of one warp hide behind the
//     stencils of three others;
//   * every statistic is accumulated in packed (image A, image B) registers: one FFMA2 per tap and site pair (the Q'
//     statistic took two scalar FMAs); B7 walks the three gradient planes one after the other (k outermost);
//   * the hot loops are straight-line code: pad columns are realised as per-site weights and masked centres, pad rows
//     in a cold pass behind a warp vote, plane / tap-row loops are runtime loops (the kernel fits the instruction cache);
//   * the centre values a phase reads from global memory are requested ahead of the CTA barrier before it;
//   * stencil weights are read from shared memory next to their use (uniform-address LDS) instead of sitting in
//     25 / 54 registers for a whole phase;
//   * no software L2 prefetch: prefetch.global.L2 (CCTL.E.PF2) fetches one 32-byte sector per instruction, and both the
//     fourth generation's one-per-line pattern and a one-per-sector pattern made this kernel SLOWER (99 / 115 us against
//     94 us without, profiles/r01_v5_summary.md) -- with 16 warps per SM the demand loads are covered well enough.
// Column layout per thread (also the order of the host emulation's array and of the CTA reduction):
//   [0,2) gamma sum per stream | [2,27) dWg | [27,36) dWs | [36 + 20 k, +18) Q'[k][tap row][col phase][b] | (+18, +2) P[k][col phase]
#pragma once
#include "isp_bwd4.cuh"
#include "isp_tmem.cuh"

namespace r2l {

constexpr int kB5Sg = 0, kB5Wg = 2, kB5Ws = 27, kB5Q = 36, kB5QStride = 20, kBwd5AccFloats = 96;

template <int TH_, int TW_, int NT_, bool GRAW_, bool TAIL_> struct Bwd5Cfg : Bwd4Cfg<TH_, TW_, NT_, GRAW_, TAIL_> {
    static constexpr int kAccCols = (NT_ / 128) * kBwd5AccFloats;          // warps w and w + 4 share TMEM lanes
    static constexpr int kTmemCols = kAccCols <= 32 ? 32 : kAccCols <= 64 ? 64 : kAccCols <= 128 ? 128 : kAccCols <= 256 ? 256 : 512;
    static_assert(NT_ % 128 == 0 && kAccCols <= 512, "whole warpgroups; the sums must fit the CTA's TMEM columns");
    static_assert(((TH_ / 2) * (TW_ / 4)) % (NT_ / 2) == 0, "B7: the same number of items for every thread");
};

inline bool bwd5_shape_ok(int H, int W) { return bwd4_shape_ok(H, W); }

#ifndef R2L_HOST_EMU
// Tensor maps of the tensors the backward streams, used ONLY to prefetch the next tile's boxes into the L2
// (cp.async.bulk.prefetch.tensor: one instruction of one thread per box -- the per-thread prefetch.global.L2 variants cost
// more issue slots than they saved).  The phases then find their demand loads in the L2 instead of DRAM.
struct alignas(64) Bwd5Maps {
    CUtensorMap gout, out;      // (W, H, 3, B) fp32, box (TW + 8, TH + 8, 3, 2)
    CUtensorMap luma;           // (2 W, H, pairs, 2) fp32, box (2 TW, TH, 1, 2)
    CUtensorMap raw;            // (W, H, B), box (TW, TH, 2)
    int on;                     // 0: no prefetch; bits 0-3 = point of the gout / out prefetch, bits 4-7 = of luma / raw
};
__device__ __forceinline__ void tma_prefetch_4d(const void* tmap, int c0, int c1, int c2, int c3) {
    asm volatile("cp.async.bulk.prefetch.tensor.4d.L2.global.tile [%0, {%1, %2, %3, %4}];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2), "r"(c3) : "memory");
}
__device__ __forceinline__ void tma_prefetch_3d(const void* tmap, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];"
                 ::"l"(tmap), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
#else
struct Bwd5Maps { int on; };
#endif

// bit pattern of a float
R2L_HD unsigned fbits(float x) {
#ifdef R2L_HOST_EMU
    unsigned u; std::memcpy(&u, &x, 4); return u;
#else
    return __float_as_uint(x);
#endif
}

// (x, y) -> one aligned register pair that stays one: a plain make_float2 of values from two different loads is
// re-packed with two MOVs at every FFMA2 that uses it when registers are tight (profiles/r01_v5_summary.md), and
// ptxas folds a mov.b64 pack / unpack away; a packed multiply by a run-time 1.0 produces the pair as an FMUL2 result.
R2L_HD f2 pack2(float x, float y, float one) { return mul2s(mk2(x, y), one); }

// a phase's view of its running sums: load at phase start, store at phase end
#ifdef R2L_HOST_EMU
#define R2L_ANY(x) (x)
#define R2L_PARK_LOAD(N, col, v)  { const float* s_ = accs[tid].sums + (col); for (int i_ = 0; i_ < (N); ++i_) (v)[i_] = s_[i_]; }
#define R2L_PARK_READY(N, v)
#define R2L_PARK_STORE(N, col, v) { float* s_ = accs[tid].sums + (col); for (int i_ = 0; i_ < (N); ++i_) s_[i_] = (v)[i_]; }
struct Bwd5Acc { float sums[kBwd5AccFloats]; };
#else
#define R2L_ANY(x) (__any_sync(0xffffffffu, (x)) != 0)
// LOAD only requests (after the thread's earlier stores have landed); the values are needed when the phase ends, so the
// TMEM round trip runs behind the phase's arithmetic and READY (wait + register tie) sits in front of their first use
#define R2L_PARK_LOAD(N, col, v)  { __syncwarp(); tmem::wait_st(); tmem::load<N>(tacc + (col), v); }
#define R2L_PARK_READY(N, v)      { tmem::ready<N>(v); }
#define R2L_PARK_STORE(N, col, v) { __syncwarp(); tmem::store<N>(tacc + (col), v); }
#endif

#ifndef R2L_HOST_EMU
// ---- one-shot all-reduce of the 132 gradients over NVLink peer memory, run by the CTA that finished them ------------
// (r2l_isp_backward_dp, include/r2l_isp.h).  Every gradient travels as one 8-byte word {value, epoch} written by ONE
// 64-bit store (single-copy atomic), so the epoch tag tells the reader that the value next to it is this step's -- no
// fence, no separate flag, one NVLink one-way latency.  Push: thread e writes word e of this rank's slot on EVERY rank.
// Pull: thread e polls word e of every rank's slot in OUR buffer until it carries the epoch and adds the values in
// rank order -- the same order everywhere, so all ranks hold bit-identical sums.  Slots are double-buffered by epoch
// parity: a rank that already pushes step i + 1 cannot overwrite what a slower rank still reads for step i, and it
// cannot reach step i + 2 before that rank has pushed step i + 1.
__device__ __forceinline__ void st_tagged_sys(float* p, float v, unsigned tag) {
    // ONE 64-bit store: {value (low word), epoch tag (high word)} -- single-copy atomic, unlike a .v2.b32 access, which
    // the PTX memory model treats as two scalar accesses in unspecified order
    const unsigned long long w = ((unsigned long long)tag << 32) | (unsigned long long)__float_as_uint(v);
    asm volatile("st.relaxed.sys.global.b64 [%0], %1;" ::"l"(p), "l"(w) : "memory");
}
__device__ __forceinline__ unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
// polls until the word carries `tag`; gives up at `deadline` (a peer that crashed or skipped its backward must not hang
// the GPU): then *ok is cleared and the caller poisons the gradients with NaN, which the host sees in the result
__device__ __forceinline__ float ld_tagged_sys(const float* p, unsigned tag, unsigned long long deadline, bool* ok) {
    unsigned long long w;
    unsigned spins = 0;
    for (;;) {
        asm volatile("ld.relaxed.sys.global.b64 %0, [%1];" : "=l"(w) : "l"(p) : "memory");
        if ((unsigned)(w >> 32) == tag) break;
        if ((++spins & 1023u) == 0u && global_timer_ns() > deadline) { *ok = false; break; }
    }
    return __uint_as_float((unsigned)w);
}
constexpr unsigned long long kExchangeTimeoutNs = 5000000000ull;        // 5 s
template <int NT> __device__ __forceinline__ void peer_allreduce(const BwdArgs& a) {
    static_assert(kSlotPitch >= R2L_NUM_PARAM_GRADS, "one tagged word per gradient");
    const int world = a.world, e0 = threadIdx.x;
    unsigned epoch = a.epoch;
    if (epoch == R2L_EPOCH_DEVICE) {
        // the count lives in the word behind this rank's slots (local memory; only this CTA of this launch touches it,
        // and launches on one stream run in order): nothing per call comes from the host, so a captured launch replays
        __shared__ unsigned s_epoch;
        if (e0 == 0) {
            unsigned* c = reinterpret_cast<unsigned*>(a.peers[a.rank] + (size_t)2 * world * (2 * kSlotPitch));
            unsigned v = *c + 1u;
            if (v == R2L_EPOCH_DEVICE || v == 0u) v = 1u;
            *c = v;
            s_epoch = v;
        }
        __syncthreads();
        epoch = s_epoch;
    }
    const size_t slot0 = (size_t)(epoch & 1u) * world * (2 * kSlotPitch);        // floats; a slot = kSlotPitch words of 8 bytes
    __syncthreads();                                                     // a.grads of this launch are written
    for (int e = e0; e < R2L_NUM_PARAM_GRADS; e += NT) {            // (one pass when the CTA has >= 132 threads)
        // Thread e % NT is the only one that touches a.grads[e] from here on: it reads the local value once, pushes it to
        // every rank, then pulls every rank's word e and writes the sum back.  (A version whose push loop ran over
        // (rank, gradient) pairs let another thread push a.grads[e] after thread e had already overwritten it with
        // the reduced sum -- ADVICE round 1.)
        const float local = a.grads[e];
        for (int p = 0; p < world; ++p)
            st_tagged_sys(a.peers[p] + slot0 + (size_t)a.rank * (2 * kSlotPitch) + 2 * e, local, epoch);
        const float* mine = a.peers[a.rank] + slot0;
        const unsigned long long deadline = global_timer_ns() + kExchangeTimeoutNs;
        bool ok = true;
        float sum = 0.f;
        for (int r = 0; r < world; ++r) sum += ld_tagged_sys(mine + (size_t)r * (2 * kSlotPitch) + 2 * e, epoch, deadline, &ok);
        a.grads[e] = ok ? sum * a.dp_scale : __int_as_float(0x7fc00000);
    }
}
#endif

template <class Cfg, typename RawT>
R2L_HD void bwd5_cta(int cta, int n_cta, const BwdArgs& a, const TileGrid& grid, float* smem, const Bwd5Maps* maps = nullptr) {
    constexpr int TH = Cfg::TH, TW = Cfg::TW, NT = Cfg::NT, PN = Cfg::PN, G = Cfg::G, HALF = Cfg::HALF;
    constexpr int GG = G + 2;                                     // runs -1 .. G
    Tables2* T2 = reinterpret_cast<Tables2*>(smem);
    Tables* T = &T2->base;
    f2* PU = reinterpret_cast<f2*>(smem + Cfg::kTableFloats);     // gU
    f2* PV = PU + Cfg::kF;                                        // gV
    f2* PG = PV + Cfg::kF;                                        // gY2, then gY0
    f2* GY1 = PG + Cfg::kF;                                       // gY1
#ifdef R2L_HOST_EMU
    std::vector<Bwd5Acc> accs(NT);
    std::memset(accs.data(), 0, sizeof(Bwd5Acc) * NT);
#else
    pdl_launch_dependents();                                      // the next kernel's CTAs may take the slots we free
    // TMEM columns for the running sums: warp 0 allocates, everybody zeroes its own cells
    __shared__ uint32_t tmem_slot;
    if (threadIdx.x < 32) tmem::alloc<Cfg::kTmemCols>(&tmem_slot);
    tmem::fence_before_sync();
    __syncthreads();
    tmem::fence_after_sync();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t tacc = tmem::addr(tmem_base, (int)(threadIdx.x >> 7) * kBwd5AccFloats);
    {
        float z[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) z[i] = 0.f;
#pragma unroll
        for (int c0 = 0; c0 < kBwd5AccFloats; c0 += 16) tmem::store<16>(tacc + c0, z);
        tmem::wait_st();
    }
#endif
    const int H = a.H, W = a.W;
    const size_t plane = (size_t)H * W;
    const size_t luma_plane = (size_t)((a.B + 1) >> 1) * plane * 2;      // floats per saved plane
    // The planes start finite where no phase ever writes: the outermost run on either side of every row (runs -2 and
    // G + 1: B4 / B5 / B6 store runs -1 .. G of every row of their plane, in or outside the image), which don't-care
    // items read.  Eight sites per row instead of the whole 100 KB (the full clear was 2 % of the kernel's samples).
#if defined(R2L_POISON_SMEM) && !defined(R2L_HOST_EMU)
    // test build: everything the clear below leaves alone starts as NaN -- a valid result that depends on it shows
    for (int i = threadIdx.x; i < Cfg::kSites; i += NT) PU[i] = mk2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
    __syncthreads();
#endif
    { R2L_FOR_THREADS(NT) {
        constexpr int kRows = Cfg::kSites / PN;
        static_assert(Cfg::kSites % PN == 0, "whole rows");
        for (int i = tid; i < kRows * 8; i += NT) {
            const int row = i >> 3, j = i & 7;
            PU[row * PN + (j & 1) + ((j >> 1) & 1) * (PN / 2 - 2) + (j >> 2) * (PN / 2)] = mk2(0.f, 0.f);
        }
    } }
    R2L_BUILD_TABLES(NT, a.P, T)
    { R2L_FOR_THREADS(NT) { build_tables2_extra(tid, NT, T2); } }
    R2L_SYNC();
#ifndef R2L_HOST_EMU
    // Under a dependent launch (R2L_ISP_BWD_PDL=1; off by default -- it measured slower, profiles/r02_summary.md) everything
    // up to here may run while the kernel before this one drains: it touched only the parameters (never written by this
    // library's kernels), shared and tensor memory.  The forward output, the luma planes and grad_out are read -- and the
    // workspace, grad_raw, the gradients written -- after the wait.  A no-op for a plain launch.
    pdl_wait();
    // The BatchNorm / additive tail, once per CTA: {gs, c1, c2, 1/ysc, -ysh/ysc} per channel in shared memory.  A deferred
    // tail (kTailDeferredTag in c1) is finished here from the statistics kernel's per-CTA sums -- one warp per channel,
    // lanes stride over the rows, fp64, xor tree: the arithmetic and order of bn_backward_finish_kernel, the same in
    // every CTA -- instead of by a one-CTA launch between the two kernels (5 us per step).
    __shared__ float s_tail[15];
    if (Cfg::TAIL) {
        const int c = (int)(threadIdx.x >> 5), lane = (int)(threadIdx.x & 31);
        if (c < 3) {
            float c1 = a.gtail[3 + c], c2 = a.gtail[6 + c];
            if (a.bn_partials && fbits(a.gtail[3]) == kTailDeferredTag) {
                double s1 = 0.0, s2 = 0.0;
                for (int i = lane; i < kBnBwdBlocks; i += 32) {
                    s1 += (double)__ldcg(a.bn_partials + ((size_t)c * kBnBwdBlocks + i) * 2);
                    s2 += (double)__ldcg(a.bn_partials + ((size_t)c * kBnBwdBlocks + i) * 2 + 1);
                }
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1) {
                    s1 += __shfl_xor_sync(0xffffffffu, s1, o);
                    s2 += __shfl_xor_sync(0xffffffffu, s2, o);
                }
                c1 = (float)(s1 / a.bn_count);
                c2 = (float)(s2 / a.bn_count);
            }
            if (lane == 0) {
                const float isc = 1.0f / a.gtail[9 + c];
                s_tail[c] = a.gtail[c]; s_tail[3 + c] = c1; s_tail[6 + c] = c2;
                s_tail[9 + c] = isc; s_tail[12 + c] = -a.gtail[12 + c] * isc;
            }
        }
        __syncthreads();
    }
    const float* tail = s_tail;
#else
    float tail_emu[15] = {0.f};
    if (Cfg::TAIL) {
        for (int c = 0; c < 3; ++c) {
            const float isc = 1.0f / a.gtail[9 + c];
            tail_emu[c] = a.gtail[c]; tail_emu[3 + c] = a.gtail[3 + c]; tail_emu[6 + c] = a.gtail[6 + c];
            tail_emu[9 + c] = isc; tail_emu[12 + c] = -a.gtail[12 + c] * isc;
        }
    }
    const float* tail = tail_emu;
#endif
    for (int tile = cta; tile < grid.n; tile += n_cta) {
        int b0, b1, ty0, tx0;
        decode_pair_tile(grid, tile, TH, TW, a.B, b0, b1, ty0, tx0);
        const bool dup = b1 == b0;
#ifndef R2L_HOST_EMU
        // the next tile's boxes -> L2 (one thread); point 1: tile start, 2: B6 start, 3: B7 start
        auto prefetch_next = [&](int point) {
            if (!maps || !maps->on || threadIdx.x != 0 || tile + n_cta >= grid.n) return;
            const bool big = (maps->on & 15) == point, small = ((maps->on >> 4) & 15) == point;
            if (!big && !small) return;
            int nb0, nb1, nty0, ntx0;
            decode_pair_tile(grid, tile + n_cta, TH, TW, a.B, nb0, nb1, nty0, ntx0);
            if (big) {
                tma_prefetch_4d(&maps->gout, ntx0 - 4, nty0 - 4, 0, nb0);
                tma_prefetch_4d(&maps->out, ntx0 - 4, nty0 - 4, 0, nb0);
            }
            if (small) {
                tma_prefetch_4d(&maps->luma, 2 * ntx0, nty0, nb0 >> 1, 0);
                tma_prefetch_3d(&maps->raw, ntx0, nty0, nb0);
            }
        };
        prefetch_next(1);
#endif
        const RawT* imgA = static_cast<const RawT*>(a.raw) + (size_t)b0 * plane;
        const RawT* imgB = static_cast<const RawT*>(a.raw) + (size_t)b1 * plane;
        const float* y0pair = a.luma + (size_t)(b0 >> 1) * plane * 2;   // Y0 of this image pair, [H][W][2]
        const float* y1pair = y0pair + luma_plane;

        // The centre values a phase reads from global memory (Y1 in B5, Y0 in B6, raw in B7) are requested at the END of
        // the phase before, ahead of the CTA barrier: the L2 round trip runs while the warp waits for the others
        // (profiles/r01_v5_summary.md: barrier 1.4 and long-scoreboard 2.1 stall cycles per issued instruction).  The
        // host emulation, which runs a phase for all threads before the next, loads them at the start of the phase.
        constexpr int NI5 = TH * G / NT, NI6 = TH * (G / 2) / NT, NI7 = (TH / 2) * G / HALF;
        static_assert((TH * G) % NT == 0 && (TH * (G / 2)) % NT == 0, "the same number of owned items for every thread");
        auto load_c5 = [&](int tid, f2 (&c)[NI5][4]) {                  // Y1 centres of B5's owned items
#pragma unroll
            for (int u = 0; u < NI5; ++u) {
                const int item = tid + u * NT;
                const int r = item / G, g = item - r * G;
                const int qy = ty0 + r, qx = tx0 + 4 * g;
#pragma unroll
                for (int j = 0; j < 4; ++j) c[u][j] = mk2(0.f, 0.f);
                if (qy < H && qx < W) ld_luma4(y1pair + ((size_t)qy * W + qx) * 2, c[u]);
            }
        };
        auto load_c6 = [&](int tid, f2 (&c)[NI6][2][4]) {               // Y0 centres of B6's owned (paired) items
#pragma unroll
            for (int v = 0; v < NI6; ++v) {
                const int item = tid + v * NT;
                const int r = item / (G / 2), g = item - r * (G / 2);
                const int qy = ty0 + r;
#pragma unroll
                for (int u = 0; u < 2; ++u) {
                    const int qx = tx0 + 4 * (g + u * (G / 2));
#pragma unroll
                    for (int j = 0; j < 4; ++j) c[v][u][j] = mk2(0.f, 0.f);
                    if (qy < H && qx < W) ld_luma4(y0pair + ((size_t)qy * W + qx) * 2, c[v][u]);
                }
            }
        };
        auto load_x7 = [&](int tid, f4 (&xa)[NI7], f4 (&xb)[NI7]) {     // fp32 raw centres of B7's items (packed there)
            const int rp = (tid >> 5) & 1, slot = ((tid >> 6) << 5) | (tid & 31);
#pragma unroll
            for (int u = 0; u < NI7; ++u) {
                const int i = slot + u * HALF;
                const int ri = i / G, g = i - ri * G;
                const int qy = ty0 + rp + 2 * ri, qx = tx0 + 4 * g;
                xa[u].x = xa[u].y = xa[u].z = xa[u].w = 0.f;
                xb[u] = xa[u];
                if (sizeof(RawT) == 4 && qy < H && qx < W) {
                    xa[u] = ld_stream4(reinterpret_cast<const float*>(imgA) + (size_t)qy * W + qx);
                    xb[u] = ld_stream4(reinterpret_cast<const float*>(imgB) + (size_t)qy * W + qx);
                }
            }
        };
#ifndef R2L_HOST_EMU
        f2 c5[NI5][4], c6[NI6][2][4];
        f4 xa7[NI7], xb7[NI7];
#endif

        // ---- B4: grad_out pulled back through gamma / clip / YUV->RGB to (gY2, gU, gV) on rows -4..TH+3, runs -1..G;
        // gamma statistic.  o = y (or (y - shift)/scale - additive behind a tail); with lo = log2(o):
        // e = cl^(1/g - 1) = 2^((1 - g) lo), log2(cl) = g lo, and the clamp passed iff o lies strictly between its two
        // clipped values (exact compare without a tail, where o is bit-identical to the forward's; a 1e-4 / 1e-6
        // relative margin behind a tail, where o is recovered by an affine inverse).  The margin at the low clip is
        // 5.3e-7 absolute; the inverse's error is ~1e-7 for a BatchNorm tail (|shift / scale| <= 1: the fma is exact
        // before its rounding, what is left are the roundings of 1/scale, shift/scale and of the stored y) and
        // ~ulp(additive)/2 more with an additive layer -- below the margin while |additive| < 2.  Testing the stored y
        // for equality with fma(o_clip + additive, scale, shift) instead would be exact for this library's own forward
        // but fragile for any other producer of `out` (tried and dropped: the host emulation's tail cases feed an fp64
        // oracle's output).
        { R2L_FOR_THREADS(NT) {
            float m2g[9];
            const float invg = T->invg, gam = T->gamma, one_m_g = 1.0f - T->gamma;
#pragma unroll
            for (int t = 0; t < 9; ++t) m2g[t] = T->M2[t] * invg;
            const float o_lo_exact = fast_exp2(invg * fast_log2(kClipLo));      // the forward's value of a low clip
            const float o_lo = Cfg::TAIL ? o_lo_exact * (1.0f + 1e-4f) : o_lo_exact;
            const float o_hi = Cfg::TAIL ? 1.0f - 1e-6f : 1.0f;
            // o_lo < o < o_hi as ONE unsigned compare on the bit patterns (positive floats order like their bits; a
            // negative or NaN o wraps to a huge difference and fails, as it fails the two float compares)
            const unsigned m_base = fbits(o_lo) + 1u, m_span = fbits(o_hi) - m_base;
            // per-tile bases: channel k sits k planes further, image B 3 planes further (32-bit element offsets inside a pair)
            const float* goT = a.gout + (size_t)b0 * 3 * plane;
            const float* yoT = a.out + (size_t)b0 * 3 * plane;
            const int plane_i = (int)plane, dB = (b1 - b0) * 3 * plane_i;      // image B = image A + dB elements (0: duplicate)
            f2 sg = mk2(0.f, 0.f);                                     // sum G o log2(o), per stream (x gamma when parked)
            // owned rectangle first, halo ring after: whole warps are inside or outside the rectangle that carries
            // the gamma statistic
            for (int item = tid; item < Cfg::FH * GG; item += NT) {
                int r, g;
                region_item<TH, G, 4>(item, r, g);
                const int gy = ty0 + r, gx = tx0 + 4 * g;
                const bool valid = gy >= 0 && gy < H && gx >= 0 && gx < W;
                const bool owned = item < TH * G;
                f2 gy2[4], gu[4], gv[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { gy2[j] = mk2(0.f, 0.f); gu[j] = mk2(0.f, 0.f); gv[j] = mk2(0.f, 0.f); }
                if (valid) {
                    const int pix = gy * W + gx;
                    f4 ga[3], gb[3], ya[3], yb[3], ad[3];
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const int off = k * plane_i + pix;
                        ga[k] = ld_stream4(goT + off);
                        ya[k] = ld_stream4(yoT + off);
                        gb[k] = ld_stream4(goT + (off + dB));
                        yb[k] = ld_stream4(yoT + (off + dB));
                        ad[k].x = ad[k].y = ad[k].z = ad[k].w = 0.f;
                        if (Cfg::TAIL && a.additive) ad[k] = *reinterpret_cast<const f4*>(a.additive + off);
                    }
#pragma unroll
                    for (int k = 0; k < 3; ++k) {
                        const float gak[4] = {ga[k].x, ga[k].y, ga[k].z, ga[k].w}, gbk[4] = {gb[k].x, gb[k].y, gb[k].z, gb[k].w};
                        const float yak[4] = {ya[k].x, ya[k].y, ya[k].z, ya[k].w}, ybk[4] = {yb[k].x, yb[k].y, yb[k].z, yb[k].w};
                        const float adk[4] = {ad[k].x, ad[k].y, ad[k].z, ad[k].w};
                        float t_gs = 1.f, t_c1 = 0.f, t_c2 = 0.f, t_isc = 1.f, t_osh = 0.f;
                        if (Cfg::TAIL) {
                            t_gs = tail[k]; t_c1 = tail[3 + k]; t_c2 = tail[6 + k]; t_isc = tail[9 + k]; t_osh = tail[12 + k];
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            float Ga = gak[j], Gb = gbk[j];             // (an odd batch's duplicate stream is zeroed below)
                            f2 o = mk2(yak[j], ybk[j]);
                            if (Cfg::TAIL) {
                                Ga = t_gs * (Ga - t_c1 - t_c2 * o.x);
                                Gb = t_gs * (Gb - t_c1 - t_c2 * o.y);
                                o = mk2(fmaf_(o.x, t_isc, t_osh) - adk[j], fmaf_(o.y, t_isc, t_osh) - adk[j]);
                            }
                            const f2 lo = mk2(fast_log2(o.x), fast_log2(o.y));
                            const f2 ex = mul2s(lo, one_m_g);
                            const f2 e = mk2(fast_exp2(ex.x), fast_exp2(ex.y));
                            if (owned) sg = fma2vv(mk2(Ga * o.x, Gb * o.y), lo, sg);
                            const f2 gr = mk2((fbits(o.x) - m_base < m_span) ? Ga * e.x : 0.f,
                                              (fbits(o.y) - m_base < m_span) ? Gb * e.y : 0.f);
                            gy2[j] = fma2s(gr, m2g[k * 3 + 0], gy2[j]);
                            gu[j] = fma2s(gr, m2g[k * 3 + 1], gu[j]);
                            gv[j] = fma2s(gr, m2g[k * 3 + 2], gv[j]);
                        }
                    }
                    if (dup) {                                          // odd batch, last pair: stream B carries no gradient
#pragma unroll
                        for (int j = 0; j < 4; ++j) { gy2[j].y = 0.f; gu[j].y = 0.f; gv[j].y = 0.f; }
                    }
                }
                st4<PN>(PG, (r + 4) * PN + 2 * (g + 2), gy2[0], gy2[1], gy2[2], gy2[3]);
                st4<PN>(PU, (r + 4) * PN + 2 * (g + 2), gu[0], gu[1], gu[2], gu[3]);
                st4<PN>(PV, (r + 4) * PN + 2 * (g + 2), gv[0], gv[1], gv[2], gv[3]);
            }
            {
                float v[2];
                R2L_PARK_LOAD(2, kB5Sg, v)
                R2L_PARK_READY(2, v)
                v[0] = fmaf_(sg.x, gam, v[0]);
                if (!dup) v[1] = fmaf_(sg.y, gam, v[1]);
                R2L_PARK_STORE(2, kB5Sg, v)
            }
#ifndef R2L_HOST_EMU
            load_c5(tid, c5);
#endif
        } }
        R2L_SYNC();

        // ---- B5: gY1 = fold_reflect2(corr^T(gY2, Wg)) on rows -2..TH+1, runs -1..G (zero outside the image); dWg ----
        // Border rules as in isp_bwd3.cuh B5: the reflect-2 fold and the pad sites' share of dWg are a few extra products
        // inside the items of rows 1,2 / H-2,H-3 and of the first / last run of the image.  Y1 centres come from the
        // plane the forward saved.
        { R2L_FOR_THREADS(NT) {
            // The weights are read next to their use (uniform-address LDS; volatile keeps the compiler from hoisting all
            // 25 into registers for the whole phase).
            const volatile float* wg = T->Wg;
            // -- owned rectangle (the items that carry the statistic): one TAP ROW A of the 5x5 at a time over all of the
            // thread's items, so only that row's five packed (image A, image B) sums are live (parked per tap row).
            // Tap row A reaches window row 4 - A; the folded pad contributions that use tap row A ride in the same pass.
            {
#ifdef R2L_HOST_EMU
                f2 c5[NI5][4];
                load_c5(tid, c5);
#endif
                f2 (&c)[NI5][4] = c5;
                f2 out[NI5][4];
                int ir[NI5], ig[NI5], rt[NI5];
                bool inside[NI5], lft[NI5], rgt[NI5];
#pragma unroll
                for (int u = 0; u < NI5; ++u) {
                    const int item = tid + u * NT;
                    const int r = item / G, g = item - r * G;
                    const int qy = ty0 + r, qx = tx0 + 4 * g;
                    ir[u] = r; ig[u] = g;
                    inside[u] = qy < H && qx < W;
                    lft[u] = qx == 0; rgt[u] = qx + 4 == W;
                    // row type of the folded-onto rows: 1 -> row 1, 2 -> row 2, 3 -> row H-2, 4 -> row H-3
                    rt[u] = qy == 1 ? 1 : (qy == 2 ? 2 : (qy == H - 2 ? 3 : (qy == H - 3 ? 4 : 0)));
#pragma unroll
                    for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                }
                // Pad columns (reflect-2 of the sharpened plane: -1 -> site 1, -2 -> site 2, W -> site W-2, W+1 -> W-3) are
                // realised on the data, so the loop is straight-line code: the adjoint of the first / last run of the image
                // uses per-site weights for its sites 1 and 2 (tap b gains the tap that reaches the same gradient through
                // the pad), the statistic adds the pad sites' products with centres masked by L / R (zero elsewhere).
                // Items outside the image (partial tiles) run with zero centres and are zeroed when stored.
                float Lf[NI5], Rf[NI5];
                f2 cl1[NI5], cl2[NI5], cr1[NI5], cr2[NI5];
                bool padrows = false;
#pragma unroll
                for (int u = 0; u < NI5; ++u) {
                    Lf[u] = (inside[u] && lft[u]) ? 1.f : 0.f;
                    Rf[u] = (inside[u] && rgt[u]) ? 1.f : 0.f;
                    cl1[u] = mul2s(c[u][1], Lf[u]); cl2[u] = mul2s(c[u][2], Lf[u]);
                    cr1[u] = mul2s(c[u][1], Rf[u]); cr2[u] = mul2s(c[u][2], Rf[u]);
                    if (!inside[u]) rt[u] = 0;
                    padrows |= rt[u] != 0;
                }
                // products of window row d (gY2 row q.y - 2 + d, columns q.x - 2 .. q.x + 5) with the tap row in w5 / acc
                // SIDE (compile time): the tile touches the left or right image border, so some item may carry pad-column
                // work; in an interior tile column (half of them at W = 256, more on wider frames) L = R = 0 everywhere
                // and the six per-site weights / six masked-centre products per item and tap row are not even issued
                // (they were 20 % of this loop).  The choice is uniform over the CTA: no divergence.
                auto full = [&](auto side_c, int u, int d, const float (&w5)[5], f2 (&acc)[5]) {
                    constexpr bool SIDE = decltype(side_c)::value;
                    // sites 1 and 2: out[1] also receives row[3] w0 + row[2] w1 (L) and row[5] w4 (R), out[2] receives
                    // row[2] w0 (L) and row[5] w3 + row[4] w4 (R); row[i] is column q.x - 2 + i
                    const float w1_0 = SIDE ? fmaf_(Rf[u], w5[4], w5[0]) : w5[0], w1_2 = SIDE ? fmaf_(Lf[u], w5[0], w5[2]) : w5[2],
                                w1_3 = SIDE ? fmaf_(Lf[u], w5[1], w5[3]) : w5[3];
                    const float w2_1 = SIDE ? fmaf_(Rf[u], w5[3], w5[1]) : w5[1], w2_2 = SIDE ? fmaf_(Rf[u], w5[4], w5[2]) : w5[2],
                                w2_4 = SIDE ? fmaf_(Lf[u], w5[0], w5[4]) : w5[4];
                    f2 row[8];
                    ld8<PN>(PG, (ir[u] + 2 + d) * PN + 2 * (ig[u] + 2), row);
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) {
                        out[u][0] = fma2s(row[4 - bb], w5[bb], out[u][0]);
                        out[u][3] = fma2s(row[7 - bb], w5[bb], out[u][3]);
                    }
                    out[u][1] = fma2s(row[5], w1_0, fma2s(row[4], w5[1], fma2s(row[3], w1_2, fma2s(row[2], w1_3, fma2s(row[1], w5[4], out[u][1])))));
                    out[u][2] = fma2s(row[6], w5[0], fma2s(row[5], w2_1, fma2s(row[4], w2_2, fma2s(row[3], w5[3], fma2s(row[2], w2_4, out[u][2])))));
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb)
#pragma unroll
                        for (int j = 0; j < 4; ++j) acc[bb] = fma2vv(c[u][j], row[j + 4 - bb], acc[bb]);
                    // pad sites' share of dWg: Y1pad(-1) = Y1(1), Y1pad(-2) = Y1(2), Y1pad(W) = Y1(W-2), Y1pad(W+1) = Y1(W-3)
                    if (SIDE) {
                        acc[0] = fma2vv(cl1[u], row[3], fma2vv(cl2[u], row[2], acc[0]));
                        acc[1] = fma2vv(cl1[u], row[2], acc[1]);
                        acc[3] = fma2vv(cr2[u], row[5], acc[3]);
                        acc[4] = fma2vv(cr2[u], row[4], fma2vv(cr1[u], row[5], acc[4]));
                    }
                };
                const bool side_tile = tx0 == 0 || tx0 + TW >= W;          // CTA-uniform
#pragma unroll 1
                for (int A = 0; A < 5; ++A) {
                    f2 acc[5];
                    float wa[5], w5[5];
                    R2L_PARK_LOAD(5, kB5Wg + 5 * A, wa)
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) { acc[bb] = mk2(0.f, 0.f); w5[bb] = wg[A * 5 + bb]; }
                    if (side_tile) {
#pragma unroll
                        for (int u = 0; u < NI5; ++u) full(std::true_type(), u, 4 - A, w5, acc);
                    } else {
#pragma unroll
                        for (int u = 0; u < NI5; ++u) full(std::false_type(), u, 4 - A, w5, acc);
                    }
                    R2L_PARK_READY(5, wa)
#pragma unroll
                    for (int bb = 0; bb < 5; ++bb) wa[bb] += acc[bb].x + acc[bb].y;
                    R2L_PARK_STORE(5, kB5Wg + 5 * A, wa)
                }
                // folded pad rows (isp_bwd3.cuh B5: pad row -1 -> row 1, -2 -> row 2, H -> row H-2, H+1 -> row H-3): only
                // warps that hold one of those four image rows come here.  Row type rt acts through tap row A on window
                // row d: (rt 1: A 0 / d 2, A 1 / d 1), (rt 2: A 0 / d 0), (rt 3: A 3 / d 3, A 4 / d 2), (rt 4: A 4 / d 4).
                if (R2L_ANY(padrows)) {
#pragma unroll 1
                    for (int A = 0; A < 5; ++A) {
                        if (A == 2) continue;
                        f2 acc[5];
                        float wa[5], w5[5];
                        R2L_PARK_LOAD(5, kB5Wg + 5 * A, wa)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) { acc[bb] = mk2(0.f, 0.f); w5[bb] = wg[A * 5 + bb]; }
#pragma unroll
                        for (int u = 0; u < NI5; ++u) {
                            const int t = rt[u];
                            const int d = A == 0 ? (t == 2 ? 0 : (t == 1 ? 2 : -1)) : A == 1 ? (t == 1 ? 1 : -1)
                                        : A == 3 ? (t == 3 ? 3 : -1) : (t == 3 ? 2 : (t == 4 ? 4 : -1));
                            if (d >= 0) full(std::true_type(), u, d, w5, acc);
                        }
                        R2L_PARK_READY(5, wa)
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) wa[bb] += acc[bb].x + acc[bb].y;
                        R2L_PARK_STORE(5, kB5Wg + 5 * A, wa)
                    }
                }
#pragma unroll
                for (int u = 0; u < NI5; ++u) {
                    if (!inside[u]) {
#pragma unroll
                        for (int j = 0; j < 4; ++j) out[u][j] = mk2(0.f, 0.f);
                    }
                    st4<PN>(GY1, (ir[u] + 2) * PN + 2 * (ig[u] + 2), out[u][0], out[u][1], out[u][2], out[u][3]);
                }
            }
#ifndef R2L_HOST_EMU
            load_c6(tid, c6);          // B6's Y0 centres, ahead of the halo ring (in-call A/B: -1.5 us against after it)
#endif
            // -- halo ring: single runs, no statistic
            for (int item = TH * G + tid; item < Cfg::G1H * GG; item += NT) {
                int r, g;
                region_item<TH, G, 2>(item, r, g);
                const int qy = ty0 + r, qx = tx0 + 4 * g;
                const bool inside = qy >= 0 && qy < H && qx >= 0 && qx < W;
                f2 out[4] = {mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f), mk2(0.f, 0.f)};
                if (inside) {
                    const bool lft = qx == 0, rgt = qx + 4 == W;
                    const int rt = qy == 1 ? 1 : (qy == 2 ? 2 : (qy == H - 2 ? 3 : (qy == H - 3 ? 4 : 0)));
                    const float Lf = lft ? 1.f : 0.f, Rf = rgt ? 1.f : 0.f;
                    auto full = [&](int d, int A) {
                        f2 row[8];
                        ld8<PN>(PG, (r + 2 + d) * PN + 2 * (g + 2), row);
                        float w5[5];
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) w5[bb] = wg[A * 5 + bb];
                        const float w1_0 = fmaf_(Rf, w5[4], w5[0]), w1_2 = fmaf_(Lf, w5[0], w5[2]), w1_3 = fmaf_(Lf, w5[1], w5[3]);
                        const float w2_1 = fmaf_(Rf, w5[3], w5[1]), w2_2 = fmaf_(Rf, w5[4], w5[2]), w2_4 = fmaf_(Lf, w5[0], w5[4]);
#pragma unroll
                        for (int bb = 0; bb < 5; ++bb) {
                            out[0] = fma2s(row[4 - bb], w5[bb], out[0]);
                            out[3] = fma2s(row[7 - bb], w5[bb], out[3]);
                        }
                        out[1] = fma2s(row[5], w1_0, fma2s(row[4], w5[1], fma2s(row[3], w1_2, fma2s(row[2], w1_3, fma2s(row[1], w5[4], out[1])))));
                        out[2] = fma2s(row[6], w5[0], fma2s(row[5], w2_1, fma2s(row[4], w2_2, fma2s(row[3], w5[3], fma2s(row[2], w2_4, out[2])))));
                    };
#pragma unroll 1
                    for (int d = 0; d < 5; ++d) full(d, 4 - d);
                    if (rt == 2) full(0, 0);
                    if (rt == 1) { full(1, 1); full(2, 0); }
                    if (rt == 3) { full(2, 4); full(3, 3); }
                    if (rt == 4) full(4, 4);
                }
                st4<PN>(GY1, (r + 2) * PN + 2 * (g + 2), out[0], out[1], out[2], out[3]);
            }
        } }
        R2L_SYNC();

        // ---- B6: gY0 = corr^T(gY1, Ws) (zero pad) on rows -1..TH, runs -1..G, zero outside the image; Ws statistic ----
#ifndef R2L_HOST_EMU
        prefetch_next(2);
#endif
        { R2L_FOR_THREADS(NT) {
            float ws[9];
#pragma unroll
            for (int t = 0; t < 9; ++t) ws[t] = T->Ws[t];
            f2 ws2[9];                                                 // packed (image A, image B) dWs sums of this tile
            float wspark[9];
            R2L_PARK_LOAD(9, kB5Ws, wspark)
#pragma unroll
            for (int t = 0; t < 9; ++t) ws2[t] = mk2(0.f, 0.f);
#ifdef R2L_HOST_EMU
            f2 c6[NI6][2][4];
            load_c6(tid, c6);
#endif
            auto b6 = [&](auto NRc, int r, int g, bool owned, const f2 (*pre)[4]) {
                constexpr int NR = decltype(NRc)::value;
                const int qy = ty0 + r;
                f2 c[NR][4], out[NR][4];
#pragma unroll
                for (int u = 0; u < NR; ++u) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) { out[u][j] = mk2(0.f, 0.f); c[u][j] = owned ? pre[u][j] : mk2(0.f, 0.f); }
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
                        for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                            for (int u = 0; u < NR; ++u)
#pragma unroll
                                for (int j = 0; j < 4; ++j)
                                    ws2[aa * 3 + bb] = fma2vv(c[u][j], row[u][j + 2 - bb], ws2[aa * 3 + bb]);
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
#pragma unroll
            for (int v = 0; v < NI6; ++v) {                             // owned rectangle: paired items
                const int item = tid + v * NT;
                b6(std::integral_constant<int, 2>(), item / (G / 2), item % (G / 2), true, c6[v]);
            }
            for (int item = TH * G + tid; item < (TH + 2) * GG; item += NT) {      // halo ring, single runs
                int r, g;
                region_item<TH, G, 1>(item, r, g);
                b6(std::integral_constant<int, 1>(), r, g, false, nullptr);
            }
            R2L_PARK_READY(9, wspark)
#pragma unroll
            for (int t = 0; t < 9; ++t) wspark[t] += ws2[t].x + ws2[t].y;
            R2L_PARK_STORE(9, kB5Ws, wspark)
#ifndef R2L_HOST_EMU
            load_x7(tid, xa7, xb7);
#endif
        } }
        R2L_SYNC();

#ifndef R2L_HOST_EMU
        prefetch_next(3);
#endif
        // ---- B7: Q' / P statistics and g_raw from the (gY0, gU, gV) windows; border rules as in isp_bwd3.cuh B7 ----
        // One gradient plane k and one TAP ROW A of the 3x3 at a time over the thread's items (rows of its CFA row phase
        // x runs): only that pass's six packed (image A, image B) Q' sums are live, a tap costs one FFMA2 for the
        // statistic and one for g_raw, and the items keep their raw centres and g_raw sums across the nine passes.
        // Tap row A reaches window row 2 - A; the reflect-1 pad row above row 1 (below row H-2) acts through tap row 0
        // (2) on window row 0 (2), pad columns through taps b = 0 / 2 of the first / last run.
        { R2L_FOR_THREADS(NT) {
            const int rp = (tid >> 5) & 1, slot = ((tid >> 6) << 5) | (tid & 31);
            constexpr int NI = NI7;
            const float one = a.B > 0 ? 1.f : 2.f;                     // 1.0 the compiler cannot see (pack2)
#ifdef R2L_HOST_EMU
            f4 xa7[NI7], xb7[NI7];
            load_x7(tid, xa7, xb7);
#endif
            f2 c[NI][4], graw[NI][4];
            int ir[NI], ig[NI];
            bool live[NI], f_top[NI], f_bot[NI], f_lft[NI], f_rgt[NI];
#pragma unroll
            for (int u = 0; u < NI; ++u) {
                const int i = slot + u * HALF;
                const int ri = i / G, g = i - ri * G;
                const int r = rp + 2 * ri;
                const int qy = ty0 + r, qx = tx0 + 4 * g;
                ir[u] = r; ig[u] = g;
                live[u] = qy < H && qx < W;                             // partial tiles: nothing there (all gradients zero)
                f_top[u] = qy == 1; f_bot[u] = qy == H - 2; f_lft[u] = qx == 0; f_rgt[u] = qx + 4 == W;
#pragma unroll
                for (int j = 0; j < 4; ++j) { c[u][j] = mk2(0.f, 0.f); graw[u][j] = mk2(0.f, 0.f); }
                if (live[u]) {                                          // raw centres of the 4 sites, packed once
                    if (sizeof(RawT) == 4) {
                        const f4 xa = xa7[u], xb = xb7[u];
                        c[u][0] = pack2(xa.x, xb.x, one); c[u][1] = pack2(xa.y, xb.y, one); c[u][2] = pack2(xa.z, xb.z, one); c[u][3] = pack2(xa.w, xb.w, one);
                    } else {
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            c[u][j] = pack2(RawLoad<RawT>::get(imgA + (size_t)qy * W + qx + j, a.denom),
                                            RawLoad<RawT>::get(imgB + (size_t)qy * W + qx + j, a.denom), one);
                    }
                }
            }
            // Pad columns (reflect-1 of the mosaic: -1 -> site 1 of the first run through tap b = 0, W -> site 2 of the last
            // run through b = 2) on the data: g_raw uses a per-site weight for those two taps, the statistic adds the
            // pad site's product with the centre masked by L / R.  The loop is straight-line code; items outside the
            // image (partial tiles) have zero centres, see zero gradients and are not stored.
            float Lf[NI], Rf[NI];
            f2 cl1[NI], cr2[NI];
            bool padrows = false;
#pragma unroll
            for (int u = 0; u < NI; ++u) {
                Lf[u] = (live[u] && f_lft[u]) ? 1.f : 0.f;
                Rf[u] = (live[u] && f_rgt[u]) ? 1.f : 0.f;
                cl1[u] = mul2s(c[u][1], Lf[u]); cr2[u] = mul2s(c[u][2], Rf[u]);
                padrows |= live[u] && (f_top[u] | f_bot[u]);
            }
            // products of window row d (g_yuv[k] row q.y - 1 + d, columns q.x - 1 .. q.x + 4) with the tap row in w / acc
            // SIDE as in B5: pad-column work only in the tiles of the first / last tile column
            auto full = [&](auto side_c, int u, const f2* pl, int d, const float (&w)[2][3], f2 (&acc)[2][3]) {
                constexpr bool SIDE = decltype(side_c)::value;
                f2 row[6];
                ld6<PN>(pl, (ir[u] + 3 + d) * PN + 2 * (ig[u] + 2), row);
#pragma unroll
                for (int bb = 0; bb < 3; ++bb)
#pragma unroll
                    for (int j = 0; j < 4; ++j) acc[j & 1][bb] = fma2vv(c[u][j], row[j + 2 - bb], acc[j & 1][bb]);
                // pad sites' share of Q': rawpad(-1) = raw(1), rawpad(W) = raw(W-2)
                if (SIDE) {
                    acc[1][0] = fma2vv(cl1[u], row[1], acc[1][0]);
                    acc[0][2] = fma2vv(cr2[u], row[4], acc[0][2]);
                }
                if (Cfg::GRAW) {                                        // g_raw(q) += sum_b g_yuv[k](q - (a-1, b-1)) AWq[par(q)][k][a][b]
                    // site 1 also receives row[1] w[1][0] (L), site 2 row[4] w[0][2] (R); row[i] is column q.x - 1 + i
                    const float w1_2 = SIDE ? fmaf_(Lf[u], w[1][0], w[1][2]) : w[1][2], w2_0 = SIDE ? fmaf_(Rf[u], w[0][2], w[0][0]) : w[0][0];
#pragma unroll
                    for (int bb = 0; bb < 3; ++bb) {
                        graw[u][0] = fma2s(row[2 - bb], w[0][bb], graw[u][0]);
                        graw[u][3] = fma2s(row[5 - bb], w[1][bb], graw[u][3]);
                    }
                    graw[u][1] = fma2s(row[3], w[1][0], fma2s(row[2], w[1][1], fma2s(row[1], w1_2, graw[u][1])));
                    graw[u][2] = fma2s(row[4], w2_0, fma2s(row[3], w[0][1], fma2s(row[2], w[0][2], graw[u][2])));
                }
            };
            const bool side_tile7 = tx0 == 0 || tx0 + TW >= W;         // CTA-uniform
#pragma unroll 1
            for (int k = 0; k < 3; ++k) {
                const f2* pl = PU + ((k + 2) % 3) * Cfg::kF;            // k = 0: gY0 (PG), 1: gU (PU), 2: gV (PV)
                const volatile float* awq = &T->AWq[2 * rp][k][0];      // [col phase * 27 + tap], read next to their use
#pragma unroll
                for (int A = 0; A < 3; ++A) {
                    f2 acc[2][3];
                    float qa[8], w[2][3];                               // 6 Q' sums [col phase][b], then (A == 1 only) 2 P sums
                    if (A == 1) { R2L_PARK_LOAD(6, kB5Q + kB5QStride * k + 6 * A, qa) R2L_PARK_LOAD(2, kB5Q + kB5QStride * k + 18, qa + 6) }
                    else { R2L_PARK_LOAD(6, kB5Q + kB5QStride * k + 6 * A, qa) }
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) { acc[cp][bb] = mk2(0.f, 0.f); w[cp][bb] = awq[cp * 27 + A * 3 + bb]; }
                    if (side_tile7) {
#pragma unroll
                        for (int u = 0; u < NI; ++u) full(std::true_type(), u, pl, 2 - A, w, acc);
                    } else {
#pragma unroll
                        for (int u = 0; u < NI; ++u) full(std::false_type(), u, pl, 2 - A, w, acc);
                    }
                    if (A == 1) { R2L_PARK_READY(8, qa) } else { R2L_PARK_READY(6, qa) }
#pragma unroll
                    for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                        for (int bb = 0; bb < 3; ++bb) qa[cp * 3 + bb] += acc[cp][bb].x + acc[cp][bb].y;
                    R2L_PARK_STORE(6, kB5Q + kB5QStride * k + 6 * A, qa)
                    if (A == 1) {                                       // P[col phase] = sum of g_yuv[k] over the owned sites
                        f2 p2[2] = {mk2(0.f, 0.f), mk2(0.f, 0.f)};
#pragma unroll
                        for (int u = 0; u < NI; ++u) {
                            f2 row[4];
                            ld4<PN>(pl, (ir[u] + 4) * PN + 2 * (ig[u] + 2), row);
#pragma unroll
                            for (int j = 0; j < 4; ++j) p2[j & 1] = fma2s(row[j], one, p2[j & 1]);      // one FFMA2 (two FADDs otherwise)
                        }
                        qa[6] += p2[0].x + p2[0].y; qa[7] += p2[1].x + p2[1].y;
                        R2L_PARK_STORE(2, kB5Q + kB5QStride * k + 18, qa + 6)
                    }
                }
            }
            // reflect-1 pad rows: the row above row 1 acts through tap row 0 on window row 0, the row below row H-2 through
            // tap row 2 on window row 2.  Only warps that hold image row 1 or H-2 come here.
            if (R2L_ANY(padrows)) {
#pragma unroll 1
                for (int k = 0; k < 3; ++k) {
                    const f2* pl = PU + ((k + 2) % 3) * Cfg::kF;
                    const volatile float* awq = &T->AWq[2 * rp][k][0];
#pragma unroll 1
                    for (int A = 0; A < 3; A += 2) {
                        f2 acc[2][3];
                        float qa[6], w[2][3];
                        R2L_PARK_LOAD(6, kB5Q + kB5QStride * k + 6 * A, qa)
#pragma unroll
                        for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) { acc[cp][bb] = mk2(0.f, 0.f); w[cp][bb] = awq[cp * 27 + A * 3 + bb]; }
#pragma unroll
                        for (int u = 0; u < NI; ++u)
                            if (live[u] && (A == 0 ? f_top[u] : f_bot[u])) full(std::true_type(), u, pl, A, w, acc);
                        R2L_PARK_READY(6, qa)
#pragma unroll
                        for (int cp = 0; cp < 2; ++cp)
#pragma unroll
                            for (int bb = 0; bb < 3; ++bb) qa[cp * 3 + bb] += acc[cp][bb].x + acc[cp][bb].y;
                        R2L_PARK_STORE(6, kB5Q + kB5QStride * k + 6 * A, qa)
                    }
                }
            }
            if (Cfg::GRAW) {
#pragma unroll
                for (int u = 0; u < NI; ++u) {
                    if (!live[u]) continue;
                    const size_t off = (size_t)(ty0 + ir[u]) * W + tx0 + 4 * ig[u];
                    float* pa = a.graw + (size_t)b0 * plane + off;
                    f4 va; va.x = graw[u][0].x; va.y = graw[u][1].x; va.z = graw[u][2].x; va.w = graw[u][3].x;
                    *reinterpret_cast<f4*>(pa) = va;
                    if (!dup) {
                        float* pb = a.graw + (size_t)b1 * plane + off;
                        f4 vb; vb.x = graw[u][0].y; vb.y = graw[u][1].y; vb.z = graw[u][2].y; vb.w = graw[u][3].y;
                        *reinterpret_cast<f4*>(pb) = vb;
                    }
                }
            }
        } }
        R2L_SYNC();   // planes are rewritten by the next tile
    }

    // ---- CTA reduction of the per-thread statistics into the kStat* layout (deterministic, fixed order), as in
    // isp_bwd3.cuh; the sums come back from tensor memory ----------------------------------------------------------
    float* part = a.partials + (size_t)cta * kStatPitch;
    constexpr int NW = NT / 32;
    constexpr int RP = kBwd5AccFloats + 1;
    float* red = reinterpret_cast<float*>(PU);                       // [NW][RP]
#ifdef R2L_HOST_EMU
    for (int w = 0; w < NW; ++w)
        for (int i = 0; i < kBwd5AccFloats; ++i) {
            float sum = 0.f;
            for (int l = 0; l < 32; ++l) sum += accs[w * 32 + l].sums[i];
            red[w * RP + i] = sum;
        }
#else
    {
        static_assert(kBwd5AccFloats == 96, "three groups of 32 running sums");
        const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
        for (int grp = 0; grp < 3; ++grp) {
            float v[32];
            __syncwarp();
            tmem::wait_st();
            tmem::load<16>(tacc + grp * 32, v);
            tmem::load<16>(tacc + grp * 32 + 16, v + 16);
            tmem::ready<32>(v);
            red[warp * RP + grp * 32 + lane] = warp_transpose_sum32(v);
        }
        tmem::fence_before_sync();
    }
#endif
    R2L_SYNC();
#ifndef R2L_HOST_EMU
    if (threadIdx.x < 32) tmem::dealloc<Cfg::kTmemCols>(tmem_base);  // every warp has read its sums back
#endif
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
                const int off = kB5Q + kB5QStride * k + (is_q ? 6 * (tt / 3) + 3 * cpq + tt % 3 : 18 + cpq);
                for (int w = rpq; w < NW; w += 2) sum += red[w * RP + off];      // warps of row phase rpq
            }
            part[s] = sum;
        }
    } }
#ifndef R2L_HOST_EMU
    if (a.ticket) {
        // Two-level finish (isp_bwd4.cuh): the last CTA of each static group sums the group's rows, the last group
        // finisher adds the group rows and runs the chain rule.  Tickets: word 0 = the groups, words 16 .. 31 = one per group
        // (the launch-tagged 64-bit words of take_ticket; bytes 128 .. 255 of the workspace's ticket block).
        __shared__ unsigned last_flag;
        const int gs = finish_group_size(n_cta), grp = cta / gs;
        const int c0 = grp * gs, c1 = min(n_cta, c0 + gs), ng = (n_cta + gs - 1) / gs;
        unsigned* gticket = a.ticket + 2 * (kFinishGroups + grp);
        double* rows = finish_group_rows(a.partials);
        double* scratch = reinterpret_cast<double*>(PU);
        static_assert((size_t)(NT / 32) * kStatPitch * 8 <= (size_t)Cfg::kSites * 8 &&
                      (size_t)(kStatPitch + 130) * 8 <= (size_t)Cfg::kSites * 8, "finish scratch fits the planes");
        __threadfence();                                             // this CTA's partial sums are visible device-wide ...
        __syncthreads();
        if (threadIdx.x == 0) last_flag = take_ticket(gticket, a.ticket_gen) == (unsigned)(c1 - c0) - 1u;   // ... before its ticket is
        __syncthreads();
        if (last_flag) {                                             // last of its group: the group's rows -> one fp64 row
            __threadfence();
            if (threadIdx.x == 0) clear_ticket(gticket);
            finish_group_sum<NT>(a.partials, c0, c1, rows + (size_t)grp * kStatPitch, scratch);
            __threadfence();                                         // the group row is visible before the group's ticket
            __syncthreads();
            if (threadIdx.x == 0) last_flag = take_ticket(a.ticket, a.ticket_gen) == (unsigned)ng - 1u;
            __syncthreads();
            if (last_flag) {                                         // last group: group rows -> gradients
                __threadfence();
                if (threadIdx.x == 0) clear_ticket(a.ticket);
                finish_from_groups<NT>(T, rows, ng, a.grads, scratch);
                if (a.world > 1) peer_allreduce<NT>(a);
            }
        }
    }
#endif
}

}  // namespace r2l
