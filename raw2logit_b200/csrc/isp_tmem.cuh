// isp_tmem.cuh -- tensor memory (TMEM) as a per-thread accumulator file.
//
// The fused backward keeps 96 running sums per thread for the whole launch (the statistics behind the 132 parameter
// gradients, SURVEY 8a: a14-a17).  Held in registers they cap the kernel at 8 warps per SM and spill
// (profiles/r01_v4_summary.md).  Blackwell's tensor memory is 128 lanes x 512 columns x 32 bit per SM, reachable only
// through tcgen05.ld / tcgen05.st, where lane i of a warp addresses TMEM lane 32 * (warp % 4) + i: exactly a private
// 32-bit cell per thread and column.  No tensor-core instruction is involved; TMEM is used as a second register file:
// a phase loads the sums it updates (tcgen05.ld.32x32b.xN: N consecutive columns -> N registers), and stores them back
// when it ends.  Warps w and w + 4 of a CTA share TMEM lanes, so each takes its own column range.
// sm_100a only (device code); the host emulation keeps the sums in a per-thread array instead.
#pragma once
#ifndef R2L_HOST_EMU
#include <cstdint>

namespace r2l {
namespace tmem {

// one warp allocates NCOLS columns (power of two >= 32) for the CTA and publishes the base address in shared memory
template <int NCOLS> __device__ __forceinline__ void alloc(uint32_t* smem_slot) {
    static_assert(NCOLS >= 32 && NCOLS <= 512 && (NCOLS & (NCOLS - 1)) == 0, "TMEM columns: power of two in [32, 512]");
    const uint32_t dst = static_cast<uint32_t>(__cvta_generic_to_shared(smem_slot));
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "n"(NCOLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS> __device__ __forceinline__ void dealloc(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(NCOLS) : "memory");
}
__device__ __forceinline__ void fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// address of column `col` (relative to the CTA's allocation) in this warp's 32 lanes
__device__ __forceinline__ uint32_t addr(uint32_t base, int col) {
    const uint32_t lane_base = ((threadIdx.x >> 5) & 3u) * 32u;
    return base + (lane_base << 16) + (uint32_t)col;
}

#define R2L_U(x) __float_as_uint(x)
__device__ __forceinline__ void ld1(uint32_t a, float* v) {
    uint32_t r0;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r0) : "r"(a));
    v[0] = __uint_as_float(r0);
}
__device__ __forceinline__ void ld2(uint32_t a, float* v) {
    uint32_t r0, r1;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(a));
    v[0] = __uint_as_float(r0); v[1] = __uint_as_float(r1);
}
__device__ __forceinline__ void ld4(uint32_t a, float* v) {
    uint32_t r[4];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
#pragma unroll
    for (int i = 0; i < 4; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld8(uint32_t a, float* v) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a));
#pragma unroll
    for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void ld16(uint32_t a, float* v) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
                 : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
                   "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
                 : "r"(a));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void st1(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(R2L_U(v[0])) : "memory");
}
__device__ __forceinline__ void st2(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(a), "r"(R2L_U(v[0])), "r"(R2L_U(v[1])) : "memory");
}
__device__ __forceinline__ void st4(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};"
                 ::"r"(a), "r"(R2L_U(v[0])), "r"(R2L_U(v[1])), "r"(R2L_U(v[2])), "r"(R2L_U(v[3])) : "memory");
}
__device__ __forceinline__ void st8(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};"
                 ::"r"(a), "r"(R2L_U(v[0])), "r"(R2L_U(v[1])), "r"(R2L_U(v[2])), "r"(R2L_U(v[3])), "r"(R2L_U(v[4])),
                 "r"(R2L_U(v[5])), "r"(R2L_U(v[6])), "r"(R2L_U(v[7])) : "memory");
}
__device__ __forceinline__ void st16(uint32_t a, const float* v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};"
                 ::"r"(a), "r"(R2L_U(v[0])), "r"(R2L_U(v[1])), "r"(R2L_U(v[2])), "r"(R2L_U(v[3])), "r"(R2L_U(v[4])),
                 "r"(R2L_U(v[5])), "r"(R2L_U(v[6])), "r"(R2L_U(v[7])), "r"(R2L_U(v[8])), "r"(R2L_U(v[9])), "r"(R2L_U(v[10])),
                 "r"(R2L_U(v[11])), "r"(R2L_U(v[12])), "r"(R2L_U(v[13])), "r"(R2L_U(v[14])), "r"(R2L_U(v[15])) : "memory");
}
#undef R2L_U

// N consecutive columns <-> v[0..N), split into power-of-two pieces (all indices compile-time after unrolling).
// The caller waits (ready<N>(v) before the first use of v, wait_st before the columns are read again).
template <int N> __device__ __forceinline__ void load(uint32_t a, float* v) {
    static_assert(N >= 0 && N < 64, "pieces up to 16 columns");
    if constexpr (N >= 16) { ld16(a, v); load<N - 16>(a + 16, v + 16); }
    else if constexpr (N >= 8) { ld8(a, v); load<N - 8>(a + 8, v + 8); }
    else if constexpr (N >= 4) { ld4(a, v); load<N - 4>(a + 4, v + 4); }
    else if constexpr (N >= 2) { ld2(a, v); load<N - 2>(a + 2, v + 2); }
    else if constexpr (N >= 1) { ld1(a, v); }
}
// Wait for every tcgen05.ld of this thread and tie v[0..N) to the wait: the compiler only sees register outputs of the
// load instructions, not that they are filled asynchronously, so without the (empty, ordered) asm statements below it
// would be free to move a use of v above the wait.
template <int N> __device__ __forceinline__ void ready(float* v) {
    wait_ld();
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("" : "+f"(v[i]));
}
template <int N> __device__ __forceinline__ void store(uint32_t a, const float* v) {
    static_assert(N >= 0 && N < 64, "pieces up to 16 columns");
    if constexpr (N >= 16) { st16(a, v); store<N - 16>(a + 16, v + 16); }
    else if constexpr (N >= 8) { st8(a, v); store<N - 8>(a + 8, v + 8); }
    else if constexpr (N >= 4) { st4(a, v); store<N - 4>(a + 4, v + 4); }
    else if constexpr (N >= 2) { st2(a, v); store<N - 2>(a + 2, v + 2); }
    else if constexpr (N >= 1) { st1(a, v); }
}

}  // namespace tmem
}  // namespace r2l
#endif
