// Tile configurations shared by the CUDA kernels (isp_kernels.cu) and the host emulation (tests/emu).
#pragma once
#include "isp_core.cuh"
#include "isp_fwd2.cuh"
#include "isp_bwd3.cuh"
#include "isp_fwd3.cuh"
#include "isp_bwd4.cuh"
#include "isp_bwd5.cuh"

namespace r2l {
using FwdDefault = FwdCfg<32, 64, 256>;          // v1 (scalar) -- kept for the emulation cross-check only
using Fwd2Default = Fwd2Cfg<32, 64, 256>;        // v2: image pairs, FFMA2, register micro-tiles (any shape)
using Fwd3Default = Fwd3Cfg<32, 64, 256>;        // v3: border rules on the data, four barriers per tile (W % 4 == 0)
using BwdNoRaw = BwdCfg<32, 64, 256, false>;     // v1 (scalar) -- emulation cross-check only
using BwdWithRaw = BwdCfg<32, 64, 256, true>;
// v3: branch-free padded-domain phases; <TH, TW, NT, GRAW, TAIL, OUT (forward output available)>
template <bool GRAW, bool TAIL, bool OUT = false> using Bwd3 = Bwd3Cfg<32, 64, 256, GRAW, TAIL, OUT>;
// v4: forward output + saved Y0/Y1 planes, nothing recomputed, two CTAs per SM; <TH, TW, NT, GRAW, TAIL>
template <bool GRAW, bool TAIL> using Bwd4 = Bwd4Cfg<32, 64, 128, GRAW, TAIL>;
constexpr int kBwd4CtasPerSm = 2;
// v5: v4 with the running sums parked in tensor memory, 256 threads x 2 CTAs per SM at <= 128 registers
#ifndef R2L_B5_TH
#define R2L_B5_TH 32
#define R2L_B5_NT 256
#define R2L_B5_CPS 2
#endif
template <bool GRAW, bool TAIL> using Bwd5 = Bwd5Cfg<R2L_B5_TH, 64, R2L_B5_NT, GRAW, TAIL>;
constexpr int kBwd5CtasPerSm = R2L_B5_CPS;
}  // namespace r2l
