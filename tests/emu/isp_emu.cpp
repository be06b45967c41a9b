// Host emulation of the CUDA CTA functions (raw2logit_b200/csrc/isp_core.cuh) -- TEST INFRASTRUCTURE ONLY.
// g++ compiles the same source the GPU kernels are made of, with threads of a CTA run one after another and
// barriers turned into loop boundaries.  tests/test_emu_logic.py compares it with the oracle so that indexing,
// border handling and the hand-derived adjoints are checked in a container that has no GPU.
// Never loaded by the raw2logit_b200 package; it is not a fallback.
#define R2L_HOST_EMU 1
#include <vector>
#include <cstring>
#include "../../include/r2l_isp.h"
#include "../../raw2logit_b200/csrc/isp_config.h"

using namespace r2l;

static Params to_params(const r2l_isp_params* p) {
    Params q;
    q.black_level = p->black_level; q.white_balance = p->white_balance; q.colour_correction = p->colour_correction;
    q.gamma_correct = p->gamma_correct; q.debayer_weight = p->debayer_weight; q.sharpen_weight = p->sharpen_weight;
    q.gauss_weight = p->gauss_weight; q.rgb2yuv = p->rgb2yuv; q.yuv2rgb = p->yuv2rgb;
    return q;
}

template <class Cfg, typename RawT>
static void run_forward_v1(const FwdArgs& a, int n_cta) {
    const TileGrid grid = make_grid(a.B, a.H, a.W, Cfg::TH, Cfg::TW);
    std::vector<float> smem(Cfg::kSmemFloats);
    for (int cta = 0; cta < n_cta; ++cta) fwd_cta<Cfg, RawT, false>(cta, n_cta, a, grid, smem.data());
}

template <typename RawT, bool STATS, bool TAIL>
static void run_forward3(const FwdArgs& a, int n_cta) {
    using Cfg = Fwd3Default;
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);
    std::vector<float> smem(Cfg::kSmemBytes / 4 + 4);
    float* base = smem.data();
    while (reinterpret_cast<uintptr_t>(base) % 16) ++base;
    for (int cta = 0; cta < n_cta; ++cta) fwd3_cta<Cfg, RawT, STATS, TAIL, false>(cta, n_cta, a, grid, base);
}

template <class Cfg, typename RawT, bool STATS>
static void run_forward(const FwdArgs& a, int n_cta) {
    const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);
    std::vector<float> smem(Cfg::kSmemBytes / 4 + 4);
    float* base = smem.data();
    while (reinterpret_cast<uintptr_t>(base) % 16) ++base;          // the kernels assume 16-byte aligned shared memory
    for (int cta = 0; cta < n_cta; ++cta) fwd2_cta<Cfg, RawT, STATS>(cta, n_cta, a, grid, base);
}

template <class Cfg, typename RawT>
static void run_backward(const BwdArgs& a, int n_cta, float* grads, int version) {
    if (version == 1) {
        const TileGrid grid = make_grid(a.B, a.H, a.W, Cfg::TH, Cfg::TW);
        std::vector<float> smem(Cfg::kSmemFloats);
        for (int cta = 0; cta < n_cta; ++cta) bwd_cta<Cfg, RawT>(cta, n_cta, a, grid, smem.data());
    } else if (version == 5 && a.out && a.luma && bwd5_shape_ok(a.H, a.W)) {     // same dispatch rule as the CUDA launcher
        auto go5 = [&](auto cfg) {
            using C5 = decltype(cfg);
            const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, C5::TH, C5::TW);
            std::vector<float> smem(C5::kSmemBytes / 4 + 4);
            float* base = smem.data();
            while (reinterpret_cast<uintptr_t>(base) % 16) ++base;
            for (int cta = 0; cta < n_cta; ++cta) bwd5_cta<C5, RawT>(cta, n_cta, a, grid, base);
        };
        if (a.gtail) go5(Bwd5<Cfg::GRAW, true>());
        else go5(Bwd5<Cfg::GRAW, false>());
    } else if (version == 4 && a.out && a.luma && bwd4_shape_ok(a.H, a.W)) {
        auto go4 = [&](auto cfg) {
            using C4 = decltype(cfg);
            const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, C4::TH, C4::TW);
            std::vector<float> smem(C4::kSmemBytes / 4 + 4);
            float* base = smem.data();
            while (reinterpret_cast<uintptr_t>(base) % 16) ++base;
            for (int cta = 0; cta < n_cta; ++cta) bwd4_cta<C4, RawT>(cta, n_cta, a, grid, base);
        };
        if (a.gtail) go4(Bwd4<Cfg::GRAW, true>());
        else go4(Bwd4<Cfg::GRAW, false>());
    } else if (version == 3 || version == 4 || version == 5) {
        if (!bwd3_shape_ok(a.H, a.W, Cfg::TH, Cfg::TW)) {           // same dispatch rule as the CUDA launcher
            run_backward<Cfg, RawT>(a, n_cta, grads, 1);
            return;
        }
        const TileGrid grid = make_grid((a.B + 1) / 2, a.H, a.W, Cfg::TH, Cfg::TW);
        auto go = [&](auto cfg) {
            using C3 = decltype(cfg);
            std::vector<float> smem(C3::kSmemBytes / 4 + 4);
            float* base = smem.data();
            while (reinterpret_cast<uintptr_t>(base) % 16) ++base;
            for (int cta = 0; cta < n_cta; ++cta) bwd3_cta<C3, RawT, false>(cta, n_cta, a, grid, base);
        };
        if (a.out) {
            if (a.gtail) go(Bwd3Cfg<Cfg::TH, Cfg::TW, Cfg::NT, Cfg::GRAW, true, true>());
            else go(Bwd3Cfg<Cfg::TH, Cfg::TW, Cfg::NT, Cfg::GRAW, false, true>());
        } else if (a.gtail) go(Bwd3Cfg<Cfg::TH, Cfg::TW, Cfg::NT, Cfg::GRAW, true>());
        else go(Bwd3Cfg<Cfg::TH, Cfg::TW, Cfg::NT, Cfg::GRAW, false>());
    }
    // finish (same arithmetic as isp_backward_finish_kernel)
    std::vector<float> tmem((sizeof(Tables) + 3) / 4);
    Tables* T = reinterpret_cast<Tables*>(tmem.data());
    R2L_BUILD_TABLES(64, a.P, T)
    double S[kNumStats];
    for (int s = 0; s < kNumStats; ++s) {
        double sum = 0.0;
        for (int c = 0; c < n_cta; ++c) sum += (double)a.partials[(size_t)c * kStatPitch + s];
        S[s] = sum;
    }
    for (int e = 0; e < R2L_NUM_PARAM_GRADS; ++e) grads[e] = finish_grad(e, S, T);
}

extern "C" {

// all pointers are HOST pointers here
int emu_isp_forward(const void* raw, int raw_dtype, float denom, int B, int H, int W, const r2l_isp_params* params,
                    const r2l_isp_tail* tail, float* out, int n_cta, int version, double* chan_sums, float* luma) {
    if (H < 3 || W < 3) return R2L_ERR_BAD_SHAPE;
    if (luma && !(version == 3 && fwd3_shape_ok(H, W))) return R2L_ERR_BAD_ARGUMENT;
    FwdArgs a;
    a.luma = luma;
    a.raw = raw; a.denom = denom; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.additive = tail ? tail->additive : nullptr; a.affine = tail ? tail->affine : nullptr; a.out = out;
    a.chan_partials = nullptr;
    if (version == 1) {
        if (raw_dtype == R2L_F32) run_forward_v1<FwdDefault, float>(a, n_cta);
        else run_forward_v1<FwdDefault, uint16_t>(a, n_cta);
    } else if (version == 3 && fwd3_shape_ok(H, W)) {                 // same dispatch rule as the CUDA launcher
        std::vector<float> partials((size_t)n_cta * kChanPitch, 0.f);
        if (chan_sums) a.chan_partials = partials.data();
        const bool tailf = a.additive || a.affine;
        if (raw_dtype == R2L_F32) {
            if (chan_sums) run_forward3<float, true, false>(a, n_cta);
            else if (tailf) run_forward3<float, false, true>(a, n_cta);
            else run_forward3<float, false, false>(a, n_cta);
        } else {
            if (chan_sums) run_forward3<uint16_t, true, false>(a, n_cta);
            else if (tailf) run_forward3<uint16_t, false, true>(a, n_cta);
            else run_forward3<uint16_t, false, false>(a, n_cta);
        }
        if (chan_sums)
            for (int k = 0; k < 6; ++k) {
                double s = 0.0;
                for (int c = 0; c < n_cta; ++c) s += partials[(size_t)c * kChanPitch + k];
                chan_sums[k] = s;
            }
    } else if (chan_sums) {
        std::vector<float> partials((size_t)n_cta * kChanPitch, 0.f);
        a.chan_partials = partials.data();
        if (raw_dtype == R2L_F32) run_forward<Fwd2Default, float, true>(a, n_cta);
        else run_forward<Fwd2Default, uint16_t, true>(a, n_cta);
        for (int k = 0; k < 6; ++k) {
            double s = 0.0;
            for (int c = 0; c < n_cta; ++c) s += partials[(size_t)c * kChanPitch + k];
            chan_sums[k] = s;
        }
    } else {
        if (raw_dtype == R2L_F32) run_forward<Fwd2Default, float, false>(a, n_cta);
        else run_forward<Fwd2Default, uint16_t, false>(a, n_cta);
    }
    return R2L_OK;
}

int emu_isp_backward(const void* raw, int raw_dtype, float denom, int B, int H, int W, const r2l_isp_params* params,
                     const float* grad_out, const float* grad_tail, const float* additive, float* grad_raw,
                     float* grad_params, int n_cta, int version, const float* out, const float* luma) {
    if (H < 3 || W < 3) return R2L_ERR_BAD_SHAPE;
    std::vector<float> partials((size_t)n_cta * kStatPitch, 0.f);
    BwdArgs a;
    a.raw = raw; a.denom = denom; a.B = B; a.H = H; a.W = W; a.P = to_params(params);
    a.gout = grad_out; a.gtail = grad_tail; a.additive = additive; a.graw = grad_raw; a.partials = partials.data();
    a.out = out;
    a.luma = luma;
    if (grad_raw) {
        if (raw_dtype == R2L_F32) run_backward<BwdWithRaw, float>(a, n_cta, grad_params, version);
        else run_backward<BwdWithRaw, uint16_t>(a, n_cta, grad_params, version);
    } else {
        if (raw_dtype == R2L_F32) run_backward<BwdNoRaw, float>(a, n_cta, grad_params, version);
        else run_backward<BwdNoRaw, uint16_t>(a, n_cta, grad_params, version);
    }
    return R2L_OK;
}

}
