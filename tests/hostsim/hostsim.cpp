// hostsim.cpp — TEST HARNESS ONLY.  Compiles the K1 device code (fit_small.cuh) as
// plain host C++ (-DQNMFIT_HOSTSIM) and runs the lanes of each CTA one after another,
// phase by phase, exactly as the CUDA kernel orders them; and the K1p device code
// (fit_pair.cuh), whose lanes exchange data with shuffles inside the block loop, with the 32
// lanes of a warp in lock step (hostsim_warp.h: one fiber per lane); and K3 (fit_struct.cuh, the
// kernel function itself) with all threads of a CTA as fibers, warp collectives and
// __syncthreads emulated.  It lets the CPU-only test
// tier check the kernel's arithmetic (streamed TSQR, anchored recurrence, R-combine,
// back-substitution, mismatch sums) against the oracle.  It is never loaded by the
// qnmfits_b200 package.
//
// Build: g++ -O2 -std=c++17 -mfma -ffp-contract=off -fPIC -shared -DQNMFIT_HOSTSIM
//        -Iinclude -Iqnmfits_b200/csrc -o tests/hostsim/libqnmfit_hostsim.so hostsim.cpp
#include <stdlib.h>
#include <string.h>
#include <vector>

#include "qnmfit.h"
#include "fit_small.cuh"
#include "k1p_config.h"
#include "fit_pair.cuh"
#include "fit_struct.cuh"
#include "fit_panel.cuh"
#include "fit_general.cuh"

#define HS_THREADS 256

static void fill_params(const qnmfit_batch *b, int lpf, bool eval, FitParams *p)
{
    memset(p, 0, sizeof(*p));
    p->n_fits = b->n_fits; p->n_modes = b->n_modes; p->n_series = b->n_series; p->n_times = b->n_times;
    p->series_stride = b->series_stride; p->first_fit = b->first_fit;
    p->times = b->times; p->data = (const double2 *)b->data;
    p->row_begin = b->row_begin; p->row_end = b->row_end; p->t0 = b->t0;
    p->row_begin_all = b->row_begin_all; p->row_end_all = b->row_end_all; p->t0_all = b->t0_all;
    p->omega = (const double2 *)b->omega; p->omega_tilde = b->omega_tilde; p->mode_ptr = b->mode_ptr;
    p->inv_Mf = b->inv_Mf; p->delta_factor = b->delta_factor; p->chi_index = b->chi_index;
    p->mf_index = b->mf_index; p->n_chi = b->n_chi > 0 ? b->n_chi : 1; p->n_mf = b->n_mf;
    p->n_constituents = b->n_constituents;
    p->anchor_rows = b->anchor_rows > 0 ? b->anchor_rows : QNMFIT_DEFAULT_ANCHOR_ROWS;
    p->dt_nominal = b->dt_nominal;
    p->C = (double2 *)b->C; p->mismatch = b->mismatch; p->residual = b->residual;
    p->R = (double2 *)b->R; p->status = b->status;
    p->model = (double2 *)b->model; p->model_stride = b->model_stride; p->omega_shared = b->omega_shared;
    p->flagged_count = b->flagged_count;
    p->series_index = b->series_index;
    p->lanes_per_fit = lpf; p->eval_only = eval ? 1 : 0;
    p->fast_mismatch = (!eval && b->uniform_weights && b->dt_nominal > 0.0 && !b->model) ? 1 : 0;
}

template <int N>
static void run(const qnmfit_batch *b, int lpf, bool eval)
{
    FitParams p;
    fill_params(b, lpf, eval, &p);
    const int fpc = HS_THREADS / lpf;
    const int ctas = (b->n_fits + fpc - 1) / fpc;
    std::vector<unsigned char> smem(SmallSmem<N, HS_THREADS>::bytes(fpc, 0) + 64);
    for (int cta = 0; cta < ctas; ++cta) {
        SmallSmem<N, HS_THREADS> sm;
        sm.carve(smem.data(), fpc, 0);
        sm.ts = p.times; sm.ds = p.data; sm.t_off = 0;
        for (int idx = 0; idx < fpc * N; ++idx) {
            const int slot = idx / N, j = idx - slot * N;
            const int fit = cta * fpc + slot;
            if (fit < p.n_fits) {
                const double2 w = fit_omega(p, fit, j);
                sm.om[slot * sm.TS + j] = w;
                if (p.dt_nominal > 0.0) {
                    const double2 q = design_entry(w, p.dt_nominal);
                    sm.qq[slot * sm.TS + j] = q;
                    sm.qw[slot * sm.TS + j] = c_mul(q, make_double2(w.y, -w.x));
                }
            }
        }
        std::vector<SmallLane> lanes(HS_THREADS);
        std::vector<int> status(HS_THREADS, 0);
        std::vector<SmallAcc> acc(HS_THREADS);
        for (auto &a : acc) a.sdd = a.res2 = a.cn2 = 0.0;
        for (int tid = 0; tid < HS_THREADS; ++tid) {
            lanes[tid] = small_lane_setup(p, cta, tid, HS_THREADS, true, SmallLayout<N>::MB);
            small_clear<N, HS_THREADS>(sm, tid);
        }
        if (!eval) {
            for (int tid = 0; tid < HS_THREADS; ++tid) small_leaf<N, HS_THREADS>(p, sm, lanes[tid], tid, acc[tid]);
            for (int s = 1; s < lpf; s <<= 1)
                for (int tid = 0; tid < HS_THREADS; ++tid)
                    small_tree_level<N, HS_THREADS>(p, sm, lanes[tid], tid, s, acc[tid]);
            for (int tid = 0; tid < HS_THREADS; ++tid)
                small_backsub<N, HS_THREADS>(p, sm, lanes[tid], tid, status[tid], acc[tid]);
        }
        if (p.fast_mismatch) {
            std::vector<double> part(HS_THREADS * 6), tmp6(HS_THREADS * 6);
            for (int tid = 0; tid < HS_THREADS; ++tid) {
                double p6[6];
                small_fast_partials<N, HS_THREADS>(p, sm, lanes[tid], tid, acc[tid], p6);
                for (int q = 0; q < 6; ++q) part[tid * 6 + q] = p6[q];
            }
            for (int s = 1; s < lpf; s <<= 1) {   // the __shfl_xor butterfly
                for (int tid = 0; tid < HS_THREADS; ++tid)
                    for (int q = 0; q < 6; ++q) tmp6[tid * 6 + q] = part[tid * 6 + q] + part[(tid ^ s) * 6 + q];
                part.swap(tmp6);
            }
            for (int tid = 0; tid < HS_THREADS; ++tid) {
                const SmallLane &L = lanes[tid];
                if (L.fit >= 0 && L.lf == 0 && L.re > L.rb) {
                    double p6[6];
                    for (int q = 0; q < 6; ++q) p6[q] = part[tid * 6 + q];
                    small_fast_finalize(p, L, sm.ds[L.rb - sm.t_off + L.d_off], sm.ds[L.re - 1 - sm.t_off + L.d_off], p6, acc[tid].cn2,
                                        status[tid]);
                }
            }
            continue;
        }
        std::vector<double> sums(HS_THREADS * 4), tmp(HS_THREADS * 4);
        for (int tid = 0; tid < HS_THREADS; ++tid) {
            double s4[4];
            small_eval<N, HS_THREADS>(p, sm, lanes[tid], tid, s4);
            for (int q = 0; q < 4; ++q) sums[tid * 4 + q] = s4[q];
        }
        for (int s = 1; s < lpf; s <<= 1) {   // the __shfl_xor butterfly
            for (int tid = 0; tid < HS_THREADS; ++tid)
                for (int q = 0; q < 4; ++q) tmp[tid * 4 + q] = sums[tid * 4 + q] + sums[(tid ^ s) * 4 + q];
            sums.swap(tmp);
        }
        for (int tid = 0; tid < HS_THREADS; ++tid) {
            double s4[4];
            for (int q = 0; q < 4; ++q) s4[q] = sums[tid * 4 + q];
            small_finalize(p, lanes[tid], s4, status[tid]);
        }
    }
}

// K1 once more, this time the kernel function itself on an emulated CTA (staging of the window,
// table fill, __syncthreads, R-combine behind __syncwarp, butterflies), threads in either order.
template <int N>
static void run_small_cta(const qnmfit_batch *b, int lpf, bool eval, bool staged, int order)
{
    FitParams p;
    fill_params(b, lpf, eval, &p);
    const int fpc = HS_THREADS / lpf;
    const int ctas = (b->n_fits + fpc - 1) / fpc;
    const int rows = b->row_end_all - b->row_begin_all;
    p.stage_begin = b->row_begin_all;
    p.stage_rows = staged ? rows : 0;
    std::vector<unsigned char> smem(SmallSmem<N, HS_THREADS>::bytes(fpc, p.stage_rows) + 64);
    for (int cta = 0; cta < ctas; ++cta) {
        if (staged) hswarp::run_cta(HS_THREADS, [&](int) { fit_small_kernel<N, HS_THREADS, true>(p); }, order, cta, smem.data());
        else hswarp::run_cta(HS_THREADS, [&](int) { fit_small_kernel<N, HS_THREADS, false>(p); }, order, cta, smem.data());
    }
}

// eval: bit 0 eval-only, bit 1 staged window, bits 8.. thread order (hostsim_warp.h: 0 ascending,
// 1 descending, >= 2 seed of a shuffled order)
extern "C" int hostsim_fit_small_cta(const qnmfit_batch *b, int lpf, int eval)
{
    if (!b || b->struct_size != (int)sizeof(qnmfit_batch)) return QNMFIT_E_ABI;
    if (b->n_series != 1 || b->n_modes < 1 || b->n_modes > 8) return QNMFIT_E_SHAPE;
    if (lpf < 1 || lpf > 32 || (lpf & (lpf - 1))) return QNMFIT_E_SHAPE;
    const int order = eval >> 8;
    const bool staged = (eval & 2) != 0 && !b->series_index;
    const bool ev = (eval & 1) != 0;
    switch (b->n_modes) {
    case 1: run_small_cta<1>(b, lpf, ev, staged, order); break;
    case 2: run_small_cta<2>(b, lpf, ev, staged, order); break;
    case 3: run_small_cta<3>(b, lpf, ev, staged, order); break;
    case 4: run_small_cta<4>(b, lpf, ev, staged, order); break;
    case 5: run_small_cta<5>(b, lpf, ev, staged, order); break;
    case 6: run_small_cta<6>(b, lpf, ev, staged, order); break;
    case 7: run_small_cta<7>(b, lpf, ev, staged, order); break;
    case 8: run_small_cta<8>(b, lpf, ev, staged, order); break;
    }
    return 0;
}

// K1p: the CTA's shared memory is set up serially, then each warp runs in lock step.
#define HSP_THREADS 64
template <int N>
static void run_pair(const qnmfit_batch *b, int lpf, bool eval, int order)
{
    constexpr int CS = k1p_cs_ct(N), MB = k1p_mb_ct(N);
    typedef PairLayout<N, CS> LY;
    static_assert(LY::E == k1p_entries_ct(N, CS), "k1p_config.h mirrors PairLayout::E");
    FitParams p;
    fill_params(b, lpf, eval, &p);
    const int fpc = HSP_THREADS / lpf;
    const int ctas = (b->n_fits + fpc - 1) / fpc;
    std::vector<unsigned char> smem(PairSmem<N, CS, HSP_THREADS>::bytes(fpc, 0) + 64);
    for (int cta = 0; cta < ctas; ++cta) {
        PairSmem<N, CS, HSP_THREADS> sm;
        sm.carve(smem.data(), fpc, 0);
        sm.ts = p.times; sm.ds = p.data; sm.t_off = 0;
        for (int idx = 0; idx < fpc * N + CS; ++idx) pair_fill_tables<N, CS, HSP_THREADS>(p, sm, cta * fpc, fpc, idx);
        for (int e = 0; e < LY::E * HSP_THREADS; ++e) sm.R[e] = make_double2(0.0, 0.0);
        for (int warp = 0; warp < HSP_THREADS / 32; ++warp)
            hswarp::run_warp([&](int lane) {
                const int tid = warp * 32 + lane;
                const SmallLane L = pair_lane_setup<CS>(p, cta, tid, HSP_THREADS, true, MB, 0);
                pair_lane_body<N, CS, MB, HSP_THREADS>(p, sm, L, tid);
            }, order);
    }
}

// All pointers in *b are HOST pointers here.  Column counts: a sample of every configuration
// (2 / 4 / 8 lanes per row slice, 4 / 5 / 6 / 8 rows per block).
// eval: bit 0 = eval-only launch, bits 8.. thread order
extern "C" int hostsim_fit_pair(const qnmfit_batch *b, int lpf, int eval)
{
    const int order = eval >> 8;
    eval &= 1;
    if (!b || b->struct_size != (int)sizeof(qnmfit_batch)) return QNMFIT_E_ABI;
    if (b->n_series != 1) return QNMFIT_E_SHAPE;
    if (lpf < 1 || lpf > 32 || (lpf & (lpf - 1)) || lpf < k1p_cs_ct(b->n_modes)) return QNMFIT_E_SHAPE;
    switch (b->n_modes) {
    case 9: run_pair<9>(b, lpf, eval, order); break;
    case 11: run_pair<11>(b, lpf, eval, order); break;
    case 12: run_pair<12>(b, lpf, eval, order); break;
    case 14: run_pair<14>(b, lpf, eval, order); break;
    case 16: run_pair<16>(b, lpf, eval, order); break;
    case 19: run_pair<19>(b, lpf, eval, order); break;
    case 24: run_pair<24>(b, lpf, eval, order); break;
    default: return QNMFIT_E_SHAPE;
    }
    return 0;
}

// K3: the unmodified kernel function, one emulated CTA per fit (form G = 4 lanes per column,
// RPT = 16 rows per thread: the library's default).
extern "C" int hostsim_fit_struct(const qnmfit_batch *b, int eval)
{
    if (!b || b->struct_size != (int)sizeof(qnmfit_batch)) return QNMFIT_E_ABI;
    const int order = eval >> 8;
    eval &= 1;
    constexpr int G = 4, RPT = 16;
    const int N = b->n_modes, L = b->n_series;
    if (N < 1 || L < 1 || N + L > 64 || b->coef_rows || b->series_index) return QNMFIT_E_SHAPE;
    FitParams p;
    fill_params(b, G, eval != 0, &p);
    p.coef = (const double2 *)b->coef; p.coef_index = b->coef_index; p.n_coef = b->n_coef;
    p.omega_rows = (const double2 *)b->omega_rows;
    p.fast_mismatch = (!eval && b->uniform_weights && b->dt_nominal > 0.0 && !b->model) ? 1 : 0;
    const int threads = G * 32 * ((N + L + 31) / 32);
    std::vector<unsigned char> smem(Struct3Smem::bytes(N, L) + 64);
    for (int fit = 0; fit < b->n_fits; ++fit)
        hswarp::run_cta(threads, [&](int) { fit_struct3_kernel<G, RPT>(p); }, order, fit, smem.data());
    return 0;
}

// K2: the unmodified kernel function (streamed dense Householder; the only kernel that takes
// per-row mixing tables), tile geometry chosen as the library does.
extern "C" int hostsim_fit_general(const qnmfit_batch *b, int eval)
{
    if (!b || b->struct_size != (int)sizeof(qnmfit_batch)) return QNMFIT_E_ABI;
    const int order = eval >> 8;
    eval &= 1;
    const int N = b->n_modes, L = b->n_series;
    if (N < 1 || L < 1 || N > 64 || b->series_index) return QNMFIT_E_SHAPE;
    FitParams p;
    fill_params(b, K2_THREADS, eval != 0, &p);
    p.coef = (const double2 *)b->coef; p.coef_index = b->coef_index; p.n_coef = b->n_coef;
    p.omega_rows = (const double2 *)b->omega_rows; p.coef_rows = (const double2 *)b->coef_rows;
    p.fast_mismatch = 0;
    int TR = 128, TK = 0;
    for (; TR >= 32; TR /= 2)
        if (TR >= L) { TK = TR / L > 0 ? TR / L : 1; break; }
    if (!TK) return QNMFIT_E_SHAPE;
    std::vector<unsigned char> smem(GeneralSmem::bytes(N, L, TR, TK) + 64);
    for (int fit = 0; fit < b->n_fits; ++fit)
        hswarp::run_cta(K2_THREADS, [&](int) { fit_general_kernel(p, TR, TK); }, order, fit, smem.data());
    return 0;
}

// K4: the unmodified kernel function (blocked structured QR); its mma.sync.m8n8k4.f64 is emulated
// as a warp collective with the PTX fragment layout.
extern "C" int hostsim_fit_panel(const qnmfit_batch *b, int eval)
{
    if (!b || b->struct_size != (int)sizeof(qnmfit_batch)) return QNMFIT_E_ABI;
    const int order = eval >> 8;
    eval &= 1;
    const int N = b->n_modes, L = b->n_series;
    if (N < 1 || L < 1 || N > 64 || L > 64 || b->coef_rows || b->series_index) return QNMFIT_E_SHAPE;
    FitParams p;
    fill_params(b, 1, eval != 0, &p);
    p.coef = (const double2 *)b->coef; p.coef_index = b->coef_index; p.n_coef = b->n_coef;
    p.omega_rows = (const double2 *)b->omega_rows;
    p.fast_mismatch = (!eval && b->uniform_weights && b->dt_nominal > 0.0 && !b->model) ? 1 : 0;
    std::vector<unsigned char> smem(PanelSmem::bytes(N, L) + 64);
    for (int fit = 0; fit < b->n_fits; ++fit)
        hswarp::run_cta(K4_THREADS, [&](int) { fit_panel_kernel(p); }, order, fit, smem.data());
    return 0;
}

extern "C" int hostsim_sizeof_batch(void) { return (int)sizeof(qnmfit_batch); }
extern "C" int hostsim_sizeof_peers(void) { return (int)sizeof(qnmfit_peers); }

// All pointers in *b are HOST pointers here.
extern "C" int hostsim_fit_small(const qnmfit_batch *b, int lpf, int eval)
{
    if (!b || b->struct_size != (int)sizeof(qnmfit_batch)) return QNMFIT_E_ABI;
    if (b->n_series != 1 || b->n_modes < 1 || b->n_modes > QNMFIT_MAX_MODES_SMALL) return QNMFIT_E_SHAPE;
    if (lpf < 1 || lpf > 32 || (lpf & (lpf - 1))) return QNMFIT_E_SHAPE;
    switch (b->n_modes) {
    case 1: run<1>(b, lpf, eval); break;
    case 2: run<2>(b, lpf, eval); break;
    case 3: run<3>(b, lpf, eval); break;
    case 4: run<4>(b, lpf, eval); break;
    case 5: run<5>(b, lpf, eval); break;
    case 6: run<6>(b, lpf, eval); break;
    case 7: run<7>(b, lpf, eval); break;
    case 8: run<8>(b, lpf, eval); break;
    case 9: run<9>(b, lpf, eval); break;
    case 10: run<10>(b, lpf, eval); break;
    case 11: run<11>(b, lpf, eval); break;
    case 12: run<12>(b, lpf, eval); break;
    }
    return 0;
}
