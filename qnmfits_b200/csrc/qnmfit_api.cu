// qnmfit_api.cu — the C ABI of libqnmfit.so (see include/qnmfit.h).
//
// Build: nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -lineinfo
//        -Xcompiler -fPIC -shared -Iinclude -o qnmfits_b200/libqnmfit.so qnmfit_api.cu
#include <cuda_runtime.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>

#include <stdlib.h>

#include "qnmfit.h"
#include "kernels.h"

struct qnmfit_ctx {
    int k3_g, k3_rpt;        // K3: lanes per column, rows per thread (developer knob QNMFIT_K3G="G,RPT")
    int device;
    int sm_count;
    int smem_optin;          // max dynamic shared memory per block
    long long launches;
    cudaEvent_t h2d_event;        // recorded after the most recent qnmfit_h2d
    int h2d_pending;
    double *peer_scratch;         // device: flagged-fit counter of qnmfit_fit_batch_peers when the caller gives none
    unsigned char *stage_up, *stage_down;   // pinned staging of qnmfit_run_host
    size_t stage_up_bytes, stage_down_bytes;
    char err[512];
};

static char g_create_err[512] = "";

static int fail(qnmfit_ctx *ctx, int code, const char *fmt, ...)
{
    char *dst = ctx ? ctx->err : g_create_err;
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(dst, 512, fmt, ap);
    va_end(ap);
    return code;
}

static int cuda_fail(qnmfit_ctx *ctx, cudaError_t e, const char *what)
{
    return fail(ctx, (int)e, "%s: %s (%s)", what, cudaGetErrorName(e), cudaGetErrorString(e));
}

// ---------------------------------------------------------------------------
// context

extern "C" int qnmfit_abi_version(void) { return QNMFIT_ABI_VERSION; }

extern "C" int qnmfit_create(int device, qnmfit_ctx **out)
{
    if (!out) return fail(nullptr, QNMFIT_E_NULL, "qnmfit_create: out is NULL");
    *out = nullptr;
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0)
        return fail(nullptr, QNMFIT_E_NOGPU, "qnmfit_create: no CUDA device (%s)",
                    e != cudaSuccess ? cudaGetErrorString(e) : "count = 0");
    if (device < 0 || device >= count)
        return fail(nullptr, QNMFIT_E_SHAPE, "qnmfit_create: device %d out of range [0, %d)", device, count);
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, QNMFIT_E_NOGPU,
                    "qnmfit_create: device %d is sm_%d%d; this library contains sm_100a code only",
                    device, prop.major, prop.minor);
    qnmfit_ctx *ctx = new qnmfit_ctx();
    ctx->device = device;
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_optin = (int)prop.sharedMemPerBlockOptin;
    ctx->launches = 0;
    ctx->peer_scratch = nullptr;
    ctx->stage_up = ctx->stage_down = nullptr;
    ctx->stage_up_bytes = ctx->stage_down_bytes = 0;
    ctx->h2d_event = nullptr;
    ctx->h2d_pending = 0;
    ctx->err[0] = 0;
    ctx->k3_g = K3C_DEFAULT_G; ctx->k3_rpt = K3C_DEFAULT_RPT;
    { const char *eg = getenv("QNMFIT_K3G"); int g2 = 0, r2 = 0;
      if (eg && sscanf(eg, "%d,%d", &g2, &r2) == 2 && k3_kernel_ptr(g2, r2)) { ctx->k3_g = g2; ctx->k3_rpt = r2; } }
    if ((e = cudaSetDevice(device)) != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaSetDevice"); delete ctx; return r; }
    // opt every kernel in to the full shared-memory carve-out once
    for (int N = 1; N <= QNMFIT_MAX_MODES_SMALL; ++N)
        for (int st = 0; st < 4; ++st) {
            const int threads = st < 2 ? k1_threads(N) : k1_alt_threads(N);
            const void *fn = threads ? k1_kernel_ptr(N, threads, (st & 1) != 0) : nullptr;
            if (!fn) continue;
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
            if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaFuncSetAttribute(K1)"); delete ctx; return r; }
        }
    for (int N = K1P_MIN_N; N <= K1P_MAX_N; ++N)
        for (int st = 0; st < 2; ++st) {
            const void *fn = k1p_kernel_ptr(N, st != 0);
            if (!fn) continue;
            e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
            if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaFuncSetAttribute(K1p)"); delete ctx; return r; }
        }
    e = cudaFuncSetAttribute(k2_kernel_ptr(), cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
    if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaFuncSetAttribute(K2)"); delete ctx; return r; }
    e = cudaFuncSetAttribute(k4_kernel_ptr(), cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(k4_kernel_ptr(), cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaFuncSetAttribute(K4)"); delete ctx; return r; }
    for (int f = 0; f < K3_FORMS; ++f) {
        const void *k3 = k3_kernel_ptr(k3_forms[f][0], k3_forms[f][1]);
        e = cudaFuncSetAttribute(k3, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_optin);
        if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaFuncSetAttribute(K3c)"); delete ctx; return r; }
        // ~45 KB per fit: ask for the full shared-memory carve-out so that 4-5 fits share an SM
        // (the driver's default carve-out for a block of this size admits only two)
        e = cudaFuncSetAttribute(k3, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e != cudaSuccess) { int r = cuda_fail(nullptr, e, "cudaFuncSetAttribute(K3c carve-out)"); delete ctx; return r; }
    }
    *out = ctx;
    return 0;
}

extern "C" int qnmfit_destroy(qnmfit_ctx *ctx)
{
    if (ctx && ctx->peer_scratch) { cudaSetDevice(ctx->device); cudaFree(ctx->peer_scratch); }
    if (ctx && ctx->h2d_event) cudaEventDestroy(ctx->h2d_event);
    if (ctx && ctx->stage_up) cudaFreeHost(ctx->stage_up);
    if (ctx && ctx->stage_down) cudaFreeHost(ctx->stage_down);
    delete ctx;
    return 0;
}

extern "C" const char *qnmfit_last_error(const qnmfit_ctx *ctx) { return ctx ? ctx->err : g_create_err; }

extern "C" int64_t qnmfit_launch_count(const qnmfit_ctx *ctx) { return ctx ? ctx->launches : 0; }

extern "C" double qnmfit_flops_per_fit(int rows, int n_modes, int n_series, int fast_mismatch)
{
    const double M = (double)rows * (double)n_series, N = (double)n_modes;
    if (fast_mismatch)
        return 8.0 * M * N * N + 22.0 * M * N + 4.0 * M - (8.0 / 3.0) * N * N * N - 4.0 * N * N + 28.0 * N;
    return 8.0 * M * N * N + 30.0 * M * N + 20.0 * M - (8.0 / 3.0) * N * N * N - 4.0 * N * N;
}

// ---------------------------------------------------------------------------
// planning

struct Plan {
    int kernel;
    int lpf;
    int grid, block;
    size_t smem;
    bool staged;
    int stage_begin, stage_rows;
    int TR, TK;
};

static int validate(qnmfit_ctx *ctx, const qnmfit_batch *b, bool eval)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!b) return fail(ctx, QNMFIT_E_NULL, "batch is NULL");
    if (b->struct_size != (int32_t)sizeof(qnmfit_batch))
        return fail(ctx, QNMFIT_E_ABI, "struct_size %d != sizeof(qnmfit_batch) %d (ABI %d)", b->struct_size,
                    (int)sizeof(qnmfit_batch), QNMFIT_ABI_VERSION);
    if (b->n_fits < 0 || b->n_modes < 1 || b->n_modes > QNMFIT_MAX_MODES || b->n_series < 1 || b->n_times < 1)
        return fail(ctx, QNMFIT_E_SHAPE, "bad sizes: n_fits=%d n_modes=%d (1..%d) n_series=%d n_times=%d",
                    b->n_fits, b->n_modes, QNMFIT_MAX_MODES, b->n_series, b->n_times);
    if (b->series_index && (b->n_series != 1 || b->series_stride < b->n_times))
        return fail(ctx, QNMFIT_E_SHAPE, "series_index needs n_series == 1 and series_stride >= n_times");
    if (b->n_series > 1 && b->series_stride < b->n_times)
        return fail(ctx, QNMFIT_E_SHAPE, "series_stride %lld < n_times %d", (long long)b->series_stride, b->n_times);
    if (!b->times || !b->data || !b->mismatch) return fail(ctx, QNMFIT_E_NULL, "times, data and mismatch are required");
    if (eval && !b->C) return fail(ctx, QNMFIT_E_NULL, "qnmfit_eval_batch needs C");
    if (!b->omega && !b->omega_rows) {
        if (!b->omega_tilde || !b->mode_ptr || !b->inv_Mf)
            return fail(ctx, QNMFIT_E_NULL, "give omega, or omega_tilde + mode_ptr + inv_Mf");
        if (b->n_chi < 1 || b->n_mf < 1 || b->n_constituents < b->n_modes)
            return fail(ctx, QNMFIT_E_SHAPE, "bad table sizes: n_chi=%d n_mf=%d n_constituents=%d", b->n_chi, b->n_mf,
                        b->n_constituents);
        if ((!b->chi_index || !b->mf_index)
            && b->first_fit + b->n_fits > (int64_t)b->n_chi * b->n_mf)
            return fail(ctx, QNMFIT_E_SHAPE, "implicit grid indexing: first_fit + n_fits = %lld exceeds n_mf*n_chi = %lld",
                        (long long)(b->first_fit + b->n_fits), (long long)b->n_chi * b->n_mf);
    }
    if (b->model && b->model_stride < (int64_t)b->n_series * (b->row_end_all - b->row_begin_all))
        return fail(ctx, QNMFIT_E_SHAPE, "model_stride %lld too small for n_series * longest window", (long long)b->model_stride);
    if (b->coef && b->n_coef < 1) return fail(ctx, QNMFIT_E_SHAPE, "coef given but n_coef = %d", b->n_coef);
    if (b->row_begin_all < 0 || b->row_end_all > b->n_times || b->row_end_all <= b->row_begin_all)
        return fail(ctx, QNMFIT_E_WINDOW,
                    "window [%d, %d) (shared window, or union of the per-fit windows) is empty or outside [0, %d)",
                    b->row_begin_all, b->row_end_all, b->n_times);
    if ((b->row_begin == nullptr) != (b->row_end == nullptr))
        return fail(ctx, QNMFIT_E_NULL, "row_begin and row_end must both be given or both be NULL");
    if (b->anchor_rows < 0 || (b->anchor_rows % 4) != 0)
        return fail(ctx, QNMFIT_E_SHAPE, "anchor_rows %d must be a non-negative multiple of 4", b->anchor_rows);
    if (!(b->dt_nominal >= 0.0)) return fail(ctx, QNMFIT_E_SHAPE, "dt_nominal must be >= 0");
    if (b->flag_list && b->flag_capacity < 1)
        return fail(ctx, QNMFIT_E_SHAPE, "flag_list needs flag_capacity >= 1 (and flagged_count, its cursor)");
    return 0;
}

static int ilog2(int v) { int l = 0; while ((1 << (l + 1)) <= v) ++l; return l; }

static int make_plan(qnmfit_ctx *ctx, const qnmfit_batch *b, Plan *pl)
{
    memset(pl, 0, sizeof(*pl));
    int kernel = b->kernel;
    const bool small_ok = b->n_series == 1 && b->n_modes <= QNMFIT_MAX_MODES_SMALL && !b->coef
        && !b->omega_rows && !b->coef_rows;
    const bool pair_ok = b->n_series == 1 && b->n_modes >= K1P_MIN_N && b->n_modes <= K1P_MAX_N && !b->coef
        && !b->omega_rows && !b->coef_rows;
    const bool struct_ok = !b->coef_rows && b->n_modes + b->n_series <= 64
        && k3_smem_bytes(b->n_modes, b->n_series) <= (size_t)ctx->smem_optin;
    const bool panel_ok = !b->coef_rows && b->n_series <= 64
        && k4_smem_bytes(b->n_modes, b->n_series) <= (size_t)ctx->smem_optin;
    if (kernel == QNMFIT_KERNEL_AUTO) {
        // K3 where it applies (N + L <= 64) except for the large stacked shapes, K4 there and
        // beyond 64 columns (QNMFIT_AUTO_PANEL=1 / 0 forces K4 / K3 wherever both apply): K4's
        // trailing update runs on the tensor cores, but its panel factorisation only pays off
        // when the trailing matrix is wide
        const char *force = getenv("QNMFIT_AUTO_PANEL");
        // measured (profiles/k3_time_r02.json, k34_sweep_r02.json): K4 wins from about 36 columns
        // with several stacked series on (cfg4: 2.50 against 2.66 ms), K3 below
        const bool prefer_panel = force ? force[0] == '1' : (b->n_modes >= 36 && b->n_series >= 8);
        const char *no_pair = getenv("QNMFIT_AUTO_PAIR");     // "0": the choice before K1p existed
        const bool prefer_pair = pair_ok && b->n_modes >= K1P_AUTO_MIN_N && !(no_pair && no_pair[0] == '0');
        kernel = prefer_pair ? QNMFIT_KERNEL_PAIR
               : small_ok ? QNMFIT_KERNEL_SMALL
               : struct_ok && !(prefer_panel && panel_ok) ? QNMFIT_KERNEL_STRUCT
               : panel_ok ? QNMFIT_KERNEL_PANEL : QNMFIT_KERNEL_GENERAL;
    }
    if (kernel == QNMFIT_KERNEL_PANEL && !panel_ok)
        return fail(ctx, QNMFIT_E_SHAPE, "K4 needs no per-row coef table and a tile that fits shared memory (n_modes=%d, n_series=%d)",
                    b->n_modes, b->n_series);
    if (kernel == QNMFIT_KERNEL_STRUCT && !struct_ok)
        return fail(ctx, QNMFIT_E_SHAPE, "K3 needs n_modes + n_series <= 64 (got %d + %d) and no per-row coef table",
                    b->n_modes, b->n_series);
    if (kernel == QNMFIT_KERNEL_PAIR && !pair_ok)
        return fail(ctx, QNMFIT_E_SHAPE, "K1p needs n_series == 1, %d <= n_modes <= %d and no coef table",
                    K1P_MIN_N, K1P_MAX_N);
    if (kernel == QNMFIT_KERNEL_SMALL && !small_ok)
        return fail(ctx, QNMFIT_E_SHAPE, "K1 needs n_series == 1, n_modes <= %d and no coef table",
                    QNMFIT_MAX_MODES_SMALL);
    if (b->series_index && kernel != QNMFIT_KERNEL_SMALL && kernel != QNMFIT_KERNEL_PAIR)
        return fail(ctx, QNMFIT_E_SHAPE, "series_index is supported by K1 / K1p only (n_modes <= %d, no coef table)",
                    K1P_MAX_N);
    if (kernel != QNMFIT_KERNEL_SMALL && kernel != QNMFIT_KERNEL_GENERAL && kernel != QNMFIT_KERNEL_STRUCT
        && kernel != QNMFIT_KERNEL_PANEL && kernel != QNMFIT_KERNEL_PAIR)
        return fail(ctx, QNMFIT_E_SHAPE, "unknown kernel id %d", b->kernel);
    pl->kernel = kernel;
    const int N = b->n_modes;
    const int Mmax = b->row_end_all - b->row_begin_all;   // longest possible window
    if (kernel == QNMFIT_KERNEL_SMALL) {
        const int stage_rows = Mmax;
        // Lanes per fit are chosen for the WHOLE sweep (plan_fits) on this device's SM count,
        // never for the slab at hand: the split fixes the order in which a fit's partial
        // factors are combined, and a fit must get the same bits on any slab / GPU count.
        const int nplan = b->plan_fits > 0 ? b->plan_fits : b->n_fits;
        double best = 1e300;
        const char *force_lpf = getenv("QNMFIT_K1_LPF");       // developer knob (tools/slab_time.py)
        for (int lpf = 1; lpf <= 32; lpf *= 2) {
            if (force_lpf && atoi(force_lpf) > 0 && lpf != atoi(force_lpf)) continue;
            const int fpc = k1_threads(N) / lpf;
            // CTAs that may be co-resident on an SM by thread count; staging the window
            // must not reduce that (8 warps per SM are needed to keep the FP64 pipe fed)
#ifdef K1_FORCE_CPS   /* developer experiments (tools/k1_variants.py) */
            const int want_cps = K1_FORCE_CPS;
#else
            const int want_cps = 256 / K1_THREADS > 0 ? 256 / K1_THREADS : 1;
#endif
            const size_t per_cta = ((size_t)ctx->smem_optin + 1024) / want_cps - 1024;
            size_t smem = k1_smem_bytes(N, k1_threads(N), fpc, stage_rows);
            bool staged = b->series_index == nullptr;   // per-fit series are read through L1/L2
            if (!staged) smem = k1_smem_bytes(N, k1_threads(N), fpc, 0);
            if (staged && smem > per_cta) { staged = false; smem = k1_smem_bytes(N, k1_threads(N), fpc, 0); }
            if (smem > per_cta) continue;
#ifdef K1_FORCE_CPS
            if (smem < per_cta * 6 / 10) smem = per_cta * 6 / 10;   // pad so that no more CTAs become resident
#endif
            const int ctas = (b->n_fits + fpc - 1) / fpc;
            const int plan_ctas = (nplan + fpc - 1) / fpc;
            const int waves = (plan_ctas + ctx->sm_count * want_cps - 1) / (ctx->sm_count * want_cps);
            const int mb = k1_block_rows(N);
            const int rpl = ((Mmax + lpf - 1) / lpf + mb - 1) / mb;
            const double blocks = rpl * (1.0 + 0.2) /* second pass ~ 20% of a first-pass block */
                                + ilog2(lpf) * ((N + mb - 1) / mb);
            const double cost = (double)(waves > 0 ? waves : 1) * blocks * (staged ? 1.0 : 1.03);
            if (cost < best) {
                best = cost;
                pl->lpf = lpf; pl->grid = ctas; pl->block = k1_threads(N); pl->smem = smem; pl->staged = staged;
                pl->stage_begin = b->row_begin_all; pl->stage_rows = staged ? stage_rows : 0;
            }
        }
        if (best >= 1e300) return fail(ctx, QNMFIT_E_SHAPE, "K1: no lanes-per-fit choice fits shared memory");
        // Block size for THIS slab (it does not enter a fit's arithmetic, kernels.h): seven warps
        // per CTA when that saves a wave or fills more SMs.  A seven-warp CTA takes 0.936 of
        // the time of an eight-warp one (measured, profiles/slab_time_r02.json).
#ifndef K1_FORCE_CPS
        const char *no_alt = getenv("QNMFIT_K1_ALT_BLOCK");     // "0": developer A/B (tools/runs/r2_ab8.sh)
        const int alt = no_alt && no_alt[0] == '0' ? 0 : k1_alt_threads(N);
        if (alt > 0 && alt % pl->lpf == 0) {
            const int fpc_alt = alt / pl->lpf;
            const size_t smem_alt = k1_smem_bytes(N, alt, fpc_alt, pl->staged ? stage_rows : 0);
            const int ctas_alt = (b->n_fits + fpc_alt - 1) / fpc_alt;
            const int waves = (pl->grid + ctx->sm_count - 1) / ctx->sm_count;
            const int waves_alt = (ctas_alt + ctx->sm_count - 1) / ctx->sm_count;
            if (smem_alt <= (size_t)ctx->smem_optin && waves_alt * 0.936 < waves * 0.999) {
                pl->grid = ctas_alt; pl->block = alt; pl->smem = smem_alt;
            }
        }
#endif
    } else if (kernel == QNMFIT_KERNEL_PAIR) {
        // as K1: lanes per fit for the WHOLE sweep (plan_fits); a fit takes lpf / CS row slices
        const int CS = k1p_cs(N), mb = k1p_mb(N), threads = k1p_threads(N);
        const int nplan = b->plan_fits > 0 ? b->plan_fits : b->n_fits;
        const char *force_lpf = getenv("QNMFIT_K1_LPF");
        double best = 1e300;
        for (int lpf = CS; lpf <= 32; lpf *= 2) {
            if (force_lpf && atoi(force_lpf) > 0 && lpf != atoi(force_lpf)) continue;
            const int fpc = threads / lpf;
            bool staged = b->series_index == nullptr;
            size_t smem = k1p_smem_bytes(N, fpc, staged ? Mmax : 0);
            if (staged && smem > (size_t)ctx->smem_optin) { staged = false; smem = k1p_smem_bytes(N, fpc, 0); }
            if (smem > (size_t)ctx->smem_optin) continue;
            const int plan_ctas = (nplan + fpc - 1) / fpc;
            const int waves = (plan_ctas + ctx->sm_count - 1) / ctx->sm_count;
            const int groups = lpf / CS;
            const int rpl = ((Mmax + groups - 1) / groups + mb - 1) / mb;
            const double blocks = rpl + ilog2(groups) * ((N + mb - 1) / mb);
            const double cost = (double)(waves > 0 ? waves : 1) * blocks * (staged ? 1.0 : 1.03);
            if (cost < best) {
                best = cost;
                pl->lpf = lpf; pl->grid = (b->n_fits + fpc - 1) / fpc; pl->block = threads; pl->smem = smem;
                pl->staged = staged; pl->stage_begin = b->row_begin_all; pl->stage_rows = staged ? Mmax : 0;
            }
        }
        if (best >= 1e300) return fail(ctx, QNMFIT_E_SHAPE, "K1p: no lanes-per-fit choice fits shared memory");
    } else if (kernel == QNMFIT_KERNEL_PANEL) {
        pl->lpf = 1; pl->smem = k4_smem_bytes(b->n_modes, b->n_series);
        { const char *pad = getenv("QNMFIT_K4_ONE_PER_SM");     // developer knob: one fit per SM (contention studies)
          if (pad && pad[0] == '1' && pl->smem < (size_t)120 * 1024) pl->smem = (size_t)120 * 1024; }
        pl->grid = b->n_fits; pl->block = k4_threads();
    } else if (kernel == QNMFIT_KERNEL_STRUCT) {
        pl->lpf = ctx->k3_g; pl->TR = ctx->k3_g * ctx->k3_rpt;
        pl->smem = k3_smem_bytes(b->n_modes, b->n_series);
        pl->grid = b->n_fits; pl->block = ctx->k3_g * 32 * ((b->n_modes + b->n_series + 31) / 32);
    } else {
        const int L = b->n_series;
        int TR = 128;
        while (TR >= 32) {
            if (TR >= L) {
                const int TK = TR / L > 0 ? TR / L : 1;
                size_t smem = k2_smem_bytes(N, L, TR, TK);
                if (smem <= (size_t)ctx->smem_optin) {
                    pl->TR = TR; pl->TK = TK; pl->smem = smem;
                    break;
                }
            }
            TR /= 2;
        }
        if (pl->TR == 0)
            return fail(ctx, QNMFIT_E_SHAPE, "K2: n_modes=%d with n_series=%d does not fit shared memory", N, L);
        pl->grid = b->n_fits; pl->block = K2_THREADS; pl->lpf = K2_THREADS;
    }
    return 0;
}

static void fill_params(const qnmfit_batch *b, const Plan &pl, bool eval, FitParams *p)
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
    p->coef = (const double2 *)b->coef; p->coef_index = b->coef_index; p->n_coef = b->n_coef;
    p->series_index = b->series_index;
    p->omega_rows = (const double2 *)b->omega_rows; p->coef_rows = (const double2 *)b->coef_rows;
    p->anchor_rows = b->anchor_rows > 0 ? b->anchor_rows : QNMFIT_DEFAULT_ANCHOR_ROWS;
    p->dt_nominal = b->dt_nominal;
    p->C = (double2 *)b->C; p->mismatch = b->mismatch; p->residual = b->residual;
    p->R = (double2 *)b->R; p->status = b->status;
    p->model = (double2 *)b->model; p->model_stride = b->model_stride; p->omega_shared = b->omega_shared;
    p->flagged_count = b->flagged_count;
    p->flag_list = b->flag_list; p->flag_capacity = b->flag_list ? b->flag_capacity : 0;
    p->fit_index = b->fit_index;
    p->lanes_per_fit = pl.lpf; p->eval_only = eval ? 1 : 0;
    p->fast_mismatch = (pl.kernel != QNMFIT_KERNEL_GENERAL && !eval && b->uniform_weights && b->dt_nominal > 0.0 && !b->model) ? 1 : 0;
    p->stage_begin = pl.stage_begin; p->stage_rows = pl.stage_rows;
}

static int validate_peers(qnmfit_ctx *ctx, const qnmfit_batch *b, const qnmfit_peers *pe)
{
    if (pe->struct_size != (int32_t)sizeof(qnmfit_peers))
        return fail(ctx, QNMFIT_E_ABI, "qnmfit_peers.struct_size %d != %d", pe->struct_size, (int)sizeof(qnmfit_peers));
    if (pe->n_peers < 2 || pe->n_peers > QNMFIT_MAX_PEERS || pe->rank < 0 || pe->rank >= pe->n_peers)
        return fail(ctx, QNMFIT_E_PEER, "n_peers=%d (2..%d), rank=%d", pe->n_peers, QNMFIT_MAX_PEERS, pe->rank);
    if (pe->epoch < 1) return fail(ctx, QNMFIT_E_PEER, "epoch %lld must be >= 1", (long long)pe->epoch);
    for (int r = 0; r < pe->n_peers; ++r)
        if (!pe->mismatch[r] || !pe->flagged[r] || !pe->flags[r])
            return fail(ctx, QNMFIT_E_NULL, "peer %d: mismatch, flagged and flags are required", r);
    if (b->fit_index) return fail(ctx, QNMFIT_E_PEER, "fit_index launches (repairs) do not take part in the exchange");
    if (b->n_fits > 0 && b->mismatch != pe->mismatch[pe->rank] + b->first_fit)
        return fail(ctx, QNMFIT_E_PEER, "b->mismatch must be peers->mismatch[rank] + first_fit");
    return 0;
}

static int launch(qnmfit_ctx *ctx, const qnmfit_batch *b, void *stream, bool eval, const qnmfit_peers *pe = nullptr)
{
    int rc = validate(ctx, b, eval);
    if (rc) return rc;
    if (pe && (rc = validate_peers(ctx, b, pe))) return rc;
    if (b->n_fits == 0 && !pe) return 0;
    Plan pl;
    memset(&pl, 0, sizeof(pl));
    if (b->n_fits > 0 && (rc = make_plan(ctx, b, &pl))) return rc;
    FitParams p;
    fill_params(b, pl, eval, &p);
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    cudaStream_t st = (cudaStream_t)stream;
    if (pe) {
        if (!ctx->peer_scratch) {
            if ((e = cudaMalloc(&ctx->peer_scratch, 16)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(peer scratch)");
            if ((e = cudaMemset(ctx->peer_scratch, 0, 16)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMemset(peer scratch)");
        }
        p.n_peers = pe->n_peers; p.peer_rank = pe->rank;
        p.peer_epoch = (unsigned long long)pe->epoch; p.peer_timeout_ns = pe->timeout_ns;
        for (int r = 0; r < pe->n_peers; ++r) {
            p.peer_mismatch[r] = pe->mismatch[r]; p.peer_flagged[r] = pe->flagged[r];
            p.peer_flags[r] = (unsigned long long *)pe->flags[r];
        }
        // the launch's flagged count travels with the epoch: count it in the ctx's scratch
        // (reset by the barrier kernel) unless the caller supplied a counter
        if (!p.flagged_count) p.flagged_count = ctx->peer_scratch;
    }
    if (b->n_fits == 0) {
        // empty slab: only the barrier below
    } else if (pl.kernel == QNMFIT_KERNEL_SMALL) {
        e = k1_launch(b->n_modes, pl.staged, pl.grid, pl.block, pl.smem, st, p);
    } else if (pl.kernel == QNMFIT_KERNEL_PAIR) {
        e = k1p_launch(b->n_modes, pl.staged, pl.grid, pl.block, pl.smem, st, p);
    } else if (pl.kernel == QNMFIT_KERNEL_PANEL) {
        e = k4_launch(pl.grid, pl.smem, st, p);
    } else if (pl.kernel == QNMFIT_KERNEL_STRUCT) {
        e = k3_launch(ctx->k3_g, ctx->k3_rpt, pl.grid, pl.block, pl.smem, st, p);
    } else {
        e = k2_launch(pl.grid, pl.smem, st, p, pl.TR, pl.TK);
    }
    if (e != cudaSuccess) return cuda_fail(ctx, e, "kernel launch");
    if (b->n_fits > 0) ctx->launches += 1;
    if (pe) {
        if ((e = peer_barrier_launch(st, p)) != cudaSuccess) return cuda_fail(ctx, e, "peer barrier launch");
        ctx->launches += 1;
    }
    return 0;
}

extern "C" int qnmfit_fit_batch(qnmfit_ctx *ctx, const qnmfit_batch *b, void *stream)
{
    return launch(ctx, b, stream, false);
}

extern "C" int qnmfit_eval_batch(qnmfit_ctx *ctx, const qnmfit_batch *b, void *stream)
{
    return launch(ctx, b, stream, true);
}

extern "C" int qnmfit_fit_batch_peers(qnmfit_ctx *ctx, const qnmfit_batch *b, const qnmfit_peers *peers, void *stream)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!peers) return fail(ctx, QNMFIT_E_NULL, "peers is NULL");
    return launch(ctx, b, stream, false, peers);
}

// ---------------------------------------------------------------------------
// one sweep with host inputs and a host result in a single call

static int grow_pinned(qnmfit_ctx *ctx, unsigned char **buf, size_t *have, size_t need)
{
    if (*have >= need) return 0;
    if (*buf) cudaFreeHost(*buf);
    *buf = nullptr; *have = 0;
    size_t bytes = need + need / 4 + 4096;
    cudaError_t e = cudaMallocHost((void **)buf, bytes);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMallocHost(staging)");
    *have = bytes;
    return 0;
}

extern "C" int qnmfit_run_host(qnmfit_ctx *ctx, const qnmfit_batch *b, const qnmfit_peers *peers,
                               const qnmfit_copy *uploads, int n_uploads, const void *result_dev,
                               void *result_host, size_t result_bytes, int flags, void *stream)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!b) return fail(ctx, QNMFIT_E_NULL, "batch is NULL");
    if (n_uploads < 0 || (n_uploads > 0 && !uploads)) return fail(ctx, QNMFIT_E_NULL, "uploads is NULL");
    if (result_bytes > 0 && (!result_dev || !result_host)) return fail(ctx, QNMFIT_E_NULL, "result pointers are NULL");
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    int rc;
    // ---- uploads: staged through the ctx's pinned buffer (free again: every call ends with a
    // stream synchronisation).  Layout of the staging = layout of the device destinations.
    if ((flags & QNMFIT_RUN_NO_SYNC) && result_bytes > 0 && !(flags & QNMFIT_RUN_RESULT_PINNED))
        return fail(ctx, QNMFIT_E_SHAPE, "QNMFIT_RUN_NO_SYNC needs QNMFIT_RUN_RESULT_PINNED");
    if (n_uploads > 0) {
        const bool coalesce = (flags & QNMFIT_RUN_COALESCE) != 0;
        const bool pinned = (flags & QNMFIT_RUN_UPLOADS_PINNED) != 0;
        size_t total = 0;
        if (coalesce) {
            for (int i = 0; i < n_uploads; ++i) {
                if (!uploads[i].dst_dev || (uploads[i].bytes && !uploads[i].src_host))
                    return fail(ctx, QNMFIT_E_NULL, "upload %d: NULL pointer", i);
                if (i > 0 && (const char *)uploads[i].dst_dev < (const char *)uploads[i - 1].dst_dev + uploads[i - 1].bytes)
                    return fail(ctx, QNMFIT_E_SHAPE, "QNMFIT_RUN_COALESCE: destinations must ascend without overlap");
                if (pinned && (const char *)uploads[i].src_host - (const char *)uploads[0].src_host
                              != (const char *)uploads[i].dst_dev - (const char *)uploads[0].dst_dev)
                    return fail(ctx, QNMFIT_E_SHAPE, "QNMFIT_RUN_UPLOADS_PINNED | COALESCE: sources must be laid out like the destinations");
            }
            total = (size_t)((const char *)uploads[n_uploads - 1].dst_dev - (const char *)uploads[0].dst_dev)
                  + uploads[n_uploads - 1].bytes;
        } else {
            for (int i = 0; i < n_uploads; ++i) {
                if (uploads[i].bytes && (!uploads[i].dst_dev || !uploads[i].src_host))
                    return fail(ctx, QNMFIT_E_NULL, "upload %d: NULL pointer", i);
                total += (uploads[i].bytes + 255) / 256 * 256;
            }
        }
        if (!pinned && (rc = grow_pinned(ctx, &ctx->stage_up, &ctx->stage_up_bytes, total))) return rc;
        if (coalesce) {
            const void *src = uploads[0].src_host;
            if (!pinned) {
                const char *base = (const char *)uploads[0].dst_dev;
                for (int i = 0; i < n_uploads; ++i)
                    memcpy(ctx->stage_up + ((const char *)uploads[i].dst_dev - base), uploads[i].src_host, uploads[i].bytes);
                src = ctx->stage_up;
            }
            if ((e = cudaMemcpyAsync(uploads[0].dst_dev, src, total, cudaMemcpyHostToDevice, st)) != cudaSuccess)
                return cuda_fail(ctx, e, "cudaMemcpyAsync(H2D)");
        } else {
            size_t off = 0;
            for (int i = 0; i < n_uploads; ++i) {
                if (!uploads[i].bytes) continue;
                const void *src = uploads[i].src_host;
                if (!pinned) {
                    memcpy(ctx->stage_up + off, uploads[i].src_host, uploads[i].bytes);
                    src = ctx->stage_up + off;
                }
                if ((e = cudaMemcpyAsync(uploads[i].dst_dev, src, uploads[i].bytes,
                                         cudaMemcpyHostToDevice, st)) != cudaSuccess)
                    return cuda_fail(ctx, e, "cudaMemcpyAsync(H2D)");
                off += (uploads[i].bytes + 255) / 256 * 256;
            }
        }
    }
    if ((flags & QNMFIT_RUN_ZERO_COUNTER) && b->flagged_count) {
        if ((e = cudaMemsetAsync(b->flagged_count, 0, sizeof(double), st)) != cudaSuccess)
            return cuda_fail(ctx, e, "cudaMemsetAsync(counter)");
    }
    if ((rc = launch(ctx, b, stream, false, peers))) return rc;
    if (result_bytes > 0) {
        void *dst = result_host;
        if (!(flags & QNMFIT_RUN_RESULT_PINNED)) {
            if ((rc = grow_pinned(ctx, &ctx->stage_down, &ctx->stage_down_bytes, result_bytes))) return rc;
            dst = ctx->stage_down;
        }
        if ((e = cudaMemcpyAsync(dst, result_dev, result_bytes, cudaMemcpyDeviceToHost, st)) != cudaSuccess)
            return cuda_fail(ctx, e, "cudaMemcpyAsync(D2H)");
    }
    if (flags & QNMFIT_RUN_NO_SYNC) return 0;
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
    if (result_bytes > 0 && !(flags & QNMFIT_RUN_RESULT_PINNED)) memcpy(result_host, ctx->stage_down, result_bytes);
    return 0;
}

// ---------------------------------------------------------------------------
// stream-ordered transfers for the host wrapper

extern "C" int qnmfit_h2d(qnmfit_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, void *stream)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (bytes == 0) return 0;
    if (!dst_dev || !src_host) return fail(ctx, QNMFIT_E_NULL, "qnmfit_h2d: NULL pointer");
    cudaError_t e;
    // the event lives on the ctx's device; one process may drive several devices
    if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    if (!ctx->h2d_event && (e = cudaEventCreateWithFlags(&ctx->h2d_event, cudaEventDisableTiming)) != cudaSuccess)
        return cuda_fail(ctx, e, "cudaEventCreate");
    if ((e = cudaMemcpyAsync(dst_dev, src_host, bytes, cudaMemcpyHostToDevice, (cudaStream_t)stream)) != cudaSuccess)
        return cuda_fail(ctx, e, "cudaMemcpyAsync(H2D)");
    if ((e = cudaEventRecord(ctx->h2d_event, (cudaStream_t)stream)) != cudaSuccess) return cuda_fail(ctx, e, "cudaEventRecord");
    ctx->h2d_pending = 1;
    return 0;
}

extern "C" int qnmfit_h2d_wait(qnmfit_ctx *ctx)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!ctx->h2d_pending) return 0;
    ctx->h2d_pending = 0;
    cudaError_t e = cudaEventSynchronize(ctx->h2d_event);
    return e == cudaSuccess ? 0 : cuda_fail(ctx, e, "cudaEventSynchronize");
}

extern "C" int qnmfit_d2h(qnmfit_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, void *stream, int sync)
{
    if (!ctx) return QNMFIT_E_NULL;
    cudaError_t e = cudaSuccess;
    if (bytes > 0) {
        if (!dst_host || !src_dev) return fail(ctx, QNMFIT_E_NULL, "qnmfit_d2h: NULL pointer");
        e = cudaMemcpyAsync(dst_host, src_dev, bytes, cudaMemcpyDeviceToHost, (cudaStream_t)stream);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaMemcpyAsync(D2H)");
    }
    if (sync && (e = cudaStreamSynchronize((cudaStream_t)stream)) != cudaSuccess) return cuda_fail(ctx, e, "cudaStreamSynchronize");
    return 0;
}

extern "C" int qnmfit_zero(qnmfit_ctx *ctx, void *dst_dev, size_t bytes, void *stream)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (bytes == 0) return 0;
    if (!dst_dev) return fail(ctx, QNMFIT_E_NULL, "qnmfit_zero: NULL pointer");
    cudaError_t e = cudaMemsetAsync(dst_dev, 0, bytes, (cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(ctx, e, "cudaMemsetAsync");
}

extern "C" int qnmfit_stream_sync(qnmfit_ctx *ctx, void *stream)
{
    if (!ctx) return QNMFIT_E_NULL;
    cudaError_t e = cudaStreamSynchronize((cudaStream_t)stream);
    return e == cudaSuccess ? 0 : cuda_fail(ctx, e, "cudaStreamSynchronize");
}

// ---------------------------------------------------------------------------
// peer-mappable device memory (cudaIpc*: one process per GPU on one node)

extern "C" int qnmfit_peer_alloc(qnmfit_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64])
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!dptr || !handle || bytes == 0) return fail(ctx, QNMFIT_E_NULL, "qnmfit_peer_alloc: NULL argument or zero size");
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "handle size");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    void *ptr = nullptr;
    if ((e = cudaMalloc(&ptr, bytes)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc(peer window)");
    if ((e = cudaMemset(ptr, 0, bytes)) != cudaSuccess) { cudaFree(ptr); return cuda_fail(ctx, e, "cudaMemset(peer window)"); }
    if ((e = cudaDeviceSynchronize()) != cudaSuccess) { cudaFree(ptr); return cuda_fail(ctx, e, "cudaDeviceSynchronize"); }
    cudaIpcMemHandle_t h;
    if ((e = cudaIpcGetMemHandle(&h, ptr)) != cudaSuccess) { cudaFree(ptr); return cuda_fail(ctx, e, "cudaIpcGetMemHandle"); }
    memcpy(handle, &h, 64);
    *dptr = ptr;
    return 0;
}

extern "C" int qnmfit_peer_open(qnmfit_ctx *ctx, const unsigned char handle[64], void **dptr)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!dptr || !handle) return fail(ctx, QNMFIT_E_NULL, "qnmfit_peer_open: NULL argument");
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void *ptr = nullptr;
    if ((e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess)) != cudaSuccess)
        return cuda_fail(ctx, e, "cudaIpcOpenMemHandle");
    *dptr = ptr;
    return 0;
}

extern "C" int qnmfit_peer_close(qnmfit_ctx *ctx, void *dptr)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!dptr) return 0;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaIpcCloseMemHandle(dptr);
    return e == cudaSuccess ? 0 : cuda_fail(ctx, e, "cudaIpcCloseMemHandle");
}

extern "C" int qnmfit_peer_free(qnmfit_ctx *ctx, void *dptr)
{
    if (!ctx) return QNMFIT_E_NULL;
    if (!dptr) return 0;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e == cudaSuccess) e = cudaFree(dptr);
    return e == cudaSuccess ? 0 : cuda_fail(ctx, e, "cudaFree(peer window)");
}

extern "C" int qnmfit_plan_batch(qnmfit_ctx *ctx, const qnmfit_batch *b, qnmfit_plan *out)
{
    if (!out) return fail(ctx, QNMFIT_E_NULL, "plan is NULL");
    int rc = validate(ctx, b, false);
    if (rc) return rc;
    Plan pl;
    if ((rc = make_plan(ctx, b, &pl))) return rc;
    memset(out, 0, sizeof(*out));
    out->kernel = pl.kernel; out->lanes_per_fit = pl.lpf; out->grid = pl.grid; out->block = pl.block;
    out->smem_bytes = (int32_t)pl.smem; out->staged = pl.staged ? 1 : 0;
    out->fast_mismatch = (pl.kernel != QNMFIT_KERNEL_GENERAL && b->uniform_weights && b->dt_nominal > 0.0 && !b->model) ? 1 : 0;
    cudaFuncAttributes fa;
    const void *fn = pl.kernel == QNMFIT_KERNEL_SMALL ? k1_kernel_ptr(b->n_modes, pl.block, pl.staged)
                   : pl.kernel == QNMFIT_KERNEL_PAIR ? k1p_kernel_ptr(b->n_modes, pl.staged)
                   : pl.kernel == QNMFIT_KERNEL_STRUCT ? k3_kernel_ptr(ctx->k3_g, ctx->k3_rpt)
                   : pl.kernel == QNMFIT_KERNEL_PANEL ? k4_kernel_ptr() : k2_kernel_ptr();
    cudaError_t e = cudaFuncGetAttributes(&fa, fn);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaFuncGetAttributes");
    out->regs_per_thread = fa.numRegs;
    return 0;
}

// ---------------------------------------------------------------------------
// FP64 peak micro-benchmarks (kernels in misc_kernels.cu)

extern "C" int qnmfit_fp64_peak(qnmfit_ctx *ctx, int kind, int iters, double *tflops)
{
    if (!ctx || !tflops) return fail(ctx, QNMFIT_E_NULL, "qnmfit_fp64_peak: NULL argument");
    if (iters < 1) iters = 1;
    cudaError_t e = cudaSetDevice(ctx->device);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "cudaSetDevice");
    if (kind >= 20) {   // latency probes: one warp, n independent chains; *tflops receives CYCLES per instruction
        double *out = nullptr;
        if ((e = cudaMalloc(&out, 2 * sizeof(double))) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
        double host[2] = {0.0, 0.0};
        for (int rep = 0; rep < 2 && e == cudaSuccess; ++rep) {
            e = fp64_latency_launch(kind, out, iters);
            if (e == cudaSuccess) e = cudaMemcpy(host, out, sizeof(host), cudaMemcpyDeviceToHost);
        }
        cudaFree(out);
        if (e != cudaSuccess) return cuda_fail(ctx, e, "fp64 latency probe");
        *tflops = host[0];
        return 0;
    }
    // kinds >= 10: the same kernels at the occupancy of K1 (one 256-thread CTA per SM,
    // two warps per scheduler) to see what that occupancy can extract from the pipe
    const bool low_occ = kind >= 10;
    if (low_occ) kind -= 10;
    const int block = 256, grid = ctx->sm_count * (low_occ ? 1 : 8);
    double *out = nullptr;
    if ((e = cudaMalloc(&out, sizeof(double) * (size_t)grid * block)) != cudaSuccess) return cuda_fail(ctx, e, "cudaMalloc");
    cudaEvent_t ev0, ev1;
    cudaEventCreate(&ev0);
    cudaEventCreate(&ev1);
    float best_ms = 1e30f;
    for (int rep = 0; rep < 6; ++rep) {   // first reps warm up; best of the rest
        cudaEventRecord(ev0, 0);
        if ((e = fp64_peak_launch(kind, grid, block, out, iters)) != cudaSuccess) break;
        cudaEventRecord(ev1, 0);
        e = cudaEventSynchronize(ev1);
        if (e != cudaSuccess) break;
        float ms = 0.f;
        cudaEventElapsedTime(&ms, ev0, ev1);
        if (rep >= 2 && ms < best_ms) best_ms = ms;
        ctx->launches += 1;
    }
    cudaEventDestroy(ev0);
    cudaEventDestroy(ev1);
    cudaFree(out);
    if (e != cudaSuccess) return cuda_fail(ctx, e, "fp64 peak kernel");
    double flops;
    if (kind == 0 || kind == 2 || kind == 3) flops = 2.0 * 64.0 * (double)iters * (double)grid * block;   // 8 chains x 8 unroll
    else if (kind == 4) flops = (2.0 * 64.0 * block + 2.0 * 256.0 * 16.0 * (block / 32)) * (double)iters * (double)grid;
    else flops = 2.0 * 256.0 * 32.0 * (double)iters * (double)grid * (block / 32);        // 32 MMAs x 256 FMA
    *tflops = flops / (best_ms * 1e-3) * 1e-12;
    return 0;
}
