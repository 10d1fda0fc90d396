/*
 * qnmfit.h — C ABI of libqnmfit.so: batched complex least-squares ringdown fits on
 * one NVIDIA B200 (sm_100a).
 *
 * The reference (eliotfinch/qnmfits) has no FFI: its boundary is a set of Python
 * functions.  This library sits *underneath* that surface.  Each entry point names
 * the reference code it replaces (paths relative to the reference checkout):
 *
 *   qnmfit_fit_batch   one launch = B independent fits, each doing what the body of
 *                      ringdown_fit (qnmfits/qnmfits.py:274-293) or
 *                      multimode_ringdown_fit (qnmfits/qnmfits.py:606-652) does for
 *                      ONE grid point / start time: form the frequencies
 *                      (qnmfits/qnm.py:235,272-280; qnmfits.py:274), build the design
 *                      matrix a[k,j] = coef * exp(-i w_j (t_k - t0))
 *                      (qnmfits.py:280-283, :628-631), solve min ||a C - d||_2
 *                      (qnmfits.py:287, :635 -> LAPACK zgelsd), evaluate the model
 *                      (qnmfits.py:290, :639) and the trapezoid mismatch
 *                      (qnmfits.py:90-97, :123-139).  The B-loop replaces the serial
 *                      Python loops of mismatch_t0_array (qnmfits.py:1271-1281) and
 *                      mismatch_M_chi_grid (qnmfits.py:1391-1410).
 *   qnmfit_eval_batch  the model + mismatch half only, for caller-supplied
 *                      amplitudes (qnmfits.py:290-293, :639-652).
 *
 * All pointers in qnmfit_batch are DEVICE pointers (the Python wrapper passes
 * torch tensors' data_ptr()); complex128 arrays are interleaved (re, im) doubles,
 * layout-compatible with numpy complex128 and CUDA double2.  The caller owns every
 * buffer and must keep it alive until the stream has been synchronised.  No C++
 * types, exceptions or torch types cross this boundary.
 *
 * Return value of every call: 0 on success; QNMFIT_E_* (< 0) for argument errors;
 * > 0 is a cudaError_t passed through.  qnmfit_last_error() gives the message.
 * Numerical conditions of individual fits never fail a batch: they are reported
 * per fit in status[] (QNMFIT_ST_* bits).
 */
#ifndef QNMFIT_H
#define QNMFIT_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define QNMFIT_ABI_VERSION 7

/* limits of the compiled kernels */
#define QNMFIT_MAX_MODES_SMALL 12    /* register-resident TSQR kernel (K1): 4-row blocks
                                        up to 8 columns, 3-row blocks for 9-12          */
#define QNMFIT_DEFAULT_ANCHOR_ROWS 256 /* measured on B200: accuracy is flat from 16 to 512 rows */
#define QNMFIT_MAX_MODES 64          /* CTA-cooperative general kernel (K2)  */

/* argument errors */
#define QNMFIT_E_NULL      (-1)      /* required pointer is NULL                  */
#define QNMFIT_E_SHAPE     (-2)      /* sizes out of range / inconsistent         */
#define QNMFIT_E_WINDOW    (-3)      /* shared window empty or outside the series */
#define QNMFIT_E_ABI       (-4)      /* struct_size does not match this library   */
#define QNMFIT_E_NOGPU     (-5)      /* no usable sm_100 device                   */
#define QNMFIT_E_PEER      (-6)      /* bad peer-exchange descriptor              */

/* per-fit status bits */
#define QNMFIT_ST_OK            0
#define QNMFIT_ST_RANK_DEFICIENT 1   /* numpy.linalg.lstsq(rcond=None) MAY truncate a singular
                                        value (numpy/linalg/_linalg.py:2553): the device's
                                        estimate of s_min (inverse iteration on the triangle)
                                        is within 8x of eps*max(M,N)*||R||_F — amplitudes are
                                        the basic QR solution; the caller decides with the SVD
                                        of the exported R (qnmfit_common.cuh)                 */
#define QNMFIT_ST_NONFINITE     2    /* non-finite value met in inputs or outputs  */
#define QNMFIT_ST_UNDERDETERMINED 4  /* rows <= columns                            */

/* kernel selection (qnmfit_batch.kernel) */
#define QNMFIT_KERNEL_AUTO    0
#define QNMFIT_KERNEL_SMALL   1      /* K1: n_series == 1 and n_modes <= 12        */
#define QNMFIT_KERNEL_GENERAL 2      /* K2: any n_series, n_modes <= 64            */
#define QNMFIT_KERNEL_STRUCT  3      /* K3: n_modes + n_series <= 64, structured QR */
#define QNMFIT_KERNEL_PANEL   4      /* K4: structured QR, blocked (compact WY) with the
                                        trailing update on the FP64 tensor cores (DMMA);
                                        any n_modes <= 64 whose tile fits shared memory */
#define QNMFIT_KERNEL_PAIR    5      /* K1p: n_series == 1, 9 <= n_modes <= 24: K1 with the
                                        columns of a row slice split over 2, 4 or 8 lanes */
#define QNMFIT_MIN_MODES_PAIR 9
#define QNMFIT_MAX_MODES_PAIR 24

typedef struct qnmfit_ctx qnmfit_ctx;

typedef struct qnmfit_batch {
    int32_t struct_size;      /* sizeof(qnmfit_batch), checked                         */
    int32_t kernel;           /* QNMFIT_KERNEL_*                                       */

    /* ---- problem sizes ---- */
    int32_t n_fits;           /* B: fits in this launch                                */
    int32_t n_modes;          /* N: columns (QNMs)                                     */
    int32_t n_series;         /* L: stacked spherical-harmonic series (1 = single)     */
    int32_t n_times;          /* K_tot: samples of the full time series                */
    int64_t series_stride;    /* complex elements between consecutive series in data   */
    int64_t first_fit;        /* global flat index of fit 0 (index sharding over GPUs) */

    /* ---- shared inputs ---- */
    const double *times;      /* f64 [n_times], ascending                              */
    const double *data;       /* c128[n_series][series_stride]                         */

    /* ---- analysis window and start time: per fit, or shared when NULL ---- */
    const int32_t *row_begin; /* i32 [B] or NULL -> row_begin_all                      */
    const int32_t *row_end;   /* i32 [B] or NULL -> row_end_all   (half open)          */
    const double  *t0;        /* f64 [B] or NULL -> t0_all                             */
    int32_t row_begin_all;
    int32_t row_end_all;
    double  t0_all;

    /* ---- frequencies: explicit, or factored (table x 1/Mf x delta) ---- */
    const double  *omega;        /* c128[B][N] or NULL                                 */
    const double  *omega_tilde;  /* c128[n_chi][n_constituents]  (Mf = 1 values)       */
    const int32_t *mode_ptr;     /* i32 [N+1] constituent range of each mode           */
    const double  *inv_Mf;       /* f64 [n_mf]  1.0/Mf                                 */
    const double  *delta_factor; /* f64 [N] or NULL (= 1)                              */
    const int32_t *chi_index;    /* i32 [B] or NULL -> (first_fit + i) % n_chi         */
    const int32_t *mf_index;     /* i32 [B] or NULL -> (first_fit + i) / n_chi         */
    int32_t n_chi;
    int32_t n_mf;
    int32_t n_constituents;
    int32_t omega_shared;        /* 1: omega is c128[1][N], the same for every fit     */

    /* ---- per-series column coefficients (mu / identity); NULL = all ones ---- */
    const double  *coef;         /* c128[n_coef][L][N]                                 */
    const int32_t *coef_index;   /* i32 [B] or NULL -> the fit's chi index             */
    int32_t n_coef;

    /* ---- design-matrix generator ---- */
    int32_t anchor_rows;      /* direct cexp every this many rows (multiple of 4);
                                 0 -> QNMFIT_DEFAULT_ANCHOR_ROWS                      */
    double  dt_nominal;       /* > 0: nearly uniform grid, rows advance by the
                                 recurrence z *= exp(-i w dt) with first-order
                                 correction for the deviation of each sample from the
                                 nominal grid; 0: direct cexp for every element        */

    /* ---- amplitudes in (eval) / out (fit) ---- */
    double  *C;               /* c128[B][N]; may be NULL for fit                       */

    /* ---- outputs (each may be NULL except mismatch) ---- */
    double  *mismatch;        /* f64 [B]                                               */
    double  *residual;        /* f64 [B]  sum |model - data|^2                         */
    double  *R;               /* c128[B][N][N+1] upper-triangular factor, last column
                                 = Q^H d (for rank / singular values on the host)      */
    int32_t *status;          /* i32 [B] QNMFIT_ST_* bits                              */
    double  *model;           /* c128[B][model_stride]: best-fit model, series-major
                                 ((i, k) -> i*rows + k, rows = the fit's window length)  */
    int64_t model_stride;     /* complex elements between consecutive fits in model    */

    /* ---- mismatch quadrature ---- */
    int32_t uniform_weights;  /* 1: the caller guarantees that every step of the window
                                 deviates from dt_nominal by < 1e-11 relative, so the
                                 trapezoid weights are uniform to that level; K1 may then
                                 take the mismatch from by-products of the factorisation
                                 (||Q^H d||^2, ||d||^2, the two end rows) instead of a
                                 second pass over the rows.  0: always the general
                                 weighted second pass.                                  */
    int32_t plan_fits;        /* > 0: number of fits of the whole sweep this launch is a slab
                                 of.  K1 chooses how many lanes share a fit (and with it the
                                 order in which their partial factors are combined) from
                                 (plan_fits, longest window, n_modes) only, so that a fit
                                 gets the same bits whichever slab, rank or GPU count runs
                                 it.  0: n_fits.                                         */

    double  *flagged_count;   /* f64 [1] or NULL: incremented (atomicAdd) once per fit
                                 whose status word is non-zero; lets a sweep detect
                                 flagged fits without copying status[] back           */

    const int32_t *series_index; /* i32 [B] or NULL.  K1 only (n_series == 1): fit b reads
                                 its data from row series_index[b] of data[][] (row
                                 stride series_stride) instead of row 0 — one launch
                                 over many waveforms, as the batched free-frequency
                                 search needs (reference qnmfits.py:1905-2043 calls the
                                 fit once per waveform and optimiser step)             */

    /* ---- time-dependent spectrum (dynamic fits, reference qnmfits.py:318-475, 676-911):
       one frequency / mixing coefficient per ROW, shared by every fit of the batch.
       K3 takes omega_rows (single series or constant coef); coef_rows needs K2.       */
    const double  *omega_rows;   /* c128[N][n_times] or NULL: w_j at sample k; replaces
                                    omega / omega_tilde                                 */
    const double  *coef_rows;    /* c128[L][N][n_times] or NULL: mu_ij at sample k;
                                    replaces coef                                       */

    /* ---- flagged fits and their repair ---- */
    int32_t *flag_list;          /* i32 [2 * flag_capacity] or NULL: every fit whose status word
                                    is non-zero appends (fit, status) at the cursor given by
                                    the old value of *flagged_count (entries beyond the
                                    capacity are dropped; the count stays exact)          */
    int32_t flag_capacity;
    int32_t reserved2;
    const int32_t *fit_index;    /* i32 [B] or NULL: a launch over a SUBSET of a sweep's fits.
                                    Fit b of the launch reads every per-fit input (windows,
                                    t0, omega, chi / Mf / coef / series indices, the implicit
                                    grid index first_fit + .) under index fit_index[b] and
                                    stores its outputs at b.  With plan_fits of the sweep the
                                    arithmetic of a fit is that of the sweep's own launch —
                                    how the host repairs exactly the flagged fits          */
} qnmfit_batch;

/* Create / destroy a context bound to one CUDA device (one process per GPU). */
int qnmfit_create(int device, qnmfit_ctx **out);
int qnmfit_destroy(qnmfit_ctx *ctx);

/* Message for the most recent non-zero return on this ctx (never NULL).
 * ctx may be NULL: returns the message of the last failed qnmfit_create. */
const char *qnmfit_last_error(const qnmfit_ctx *ctx);

/* Launch the fits of *b asynchronously on `stream` (a cudaStream_t; NULL = legacy
 * default stream).  Replaces the loop bodies cited at the top of this file. */
int qnmfit_fit_batch(qnmfit_ctx *ctx, const qnmfit_batch *b, void *stream);

/* Model + mismatch (+ residual) only, with amplitudes read from b->C. */
int qnmfit_eval_batch(qnmfit_ctx *ctx, const qnmfit_batch *b, void *stream);

/* ---------------------------------------------------------------------------------
 * Multi-GPU result exchange fused into the fit kernels (one process per GPU).
 *
 * The reference's sweeps are serial loops over independent fits
 * (qnmfits/qnmfits.py:1271-1281, :1391-1410); here the flat fit index is split into one
 * slab per GPU, and every GPU needs the whole mismatch array at the end.  Instead of a
 * collective after the kernel, the thread that finishes a fit stores its mismatch
 * straight into the result array of EVERY peer GPU (peer-mapped device memory, posted
 * stores over NVLink / NVSwitch that overlap the remaining fits).  A one-warp barrier
 * kernel follows on the same stream: lane r publishes the launch's epoch (and count of
 * flagged fits) to peer r with system-scope release and waits for peer r's epoch.  When
 * it has completed on a GPU, that GPU holds the complete array.
 *
 *   qnmfit_peer_alloc   device allocation (zero-filled) that other processes of this
 *                       node can map, and its 64-byte handle (cudaIpcMemHandle_t);
 *   qnmfit_peer_open    map the allocation behind a handle received from another rank;
 *   qnmfit_peer_close / qnmfit_peer_free   undo the above;
 *   qnmfit_fit_batch_peers   qnmfit_fit_batch + the exchange.  Element first_fit + i of
 *                       mismatch[r] receives fit i for every r; b->mismatch must be
 *                       mismatch[rank] + first_fit (the local copy).  All ranks must
 *                       issue the launches of one epoch (SPMD); a rank with n_fits == 0
 *                       still calls (only the barrier kernel runs).
 */
#define QNMFIT_MAX_PEERS 8

typedef struct qnmfit_peers {
    int32_t  struct_size;                   /* sizeof(qnmfit_peers), checked                  */
    int32_t  n_peers;                       /* W: ranks incl. this one, 2..QNMFIT_MAX_PEERS   */
    int32_t  rank;                          /* this rank, 0..W-1                              */
    int32_t  reserved;
    int64_t  epoch;                         /* > every earlier epoch used with these flags;
                                               the same value on every rank                   */
    int64_t  timeout_ns;                    /* give up waiting for a peer after this long
                                               (<= 0: 30 s); the peer's slot of the local
                                               flagged[] array is then set to NaN             */
    double   *mismatch[QNMFIT_MAX_PEERS];   /* f64 [n_total] on every rank ([rank] = local)   */
    double   *flagged[QNMFIT_MAX_PEERS];    /* f64 [W] on every rank: element `rank` of each
                                               receives this launch's count of flagged fits   */
    uint64_t *flags[QNMFIT_MAX_PEERS];      /* u64 [W] on every rank: element `rank` of each
                                               is set to epoch once this launch's stores are
                                               visible system-wide                            */
} qnmfit_peers;

int qnmfit_peer_alloc(qnmfit_ctx *ctx, size_t bytes, void **dptr, unsigned char handle[64]);
int qnmfit_peer_open(qnmfit_ctx *ctx, const unsigned char handle[64], void **dptr);
int qnmfit_peer_close(qnmfit_ctx *ctx, void *dptr);
int qnmfit_peer_free(qnmfit_ctx *ctx, void *dptr);
int qnmfit_fit_batch_peers(qnmfit_ctx *ctx, const qnmfit_batch *b, const qnmfit_peers *peers, void *stream);

/* Stream-ordered transfers for the host wrapper, so that a sweep through the Python API
 * costs a handful of C calls (the reference pays none: its arrays never leave the host).
 *   qnmfit_h2d       cudaMemcpyAsync host -> device on `stream`; `src` should be pinned.
 *                    The ctx remembers the copy; qnmfit_h2d_wait blocks until the most
 *                    recent one has left the host buffer (so a staging buffer can be
 *                    refilled).
 *   qnmfit_d2h       cudaMemcpyAsync device -> host; sync != 0: also wait for the stream.
 *   qnmfit_zero      cudaMemsetAsync(dst, 0, bytes).
 *   qnmfit_stream_sync   cudaStreamSynchronize. */
int qnmfit_h2d(qnmfit_ctx *ctx, void *dst_dev, const void *src_host, size_t bytes, void *stream);
int qnmfit_h2d_wait(qnmfit_ctx *ctx);
int qnmfit_d2h(qnmfit_ctx *ctx, void *dst_host, const void *src_dev, size_t bytes, void *stream, int sync);
int qnmfit_zero(qnmfit_ctx *ctx, void *dst_dev, size_t bytes, void *stream);
int qnmfit_stream_sync(qnmfit_ctx *ctx, void *stream);

/* One sweep with HOST inputs and a HOST result in a single call — what the loops of
 * mismatch_M_chi_grid / mismatch_t0_array (reference qnmfits/qnmfits.py:1271-1281,
 * :1391-1410) cost through this library: copy `n_uploads` host arrays into their device
 * locations (staged through pinned memory owned by the ctx; QNMFIT_RUN_COALESCE: the
 * destinations ascend inside ONE device allocation and everything between them is padding,
 * so they travel in one cudaMemcpyAsync), zero *b->flagged_count (QNMFIT_RUN_ZERO_COUNTER),
 * launch the fits (with the peer exchange when `peers` is not NULL), copy `result_bytes`
 * from `result_dev` back to `result_host` (QNMFIT_RUN_RESULT_PINNED: `result_host` is
 * page-locked, copy straight into it) and wait for the stream.
 * One host thread driving several devices issues the call once per device with
 * QNMFIT_RUN_NO_SYNC (return as soon as everything is enqueued; needs a page-locked
 * `result_host`; the caller waits with qnmfit_stream_sync before it reads the result or uses
 * the ctx again) and QNMFIT_RUN_UPLOADS_PINNED (the `src_host` arrays are page-locked already
 * — one staging copy shared by all devices — and, with QNMFIT_RUN_COALESCE, laid out like
 * their destinations). */
typedef struct qnmfit_copy {
    void       *dst_dev;
    const void *src_host;
    size_t      bytes;
} qnmfit_copy;
#define QNMFIT_RUN_COALESCE      1
#define QNMFIT_RUN_ZERO_COUNTER  2
#define QNMFIT_RUN_RESULT_PINNED 4
#define QNMFIT_RUN_NO_SYNC       8
#define QNMFIT_RUN_UPLOADS_PINNED 16
int qnmfit_run_host(qnmfit_ctx *ctx, const qnmfit_batch *b, const qnmfit_peers *peers,
                    const qnmfit_copy *uploads, int n_uploads,
                    const void *result_dev, void *result_host, size_t result_bytes,
                    int flags, void *stream);

/* Number of kernels this ctx has launched so far (bench.py's gpu_launches). */
int64_t qnmfit_launch_count(const qnmfit_ctx *ctx);

/* What qnmfit_fit_batch would launch for *b: kernel id, lanes per fit, grid and
 * block size, dynamic shared memory bytes, registers per thread. */
typedef struct qnmfit_plan {
    int32_t kernel;
    int32_t lanes_per_fit;
    int32_t grid;
    int32_t block;
    int32_t smem_bytes;
    int32_t regs_per_thread;
    int32_t staged;           /* waveform window staged in shared memory */
    int32_t fast_mismatch;    /* mismatch from factorisation by-products (no 2nd pass) */
} qnmfit_plan;
int qnmfit_plan_batch(qnmfit_ctx *ctx, const qnmfit_batch *b, qnmfit_plan *plan);

/* FP64 peak micro-benchmarks on the ctx's device.  kind 0 = DFMA, operands mostly from
 * the reuse cache (vector pipe peak); 1 = DMMA m8n8k4 (tensor pipe); 2 = DFMA with three
 * distinct register operands per instruction (register-file read bound); 3 = DFMA with
 * two register reads + one reused operand; 4 = DFMA and DMMA interleaved in one warp.
 * Writes TFLOP/s (2 flops per FMA).  Kinds 20 + n / 30 + n (n = 1, 2, 4, 8) are latency
 * probes instead: ONE warp issuing DMMA / DFMA on n independent accumulators; they write
 * the measured CYCLES per instruction. */
int qnmfit_fp64_peak(qnmfit_ctx *ctx, int kind, int iters, double *tflops);

/* Algorithmic FP64 flops credited to one fit (DESIGN.md "flop accounting"), with
 * M = n_series * rows:
 *   fast_mismatch = 0:  F = 8 M N^2 + 30 M N + 20 M - (8/3) N^3 - 4 N^2
 *                       (QR, Q^H d, back-substitution, row generation, model, three
 *                        trapezoid inner products — SURVEY.md section 8d);
 *   fast_mismatch = 1:  F = 8 M N^2 + 22 M N + 4 M - (8/3) N^3 - 4 N^2 + 28 N
 *                       (no model pass: ||d||^2 and two end rows instead). */
double qnmfit_flops_per_fit(int rows, int n_modes, int n_series, int fast_mismatch);

/* ---- lock-step Nelder-Mead (host code; no device is touched) ---------------------------
 * The reference's free_frequency_fit (qnmfits/qnmfits.py:1992-2041) and calculate_epsilon
 * (:1520-1560) call scipy.optimize.minimize(method='Nelder-Mead', bounds=...) once per
 * waveform, one least-squares fit per objective call.  qnmfit_nm_* advances B such searches
 * together so that every optimiser step is ONE qnmfit_fit_batch launch: the caller loops
 *
 *     n = qnmfit_nm_step(nm, NULL, x, idx);
 *     while (n > 0) { f[0..n) = objective of point x[i*N .. i*N+N) of problem idx[i];
 *                     n = qnmfit_nm_step(nm, f, x, idx); }
 *
 * Each problem follows scipy's bounded Nelder-Mead (scipy 1.18 _minimize_neldermead,
 * non-adaptive coefficients, start simplex 5 % / 0.00025, clipping to the bounds, xatol and
 * fatol test, status 0 converged / 1 maxfun / 2 maxiter) operation for operation; idx is
 * ascending and holds a problem at most once per step.  `order`, when not NULL, is asked for
 * the sorted order of the simplices whose order is ambiguous (equal values or NaNs; all such
 * rows of a step in one call): numpy's argsort is not stable, so the Python wrapper passes
 * numpy's own.  x and idx
 * must hold n_problems points.  qnmfit_nm_step returns the number of points handed out, 0
 * when every search has ended, QNMFIT_E_* (< 0) on an argument error. */
#define QNMFIT_NM_MAX_VARS 64
typedef struct qnmfit_nm qnmfit_nm;
/* values: [n_rows][n] simplex values; order: [n_rows][n], to be filled with the permutation that
 * sorts each row ascending (NaNs last). */
typedef void (*qnmfit_nm_order_fn)(const double *values, int64_t n_rows, int n, int64_t *order, void *user);
int qnmfit_nm_create(int64_t n_problems, int n_vars, const double *x0, const double *lower,
                     const double *upper, double xatol, double fatol, double maxiter, double maxfun,
                     qnmfit_nm_order_fn order, void *order_user, qnmfit_nm **out);
int64_t qnmfit_nm_step(qnmfit_nm *nm, const double *f_prev, double *x_out, int64_t *idx_out);
/* Any output pointer may be NULL.  x: [n_problems][n_vars] best vertices; fun: their values. */
int qnmfit_nm_result(const qnmfit_nm *nm, double *x, double *fun, int64_t *nit, int64_t *nfev,
                     int64_t *status, int64_t *n_calls);
int qnmfit_nm_destroy(qnmfit_nm *nm);

int qnmfit_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif /* QNMFIT_H */
