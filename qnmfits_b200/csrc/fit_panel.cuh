// fit_panel.cuh — K4: the structured two-phase QR of K3 (fit_struct.cuh) as a BLOCKED
// Householder factorisation whose trailing update runs on the FP64 tensor cores.
//
// Replaces, for one start time / grid point, the body of the reference's
// multimode_ringdown_fit (qnmfits/qnmfits.py:606-652) and of ringdown_fit
// (qnmfits.py:274-293) when N > 12.  Same mathematics as K3:
//
//   phase 1   QR of [E | d_1 ... d_L]   (K rows, N pivot columns + L right-hand sides)
//   phase 2   QR of the stacked triangles [R_E D_i | Y_i], i = 1..L
//   then      back-substitution, residual, mismatch      (struct_finish, shared with K3)
//
// but a 64-row tile of the matrix lives in SHARED memory (column-major), and the N
// reflections of [R; tile] are applied panel by panel (8 columns, compact WY):
//
//   panel     warp 0 factors the 8 panel columns (two rows per lane in registers, column
//             norms and dot products by warp shuffles), and builds the 8 x 8 triangular T
//             of  H_7 ... H_0 = I - V T^H V^H  from the Gram matrix of the reflector tails,
//             whose entries ride along in the same shuffle reductions;
//   update    every warp takes 8-column chunks of the trailing matrix [E_trailing | d_1..d_L]:
//                 X = diag(v0) R_panel,chunk + V_b^H B_chunk      (8 x 64 x 8)   DMMA m8n8k4
//                 W = T^H X                                        (8 x 8 x 8)    DMMA
//                 R_panel,chunk -= diag(v0) W,   B_chunk -= V_b W  (64 x 8 x 8)   DMMA
//             — the genuinely dense contractions of this path (BASELINE.json north_star); a
//             complex product is four real MMAs.  One CTA barrier per panel phase instead of
//             one per reflection, and one LDS.128 feeds 128-256 FMAs instead of four.
//
// One CTA of 128 threads per fit, ~102 KB of shared memory at L = 21, N = 40: two fits per
// SM, so that the serial panel of one overlaps the tensor-core update of the other.
#pragma once
#include "qnmfit_common.cuh"
#include "fit_struct.cuh"

#define K4_THREADS 128
#define K4_WARPS (K4_THREADS / 32)
#define K4_M 64                   // tile rows
#define K4_NB 8                   // panel width
#define K4_S 65                   // column stride of the tile in complex elements (odd: see k4_update_chunk)
#define K4_XS 10                  // row stride of a warp's 8 x 8 exchange buffer

struct PanelSmem {
    PackedR R1;       // N rows, N+L columns: R_E | Y
    PackedR R2;       // N rows, N+1 columns: second factor | Q^H d  (lives in unused tile columns when it fits)
    double2 *tile;    // [NC][K4_S] column-major
    double2 *Tm;      // [2][8][8]  T of the panel being applied / being factored (upper triangular, zero below)
    double2 *gbuf;    // [8][8]  Gram of the reflector tails
    double2 *xw;      // [K4_WARPS][8][K4_XS]  fragment-layout exchange, one per warp
    double2 *om, *qq, *qw, *Cv;   // [N]
    double2 *scratch; // struct_finish: cc [L][N] then E [K3C_TK][N]   (overlays the tile)
    double *diag1, *diag2;        // [N]
    double *v0;       // [2][8]: leading reflector entries of the same two panels
    double *red;      // [16][8]
    double *ends;     // [64][3]
    // The same arrays as element offsets from the start of the CTA's shared memory (double2
    // units; o_diag*, o_v0 in doubles).  The hot code addresses shared memory as
    // k4_shared()[offset]: through the pointers above nvcc cannot prove the address space once
    // struct_finish is part of the kernel, and every tile access becomes a generic LD.E / ST.E.
    int o_R1, o_R2, o_tile, o_Tm, o_gbuf, o_xw, o_om, o_qq, o_qw, o_diag1, o_diag2, o_v0;

    // does R2 fit behind the N+1 columns phase 2 uses?
    __host__ __device__ static bool r2_in_tile(int N, int L)
    {
        return PackedR::entries(N, N) <= (size_t)(L - 1) * K4_S                      // behind column N of the tile
            && (size_t)L * N + (size_t)K3C_TK * N <= (size_t)(N + 1) * K4_S;        // and clear of struct_finish's scratch
    }
    __host__ __device__ static size_t bytes(int N, int L)
    {
        const size_t r1 = PackedR::entries(N, N + L - 1);
        const size_t r2 = r2_in_tile(N, L) ? 0 : PackedR::entries(N, N);
        size_t tile = (size_t)(N + L) * K4_S;
        const size_t scr = (size_t)L * N + (size_t)K3C_TK * N;
        if (tile < scr) tile = scr;
        return sizeof(double2) * (r1 + r2 + tile + 3 * 64 + (size_t)K4_WARPS * 8 * K4_XS + 4 * (size_t)N)
             + sizeof(double) * (2 * (size_t)N + 16 + 128 + 64 * 3);
    }
    __device__ void carve(void *base, int N, int L)
    {
        const size_t r1 = PackedR::entries(N, N + L - 1);
        size_t tl = (size_t)(N + L) * K4_S;
        const size_t scr = (size_t)L * N + (size_t)K3C_TK * N;
        if (tl < scr) tl = scr;
        double2 *p = (double2 *)base;
        R1.a = p; R1.W = N + L - 1; p += r1;
        tile = p; p += tl;
        scratch = tile;
        R2.W = N;
        if (r2_in_tile(N, L)) R2.a = tile + (size_t)(N + 1) * K4_S;
        else { R2.a = p; p += PackedR::entries(N, N); }
        Tm = p; p += 128; gbuf = p; p += 64;
        xw = p; p += K4_WARPS * 8 * K4_XS;
        om = p; p += N; qq = p; p += N; qw = p; p += N; Cv = p; p += N;
        double *d = (double *)p;
        diag1 = d; d += N;
        diag2 = d; d += N;
        v0 = d; d += 16;
        red = d; d += 128;
        ends = d;
        const double2 *b2 = (const double2 *)base;
        const double *b1 = (const double *)base;
        o_R1 = (int)(R1.a - b2); o_R2 = (int)(R2.a - b2); o_tile = (int)(tile - b2);
        o_Tm = (int)(Tm - b2); o_gbuf = (int)(gbuf - b2); o_xw = (int)(xw - b2);
        o_om = (int)(om - b2); o_qq = (int)(qq - b2); o_qw = (int)(qw - b2);
        o_diag1 = (int)(diag1 - b1); o_diag2 = (int)(diag2 - b1); o_v0 = (int)(v0 - b1);
    }
};

// the CTA's dynamic shared memory, typed: indexing THIS pointer keeps the accesses LDS / STS
__device__ __forceinline__ double2 *k4_shared()
{
    QF_DYN_SMEM(smem_raw);
    return (double2 *)smem_raw;
}

// entry (j, k), j < k, of a packed factor with W columns right of column 0 (PackedR::at)
__device__ __forceinline__ int k4_ridx(const int o_R, const int W, const int j, const int k)
{
    return o_R + j * W - (j * (j - 1)) / 2 + (k - j - 1);
}

// D (8 x 8) += A (8 x 4, row) * B (4 x 8, col), FP64 tensor core.  Fragments (PTX ISA, m8n8k4
// .f64): a = A[lane >> 2][lane & 3], b = B[lane & 3][lane >> 2], d0/d1 = D[lane >> 2][2 (lane & 3) + 0/1].
__device__ __forceinline__ void k4_dmma(double &d0, double &d1, const double a, const double b)
{
#ifndef QNMFIT_HOSTSIM
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
#else
    hs_dmma_m8n8k4(d0, d1, a, b);      // tests/hostsim/hostsim_warp.h: the same fragments, on the host
#endif
}

__device__ __forceinline__ double2 k4_cmul(const double2 a, const double2 b)
{
    return make_double2(fma(a.x, b.x, -(a.y * b.y)), fma(a.x, b.y, a.y * b.x));
}

#ifdef K4_TRACE
// developer build (tools/build_variant.sh k4trace -DK4_TRACE): per-warp clock stamps of CTA 0's
// pipeline, read back with qnmfit_debug_trace (tools/k4_trace.py)
__device__ long long k4_trace_buf[8192];
__device__ int k4_trace_n;
#define K4_STAMP(tag)                                                                                     \
    do {                                                                                                  \
        if (blockIdx.x == 0 && lane == 0) {                                                               \
            const int slot_ = atomicAdd(&k4_trace_n, 1);                                                  \
            if (slot_ < 2048) { k4_trace_buf[4 * slot_] = clock64(); k4_trace_buf[4 * slot_ + 1] = warp;  \
                                k4_trace_buf[4 * slot_ + 2] = (tag); k4_trace_buf[4 * slot_ + 3] = pi; }   \
        }                                                                                                 \
    } while (0)
#ifndef K4_TRACE_PANEL
#define K4_PSTAMP(tag, jj) do { } while (0)
#else
#define K4_PSTAMP(tag, jj)                                                                                \
    do {                                                                                                  \
        if (blockIdx.x == 0 && lane == 0) {                                                               \
            const int slot_ = atomicAdd(&k4_trace_n, 1);                                                  \
            if (slot_ < 2048) { k4_trace_buf[4 * slot_] = clock64(); k4_trace_buf[4 * slot_ + 1] = 100;   \
                                k4_trace_buf[4 * slot_ + 2] = (tag); k4_trace_buf[4 * slot_ + 3] = (jj); } \
        }                                                                                                 \
    } while (0)
#endif
#else
#define K4_STAMP(tag) do { } while (0)
#define K4_PSTAMP(tag, jj) do { } while (0)
#endif

// sum over the 4 lanes that share a panel column (every lane of the warp must call it)
__device__ __forceinline__ double k4_sum4(double v)
{
    v += __shfl_xor_sync(0xffffffffu, v, 2);
    v += __shfl_xor_sync(0xffffffffu, v, 1);
    return v;
}

// ---------------------------------------------------------------------------------------
// Panel factorisation by ONE warp: reflections first..w-1 of the panel starting at column j0
// (columns before `first` are structurally zero in this tile: phase 2).  Lane (c = lane / 4,
// q = lane % 4) owns rows q + 4 a, a = 0..15, of panel column c in registers: a column's dot
// product needs two shuffle levels over its four lanes, and the whole warp runs the same
// instruction stream (eight columns side by side).  Reflection jj:
//   * every column keeps |its rows|^2 current (accumulated by the update pass); column jj's
//     lanes publish their rows — the final reflector tail b_jj — to the tile, and every lane
//     forms the reflector scalars from the broadcast norm;
//   * every column forms b_jj^H x over its rows: later columns update their entry of row jj of
//     the factor and their rows; earlier columns (final tails) deliver the Gram entry g_c,jj =
//     b_c^H b_jj that T needs (formed from the final tails: building it from by-products of
//     the reflections loses ||B_k|| / ||b_k|| digits on ill-conditioned overtone columns);
//   * lanes 0..7 extend row `lane` of T by column jj:  T_ii = beta_i,
//     T_i,jj = -beta_jj sum_{l<jj} T_il g_l,jj.
// On return the tile's panel columns hold the tails (the V_b of the update), the factor
// (o_R, RW, o_diag) the panel's rows within the panel, v0[buf] / Tm[buf] the leading reflector
// entries and T.  Two earlier forms are in profiles/README.md: rows over lanes with all eight
// columns per lane (15-value five-level butterflies: issue-bound on one scheduler, 16.5 k
// cycles per panel) and the whole CTA with one barrier per reflection (10 k cycles, and no
// warp left to apply the previous panel meanwhile).
__device__ __forceinline__ void k4_panel(const PanelSmem &sm, const int o_R, const int RW, const int o_diag,
                                         const int j0, const int w, const int first, const int lane, const int buf)
{
    double2 *const S2 = k4_shared();
    double *const S1 = (double *)S2;
    const int c = lane >> 2, q = lane & 3;
    const bool in_panel = c < w;
    const int colbase = sm.o_tile + (j0 + (in_panel ? c : 0)) * K4_S + q;
    double2 x[16];
    double part0 = 0.0, part1 = 0.0;
#pragma unroll
    for (int a = 0; a < 16; ++a) {
        const double2 v = S2[colbase + 4 * a];
        x[a] = in_panel ? v : make_double2(0.0, 0.0);
    }
#pragma unroll
    for (int a = 0; a < 16; a += 2) {
        part0 = fma(x[a].x, x[a].x, part0);         part1 = fma(x[a + 1].x, x[a + 1].x, part1);
        part0 = fma(x[a].y, x[a].y, part0);         part1 = fma(x[a + 1].y, x[a + 1].y, part1);
    }
    double part = part0 + part1;
    double betas[K4_NB];                          // beta of the panel's reflections (0: absent)
#pragma unroll
    for (int kk = 0; kk < K4_NB; ++kk) betas[kk] = 0.0;
    S2[sm.o_gbuf + lane] = make_double2(0.0, 0.0);
    S2[sm.o_gbuf + lane + 32] = make_double2(0.0, 0.0);
    __syncwarp();

#pragma unroll
    for (int jj = 0; jj < K4_NB; ++jj) {
        double v0 = 0.0;
        if (jj >= first && jj < w) {             // warp-uniform
            const int j = j0 + jj;
            // One barrier per reflection (the tail must be in the tile before it is read); everything
            // behind it is ONE basic block, so that the scalar chain (rsqrt, rcp: ~130 cycles of
            // dependent operations), the dot products and the update interleave — a lone in-order
            // warp stalls on every chain the compiler cannot overlap.
            const double tot = __shfl_sync(0xffffffffu, k4_sum4(part), 4 * jj);     // |b_jj|^2
            if (c == jj) {
#pragma unroll
                for (int a = 0; a < 16; ++a) S2[colbase + 4 * a] = x[a];             // the final tail b_jj
            }
            const bool later = c > jj && in_panel;
            const int ri = k4_ridx(o_R, RW, j, j0 + (later ? c : jj + 1));
            const double r = S1[o_diag + j];
            const double2 old = S2[ri];
            __syncwarp();
            // b_jj^H x over this lane's rows (sixteen independent accumulators)
            const int bbase = sm.o_tile + j * K4_S + q;
            double2 b[16];
            double s_xx[4] = {0.0, 0.0, 0.0, 0.0}, s_yy[4] = {0.0, 0.0, 0.0, 0.0};
            double s_xy[4] = {0.0, 0.0, 0.0, 0.0}, s_yx[4] = {0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int a = 0; a < 16; ++a) {
                b[a] = S2[bbase + 4 * a];
                s_xx[a & 3] = fma(b[a].x, x[a].x, s_xx[a & 3]);
                s_yy[a & 3] = fma(b[a].y, x[a].y, s_yy[a & 3]);
                s_xy[a & 3] = fma(b[a].x, x[a].y, s_xy[a & 3]);
                s_yx[a & 3] = fma(b[a].y, x[a].x, s_yx[a & 3]);
            }
            // reflector scalars (every lane the same)
            const double t = fma(r, r, tot) + 1e-300;        // fit_small.cuh: an all-zero column needs no branch
            const double y = qf_rsqrt(t);
            const double nrm = t * y;
            const double ar = fabs(r);
            v0 = copysign(ar + nrm, r);
            const double beta = qf_rcp(nrm * (ar + nrm));
            const double dr = k4_sum4(((s_xx[0] + s_yy[0]) + (s_xx[1] + s_yy[1])) + ((s_xx[2] + s_yy[2]) + (s_xx[3] + s_yy[3])));
            const double di = k4_sum4(((s_xy[0] - s_yx[0]) + (s_xy[1] - s_yx[1])) + ((s_xy[2] - s_yx[2]) + (s_xy[3] - s_yx[3])));
            betas[jj] = beta;
            if (c < jj && q == 0) S2[sm.o_gbuf + c * 8 + jj] = make_double2(dr, -di);     // g_c,jj = b_c^H b_jj
            // the panel's row of R and the rank-1 update of the later columns
            const double pr = fma(v0, old.x, dr) * beta, pi = fma(v0, old.y, di) * beta;
            if (later) {
                if (q == 0) S2[ri] = make_double2(fma(-v0, pr, old.x), fma(-v0, pi, old.y));
                double pn[8] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                for (int a = 0; a < 16; ++a) {
                    double bx = x[a].x, by = x[a].y;
                    bx = fma(-pr, b[a].x, bx);
                    by = fma(-pr, b[a].y, by);
                    bx = fma(pi, b[a].y, bx);
                    by = fma(-pi, b[a].x, by);
                    x[a] = make_double2(bx, by);
                    pn[a & 3] = fma(bx, bx, pn[a & 3]);
                    pn[4 + (a & 3)] = fma(by, by, pn[4 + (a & 3)]);
                }
                part = ((pn[0] + pn[4]) + (pn[1] + pn[5])) + ((pn[2] + pn[6]) + (pn[3] + pn[7]));
            }
            if (lane == 0) S1[o_diag + j] = -copysign(nrm, r);
        }
        if (lane == 0) S1[sm.o_v0 + buf * 8 + jj] = v0;
    }
    // the tails are in the tile (the V_b of the update).  T from the Gram of the tails (strict
    // upper part in gbuf), column j by back-substitution on e_j — every column independent of the
    // others, lane j (< 8) builds column j in registers, all lanes run the same predicated code:
    //   T_jj = beta_j,   T_ij = -beta_i sum_{l = i+1..j} g_il T_lj   (i = j-1 .. 0)
    // (columns of absent reflections come out zero: beta = 0).
    __syncwarp();
    {
        const int jc = lane & 7;
        double2 Tc[K4_NB];
#pragma unroll
        for (int i = K4_NB - 1; i >= 0; --i) {
            double2 acc = make_double2(0.0, 0.0);
#pragma unroll
            for (int l = i + 1; l < K4_NB; ++l) {
                const double2 g = S2[sm.o_gbuf + i * 8 + l];
                acc.x = fma(g.x, Tc[l].x, acc.x); acc.x = fma(-g.y, Tc[l].y, acc.x);
                acc.y = fma(g.x, Tc[l].y, acc.y); acc.y = fma(g.y, Tc[l].x, acc.y);
            }
            const double bi = betas[i];
            Tc[i] = i > jc ? make_double2(0.0, 0.0) : i == jc ? make_double2(bi, 0.0) : make_double2(-bi * acc.x, -bi * acc.y);
        }
        if (lane < K4_NB) {
#pragma unroll
            for (int i = 0; i < K4_NB; ++i) S2[sm.o_Tm + buf * 64 + i * 8 + jc] = Tc[i];
        }
    }
}

// ---------------------------------------------------------------------------------------
// Trailing update of one 8-column chunk [c0, c0 + 8) by one warp (columns >= ncols are
// padding: read as zero, never stored).  Vr / Vi: the warp's A fragments of V_b^H for the 16
// k-steps of the tile's 64 rows (thread (g, t): V_b[4 s + t][g]).  Shared-memory access
// patterns with the odd column stride K4_S: the C fragments of the second product are
// conflict-free, its A fragments and the B fragments of the first product are two-way.
__device__ __forceinline__ void k4_update_chunk(const PanelSmem &sm, const int o_R, const int RW, const int j0,
                                                const int w, const int first, const int c0, const int ncols,
                                                const double (&Vr)[16], const double (&Vi)[16], const int o_xw,
                                                const int lane, const int buf, const int r_begin = 0,
                                                const int r_end = K4_M / 8, const bool write_R = true)
{
    double2 *const S2 = k4_shared();
    double *const S1 = (double *)S2;
    const int g = lane >> 2, t = lane & 3;
    const int col0 = c0 + 2 * t, col1 = col0 + 1;
    const bool ok0 = col0 < ncols, ok1 = col1 < ncols;
    const bool row_ok = g < w && g >= first;                  // reflector g exists
    const double v0g = S1[sm.o_v0 + buf * 8 + g];              // zero for absent reflectors
    const int r0i = k4_ridx(o_R, RW, j0 + (row_ok ? g : 0), ok0 ? col0 : c0);
    const int r1i = k4_ridx(o_R, RW, j0 + (row_ok ? g : 0), ok1 ? col1 : c0);
    // ---- X = diag(v0) R_panel,chunk + V_b^H B_chunk
    double xr0 = 0.0, xr1 = 0.0, xi0 = 0.0, xi1 = 0.0;         // even k-steps (and the R term)
    double yr0 = 0.0, yr1 = 0.0, yi0 = 0.0, yi1 = 0.0;         // odd k-steps
    {
        const double2 ra = S2[r0i], rb = S2[r1i];
        if (row_ok && ok0) { xr0 = v0g * ra.x; xi0 = v0g * ra.y; }
        if (row_ok && ok1) { xr1 = v0g * rb.x; xi1 = v0g * rb.y; }
    }
    {
        const int cb = c0 + g;
        const bool okb = cb < ncols;
        const int bcol = sm.o_tile + (okb ? cb : c0) * K4_S + t;
#pragma unroll
        for (int s = 0; s < 16; s += 2) {
            double2 b0 = S2[bcol + 4 * s], b1 = S2[bcol + 4 * s + 4];
            if (!okb) { b0 = make_double2(0.0, 0.0); b1 = b0; }
            // X += conj(V)^T B:  Xr += Vr Br + Vi Bi,  Xi += Vr Bi - Vi Br
            k4_dmma(xr0, xr1, Vr[s], b0.x);         k4_dmma(yr0, yr1, Vr[s + 1], b1.x);
            k4_dmma(xi0, xi1, Vr[s], b0.y);         k4_dmma(yi0, yi1, Vr[s + 1], b1.y);
            k4_dmma(xr0, xr1, Vi[s], b0.y);         k4_dmma(yr0, yr1, Vi[s + 1], b1.y);
            k4_dmma(xi0, xi1, Vi[s], -b0.x);        k4_dmma(yi0, yi1, Vi[s + 1], -b1.x);
        }
    }
    xr0 += yr0; xr1 += yr1; xi0 += yi0; xi1 += yi1;
    // ---- W = T^H X: X goes from the accumulator layout to the B layout through the warp's buffer
    __syncwarp();
    S2[o_xw + g * K4_XS + 2 * t] = make_double2(xr0, xi0);
    S2[o_xw + g * K4_XS + 2 * t + 1] = make_double2(xr1, xi1);
    __syncwarp();
    double wr0 = 0.0, wr1 = 0.0, wi0 = 0.0, wi1 = 0.0;
#pragma unroll
    for (int s = 0; s < 2; ++s) {
        const double2 xb = S2[o_xw + (4 * s + t) * K4_XS + g];     // X[4 s + t][g]
        const double2 tt = S2[sm.o_Tm + buf * 64 + (4 * s + t) * 8 + g];   // T^H[g][4 s + t] = conj(T[4 s + t][g])
        // W += conj(T)^T X:  Wr += Tr Xr + Ti Xi,  Wi += Tr Xi - Ti Xr
        k4_dmma(wr0, wr1, tt.x, xb.x);
        k4_dmma(wi0, wi1, tt.x, xb.y);
        k4_dmma(wr0, wr1, tt.y, xb.y);
        k4_dmma(wi0, wi1, tt.y, -xb.x);
    }
    // ---- the panel's rows of R (one warp only when several share the chunk by row tiles)
    if (row_ok && write_R) {
        if (ok0) { const double2 r = S2[r0i]; S2[r0i] = make_double2(fma(-v0g, wr0, r.x), fma(-v0g, wi0, r.y)); }
        if (ok1) { const double2 r = S2[r1i]; S2[r1i] = make_double2(fma(-v0g, wr1, r.x), fma(-v0g, wi1, r.y)); }
    }
    // ---- B_chunk -= V_b W: -W in the B layout
    __syncwarp();
    S2[o_xw + g * K4_XS + 2 * t] = make_double2(-wr0, -wi0);
    S2[o_xw + g * K4_XS + 2 * t + 1] = make_double2(-wr1, -wi1);
    __syncwarp();
    const double2 n0 = S2[o_xw + t * K4_XS + g], n1 = S2[o_xw + (4 + t) * K4_XS + g];     // -W[t][g], -W[4 + t][g]
    const bool va0 = t < w && t >= first, va1 = 4 + t < w && 4 + t >= first;    // panel columns that are reflectors
    const int vcol0 = sm.o_tile + (j0 + (va0 ? t : 0)) * K4_S + g;
    const int vcol1 = sm.o_tile + (j0 + (va1 ? 4 + t : 0)) * K4_S + g;
    const int ccol0 = sm.o_tile + (ok0 ? col0 : c0) * K4_S + g;
    const int ccol1 = sm.o_tile + (ok1 ? col1 : c0) * K4_S + g;
#pragma unroll
    for (int r = 0; r < K4_M / 8; ++r) {
        if (r < r_begin || r >= r_end) continue;
        double2 a0 = S2[vcol0 + 8 * r], a1 = S2[vcol1 + 8 * r];
        if (!va0) a0 = make_double2(0.0, 0.0);
        if (!va1) a1 = make_double2(0.0, 0.0);
        const double2 c_0 = S2[ccol0 + 8 * r], c_1 = S2[ccol1 + 8 * r];
        double cr0 = c_0.x, cr1 = c_1.x, ci0 = c_0.y, ci1 = c_1.y;
        // C += V (-W):  Cr += Vr Nr - Vi Ni,  Ci += Vr Ni + Vi Nr
        k4_dmma(cr0, cr1, a0.x, n0.x);
        k4_dmma(ci0, ci1, a0.x, n0.y);
        k4_dmma(cr0, cr1, a0.y, -n0.y);
        k4_dmma(ci0, ci1, a0.y, n0.x);
        k4_dmma(cr0, cr1, a1.x, n1.x);
        k4_dmma(ci0, ci1, a1.x, n1.y);
        k4_dmma(cr0, cr1, a1.y, -n1.y);
        k4_dmma(ci0, ci1, a1.y, n1.x);
        if (ok0) S2[ccol0 + 8 * r] = make_double2(cr0, ci0);
        if (ok1) S2[ccol1 + 8 * r] = make_double2(cr1, ci1);
    }
}

// The A fragments of V_b^H for the 16 k-steps of the tile's 64 rows (thread (g, t): V_b[4 s + t][g]).
__device__ __forceinline__ void k4_load_v(const PanelSmem &sm, const int j0, const int w, const int first,
                                          const int lane, double (&Vr)[16], double (&Vi)[16])
{
    double2 *const S2 = k4_shared();
    const int g = lane >> 2, t = lane & 3;
    const bool live = g < w && g >= first;
    const int vcol = sm.o_tile + (j0 + (live ? g : 0)) * K4_S + t;
#pragma unroll
    for (int s = 0; s < 16; ++s) {
        const double2 v = S2[vcol + 4 * s];
        Vr[s] = live ? v.x : 0.0;
        Vi[s] = live ? v.y : 0.0;
    }
}

// All reflections jstart..N-1 of [R; tile] (tile: K4_M rows x ncols columns; factor at o_R with
// RW columns right of column 0, diagonal at o_diag), software-pipelined over the panels: while
// the other warps apply panel p to the trailing chunks, the panel warp applies it to the chunk
// that holds panel p + 1 and factors that panel right away (T / v0 double-buffered), so the
// latency-bound panel chain and the tensor-core update of one fit overlap; two CTA barriers per
// panel.  The panel warp rotates with the panel index (and the CTA), so that the two
// co-resident fits do not queue their panels on the same scheduler.  `warp` must be
// warp-uniform for the compiler (see the kernel).
__device__ __forceinline__ void k4_factor_tile(const PanelSmem &sm, const int o_R, const int RW, const int o_diag,
                                               const int ncols, const int N, const int jstart, const int lane,
                                               const int warp)
{
    const int pfirst = jstart / K4_NB, plast = (N + K4_NB - 1) / K4_NB - 1;
    // the two CTAs of an SM sit in warp slots 0-3 and 4-7 (scheduler = slot % 4): offset their
    // panel warps by two so that both panels never queue on one scheduler
    unsigned slot0;
#ifndef QNMFIT_HOSTSIM
    asm("mov.u32 %0, %%warpid;" : "=r"(slot0));
#else
    slot0 = (unsigned)(threadIdx.x >> 5);
#endif
    const int rot = 2 * (int)((__shfl_sync(0xffffffffu, slot0, 0) >> 2) & 1);
    if (warp == ((pfirst + rot) & (K4_WARPS - 1)))
        k4_panel(sm, o_R, RW, o_diag, pfirst * K4_NB, N - pfirst * K4_NB < K4_NB ? N - pfirst * K4_NB : K4_NB,
                 jstart - pfirst * K4_NB, lane, pfirst & 1);
    __syncthreads();
#pragma unroll 1
    for (int pi = pfirst; pi <= plast; ++pi) {
        const int j0 = pi * K4_NB;
        const int w = N - j0 < K4_NB ? N - j0 : K4_NB;
        const int first = jstart > j0 ? jstart - j0 : 0;
        const int cbeg = j0 + w;
        const int nchunks = (ncols - cbeg + 7) / 8;
        const int buf = pi & 1;
        const int xw = sm.o_xw + warp * 8 * K4_XS;
        double Vr[16], Vi[16];
        if (pi < plast) {
            // chunk 0 = the columns of panel pi + 1, which its warp factors next: it is on the fit's
            // critical chain, so ALL warps apply panel pi to it — each forms X and W for the chunk
            // (the same 72 DMMA, in parallel on the four schedulers) and updates two of its eight
            // row tiles; the panel warp also stores the chunk's rows of R
            const int pw = (pi + 1 + rot) & (K4_WARPS - 1);
            const int slot = (warp - pw - 1) & (K4_WARPS - 1);      // 0..2 for the other warps, 3 for the panel warp
            K4_STAMP(0);
            k4_load_v(sm, j0, w, first, lane, Vr, Vi);
            K4_STAMP(1);
            k4_update_chunk(sm, o_R, RW, j0, w, first, cbeg, ncols, Vr, Vi, xw, lane, buf, 2 * warp, 2 * warp + 2, warp == pw);
            K4_STAMP(2);
            __syncthreads();
            K4_STAMP(3);
            if (warp == pw) {
                const int j1 = j0 + K4_NB;
                k4_panel(sm, o_R, RW, o_diag, j1, N - j1 < K4_NB ? N - j1 : K4_NB, 0, lane, buf ^ 1);
                K4_STAMP(4);
            } else {
                for (int ch = 1 + slot; ch < nchunks; ch += K4_WARPS - 1)
                    k4_update_chunk(sm, o_R, RW, j0, w, first, cbeg + 8 * ch, ncols, Vr, Vi, xw, lane, buf);
                K4_STAMP(5);
            }
        } else if (warp < nchunks) {                                 // last panel: every warp takes chunks
            K4_STAMP(6);
            k4_load_v(sm, j0, w, first, lane, Vr, Vi);
            for (int ch = warp; ch < nchunks; ch += K4_WARPS)
                k4_update_chunk(sm, o_R, RW, j0, w, first, cbeg + 8 * ch, ncols, Vr, Vi, xw, lane, buf);
            K4_STAMP(7);
        }
        __syncthreads();
        K4_STAMP(8);
    }
}

// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(K4_THREADS, 2) fit_panel_kernel(const __grid_constant__ FitParams p)
{
    QF_DYN_SMEM(smem_raw);
    const int N = p.n_modes, L = p.n_series, NC = N + L;
    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);      // warp-uniform by construction
    const int fit = blockIdx.x;
    PanelSmem sm;
    sm.carve(smem_raw, N, L);
    double2 *const S2 = k4_shared();
    double *const S1 = (double *)S2;

    const int fi = input_fit(p, fit);
    int rb = p.row_begin ? p.row_begin[fi] : p.row_begin_all;
    int re = p.row_end ? p.row_end[fi] : p.row_end_all;
    const double t0 = p.t0 ? p.t0[fi] : p.t0_all;
    if (rb < 0) rb = 0;
    if (re > p.n_times) re = p.n_times;
    if (re < rb) re = rb;
    const int K = re - rb;
    const double2 *coef = nullptr;
    if (p.coef) {
        const int ci = p.coef_index ? p.coef_index[fi] : fit_chi_index(p, fi);
        coef = p.coef + (long long)ci * L * N;
    }
    const bool two_phase = (coef != nullptr) || L > 1;
    const bool recur = p.dt_nominal > 0.0 && !p.omega_rows;

    for (int j = tid; j < N; j += K4_THREADS) {
        const double2 w = fit_omega(p, fi, j);
        S2[sm.o_om + j] = w;
        if (p.dt_nominal > 0.0) {
            const double2 q = design_entry(w, p.dt_nominal);
            S2[sm.o_qq + j] = q;
            S2[sm.o_qw + j] = c_mul(q, make_double2(w.y, -w.x));
        }
        S1[sm.o_diag1 + j] = 0.0;
        S1[sm.o_diag2 + j] = 0.0;
    }
    {
        const int n1 = (int)PackedR::entries(N, NC - 1);
        for (int e = tid; e < n1; e += K4_THREADS) S2[sm.o_R1 + e] = make_double2(0.0, 0.0);
        if (!PanelSmem::r2_in_tile(N, L)) {
            const int n2 = (int)PackedR::entries(N, N);
            for (int e = tid; e < n2; e += K4_THREADS) S2[sm.o_R2 + e] = make_double2(0.0, 0.0);
        }
    }
    __syncthreads();

    double sdd = 0.0, res2 = 0.0;
    if (!p.eval_only) {
        // ---------------- phase 1: [E | d_1..d_L], 64 rows at a time ----------------
        const int ntiles = (K + K4_M - 1) / K4_M;
#pragma unroll 1
        for (int tl = 0; tl < ntiles; ++tl) {
            const int row0 = rb + tl * K4_M;
            // exponentials: a (column, 16-row segment) pair per thread — anchor + corrected recurrence
            // (fit_small.cuh, small_leaf_uniform), or direct evaluation of every element
            for (int e = tid; e < N * 4; e += K4_THREADS) {
                const int c = e >> 2, seg = e & 3;
                const int col = sm.o_tile + c * K4_S + seg * 16;
                const int first = row0 + seg * 16;
                if (recur) {
                    if (first < re) {
                        const double dt = p.dt_nominal;
                        double tau = qf_sub_rn(p.times[first], t0);
                        double2 z = design_entry(S2[sm.o_om + c], tau);
                        const double2 q = S2[sm.o_qq + c], wq = S2[sm.o_qw + c];
#pragma unroll 4
                        for (int r = 0; r < 16; ++r) {
                            S2[col + r] = first + r < re ? z : make_double2(0.0, 0.0);
                            const int kn = first + r + 1 < re ? first + r + 1 : re - 1;
                            const double tau_n = qf_sub_rn(p.times[kn], t0);
                            const double de = qf_sub_rn(qf_sub_rn(tau_n, tau), dt);
                            tau = tau_n;
                            z = c_mul(z, make_double2(fma(wq.x, de, q.x), fma(wq.y, de, q.y)));
                        }
                    } else {
                        for (int r = 0; r < 16; ++r) S2[col + r] = make_double2(0.0, 0.0);
                    }
                } else {
                    for (int r = 0; r < 16; ++r) {
                        double2 v = make_double2(0.0, 0.0);
                        if (first + r < re) {
                            const double2 om = p.omega_rows ? p.omega_rows[(long long)c * p.n_times + first + r] : S2[sm.o_om + c];
                            v = design_entry(om, qf_sub_rn(p.times[first + r], t0));
                        }
                        S2[col + r] = v;
                    }
                }
            }
            // right-hand sides
            for (int e = tid; e < L * K4_M; e += K4_THREADS) {
                const int i = e / K4_M, r = e - i * K4_M;
                double2 v = make_double2(0.0, 0.0);
                if (row0 + r < re) v = p.data[(long long)i * p.series_stride + row0 + r];
                S2[sm.o_tile + (N + i) * K4_S + r] = v;
                sdd = fma(v.x, v.x, sdd);
                sdd = fma(v.y, v.y, sdd);
            }
            __syncthreads();
            k4_factor_tile(sm, sm.o_R1, NC - 1, sm.o_diag1, NC, N, 0, lane, warp);
            for (int e = tid; e < L * K4_M; e += K4_THREADS) {     // what is left of the right-hand sides
                const int i = e / K4_M, r = e - i * K4_M;
                const double2 v = S2[sm.o_tile + (N + i) * K4_S + r];
                res2 = fma(v.x, v.x, res2);
                res2 = fma(v.y, v.y, res2);
            }
            __syncthreads();
        }

        // ---------------- phase 2: stacked [R_E D_i | Y_i] ----------------
        if (two_phase) {
            if (PanelSmem::r2_in_tile(N, L)) {      // R2 lives behind the N + 1 columns of the phase-2 tile
                const int n2 = (int)PackedR::entries(N, N);
                for (int e = tid; e < n2; e += K4_THREADS) S2[sm.o_R2 + e] = make_double2(0.0, 0.0);
            }
            const int rows2 = N * L;
            const int ntiles2 = (rows2 + K4_M - 1) / K4_M;
#pragma unroll 1
            for (int tl = 0; tl < ntiles2; ++tl) {
                const int jstart = (tl * K4_M) / L;             // first non-zero column of the tile
                for (int e = tid; e < (N + 1) * K4_M; e += K4_THREADS) {
                    const int c = e / K4_M, r = e - c * K4_M;
                    const int q = tl * K4_M + r;
                    double2 v = make_double2(0.0, 0.0);
                    if (q < rows2) {
                        const int rr = q / L, i = q - rr * L;
                        if (c == N) v = S2[k4_ridx(sm.o_R1, NC - 1, rr, N + i)];
                        else if (c >= rr) {
                            const double2 cf = coef ? coef[i * N + c] : make_double2(1.0, 0.0);
                            const double d = S1[sm.o_diag1 + rr];
                            v = c == rr ? make_double2(cf.x * d, cf.y * d) : k4_cmul(cf, S2[k4_ridx(sm.o_R1, NC - 1, rr, c)]);
                        }
                    }
                    S2[sm.o_tile + c * K4_S + r] = v;
                }
                __syncthreads();
                k4_factor_tile(sm, sm.o_R2, N, sm.o_diag2, N + 1, N, jstart, lane, warp);
                for (int r = tid; r < K4_M; r += K4_THREADS) {
                    const double2 v = S2[sm.o_tile + N * K4_S + r];
                    res2 = fma(v.x, v.x, res2);
                    res2 = fma(v.y, v.y, res2);
                }
                __syncthreads();
            }
        }
    }
    struct_finish(p, sm, fit, N, L, rb, re, t0, coef, two_phase, sdd, res2);
}
