/*
 * xinv.h -- C-ABI of libxinv_b200.so: the B200-native SOR elliptic inverter.
 *
 * Drop-in boundary: these entry points replace the call
 *     core.inv_*  ->  numbas.invert_*
 * of the reference (/root/reference/xinvert/core.py:130-139, :419-428, :60-69),
 * batched over all non-core slices (the serial `for selDict in loop_noncore`
 * loop of core.py:129/418/59 becomes the `batch` argument).
 *
 * Conventions
 *   - plain pointers and sizes only; no torch / numpy types.
 *   - all arrays are C-contiguous float64, x fastest: [batch][ny][nx] or
 *     [batch][nz][ny][nx].  S is in/out (warm start: the solve continues from
 *     whatever S holds, as numbas.py does); coefficients / forcing are read-only.
 *   - `flags` is a HOST array double[batch][3] = (overflow 0/1, last relative
 *     change of mean|S|, last loop index), the per-slice equivalent of the
 *     reference's flags[3] (numbas.py:403-408).  It is read on entry (values a
 *     slice keeps when it overflows in its first sweep) and written on return.
 *   - return value: 0 = ok, <0 = error (XINV_E_*); xinv_last_error() gives the
 *     text.  Numerical blow-up is NOT an error: it sets flags[b][0] = 1 and
 *     stops that slice, exactly like numbas.py:403-405.
 *   - pointers are borrowed for the duration of the call only.
 *   - one xinv_ctx per (thread, device); calls on one ctx are stream-ordered
 *     and must not be issued concurrently from several threads.
 */
#ifndef XINV_H
#define XINV_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define XINV_VERSION 102

/* boundary conditions (numbas.py BCy/BCx strings 'fixed' / 'extend' / 'periodic') */
#define XINV_BC_FIXED    0
#define XINV_BC_EXTEND   1
#define XINV_BC_PERIODIC 2

/* orderings */
#define XINV_ORDER_COLOUR 0   /* red-black (5/7-pt) or 4-colour (9-pt): the fast path   */
#define XINV_ORDER_LEX    1   /* reference's lexicographic order, wavefront-parallel:   */
                              /* bit-identical trajectory to numbas.py, slower          */

/* where S / coefficient / forcing pointers live */
#define XINV_MEM_HOST   0     /* library stages H2D / D2H itself                        */
#define XINV_MEM_DEVICE 1     /* device pointers on ctx's device (e.g. tensor.data_ptr) */

/* engines (opts.engine); 0 lets the library choose */
#define XINV_ENGINE_AUTO    0
#define XINV_ENGINE_COLOUR  1 /* one in-place kernel per colour + fused norm/decide     */
#define XINV_ENGINE_FUSED   2 /* TMA-staged fused red+black iteration kernel (2-D, B==0)*/
#define XINV_ENGINE_RESIDENT 3 /* small 2-D slices: the whole solve in one CTA, operands in shared memory */
#define XINV_ENGINE_CLUSTER 4 /* 2-D, B==0, row coefficients: the whole solve in one thread-block cluster (psi in registers + DSMEM) */

#define XINV_IO_F32_IN  1
#define XINV_IO_F32_OUT 2

#define XINV_ACCEL_NONE      0
#define XINV_ACCEL_CHEBYSHEV 1

/* error codes */
#define XINV_OK          0
#define XINV_E_ARG      -1    /* bad argument                                           */
#define XINV_E_CUDA     -2    /* CUDA runtime / driver error                            */
#define XINV_E_STATE    -3    /* begin/step/end protocol violated                       */
#define XINV_E_NOMEM    -4
#define XINV_E_UNSUPPORTED -5 /* e.g. lexicographic ordering with periodic-x 9-point    */
#define XINV_E_NCCL     -6

typedef struct xinv_ctx xinv_ctx;

/* Options; pass NULL for defaults (host pointers, colour ordering, dense
 * coefficients, automatic engine and check interval). */
typedef struct xinv_opts {
    int32_t struct_size;      /* = sizeof(xinv_opts)                                    */
    int32_t ordering;         /* XINV_ORDER_*                                           */
    int32_t mem_space;        /* XINV_MEM_*                                             */
    int32_t engine;           /* XINV_ENGINE_*                                          */
    int32_t check_every;      /* sweeps between host polls of the active count; 0=auto  */
    int32_t profile;          /* 1: time the dominant (sweep) kernels with CUDA events   */
    /* batch stride (in elements) of each coefficient/forcing array, in argument
     * order (A,B,C,F | A..G | A,B,C,F).  -1 = dense (one slice per batch entry,
     * as the reference materialises them); 0 = one slice shared by the whole
     * batch (what xarray broadcasting of time-independent coefficients means). */
    int64_t coef_stride[8];
    /* Convergence acceleration (NOT in the reference, off by default): XINV_ACCEL_CHEBYSHEV varies the
     * relaxation factor from half sweep to half sweep -- omega_0 = 1, omega_1 = 1/(1 - rho^2/2),
     * omega_{h+1} = 1/(1 - rho^2 omega_h/4) -> optArg, with rho^2 = 1 - (2/optArg - 1)^2 the squared Jacobi spectral
     * radius that optArg is the optimal SOR factor of -- instead of using optArg from the first sweep on.  Colour
     * ordering only; runs on the cluster, resident and colour engines. */
    int32_t accel;            /* XINV_ACCEL_*                                            */
    /* float32 I/O of the device front ends (xinv_std2d_rows / xinv_gen2d_rows / xinv_std3d_rows, host pointers only;
     * SURVEY 8f #4 "float32 storage"): XINV_IO_F32_IN -- the user's forcing is float32 in host memory;
     * XINV_IO_F32_OUT -- S_out is a float32 array.  Half the bytes cross PCIe; the values are widened / narrowed
     * (round to nearest) on the device and the solve is the same float64 solve, so the result equals what promoting
     * the forcing on the host and casting the result back gives (what xinvert_b200.apps does with float32 input). */
    int32_t io_f32;
} xinv_opts;

/* Statistics of the last solve on a ctx. */
typedef struct xinv_stats {
    int64_t sweeps_launched;  /* iterations launched (>= max_b(flags[b][2]+1))          */
    int64_t kernel_launches;  /* kernels of this library launched                        */
    int64_t cell_updates;     /* sum_b (flags[b][2]+1) * cells-of-slice                  */
    double  solve_ms;         /* device time of the iteration loop (CUDA events)         */
    double  h2d_ms, d2h_ms;   /* staging time when mem_space == HOST                     */
    int64_t h2d_bytes, d2h_bytes;
    int32_t engine;           /* engine that ran                                         */
    int32_t ncolours;
    double  sweep_ms;         /* mean device time of one full sweep (all colours)        */
    double  dom_ms;           /* opts.profile: summed device time of the dominant kernels */
    int64_t dom_launches;     /*               ... and how many launches that covers      */
    int64_t slow_strips;      /* reserved (always 0)                                      */
    int32_t iters_per_pass;   /* fused engine: SOR iterations per pass over HBM (T)       */
    int32_t row_coeffs;       /* fused engine: 1 = A and C were constant along x (RC kernels; 3-D: A only),
                                 2 = 3-D with A, B and C all constant along x (row-value kernels) */
} xinv_stats;

/* ---- context ---------------------------------------------------------- */
int  xinv_create(xinv_ctx **out, int device);
/* same, but work is issued on the caller's cudaStream_t (e.g. torch's current stream) */
int  xinv_create_on_stream(xinv_ctx **out, int device, void *cuda_stream);
void xinv_destroy(xinv_ctx *ctx);
const char *xinv_last_error(void);
int  xinv_version(void);
int  xinv_get_stats(const xinv_ctx *ctx, xinv_stats *out);
int  xinv_device_count(int *out);
int  xinv_synchronize(xinv_ctx *ctx);
/* CUDA-event stopwatch on the ctx's stream (what bench.py brackets its timed region with) */
int  xinv_timer_start(xinv_ctx *ctx);
int  xinv_timer_stop(xinv_ctx *ctx, double *ms_out);   /* records, synchronises, returns ms */

/* pinned host memory helpers (so callers can stage at full PCIe rate) */
int  xinv_host_alloc(void **out, int64_t bytes);
int  xinv_host_free(void *p);
/* *out = 1 if p lies in page-locked (cudaHostAlloc'ed / registered) memory, 0 for ordinary pageable memory:
 * callers stage pageable inputs through their own pinned buffers (a copy from pageable memory runs at a
 * fraction of the PCIe rate and cannot overlap) */
int  xinv_host_is_pinned(const void *p, int *out);
/* device memory helpers for callers without torch (bench / tests) */
int  xinv_dev_alloc(xinv_ctx *ctx, void **out, int64_t bytes);
int  xinv_dev_free(xinv_ctx *ctx, void *p);
int  xinv_memcpy_h2d(xinv_ctx *ctx, void *dst, const void *src, int64_t bytes);
int  xinv_memcpy_d2h(xinv_ctx *ctx, void *dst, const void *src, int64_t bytes);

/* ---- one-call solvers -------------------------------------------------- */

/* Replaces core.inv_standard2D -> numbas.invert_standard_2D
 * (core.py:129-139, numbas.py:215-416).  B == NULL means B is identically 0
 * (5-point stencil, red-black); otherwise the 9-point stencil with a 4-colour
 * ordering is used. */
int xinv_std2d(xinv_ctx *ctx, double *S, const double *A, const double *B,
               const double *C, const double *F,
               int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
               double delxSqr, double ratioQtr, double ratioSqr,
               double optArg, double undef, double *flags,
               int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* Poisson-type front end: the host-side preparation of apps.__mask_FS,
 * apps.__coeffs_Poisson and the de-masking of apps.__template (apps.py:2112-2159, :1397-1437,
 * :1386-1392) done on the device, for the common case icbc == None:
 *   F_user      the user's forcing [batch][ny][nx]; cells equal to user_undef (any NaN when
 *               user_undef is NaN) are masked ("land");
 *   F_row_scale [ny] or NULL: the forcing of row j is multiplied by F_row_scale[j]
 *               (lat-lon: cos(lat), apps.py:1409) -- one IEEE multiply, as the reference does;
 *   A_rows, C_rows  [ny]: coefficients constant along x (apps.py:1405-1408: cosH, 1/cosG;
 *               cartesian: 1), shared by the whole batch;
 *   S_out       [batch][ny][nx], output only: the solve starts from 0 (initS, apps.py:2145) and
 *               masked cells are set to out_undef on return (apps.py:1389-1392).
 * Numerically identical to building full arrays on the host and calling xinv_std2d.  Needs the
 * fused engine (2-D, even nx when periodic-x): otherwise XINV_E_UNSUPPORTED and the caller falls
 * back to xinv_std2d; also when an unmasked forcing value is not finite (the reference's
 * `maskF - maskF` template is not zero then and the host path reproduces its quirks). */
int xinv_std2d_rows(xinv_ctx *ctx, double *S_out, const double *A_rows, const double *C_rows,
                    const double *F_user, const double *F_row_scale, double user_undef, double out_undef,
                    int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                    double delxSqr, double ratioQtr, double ratioSqr,
                    double optArg, double undef, double *flags,
                    int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* The same front end for the general form with coefficients constant along x (invert_GillMatsuno,
 * invert_Stommel; apps.py:1609-1657, :1712-1748): rows = [5][ny] = A, C, D, E, F of every row;
 * G_user the user's forcing (user_undef / NaN = land).  g_mode 0: G = forcing (Gill-Matsuno,
 * apps.py:1655); g_mode 1: G = ((-forcing) / g_p1) / g_p2 (Stommel: -curl / D / rho0, apps.py:1746).
 * S_out is output only (zero initial guess), land is set to out_undef.  XINV_E_UNSUPPORTED as above. */
int xinv_gen2d_rows(xinv_ctx *ctx, double *S_out, const double *rows, const double *G_user,
                    int g_mode, double g_p1, double g_p2, double user_undef, double out_undef,
                    int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                    double delx, double delxSqr, double ratio, double ratioQtr, double ratioSqr,
                    double optArg, double undef, double *flags,
                    int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* The same front end for invert_omega (apps.py:766-827, :2016-2052): rows = [4][ny]:
 *   rows[0] = A of every row (lat-lon: f^2 cos(lat); cartesian: f^2), the same on every level;
 *   rows[1] = the factor of B = N2 * rows[1] (cosH | 1);  rows[2] = the divisor of C = N2 / rows[2] (cosG | 1);
 *   rows[3] = the forcing scale (cosG | 1).
 * N2 is read as N2[b*n2_strides[0] + k*n2_strides[1] + j*n2_strides[2] + i*n2_strides[3]] (elements; 0 =
 * broadcast): a scalar, a profile along one core dimension, one volume shared by the batch or a full array;
 * n2_count = elements in the buffer.  F_user: the user's forcing [batch][nz][ny][nx] (user_undef / NaN = land);
 * S_out is output only (zero initial guess), land is set to out_undef.  Needs the 3-D fused engine
 * (even nx when periodic-x): otherwise XINV_E_UNSUPPORTED and the caller falls back to xinv_std3d. */
int xinv_std3d_rows(xinv_ctx *ctx, double *S_out, const double *rows, const double *N2,
                    const int64_t *n2_strides, int64_t n2_count, const double *F_user,
                    double user_undef, double out_undef,
                    int64_t batch, int64_t nz, int64_t ny, int64_t nx, int bcz, int bcy, int bcx,
                    double delxSqr, double ratio2Sqr, double ratio1Sqr,
                    double optArg, double undef, double *flags,
                    int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* Replaces core.inv_general2D -> numbas.invert_general_2D
 * (core.py:418-428, numbas.py:987-1201).  B == NULL means B == 0. */
int xinv_gen2d(xinv_ctx *ctx, double *S, const double *A, const double *B,
               const double *C, const double *D, const double *E,
               const double *F, const double *G,
               int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
               double delx, double delxSqr, double ratio, double ratioQtr,
               double ratioSqr, double optArg, double undef, double *flags,
               int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* Replaces core.inv_standard3D -> numbas.invert_standard_3D
 * (core.py:59-69, numbas.py:15-212).  bcz is accepted and ignored, as in the
 * reference (numbas.py never reads BCz). */
int xinv_std3d(xinv_ctx *ctx, double *S, const double *A, const double *B,
               const double *C, const double *F,
               int64_t batch, int64_t nz, int64_t ny, int64_t nx,
               int bcz, int bcy, int bcx,
               double delxSqr, double ratio2Sqr, double ratio1Sqr,
               double optArg, double undef, double *flags,
               int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* ---- epilogue: flow components from the inverted field ----------------------------------
 * Replaces the array work of apps.cal_flow (apps.py:1181-1317): centred differences of S along
 * the two core dimensions -- numpy.gradient's formulas, chosen by the caller exactly as numpy
 * chooses them -- and the per-row combination of the 'GillMatsuno' branch (apps.py:1277-1317) or
 * the metric division and signs of the 'streamfunction' / 'velocitypotential' branches
 * (apps.py:1207-1271, finitediffs.py:151-207, :548-659).  All pointers follow opts->mem_space. */
#define XINV_EDGE_ONESIDED 0  /* no padding: one-sided differences at the ends (numpy.gradient, edge_order 1) */
#define XINV_EDGE_FIXED    1  /* padBCs 'fixed': the value beyond the end is lo / hi                              */
#define XINV_EDGE_EXTEND   2  /* 'extend': the edge value                                                         */
#define XINV_EDGE_REFLECT  3  /* 'reflect': the first inner value                                                 */
#define XINV_EDGE_PERIODIC 4  /* 'periodic': the value from the other end                                         */
#define XINV_FLOW_GRAD     0  /* out1 = s1 * (dS/dy / rows[0][j]), out2 = s2 * (dS/dx / rows[1][j]); swap: (x, y) */
#define XINV_FLOW_GM_LL    1  /* Gill-Matsuno lat-lon: rows = coef1, coef2, cosLat                                */
#define XINV_FLOW_GM_CART  2  /* Gill-Matsuno cartesian: rows = coef1, coef2                                      */
typedef struct xinv_flow_axis {
    int32_t uniform;          /* 1: (f[i+1] - f[i-1]) / den;  0: w[0][i] f[i-1] + w[1][i] f[i] + w[2][i] f[i+1]   */
    int32_t edge;             /* XINV_EDGE_*                                                                      */
    double den;               /* uniform: 2 dx                                                                    */
    double lo, hi;            /* ONESIDED: spacing of the one-sided differences; FIXED: the fill values           */
    const double *w;          /* non-uniform: [3][n]                                                              */
} xinv_flow_axis;
typedef struct xinv_flow_desc {
    int32_t struct_size;      /* = sizeof(xinv_flow_desc)                                                         */
    int32_t comb;             /* XINV_FLOW_*                                                                      */
    int32_t swap;             /* GRAD: 1 = return (x-derivative, y-derivative)                                    */
    int32_t nrows;            /* rows of `rows` (2 or 3)                                                          */
    double s1, s2;            /* GRAD: +1 / -1                                                                    */
    double deg2m;             /* GM_LL                                                                            */
    xinv_flow_axis y, x;
    const double *rows;       /* Dense front end (invert_Eliassen, apps.py:300-346 with apps.__mask_FS :2112-2159, apps.__coeffs_Eliassen :1582-1606
 * and the de-masking of apps.__template :1386-1392, for icbc == None): A, B, C as the caller holds them (dense, or one
 * slice shared by the batch through opts.coef_stride), the user's forcing with user_undef (any NaN when that is NaN)
 * marking land; the masked forcing, the zero initial guess and the de-masked result (land = out_undef) are formed on
 * the device.  S_out is output only.  XINV_E_UNSUPPORTED when an unmasked forcing value is not finite. */
int xinv_std2d_front(xinv_ctx *ctx, double *S_out, const double *A, const double *B, const double *C,
                     const double *F_user, double user_undef, double out_undef,
                     int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                     double delxSqr, double ratioQtr, double ratioSqr, double optArg, double undef,
                     double *flags, int64_t mxLoop, double tolerance, const xinv_opts *opts);

/* ---- The remaining SOR kernels of numbas.py, on the generic colour engine ---------------------
 * Same conventions as above (batch axis first, S in/out, flags[batch][3], opts may be NULL; XINV_ORDER_COLOUR
 * only).  Each mirrors the numba signature it replaces, minus the arguments that kernel never reads:
 *   xinv_std2d_test  numbas.invert_standard_2D_test (numbas.py:421-424): nine-point stencil with two cross-term
 *                    coefficients (B, C), D in the role of invert_standard_2D's C and a linear term E;
 *   xinv_gen3d       numbas.invert_general_3D (numbas.py:746-749);  bcz accepted and ignored as BCz is there;
 *   xinv_std1d       numbas.invert_standard_1D (numbas.py:633-635): batch series of nx points; bcx may be
 *                    XINV_BC_EXTEND here (end points copy their neighbour before every sweep). */
int xinv_std2d_test(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *C,
                    const double *D, const double *E, const double *F,
                    int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                    double delxSqr, double ratioQtr, double ratioSqr, double optArg, double undef,
                    double *flags, int64_t mxLoop, double tolerance, const xinv_opts *opts);
int xinv_gen3d(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *C,
               const double *D, const double *E, const double *F, const double *G, const double *H,
               int64_t batch, int64_t nz, int64_t ny, int64_t nx, int bcz, int bcy, int bcx,
               double delx, double delxSqr, double ratio2, double ratio1, double ratio2Sqr, double ratio1Sqr,
               double optArg, double undef, double *flags, int64_t mxLoop, double tolerance,
               const xinv_opts *opts);
int xinv_std1d(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *F,
               int64_t batch, int64_t nx, int bcx, double delxSqr, double optArg, double undef,
               double *flags, int64_t mxLoop, double tolerance, const xinv_opts *opts);
/*   xinv_bih2d       numbas.invert_general_bih_2D (numbas.py:1205-1210): 13-point biharmonic (invert_StommelMunk);
 *                    nine colours (fifteen with periodic-x when nx is not a multiple of 3); rows 2 .. ny-3 and --
 *                    unless x is periodic -- columns 2 .. nx-3 are updated; the two-row extend condition and the
 *                    reference's edge-column arithmetic are reproduced as they are (xinv_device.cuh: xd_update_bih).
 *                    The ten arrays are dense ([batch][ny][nx]; opts.coef_stride applies to the first eight). */
int xinv_bih2d(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *C, const double *D,
               const double *E, const double *F, const double *G, const double *H, const double *I,
               const double *J, int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
               double delxSSr, double delxTr, double delxSqr, double ratio, double ratioSSr, double ratioQtr,
               double ratioSqr, double optArg, double undef, double *flags, int64_t mxLoop, double tolerance,
               const xinv_opts *opts);

/* [nrows][ny]                                                                      */
} xinv_flow_desc;
int xinv_flow2d(xinv_ctx *ctx, double *out1, double *out2, const double *S,
                int64_t batch, int64_t ny, int64_t nx, const xinv_flow_desc *desc, const xinv_opts *opts);

/* ---- stepwise protocol (used by the multi-GPU driver so that ranks can
 *      exchange their active-slice counts between chunks of sweeps) ---------
 * xinv_*_begin stages the problem and initialises per-slice loop state;
 * xinv_step runs up to `sweeps` more iterations and returns how many slices of
 * this ctx are still active; xinv_end writes S / flags back and releases the
 * problem.  xinv_std2d(...) == begin + step-until-0 + end. */
int xinv_std2d_begin(xinv_ctx *ctx, double *S, const double *A, const double *B,
                     const double *C, const double *F,
                     int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                     double delxSqr, double ratioQtr, double ratioSqr,
                     double optArg, double undef, double *flags,
                     int64_t mxLoop, double tolerance, const xinv_opts *opts);
int xinv_gen2d_begin(xinv_ctx *ctx, double *S, const double *A, const double *B,
                     const double *C, const double *D, const double *E,
                     const double *F, const double *G,
                     int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                     double delx, double delxSqr, double ratio, double ratioQtr,
                     double ratioSqr, double optArg, double undef, double *flags,
                     int64_t mxLoop, double tolerance, const xinv_opts *opts);
int xinv_std3d_begin(xinv_ctx *ctx, double *S, const double *A, const double *B,
                     const double *C, const double *F,
                     int64_t batch, int64_t nz, int64_t ny, int64_t nx,
                     int bcz, int bcy, int bcx,
                     double delxSqr, double ratio2Sqr, double ratio1Sqr,
                     double optArg, double undef, double *flags,
                     int64_t mxLoop, double tolerance, const xinv_opts *opts);
int xinv_step(xinv_ctx *ctx, int64_t sweeps, int64_t *n_active_out);
int xinv_end(xinv_ctx *ctx);

/* ---- multi-GPU: scalar all-reduce of the active-slice count over NCCL ----
 * (SURVEY.md 8e).  The unique id is created on rank 0 and distributed by the
 * caller (torch.distributed broadcast, MPI, a file ...). */
int xinv_nccl_unique_id(void *id128);                 /* writes 128 bytes */
int xinv_nccl_init(xinv_ctx *ctx, const void *id128, int rank, int world);
int xinv_nccl_allreduce_active(xinv_ctx *ctx, int64_t local, int64_t *global_out);
int xinv_nccl_finalize(xinv_ctx *ctx);

#ifdef __cplusplus
}
#endif
#endif /* XINV_H */
