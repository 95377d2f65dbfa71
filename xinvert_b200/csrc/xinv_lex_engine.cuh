// xinv_lex_engine.cuh -- XINV_ORDER_LEX: the reference's lexicographic in-place
// Gauss-Seidel order (numbas.py:312-399 / :117-195 / :1092-1184), executed as a
// skewed wavefront so that every cell sees exactly the neighbour values the
// serial loop nest would give it.  The trajectory (every iterate, the loop
// count, flags) is therefore the reference's own; this is the parity bridge,
// not the speed path (one CTA per slice, one __syncthreads per wavefront step).
//
// Wavefront schedule: cell (k,j,i) is updated at step
//        t = (i - i0) + sj*(j-1) + sk*(k-1)
//   5-/7-point stencils: sj = sk = 1.  "New" neighbours (k,j,i-1), (k,j-1,i),
//     (k-1,j,i) sit on step t-1; "old" neighbours (i+1, j+1, k+1) on t+1.  With
//     periodic-x the west column reads old (j,nx-1) [updated nx-1 steps later] and
//     the east column reads new (j,0) [nx-1 steps earlier]: both consistent.
//   9-point stencil (B != 0), non-periodic x: sj = 2, so that (j-1,i+1) [new] is
//     on step t-1 and (j+1,i-1) [old] on t+1.
//   9-point with periodic-x needs new (j-1,nx-1) for (j,0), i.e. whole rows in
//     sequence: rejected with XINV_E_UNSUPPORTED in problem_begin.
#pragma once
#include "xinv_device.cuh"
#include "xinv_colour_engine.cuh"

#define XD_LEX_THREADS 1024

template <int KIND, bool HASB>
__global__ void __launch_bounds__(XD_LEX_THREADS)
xd_lex_sweep_kernel(double *Sall, XdCoef q, XdGeom g, const XdSliceState *__restrict__ st)
{
    const int b = blockIdx.x;
    if (!st[b].active) return;
    double *S = Sall + (i64)b * g.N;
    const i64 nyi = g.ny - 2;                                   // interior rows per level
    const i64 nzi = (KIND == XD_STD3D) ? g.nz - 2 : 1;
    const i64 R = nyi * nzi;
    const int sj = (HASB && KIND != XD_STD3D) ? 2 : 1;
    const i64 ncols = g.i1 - g.i0;
    const i64 T = ncols + sj * (nyi - 1) + (nzi - 1);

    // y-"extend" rows first (numbas.py:284-310 / :87-115), by the same CTA
    if (g.bcy == XD_BC_EXTEND) {
        const i64 levels = (KIND == XD_STD3D) ? g.nz - 2 : 1;
        for (i64 w = threadIdx.x; w < levels * g.nx; w += blockDim.x) {
            const i64 lv = w / g.nx, i = w - lv * g.nx;
            double *P = S + ((KIND == XD_STD3D) ? (lv + 1) * g.ny * g.nx : 0);
            i64 src = i;
            if (g.bcx != XD_BC_PERIODIC) { if (i == 0) src = 1; else if (i == g.nx - 1) src = g.nx - 2; }
            const double a = P[g.nx + src];
            if (a != q.undef) P[i] = a;
        }
        __syncthreads();
        for (i64 w = threadIdx.x; w < levels * g.nx; w += blockDim.x) {
            const i64 lv = w / g.nx, i = w - lv * g.nx;
            double *P = S + ((KIND == XD_STD3D) ? (lv + 1) * g.ny * g.nx : 0);
            i64 src = i;
            if (g.bcx != XD_BC_PERIODIC) { if (i == 0) src = 1; else if (i == g.nx - 1) src = g.nx - 2; }
            const double z = P[(g.ny - 2) * g.nx + src];
            if (z != q.undef) P[(g.ny - 1) * g.nx + i] = z;
        }
        __syncthreads();
    }
    if (R <= 0 || ncols <= 0) return;

    const double *cA = q.c[0] + b * q.cs[0];
    const double *cB = q.c[1] ? q.c[1] + b * q.cs[1] : nullptr;
    const double *c2 = q.c[2] + b * q.cs[2];
    const double *c3 = q.c[3] + b * q.cs[3];
    const double *c4 = (KIND == XD_GEN2D) ? q.c[4] + b * q.cs[4] : nullptr;
    const double *c5 = (KIND == XD_GEN2D) ? q.c[5] + b * q.cs[5] : nullptr;
    const double *c6 = (KIND == XD_GEN2D) ? q.c[6] + b * q.cs[6] : nullptr;

    for (i64 t = 0; t < T; ++t) {
        for (i64 r = threadIdx.x; r < R; r += blockDim.x) {
            const i64 kk = r / nyi, jj = r - kk * nyi;          // zero-based interior level / row
            const i64 off = sj * jj + kk;
            const i64 ci = t - off;
            if (ci < 0 || ci >= ncols) continue;
            const i64 i = g.i0 + ci, j = jj + 1, k = kk + 1;
            const i64 ip = (i == g.nx - 1) ? 0 : i + 1;
            const i64 im = (i == 0) ? g.nx - 1 : i - 1;
            if (KIND == XD_STD2D)
                xd_update_std2d<HASB>(S, cA, cB, c2, c3, g.nx, j, i, ip, im,
                                      q.p[0], q.p[1], q.p[2], q.optArg, q.undef);
            else if (KIND == XD_GEN2D)
                xd_update_gen2d<HASB>(S, cA, cB, c2, c3, c4, c5, c6, g.nx, j, i, ip, im,
                                      q.p[0], q.p[1], q.p[2], q.p[3], q.p[4], q.optArg, q.undef);
            else
                xd_update_std3d(S, cA, cB, c2, c3, g.ny, g.nx, k, j, i, ip, im,
                                q.p[0], q.p[1], q.p[2], q.optArg, q.undef);
        }
        __syncthreads();
    }
}

static int lex_sweep(cudaStream_t stream, int kind, bool hasB, const XdGeom &g, const XdCoef &q, i64 batch,
                     double *dS, XdSliceState *st, int nblk_norm, double *psum, i64 *pcnt, unsigned *ticket,
                     int *nactive, double tol, i64 mxLoop, int zero_exit, int64_t *launches)
{
    dim3 grid((unsigned)batch);
    if (kind == XD_STD2D) {
        if (hasB) xd_lex_sweep_kernel<XD_STD2D, true><<<grid, XD_LEX_THREADS, 0, stream>>>(dS, q, g, st);
        else      xd_lex_sweep_kernel<XD_STD2D, false><<<grid, XD_LEX_THREADS, 0, stream>>>(dS, q, g, st);
    } else if (kind == XD_GEN2D) {
        if (hasB) xd_lex_sweep_kernel<XD_GEN2D, true><<<grid, XD_LEX_THREADS, 0, stream>>>(dS, q, g, st);
        else      xd_lex_sweep_kernel<XD_GEN2D, false><<<grid, XD_LEX_THREADS, 0, stream>>>(dS, q, g, st);
    } else {
        xd_lex_sweep_kernel<XD_STD3D, false><<<grid, XD_LEX_THREADS, 0, stream>>>(dS, q, g, st);
    }
    dim3 ngrid((unsigned)nblk_norm, (unsigned)batch, 1);
    xd_norm_decide_kernel<<<ngrid, XD_NORM_THREADS, 0, stream>>>(dS, g.N, q.undef, nblk_norm, psum, pcnt, ticket,
                                                               st, nactive, tol, mxLoop, zero_exit);
    *launches += 2;
    return 0;
}
