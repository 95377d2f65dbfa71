// xinv_colour_engine.cuh -- the generic "one in-place kernel per colour" engine.
// Handles every stencil (5-, 7-, 9-point), every boundary condition, every grid shape and every
// kernel of numbas.py offered here (XdKind); the marching engines (xinv_march2d.cuh, xinv_march3d.cuh),
// the cluster engine (xinv_cluster2d.cuh) and the resident engine (xinv_resident.cuh) take over
// wherever they apply ("Which engine runs").
#pragma once
#include "xinv_device.cuh"

enum XdKind { XD_STD2D = 0, XD_GEN2D = 1, XD_STD3D = 2, XD_STD2DT = 3, XD_GEN3D = 4, XD_STD1D = 5, XD_BIH2D = 6 };
#define XD_IS3D(kind) ((kind) == XD_STD3D || (kind) == XD_GEN3D)

#define XD_SWEEP_THREADS 128

// y-"extend" boundary rows, applied before each sweep on active slices:
// numbas.py:284-310 (2-D) and :87-115 (3-D: levels 1..nz-2 only).
__global__ void xd_extend_kernel(double *__restrict__ S, XdGeom g, double undef,
                                 const XdSliceState *__restrict__ st)
{
    const int b = blockIdx.z;
    if (!st[b].active) return;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.nx) return;
    // level handled by this block row
    i64 k = blockIdx.y;
    if (g.nz > 1) k += 1;                       // 3-D: interior levels only
    double *P = S + (i64)b * g.N + k * g.ny * g.nx;
    const i64 nx = g.nx, ny = g.ny;
    if (g.bcx == XD_BC_PERIODIC || (i >= 1 && i <= nx - 2)) {
        const double a = P[nx + i], z = P[(ny - 2) * nx + i];
        if (a != undef) P[i] = a;
        if (z != undef) P[(ny - 1) * nx + i] = z;
    } else {
        // non-periodic corners copy the diagonal neighbour (numbas.py:303-310)
        const i64 src = (i == 0) ? 1 : nx - 2;
        const double a = P[nx + src], z = P[(ny - 2) * nx + src];
        if (a != undef) P[i] = a;
        if (z != undef) P[(ny - 1) * nx + i] = z;
    }
}

// invert_standard_1D: the 'extend' condition is along x -- the end points take their neighbour's value before every
// sweep (numbas.py:686-690).  One thread per series.
__global__ void xd_extend1d_kernel(double *__restrict__ S, i64 nx, int batch, double undef, const XdSliceState *__restrict__ st)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch || !st[b].active) return;
    double *P = S + (i64)b * nx;
    const double a = P[1], z = P[nx - 2];
    if (a != undef) P[0] = a;
    if (z != undef) P[nx - 1] = z;
}

// invert_general_bih_2D: the two-row y-"extend" copy (numbas.py:1298-1343).  The periodic branch copies row 1 into row 0
// BEFORE it refreshes row 1 from row 2; the non-periodic branch copies row 2 into both and sets the 2 x 2 corners from
// the diagonal cell (2, 2) etc.  One thread per column, the corners by the threads of columns 0 and nx-1 AFTER the rows
// (the reference's order: a corner assignment overrides what the row loop wrote to (0, 1), (1, 1), ...).
__global__ void xd_extend_bih_kernel(double *__restrict__ S, XdGeom g, double undef, const XdSliceState *__restrict__ st)
{
    const int b = blockIdx.y;
    if (!st[b].active) return;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 nx = g.nx, ny = g.ny;
    if (i >= nx) return;
    double *P = S + (i64)b * g.N;
    double *r0 = P, *r1 = P + nx, *r2 = P + 2 * nx, *m1 = P + (ny - 1) * nx, *m2 = P + (ny - 2) * nx, *m3 = P + (ny - 3) * nx;
    if (g.bcx == XD_BC_PERIODIC) {
        if (r2[i] != undef) { r0[i] = r1[i]; r1[i] = r2[i]; }
        if (m3[i] != undef) { m1[i] = m3[i]; m2[i] = m3[i]; }
        return;
    }
    // columns 0, 1, nx-2, nx-1 end up with the corner values where the corner source is valid, else with what the row
    // loop gave them (columns 1 and nx-2) or untouched (columns 0 and nx-1)
    const bool wcorner = (i <= 1), ecorner = (i >= nx - 2);
    const i64 src = wcorner ? 2 : (ecorner ? nx - 3 : i);
    const bool inloop = (i >= 1 && i <= nx - 2);
    {
        const double a = r2[i], ac = r2[src];
        if (inloop && a != undef) { r0[i] = a; r1[i] = a; }
        if ((wcorner || ecorner) && ac != undef) { r0[i] = ac; r1[i] = ac; }
    }
    {
        const double z = m3[i], zc = m3[src];
        if (inloop && z != undef) { m1[i] = z; m2[i] = z; }
        if ((wcorner || ecorner) && zc != undef) { m1[i] = zc; m2[i] = zc; }
    }
}

// One colour of one biharmonic sweep, in place: blockIdx.x = row * nxblk + xblk over the rows 2 .. ny-3, thread t of a
// row handles column 3 t + (colour mod 3) (colours 0..8) or one of the last two columns (wrap-fix colours 9..14).
__global__ void __launch_bounds__(XD_SWEEP_THREADS)
xd_sweep_bih_kernel(double *__restrict__ Sall, XdCoef q, XdGeom g, int colour, int nxblk, const XdSliceState *__restrict__ st)
{
    const int b = blockIdx.y;
    if (!st[b].active) return;
    const i64 row = blockIdx.x / nxblk;
    const int xb = (int)(blockIdx.x - row * nxblk);
    const i64 j = 2 + row;
    i64 i;
    if (colour >= 9) {
        if (xb != 0 || threadIdx.x != 0) return;
        i = g.nx - 2 + (colour - 9) / 3;
        if ((int)(j % 3) != (colour - 9) % 3) return;
    } else {
        if ((int)(j % 3) != colour / 3) return;
        i = 3 * ((i64)xb * blockDim.x + threadIdx.x) + (colour % 3);
    }
    if (i < g.i0 || i >= g.i1) return;
    if (xd_colour_bih(g.wrapfix, g.nx, j, i) != colour) return;
    xd_update_bih(Sall + (i64)b * g.N, q, b, g.nx, j, i, g.bcx == XD_BC_PERIODIC);
}

// One colour of one sweep, in place.  Thread t of a row handles the t-th cell
// of that colour in the row.  blockIdx.x = row * nxblk + xblk, blockIdx.y = slice.
template <int KIND, bool HASB>
__global__ void __launch_bounds__(XD_SWEEP_THREADS)
xd_sweep_colour_kernel(double *__restrict__ Sall, XdCoef q, XdGeom g, int colour, int nxblk,
                       const XdSliceState *__restrict__ st)
{
    const int b = blockIdx.y;
    if (!st[b].active) return;
    const i64 row = blockIdx.x / nxblk;
    const int xb = (int)(blockIdx.x - row * nxblk);
    i64 k = 0, j;
    if (XD_IS3D(KIND))         { k = 1 + row / (g.ny - 2); j = 1 + row % (g.ny - 2); }
    else if (KIND == XD_STD1D) { j = 0; }
    else                       { j = 1 + row; }
    const i64 jk = j + k;
    const int base = (g.scheme == 4) ? 4 : 2;
    i64 i;
    if (colour >= base) {                       // wrap-fix colours: column nx-1 only
        if (xb != 0 || threadIdx.x != 0) return;
        i = g.nx - 1;
    } else if (g.scheme == 4) {
        if ((int)(j & 1) != (colour >> 1)) return;
        i = 2 * ((i64)xb * blockDim.x + threadIdx.x) + (colour & 1);
    } else {
        i = 2 * ((i64)xb * blockDim.x + threadIdx.x) + ((jk + colour) & 1);
    }
    if (i < g.i0 || i >= g.i1) return;
    if (xd_colour(g.scheme, g.wrapfix, g.nx, jk, j, i) != colour) return;
    const i64 ip = (i == g.nx - 1) ? 0 : i + 1;
    const i64 im = (i == 0) ? g.nx - 1 : i - 1;
    double *S = Sall + (i64)b * g.N;
    if (KIND == XD_STD2D) {
        xd_update_std2d<HASB>(S, q.c[0] + b * q.cs[0], HASB ? q.c[1] + b * q.cs[1] : nullptr,
                              q.c[2] + b * q.cs[2], q.c[3] + b * q.cs[3],
                              g.nx, j, i, ip, im, q.p[0], q.p[1], q.p[2], q.optArg, q.undef);
    } else if (KIND == XD_GEN2D) {
        xd_update_gen2d<HASB>(S, q.c[0] + b * q.cs[0], HASB ? q.c[1] + b * q.cs[1] : nullptr,
                              q.c[2] + b * q.cs[2], q.c[3] + b * q.cs[3], q.c[4] + b * q.cs[4],
                              q.c[5] + b * q.cs[5], q.c[6] + b * q.cs[6],
                              g.nx, j, i, ip, im, q.p[0], q.p[1], q.p[2], q.p[3], q.p[4],
                              q.optArg, q.undef);
    } else if (KIND == XD_STD3D) {
        xd_update_std3d(S, q.c[0] + b * q.cs[0], q.c[1] + b * q.cs[1], q.c[2] + b * q.cs[2],
                        q.c[3] + b * q.cs[3], g.ny, g.nx, k, j, i, ip, im,
                        q.p[0], q.p[1], q.p[2], q.optArg, q.undef);
    } else if (KIND == XD_STD2DT) {
        xd_update_std2dt(S, q.c[0] + b * q.cs[0], q.c[1] + b * q.cs[1], q.c[2] + b * q.cs[2], q.c[3] + b * q.cs[3],
                         q.c[4] + b * q.cs[4], q.c[5] + b * q.cs[5], g.nx, j, i, ip, im, q.p[0], q.p[1], q.p[2], q.optArg, q.undef);
    } else if (KIND == XD_GEN3D) {
        xd_update_gen3d(S, q, b, g.ny, g.nx, k, j, i, ip, im);
    } else {
        xd_update_std1d(S, q.c[0] + b * q.cs[0], q.c[1] + b * q.cs[1], q.c[2] + b * q.cs[2], i, ip, im, q.p[0], q.optArg, q.undef);
    }
}

// mean|S| over S != undef (numbas.py:1689-1728) + loop control, one pass.
// grid = (nblk, batch).  Every block reduces a contiguous chunk to one partial;
// the last block of a slice to finish (ticket) adds the partials in index order,
// so the sum does not depend on block scheduling, and applies xd_decide.
#define XD_NORM_THREADS 256
__global__ void __launch_bounds__(XD_NORM_THREADS)
xd_norm_decide_kernel(const double *__restrict__ Sall, i64 N, double undef, int nblk,
                      double *__restrict__ psum, i64 *__restrict__ pcnt,
                      unsigned *__restrict__ ticket, XdSliceState *__restrict__ st,
                      int *__restrict__ nactive, double tol, i64 mxLoop, int zero_exit)
{
    __shared__ double sm_sum[32];
    __shared__ i64 sm_cnt[32];
    __shared__ int is_last;
    const int b = blockIdx.y;
    if (!st[b].active) return;
    const double *S = Sall + (i64)b * N;
    const i64 chunk = (N + nblk - 1) / nblk;
    const i64 lo = (i64)blockIdx.x * chunk;
    const i64 hi = (lo + chunk < N) ? lo + chunk : N;
    double sum = 0.0;
    i64 cnt = 0;
    for (i64 p = lo + threadIdx.x; p < hi; p += blockDim.x) {
        const double v = S[p];
        if (v != undef) { sum += fabs(v); cnt += 1; }
    }
    xd_block_reduce(sum, cnt, sm_sum, sm_cnt);
    if (threadIdx.x == 0) {
        psum[(i64)b * nblk + blockIdx.x] = sum;
        pcnt[(i64)b * nblk + blockIdx.x] = cnt;
        __threadfence();
        const unsigned t = atomicAdd(&ticket[b], 1u);
        is_last = (t == (unsigned)nblk - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < 32) {
        double s = 0.0;
        i64 c = 0;
        const volatile double *vs = psum + (i64)b * nblk;
        const volatile i64 *vc = pcnt + (i64)b * nblk;
        for (int p = threadIdx.x; p < nblk; p += 32) { s += vs[p]; c += vc[p]; }
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s += __shfl_down_sync(0xffffffffu, s, o);
            c += __shfl_down_sync(0xffffffffu, c, o);
        }
        if (threadIdx.x == 0) {
            XdSliceState s_ = st[b];
            xd_decide(s_, s, c, tol, mxLoop, zero_exit);
            st[b] = s_;
            ticket[b] = 0u;
            if (!s_.active) atomicSub(nactive, 1);
        }
    }
}

// decide-only kernel for engines that already produced (sum,count) partials
__global__ void xd_decide_kernel(int nblk, const double *__restrict__ psum,
                                 const i64 *__restrict__ pcnt, XdSliceState *__restrict__ st,
                                 int *__restrict__ nactive, double tol, i64 mxLoop, int zero_exit)
{
    const int b = blockIdx.x;
    if (!st[b].active) return;
    double s = 0.0;
    i64 c = 0;
    for (int p = threadIdx.x; p < nblk; p += 32) { s += psum[(i64)b * nblk + p]; c += pcnt[(i64)b * nblk + p]; }
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        s += __shfl_down_sync(0xffffffffu, s, o);
        c += __shfl_down_sync(0xffffffffu, c, o);
    }
    if (threadIdx.x == 0) {
        XdSliceState s_ = st[b];
        xd_decide(s_, s, c, tol, mxLoop, zero_exit);
        st[b] = s_;
        if (!s_.active) atomicSub(nactive, 1);
    }
}
