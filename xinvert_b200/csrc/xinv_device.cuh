// xinv_device.cuh -- device-side building blocks shared by every engine:
// colourings, the per-cell SOR updates (one IEEE binary64 operation per
// +,-,*,/ of the reference expression, in the reference's evaluation order;
// the translation unit is compiled with -fmad=false so nothing is contracted),
// and deterministic block reductions.
//
// Reference semantics followed here (file:line in /root/reference/xinvert):
//   invert_standard_2D  numbas.py:344-369 (+ west/east columns :315-340, :374-399)
//   invert_general_2D   numbas.py:1126-1153 (+ :1095-1122, :1157-1184)
//   invert_standard_3D  numbas.py:147-169 (+ :121-143, :173-195)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef long long i64;

#define XD_BC_FIXED    0
#define XD_BC_EXTEND   1
#define XD_BC_PERIODIC 2

// ---------------------------------------------------------------------------
// Colourings.  2-colour scheme: parity of the index sum (5-/7-point stencils).
// 4-colour scheme: 2*(j&1)+(i&1) (9-point stencil).  With periodic-x and odd nx
// the wrap neighbours (columns 0 and nx-1) would share a colour, so column nx-1
// is moved to two extra colours indexed by row parity ("wrap-fix").
// (The test oracle restates this colouring independently; tests compare them.)
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ int xd_num_colours(int scheme, int wrapfix)
{
    return (scheme == 4 ? 4 : 2) + (wrapfix ? 2 : 0);
}

__host__ __device__ __forceinline__ int xd_colour(int scheme, int wrapfix, i64 nx, i64 jk, i64 j, i64 i)
{
    // jk = j (2-D) or j+k (3-D): the row-parity that enters the 2-colour scheme
    if (scheme == 4) {
        if (wrapfix && i == nx - 1) return 4 + (int)(j & 1);
        return 2 * (int)(j & 1) + (int)(i & 1);
    }
    if (wrapfix && i == nx - 1) return 2 + (int)(jk & 1);
    return (int)((i + jk) & 1);
}

// ---------------------------------------------------------------------------
// Kernel argument blocks (passed by value)
// ---------------------------------------------------------------------------
struct XdGeom {
    i64 nz, ny, nx;       // nz == 1 for 2-D problems
    i64 N;                // cells per slice
    int bcy, bcx;
    int i0, i1;           // updated columns [i0, i1)
    int scheme, wrapfix, ncol;
};

struct XdCoef {
    const double *c[12];  // coefficient / forcing arrays in C-ABI argument order
    i64 cs[12];           // batch stride of each (elements); 0 = shared by batch
    double p[8];          // scalar parameters, meaning per problem kind
    double optArg, undef;
};

// ---------------------------------------------------------------------------
// Per-cell updates.  S points at the slice; idx helpers take (row offset, col).
// ---------------------------------------------------------------------------

// standard 2-D.  c[] = {A,B,C,F}; p[] = {delxSqr, ratioQtr, ratioSqr}
template <bool HASB>
__device__ __forceinline__ void xd_update_std2d(double *__restrict__ S,
    const double *__restrict__ A, const double *__restrict__ B,
    const double *__restrict__ C, const double *__restrict__ F,
    i64 nx, i64 j, i64 i, i64 ip, i64 im,
    double delxSqr, double ratioQtr, double ratioSqr, double optArg, double undef)
{
    const i64 c = j * nx + i, n = c + nx, s = c - nx;
    const i64 e = j * nx + ip, w = j * nx + im;
    const double Fc = F[c], An = A[n], Ac = A[c], Ce = C[e], Cc = C[c];
    bool cond = (Fc != undef) & (An != undef) & (Ac != undef) & (Ce != undef) & (Cc != undef);
    double Be = 0, Bw = 0, Bn = 0, Bs = 0, Bq = 0;
    if (HASB) {
        Be = B[e]; Bw = B[w]; Bn = B[n]; Bs = B[s];
        cond = cond & (Be != undef) & (Bw != undef) & (Bn != undef) & (Bs != undef);
        Bq = (i == 0) ? B[n + 1] : Bn;            // numbas.py:327 west-column quirk
    }
    if (!cond) return;
    const double Sc = S[c], Sn = S[n], Ss = S[s], Se = S[e], Sw = S[w];
    const double t1 = (An * (Sn - Sc) - Ac * (Sc - Ss)) * ratioSqr;
    const double t4 = (Ce * (Se - Sc) - Cc * (Sc - Sw));
    double temp;
    if (HASB) {
        const double Sne = S[n - i + ip], Snw = S[n - i + im];
        const double Sse = S[s - i + ip], Ssw = S[s - i + im];
        const double Sse2 = (i == 0) ? Ss : Sse;  // numbas.py:328 west-column quirk
        const double t2 = (Bq * (Sne - Snw) - Bs * (Sse2 - Ssw)) * ratioQtr;
        const double t3 = (Be * (Sne - Sse) - Bw * (Snw - Ssw)) * ratioQtr;
        temp = (((t1 + t2) + t3) + t4) - Fc * delxSqr;
    } else {
        temp = (t1 + t4) - Fc * delxSqr;
    }
    temp = temp * (optArg / ((An + Ac) * ratioSqr + (Ce + Cc)));
    S[c] = Sc + temp;
}

// The same update with everything that does not change between sweeps formed beforehand, by the reference's own
// operations (as the marching engines do): Fd = F * delxSqr, or a marker NaN (high word XD_SKIP_HI) where the cell must
// never be updated; fac = optArg / denominator.  Same operations on the same operands as xd_update_std2d: same bits.
#define XD_SKIP_HI 0x7ff4dead
// Storage: a row holds its even columns first, then its odd ones (XD_SPLIT_POS) -- the cells of one colour of a row
// are contiguous, so a warp's loads of a colour step are unit-stride instead of stride-2 (no shared-memory bank
// conflicts); he = (nx + 1) / 2.
#define XD_SPLIT_POS(i, he) (((i) >> 1) + ((i) & 1) * (he))
template <bool HASB>
__device__ __forceinline__ void xd_update_std2d_pre(double *__restrict__ S,
    const double *__restrict__ A, const double *__restrict__ B, const double *__restrict__ C,
    const double *__restrict__ Fd, const double *__restrict__ fac,
    int nx, int he, int j, int i, int ip, int im, double ratioQtr, double ratioSqr)
{
    const int row = j * nx;
    const int pi = XD_SPLIT_POS(i, he), pe = XD_SPLIT_POS(ip, he), pw = XD_SPLIT_POS(im, he);
    const int c = row + pi, n = c + nx, s = c - nx;
    const int e = row + pe, w = row + pw;
    const double Fdc = Fd[c];
    if (__double2hiint(Fdc) == XD_SKIP_HI) return;
    const double An = A[n], Ac = A[c], Ce = C[e], Cc = C[c];
    const double Sc = S[c], Sn = S[n], Ss = S[s], Se = S[e], Sw = S[w];
    const double t1 = (An * (Sn - Sc) - Ac * (Sc - Ss)) * ratioSqr;
    const double t4 = (Ce * (Se - Sc) - Cc * (Sc - Sw));
    double temp;
    if (HASB) {
        const double Be = B[e], Bw = B[w], Bn = B[n], Bs = B[s];
        const double Bq = (i == 0) ? B[row + nx + XD_SPLIT_POS(1, he)] : Bn;   // numbas.py:327 west-column quirk: B[j+1,1]
        const double Sne = S[e + nx], Snw = S[w + nx];
        const double Sse = S[e - nx], Ssw = S[w - nx];
        const double Sse2 = (i == 0) ? Ss : Sse;               // numbas.py:328 west-column quirk
        const double t2 = (Bq * (Sne - Snw) - Bs * (Sse2 - Ssw)) * ratioQtr;
        const double t3 = (Be * (Sne - Sse) - Bw * (Snw - Ssw)) * ratioQtr;
        temp = (((t1 + t2) + t3) + t4) - Fdc;
    } else {
        temp = (t1 + t4) - Fdc;
    }
    temp = temp * fac[c];
    S[c] = Sc + temp;
}

// general 2-D.  c[] = {A,B,C,D,E,F,G}; p[] = {delx, delxSqr, ratio, ratioQtr, ratioSqr}
template <bool HASB>
__device__ __forceinline__ void xd_update_gen2d(double *__restrict__ S,
    const double *__restrict__ A, const double *__restrict__ B,
    const double *__restrict__ C, const double *__restrict__ D,
    const double *__restrict__ E, const double *__restrict__ F,
    const double *__restrict__ G,
    i64 nx, i64 j, i64 i, i64 ip, i64 im,
    double delx, double delxSqr, double ratio, double ratioQtr, double ratioSqr,
    double optArg, double undef)
{
    const i64 c = j * nx + i, n = c + nx, s = c - nx;
    const i64 e = j * nx + ip, w = j * nx + im;
    const double Gc = G[c], Ac = A[c], Cc = C[c], Dc = D[c], Ec = E[c], Fc = F[c];
    bool cond = (Gc != undef) & (Ac != undef) & (Cc != undef) & (Dc != undef) &
                (Ec != undef) & (Fc != undef);
    double Bc = 0;
    if (HASB) { Bc = B[c]; cond = cond & (Bc != undef); }
    if (!cond) return;
    const double Sc = S[c], Sn = S[n], Ss = S[s], Se = S[e], Sw = S[w];
    double temp = Ac * ((Sn - Sc) - (Sc - Ss)) * ratioSqr;
    if (HASB) {
        const double Sne = S[n - i + ip], Snw = S[n - i + im];
        const double Sse = S[s - i + ip], Ssw = S[s - i + im];
        temp = temp + Bc * ((Sne - Sse) - (Snw - Ssw)) * ratioQtr;
    }
    temp = temp + Cc * ((Se - Sc) - (Sc - Sw));
    temp = temp + (Dc * (Sn - Ss) * ratio + Ec * (Se - Sw)) * delx / 2.0;
    temp = temp + (Fc * Sc - Gc) * delxSqr;
    temp = temp * (optArg / ((Ac * ratioSqr + Cc) * 2.0 - Fc * delxSqr));
    S[c] = Sc + temp;
}

// standard 3-D.  c[] = {A,B,C,F}; p[] = {delxSqr, ratio2Sqr, ratio1Sqr}
__device__ __forceinline__ void xd_update_std3d(double *__restrict__ S,
    const double *__restrict__ A, const double *__restrict__ B,
    const double *__restrict__ C, const double *__restrict__ F,
    i64 ny, i64 nx, i64 k, i64 j, i64 i, i64 ip, i64 im,
    double delxSqr, double ratio2Sqr, double ratio1Sqr, double optArg, double undef)
{
    const i64 pl = ny * nx;
    const i64 row = k * pl + j * nx;
    const i64 c = row + i, e = row + ip, w = row + im;
    const i64 n = c + nx, s = c - nx, u = c + pl, d = c - pl;
    const double Fc = F[c], Au = A[u], Ac = A[c], Bn = B[n], Bc = B[c], Ce = C[e], Cc = C[c];
    const bool cond = (Fc != undef) & (Au != undef) & (Ac != undef) & (Bn != undef) &
                      (Bc != undef) & (Ce != undef) & (Cc != undef);
    if (!cond) return;
    const double Sc = S[c];
    double temp = (
        (Au * (S[u] - Sc) - Ac * (Sc - S[d])) * ratio2Sqr +
        (Bn * (S[n] - Sc) - Bc * (Sc - S[s])) * ratio1Sqr +
        (Ce * (S[e] - Sc) - Cc * (Sc - S[w]))
    ) - Fc * delxSqr;
    temp = temp * (optArg / ((Au + Ac) * ratio2Sqr + (Bn + Bc) * ratio1Sqr + (Ce + Cc)));
    S[c] = Sc + temp;
}

// ---- the remaining kernels of numbas.py on the same skeleton (colour engine) ----

// invert_standard_2D_test (numbas.py:420-629; interior :560-583, west :525-553 with the B[j+1,1] / S[j-1,0] quirks of
// :538-539, east :588-611).  c[] = {A,B,C,D,E,F}; p[] = {delxSqr, ratioQtr, ratioSqr}
__device__ __forceinline__ void xd_update_std2dt(double *__restrict__ S,
    const double *__restrict__ A, const double *__restrict__ B, const double *__restrict__ C,
    const double *__restrict__ D, const double *__restrict__ E, const double *__restrict__ F,
    i64 nx, i64 j, i64 i, i64 ip, i64 im,
    double delxSqr, double ratioQtr, double ratioSqr, double optArg, double undef)
{
    const i64 c = j * nx + i, n = c + nx, s = c - nx;
    const i64 e = j * nx + ip, w = j * nx + im;
    const double Fc = F[c], An = A[n], Ac = A[c], Bn = B[n], Bs = B[s], Ce = C[e], Cw = C[w], De = D[e], Dc = D[c], Ec = E[c];
    const bool cond = (Fc != undef) & (An != undef) & (Ac != undef) & (Bn != undef) & (Bs != undef) &
                      (Ce != undef) & (Cw != undef) & (De != undef) & (Dc != undef) & (Ec != undef);
    if (!cond) return;
    const double Bq = (i == 0) ? B[n + 1] : Bn;
    const double Sc = S[c], Sn = S[n], Ss = S[s], Se = S[e], Sw = S[w];
    const double Sne = S[n - i + ip], Snw = S[n - i + im];
    const double Sse = S[s - i + ip], Ssw = S[s - i + im];
    const double Sse2 = (i == 0) ? Ss : Sse;
    const double t1 = (An * (Sn - Sc) - Ac * (Sc - Ss)) * ratioSqr;
    const double t2 = (Bq * (Sne - Snw) - Bs * (Sse2 - Ssw)) * ratioQtr;
    const double t3 = (Ce * (Sne - Sse) - Cw * (Snw - Ssw)) * ratioQtr;
    const double t4 = (De * (Se - Sc) - Dc * (Sc - Sw));
    double temp = (((t1 + t2) + t3) + t4) + (Ec * Sc - Fc) * delxSqr;
    temp = temp * (optArg / (((An + Ac) * ratioSqr + (De + Dc)) - Ec * delxSqr));
    S[c] = Sc + temp;
}

// invert_general_3D (numbas.py:745-984; interior :905-935, west :868-900 -- its condition names G twice and never
// tests H, :869 --, east :940-972).  c[] = {A..H}; p[] = {delx, delxSqr, ratio2, ratio1, ratio2Sqr, ratio1Sqr}
__device__ __forceinline__ void xd_update_gen3d(double *__restrict__ S, const XdCoef &q, i64 b,
    i64 ny, i64 nx, i64 k, i64 j, i64 i, i64 ip, i64 im)
{
    const i64 pl = ny * nx;
    const i64 row = k * pl + j * nx;
    const i64 c = row + i, e = row + ip, w = row + im;
    const i64 n = c + nx, s = c - nx, u = c + pl, d = c - pl;
    const double undef = q.undef;
    const double Ac = q.c[0][b * q.cs[0] + c], Bc = q.c[1][b * q.cs[1] + c], Cc = q.c[2][b * q.cs[2] + c];
    const double Dc = q.c[3][b * q.cs[3] + c], Ec = q.c[4][b * q.cs[4] + c], Fc = q.c[5][b * q.cs[5] + c];
    const double Gc = q.c[6][b * q.cs[6] + c], Hc = q.c[7][b * q.cs[7] + c];
    bool cond = (Gc != undef) & (Ac != undef) & (Bc != undef) & (Cc != undef) & (Dc != undef) & (Ec != undef) & (Fc != undef);
    if (i != 0) cond = cond & (Hc != undef);
    if (!cond) return;
    const double delx = q.p[0], delxSqr = q.p[1], ratio2 = q.p[2], ratio1 = q.p[3], ratio2Sqr = q.p[4], ratio1Sqr = q.p[5];
    const double Sc = S[c], Su = S[u], Sd = S[d], Sn = S[n], Ss = S[s], Se = S[e], Sw = S[w];
    double temp = Ac * ((Su - Sc) - (Sc - Sd)) * ratio2Sqr;
    temp = temp + Bc * ((Sn - Sc) - (Sc - Ss)) * ratio1Sqr;
    temp = temp + Cc * ((Se - Sc) - (Sc - Sw));
    temp = temp + ((Dc * (Su - Sd) * ratio2 + Ec * (Sn - Ss) * ratio1) + Fc * (Se - Sw)) * delx / 2.0;
    temp = temp + (Gc * Sc - Hc) * delxSqr;
    temp = temp * (q.optArg / (((Ac * ratio2Sqr + Bc * ratio1Sqr) + Cc) * 2.0 - Gc * delxSqr));
    S[c] = Sc + temp;
}

// invert_standard_1D (numbas.py:632-742; interior :703-710, west :694-700, east :713-719).  c[] = {A,B,F}; p[] = {delxSqr}
__device__ __forceinline__ void xd_update_std1d(double *__restrict__ S,
    const double *__restrict__ A, const double *__restrict__ B, const double *__restrict__ F,
    i64 i, i64 ip, i64 im, double delxSqr, double optArg, double undef)
{
    const double Fc = F[i], Ac = A[i], Ae = A[ip], Bc = B[i];
    if (!((Fc != undef) & (Ac != undef) & (Ae != undef) & (Bc != undef))) return;
    const double Sc = S[i];
    double temp = (Ae * (S[ip] - Sc) - Ac * (Sc - S[im])) / delxSqr + (Bc * Sc - Fc);
    temp = temp * (optArg / ((Ae + Ac) / delxSqr - Bc));
    S[i] = Sc + temp;
}

// invert_general_bih_2D (numbas.py:1204-1586): 13-point biharmonic.  c[] = {A..J}; p[] = {delxSSr, delxTr, delxSqr, ratio,
// ratioSSr, ratioQtr, ratioSqr}.  Nine colours 3*(j mod 3) + (i mod 3) (everything a cell reads lies within +-2 rows /
// columns); periodic-x with nx not a multiple of 3: the last two columns move to six extra colours.
__host__ __device__ __forceinline__ int xd_colour_bih(int wrapfix, i64 nx, i64 j, i64 i)
{
    if (wrapfix && i >= nx - 2) return 9 + 3 * (int)(i - (nx - 2)) + (int)(j % 3);
    return 3 * (int)(j % 3) + (int)(i % 3);
}
__device__ __forceinline__ i64 xd_wrapcol(i64 i, i64 nx) { i %= nx; return i < 0 ? i + nx : i; }
// Interior numbas.py:1438-1479; periodic edge columns :1348-1390 (0), :1392-1434 (1), :1481-1524 (nx-2), :1526-1569 (nx-1),
// with their quirks: the G term's operation order, and -- in the two east columns -- the B term's "two columns west"
// operand taken at the stale inner-loop index (columns nx-7 / nx-6, negative indices wrapping) instead of nx-4 / nx-3.
__device__ __forceinline__ void xd_update_bih(double *__restrict__ S, const XdCoef &q, i64 b, i64 nx, i64 j, i64 i, bool periodic)
{
    const i64 c = j * nx + i;
    double v[10];
    bool cond = true;
    #pragma unroll
    for (int m = 0; m < 10; ++m) { v[m] = q.c[m][b * q.cs[m] + c]; cond = cond & (v[m] != q.undef); }
    if (!cond) return;
    const double A = v[0], B = v[1], C = v[2], D = v[3], E = v[4], F = v[5], G = v[6], H = v[7], I = v[8], J = v[9];
    const double delxSSr = q.p[0], delxTr = q.p[1], delxSqr = q.p[2], ratio = q.p[3], ratioSSr = q.p[4], ratioQtr = q.p[5],
                 ratioSqr = q.p[6];
    const bool edge = periodic && (i < 2 || i >= nx - 2);
    i64 e1 = i + 1, e2 = i + 2, w1 = i - 1, w2 = i - 2, w2b = i - 2;
    if (periodic) {
        e1 = xd_wrapcol(e1, nx); e2 = xd_wrapcol(e2, nx); w1 = xd_wrapcol(w1, nx); w2 = xd_wrapcol(w2, nx); w2b = w2;
        if (i == nx - 2) w2b = xd_wrapcol(nx - 7, nx);
        if (i == nx - 1) w2b = xd_wrapcol(nx - 6, nx);
    }
    double *r0 = S + j * nx;
    const double *rp1 = r0 + nx, *rp2 = r0 + 2 * nx, *rm1 = r0 - nx, *rm2 = r0 - 2 * nx;
    const double Sc = r0[i];
    double temp = A * ((((rp2[i] - 4.0 * rp1[i]) + 6.0 * Sc) - 4.0 * rm1[i]) + rm2[i]) * ratioSSr;
    temp = temp + B * ((((((((rp2[e2] - 2.0 * rp2[i]) + rp2[w2b]) + -2.0 * r0[e2]) + 4.0 * Sc) - 2.0 * r0[w2b]) + rm2[e2]) -
                        2.0 * rm2[i]) + rm2[w2b]) * ratioSqr / 16.0;
    temp = temp + C * ((((r0[e2] - 4.0 * r0[e1]) + 6.0 * Sc) - 4.0 * r0[w1]) + r0[w2]);
    temp = temp + D * ((rp1[i] - Sc) - (Sc - rm1[i])) * ratioSqr * delxSqr;
    temp = temp + E * ((rp1[e1] - rm1[e1]) - (rp1[w1] - rm1[w1])) * ratioQtr * delxSqr;
    temp = temp + F * ((r0[e1] - Sc) - (Sc - r0[w1])) * delxSqr;
    if (edge) temp = temp + G * (rp1[i] - rm1[i]) * delxTr / 2.0 * ratio;
    else      temp = temp + G * (rp1[i] - rm1[i]) * delxTr * ratio / 2.0;
    temp = temp + H * (r0[e1] - r0[w1]) * delxTr / 2.0;
    temp = temp + (I * Sc - J) * delxSSr;
    temp = temp * (-q.optArg / ((((A * ratioSSr + C) * 6.0 + B * ratioSqr / 4.0) - (D * ratioSqr + F) * 2.0 * delxSqr) + I * delxSSr));
    r0[i] = Sc + temp;
}

// ---------------------------------------------------------------------------
// Deterministic block reduction of (sum, count): fixed shuffle tree inside a
// warp, fixed order across warps.  Result valid in thread 0.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void xd_block_reduce(double &sum, i64 &cnt, double *sm_sum, i64 *sm_cnt)
{
    #pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        sum += __shfl_down_sync(0xffffffffu, sum, o);
        cnt += __shfl_down_sync(0xffffffffu, cnt, o);
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int nw = (blockDim.x + 31) >> 5;
    if (lane == 0) { sm_sum[wid] = sum; sm_cnt[wid] = cnt; }
    __syncthreads();
    if (wid == 0) {
        sum = (lane < nw) ? sm_sum[lane] : 0.0;
        cnt = (lane < nw) ? sm_cnt[lane] : 0;
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            sum += __shfl_down_sync(0xffffffffu, sum, o);
            cnt += __shfl_down_sync(0xffffffffu, cnt, o);
        }
    }
}

// Per-slice loop state kept on the device between sweeps.
struct XdSliceState {
    double normPrev;
    double flags[3];      // overflow, last relative change, last loop index
    i64    loop;
    int    active;
    int    sweeps_done;   // sweeps actually executed on this slice
    // fused engine only (xinv_march2d.cuh):
    int    cur;           // which ping-pong buffer holds the slice's current psi
    int    nit;           // iterations the next pass runs on this slice (1..T)
    int    redo;          // next pass re-runs the final iteration(s); loop control already done
    int    pad_;
    double omega;         // XINV_ACCEL_CHEBYSHEV: relaxation factor of the slice's next half sweep (resident / cluster engines)
};

// Chebyshev acceleration of SOR (xinv.h, xinv_opts.accel): the factor after `omega`; `first`: omega is omega_0 = 1.
// One IEEE operation per step, the same on the host, on the device and in the test oracle.
__host__ __device__ __forceinline__ double xd_cheb_next(double omega, double rho2, bool first)
{
    double t = first ? rho2 / 2.0 : (rho2 * omega) / 4.0;
    t = 1.0 - t;
    return 1.0 / t;
}

// Loop control of the reference after each sweep: numbas.py:401-414 (2-D
// standard, with the norm==0 exit), :197-210 and :1186-1199 (no such exit).
__device__ __forceinline__ void xd_decide(XdSliceState &st, double sum, i64 cnt,
                                          double tol, i64 mxLoop, int zero_exit)
{
    double norm;
    if (cnt != 0) norm = sum / (double)cnt;
    else          norm = nan("");
    st.sweeps_done += 1;
    if (isnan(norm) || norm > 1e100) {
        st.flags[0] = 1.0;
        st.active = 0;
        return;
    }
    st.flags[1] = fabs(norm - st.normPrev) / st.normPrev;
    st.flags[2] = (double)st.loop;
    if (st.flags[1] < tol || st.loop >= mxLoop || (zero_exit && norm == 0.0)) {
        st.active = 0;
        return;
    }
    st.normPrev = norm;
    st.loop += 1;
}
