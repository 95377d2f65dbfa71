// xinv_resident.cuh -- XINV_ENGINE_RESIDENT: the whole solve of a SMALL 2-D slice in one CTA.
//
// Slices whose operands fit into the shared memory of one SM (a few thousand cells: the z-lat
// sections of invert_Eliassen, tests/test_Eliassen.py:15-232 of the reference -- 37 x 73 --, small
// Gill-Matsuno / Stommel / Poisson domains) are bound by launch latency on every other engine:
// one sweep is 3-6 kernel launches of ~2 us each for a few hundred nanoseconds of work.  Here a
// CTA stages psi and every coefficient array of its slice into shared memory once (bulk TMA
// copies, cp.async.bulk, completion on an mbarrier), then iterates without leaving the SM:
//   per sweep:  [y-extend rows] -> colour 0 .. ncol-1 in place (__syncthreads in between)
//               -> mean|psi| by a deterministic block reduction -> numbas.py:401-414 by thread 0
// until the slice stops or the launch's sweep budget is used up, and writes psi back.
// All stencils of the 2-D forms are handled (5-point, 9-point with B != 0: four colours, the
// general form; wrap-fix colours for periodic-x with odd nx): the per-cell updates are the very
// functions of the colour engine (xd_update_std2d / xd_update_gen2d, reference expressions
// numbas.py:344-369 / :1126-1153) applied to shared-memory pointers, the colouring is
// xd_colour() -- so the iterates are bit-identical to the colour engine's and the oracle's.
// A batch runs one slice per CTA (persistent over the batch), each stopping on its own test.
#pragma once
#include "xinv_colour_engine.cuh"
#include "xinv_march2d.cuh"          // mbarrier / bulk-copy helpers

#define XR_THREADS 512

__device__ __forceinline__ void xr_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(xf_smem_u32(dst)), "l"(src), "r"(bytes), "r"(xf_smem_u32(bar)) : "memory");
}

struct XrArgs {
    double *S;                // [batch][N]
    XdCoef q;
    XdGeom g;
    int narr;                 // arrays staged per slice besides psi (coefficients + forcing, NULL B left out)
    int slot[8];              // slot[m] = index of coefficient m in shared memory (1-based after psi), -1 = absent
    int batch;
    XdSliceState *st;
    int *nactive;
    double tol;
    i64 mxLoop;
    int zero_exit;
    int nsweeps;              // sweep budget of this launch
    int accel;                // XINV_ACCEL_CHEBYSHEV: omega varies from half sweep to half sweep (xinv.h)
    double rho2;
};

// PRE (standard form, no omega schedule, one more array fits): after staging, F is replaced in place by Fd and a
// factor array is formed once per launch (xd_update_std2d_pre) -- no undef test and no division in the sweeps -- and
// every array is kept colour-split in shared memory (XD_SPLIT_POS: a row's even columns, then its odd ones), so that
// the loads of a colour step are unit-stride across a warp instead of stride-2 (two-way bank conflicts on every LDS.64).
template <int KIND, bool HASB, bool PRE>
__global__ void __launch_bounds__(XR_THREADS, 1)
xr_resident_kernel(const XrArgs a)
{
    extern __shared__ __align__(128) unsigned char xr_smem[];
    const XdGeom &g = a.g;
    const i64 N = g.N;
    const i64 Np = (N + 1) & ~(i64)1;            // arrays start 16-byte aligned
    double *sm = reinterpret_cast<double *>(xr_smem);
    double *sS = sm;
    double *red_sum = sm + (size_t)(a.narr + 1 + (PRE ? 1 : 0)) * Np;   // [32]  (PRE: the factor array sits before it)
    i64 *red_cnt = reinterpret_cast<i64 *>(red_sum + 32);       // [32]
    uint64_t *bar = reinterpret_cast<uint64_t *>(red_cnt + 32);
    int *flag = reinterpret_cast<int *>(bar + 1);
    const int tid = threadIdx.x, nth = blockDim.x;
    unsigned phase = 0;
    if (tid == 0) { xf_mbar_init(bar, 1); xf_fence_barrier_init(); }
    __syncthreads();

    const i64 nx = g.nx, ny = g.ny;
    const int base = (g.scheme == 4) ? 4 : 2;
    const i64 half = (nx + 1) / 2;
    const double undef = a.q.undef;
    const int he = ((int)nx + 1) / 2;            // PRE: even columns of a row first, then the odd ones
    #define XR_POS(i) (PRE ? XD_SPLIT_POS((int)(i), he) : (int)(i))
    // thread layout of a colour step: W lanes per row, rpp rows at a time
    int W = ((int)half + 7) & ~7;
    if (W > nth) W = nth;
    const int ty = tid / W, tx = tid - ty * W;
    const int rpp = (nth / W) > 0 ? (nth / W) : 1;

    for (int b = blockIdx.x; b < a.batch; b += gridDim.x) {
        if (tid == 0) flag[0] = a.st[b].active;
        __syncthreads();
        const int active0 = flag[0];
        __syncthreads();
        if (!active0) continue;

        // ---- stage psi and the coefficient arrays of slice b (bulk TMA where 16-byte alignment allows) ----
        const double *src[9];
        double *dst[9];
        src[0] = a.S + (i64)b * N; dst[0] = sS;
        int na = 1;
        for (int m = 0; m < 8; ++m) {
            if (a.slot[m] < 0) continue;
            src[na] = a.q.c[m] + (i64)b * a.q.cs[m];
            dst[na] = sm + (size_t)a.slot[m] * Np;
            ++na;
        }
        const uint32_t nb16 = (uint32_t)((N / 2) * 16);          // bytes that can go as one bulk copy per array
        bool bulk = nb16 >= 16 && !PRE;                          // (PRE: permuted on the way in, by ordinary loads)
        for (int m = 0; m < na; ++m) bulk = bulk && ((reinterpret_cast<uintptr_t>(src[m]) & 15) == 0);
        if (bulk) {
            if (tid == 0) {
                xf_fence_proxy_async();          // the previous slice was read / written through the generic proxy
                xf_mbar_expect_tx(bar, nb16 * (uint32_t)na);
                for (int m = 0; m < na; ++m) xr_bulk_g2s(dst[m], src[m], nb16, bar);
            }
            if ((N & 1) && tid < na) dst[tid][N - 1] = src[tid][N - 1];     // the odd tail
            xf_mbar_wait(bar, phase);
            phase ^= 1u;
        } else if (PRE) {
            for (int p = tid; p < (int)N; p += nth) {
                const int j = p / (int)nx, i = p - j * (int)nx;
                const int pp = j * (int)nx + XD_SPLIT_POS(i, he);
                for (int m = 0; m < na; ++m) dst[m][pp] = src[m][p];
            }
        } else {
            for (int m = 0; m < na; ++m)
                for (i64 p = tid; p < N; p += nth) dst[m][p] = src[m][p];
        }
        __syncthreads();
        const double *cA = (a.slot[0] >= 0) ? sm + (size_t)a.slot[0] * Np : nullptr;
        const double *cB = (HASB && a.slot[1] >= 0) ? sm + (size_t)a.slot[1] * Np : nullptr;
        const double *c2 = (a.slot[2] >= 0) ? sm + (size_t)a.slot[2] * Np : nullptr;
        const double *c3 = (a.slot[3] >= 0) ? sm + (size_t)a.slot[3] * Np : nullptr;
        const double *c4 = (KIND == XD_GEN2D) ? sm + (size_t)a.slot[4] * Np : nullptr;
        const double *c5 = (KIND == XD_GEN2D) ? sm + (size_t)a.slot[5] * Np : nullptr;
        const double *c6 = (KIND == XD_GEN2D) ? sm + (size_t)a.slot[6] * Np : nullptr;

        double *sFac = sm + (size_t)(a.narr + 1) * Np;          // PRE only
        if (PRE) {
            double *sF = sm + (size_t)a.slot[3] * Np;
            for (int p = tid; p < (int)N; p += nth) {
                const int j = p / (int)nx, i = p - j * (int)nx;
                double fdv = __hiloint2double(XD_SKIP_HI, 0), fv = 0.0;
                const int pc = j * (int)nx + XD_SPLIT_POS(i, he);
                if (j >= 1 && j <= (int)ny - 2 && i >= g.i0 && i < g.i1) {
                    const int ip = (i == (int)nx - 1) ? 0 : i + 1, im = (i == 0) ? (int)nx - 1 : i - 1;
                    const int n = pc + (int)nx, s = pc - (int)nx, e = j * (int)nx + XD_SPLIT_POS(ip, he), w = j * (int)nx + XD_SPLIT_POS(im, he);
                    const double Fc = sF[pc], An = cA[n], Ac = cA[pc], Ce = c2[e], Cc = c2[pc];
                    bool cond = (Fc != undef) & (An != undef) & (Ac != undef) & (Ce != undef) & (Cc != undef);
                    if (HASB) cond = cond & (cB[e] != undef) & (cB[w] != undef) & (cB[n] != undef) & (cB[s] != undef);
                    if (cond) {
                        fdv = Fc * a.q.p[0];
                        fv = a.q.optArg / ((An + Ac) * a.q.p[2] + (Ce + Cc));
                    }
                }
                sF[pc] = fdv;                                // (a cell's condition reads F at the cell itself only)
                sFac[pc] = fv;
            }
            __syncthreads();
        }
        XdSliceState st_;
        if (tid == 0) st_ = a.st[b];
        double om = a.st[b].omega;                   // Chebyshev: factor of the next half sweep (every thread keeps it)
        bool om_first = (a.st[b].sweeps_done == 0);
        for (int sweep = 0; sweep < a.nsweeps; ++sweep) {
            double w0 = a.q.optArg, w1 = a.q.optArg;
            if (a.accel) { w0 = om; w1 = xd_cheb_next(om, a.rho2, om_first); om = xd_cheb_next(w1, a.rho2, false); om_first = false; }
            // ---- y-"extend" rows (numbas.py:284-310), as xd_extend_kernel ----
            if (g.bcy == XD_BC_EXTEND) {
                for (i64 i = tid; i < nx; i += nth) {
                    i64 s_ = i;
                    if (g.bcx != XD_BC_PERIODIC) { if (i == 0) s_ = 1; else if (i == nx - 1) s_ = nx - 2; }
                    const double v0 = sS[nx + XR_POS(s_)], v1 = sS[(ny - 2) * nx + XR_POS(s_)];
                    if (v0 != undef) sS[XR_POS(i)] = v0;
                    if (v1 != undef) sS[(ny - 1) * nx + XR_POS(i)] = v1;
                }
                __syncthreads();
            }
            // ---- the colours, in place.  Threads are laid out as rows of W lanes (W = cells of one colour in a row,
            //      rounded up to a multiple of 8): no division in the loop ----
            for (int colour = 0; colour < g.ncol; ++colour) {
                const double wq = (colour < base / 2) ? w0 : w1;
                if (colour >= base) {                              // wrap-fix colours: column nx-1 only
                    for (int j = 1 + tid; j < (int)ny - 1; j += nth) {
                        const i64 i = nx - 1;
                        if (i < g.i0 || i >= g.i1) continue;
                        if (xd_colour(g.scheme, g.wrapfix, nx, j, j, i) != colour) continue;
                        if (KIND == XD_STD2D && PRE)
                            xd_update_std2d_pre<HASB>(sS, cA, cB, c2, c3, sFac, (int)nx, he, j, (int)i, 0, (int)i - 1, a.q.p[1], a.q.p[2]);
                        else if (KIND == XD_STD2D)
                            xd_update_std2d<HASB>(sS, cA, cB, c2, c3, nx, j, i, 0, i - 1, a.q.p[0], a.q.p[1], a.q.p[2], wq, undef);
                        else
                            xd_update_gen2d<HASB>(sS, cA, cB, c2, c3, c4, c5, c6, nx, j, i, 0, i - 1, a.q.p[0], a.q.p[1], a.q.p[2],
                                                  a.q.p[3], a.q.p[4], wq, undef);
                    }
                } else {
                    // 4-colour scheme: only the rows of this colour's parity (every second row) are visited
                    const int jstep = (g.scheme == 4) ? 2 : 1;
                    const int jfirst = (g.scheme == 4) ? (((colour >> 1) & 1) ? 1 : 2) : 1;
                    for (int j = jfirst + jstep * ty; ty < rpp && j < (int)ny - 1; j += jstep * rpp) {     // (the last, partial row of threads idles)
                        if (tx >= (int)half) continue;
                        const int i = (g.scheme == 4) ? 2 * tx + (colour & 1) : 2 * tx + ((j + colour) & 1);
                        if (i < g.i0 || i >= g.i1) continue;
                        if (xd_colour(g.scheme, g.wrapfix, nx, j, j, i) != colour) continue;
                        const i64 ip = (i == (int)nx - 1) ? 0 : i + 1;
                        const i64 im = (i == 0) ? nx - 1 : i - 1;
                        if (KIND == XD_STD2D && PRE)
                            xd_update_std2d_pre<HASB>(sS, cA, cB, c2, c3, sFac, (int)nx, he, j, i, (int)ip, (int)im, a.q.p[1], a.q.p[2]);
                        else if (KIND == XD_STD2D)
                            xd_update_std2d<HASB>(sS, cA, cB, c2, c3, nx, j, i, ip, im, a.q.p[0], a.q.p[1], a.q.p[2], wq, undef);
                        else
                            xd_update_gen2d<HASB>(sS, cA, cB, c2, c3, c4, c5, c6, nx, j, i, ip, im, a.q.p[0], a.q.p[1], a.q.p[2],
                                                  a.q.p[3], a.q.p[4], wq, undef);
                    }
                }
                __syncthreads();
            }
            // ---- mean|psi| over psi != undef (numbas.py:1710-1728), loop control by thread 0 ----
            double sum = 0.0;
            i64 cnt = 0;
            for (i64 p = tid; p < N; p += nth) {
                const double v = sS[p];
                if (v != undef) { sum += fabs(v); cnt += 1; }
            }
            xd_block_reduce(sum, cnt, red_sum, red_cnt);
            if (tid == 0) {
                xd_decide(st_, sum, cnt, a.tol, a.mxLoop, a.zero_exit);
                flag[0] = st_.active;
            }
            __syncthreads();
            const int go_on = flag[0];
            __syncthreads();
            if (!go_on) break;
        }
        // ---- psi back to HBM, state back ----
        double *out = a.S + (i64)b * N;
        if (PRE) {
            for (int p = tid; p < (int)N; p += nth) {
                const int j = p / (int)nx, i = p - j * (int)nx;
                out[p] = sS[j * (int)nx + XD_SPLIT_POS(i, he)];
            }
        } else {
            for (i64 p = tid; p < N; p += nth) out[p] = sS[p];
        }
        if (tid == 0) {
            st_.omega = om;
            a.st[b] = st_;
            if (!st_.active) atomicSub(a.nactive, 1);
        }
        xf_fence_proxy_async();              // this slice's generic-proxy accesses before the next slice's bulk copies
        __syncthreads();
    }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
struct ResidentPlan {
    bool built = false;
    XrArgs args{};
    size_t smem = 0, smem_pre = 0;   // without / with the factor array of the PRE kernels
    int grid = 0;
    int kind = 0;
    bool hasB = false;
    bool pre_ok = false;             // the PRE kernel fits (standard form)
};

static inline void resident_plan_release(ResidentPlan &p) { p = ResidentPlan(); }

template <int KIND, bool HASB, bool PRE>
static cudaError_t xr_prepare(size_t smem, int *blocks_per_sm)
{
    cudaError_t e = cudaFuncSetAttribute(xr_resident_kernel<KIND, HASB, PRE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, xr_resident_kernel<KIND, HASB, PRE>, XR_THREADS, smem);
}

// Does a slice with all its operands fit into the shared memory of one SM?
static inline int resident_plan_build(ResidentPlan &p, int sm_count, int kind, bool hasB, const XdGeom &g, const XdCoef &q,
                                      i64 batch, double *dS, std::string &why)
{
    resident_plan_release(p);
    if (kind != XD_STD2D && kind != XD_GEN2D) { why = "2-D problems only"; return -1; }
    if (g.ny < 3 || g.nx < 3) { why = "grid too small"; return -1; }
    const int ncoef = (kind == XD_STD2D) ? 4 : 7;
    XrArgs &a = p.args;
    a.narr = 0;
    for (int m = 0; m < 8; ++m) a.slot[m] = -1;
    for (int m = 0; m < ncoef; ++m) {
        if (!q.c[m]) continue;                   // B == NULL
        a.slot[m] = ++a.narr;
    }
    const i64 Np = (g.N + 1) & ~(i64)1;
    p.smem = (size_t)(a.narr + 1) * Np * sizeof(double) + 32 * 16 + 64;
    if (p.smem > 220 * 1024) { why = "slice does not fit into shared memory"; return -1; }
    int bps = 0;
    cudaError_t e;
    if (kind == XD_STD2D) e = hasB ? xr_prepare<XD_STD2D, true, false>(p.smem, &bps) : xr_prepare<XD_STD2D, false, false>(p.smem, &bps);
    else                  e = hasB ? xr_prepare<XD_GEN2D, true, false>(p.smem, &bps) : xr_prepare<XD_GEN2D, false, false>(p.smem, &bps);
    p.smem_pre = p.smem + (size_t)Np * sizeof(double);
    if (e == cudaSuccess && bps >= 1 && kind == XD_STD2D && p.smem_pre <= 220 * 1024 && g.N < ((i64)1 << 30)) {
        int bps2 = 0;
        cudaError_t e2 = hasB ? xr_prepare<XD_STD2D, true, true>(p.smem_pre, &bps2) : xr_prepare<XD_STD2D, false, true>(p.smem_pre, &bps2);
        p.pre_ok = (e2 == cudaSuccess && bps2 >= 1);
        if (!p.pre_ok) (void)cudaGetLastError();
    }
    if (e != cudaSuccess || bps < 1) { why = std::string("resident kernel does not fit: ") + cudaGetErrorString(e); (void)cudaGetLastError(); return -1; }
    a.S = dS; a.q = q; a.g = g; a.batch = (int)batch;
    const i64 slots = (i64)sm_count * bps;
    p.grid = (int)(batch < slots ? batch : slots);
    p.kind = kind; p.hasB = hasB;
    p.built = true;
    return 0;
}

static inline int resident_sweep(ResidentPlan &p, cudaStream_t stream, XdSliceState *st, int *nactive, double tol, i64 mxLoop,
                                 int zero_exit, int nsweeps, int accel, double rho2, int64_t *launches)
{
    XrArgs &a = p.args;
    a.st = st; a.nactive = nactive; a.tol = tol; a.mxLoop = mxLoop; a.zero_exit = zero_exit; a.nsweeps = nsweeps;
    a.accel = accel; a.rho2 = rho2;
    if (p.kind == XD_STD2D && p.pre_ok && !accel) {
        if (p.hasB) xr_resident_kernel<XD_STD2D, true, true><<<p.grid, XR_THREADS, p.smem_pre, stream>>>(a);
        else        xr_resident_kernel<XD_STD2D, false, true><<<p.grid, XR_THREADS, p.smem_pre, stream>>>(a);
    } else if (p.kind == XD_STD2D) {
        if (p.hasB) xr_resident_kernel<XD_STD2D, true, false><<<p.grid, XR_THREADS, p.smem, stream>>>(a);
        else        xr_resident_kernel<XD_STD2D, false, false><<<p.grid, XR_THREADS, p.smem, stream>>>(a);
    } else {
        if (p.hasB) xr_resident_kernel<XD_GEN2D, true, false><<<p.grid, XR_THREADS, p.smem, stream>>>(a);
        else        xr_resident_kernel<XD_GEN2D, false, false><<<p.grid, XR_THREADS, p.smem, stream>>>(a);
    }
    *launches += 1;
    return cudaGetLastError() == cudaSuccess ? 0 : -1;
}
