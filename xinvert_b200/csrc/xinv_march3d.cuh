// xinv_march3d.cuh -- XINV_ENGINE_FUSED for the 3-D standard form (invert_omega;
// numbas.py:15-212): ONE pass over the volume performs a complete red+black SOR
// iteration, the y-"extend" rows, sum|omega| / count and the loop control.
//
// Design ("plane marching"): a CTA owns a column of tiles -- TJ rows x 64 columns of
// every level -- and marches along z.  One warp per tile row, one column pair per lane
// (as in the 2-D engine: x-neighbours by warp shuffle).
//   * Once per solve the engine builds padded copies of the operands (same layout as the
//     2-D engine: XM_PADL ghost columns left, >= XM_GHOST right, holding the periodic
//     wrap-around neighbours) and two derived arrays, with the reference's own operations:
//        Fd  = F * delxSqr, or a marker where the cell must never be updated (boundary
//              level / row / fixed column, an undef operand: numbas.py:117-118, :147-150)
//        fac = optArg / ((A[k+1]+A[k])*ratio2Sqr + (B[j+1]+B[j])*ratio1Sqr + (C[i+1]+C[i]))
//                                                                     (numbas.py:166-168)
//   * Thread 0 feeds a K-stage shared-memory ring with TMA box loads: per level one box of
//     omega, A, C, Fd, fac (TJ x 64) and B (TJ+1 x 64: the row north of the tile too),
//     completion on one mbarrier per stage, K-1 levels ahead of the consumers.
//   * z-neighbours live in registers (the lane keeps its column pair of the last three
//     levels and A of the last two); y-neighbours of the red half step are read from the
//     staged (still untouched) level, those of the black half step from a small exchange
//     buffer into which every warp publishes its row after the red half step.
//   * Schedule per step s (level s has just arrived): red cells of level s-1 (colour 0 =
//     (i+j+k) even, as the colour engine and the oracle), black cells of level s-2, which
//     is then complete: norm accumulation, store to the OTHER omega buffer (ping-pong:
//     neighbouring tiles still need the old values).  One __syncthreads per step.
//   * A tile has a 2-cell halo in y and x (red results of the halo are recomputed; black
//     results only exist for the owned TJ-4 rows x 60 columns) and none in z.
//   * Per-tile (sum, count) partials are combined in fixed order by the CTA that finishes
//     a slice last (atomic ticket), which then runs numbas.py:197-210.
// HBM/L2 traffic per pass (= per iteration): omega r+w, A, B, C, Fd, fac = 56 N bytes
// (x the halo overhead of the tiling), of which 48 N are algorithmic (omega r+w, A, B, C, F).
#pragma once
#include "xinv_march2d.cuh"

#define X3_W 64

struct X3Args {
    double *Sbuf[2];          // padded omega buffers [batch][nz][ny][pitch]
    i64 pitch, plane, slice;  // plane = ny * pitch, slice = nz * plane
    int nz, ny, nx;
    int ntx, nty, RB;         // column tiles, row tiles, owned rows per tile (TJ - 4)
    int batch;
    int bcy, bcx;
    int cbA, cbB, cbC, cbFd, cbFac;   // 1: the array has a batch axis, 0: one volume shared by the batch
    double r2, r1, undef;     // ratio2Sqr, ratio1Sqr
    XdSliceState *st;
    double *psum;             // [batch][ntx*nty]
    i64 *pcnt;
    unsigned *ticket;
    int *nactive;
    double tol;
    i64 mxLoop;
    int npass;
    unsigned long long *gbar;
    unsigned long long gbar_base;
};

// numbas.py:153-169, operation for operation (cf. xd_update_std3d), with F * delxSqr and
// optArg / denominator taken from the precomputed arrays
__device__ __forceinline__ double x3_cell(double Sc, double Su, double Sd, double Sn, double Ss, double Se, double Sw,
                                          double Au, double Ac, double Bn, double Bc, double Ce, double Cc,
                                          double Fd, double fac, double r2, double r1)
{
    double temp = ((Au * (Su - Sc) - Ac * (Sc - Sd)) * r2 + (Bn * (Sn - Sc) - Bc * (Sc - Ss)) * r1 +
                   (Ce * (Se - Sc) - Cc * (Sc - Sw))) - Fd;
    temp = temp * fac;
    const bool upd = __double2hiint(Fd) != XM_SKIP_HI;
    const double nv = Sc + temp;
    return upd ? nv : Sc;
}

__device__ __forceinline__ double2 x3_ld2(const double *p) { return *reinterpret_cast<const double2 *>(p); }

template <int TJ, int K>
__global__ void __launch_bounds__(TJ * 32, 1)
xm3_std3d_kernel(const __grid_constant__ CUtensorMap mS0, const __grid_constant__ CUtensorMap mS1,
                 const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                 const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mFd,
                 const __grid_constant__ CUtensorMap mFac, const X3Args a)
{
    constexpr int W = X3_W;
    constexpr int TILE = TJ * W;                 // doubles per array per stage
    constexpr int OFF_S = 0, OFF_A = TILE, OFF_C = 2 * TILE, OFF_FD = 3 * TILE, OFF_FAC = 4 * TILE, OFF_B = 5 * TILE;
    constexpr int STAGE = 5 * TILE + (TJ + 1) * W;
    constexpr uint32_t STAGE_BYTES = STAGE * sizeof(double);

    extern __shared__ __align__(1024) unsigned char x3_smem[];
    double *ring = reinterpret_cast<double *>(x3_smem);
    double *X = ring + (size_t)K * STAGE;        // two exchange buffers of TILE doubles
    double *red_sum = X + 2 * TILE;              // [TJ]
    i64 *red_cnt = reinterpret_cast<i64 *>(red_sum + TJ);   // [TJ]
    uint64_t *bars = reinterpret_cast<uint64_t *>(red_cnt + TJ);   // [K]
    int *box = reinterpret_cast<int *>(bars + K);            // [4] CTA-wide broadcasts

    const int lane = threadIdx.x & 31;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // tile row of this warp (warp-uniform)
    if (threadIdx.x == 0) {
        #pragma unroll
        for (int s = 0; s < K; ++s) xf_mbar_init(&bars[s], 1);
        xf_fence_barrier_init();
    }
    __syncthreads();

    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const bool periodic = (a.bcx == XD_BC_PERIODIC);
    const bool extend = (a.bcy == XD_BC_EXTEND);
    const int tps = a.ntx * a.nty;               // tiles per slice
    const int total = tps * a.batch;
    const double undef = a.undef, r2 = a.r2, r1 = a.r1;
    const double2 zero2 = make_double2(0.0, 0.0);
    unsigned q0 = 0;                             // levels consumed by this CTA so far (ring position)

    for (int pp = 0; pp < a.npass; ++pp) {
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tps;
        const int tidx = tile - b * tps;
        const int yb = tidx / a.ntx, xb = tidx - yb * a.ntx;
        // slice state: constant while any tile of the slice is still to do in this pass
        // (read through L2: another SM rewrites it between two passes of one launch)
        if (threadIdx.x == 0) { box[0] = __ldcg(&a.st[b].active); box[1] = __ldcg(&a.st[b].cur); }
        __syncthreads();
        const int active = box[0], cur = box[1];
        __syncthreads();                         // box is rewritten by the next tile
        if (!active) continue;                   // frozen slice

        const int x0 = xb * (W - 4), y0 = yb * a.RB;
        const int j = y0 - 2 + w;                // global row of this warp
        const int gx = x0 - 2 + 2 * lane;        // global (even) column of this lane's pair
        const int bx = x0 - 2 + XM_PADL;         // padded x coordinate of the tile's first column
        const CUtensorMap *mS = cur ? &mS1 : &mS0;
        const bool own_row = (w >= 2) & (w < TJ - 2) & (j < ny);
        const bool own_lane = (lane >= 1) & (lane < 31);
        const bool own_x = own_row & own_lane & (gx < nx);
        const bool own_y = own_row & own_lane & (gx + 1 < nx);
        const bool ghe = own_y & periodic & (gx < XM_GHOST);          // also write the east ghost copy
        const bool ghw = own_y & periodic & (gx >= nx - XM_GHOST);    // also write the west ghost copy
        const int wn = (w + 1 < TJ) ? W : 0, ws = (w > 0) ? -W : 0;  // rows beyond the tile only ever feed halo rows
        double *const outS = a.Sbuf[cur ^ 1] + (i64)b * a.slice + (i64)j * a.pitch + XM_PADL + gx;
        const int rowoff = w * W + 2 * lane;

        auto issue = [&](int k) {                // thread 0: TMA loads of level k of this tile
            const unsigned q = q0 + (unsigned)k;
            double *dst = ring + (size_t)(q % K) * STAGE;
            uint64_t *bar = &bars[q % K];
            const int ys = y0 - 2;
            xf_mbar_expect_tx(bar, STAGE_BYTES);
            xf_tma_load_3d(dst + OFF_S, mS, bar, bx, ys, b * nz + k);
            xf_tma_load_3d(dst + OFF_A, &mA, bar, bx, ys, b * a.cbA * nz + k);
            xf_tma_load_3d(dst + OFF_C, &mC, bar, bx, ys, b * a.cbC * nz + k);
            xf_tma_load_3d(dst + OFF_FD, &mFd, bar, bx, ys, b * a.cbFd * nz + k);
            xf_tma_load_3d(dst + OFF_FAC, &mFac, bar, bx, ys, b * a.cbFac * nz + k);
            xf_tma_load_3d(dst + OFF_B, &mB, bar, bx, ys, b * a.cbB * nz + k);     // TJ+1 rows
        };
        if (threadIdx.x == 0) {
            // every thread has finished reading the ring (barriers of the previous tile); order those
            // generic-proxy reads before the async-proxy writes of the new loads
            xf_fence_proxy_async();
            const int pre = (nz < K - 1) ? nz : K - 1;
            for (int k = 0; k < pre; ++k) issue(k);
        }

        double2 P1 = zero2, P2 = zero2, P3 = zero2;      // omega of levels s-1, s-2, s-3
        double2 A1 = zero2, A2 = zero2;                  // A of levels s-1, s-2
        double bBc = 0.0, bBn = 0.0, bCw = 0.0, bCe = 0.0, bFac = 0.0;   // operands of the black cell of level s-2
        double bFd = xm_skip_value();
        double nsum = 0.0;
        int ncnt = 0;

        for (int s = 0; s < nz + 2; ++s) {
            // ---- level s arrives ----
            double2 Pn = zero2, An = zero2;
            if (s < nz) {
                const unsigned q = q0 + (unsigned)s;
                xf_mbar_wait(&bars[q % K], (q / K) & 1u);
                const double *g = ring + (size_t)(q % K) * STAGE + rowoff;
                Pn = x3_ld2(g + OFF_S);
                An = x3_ld2(g + OFF_A);
                if (extend & (s >= 1) & (s <= nz - 2)) {           // numbas.py:87-115: levels 1..nz-2 only
                    if (j == 0) Pn = xm_extend(Pn, x3_ld2(g + OFF_S + W), gx, nx, periodic, undef);
                    if (j == ny - 1) Pn = xm_extend(Pn, x3_ld2(g + OFF_S - W), gx, nx, periodic, undef);
                }
            }
            // ---- red cells of level s-1 (neighbours in y: the staged level, still untouched) ----
            const int kr = s - 1;
            // operands of the black cell of level s-1: used by the black half step of the NEXT step
            double nBc = 0.0, nBn = 0.0, nCw = 0.0, nCe = 0.0, nFac = 0.0, nFd = xm_skip_value();
            if ((kr >= 1) & (kr <= nz - 2)) {
                const unsigned q = q0 + (unsigned)kr;
                const double *r = ring + (size_t)(q % K) * STAGE + rowoff;
                const double2 Bc = x3_ld2(r + OFF_B), Bn = x3_ld2(r + OFF_B + W), Cc = x3_ld2(r + OFF_C);
                const double2 Fd = x3_ld2(r + OFF_FD), Fc = x3_ld2(r + OFF_FAC);
                double2 Sn = x3_ld2(r + OFF_S + wn), Ss = x3_ld2(r + OFF_S + ws);
                if (extend) {
                    // the extended boundary row as the cells of rows 1 / ny-2 see it: their own old value
                    if (j == 1) { if (P1.x != undef) Ss.x = P1.x; if (P1.y != undef) Ss.y = P1.y; }
                    if (j == ny - 2) { if (P1.x != undef) Sn.x = P1.x; if (P1.y != undef) Sn.y = P1.y; }
                }
                const double Cnext = xm_shfl_down1(Cc.x);          // C of the column east of the pair
                if (((j + kr) & 1) == 0) {                         // red = the even column of the pair
                    const double nb = xm_shfl_up1(P1.y);
                    P1.x = x3_cell(P1.x, Pn.x, P2.x, Sn.x, Ss.x, P1.y, nb, An.x, A1.x, Bn.x, Bc.x, Cc.y, Cc.x, Fd.x, Fc.x, r2, r1);
                    nBc = Bc.y; nBn = Bn.y; nCw = Cc.y; nCe = Cnext; nFd = Fd.y; nFac = Fc.y;
                } else {
                    const double nb = xm_shfl_down1(P1.x);
                    P1.y = x3_cell(P1.y, Pn.y, P2.y, Sn.y, Ss.y, nb, P1.x, An.y, A1.y, Bn.y, Bc.y, Cnext, Cc.y, Fd.y, Fc.y, r2, r1);
                    nBc = Bc.x; nBn = Bn.x; nCw = Cc.x; nCe = Cc.y; nFd = Fd.x; nFac = Fc.x;
                }
            }
            // publish the row (red cells final for this iteration) for the black half step of the next step
            *reinterpret_cast<double2 *>(X + (size_t)(s & 1) * TILE + rowoff) = P1;
            // ---- black cells of level s-2 (neighbours in y: rows published in the previous step) ----
            const int kb = s - 2;
            if ((kb >= 1) & (kb <= nz - 2)) {
                const double *xr = X + (size_t)((s - 1) & 1) * TILE + rowoff;
                const double2 Sn = x3_ld2(xr + wn), Ss = x3_ld2(xr + ws);
                if (((j + kb) & 1) == 1) {                         // black = the even column of the pair
                    const double nb = xm_shfl_up1(P2.y);
                    P2.x = x3_cell(P2.x, P1.x, P3.x, Sn.x, Ss.x, P2.y, nb, A1.x, A2.x, bBn, bBc, bCe, bCw, bFd, bFac, r2, r1);
                } else {
                    const double nb = xm_shfl_down1(P2.x);
                    P2.y = x3_cell(P2.y, P1.y, P3.y, Sn.y, Ss.y, nb, P2.x, A1.y, A2.y, bBn, bBc, bCe, bCw, bFd, bFac, r2, r1);
                }
            }
            // ---- level s-2 is complete: norm over owned cells (numbas.py:1689-1708), store ----
            if ((kb >= 0) & (kb <= nz - 1)) {
                xm_norm_acc_lane(nsum, ncnt, P2.x, own_x, undef);
                xm_norm_acc_lane(nsum, ncnt, P2.y, own_y, undef);
                if ((kb >= 1) & (kb <= nz - 2)) {                  // levels 0 and nz-1 never change
                    double *dst = outS + (i64)kb * a.plane;
                    xm_store2_if(own_y, dst, P2);
                    xm_store1_if(own_x & !own_y, dst, P2.x);       // odd nx: the last column stands alone
                    xm_store2_if(ghe, dst + nx, P2);
                    xm_store2_if(ghw, dst - nx, P2);
                }
            }
            __syncthreads();                     // rows published; stage of level s-1 free
            if (threadIdx.x == 0 && s + K - 1 < nz) { xf_fence_proxy_async(); issue(s + K - 1); }
            P3 = P2; P2 = P1; P1 = Pn;
            A2 = A1; A1 = An;
            bBc = nBc; bBn = nBn; bCw = nCw; bCe = nCe; bFd = nFd; bFac = nFac;
        }
        q0 += (unsigned)nz;

        // ---- per-tile norm partial, ticket, loop control by the last tile of the slice ----
        #pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            nsum += __shfl_down_sync(0xffffffffu, nsum, o);
            ncnt += __shfl_down_sync(0xffffffffu, ncnt, o);
        }
        if (lane == 0) { red_sum[w] = nsum; red_cnt[w] = (i64)ncnt; }
        __syncthreads();
        if (threadIdx.x == 0) {
            double ts = 0.0;
            i64 tc = 0;
            #pragma unroll
            for (int r = 0; r < TJ; ++r) { ts += red_sum[r]; tc += red_cnt[r]; }
            a.psum[(i64)b * tps + tidx] = ts;
            a.pcnt[(i64)b * tps + tidx] = tc;
            __threadfence();
            const unsigned tk = atomicAdd(&a.ticket[b], 1u);
            box[2] = (tk == (unsigned)tps - 1u);
        }
        __syncthreads();
        const int last = box[2];
        if (last && w == 0) {
            __threadfence();
            // fixed assignment of partials to lanes and a fixed shuffle tree: the sum does not depend on
            // which tile happened to finish last
            double s_ = 0.0;
            i64 c_ = 0;
            for (int p = lane; p < tps; p += 32) {
                s_ += __ldcg(a.psum + (i64)b * tps + p);
                c_ += __ldcg(a.pcnt + (i64)b * tps + p);
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                s_ += __shfl_down_sync(0xffffffffu, s_, o);
                c_ += __shfl_down_sync(0xffffffffu, c_, o);
            }
            if (lane == 0) {
                XdSliceState st_ = a.st[b];
                xd_decide(st_, s_, c_, a.tol, a.mxLoop, 0);       // no norm == 0 exit in 3-D (numbas.py:206)
                st_.cur ^= 1;
                a.st[b] = st_;
                a.ticket[b] = 0u;
                if (!st_.active) atomicSub(a.nactive, 1);
            }
        }
        __syncthreads();                         // box[2] / reduction scratch are rewritten by the next tile
    }
    // ---- grid-wide barrier before the next pass of this launch (cooperative launch) ----
    if (pp + 1 < a.npass) {
        __syncthreads();
        int go_on = 1;
        if (threadIdx.x == 0) {
            __threadfence();
            atomicAdd(a.gbar, 1ULL);
            const unsigned long long want = a.gbar_base + (unsigned long long)(pp + 1) * gridDim.x;
            unsigned long long seen;
            do {
                asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(seen) : "l"(a.gbar) : "memory");
            } while (seen < want);
            int na;
            asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(na) : "l"(a.nactive) : "memory");
            go_on = (na != 0);
            if (!go_on && pp + 2 < a.npass) atomicAdd(a.gbar, (unsigned long long)(a.npass - 2 - pp));
        }
        go_on = __syncthreads_or(go_on && threadIdx.x == 0);
        if (!go_on) break;
        asm volatile("fence.proxy.async.global;" ::: "memory");
    }
    }
}

// ----------------------------------------------------------------------------
// dense <-> padded layout (rows = nz * ny of every volume; grid.y strides over them)
// ----------------------------------------------------------------------------
__global__ void x3_pack_kernel(double *__restrict__ dst, const double *__restrict__ src, i64 rows, i64 nx, i64 pitch,
                               i64 src_bstride, i64 nb, int periodic)
{
    const i64 pc = (i64)blockIdx.x * blockDim.x + threadIdx.x;     // padded column
    if (pc >= pitch) return;
    const i64 i = pc - XM_PADL;
    for (i64 row = blockIdx.y; row < rows * nb; row += gridDim.y) {
        const i64 b = row / rows, r = row - b * rows;
        const double *s = src + b * src_bstride + r * nx;
        double v = 0.0;
        if (i >= 0 && i < nx) v = s[i];
        else if (periodic && i >= -XM_GHOST && i < nx + XM_GHOST) v = s[((i % nx) + nx) % nx];
        dst[row * pitch + pc] = v;
    }
}

// Fd and fac of the padded layout (see the header).  Fd exists for nbFd volumes, fac for nbFac <= nbFd.
// Plain IEEE operations (-fmad=false), div.rn.f64: bit-identical to what the reference computes.
__global__ void x3_pack_derived_kernel(double *__restrict__ Fd, double *__restrict__ fac, XdCoef q, i64 nz, i64 ny, i64 nx,
                                       i64 pitch, i64 nbFd, i64 nbFac, int periodic)
{
    const i64 pc = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (pc >= pitch) return;
    const i64 i = pc - XM_PADL;
    const i64 rows = nz * ny;
    bool col = true;
    i64 iw = i, ie = i + 1;
    if (periodic) {
        col = (i >= -XM_GHOST) && (i < nx + XM_GHOST);
        iw = ((i % nx) + nx) % nx;
        ie = (iw + 1 == nx) ? 0 : iw + 1;
    } else {
        col = (i >= 1) && (i <= nx - 2);
    }
    const double ratio2Sqr = q.p[1], ratio1Sqr = q.p[2], delxSqr = q.p[0];
    for (i64 row = blockIdx.y; row < rows * nbFd; row += gridDim.y) {
        const i64 b = row / rows, r = row - b * rows;
        const i64 k = r / ny, j = r - k * ny;
        const bool cell = col && (k >= 1) && (k <= nz - 2) && (j >= 1) && (j <= ny - 2);
        double vF = __hiloint2double(XM_SKIP_HI, 0), vf = 0.0;
        if (cell) {
            const i64 o = (k * ny + j) * nx;
            const double *A = q.c[0] + b * q.cs[0], *B = q.c[1] + b * q.cs[1], *C = q.c[2] + b * q.cs[2];
            const double Au = A[o + ny * nx + iw], Ac = A[o + iw], Bn = B[o + nx + iw], Bc = B[o + iw];
            const double Ce = C[o + ie], Cc = C[o + iw];
            const double Fc = q.c[3][b * q.cs[3] + o + iw];
            if ((Fc != q.undef) & (Au != q.undef) & (Ac != q.undef) & (Bn != q.undef) & (Bc != q.undef) &
                (Ce != q.undef) & (Cc != q.undef))
                vF = Fc * delxSqr;
            vf = q.optArg / ((Au + Ac) * ratio2Sqr + (Bn + Bc) * ratio1Sqr + (Ce + Cc));
        }
        Fd[row * pitch + pc] = vF;
        if (b < nbFac) fac[row * pitch + pc] = vf;
    }
}

__global__ void x3_unpack_kernel(double *__restrict__ dst, const double *__restrict__ buf0,
                                 const double *__restrict__ buf1, i64 rows, i64 nx, i64 pitch, i64 nb,
                                 const XdSliceState *__restrict__ st)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    for (i64 row = blockIdx.y; row < rows * nb; row += gridDim.y) {
        const i64 b = row / rows;
        const double *src = st[b].cur ? buf1 : buf0;
        dst[row * nx + i] = src[row * pitch + XM_PADL + i];
    }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
struct X3Variant { int TJ, K; };
static const X3Variant X3_VARIANTS[] = {
    {16, 4},   // 0: 12 owned rows per tile, 512 threads, 210 KB
    {12, 4},   // 1:  8 owned rows, 384 threads, 158 KB
    {8, 4},    // 2:  4 owned rows, 256 threads, 106 KB (tiny grids)
    {16, 3},   // 3: shallower ring
    {12, 3},   // 4
};
#define X3_NVARIANTS ((int)(sizeof(X3_VARIANTS) / sizeof(X3_VARIANTS[0])))

struct Fused3Plan {
    bool built = false;
    int variant = 0;
    bool coop = false;
    int ppl = 1;
    unsigned long long gbar_base = 0;
    void *bufS[2] = {nullptr, nullptr};
    void *bufA = nullptr, *bufB = nullptr, *bufC = nullptr, *bufFd = nullptr, *bufFac = nullptr;
    CUtensorMap mS[2], mA, mB, mC, mFd, mFac;
    X3Args args{};
    i64 batch = 0;
    int nblk_partials = 0;
    size_t smem = 0;
    int grid = 0;
};

static inline void fused3_plan_release(Fused3Plan &p) { p = Fused3Plan(); }

static inline bool fused3_plan_supported(const XdGeom &g, std::string &why)
{
    if (g.wrapfix) { why = "periodic-x with odd nx needs the wrap-fix colours"; return false; }
    if (g.nz < 3 || g.ny < 3 || g.nx < 4) { why = "grid too small"; return false; }
    if (g.ny > 0x3ffffff0 || g.nx > 0x3ffffff0 || g.nz > 0x3ffffff0) { why = "grid too large"; return false; }
    return true;
}

template <int TJ, int K>
static size_t x3_smem_bytes()
{
    const size_t stage = (size_t)(5 * TJ * X3_W + (TJ + 1) * X3_W) * sizeof(double);
    return (size_t)K * stage + (size_t)2 * TJ * X3_W * sizeof(double) + (size_t)TJ * 16 + (size_t)K * 8 + 16;
}
template <int TJ, int K>
static cudaError_t x3_prepare(size_t *smem, int *blocks_per_sm)
{
    *smem = x3_smem_bytes<TJ, K>();
    cudaError_t e = cudaFuncSetAttribute(xm3_std3d_kernel<TJ, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, xm3_std3d_kernel<TJ, K>, TJ * 32, *smem);
}
template <int TJ, int K>
static cudaError_t x3_launch(const Fused3Plan &p, cudaStream_t stream)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)p.grid);
    cfg.blockDim = dim3(TJ * 32);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeCooperative;
    attr[0].val.cooperative = (p.args.npass > 1) ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, xm3_std3d_kernel<TJ, K>, p.mS[0], p.mS[1], p.mA, p.mB, p.mC, p.mFd, p.mFac, p.args);
}

#define X3_DISPATCH(v, CALL)              \
    switch (v) {                          \
    case 0: CALL(16, 4); break;           \
    case 1: CALL(12, 4); break;           \
    case 2: CALL(8, 4); break;            \
    case 3: CALL(16, 3); break;           \
    default: CALL(12, 3); break;          \
    }

// tile height: as few halo rows as possible while the tiles still fill the SMs evenly
static int x3_choose_variant(i64 ny, i64 nx, i64 batch, int sm_count)
{
    const i64 ntx = (nx + X3_W - 5) / (X3_W - 4);
    double best = -1.0;
    int bestv = 0;
    for (int v = 0; v < 3; ++v) {
        const int RB = X3_VARIANTS[v].TJ - 4;
        const i64 tiles = ntx * ((ny + RB - 1) / RB) * batch;
        const i64 rounds = (tiles + sm_count - 1) / sm_count;
        // useful rows per tile row x how evenly the tiles fill the SMs; a taller tile also spends more
        // shared-memory and FP64 cycles per step on one SM, which the first factor already prices in
        const double eff = ((double)RB / (double)X3_VARIANTS[v].TJ) * ((double)tiles / (double)(rounds * sm_count));
        if (eff > best + 1e-9) { best = eff; bestv = v; }
    }
    return bestv;
}

static inline int fused3_plan_build(Fused3Plan &p, XmWork &work, int sm_count, const XdGeom &g, const XdCoef &q, i64 batch,
                                    double *dS, cudaStream_t stream, std::string &why)
{
    fused3_plan_release(p);
    const i64 nz = g.nz, ny = g.ny, nx = g.nx;
    const i64 pitch = ((XM_PADL + nx + XM_GHOST) + 3) / 4 * 4;
    const int periodic = (g.bcx == XD_BC_PERIODIC);
    const i64 rows = nz * ny;
    const size_t vol_bytes = (size_t)rows * pitch * sizeof(double);
    const int cb[4] = {q.cs[0] != 0, q.cs[1] != 0, q.cs[2] != 0, q.cs[3] != 0};
    const int cbFac = cb[0] | cb[1] | cb[2];
    const int cbFd = cbFac | cb[3];
    if ((i64)nz * batch > 0x7ffffff0) { why = "too many levels x slices for one tensor map"; return -1; }
    cudaError_t e;
#define X3_ALLOC(ptr, idx, bytes)                                                   \
    if ((e = xm_work_ensure(work, (idx), (bytes))) != cudaSuccess) {                \
        why = std::string("cudaMalloc: ") + cudaGetErrorString(e);                  \
        fused3_plan_release(p);                                                     \
        return -1;                                                                  \
    }                                                                               \
    (ptr) = work.p[idx];
    X3_ALLOC(p.bufS[0], 0, vol_bytes * batch);
    X3_ALLOC(p.bufS[1], 1, vol_bytes * batch);
    X3_ALLOC(p.bufA, 2, vol_bytes * (cb[0] ? batch : 1));
    X3_ALLOC(p.bufC, 3, vol_bytes * (cb[2] ? batch : 1));
    X3_ALLOC(p.bufFd, 4, vol_bytes * (cbFd ? batch : 1));
    X3_ALLOC(p.bufFac, 5, vol_bytes * (cbFac ? batch : 1));
    X3_ALLOC(p.bufB, 6, vol_bytes * (cb[1] ? batch : 1));
    void *flag;
    X3_ALLOC(flag, 7, 16);
#undef X3_ALLOC
    {
        const char *env = getenv("XINV_FUSED3_VARIANT");
        p.variant = env ? atoi(env) : x3_choose_variant(ny, nx, batch, sm_count);
        if (p.variant < 0 || p.variant >= X3_NVARIANTS) p.variant = 0;
    }
    const X3Variant v = X3_VARIANTS[p.variant];
    dim3 blk(128);
    auto gridfor = [&](i64 cols, i64 nb) {
        i64 gy = rows * nb;
        if (gy > 32768) gy = 32768;
        return dim3((unsigned)((cols + 127) / 128), (unsigned)gy, 1);
    };
    auto pack = [&](void *dst, const double *src, i64 bstride, i64 nb) {
        x3_pack_kernel<<<gridfor(pitch, nb), blk, 0, stream>>>((double *)dst, src, rows, nx, pitch, bstride, nb, periodic);
    };
    pack(p.bufS[0], dS, g.N, batch);
    pack(p.bufS[1], dS, g.N, batch);       // levels 0 / nz-1 and all pad columns of both buffers start identical
    pack(p.bufA, q.c[0], q.cs[0], cb[0] ? batch : 1);
    pack(p.bufB, q.c[1], q.cs[1], cb[1] ? batch : 1);
    pack(p.bufC, q.c[2], q.cs[2], cb[2] ? batch : 1);
    x3_pack_derived_kernel<<<gridfor(pitch, cbFd ? batch : 1), blk, 0, stream>>>(
        (double *)p.bufFd, (double *)p.bufFac, q, nz, ny, nx, pitch, cbFd ? batch : 1, cbFac ? batch : 1, periodic);
    if ((e = cudaGetLastError()) != cudaSuccess) {
        why = std::string("pack kernels: ") + cudaGetErrorString(e);
        fused3_plan_release(p);
        return -1;
    }
    // tensor maps: (pitch, ny, levels x volumes), box 64 x TJ (B: TJ+1) x 1
    if (xf_make_map(&p.mS[0], p.bufS[0], pitch, ny, nz * batch, X3_W, v.TJ, why) ||
        xf_make_map(&p.mS[1], p.bufS[1], pitch, ny, nz * batch, X3_W, v.TJ, why) ||
        xf_make_map(&p.mA, p.bufA, pitch, ny, nz * (cb[0] ? batch : 1), X3_W, v.TJ, why) ||
        xf_make_map(&p.mB, p.bufB, pitch, ny, nz * (cb[1] ? batch : 1), X3_W, v.TJ + 1, why) ||
        xf_make_map(&p.mC, p.bufC, pitch, ny, nz * (cb[2] ? batch : 1), X3_W, v.TJ, why) ||
        xf_make_map(&p.mFd, p.bufFd, pitch, ny, nz * (cbFd ? batch : 1), X3_W, v.TJ, why) ||
        xf_make_map(&p.mFac, p.bufFac, pitch, ny, nz * (cbFac ? batch : 1), X3_W, v.TJ, why)) {
        fused3_plan_release(p);
        return -1;
    }
    X3Args &a = p.args;
    a.Sbuf[0] = (double *)p.bufS[0];
    a.Sbuf[1] = (double *)p.bufS[1];
    a.pitch = pitch; a.plane = ny * pitch; a.slice = nz * ny * pitch;
    a.nz = (int)nz; a.ny = (int)ny; a.nx = (int)nx;
    a.RB = v.TJ - 4;
    a.ntx = (int)((nx + X3_W - 5) / (X3_W - 4));
    a.nty = (int)((ny + a.RB - 1) / a.RB);
    a.batch = (int)batch;
    a.bcy = g.bcy; a.bcx = g.bcx;
    a.cbA = cb[0]; a.cbB = cb[1]; a.cbC = cb[2]; a.cbFd = cbFd; a.cbFac = cbFac;
    a.r2 = q.p[1]; a.r1 = q.p[2]; a.undef = q.undef;
    p.batch = batch;
    p.nblk_partials = a.ntx * a.nty;
    const i64 tiles = (i64)a.ntx * a.nty * batch;
    p.grid = (int)(tiles < sm_count ? tiles : sm_count);
    if (p.grid < 1) p.grid = 1;
    int blocks_per_sm = 0;
#define X3_PREP(TJ_, K_) e = x3_prepare<TJ_, K_>(&p.smem, &blocks_per_sm)
    X3_DISPATCH(p.variant, X3_PREP);
#undef X3_PREP
    if (e != cudaSuccess || blocks_per_sm < 1) {
        why = std::string("3-D fused kernel does not fit: ") + cudaGetErrorString(e);
        fused3_plan_release(p);
        return -1;
    }
    {
        int can_coop = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&can_coop, cudaDevAttrCooperativeLaunch, dev);
        p.coop = can_coop && ((i64)blocks_per_sm * sm_count >= p.grid);
        const char *eppl = getenv("XINV_FUSED_PPL");
        p.ppl = p.coop ? (eppl ? atoi(eppl) : 32) : 1;
        if (p.ppl < 1) p.ppl = 1;
        if ((e = cudaMemsetAsync((char *)work.p[7] + 8, 0, 8, stream)) != cudaSuccess) {
            why = std::string("barrier counter: ") + cudaGetErrorString(e);
            fused3_plan_release(p);
            return -1;
        }
        a.gbar = reinterpret_cast<unsigned long long *>((char *)work.p[7] + 8);
        p.gbar_base = 0;
    }
    p.built = true;
    return 0;
}

// one launch = npass passes (one iteration each on every active slice)
static inline int fused3_sweep(Fused3Plan &p, cudaStream_t stream, XdSliceState *st, double *psum, i64 *pcnt,
                               unsigned *ticket, int *nactive, double tol, i64 mxLoop, int npass, int64_t *launches)
{
    X3Args &a = p.args;
    a.st = st; a.psum = psum; a.pcnt = pcnt; a.ticket = ticket; a.nactive = nactive;
    a.tol = tol; a.mxLoop = mxLoop;
    cudaError_t e = cudaSuccess;
#define X3_GO(TJ_, K_) e = x3_launch<TJ_, K_>(p, stream)
    if (npass > 1) {
        a.npass = npass;
        a.gbar_base = p.gbar_base;
        X3_DISPATCH(p.variant, X3_GO);
        if (e == cudaSuccess) {
            p.gbar_base += (unsigned long long)p.grid * (unsigned long long)(npass - 1);
            *launches += 1;
            return 0;
        }
        (void)cudaGetLastError();                // cooperative launch refused: one pass per launch from here on
        p.coop = false;
        p.ppl = 1;
    }
    a.npass = 1;
    a.gbar_base = p.gbar_base;
    for (int n = 0; n < npass; ++n) {
        X3_DISPATCH(p.variant, X3_GO);
        if (e != cudaSuccess) return -1;
        *launches += 1;
    }
#undef X3_GO
    return 0;
}

static inline int fused3_unpack(Fused3Plan &p, double *dS, const XdSliceState *st, cudaStream_t stream)
{
    const X3Args &a = p.args;
    const i64 rows = (i64)a.nz * a.ny;
    i64 gy = rows * p.batch;
    if (gy > 32768) gy = 32768;
    dim3 grid((unsigned)((a.nx + 127) / 128), (unsigned)gy, 1);
    x3_unpack_kernel<<<grid, 128, 0, stream>>>(dS, a.Sbuf[0], a.Sbuf[1], rows, a.nx, a.pitch, p.batch, st);
    return 0;
}
