// xinv_march3d.cuh -- XINV_ENGINE_FUSED for the 3-D standard form (invert_omega;
// numbas.py:15-212): ONE pass over the volume performs a complete red+black SOR
// iteration, the y-"extend" rows, sum|omega| / count and the loop control.
//
// Design ("plane marching"): a CTA owns a tile of TJ rows x 64 columns x a range of levels
// and marches along z.  One warp per tile row, one column pair per lane (as in the 2-D
// engine: x-neighbours by warp shuffle).
//   * Once per solve the engine builds padded copies of the operands (same layout as the
//     2-D engine: XM_PADL ghost columns left, >= XM_GHOST right, holding the periodic
//     wrap-around neighbours) and two derived arrays, with the reference's own operations:
//        Fd  = F * delxSqr, or a marker where the cell must never be updated (boundary
//              level / row / fixed column, an undef operand: numbas.py:117-118, :147-150)
//        fac = optArg / ((A[k+1]+A[k])*ratio2Sqr + (B[j+1]+B[j])*ratio1Sqr + (C[i+1]+C[i]))
//                                                                     (numbas.py:166-168)
//     (forming the factor in the kernel instead was measured: 31 -> 44 us per sweep on the
//     37 x 180 x 360 case -- the march is bound by dependent latency, not by bytes.)
//   * AROW kernels: A constant along x (detected on the device; invert_omega's A = f^2 cos(lat),
//     apps.py:2025-2036) arrives as one value per row instead of a 64-column tile.
//   * Warp 0 is the producer: it feeds a K-stage shared-memory ring with TMA box loads, per level
//     one box of omega (TJ rows x 64 columns), B (TJ-1 rows: the updated rows and the row north of
//     them) and A, C, Fd, fac (the TJ-2 rows on which anything is ever updated); a "landed"
//     mbarrier per stage completes on the TMA byte count, a "free" mbarrier per stage collects one
//     arrival per compute warp, so the producer runs up to K levels ahead and the issue of the
//     loads is off the critical path of the march.  (Rows 0 and TJ-1 of a tile are never updated and
//     their warps would compute nothing anyone reads: warp 0 produces, warp TJ-1 is spare, the TJ-2
//     compute warps meet at a named barrier.)
//   * z-neighbours live in registers (the lane keeps its column pair of the last four
//     levels and A of the last three); y-neighbours of the red half step are read from the
//     staged (still untouched) level, those of the black half step from a small exchange
//     buffer into which every warp publishes its row after the red half step.
//   * Schedule per step (level kl has just arrived): red cells of level kl-1 (colour 0 =
//     (i+j+k) even, as the colour engine and the oracle), black cells of level kl-2, which
//     is then complete: norm accumulation, store to the OTHER omega buffer (ping-pong:
//     neighbouring tiles still need the old values).  One __syncthreads per step.  The march is
//     unrolled four steps deep, so the colour of the lane's even column and every slot of the
//     register windows are compile-time constants: no register is ever moved.
//   * A tile has a 2-cell halo in y and x and -- when the levels are split over several tiles
//     to fill the machine (small volumes) -- 2 levels in z: red results of the halo are
//     recomputed, black results only exist for the owned TJ-4 rows x 60 columns x levels.
//   * Per-tile (sum, count) partials are combined in fixed order by the CTA that finishes
//     a slice last (atomic ticket), which then runs numbas.py:197-210.
// HBM/L2 traffic per pass (= per iteration): omega r+w, A, B, C, Fd, fac = 56 N bytes
// (AROW: 48 N), x the halo overhead of the tiling on the reads; 48 N (40 N) are algorithmic
// (omega r+w, A, B, C, F).
#pragma once
#include "xinv_march2d.cuh"

#define X3_W 64

struct X3Args {
    double *Sbuf[2];          // padded omega buffers [batch][nz][ny][pitch]
    i64 pitch, plane, slice;  // plane = ny * pitch, slice = nz * plane
    int nz, ny, nx;
    int ntx, nty, ntz, RB, ZB;   // column / row / level tiles; owned rows (TJ - 4) and levels (even) per tile
    int batch;
    int bcy, bcx;
    int cbA, cbB, cbC, cbFd, cbFac;   // 1: the array has a batch axis, 0: one volume shared by the batch
    double r2, r1, undef;     // ratio2Sqr, ratio1Sqr
    XdSliceState *st;
    double *psum;             // [batch][ntx*nty*ntz]
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

// Shared-memory layout of one ring stage (offsets in doubles).  Row r of the omega box is global row
// y0 - 2 + r; the A, C, Fd, fac and B boxes start one row later (row r <-> y0 - 1 + r).
// CM = coefficient mode: X3_DENSE (A, B, C, fac as volumes), X3_AROW (A as one value per row), X3_ROWS (A, B, C and
// fac all as one value per row: the stage holds omega, four vectors of TJ row values and Fd, nothing else).
#define X3_DENSE 0
#define X3_AROW 1
#define X3_ROWS 2
template <int TJ, int CM>
struct X3Lay {
    static constexpr int W = X3_W;
    static constexpr int RC = TJ - 2;            // rows of A, C, Fd, fac: everything that is ever updated
    static constexpr int RBN = TJ - 1;           // rows of B: those and the row north of them
    static constexpr int OFF_S = 0;
    static constexpr int OFF_A = TJ * W;
    // row values: TJ per vector (rows y0-2 ...: TMA start coordinates must be even), padded to 128 bytes
    static constexpr int A_SZ = (CM == X3_ROWS) ? (4 * TJ + 15) / 16 * 16 : (CM == X3_AROW) ? 32 : RC * W;
    static constexpr int C_SZ = (CM == X3_ROWS) ? 0 : RC * W;
    static constexpr int OFF_C = OFF_A + A_SZ;
    static constexpr int OFF_FD = OFF_C + C_SZ;
    static constexpr int OFF_FAC = OFF_FD + RC * W;
    static constexpr int OFF_B = OFF_FAC + C_SZ;
    static constexpr int STAGE = OFF_B + ((CM == X3_ROWS) ? 0 : RBN * W);
    static constexpr uint32_t TX_BYTES =
        (uint32_t)(((CM == X3_ROWS) ? TJ * W + 4 * TJ + RC * W
                                    : TJ * W + ((CM == X3_AROW) ? TJ : RC * W) + 3 * RC * W + RBN * W) * sizeof(double));
    static_assert(TJ <= 32 && (TJ % 2) == 0, "row-value box: at most 256 bytes, a multiple of 16");
};

// Everything the march of one tile needs besides its registers.
struct X3Tile {
    const double *ringrow;   // ring + (this warp's row) * W + 2 * lane
    const double *arow;      // AROW: ring + OFF_A + this warp's row
    double *xrow;            // exchange buffers, same offset as ringrow
    uint64_t *bars;          // [K] "level has landed" (TMA transaction barriers) | [K] "stage is free again" (compute warps arrive)
    double *outp;            // output row of level kbase - 2 (advanced by one level per step)
    i64 plane;
    int nsteps;              // steps of the march; step s brings level kbase + s (kbase even)
    int s_load;              // steps [0, s_load) bring a level
    int s_red0, s_red1;      // steps in which red cells (of the level that arrived one step earlier) are updated
    int s_blk0, s_blk1;      // ... black cells (two steps earlier): the owned, updatable levels
    int s_fin0, s_fin1;      // steps in which an owned level is complete (norm)
    int s_ext0, s_ext1;      // arrival steps of the levels 1 .. nz-2 (y-extend)
    int nx, gx;
    int wn, ws;              // smem offsets of the rows north / south of this warp's row
    bool periodic, ext_j0, ext_jN, ext_j1, ext_jM;
    bool own_x, own_y, ghe, ghw, edge;
    double undef, r2, r1;
};

// ring position: carried from tile to tile (the mbarrier phases go on)
struct X3Ring {
    int cons;                // stage the next arriving / issued level sits in
    unsigned phase;          // bit k: parity to wait for on stage k's barrier (compute warps: "landed"; producer: "free")
};

__device__ __forceinline__ void x3_mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(xf_smem_u32(bar)) : "memory");
}
// the compute warps of a CTA meet at named barrier 1 (the producer warp and the spare warp do not take part)
// Warps whose row is odd and warps whose row is even run different instantiations of x3_march, so they reach the
// same barrier from different instruction addresses.  PTX allows that (bar.sync is "aligned" within a warp only), but
// compute-sanitizer's synccheck reports it as "divergent thread(s) in block".  -DX3_BARRIER_NOINLINE=1 puts the barrier
// behind one function address for the tool (0 reports; 29.5 -> 30.6 us per sweep at 360x180x37, so not the default).
#ifndef X3_BARRIER_NOINLINE
#define X3_BARRIER_NOINLINE 0
#endif
#if X3_BARRIER_NOINLINE
__device__ __noinline__ void x3_bar_compute(int nthreads)
#else
__device__ __forceinline__ void x3_bar_compute(int nthreads)
#endif
{
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

// One tile.  JODD = parity of this warp's global row: with the step index known modulo 4 (4-step unrolled
// loop, kbase even) the colour of the lane's even column is a compile-time constant and the register windows
// (omega of the last four levels, A of the last three, the operands saved for the black half step) are
// indexed by constants.
template <int TJ, int K, int CM, bool JODD>
__device__ __forceinline__ void x3_march(const X3Tile &t, X3Ring &rg, double &nsum, int &ncnt)
{
    using L = X3Lay<TJ, CM>;
    constexpr bool AROW = (CM != X3_DENSE), ROWS = (CM == X3_ROWS);
    constexpr int W = X3_W;
    constexpr int TILE = TJ * W;
    // the coefficient boxes start one row below the omega box: fold the -W into the offsets
    constexpr int OFF_S = L::OFF_S, OFF_A = L::OFF_A - W, OFF_C = L::OFF_C - W, OFF_FD = L::OFF_FD - W, OFF_FAC = L::OFF_FAC - W;
    constexpr int OFF_BC = L::OFF_B - W, OFF_BN = L::OFF_B;
    constexpr int STAGE = L::STAGE;
    const double2 zero2 = make_double2(0.0, 0.0);
    const double undef = t.undef, r2 = t.r2, r1 = t.r1;
    double2 P[4] = {zero2, zero2, zero2, zero2};     // omega: the level of step s in P[s & 3]
    double2 Aw[4] = {zero2, zero2, zero2, zero2};    // A:     likewise
    // operands of the black cell of the level that arrived one step earlier, saved at step s in slot s & 1
    double kBc[2] = {0.0, 0.0}, kBn[2] = {0.0, 0.0}, kCw[2] = {0.0, 0.0}, kCe[2] = {0.0, 0.0}, kFac[2] = {0.0, 0.0};
    double kFd[2] = {xm_skip_value(), xm_skip_value()};
    double *outp = t.outp;
    int prev_off = 0;                                // stage offset (doubles) of the level that arrived in the previous step
    int prev_stage = -1;                             // ... and its stage (-1: none to hand back)
    const int lane = threadIdx.x & 31;

    // ST ("steady"): a step in which a level arrives, red, black and a finished level all take place and the level
    // is one of 1 .. nz-2: every range test below is true at compile time (most steps of a march)
    auto step = [&](auto u_tag, auto st_tag, const int s) {
        constexpr int U = decltype(u_tag)::value;    // s & 3
        constexpr bool ST = decltype(st_tag)::value;
        constexpr int U1 = (U + 3) & 3, U2 = (U + 2) & 3, U3 = (U + 1) & 3;   // slots of the levels of steps s-1, s-2, s-3
        // red cell of the level of step s-1 = the even column of the pair  <=>  j + kbase + s - 1 even
        constexpr bool RX = ((int(JODD) + U + 1) & 1) == 0;
        int cur_off = prev_off, cur_stage = -1;
        // ---- a level arrives ----
        if (ST || (s < t.s_load)) {
            xf_mbar_wait(&t.bars[rg.cons], (rg.phase >> rg.cons) & 1u);
            rg.phase ^= (1u << rg.cons);
            cur_off = rg.cons * STAGE;
            cur_stage = rg.cons;
            rg.cons = (rg.cons + 1 == K) ? 0 : rg.cons + 1;
            const double *g = t.ringrow + cur_off;
            P[U] = x3_ld2(g + OFF_S);
            if (AROW) { const double av = t.arow[cur_off]; Aw[U] = make_double2(av, av); }
            else      Aw[U] = x3_ld2(g + OFF_A);
            if ((t.ext_j0 | t.ext_jN) && (ST || ((s >= t.s_ext0) & (s <= t.s_ext1)))) {   // numbas.py:87-115: levels 1..nz-2 only
                if (t.ext_j0) P[U] = xm_extend(P[U], x3_ld2(g + OFF_S + W), t.gx, t.nx, t.periodic, undef);
                else          P[U] = xm_extend(P[U], x3_ld2(g + OFF_S - W), t.gx, t.nx, t.periodic, undef);
            }
        } else {
            P[U] = zero2; Aw[U] = zero2;
        }
        // ---- red cells of the level of step s-1 (neighbours in y: the staged level, still untouched) ----
        if (ST || ((s >= t.s_red0) & (s <= t.s_red1))) {
            const double *r = t.ringrow + prev_off;
            double2 Bc, Bn, Cc, Fc;
            if (ROWS) {                                            // one value per row: B of this row and of the row north, C, fac
                const double *rv = t.arow + prev_off;
                const double vb = rv[TJ], vn = rv[TJ + 1], vc = rv[2 * TJ], vf = rv[3 * TJ];
                Bc = make_double2(vb, vb); Bn = make_double2(vn, vn); Cc = make_double2(vc, vc); Fc = make_double2(vf, vf);
            } else {
                Bc = x3_ld2(r + OFF_BC); Bn = x3_ld2(r + OFF_BN); Cc = x3_ld2(r + OFF_C); Fc = x3_ld2(r + OFF_FAC);
            }
            const double2 Fd = x3_ld2(r + OFF_FD);
            double2 Sn = x3_ld2(r + OFF_S + t.wn), Ss = x3_ld2(r + OFF_S + t.ws);
            if (t.ext_j1 | t.ext_jM) {
                // the extended boundary row as the cells of rows 1 / ny-2 see it: their own old value
                if (t.ext_j1) { if (P[U1].x != undef) Ss.x = P[U1].x; if (P[U1].y != undef) Ss.y = P[U1].y; }
                if (t.ext_jM) { if (P[U1].x != undef) Sn.x = P[U1].x; if (P[U1].y != undef) Sn.y = P[U1].y; }
            }
            const double Cnext = ROWS ? Cc.x : xm_shfl_down1(Cc.x);    // C of the column east of the pair
            if (RX) {
                const double nb = xm_shfl_up1(P[U1].y);
                P[U1].x = x3_cell(P[U1].x, P[U].x, P[U2].x, Sn.x, Ss.x, P[U1].y, nb, Aw[U].x, Aw[U1].x, Bn.x, Bc.x, Cc.y, Cc.x,
                                  Fd.x, Fc.x, r2, r1);
                kBc[U & 1] = Bc.y; kBn[U & 1] = Bn.y; kCw[U & 1] = Cc.y; kCe[U & 1] = Cnext; kFd[U & 1] = Fd.y; kFac[U & 1] = Fc.y;
            } else {
                const double nb = xm_shfl_down1(P[U1].x);
                P[U1].y = x3_cell(P[U1].y, P[U].y, P[U2].y, Sn.y, Ss.y, nb, P[U1].x, Aw[U].y, Aw[U1].y, Bn.y, Bc.y, Cnext, Cc.y,
                                  Fd.y, Fc.y, r2, r1);
                kBc[U & 1] = Bc.x; kBn[U & 1] = Bn.x; kCw[U & 1] = Cc.x; kCe[U & 1] = Cc.y; kFd[U & 1] = Fd.x; kFac[U & 1] = Fc.x;
            }
        } else {
            kFd[U & 1] = xm_skip_value();
        }
        // publish the row (red cells final for this iteration) for the black half step of the next step
        *reinterpret_cast<double2 *>(t.xrow + (U & 1) * TILE) = P[U1];
        // ---- black cells of the level of step s-2 (neighbours in y: rows published in the previous step) ----
        const bool blk = ST || ((s >= t.s_blk0) & (s <= t.s_blk1));
        if (blk) {
            constexpr int Q = (U + 1) & 1;                         // slot written in the previous step
            const double *xr = t.xrow + Q * TILE;
            const double2 Sn = x3_ld2(xr + t.wn), Ss = x3_ld2(xr + t.ws);
            if (RX) {                                              // red of step s-1's level even column <=> black of step s-2's
                const double nb = xm_shfl_up1(P[U2].y);
                P[U2].x = x3_cell(P[U2].x, P[U1].x, P[U3].x, Sn.x, Ss.x, P[U2].y, nb, Aw[U1].x, Aw[U2].x, kBn[Q], kBc[Q], kCe[Q],
                                  kCw[Q], kFd[Q], kFac[Q], r2, r1);
            } else {
                const double nb = xm_shfl_down1(P[U2].x);
                P[U2].y = x3_cell(P[U2].y, P[U1].y, P[U3].y, Sn.y, Ss.y, nb, P[U2].x, Aw[U1].y, Aw[U2].y, kBn[Q], kBc[Q], kCe[Q],
                                  kCw[Q], kFd[Q], kFac[Q], r2, r1);
            }
        }
        // ---- an owned level is complete: norm over owned cells (numbas.py:1689-1708), store ----
        if (ST || ((s >= t.s_fin0) & (s <= t.s_fin1))) {
            xm_norm_acc_lane(nsum, ncnt, P[U2].x, t.own_x, undef);
            xm_norm_acc_lane(nsum, ncnt, P[U2].y, t.own_y, undef);
            if (blk) {                                             // levels 0 and nz-1 never change
                xm_store2_if(t.own_y, outp, P[U2]);
                if (t.edge) {                                      // odd nx / periodic ghost columns (warp-uniform)
                    xm_store1_if(t.own_x & !t.own_y, outp, P[U2].x);
                    xm_store2_if(t.ghe, outp + t.nx, P[U2]);
                    xm_store2_if(t.ghw, outp - t.nx, P[U2]);
                }
            }
        }
        outp += t.plane;
        // the stage of the previous step's level has been read for the last time: hand it back to the producer
        if (ST || prev_stage >= 0) {
            __syncwarp();
            if (lane == 0) x3_mbar_arrive(&t.bars[K + prev_stage]);
        }
        prev_off = cur_off;
        prev_stage = cur_stage;
        x3_bar_compute((TJ - 2) * 32);           // rows published
    };

    // steady steps [lo, hi]: groups of four steps inside that range run without the range tests
    const int lo = max(max(max(1, t.s_red0), max(t.s_blk0, t.s_fin0)), t.s_ext0);
    const int hi = min(min(min(t.s_load - 1, t.s_red1), min(t.s_blk1, t.s_fin1)), t.s_ext1);
    for (int s0 = 0; s0 < t.nsteps; s0 += 4) {
        if ((s0 >= lo) & (s0 + 3 <= hi)) {
            step(std::integral_constant<int, 0>{}, std::true_type{}, s0);
            step(std::integral_constant<int, 1>{}, std::true_type{}, s0 + 1);
            step(std::integral_constant<int, 2>{}, std::true_type{}, s0 + 2);
            step(std::integral_constant<int, 3>{}, std::true_type{}, s0 + 3);
        } else {
            step(std::integral_constant<int, 0>{}, std::false_type{}, s0);
            if (s0 + 1 < t.nsteps) step(std::integral_constant<int, 1>{}, std::false_type{}, s0 + 1);
            if (s0 + 2 < t.nsteps) step(std::integral_constant<int, 2>{}, std::false_type{}, s0 + 2);
            if (s0 + 3 < t.nsteps) step(std::integral_constant<int, 3>{}, std::false_type{}, s0 + 3);
        }
    }
    if (prev_stage >= 0) {                           // a level that arrived in the very last step
        __syncwarp();
        if (lane == 0) x3_mbar_arrive(&t.bars[K + prev_stage]);
    }
}

template <int TJ, int K, int MINB, int CM>
__global__ void __launch_bounds__(TJ * 32, MINB)
xm3_std3d_kernel(const __grid_constant__ CUtensorMap mS0, const __grid_constant__ CUtensorMap mS1,
                 const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mB,
                 const __grid_constant__ CUtensorMap mC, const __grid_constant__ CUtensorMap mFd,
                 const __grid_constant__ CUtensorMap mFac, const X3Args a)
{
    using L = X3Lay<TJ, CM>;
    constexpr bool AROW = (CM != X3_DENSE), ROWS = (CM == X3_ROWS);
    constexpr int W = X3_W;
    constexpr int TILE = TJ * W;                 // doubles per exchange buffer
    constexpr int STAGE = L::STAGE;

    extern __shared__ __align__(1024) unsigned char x3_smem[];
    double *ring = reinterpret_cast<double *>(x3_smem);
    double *X = ring + (size_t)K * STAGE;        // two exchange buffers of TILE doubles
    double *red_sum = X + 2 * TILE;              // [TJ]
    i64 *red_cnt = reinterpret_cast<i64 *>(red_sum + TJ);   // [TJ]
    uint64_t *bars = reinterpret_cast<uint64_t *>(red_cnt + TJ);   // [2 K]: landed | free
    int *box = reinterpret_cast<int *>(bars + 2 * K);        // [4] CTA-wide broadcasts
    XdSliceState *lst = reinterpret_cast<XdSliceState *>(box + 4);   // replicated loop control: this CTA's copy of its slice's state

    const int lane = threadIdx.x & 31;
    const int w = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);   // tile row of this warp (warp-uniform)
    // Warp roles: the rows 0 and TJ-1 of a tile are never updated (their red results would be wrong and nothing
    // reads them), so those two warps do not march: warp 0 is the TMA producer, warp TJ-1 is spare.
    if (threadIdx.x == 0) {
        #pragma unroll
        for (int s = 0; s < K; ++s) { xf_mbar_init(&bars[s], 1); xf_mbar_init(&bars[K + s], TJ - 2); }
        xf_fence_barrier_init();
    }
    __syncthreads();

    const int nx = a.nx, ny = a.ny, nz = a.nz;
    const bool periodic = (a.bcx == XD_BC_PERIODIC);
    const bool extend = (a.bcy == XD_BC_EXTEND);
    const int tpl = a.ntx * a.nty;               // tiles per level range
    const int tps = tpl * a.ntz;                 // tiles per slice
    const int total = tps * a.batch;
    // Replicated loop control (every CTA has at most one tile per pass, several passes per launch; cf. xinv_march2d.cuh):
    // no CTA sums the partials and runs numbas.py:197-210 BEFORE the grid barrier; behind it warp 0 of every CTA does
    // so for its own slice, from the same partials in the same order, on a copy of the slice's state kept in shared
    // memory for the whole launch; the CTA with the slice's first tile writes the state back for the host.
    const bool repl = (total <= (int)gridDim.x) && a.npass > 1;
    X3Ring rg;
    rg.cons = 0;
    rg.phase = (w == 0) ? 0xffffffffu : 0u;      // producer: a fresh "free" barrier counts as completed (parity 1)

    for (int pp = 0; pp < a.npass; ++pp) {
    for (int tile = blockIdx.x; tile < total; tile += gridDim.x) {
        const int b = tile / tps;
        const int tidx = tile - b * tps;
        const int zb = tidx / tpl, trem = tidx - zb * tpl;
        const int yb = trem / a.ntx, xb = trem - yb * a.ntx;
        // slice state: constant while any tile of the slice is still to do in this pass
        // (read through L2: another SM rewrites it between two passes of one launch)
        if (threadIdx.x == 0) {
            if (repl) {
                if (pp == 0) *lst = a.st[b];     // (global state: first pass of the launch only)
                box[0] = lst->active; box[1] = lst->cur;
            } else {
                box[0] = __ldcg(&a.st[b].active); box[1] = __ldcg(&a.st[b].cur);
            }
        }
        __syncthreads();
        const int active = box[0], cur = box[1];
        __syncthreads();                         // box is rewritten by the next tile
        if (!active) continue;                   // frozen slice

        const int x0 = xb * (W - 4), y0 = yb * a.RB;     // RB is even: the parity of a warp's row never changes
        const int k0 = zb * a.ZB, k1 = min(k0 + a.ZB, nz);   // owned levels [k0, k1); ZB is even
        const int kbase = (k0 >= 2) ? k0 - 2 : 0;            // first level loaded (even)
        const int klast = min(k1 + 1, nz - 1);               // last level loaded
        const int j = y0 - 2 + w;                // global row of this warp
        const int gx = x0 - 2 + 2 * lane;        // global (even) column of this lane's pair
        const int bx = x0 - 2 + XM_PADL;         // padded x coordinate of the tile's first column
        const CUtensorMap *mS = cur ? &mS1 : &mS0;
        const bool own_row = (w >= 2) & (w < TJ - 2) & (j < ny);
        const bool own_lane = (lane >= 1) & (lane < 31);
        X3Tile t;
        t.own_x = own_row & own_lane & (gx < nx);
        t.own_y = own_row & own_lane & (gx + 1 < nx);
        t.ghe = t.own_y & periodic & (gx < XM_GHOST);          // also write the east ghost copy
        t.ghw = t.own_y & periodic & (gx >= nx - XM_GHOST);    // also write the west ghost copy
        // warp-uniform: this tile has single-column stores (odd nx) or ghost-column duty
        t.edge = ((nx & 1) & (x0 + W - 2 > nx)) | (periodic & ((x0 < XM_GHOST + 2) | (x0 + W - 2 > nx - XM_GHOST)));
        t.wn = (w + 1 < TJ) ? W : 0;             // rows beyond the tile only ever feed halo rows
        t.ws = (w > 0) ? -W : 0;
        t.ringrow = ring + w * W + 2 * lane;
        t.arow = ring + L::OFF_A + w;
        t.xrow = X + w * W + 2 * lane;
        t.bars = bars;
        t.plane = a.plane;
        // step s brings level kbase + s; red works on level kbase + s - 1, black on kbase + s - 2
        t.s_load = klast - kbase + 1;
        t.s_red0 = max(1, k0 - 1) - kbase + 1;   t.s_red1 = min(nz - 2, k1) - kbase + 1;        // halo levels k0-1 and k1 too
        t.s_blk0 = max(1, k0) - kbase + 2;       t.s_blk1 = min(nz - 2, k1 - 1) - kbase + 2;    // owned levels only
        t.s_fin0 = k0 - kbase + 2;               t.s_fin1 = k1 - 1 - kbase + 2;
        t.s_ext0 = 1 - kbase;                    t.s_ext1 = nz - 2 - kbase;
        t.nsteps = t.s_fin1 + 1;
        t.outp = a.Sbuf[cur ^ 1] + (i64)b * a.slice + (i64)(kbase - 2) * a.plane + (i64)j * a.pitch + XM_PADL + gx;
        t.nx = nx; t.gx = gx;
        t.periodic = periodic;
        t.ext_j0 = extend & (j == 0); t.ext_jN = extend & (j == ny - 1);
        t.ext_j1 = extend & (j == 1); t.ext_jM = extend & (j == ny - 2);
        t.undef = a.undef; t.r2 = a.r2; t.r1 = a.r1;

        auto issue = [&](int s, int stage) {     // producer lane: TMA loads of the level of step s into `stage`
            const int k = kbase + s;
            double *dst = ring + (size_t)stage * STAGE;
            uint64_t *bar = &bars[stage];
            const int ys = y0 - 2;
            xf_mbar_expect_tx(bar, L::TX_BYTES);
            xf_tma_load_3d(dst + L::OFF_S, mS, bar, bx, ys, b * nz + k);                        // TJ rows
            if (ROWS) {                          // A, B, C, fac: four vectors of TJ row values in one box
                xf_tma_load_3d(dst + L::OFF_A, &mA, bar, ys, 0, b * a.cbA * nz + k);
                xf_tma_load_3d(dst + L::OFF_FD, &mFd, bar, bx, ys + 1, b * a.cbFd * nz + k);
                return;
            }
            if (AROW) xf_tma_load_3d(dst + L::OFF_A, &mA, bar, ys, 0, b * a.cbA * nz + k);      // TJ row values (even start)
            else      xf_tma_load_3d(dst + L::OFF_A, &mA, bar, bx, ys + 1, b * a.cbA * nz + k); // TJ-2 rows
            xf_tma_load_3d(dst + L::OFF_C, &mC, bar, bx, ys + 1, b * a.cbC * nz + k);
            xf_tma_load_3d(dst + L::OFF_FD, &mFd, bar, bx, ys + 1, b * a.cbFd * nz + k);
            xf_tma_load_3d(dst + L::OFF_FAC, &mFac, bar, bx, ys + 1, b * a.cbFac * nz + k);
            xf_tma_load_3d(dst + L::OFF_B, &mB, bar, bx, ys + 1, b * a.cbB * nz + k);           // TJ-1 rows
        };
        double nsum = 0.0;
        int ncnt = 0;
        if (w == 0) {
            // producer: one level per stage, as far ahead as the ring allows
            for (int s = 0; s < t.s_load; ++s) {
                xf_mbar_wait(&bars[K + rg.cons], (rg.phase >> rg.cons) & 1u);    // the compute warps are done with the stage
                rg.phase ^= (1u << rg.cons);
                if (lane == 0) {
                    xf_fence_proxy_async();      // their generic-proxy reads before the async-proxy writes
                    issue(s, rg.cons);
                }
                rg.cons = (rg.cons + 1 == K) ? 0 : rg.cons + 1;
                __syncwarp();
            }
        } else if (w < TJ - 1) {
            if (w & 1) x3_march<TJ, K, CM, true>(t, rg, nsum, ncnt);
            else       x3_march<TJ, K, CM, false>(t, rg, nsum, ncnt);
        }

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
            // (replicated loop control: two sets of slots by the parity of the pass, published by the grid barrier)
            const i64 slot = (repl ? (i64)(pp & 1) * a.batch * tps : 0) + (i64)b * tps + tidx;
            a.psum[slot] = ts;
            a.pcnt[slot] = tc;
            box[2] = 0;
            if (!repl) {
                __threadfence();
                const unsigned tk = atomicAdd(&a.ticket[b], 1u);
                box[2] = (tk == (unsigned)tps - 1u);
            }
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
    if (pp + 1 < a.npass || repl) {              // (replicated loop control: also behind the last pass of the launch)
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
            if (!repl || a.batch > 1) {          // (replicated: read before this boundary's verdicts lower it -- a stale
                int na;                          // count only delays the exit by a pass; a single slice needs no count)
                asm volatile("ld.relaxed.gpu.global.s32 %0, [%1];" : "=r"(na) : "l"(a.nactive) : "memory");
                go_on = (na != 0);
                if (!repl && !go_on && pp + 2 < a.npass) atomicAdd(a.gbar, (unsigned long long)(a.npass - 2 - pp));
            }
        }
        if (repl) {
            __syncthreads();                     // the barrier has been passed: known to every warp
            const int tile0 = blockIdx.x;
            if (w == 0 && tile0 < total && __shfl_sync(0xffffffffu, lst->active, 0)) {   // the CTA's slice ran in this pass
                const int b = tile0 / tps;
                const i64 poff = (i64)(pp & 1) * a.batch * tps + (i64)b * tps;
                // fixed assignment of partials to lanes and a fixed shuffle tree, loads of eight steps in flight
                double s_ = 0.0;
                i64 c_ = 0;
                for (int p0 = lane; p0 < tps; p0 += 32 * 8) {
                    double vs[8];
                    i64 vc[8];
                    #pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        const int p = p0 + 32 * m;
                        vs[m] = (p < tps) ? __ldcg(a.psum + poff + p) : 0.0;
                        vc[m] = (p < tps) ? __ldcg(a.pcnt + poff + p) : 0;
                    }
                    #pragma unroll
                    for (int m = 0; m < 8; ++m) {
                        if (p0 + 32 * m < tps) { s_ += vs[m]; c_ += vc[m]; }
                    }
                }
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) {
                    s_ += __shfl_down_sync(0xffffffffu, s_, o);
                    c_ += __shfl_down_sync(0xffffffffu, c_, o);
                }
                if (lane == 0) {
                    XdSliceState st_ = *lst;
                    xd_decide(st_, s_, c_, a.tol, a.mxLoop, 0);       // no norm == 0 exit in 3-D (numbas.py:206)
                    st_.cur ^= 1;
                    *lst = st_;
                    if (tile0 == b * tps) {      // the CTA with the slice's first tile keeps the global copy current
                        a.st[b] = st_;
                        if (!st_.active) atomicSub(a.nactive, 1);
                    }
                }
            }
            __syncthreads();                     // the verdict is known to every warp of the CTA
            // one slice: every CTA knows the verdict; several: all CTAs leave at the same boundary (monotonic arrival
            // counter), so they go by the count of active slices
            if (a.batch == 1) go_on = (tile0 < total) && lst->active;
            go_on = __syncthreads_or((a.batch == 1) ? (go_on && threadIdx.x == 0) : (go_on && threadIdx.x == 0));
            if (!go_on && threadIdx.x == 0 && pp + 1 < a.npass) atomicAdd(a.gbar, (unsigned long long)(a.npass - 1 - pp));
        } else {
            go_on = __syncthreads_or(go_on && threadIdx.x == 0);
        }
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
        if (fac && b < nbFac) fac[row * pitch + pc] = vf;
    }
}

// AROW detection: flag[0] |= 1 if some X[row][i] differs (bitwise) from X[row][0]
__global__ void x3_rowconst_kernel(const double *__restrict__ X, i64 rows, i64 nx, int *flag)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const long long *P = reinterpret_cast<const long long *>(X);
    int bad = 0;
    for (i64 row = blockIdx.y; row < rows; row += gridDim.y) bad |= (P[row * nx + i] != P[row * nx]);
    if (bad) flag[0] = 1;
}
// AROW operand: vals[v][j] = A[v][j][0] for every (volume x level) v, row pitch rpitch
__global__ void x3_pack_rowvals_kernel(double *__restrict__ vals, const double *__restrict__ A, i64 nv, i64 ny, i64 nx,
                                       i64 rpitch)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rpitch) return;
    for (i64 v = blockIdx.y; v < nv; v += gridDim.y) vals[v * rpitch + j] = (j < ny) ? A[(v * ny + j) * nx] : 0.0;
}

// X3_ROWS operand: vals[v][0..3][j] = A, B, C and the factor of row j of (volume x level) v, from column 0 of the
// dense arrays (which x3_rowconst_kernel has found constant along x); the factor with the operations of numbas.py:166-168
__global__ void x3_pack_rows4_kernel(double *__restrict__ vals, XdCoef q, i64 nz, i64 ny, i64 nx, i64 rpitch, i64 nb)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rpitch) return;
    const double ratio2Sqr = q.p[1], ratio1Sqr = q.p[2];
    for (i64 v = blockIdx.y; v < nz * nb; v += gridDim.y) {
        const i64 b = v / nz, k = v - b * nz;
        double va = 0.0, vb = 0.0, vc = 0.0, vf = 0.0;
        if (j < ny) {
            const double *A = q.c[0] + b * q.cs[0], *B = q.c[1] + b * q.cs[1], *C = q.c[2] + b * q.cs[2];
            const i64 o = (k * ny + j) * nx;
            va = A[o]; vb = B[o]; vc = C[o];
            if ((k >= 1) && (k <= nz - 2) && (j >= 1) && (j <= ny - 2))
                vf = q.optArg / ((A[o + ny * nx] + va) * ratio2Sqr + (B[o + nx] + vb) * ratio1Sqr + (vc + vc));
        }
        double *d = vals + v * 4 * rpitch + j;
        d[0] = va; d[rpitch] = vb; d[2 * rpitch] = vc; d[3 * rpitch] = vf;
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
// Front end for invert_omega (xinv_std3d_rows): what apps.__mask_FS, __coeffs_omega and the
// de-masking of __template do on the host (apps.py:2112-2159, :2016-2052, :1386-1392), on the device.
//   A[k][j][i] = rows[0][j]                       (f^2 cosG | f^2)
//   B[k][j][i] = N2(b,k,j,i) * rows[1][j]          (n2 * cosH | n2)            one IEEE multiply, as numpy does
//   C[k][j][i] = N2(b,k,j,i) / rows[2][j]          (n2 / cosG | n2)            one IEEE division
//   F          = forcing * rows[3][j] on valid cells (cosG | 1), undef on land
// N2 is read through four element strides (batch, level, row, column; 0 = broadcast), so a scalar, a profile
// along any core dimension, a volume shared by the batch and a full array are all the same code.
// ----------------------------------------------------------------------------
struct X3Front {
    bool on = false;
    const double *rows = nullptr;        // [4][ny] (device)
    const double *N2 = nullptr;          // (device)
    i64 ns[4] = {0, 0, 0, 0};
    const double *F = nullptr;           // user forcing [batch][nz][ny][nx] (device)
    double user_undef = 0.0, out_undef = 0.0;
};

__device__ __forceinline__ bool x3_land(double f, double user_undef, double undef)
{
    // (a raw value equal to the internal marker is land too: maskF != -9.99e8, apps.py:2036, :1389)
    return ((user_undef != user_undef) ? (f != f) : (f == user_undef)) | (f == undef);
}

// B and C of the padded layout for nbBC volumes
__global__ void x3_front_bc_kernel(double *__restrict__ Bp, double *__restrict__ Cp, const double *__restrict__ rows,
                                   const double *__restrict__ N2, i64 s0, i64 s1, i64 s2, i64 s3, i64 nz, i64 ny, i64 nx,
                                   i64 pitch, i64 nbBC, int periodic)
{
    const i64 pc = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (pc >= pitch) return;
    const i64 i = pc - XM_PADL;
    const i64 nrows = nz * ny;
    bool col = (i >= 0) && (i < nx);
    i64 iw = i;
    if (periodic && i >= -XM_GHOST && i < nx + XM_GHOST) { col = true; iw = ((i % nx) + nx) % nx; }
    for (i64 row = blockIdx.y; row < nrows * nbBC; row += gridDim.y) {
        const i64 b = row / nrows, r = row - b * nrows;
        const i64 k = r / ny, j = r - k * ny;
        double vb = 0.0, vc = 0.0;
        if (col) {
            const double n2 = N2[b * s0 + k * s1 + j * s2 + iw * s3];
            vb = n2 * rows[ny + j];
            vc = n2 / rows[2 * ny + j];
        }
        Bp[row * pitch + pc] = vb;
        Cp[row * pitch + pc] = vc;
    }
}

// Fd (every volume) and fac (nbFac volumes) straight from the user's forcing, N2 and the row vectors.
// flag[1] |= 1 if a valid forcing value is not finite (the caller then falls back to the host path).
__global__ void x3_front_derived_kernel(double *__restrict__ Fd, double *__restrict__ fac, const double *__restrict__ rows,
                                        const double *__restrict__ N2, i64 s0, i64 s1, i64 s2, i64 s3,
                                        const double *__restrict__ F, i64 nz, i64 ny, i64 nx, i64 pitch, i64 nb, i64 nbFac,
                                        int periodic, double user_undef, double undef, double delxSqr, double ratio2Sqr,
                                        double ratio1Sqr, double optArg, int *flag)
{
    const i64 pc = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (pc >= pitch) return;
    const i64 i = pc - XM_PADL;
    const i64 nrows = nz * ny;
    bool col = true;
    i64 iw = i, ie = i + 1;
    if (periodic) {
        col = (i >= -XM_GHOST) && (i < nx + XM_GHOST);
        iw = ((i % nx) + nx) % nx;
        ie = (iw + 1 == nx) ? 0 : iw + 1;
    } else {
        col = (i >= 1) && (i <= nx - 2);
    }
    for (i64 row = blockIdx.y; row < nrows * nb; row += gridDim.y) {
        const i64 b = row / nrows, r = row - b * nrows;
        const i64 k = r / ny, j = r - k * ny;
        const bool cell = col && (k >= 1) && (k <= nz - 2) && (j >= 1) && (j <= ny - 2);
        double vF = __hiloint2double(XM_SKIP_HI, 0), vf = 0.0;
        if (cell) {
            const double *n2 = N2 + b * s0 + k * s1;
            const double Au = rows[j], Ac = rows[j];
            const double Bn = n2[(j + 1) * s2 + iw * s3] * rows[ny + j + 1], Bc = n2[j * s2 + iw * s3] * rows[ny + j];
            const double Ce = n2[j * s2 + ie * s3] / rows[2 * ny + j], Cc = n2[j * s2 + iw * s3] / rows[2 * ny + j];
            const double f = F[row * nx + iw];
            if (!x3_land(f, user_undef, undef)) {
                if (!isfinite(f)) flag[1] = 1;
                const double fm = f * rows[3 * ny + j];
                if ((fm != undef) & (Au != undef) & (Bn != undef) & (Bc != undef) & (Ce != undef) & (Cc != undef))
                    vF = fm * delxSqr;
            }
            vf = optArg / ((Au + Ac) * ratio2Sqr + (Bn + Bc) * ratio1Sqr + (Ce + Cc));
        }
        Fd[row * pitch + pc] = vF;
        if (fac && b < nbFac) fac[row * pitch + pc] = vf;
    }
}

// X3_ROWS operand of the front end (N2 without a column axis): the same values x3_front_bc_kernel and
// x3_front_derived_kernel form cell by cell, once per row
__global__ void x3_front_rows4_kernel(double *__restrict__ vals, const double *__restrict__ rows, const double *__restrict__ N2,
                                      i64 s0, i64 s1, i64 s2, i64 nz, i64 ny, i64 rpitch, i64 nb, double ratio2Sqr,
                                      double ratio1Sqr, double optArg)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rpitch) return;
    for (i64 v = blockIdx.y; v < nz * nb; v += gridDim.y) {
        const i64 b = v / nz, k = v - b * nz;
        double va = 0.0, vb = 0.0, vc = 0.0, vf = 0.0;
        if (j < ny) {
            const double *n2 = N2 + b * s0 + k * s1;
            va = rows[j];
            vb = n2[j * s2] * rows[ny + j];
            vc = n2[j * s2] / rows[2 * ny + j];
            if ((k >= 1) && (k <= nz - 2) && (j >= 1) && (j <= ny - 2)) {
                const double Bn = n2[(j + 1) * s2] * rows[ny + j + 1];
                vf = optArg / ((va + va) * ratio2Sqr + (Bn + vb) * ratio1Sqr + (vc + vc));
            }
        }
        double *d = vals + v * 4 * rpitch + j;
        d[0] = va; d[rpitch] = vb; d[2 * rpitch] = vc; d[3 * rpitch] = vf;
    }
}

// A as row values: vals[v][j] = rows[0][j] for every level v
__global__ void x3_front_rowvals_kernel(double *__restrict__ vals, const double *__restrict__ rows, i64 nv, i64 ny, i64 rpitch)
{
    const i64 j = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= rpitch) return;
    for (i64 v = blockIdx.y; v < nv; v += gridDim.y) vals[v * rpitch + j] = (j < ny) ? rows[j] : 0.0;
}

// ... and the way back: dense S := omega where the forcing was valid, out_undef on land
__global__ void x3_unpack_front_kernel(double *__restrict__ dst, const double *__restrict__ buf0,
                                       const double *__restrict__ buf1, const double *__restrict__ F, i64 rows, i64 nx,
                                       i64 pitch, i64 nb, double user_undef, double out_undef, double undef,
                                       const XdSliceState *__restrict__ st)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    for (i64 row = blockIdx.y; row < rows * nb; row += gridDim.y) {
        const i64 b = row / rows;
        const double *src = st[b].cur ? buf1 : buf0;
        const double f = F[row * nx + i];
        dst[row * nx + i] = x3_land(f, user_undef, undef) ? out_undef : src[row * pitch + XM_PADL + i];
    }
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
// kernel variants: tile height TJ (TJ - 4 owned rows, TJ warps), ring depth K, CTAs per SM the register
// budget is cut for.  Shared memory per CTA = K stages of (6 TJ - 9) rows of 512 bytes (AROW: 5 TJ - 7)
// + 2 exchange buffers of TJ rows.
struct X3Variant { int TJ, K, MINB; };
static const X3Variant X3_VARIANTS[] = {
    {16, 4, 1},   // 0: 12 owned rows, 512 threads, 190 KB: large volumes
    {12, 4, 1},   // 1:  8 owned rows, 384 threads, 138 KB
    {12, 3, 2},   // 2:  8 owned rows, 2 CTAs per SM (24 warps), 106 KB each
    {8, 3, 3},    // 3:  4 owned rows, 3 CTAs per SM (24 warps), 67 KB each
    {8, 4, 2},    // 4:  4 owned rows, 2 CTAs per SM, 87 KB each
    {16, 3, 1},   // 5: 12 owned rows, shallower ring, 146 KB
    {12, 5, 1},   // 6: deeper rings
    {12, 6, 1},   // 7
    {8, 8, 1},    // 8
    {8, 5, 2},    // 9
    {20, 3, 1},   // 10: 16 owned rows, 640 threads
    {24, 2, 1},   // 11
};
#define X3_NVARIANTS ((int)(sizeof(X3_VARIANTS) / sizeof(X3_VARIANTS[0])))

struct Fused3Plan {
    bool built = false;
    X3Front front;
    int variant = 0;
    bool arow = false;             // A constant along x: AROW kernels
    bool rowsmode = false;         // ... and so are B and C: X3_ROWS kernels (A, B, C, fac as row values)
    int cm() const { return rowsmode ? X3_ROWS : arow ? X3_AROW : X3_DENSE; }
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

template <int TJ, int K, int CM>
static size_t x3_smem_bytes()
{
    const size_t stage = (size_t)X3Lay<TJ, CM>::STAGE * sizeof(double);
    return (size_t)K * stage + (size_t)2 * TJ * X3_W * sizeof(double) + (size_t)TJ * 16 + (size_t)K * 16 + 16 + sizeof(XdSliceState) + 8;
}
template <int TJ, int K, int MINB, int CM>
static cudaError_t x3_prepare(size_t *smem, int *blocks_per_sm)
{
    *smem = x3_smem_bytes<TJ, K, CM>();
    cudaError_t e = cudaFuncSetAttribute(xm3_std3d_kernel<TJ, K, MINB, CM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)*smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, xm3_std3d_kernel<TJ, K, MINB, CM>, TJ * 32, *smem);
}
template <int TJ, int K, int MINB, int CM>
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
    return cudaLaunchKernelEx(&cfg, xm3_std3d_kernel<TJ, K, MINB, CM>, p.mS[0], p.mS[1], p.mA, p.mB, p.mC, p.mFd, p.mFac, p.args);
}

#define X3_DISPATCH1(v, AR, CALL)                     \
    switch (v) {                                      \
    case 0: CALL(16, 4, 1, AR); break;                \
    case 1: CALL(12, 4, 1, AR); break;                \
    case 2: CALL(12, 3, 2, AR); break;                \
    case 3: CALL(8, 3, 3, AR); break;                 \
    case 4: CALL(8, 4, 2, AR); break;                 \
    case 5: CALL(16, 3, 1, AR); break;                \
    case 6: CALL(12, 5, 1, AR); break;                \
    case 7: CALL(12, 6, 1, AR); break;                \
    case 8: CALL(8, 8, 1, AR); break;                 \
    case 9: CALL(8, 5, 2, AR); break;                 \
    case 10: CALL(20, 3, 1, AR); break;               \
    default: CALL(24, 2, 1, AR); break;               \
    }
#define X3_DISPATCH(v, cm, CALL)                                              \
    if ((cm) == X3_ROWS) { X3_DISPATCH1(v, X3_ROWS, CALL) }                    \
    else if ((cm) == X3_AROW) { X3_DISPATCH1(v, X3_AROW, CALL) }               \
    else { X3_DISPATCH1(v, X3_DENSE, CALL) }

// Tile shape.  Measured on B200 (37 x 180 x 360 and 300 x 300 x 602 volumes): a step of the march costs about
// 0.83 us with 16-row tiles (ring of 4), 0.72 us with 12-row tiles (ring of 5) and 0.78 us with two co-resident
// 8-row tiles per SM (ring of 5) -- nearly independent of the volume -- so the cost of a pass is about
// (rounds of tiles over the SM slots) x (levels + drain + pipeline fill) x that.  Splitting the levels over
// several tiles (XINV_FUSED3_NTZ) never won in the measurements and is off by default.
static void x3_choose(i64 nz, i64 ny, i64 nx, i64 batch, int sm_count, int *variant, int *ntz)
{
    const i64 ntx = (nx + X3_W - 5) / (X3_W - 4);
    static const int cand[] = {0, 6, 9};
    static const double t_step[] = {0.83, 0.72, 0.78};
    double best = 1e300;
    *variant = 0; *ntz = 1;
    for (int ci = 0; ci < 3; ++ci) {
        const X3Variant v = X3_VARIANTS[cand[ci]];
        const int RB = v.TJ - 4;
        const i64 tiles = ntx * ((ny + RB - 1) / RB) * batch;
        const double slots = (double)sm_count * v.MINB;
        const double rounds = (tiles <= slots) ? 1.0 : (double)tiles / slots;
        const double cost = rounds * ((double)nz + 5.0) * t_step[ci];
        if (cost < best - 1e-9) { best = cost; *variant = cand[ci]; }
    }
}

// Row-value kernels (X3_ROWS): a step costs 0.63 us with 16-row tiles, 0.54 us with 12-row tiles, 0.85 us with 20-row
// tiles, plus about 2.7 us per pass (measured at 37 x 180 x 360 and 300 x 300 x 602); tiles are dealt round-robin, so
// the rounds are whole; with omega and F as the only volumes the halo levels of a level range are cheap, and splitting
// the levels in two pays when it brings the tiles of a small volume to one round of the machine
// (37 x 180 x 360: 24.6 -> 21.6 us per sweep).
static void x3_choose_rows(i64 nz, i64 ny, i64 nx, i64 batch, int sm_count, int *variant, int *ntz)
{
    const i64 ntx = (nx + X3_W - 5) / (X3_W - 4);
    static const int cand[] = {6, 0, 10};
    static const double t_step[] = {0.54, 0.63, 0.85};
    double best = 1e300;
    *variant = 6; *ntz = 1;
    for (int nt = 1; nt <= 2; ++nt) {
        i64 ZB = (nz + nt - 1) / nt;
        ZB += (ZB & 1);
        if (nt > 1 && ZB < 8) continue;
        const i64 nzt = (nz + ZB - 1) / ZB;
        const double steps = (nzt == 1) ? (double)nz + 2.0 : (double)ZB + 2.0;
        for (int ci = 0; ci < 3; ++ci) {
            const X3Variant v = X3_VARIANTS[cand[ci]];
            const int RB = v.TJ - 4;
            const i64 tiles = ntx * ((ny + RB - 1) / RB) * nzt * batch;
            const i64 slots = (i64)sm_count * v.MINB;
            const double rounds = (double)((tiles + slots - 1) / slots);
            const double cost = rounds * steps * t_step[ci] + 2.7;
            if (cost < best * 0.98) { best = cost; *variant = cand[ci]; *ntz = nt; }
        }
    }
}

static inline int fused3_plan_build(Fused3Plan &p, XmWork &work, int sm_count, const XdGeom &g, const XdCoef &q, i64 batch,
                                    double *dS, cudaStream_t stream, std::string &why, const X3Front *front = nullptr)
{
    fused3_plan_release(p);
    if (front) p.front = *front;
    const bool fe = p.front.on;                  // front end: row vectors + N2 + user forcing, zero initial guess
    const i64 nz = g.nz, ny = g.ny, nx = g.nx;
    const i64 pitch = ((XM_PADL + nx + XM_GHOST) + 3) / 4 * 4;
    const int periodic = (g.bcx == XD_BC_PERIODIC);
    const i64 rows = nz * ny;
    const size_t vol_bytes = (size_t)rows * pitch * sizeof(double);
    const int feb = fe && p.front.ns[0] != 0;    // front end: N2 (hence B, C and the factor) has a batch axis
    const int cb[4] = {fe ? 0 : q.cs[0] != 0, fe ? feb : q.cs[1] != 0, fe ? feb : q.cs[2] != 0, fe ? 1 : q.cs[3] != 0};
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
    void *flag;
    X3_ALLOC(flag, 7, 16);
    dim3 blk(128);
    auto gridfor = [&](i64 cols, i64 nrows) {
        i64 gy = nrows;
        if (gy > 32768) gy = 32768;
        return dim3((unsigned)((cols + 127) / 128), (unsigned)gy, 1);
    };
    // ---- is A constant along x?  (one pass over it, one 4-byte read-back) ----
    {
        const char *ea = getenv("XINV_FUSED3_AROW");
        p.arow = fe || !(ea && atoi(ea) == 0);
        if (fe) {
            cudaMemsetAsync(flag, 0, 8, stream);
        } else if (p.arow) {
            const i64 nrA = rows * (cb[0] ? batch : 1);
            cudaMemsetAsync(flag, 0, 4, stream);
            x3_rowconst_kernel<<<gridfor(nx, nrA), blk, 0, stream>>>(q.c[0], nrA, nx, (int *)flag);
            int h = 1;
            if ((e = cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(stream)) != cudaSuccess) {
                why = std::string("row-constancy check: ") + cudaGetErrorString(e);
                fused3_plan_release(p);
                return -1;
            }
            p.arow = (h == 0);
        }
        // ---- ... and B and C too?  (front end: N2 without a column axis; XINV_FUSED3_ROWS=0 switches the mode off) ----
        const char *er = getenv("XINV_FUSED3_ROWS");
        p.rowsmode = p.arow && !(er && atoi(er) == 0);
        if (fe) {
            p.rowsmode = p.rowsmode && (p.front.ns[3] == 0);
        } else if (p.rowsmode) {
            const i64 nrB = rows * (cb[1] ? batch : 1), nrC = rows * (cb[2] ? batch : 1);
            x3_rowconst_kernel<<<gridfor(nx, nrB), blk, 0, stream>>>(q.c[1], nrB, nx, (int *)flag);
            x3_rowconst_kernel<<<gridfor(nx, nrC), blk, 0, stream>>>(q.c[2], nrC, nx, (int *)flag);
            int h = 1;
            if ((e = cudaMemcpyAsync(&h, flag, 4, cudaMemcpyDeviceToHost, stream)) != cudaSuccess ||
                (e = cudaStreamSynchronize(stream)) != cudaSuccess) {
                why = std::string("row-constancy check: ") + cudaGetErrorString(e);
                fused3_plan_release(p);
                return -1;
            }
            p.rowsmode = (h == 0);
        }
    }
    const bool rm = p.rowsmode;
    const i64 rpitch = (ny + 1) / 2 * 2;         // row-value vectors: TMA strides are multiples of 16 bytes
    const i64 nvA = nz * ((rm ? cbFac : cb[0]) ? batch : 1);    // (volume x level) planes of A (X3_ROWS: of A, B, C, fac)
    X3_ALLOC(p.bufA, 2, p.arow ? (size_t)nvA * (rm ? 4 : 1) * rpitch * sizeof(double) : vol_bytes * (cb[0] ? batch : 1));
    X3_ALLOC(p.bufC, 3, rm ? 16 : vol_bytes * (cb[2] ? batch : 1));
    X3_ALLOC(p.bufFd, 4, vol_bytes * (cbFd ? batch : 1));
    X3_ALLOC(p.bufFac, 5, rm ? 16 : vol_bytes * (cbFac ? batch : 1));
    X3_ALLOC(p.bufB, 6, rm ? 16 : vol_bytes * (cb[1] ? batch : 1));
#undef X3_ALLOC
    int ntz = 1;
    {
        if (rm) x3_choose_rows(nz, ny, nx, batch, sm_count, &p.variant, &ntz);
        else    x3_choose(nz, ny, nx, batch, sm_count, &p.variant, &ntz);
        const char *env = getenv("XINV_FUSED3_VARIANT");
        if (env) { p.variant = atoi(env); ntz = 1; }
        if (p.variant < 0 || p.variant >= X3_NVARIANTS) p.variant = 0;
        const char *ez = getenv("XINV_FUSED3_NTZ");
        if (ez && atoi(ez) > 0) ntz = atoi(ez);
    }
    const X3Variant v = X3_VARIANTS[p.variant];
    auto pack = [&](void *dst, const double *src, i64 bstride, i64 nb) {
        x3_pack_kernel<<<gridfor(pitch, rows * nb), blk, 0, stream>>>((double *)dst, src, rows, nx, pitch, bstride, nb, periodic);
    };
    if (fe) {
        const X3Front &f = p.front;
        cudaMemsetAsync(p.bufS[0], 0, vol_bytes * batch, stream);      // zero initial guess (apps.py:2145), ghosts included
        cudaMemsetAsync(p.bufS[1], 0, vol_bytes * batch, stream);
        if (rm) {
            x3_front_rows4_kernel<<<gridfor(rpitch, nvA), blk, 0, stream>>>((double *)p.bufA, f.rows, f.N2, f.ns[0], f.ns[1], f.ns[2],
                                                                            nz, ny, rpitch, feb ? batch : 1, q.p[1], q.p[2], q.optArg);
        } else {
            x3_front_rowvals_kernel<<<gridfor(rpitch, nvA), blk, 0, stream>>>((double *)p.bufA, f.rows, nvA, ny, rpitch);
            x3_front_bc_kernel<<<gridfor(pitch, rows * (feb ? batch : 1)), blk, 0, stream>>>(
                (double *)p.bufB, (double *)p.bufC, f.rows, f.N2, f.ns[0], f.ns[1], f.ns[2], f.ns[3], nz, ny, nx, pitch,
                feb ? batch : 1, periodic);
        }
        x3_front_derived_kernel<<<gridfor(pitch, rows * batch), blk, 0, stream>>>(
            (double *)p.bufFd, rm ? nullptr : (double *)p.bufFac, f.rows, f.N2, f.ns[0], f.ns[1], f.ns[2], f.ns[3], f.F, nz, ny, nx, pitch,
            batch, feb ? batch : 1, periodic, f.user_undef, q.undef, q.p[0], q.p[1], q.p[2], q.optArg, (int *)flag);
    } else {
        pack(p.bufS[0], dS, g.N, batch);
        pack(p.bufS[1], dS, g.N, batch);   // levels 0 / nz-1 and all pad columns of both buffers start identical
        if (rm) {
            x3_pack_rows4_kernel<<<gridfor(rpitch, nvA), blk, 0, stream>>>((double *)p.bufA, q, nz, ny, nx, rpitch, cbFac ? batch : 1);
        } else {
            if (p.arow) x3_pack_rowvals_kernel<<<gridfor(rpitch, nvA), blk, 0, stream>>>((double *)p.bufA, q.c[0], nvA, ny, nx, rpitch);
            else        pack(p.bufA, q.c[0], q.cs[0], cb[0] ? batch : 1);
            pack(p.bufB, q.c[1], q.cs[1], cb[1] ? batch : 1);
            pack(p.bufC, q.c[2], q.cs[2], cb[2] ? batch : 1);
        }
        x3_pack_derived_kernel<<<gridfor(pitch, rows * (cbFd ? batch : 1)), blk, 0, stream>>>(
            (double *)p.bufFd, rm ? nullptr : (double *)p.bufFac, q, nz, ny, nx, pitch, cbFd ? batch : 1, cbFac ? batch : 1, periodic);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) {
        why = std::string("pack kernels: ") + cudaGetErrorString(e);
        fused3_plan_release(p);
        return -1;
    }
    // tensor maps: (pitch, ny, levels x volumes); boxes 64 columns x TJ rows (omega), TJ-1 (B), TJ-2 (A, C, Fd, fac)
    if (xf_make_map(&p.mS[0], p.bufS[0], pitch, ny, nz * batch, X3_W, v.TJ, why) ||
        xf_make_map(&p.mS[1], p.bufS[1], pitch, ny, nz * batch, X3_W, v.TJ, why) ||
        (p.arow ? xf_make_row_map(&p.mA, p.bufA, ny, rpitch, nvA, v.TJ, rm ? 4 : 1, why)
                : xf_make_map(&p.mA, p.bufA, pitch, ny, nvA, X3_W, v.TJ - 2, why)) ||
        xf_make_map(&p.mFd, p.bufFd, pitch, ny, nz * (cbFd ? batch : 1), X3_W, v.TJ - 2, why) ||
        (!rm && (xf_make_map(&p.mB, p.bufB, pitch, ny, nz * (cb[1] ? batch : 1), X3_W, v.TJ - 1, why) ||
                 xf_make_map(&p.mC, p.bufC, pitch, ny, nz * (cb[2] ? batch : 1), X3_W, v.TJ - 2, why) ||
                 xf_make_map(&p.mFac, p.bufFac, pitch, ny, nz * (cbFac ? batch : 1), X3_W, v.TJ - 2, why)))) {
        fused3_plan_release(p);
        return -1;
    }
    if (rm) p.mB = p.mC = p.mFac = p.mFd;        // unused by the X3_ROWS kernels
    X3Args &a = p.args;
    a.Sbuf[0] = (double *)p.bufS[0];
    a.Sbuf[1] = (double *)p.bufS[1];
    a.pitch = pitch; a.plane = ny * pitch; a.slice = nz * ny * pitch;
    a.nz = (int)nz; a.ny = (int)ny; a.nx = (int)nx;
    a.RB = v.TJ - 4;
    a.ntx = (int)((nx + X3_W - 5) / (X3_W - 4));
    a.nty = (int)((ny + a.RB - 1) / a.RB);
    if (ntz < 1) ntz = 1;
    a.ZB = (int)((nz + ntz - 1) / ntz);
    a.ZB += (a.ZB & 1);                         // level ranges start on even levels
    if (a.ZB < 2) a.ZB = 2;
    a.ntz = (int)((nz + a.ZB - 1) / a.ZB);
    a.batch = (int)batch;
    a.bcy = g.bcy; a.bcx = g.bcx;
    a.cbA = rm ? cbFac : cb[0]; a.cbB = cb[1]; a.cbC = cb[2]; a.cbFd = cbFd; a.cbFac = cbFac;
    a.r2 = q.p[1]; a.r1 = q.p[2]; a.undef = q.undef;
    p.batch = batch;
    p.nblk_partials = 2 * a.ntx * a.nty * a.ntz;       // (two sets of slots: replicated loop control alternates between them)
    const i64 tiles = (i64)a.ntx * a.nty * a.ntz * batch;
    int blocks_per_sm = 0;
#define X3_PREP(TJ_, K_, MB_, AR_) e = x3_prepare<TJ_, K_, MB_, AR_>(&p.smem, &blocks_per_sm)
    X3_DISPATCH(p.variant, p.cm(), X3_PREP);
#undef X3_PREP
    if (e != cudaSuccess || blocks_per_sm < 1) {
        why = std::string("3-D fused kernel does not fit: ") + cudaGetErrorString(e);
        fused3_plan_release(p);
        return -1;
    }
    {
        const i64 slots = (i64)sm_count * blocks_per_sm;
        p.grid = (int)(tiles < slots ? tiles : slots);
        if (p.grid < 1) p.grid = 1;
        int can_coop = 0, dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&can_coop, cudaDevAttrCooperativeLaunch, dev);
        p.coop = can_coop != 0;
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
#define X3_GO(TJ_, K_, MB_, AR_) e = x3_launch<TJ_, K_, MB_, AR_>(p, stream)
    if (npass > 1) {
        a.npass = npass;
        a.gbar_base = p.gbar_base;
        X3_DISPATCH(p.variant, p.cm(), X3_GO);
        if (e == cudaSuccess) {
            // arrivals per CTA and launch: one per pass boundary, one more behind the last pass when the loop control is replicated
            const bool repl = (i64)a.ntx * a.nty * a.ntz * a.batch <= (i64)p.grid;
            p.gbar_base += (unsigned long long)p.grid * (unsigned long long)(repl ? npass : npass - 1);
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
        X3_DISPATCH(p.variant, p.cm(), X3_GO);
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
    if (p.front.on)
        x3_unpack_front_kernel<<<grid, 128, 0, stream>>>(dS, a.Sbuf[0], a.Sbuf[1], p.front.F, rows, a.nx, a.pitch, p.batch,
                                                         p.front.user_undef, p.front.out_undef, a.undef, st);
    else
        x3_unpack_kernel<<<grid, 128, 0, stream>>>(dS, a.Sbuf[0], a.Sbuf[1], rows, a.nx, a.pitch, p.batch, st);
    return 0;
}
