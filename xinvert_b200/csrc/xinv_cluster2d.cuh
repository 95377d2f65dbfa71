// xinv_cluster2d.cuh -- XINV_ENGINE_CLUSTER: the whole solve of a small or medium 2-D slice inside
// ONE THREAD-BLOCK CLUSTER, psi resident in registers and distributed shared memory.
//
// For slices of up to ~10^5 cells (BASELINE configs[0], 360 x 180; reanalysis-sized 144 x 73 ...)
// the marching engine (xinv_march2d.cuh) is bound by its fixed cost per pass -- a grid-wide barrier,
// the final reduction, the pipeline refill: 4.8 us per sweep on 360 x 180 where the arithmetic needs
// a fraction of a microsecond.  Here the slice never leaves the SMs between sweeps:
//   * the rows of the slice are dealt out to the R CTAs of a cluster (R = 1, 2, 4, 8 or 16), a row to
//     WPR warps, a run of K consecutive cells to a thread;
//   * a thread keeps psi and Fd (general form: Gm) of its K cells IN REGISTERS for the whole solve,
//     together with the row values A[j], A[j+1], C[j], fac[j] (general form: A, C, D, E, F, fac):
//     the east/west neighbours of a cell are registers of the same thread, only the first and last
//     cell of the run look into the neighbouring thread's run;
//   * updated values are published to a shared-memory tile of the CTA's rows (+ one halo row above
//     and below) for the north/south neighbours; a value in the first / last row of a CTA is also
//     pushed straight into the halo row of the neighbouring CTA's tile through distributed shared
//     memory (st to a mapa-translated address), so every load of the hot loop is a local LDS;
//   * no cluster barrier in the loop: the pushes are st.async stores that complete a transaction count
//     on an mbarrier in the RECEIVER's shared memory, so a CTA only waits -- in the warps of its first
//     and last row -- for the halo values of its two neighbours; inside the CTA one __syncthreads per
//     colour orders the tile;
//   * the partial sums of |psi| travel the same way: every warp pushes its (sum, count) to a slot in
//     EVERY CTA of the cluster, every warp adds the slots in a fixed order once they have arrived, so
//     all threads of the cluster hold the same mean|psi| and run the reference's loop control
//     (numbas.py:401-414) redundantly -- no further synchronisation for the stop test;
//   * the tile is laid out with one pad slot per run (row pitch of a run: K + 1 doubles, odd), which
//     makes the strided accesses of a half-warp fall into distinct banks.
// Operands are the padded / derived copies of the fused plan (xinv_march2d.cuh: Fd with the skip
// marker, row values with the factor), so everything the marching engine accepts in its RC form
// -- invert_Poisson, invert_GillMatsuno, invert_Stommel and their device front ends -- runs here
// when the slice fits; arithmetic per cell: xm_eval / xm_eval_gen operation for operation
// (numbas.py:351-369 with B == 0; :1132-1153), compiled with -fmad=false: the iterates are
// bit-identical to the marching engine's, the colour engine's and the oracle's.
// A batch runs one slice per cluster at a time (persistent clusters), each stopping on its own test.
#pragma once
#include <cooperative_groups.h>
#include "xinv_march2d.cuh"

namespace cg = cooperative_groups;

// Norm partials of sweep n land in slot set n % XC_SLOTS of every CTA.  A CTA's control warp reads set n during its
// sweep n+1; a CTA d hops away can run at most ~d/2 sweeps ahead of this CTA's compute warps (the halo exchange chains
// them), which wait for the control warp at every sweep's CTA barrier -- so a set is rewritten (by sweep n + XC_SLOTS)
// only if the control warp lagged its own compute warps by microseconds while holding no lock; four sets put that
// beyond what a resident, non-preempted kernel can do (two would do in practice).
#define XC_SLOTS 4u

struct XcArgs {
    double *Sbuf[2];          // padded psi buffers of the fused plan [batch][ny][pitch]
    const double *Fd;         // [cbFd ? batch : 1][ny][pitch]
    const double *rows;       // [cbRow ? batch : 1][NV][rpitch]
    i64 pitch, slice, rpitch;
    int ny, nx, batch;
    int bcy, bcx;
    int cbFd, cbRow;
    double ratioSqr, undef;
    double ratio, delx, delxSqr;      // general form only
    XdSliceState *st;
    int *nactive;
    double tol;
    i64 mxLoop;
    int zero_exit;
    int nsweeps;              // sweep budget of this launch
    int R;                    // CTAs per cluster
    int WPR, RPR;             // warps per row, runs per row
    int RPmax;                // rows of the tile besides the two halo rows (= max rows per CTA)
    int TP;                   // tile pitch (doubles)
    int NW;                   // warps per CTA
    int accel;                // XINV_ACCEL_CHEBYSHEV: omega varies from half sweep to half sweep (xinv.h)
    double rho2;
};

template <int KIND> struct XcRow;
template <> struct XcRow<0> { double Ac, An, C, fac; };
template <> struct XcRow<1> { double A, C, D, E, F, fac; };

// ---- DSMEM plumbing --------------------------------------------------------------------------------
// A value for another CTA of the cluster travels as st.async (SASS: STAS): a store into the peer's shared
// memory that completes a transaction count on an mbarrier IN THE PEER'S shared memory.  The consumer waits
// on its own mbarrier, so data and "it has arrived" come as one message and no fence is needed -- the
// release/acquire pair of barrier.cluster costs a MEMBAR.ALL.GPU per barrier with this toolchain
// (cuobjdump), 3.5 us per sweep on 360 x 180 against 1.x us this way.
__device__ __forceinline__ uint32_t xc_mapa(uint32_t saddr, uint32_t rank)
{
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(saddr), "r"(rank));
    return r;
}
__device__ __forceinline__ void xc_st_async(uint32_t raddr, double v, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.f64 [%0], %1, [%2];"
                 ::"r"(raddr), "d"(v), "r"(rbar) : "memory");
}
__device__ __forceinline__ void xc_st_async2(uint32_t raddr, double v, double w, uint32_t rbar)
{
    asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v2.f64 [%0], {%1, %2}, [%3];"
                 ::"r"(raddr), "d"(v), "d"(w), "r"(rbar) : "memory");
}

// named barrier 1: the compute warps of a CTA among themselves (the control warp is not held up by it, nor they by it)
__device__ __forceinline__ void xc_compute_sync(int nthreads) { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); }

// one colour of a thread's run: cells m = PAR, PAR + 2, ... (< K)
template <int KIND, int K, int PAR>
__device__ __forceinline__ void xc_half(double (&psi)[K], const double (&fd)[K], const XcRow<KIND> &cr, const XcArgs &a,
                                        double *Tc, const double *Tn, const double *Ts, int col0, int westIdx, int eastIdx,
                                        bool act, bool ghostE, bool ghostW, int ghostEidx,
                                        uint32_t rN, uint32_t rbN, uint32_t rS, uint32_t rbS, double fac)
{
    constexpr int NC = (K - PAR + 1) / 2;
    constexpr bool LASTIN = ((K - 1 - PAR) % 2) == 0;          // is cell K-1 of this colour?
    double sn[NC], ss[NC];
    #pragma unroll
    for (int c = 0; c < NC; ++c) {
        sn[c] = Tn[col0 + PAR + 2 * c];
        ss[c] = Ts[col0 + PAR + 2 * c];
    }
    double sw = 0.0, se = 0.0;
    if (PAR == 0) sw = Tc[westIdx];
    if (LASTIN) se = Tc[eastIdx];
    // The cells of a colour are independent: the arithmetic is written stage by stage ACROSS the cells so that the
    // dependent chain of one cell (8 operations deep) is interleaved with the others' (the compiler keeps this order).
    // Per cell the operations and their order are those of xm_eval / xm_eval_gen (numbas.py:351-369, :1132-1153).
    double t[NC];
    if constexpr (KIND == 0) {
        const XcRow<0> &r = cr;
        double a1[NC], a2[NC], a3[NC], a4[NC];
        #pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int m = PAR + 2 * c;
            const double So = psi[m];
            const double Sw = (m == 0) ? sw : psi[m > 0 ? m - 1 : 0];
            const double Se = (m == K - 1) ? se : psi[m < K - 1 ? m + 1 : K - 1];
            a1[c] = sn[c] - So; a2[c] = So - ss[c]; a3[c] = Se - So; a4[c] = So - Sw;
        }
        #pragma unroll
        for (int c = 0; c < NC; ++c) { a1[c] = r.An * a1[c]; a2[c] = r.Ac * a2[c]; a3[c] = r.C * a3[c]; a4[c] = r.C * a4[c]; }
        #pragma unroll
        for (int c = 0; c < NC; ++c) { a1[c] = a1[c] - a2[c]; a3[c] = a3[c] - a4[c]; }
        #pragma unroll
        for (int c = 0; c < NC; ++c) a1[c] = a1[c] * a.ratioSqr;
        #pragma unroll
        for (int c = 0; c < NC; ++c) t[c] = a1[c] + a3[c];
        #pragma unroll
        for (int c = 0; c < NC; ++c) t[c] = t[c] - fd[PAR + 2 * c];
        #pragma unroll
        for (int c = 0; c < NC; ++c) t[c] = t[c] * fac;
    } else {
        const XcRow<1> &g = cr;
        #pragma unroll
        for (int c = 0; c < NC; ++c) {
            const int m = PAR + 2 * c;
            const double So = psi[m];
            const double Sw = (m == 0) ? sw : psi[m > 0 ? m - 1 : 0];
            const double Se = (m == K - 1) ? se : psi[m < K - 1 ? m + 1 : K - 1];
            const double Snn = sn[c], Sss = ss[c];
            double temp = g.A * ((Snn - So) - (So - Sss)) * a.ratioSqr;
            temp = temp + g.C * ((Se - So) - (So - Sw));
            temp = temp + (g.D * (Snn - Sss) * a.ratio + g.E * (Se - Sw)) * a.delx / 2.0;
            temp = temp + (g.F * So - fd[m]) * a.delxSqr;
            t[c] = temp * fac;
        }
    }
    #pragma unroll
    for (int c = 0; c < NC; ++c) {
        const int m = PAR + 2 * c;
        const bool upd = __double2hiint(fd[m]) != XM_SKIP_HI;
        const double nv = psi[m] + t[c];
        psi[m] = upd ? nv : psi[m];
    }
    if (act) {
        #pragma unroll
        for (int c = 0; c < NC; ++c) Tc[col0 + PAR + 2 * c] = psi[PAR + 2 * c];
        if (PAR == 0 && ghostE) Tc[ghostEidx] = psi[0];            // column 0 -> the east ghost (periodic-x)
        if (LASTIN && ghostW) Tc[0] = psi[K - 1];                  // column nx-1 -> the west ghost
        if (rN) {                                                  // into the neighbouring CTAs' halo rows (DSMEM)
            #pragma unroll
            for (int c = 0; c < NC; ++c) xc_st_async(rN + 8u * (uint32_t)(col0 + PAR + 2 * c), psi[PAR + 2 * c], rbN);
        }
        if (rS) {
            #pragma unroll
            for (int c = 0; c < NC; ++c) xc_st_async(rS + 8u * (uint32_t)(col0 + PAR + 2 * c), psi[PAR + 2 * c], rbS);
        }
    }
}

// The block has NW compute warps (a row segment each) and ONE CONTROL WARP (the last one): it owns no cells, posts the
// transaction counts, and -- while the compute warps are busy with colour 0 of the next sweep -- waits for the norm
// partials of the sweep just finished, adds them up and runs the reference's loop control; the compute warps pick the
// verdict up at the CTA barrier that ends colour 0.
template <int KIND, int K>
__global__ void __launch_bounds__((K <= 6 ? 512 : K <= 8 ? 448 : K <= 12 ? 416 : 320), 1)
xc_cluster_kernel(const XcArgs a)
{
    static_assert(K % 2 == 0, "runs hold whole red/black pairs");
    extern __shared__ __align__(16) unsigned char xc_smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int R = a.R;
    const int rank = (int)cluster.block_rank();
    const int cid = blockIdx.x / R, ncl = gridDim.x / R;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, nth = blockDim.x;
    const int ny = a.ny, nx = a.nx, TP = a.TP, NW = a.NW;
    const bool ctrl = (warp == NW);
    double *T = reinterpret_cast<double *>(xc_smem);                       // [(RPmax + 2)][TP]
    double2 *slot = reinterpret_cast<double2 *>(T + (size_t)(a.RPmax + 2) * TP);   // [XC_SLOTS][R * NW]: (sum, count) per compute warp of the cluster
    double *bkT = reinterpret_cast<double *>(slot + XC_SLOTS * R * NW);           // [K][threads]: every cell's value before a speculative sweep
    uint64_t *hb = reinterpret_cast<uint64_t *>(bkT + (size_t)K * nth);   // [2] halo values of colour 0 / 1 have arrived
    uint64_t *nb = hb + 2;                                                 // [2] norm partials have arrived (by parity of the sweep)
    int *verdict = reinterpret_cast<int *>(nb + 2);                        // [2] "the slice goes on" by parity of the judged sweep, written by the control warp
    if (tid == 0) {
        xf_mbar_init(hb, 1); xf_mbar_init(hb + 1, 1); xf_mbar_init(nb, 1); xf_mbar_init(nb + 1, 1);
        xf_fence_barrier_init();
    }
    cluster.sync();

    const int j0 = (int)(((i64)rank * ny) / R), j1 = (int)(((i64)(rank + 1) * ny) / R);
    const int nrows = j1 - j0;
    const int jl = warp / a.WPR;
    const int q = (warp - jl * a.WPR) * 32 + lane;
    const bool act = !ctrl && (jl < nrows) && (q < a.RPR);
    const int j = j0 + (jl < nrows ? jl : 0);
    const int i0 = q * K;
    int nvalid = nx - i0;
    if (nvalid > K) nvalid = K;
    if (!act || nvalid < 0) nvalid = 0;
    const bool periodic = (a.bcx == XD_BC_PERIODIC);
    const int col0 = 1 + q * (K + 1);
    const int westIdx = (q > 0) ? col0 - 2 : 0;
    const bool lastrun = act && (i0 + nvalid == nx);
    const int ghostEidx = 1 + (a.RPR - 1) * (K + 1) + (nx - (a.RPR - 1) * K);   // the slot east of column nx-1
    const int eastIdx = lastrun ? ghostEidx : col0 + K + 1;
    const bool ghostE = periodic && act && (q == 0);
    const bool ghostW = periodic && lastrun;
    double *Tc = T + (size_t)((ctrl ? 0 : jl) + 1) * TP;
    const double *Tn = Tc + TP, *Ts = Tc - TP;
    // halo rows of the neighbouring CTAs that mirror this thread's row, and the neighbours' halo barriers
    const bool hasN = rank + 1 < R, hasS = rank > 0;
    const bool edgeN = !ctrl && hasN && jl == nrows - 1, edgeS = !ctrl && hasS && jl == 0;      // warp-uniform
    uint32_t rN = 0, rS = 0, rbN[2] = {0, 0}, rbS[2] = {0, 0};
    if (edgeN) {
        rN = xc_mapa(xf_smem_u32(T), rank + 1);                                  // their row j0' - 1
        rbN[0] = xc_mapa(xf_smem_u32(hb), rank + 1); rbN[1] = xc_mapa(xf_smem_u32(hb + 1), rank + 1);
    }
    if (edgeS) {
        const int pj0 = (int)(((i64)(rank - 1) * ny) / R);
        rS = xc_mapa(xf_smem_u32(T + (size_t)(j0 - pj0 + 1) * TP), rank - 1);    // their row j1'
        rbS[0] = xc_mapa(xf_smem_u32(hb), rank - 1); rbS[1] = xc_mapa(xf_smem_u32(hb + 1), rank - 1);
    }
    const uint32_t halo_bytes = (uint32_t)((hasN ? 1 : 0) + (hasS ? 1 : 0)) * (uint32_t)a.RPR * (K / 2) * 8u;
    const uint32_t norm_bytes = (uint32_t)(R * NW) * 16u;
    const double undef = a.undef;
    const int par0 = j & 1;                                       // colour 0 cells of this row: m = par0, par0 + 2, ...
    const bool upd_row = act && j > 0 && j < ny - 1;              // rows 0 and ny-1 are never updated (warp-uniform)
    const bool ext_lo = act && a.bcy == XD_BC_EXTEND && j == 0;
    const bool ext_hi = act && a.bcy == XD_BC_EXTEND && j == ny - 1;
    const bool ext_cta = a.bcy == XD_BC_EXTEND && (rank == 0 || rank == R - 1);
    const bool post = ctrl && lane == 0;                          // the thread that posts transaction counts
    // (Letting the first / last row's warps go first in a half sweep -- a named barrier holding the interior warps back
    // until the pushes are on their way -- was measured slower: 2.87 vs 2.13 us per sweep on 360 x 180.)
    unsigned sp = 0;                                              // completed sweeps so far: which norm barrier, which parity
    unsigned ph = 0;                                              // sweeps so far: parity of the two halo barriers

    for (int b = cid; b < a.batch; b += ncl) {
        if (!a.st[b].active) continue;                            // the same for every thread of the cluster
        const int cur = a.st[b].cur;
        XdSliceState st_;                                         // lives in the control warp
        if (ctrl) st_ = a.st[b];
        // ---- operands of this thread: psi, Fd of its run; the row values ----
        double psi[K], fd[K];
        {
            const double *src = a.Sbuf[cur] + (i64)b * a.slice + (i64)j * a.pitch + XM_PADL + i0;
            const double *sf = a.Fd + (i64)(a.cbFd ? b : 0) * a.slice + (i64)j * a.pitch + XM_PADL + i0;
            #pragma unroll
            for (int m = 0; m < K; ++m) {
                psi[m] = (m < nvalid) ? src[m] : undef;
                fd[m] = (m < nvalid) ? sf[m] : xm_skip_value();
            }
        }
        XcRow<KIND> cr;
        {
            const double *rv = a.rows + (i64)(a.cbRow ? b : 0) * (KIND == 0 ? 3 : 6) * a.rpitch;
            if constexpr (KIND == 0) {
                XcRow<0> &r = cr;
                r.Ac = rv[j]; r.An = rv[j + 1 < ny ? j + 1 : j]; r.C = rv[a.rpitch + j]; r.fac = rv[2 * a.rpitch + j];
            } else {
                XcRow<1> &g = cr;
                g.A = rv[j]; g.C = rv[a.rpitch + j]; g.D = rv[2 * a.rpitch + j]; g.E = rv[3 * a.rpitch + j];
                g.F = rv[4 * a.rpitch + j]; g.fac = rv[5 * a.rpitch + j];
            }
        }
        // Chebyshev (xinv_opts.accel): the factor of a half sweep is omega_h / den with the denominator of the row,
        // formed as the reference forms it in every cell (numbas.py:364-367, :1151-1153; row-constant operands)
        double den = 1.0;
        if (a.accel) {
            if constexpr (KIND == 0) den = (cr.An + cr.Ac) * a.ratioSqr + (cr.C + cr.C);
            else                     den = (cr.A * a.ratioSqr + cr.C) * 2.0 - cr.F * a.delxSqr;
        }
        double om = a.st[b].omega, om_prev = om;     // omega of the next half sweep / before the sweep that may be undone
        bool om_first = (a.st[b].sweeps_done == 0), om_first_prev = om_first;
        // ---- fill the tile: own rows (with the periodic ghosts), halo rows straight from HBM ----
        if (act) {
            #pragma unroll
            for (int m = 0; m < K; ++m) Tc[col0 + m] = psi[m];
            if (ghostE) Tc[ghostEidx] = psi[0];
            if (ghostW) Tc[0] = psi[K - 1];
            if (jl == nrows - 1 && j + 1 < ny) {
                const double *src = a.Sbuf[cur] + (i64)b * a.slice + (i64)(j + 1) * a.pitch + XM_PADL + i0;
                #pragma unroll
                for (int m = 0; m < K; ++m) if (m < nvalid) Tc[TP + col0 + m] = src[m];
            }
            if (jl == 0 && j > 0) {
                const double *src = a.Sbuf[cur] + (i64)b * a.slice + (i64)(j - 1) * a.pitch + XM_PADL + i0;
                #pragma unroll
                for (int m = 0; m < K; ++m) if (m < nvalid) Tc[-TP + col0 + m] = src[m];
            }
        }
        cluster.sync();                    // once per slice: every tile is filled before a neighbour's first push can land

        // The stop test of sweep n is LAGGED by one sweep: its norm partials travel through the cluster and are added up
        // and judged by the control warp while the compute warps are already busy with sweep n+1, every cell's old value
        // parked in shared memory; if the verdict is "stop", the cells are restored and the slice ends exactly as if it
        // had been tested right away (numbas.py:401-414: fields, flags and loop counts are unchanged by the lag).
        bool pending = false;
        for (int sweep = 0; sweep < a.nsweeps; ++sweep) {
            double f0 = cr.fac, f1 = cr.fac;
            if (a.accel) {
                om_prev = om; om_first_prev = om_first;
                const double w1 = xd_cheb_next(om, a.rho2, om_first);
                f0 = om / den; f1 = w1 / den;
                om = xd_cheb_next(w1, a.rho2, false); om_first = false;
            }
            if (ctrl) {
                if (post) {
                    if (halo_bytes) { xf_mbar_expect_tx(hb, halo_bytes); xf_mbar_expect_tx(hb + 1, halo_bytes); }
                    if (R > 1) xf_mbar_expect_tx(nb + (sp & 1u), norm_bytes);
                }
                if (pending) {                                     // the stop test of the previous sweep
                    const unsigned pp = sp - 1u;
                    if (R > 1) xf_mbar_wait(nb + (pp & 1u), (pp >> 1) & 1u);
                    double ts = 0.0, tn = 0.0;
                    {
                        const double2 *sl = slot + (pp % XC_SLOTS) * (R * NW);
                        for (int k = lane; k < R * NW; k += 32) { const double2 v = sl[k]; ts += v.x; tn += v.y; }
                    }
                    #pragma unroll
                    for (int o = 16; o > 0; o >>= 1) {
                        ts += __shfl_xor_sync(0xffffffffu, ts, o);
                        tn += __shfl_xor_sync(0xffffffffu, tn, o);
                    }
                    xd_decide(st_, ts, (i64)tn, a.tol, a.mxLoop, a.zero_exit);   // the control warps of all CTAs, identically
                    if (lane == 0) verdict[pp & 1u] = st_.active;   // two slots: the control warp is a sweep ahead of a slow reader
                }
                if (halo_bytes) xf_mbar_wait(hb, ph & 1u);         // (a barrier is armed again only after its phase is over)
            } else {
                #pragma unroll
                for (int m = 0; m < K; ++m) bkT[m * nth + tid] = psi[m];
                // ---- y-"extend" rows (numbas.py:284-310): row 0 := row 1, row ny-1 := row ny-2 where != undef ----
                if (ext_cta) {
                    if (ext_lo || ext_hi) {
                        const double *Tr = ext_lo ? Tn : Ts;
                        double sv[K + 2];
                        sv[0] = Tr[westIdx];
                        #pragma unroll
                        for (int m = 0; m < K; ++m) sv[m + 1] = Tr[col0 + m];
                        sv[K + 1] = Tr[eastIdx];
                        #pragma unroll
                        for (int m = 0; m < K; ++m) {
                            double v = sv[m + 1];
                            if (!periodic) {
                                if (i0 + m == 0) v = sv[m + 2];             // S[0,0] = S[1,1]
                                if (i0 + m == nx - 1) v = sv[m];            // S[0,nx-1] = S[1,nx-2]
                            }
                            if (m < nvalid && v != undef) psi[m] = v;
                            Tc[col0 + m] = psi[m];
                        }
                    }
                    xc_compute_sync(NW * 32);                      // row 1 / ny-2 read the new rows (same CTA)
                }
                // ---- colour 0 ----
                if (upd_row) {
                    if (par0 == 0)
                        xc_half<KIND, K, 0>(psi, fd, cr, a, Tc, Tn, Ts, col0, westIdx, eastIdx, act, ghostE, ghostW, ghostEidx, rN, rbN[0], rS, rbS[0], f0);
                    else
                        xc_half<KIND, K, 1>(psi, fd, cr, a, Tc, Tn, Ts, col0, westIdx, eastIdx, act, ghostE, ghostW, ghostEidx, rN, rbN[0], rS, rbS[0], f0);
                }
                xc_compute_sync(NW * 32);
                if (edgeN || edgeS) xf_mbar_wait(hb, ph & 1u);
                // ---- colour 1, and sum|psi| / count over psi != undef (numbas.py:1710-1728): thread -> warp -> a slot in every CTA ----
                if (upd_row) {
                    if (par0 == 0)
                        xc_half<KIND, K, 1>(psi, fd, cr, a, Tc, Tn, Ts, col0, westIdx, eastIdx, act, ghostE, ghostW, ghostEidx, rN, rbN[1], rS, rbS[1], f1);
                    else
                        xc_half<KIND, K, 0>(psi, fd, cr, a, Tc, Tn, Ts, col0, westIdx, eastIdx, act, ghostE, ghostW, ghostEidx, rN, rbN[1], rS, rbS[1], f1);
                }
                double s0 = 0.0, s1 = 0.0;
                int n = 0;
                #pragma unroll
                for (int m = 0; m < K; m += 2) {                   // slots beyond the run hold undef
                    const double v = psi[m], w = psi[m + 1];
                    if (v != undef) { s0 += fabs(v); n += 1; }
                    if (w != undef) { s1 += fabs(w); n += 1; }
                }
                double sw_ = s0 + s1;
                #pragma unroll
                for (int o = 16; o > 0; o >>= 1) sw_ += __shfl_xor_sync(0xffffffffu, sw_, o);
                n = __reduce_add_sync(0xffffffffu, n);
                if (R == 1) {                                      // a cluster of one CTA: plain stores, ordered by the CTA barrier below
                    if (lane == 0) slot[(sp % XC_SLOTS) * NW + warp] = make_double2(sw_, (double)n);
                } else if (lane < R)
                    xc_st_async2(xc_mapa(xf_smem_u32(slot + (sp % XC_SLOTS) * (R * NW) + rank * NW + warp), lane), sw_, (double)n,
                                 xc_mapa(xf_smem_u32(nb + (sp & 1u)), lane));
            }
            __syncthreads();
            if (edgeN || edgeS || (ctrl && halo_bytes)) xf_mbar_wait(hb + 1, ph & 1u);
            ++ph;
            ++sp;
            if (pending && !verdict[sp & 1u]) {                    // (slot of sweep sp - 2) the previous sweep was the last one: undo this one
                if (ctrl) {
                    const unsigned pp = sp - 1u;                   // its partials are on their way: let them land
                    if (R > 1) xf_mbar_wait(nb + (pp & 1u), (pp >> 1) & 1u);
                } else {
                    #pragma unroll
                    for (int m = 0; m < K; ++m) psi[m] = bkT[m * nth + tid];
                }
                om = om_prev; om_first = om_first_prev;
                pending = false;
                break;
            }
            pending = true;
        }
        if (pending && ctrl) {                                     // the sweep budget of this launch is used up: test the last sweep now
            const unsigned pp = sp - 1u;
            if (R > 1) xf_mbar_wait(nb + (pp & 1u), (pp >> 1) & 1u);
            double ts = 0.0, tn = 0.0;
            {
                const double2 *sl = slot + (pp % XC_SLOTS) * (R * NW);
                for (int k = lane; k < R * NW; k += 32) { const double2 v = sl[k]; ts += v.x; tn += v.y; }
            }
            #pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
                ts += __shfl_xor_sync(0xffffffffu, ts, o);
                tn += __shfl_xor_sync(0xffffffffu, tn, o);
            }
            xd_decide(st_, ts, (i64)tn, a.tol, a.mxLoop, a.zero_exit);
        }
        // ---- psi back to the plan's buffer, state back ----
        if (act) {
            double *dst = a.Sbuf[cur] + (i64)b * a.slice + (i64)j * a.pitch + XM_PADL + i0;
            #pragma unroll
            for (int m = 0; m < K; ++m) if (m < nvalid) dst[m] = psi[m];
        }
        if (rank == 0 && post) {
            st_.omega = om;
            a.st[b] = st_;
            if (!st_.active) atomicSub(a.nactive, 1);
        }
    }
    cluster.sync();                                                // no CTA leaves while its shared memory may be addressed
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
struct ClusterPlan {
    bool built = false;
    XcArgs args{};
    int kind = 0, K = 0, R = 1;
    int threads = 0, grid = 0;
    size_t smem = 0;
};

static inline void cluster_plan_release(ClusterPlan &p) { p = ClusterPlan(); }

static const int XC_KS[] = {4, 6, 8, 10, 12, 16};    // (10: periodic grids of 50, 250, 350 ... columns)
#define XC_NK ((int)(sizeof(XC_KS) / sizeof(XC_KS[0])))

#define XC_DISPATCH(kind, K, CALL)                       \
    if ((kind) == 0) switch (K) {                        \
    case 4: CALL(0, 4); break;                           \
    case 6: CALL(0, 6); break;                           \
    case 8: CALL(0, 8); break;                           \
    case 10: CALL(0, 10); break;                         \
    case 12: CALL(0, 12); break;                         \
    default: CALL(0, 16); break;                         \
    } else switch (K) {                                  \
    case 4: CALL(1, 4); break;                           \
    case 6: CALL(1, 6); break;                           \
    case 8: CALL(1, 8); break;                           \
    case 10: CALL(1, 10); break;                         \
    case 12: CALL(1, 12); break;                         \
    default: CALL(1, 16); break;                         \
    }

template <int KIND, int K>
static cudaError_t xc_attr(size_t smem, int R, int *max_threads)
{
    cudaFuncAttributes fa;
    cudaError_t e = cudaFuncGetAttributes(&fa, xc_cluster_kernel<KIND, K>);
    if (e != cudaSuccess) return e;
    *max_threads = fa.maxThreadsPerBlock;
    if ((e = cudaFuncSetAttribute(xc_cluster_kernel<KIND, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess)
        return e;
    if (R > 8) e = cudaFuncSetAttribute(xc_cluster_kernel<KIND, K>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    return e;
}

template <int KIND, int K>
static cudaError_t xc_launch(const ClusterPlan &p, cudaStream_t stream, int *max_clusters)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(max_clusters ? p.R : p.grid));
    cfg.blockDim = dim3((unsigned)p.threads);
    cfg.dynamicSmemBytes = p.smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)p.R;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (max_clusters) return cudaOccupancyMaxActiveClusters(max_clusters, xc_cluster_kernel<KIND, K>, &cfg);
    return cudaLaunchKernelEx(&cfg, xc_cluster_kernel<KIND, K>, p.args);
}

// Fill the shape-dependent fields of a plan for cluster size R and run length K; false when the shape does not fit.
static inline bool xc_shape(ClusterPlan &cp, const XdGeom &g, int R, int K, int max_threads)
{
    const int ny = (int)g.ny, nx = (int)g.nx;
    if (R > 1 && ny / R < 2) return false;                  // every CTA holds at least two rows (y-extend stays inside a CTA)
    if (g.bcx == XD_BC_PERIODIC && nx % K != 0) return false;   // the last run ends at column nx-1 (its east neighbour is the ghost)
    XcArgs &a = cp.args;
    a.R = R;
    a.RPR = (nx + K - 1) / K;
    a.WPR = (a.RPR + 31) / 32;
    a.RPmax = (ny + R - 1) / R;
    a.TP = 2 + a.WPR * 32 * (K + 1);
    a.NW = a.RPmax * a.WPR;
    if ((a.NW + 1) * 32 > max_threads || a.NW + 1 > 32) return false;       // + the control warp
    cp.K = K; cp.R = R;
    cp.threads = (a.NW + 1) * 32;
    cp.smem = ((size_t)(a.RPmax + 2) * a.TP + (size_t)R * a.NW * 8 + (size_t)K * cp.threads) * sizeof(double) + 64;
    return cp.smem <= 200 * 1024;
}

// Build on top of a fused plan with row coefficients (p.rc): same operands, another way through them.
// (R, K) by a cost model fitted to measurements on B200 (scripts/bench_cluster_batch.py, scripts/prof_c1.py):
//   one sweep of a cluster ~ 2 x [550 + 70 (K/2) ceil(warps / 4)] cycles  (DSMEM flight + CTA barrier + stop-test share;
//                                                                          K/2 cells per thread and colour, the warps of
//                                                                          an SM sub-partition one after the other)
//   a batch takes ceil(batch / clusters resident at once) such sweeps;
//   the marching engine ~ 5 us + cells / 2.6e5 per us  (its fixed cost per pass, then 2.6e11 cell-updates/s).
// `force`: engine = 'cluster' was asked for -- take the best shape even where the marching engine is expected to win.
static inline int cluster_plan_build(ClusterPlan &cp, const FusedPlan &fp, int sm_count, const XdGeom &g, const XdCoef &q,
                                     i64 batch, bool force, std::string &why)
{
    (void)sm_count;
    cluster_plan_release(cp);
    if (!fp.built || !fp.rc) { why = "cluster engine needs coefficients constant along x"; return -1; }
    const XmArgs &fa = fp.args;
    const char *eR = getenv("XINV_CLUSTER_R"), *eK = getenv("XINV_CLUSTER_K");
    cudaError_t e = cudaSuccess;
    ClusterPlan best;
    double best_us = 1e300;
    for (int k = 0; k < XC_NK; ++k) {
        const int K = XC_KS[k];
        if (eK && atoi(eK) != K) continue;
        int mt = 0;
#define XC_ATTR(KD, K_) e = xc_attr<KD, K_>(200 * 1024, 16, &mt)
        XC_DISPATCH(fp.kind, K, XC_ATTR);
#undef XC_ATTR
        if (e != cudaSuccess) { why = std::string("cluster kernel attributes: ") + cudaGetErrorString(e); (void)cudaGetLastError(); return -1; }
        for (int R = 16; R >= 1; R >>= 1) {
            if (eR && atoi(eR) != R) continue;
            ClusterPlan c;
            c.kind = fp.kind;
            if (!xc_shape(c, g, R, K, mt)) continue;
            int maxcl = 0;
#define XC_OCC(KD, K_) e = xc_launch<KD, K_>(c, 0, &maxcl)
            XC_DISPATCH(c.kind, c.K, XC_OCC);
#undef XC_OCC
            if (e != cudaSuccess || maxcl < 1) { (void)cudaGetLastError(); continue; }
            const i64 waves = (batch + maxcl - 1) / maxcl;
            const double half = (R > 1 ? 550.0 : 250.0) + 70.0 * (K / 2) * ((c.args.NW + 3) / 4);
            const double us = (double)waves * 2.0 * half / 1900.0;
            if (us < best_us) {
                best_us = us; best = c;
                best.grid = (int)((batch < maxcl ? batch : maxcl) * R);
            }
        }
    }
    if (best_us >= 1e300) { why = "slice does not fit a thread-block cluster"; return -1; }
    const double march_us = 5.0 + (double)g.N * (double)batch / 2.6e5;
    if (!force && !(eR || eK) && best_us >= march_us) {
        why = "the marching engine is expected to be faster for this batch";
        return -1;
    }
    cp = best;
    XcArgs &a = cp.args;
    a.Sbuf[0] = fa.Sbuf[0]; a.Sbuf[1] = fa.Sbuf[1];
    a.Fd = (const double *)fp.bufFd; a.rows = (const double *)fp.bufRow;
    a.pitch = fa.pitch; a.slice = fa.slice; a.rpitch = (g.ny + 3) / 4 * 4;
    a.ny = (int)g.ny; a.nx = (int)g.nx; a.batch = (int)batch;
    a.bcy = g.bcy; a.bcx = g.bcx;
    a.cbFd = fa.cbFd; a.cbRow = fa.cbRow;
    a.ratioSqr = fa.ratioSqr; a.undef = q.undef; a.ratio = fa.ratio; a.delx = fa.delx; a.delxSqr = fa.delxSqr;
    cp.built = true;
    return 0;
}

static inline int cluster_sweep(ClusterPlan &p, cudaStream_t stream, XdSliceState *st, int *nactive, double tol, i64 mxLoop,
                                int zero_exit, int nsweeps, int accel, double rho2, int64_t *launches)
{
    XcArgs &a = p.args;
    a.st = st; a.nactive = nactive; a.tol = tol; a.mxLoop = mxLoop; a.zero_exit = zero_exit; a.nsweeps = nsweeps;
    a.accel = accel; a.rho2 = rho2;
    cudaError_t e = cudaSuccess;
#define XC_GO(KD, K_) e = xc_launch<KD, K_>(p, stream, nullptr)
    XC_DISPATCH(p.kind, p.K, XC_GO);
#undef XC_GO
    if (e != cudaSuccess) return -1;
    *launches += 1;
    return 0;
}
