// xinv_fused2d.cuh -- XINV_ENGINE_FUSED: one kernel launch per SOR iteration for
// the 2-D standard-form problem with B == 0 (invert_Poisson & friends; numbas.py:215-416).
//
// What one launch does, per tile of TY x TX cells of one slice:
//   1. one elected thread arms an mbarrier and issues four TMA box loads
//      (cp.async.bulk.tensor.3d) of the psi tile with a 2-cell halo and of the A, C
//      and F tiles into shared memory;
//   2. y-"extend" boundary rows are applied in shared memory (numbas.py:284-310);
//   3. red cells ((i+j) even) of the tile + 1-cell ring are updated in shared memory;
//   4. black cells of the tile are updated from the new red values, and the finished
//      tile is written with coalesced 16-byte stores to the *other* psi buffer
//      (ping-pong: a neighbouring tile still needs this tile's old values);
//   5. sum|psi| and the count of psi != undef of the tile are reduced with warp
//      shuffles into one partial per tile; the last tile of a slice to finish
//      (atomic ticket) adds the partials in index order and runs the reference's
//      loop control (numbas.py:401-414) for that slice.
// HBM traffic per iteration: psi read + psi write + A + C + F once each = 40 N bytes
// (the colour engine moves 2 x 40 N + 8 N).
//
// Layout in HBM: the engine works on its own copies with a padded pitch
// (nx + 4, rounded up to even): two ghost columns on either side of every row.
// For periodic-x they hold the wrap-around neighbours (the east edge tile writes
// its last two columns also into the west ghosts and vice versa, so they are
// always current); TMA boxes then never need wrap logic and box starts are
// 32-byte aligned.  Rows outside [0, ny) are zero-filled by TMA and never used.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <string>
#include "xinv_device.cuh"

// ----------------------------------------------------------------------------
// PTX helpers: mbarrier + TMA
// ----------------------------------------------------------------------------
__device__ __forceinline__ uint32_t xf_smem_u32(const void *p)
{
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void xf_mbar_init(uint64_t *bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(xf_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void xf_fence_barrier_init()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void xf_mbar_expect_tx(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(xf_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void xf_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "XF_WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra XF_DONE_%=;\n\t"
        "bra XF_WAIT_%=;\n\t"
        "XF_DONE_%=:\n\t"
        "}\n" ::"r"(xf_smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void xf_tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%3, %4, %5}], [%2];" ::"r"(xf_smem_u32(dst)),
        "l"((uint64_t)map), "r"(xf_smem_u32(bar)), "r"(x), "r"(y), "r"(z)
        : "memory");
}
__device__ __forceinline__ void xf_prefetch_tmap(const CUtensorMap *map)
{
    asm volatile("prefetch.tensormap [%0];" ::"l"((uint64_t)map) : "memory");
}

// ----------------------------------------------------------------------------
struct XfArgs {
    double *Sbuf[2];          // padded psi buffers [batch][ny][pitch]
    i64 pitch, ny, nx;
    i64 slice;                // ny * pitch
    int ntx, nty;
    int bcy, bcx;
    int cbA, cbC, cbF;        // 1: coefficient has a batch axis, 0: shared slice
    double delxSqr, ratioSqr, optArg, undef;
    int parity;               // which buffer holds the current psi
    XdSliceState *st;
    double *psum;
    i64 *pcnt;
    unsigned *ticket;
    int *nactive;
    double tol;
    i64 mxLoop;
    int zero_exit;
};

#define XF_THREADS 256

// SOR update of one cell from shared-memory tiles; identical operation order to
// xd_update_std2d<false> (numbas.py:351-369 with B == 0).
template <int W>
__device__ __forceinline__ double xf_cell(const double *__restrict__ sS, const double *__restrict__ sA,
                                          const double *__restrict__ sC, const double *__restrict__ sF, int r,
                                          int c, double Sc, double Sw, double Se, double delxSqr, double ratioSqr,
                                          double optArg, double undef)
{
    const int o = r * W + c;
    const double Fc = sF[o], An = sA[o + W], Ac = sA[o], Ce = sC[o + 1], Cc = sC[o];
    const bool cond = (Fc != undef) & (An != undef) & (Ac != undef) & (Ce != undef) & (Cc != undef);
    if (!cond) return Sc;
    const double Sn = sS[o + W], Ss = sS[o - W];
    const double t1 = (An * (Sn - Sc) - Ac * (Sc - Ss)) * ratioSqr;
    const double t4 = (Ce * (Se - Sc) - Cc * (Sc - Sw));
    double temp = (t1 + t4) - Fc * delxSqr;
    temp = temp * (optArg / ((An + Ac) * ratioSqr + (Ce + Cc)));
    return Sc + temp;
}

template <int TY, int W, int MINB>
__global__ void __launch_bounds__(XF_THREADS, MINB)
xf_std2d_kernel(const __grid_constant__ CUtensorMap mS0, const __grid_constant__ CUtensorMap mS1,
                const __grid_constant__ CUtensorMap mA, const __grid_constant__ CUtensorMap mC,
                const __grid_constant__ CUtensorMap mF, const XfArgs a)
{
    constexpr int TX = W - 4;
    constexpr int ROWS = TY + 4;
    constexpr int TILE = ROWS * W;
    constexpr int NWARP = XF_THREADS / 32;
    constexpr int PAIRS = W / 2;                 // column pairs per row
    constexpr int WPR = PAIRS / 32;              // warps needed to cover one row (1 or 2)
    static_assert(PAIRS % 32 == 0, "W must be 64 or 128");

    extern __shared__ __align__(1024) unsigned char xf_smem[];
    double *sS = reinterpret_cast<double *>(xf_smem);
    double *sA = sS + TILE;
    double *sC = sA + TILE;
    double *sF = sC + TILE;
    uint64_t *bar = reinterpret_cast<uint64_t *>(sF + TILE);
    __shared__ double red_sum[32];
    __shared__ i64 red_cnt[32];
    __shared__ int is_last;

    const int b = blockIdx.y;
    if (!a.st[b].active) return;
    const int tile = blockIdx.x;
    const int tyi = tile / a.ntx, txi = tile - tyi * a.ntx;
    const int x0 = txi * TX, y0 = tyi * TY;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int ny = (int)a.ny, nx = (int)a.nx;
    const bool periodic = (a.bcx == XD_BC_PERIODIC);

    if (tid == 0) {
        xf_mbar_init(bar, 1);
        xf_fence_barrier_init();
    }
    __syncthreads();
    if (tid == 0) {
        xf_mbar_expect_tx(bar, 4u * TILE * (uint32_t)sizeof(double));
        // padded x of global column (x0 - 2) is x0; rows start at y0 - 2 (OOB rows -> 0)
        xf_tma_load_3d(sS, a.parity ? &mS1 : &mS0, bar, x0, y0 - 2, b);
        xf_tma_load_3d(sA, &mA, bar, x0, y0 - 2, b * a.cbA);
        xf_tma_load_3d(sC, &mC, bar, x0, y0 - 2, b * a.cbC);
        xf_tma_load_3d(sF, &mF, bar, x0, y0 - 2, b * a.cbF);
    }
    xf_mbar_wait(bar, 0);

    // ---- y-"extend" rows (numbas.py:284-310), on the pre-sweep values ----------
    if (a.bcy == XD_BC_EXTEND) {
        // smem row of global row 0 is 2 - y0; of global row ny-1 is ny + 1 - y0
        const int rb = 2 - y0, rt = ny + 1 - y0;
        for (int c = tid; c < 2 * W; c += XF_THREADS) {
            const bool top = (c >= W);
            const int cc = top ? c - W : c;
            const int rdst = top ? rt : rb;
            const int rsrc = top ? rt - 1 : rb + 1;
            if (rdst < 0 || rdst >= ROWS || rsrc < 0 || rsrc >= ROWS) continue;
            const int gi = x0 - 2 + cc;
            int cs = cc;
            if (!periodic) {
                if (gi < 0 || gi > nx - 1) continue;
                if (gi == 0) cs = cc + 1;                 // corners copy the diagonal neighbour
                else if (gi == nx - 1) cs = cc - 1;
            } else if (gi < -2 || gi > nx + 1) continue;
            if (cs < 0 || cs >= W) continue;
            const double v = sS[rsrc * W + cs];
            if (v != a.undef) sS[rdst * W + cc] = v;
        }
        __syncthreads();
    }

    // ---- red cells ((i + j) even) on the tile + 1-cell ring ---------------------
    const int ilo = periodic ? -1 : 1, ihi = periodic ? nx : nx - 2;   // columns that get updated
    for (int rr = warp / WPR; rr < ROWS - 2; rr += NWARP / WPR) {
        const int r = rr + 1;
        const int j = y0 - 2 + r;
        if (j < 1 || j > ny - 2) continue;
        const int pp = lane + 32 * (warp % WPR);
        const int c = 2 * pp + (j & 1);            // x0 is even: parity of global i == parity of c
        const int gi = x0 - 2 + c;
        if (c >= 1 && c <= W - 2 && gi >= ilo && gi <= ihi) {
            const int o = r * W + c;
            sS[o] = xf_cell<W>(sS, sA, sC, sF, r, c, sS[o], sS[o - 1], sS[o + 1], a.delxSqr, a.ratioSqr, a.optArg,
                               a.undef);
        }
    }
    __syncthreads();

    // ---- black cells of the tile; write the finished tile; norm partial -------
    double *out = a.Sbuf[a.parity ^ 1] + (i64)b * a.slice;
    const int i0 = periodic ? 0 : 1, i1 = periodic ? nx : nx - 1;
    double nsum = 0.0;
    i64 ncnt = 0;
    for (int rr = warp / WPR; rr < TY; rr += NWARP / WPR) {
        const int r = rr + 2;
        const int j = y0 + rr;
        if (j > ny - 1) break;
        const int pp = lane + 32 * (warp % WPR);
        const int c0 = 2 * pp;
        if (c0 < 2 || c0 > W - 4) continue;
        const int gi = x0 - 2 + c0;               // even global column of this pair
        if (gi >= nx) continue;
        const int o = r * W + c0;
        double2 v = *reinterpret_cast<const double2 *>(sS + o);
        if (j >= 1 && j <= ny - 2) {
            if (j & 1) {                           // odd row: black is the even column
                if (gi >= i0 && gi < i1)
                    v.x = xf_cell<W>(sS, sA, sC, sF, r, c0, v.x, sS[o - 1], v.y, a.delxSqr, a.ratioSqr, a.optArg,
                                     a.undef);
            } else {                               // even row: black is the odd column
                if (gi + 1 >= i0 && gi + 1 < i1)
                    v.y = xf_cell<W>(sS, sA, sC, sF, r, c0 + 1, v.y, v.x, sS[o + 2], a.delxSqr, a.ratioSqr,
                                     a.optArg, a.undef);
            }
        }
        double *dst = out + (i64)j * a.pitch + 2 + gi;
        if (gi + 1 < nx) {
            *reinterpret_cast<double2 *>(dst) = v;
            if (periodic) {                        // keep the ghost columns current
                if (gi == 0) *reinterpret_cast<double2 *>(dst + nx) = v;
                if (gi == nx - 2) *reinterpret_cast<double2 *>(dst - nx) = v;
            }
            if (v.x != a.undef) { nsum += fabs(v.x); ncnt += 1; }
            if (v.y != a.undef) { nsum += fabs(v.y); ncnt += 1; }
        } else {                                   // odd nx (non-periodic only): last column alone
            dst[0] = v.x;
            if (v.x != a.undef) { nsum += fabs(v.x); ncnt += 1; }
        }
    }

    // ---- norm partial of this tile, then ticket; last tile runs the loop control ----
    xd_block_reduce(nsum, ncnt, red_sum, red_cnt);
    const int ntiles = a.ntx * a.nty;
    if (tid == 0) {
        a.psum[(i64)b * ntiles + tile] = nsum;
        a.pcnt[(i64)b * ntiles + tile] = ncnt;
        __threadfence();
        const unsigned t = atomicAdd(&a.ticket[b], 1u);
        is_last = (t == (unsigned)ntiles - 1u);
    }
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    {
        const volatile double *vs = a.psum + (i64)b * ntiles;
        const volatile i64 *vc = a.pcnt + (i64)b * ntiles;
        double s = 0.0;
        i64 cn = 0;
        // fixed assignment of partials to threads and fixed reduction tree: the
        // result does not depend on which tile happened to finish last
        for (int p = tid; p < ntiles; p += XF_THREADS) { s += vs[p]; cn += vc[p]; }
        __syncthreads();
        xd_block_reduce(s, cn, red_sum, red_cnt);
        if (tid == 0) {
            XdSliceState s_ = a.st[b];
            xd_decide(s_, s, cn, a.tol, a.mxLoop, a.zero_exit);
            a.st[b] = s_;
            a.ticket[b] = 0u;
            if (!s_.active) atomicSub(a.nactive, 1);
        }
    }
}

// ----------------------------------------------------------------------------
// dense <-> padded layout
// ----------------------------------------------------------------------------
__global__ void xf_pack_kernel(double *__restrict__ dst, const double *__restrict__ src, i64 ny, i64 nx,
                               i64 pitch, i64 src_bstride, int periodic)
{
    const i64 j = blockIdx.y;
    const int b = blockIdx.z;
    const i64 pc = (i64)blockIdx.x * blockDim.x + threadIdx.x;     // padded column
    if (pc >= pitch) return;
    const double *s = src + (i64)b * src_bstride + j * nx;
    i64 i = pc - 2;
    double v = 0.0;
    if (i >= 0 && i < nx) v = s[i];
    else if (periodic && i >= -2 && i < nx + 2) v = s[(i + nx) % nx];
    dst[((i64)b * ny + j) * pitch + pc] = v;
}

__global__ void xf_unpack_kernel(double *__restrict__ dst, const double *__restrict__ buf0,
                                 const double *__restrict__ buf1, i64 ny, i64 nx, i64 pitch,
                                 const XdSliceState *__restrict__ st)
{
    const i64 j = blockIdx.y;
    const int b = blockIdx.z;
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nx) return;
    const double *src = (st[b].sweeps_done & 1) ? buf1 : buf0;
    dst[((i64)b * ny + j) * nx + i] = src[((i64)b * ny + j) * pitch + 2 + i];
}

// ----------------------------------------------------------------------------
// host side
// ----------------------------------------------------------------------------
typedef CUresult (*xf_encode_fn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                 const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                 CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

struct FusedPlan {
    bool built = false;
    int nblk_partials = 0;
    int variant = 0;               // 0: TY=28,W=64   1: TY=12,W=128
    void *bufS[2] = {nullptr, nullptr};
    void *bufA = nullptr, *bufC = nullptr, *bufF = nullptr;
    CUtensorMap mS[2], mA, mC, mF;
    XfArgs args{};
    i64 batch = 0;
    i64 sweeps = 0;
    size_t smem = 0;
};

static inline void fused_plan_release(FusedPlan &p)
{
    if (p.bufS[0]) cudaFree(p.bufS[0]);
    if (p.bufS[1]) cudaFree(p.bufS[1]);
    if (p.bufA) cudaFree(p.bufA);
    if (p.bufC) cudaFree(p.bufC);
    if (p.bufF) cudaFree(p.bufF);
    p = FusedPlan();
}

static inline bool fused_plan_supported(int kind, bool hasB, const XdGeom &g, std::string &why)
{
    if (kind != 0 /* XD_STD2D */) { why = "fused engine covers the 2-D standard form only"; return false; }
    if (hasB) { why = "fused engine needs B == 0 (5-point stencil)"; return false; }
    if (g.wrapfix) { why = "periodic-x with odd nx needs the wrap-fix colours"; return false; }
    if (g.ny < 3 || g.nx < 4) { why = "grid too small"; return false; }
    return true;
}

static xf_encode_fn xf_get_encode()
{
    static xf_encode_fn fn = nullptr;
    if (fn) return fn;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = (xf_encode_fn)p;
    return fn;
}

static int xf_make_map(CUtensorMap *m, void *base, i64 pitch, i64 ny, i64 nb, int W, int ROWS, std::string &why)
{
    xf_encode_fn enc = xf_get_encode();
    if (!enc) { why = "cuTensorMapEncodeTiled not available from the driver"; return -1; }
    cuuint64_t dims[3] = {(cuuint64_t)pitch, (cuuint64_t)ny, (cuuint64_t)nb};
    cuuint64_t strides[2] = {(cuuint64_t)pitch * 8, (cuuint64_t)pitch * (cuuint64_t)ny * 8};
    cuuint32_t box[3] = {(cuuint32_t)W, (cuuint32_t)ROWS, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, base, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { why = "cuTensorMapEncodeTiled failed (" + std::to_string((int)r) + ")"; return -1; }
    return 0;
}

template <int TY, int W, int MINB>
static int xf_launch(FusedPlan &p, cudaStream_t stream)
{
    dim3 grid((unsigned)(p.args.ntx * p.args.nty), (unsigned)p.batch);
    xf_std2d_kernel<TY, W, MINB><<<grid, XF_THREADS, p.smem, stream>>>(p.mS[0], p.mS[1], p.mA, p.mC, p.mF, p.args);
    return 0;
}

#define XF_V0_TY 28
#define XF_V0_W 64
#define XF_V1_TY 12
#define XF_V1_W 128

static inline int fused_plan_build(FusedPlan &p, int sm_count, int kind, const XdGeom &g, const XdCoef &q, i64 batch,
                                   double *dS, double * /*unused*/, cudaStream_t stream, std::string &why)
{
    (void)sm_count; (void)kind;
    fused_plan_release(p);
    const i64 ny = g.ny, nx = g.nx;
    i64 pitch = nx + 4;
    if (pitch & 1) pitch += 1;
    const char *env = getenv("XINV_FUSED_VARIANT");
    p.variant = env ? atoi(env) : 0;
    const int TY = p.variant == 1 ? XF_V1_TY : XF_V0_TY;
    const int W = p.variant == 1 ? XF_V1_W : XF_V0_W;
    const int TX = W - 4, ROWS = TY + 4;
    const int periodic = (g.bcx == XD_BC_PERIODIC);
    const size_t slice_bytes = (size_t)ny * pitch * sizeof(double);
    const int cb[3] = {q.cs[0] != 0, q.cs[2] != 0, q.cs[3] != 0};
    cudaError_t e;
#define XF_ALLOC(ptr, bytes)                                                        \
    if ((e = cudaMalloc(&(ptr), (bytes))) != cudaSuccess) {                         \
        why = std::string("cudaMalloc: ") + cudaGetErrorString(e);                  \
        fused_plan_release(p);                                                      \
        return -1;                                                                  \
    }
    XF_ALLOC(p.bufS[0], slice_bytes * batch);
    XF_ALLOC(p.bufS[1], slice_bytes * batch);
    XF_ALLOC(p.bufA, slice_bytes * (cb[0] ? batch : 1));
    XF_ALLOC(p.bufC, slice_bytes * (cb[1] ? batch : 1));
    XF_ALLOC(p.bufF, slice_bytes * (cb[2] ? batch : 1));
#undef XF_ALLOC
    dim3 blk(128);
    auto pack = [&](void *dst, const double *src, i64 bstride, i64 nb) {
        dim3 grid((unsigned)((pitch + 127) / 128), (unsigned)ny, (unsigned)nb);
        xf_pack_kernel<<<grid, blk, 0, stream>>>((double *)dst, src, ny, nx, pitch, bstride, periodic);
    };
    pack(p.bufS[0], dS, g.N, batch);
    pack(p.bufS[1], dS, g.N, batch);       // boundary rows/cols of both buffers start identical
    pack(p.bufA, q.c[0], q.cs[0], cb[0] ? batch : 1);
    pack(p.bufC, q.c[2], q.cs[2], cb[1] ? batch : 1);
    pack(p.bufF, q.c[3], q.cs[3], cb[2] ? batch : 1);
    if ((e = cudaGetLastError()) != cudaSuccess) {
        why = std::string("pack kernels: ") + cudaGetErrorString(e);
        fused_plan_release(p);
        return -1;
    }
    if (xf_make_map(&p.mS[0], p.bufS[0], pitch, ny, batch, W, ROWS, why) ||
        xf_make_map(&p.mS[1], p.bufS[1], pitch, ny, batch, W, ROWS, why) ||
        xf_make_map(&p.mA, p.bufA, pitch, ny, cb[0] ? batch : 1, W, ROWS, why) ||
        xf_make_map(&p.mC, p.bufC, pitch, ny, cb[1] ? batch : 1, W, ROWS, why) ||
        xf_make_map(&p.mF, p.bufF, pitch, ny, cb[2] ? batch : 1, W, ROWS, why)) {
        fused_plan_release(p);
        return -1;
    }
    XfArgs &a = p.args;
    a.Sbuf[0] = (double *)p.bufS[0];
    a.Sbuf[1] = (double *)p.bufS[1];
    a.pitch = pitch; a.ny = ny; a.nx = nx; a.slice = ny * pitch;
    a.ntx = (int)((nx + TX - 1) / TX);
    a.nty = (int)((ny + TY - 1) / TY);
    a.bcy = g.bcy; a.bcx = g.bcx;
    a.cbA = cb[0]; a.cbC = cb[1]; a.cbF = cb[2];
    a.delxSqr = q.p[0]; a.ratioSqr = q.p[2]; a.optArg = q.optArg; a.undef = q.undef;
    a.parity = 0;
    p.batch = batch;
    p.sweeps = 0;
    p.nblk_partials = a.ntx * a.nty;
    p.smem = (size_t)4 * ROWS * W * sizeof(double) + 64;
    if (p.variant == 1)
        e = cudaFuncSetAttribute(xf_std2d_kernel<XF_V1_TY, XF_V1_W, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    else
        e = cudaFuncSetAttribute(xf_std2d_kernel<XF_V0_TY, XF_V0_W, 3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)p.smem);
    if (e != cudaSuccess) {
        why = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        fused_plan_release(p);
        return -1;
    }
    p.built = true;
    return 0;
}

static inline int fused_sweep(FusedPlan &p, cudaStream_t stream, XdSliceState *st, double *psum, i64 *pcnt,
                              unsigned *ticket, int *nactive, double tol, i64 mxLoop, int zero_exit, int64_t *launches)
{
    XfArgs &a = p.args;
    a.st = st; a.psum = psum; a.pcnt = pcnt; a.ticket = ticket; a.nactive = nactive;
    a.tol = tol; a.mxLoop = mxLoop; a.zero_exit = zero_exit;
    a.parity = (int)(p.sweeps & 1);
    if (p.variant == 1) xf_launch<XF_V1_TY, XF_V1_W, 3>(p, stream);
    else                xf_launch<XF_V0_TY, XF_V0_W, 3>(p, stream);
    p.sweeps += 1;
    *launches += 1;
    return 0;
}

// copy every slice's final psi (whichever buffer holds it) back to the dense array
static inline int fused_unpack(FusedPlan &p, double *dS, const XdSliceState *st, cudaStream_t stream)
{
    const XfArgs &a = p.args;
    dim3 grid((unsigned)((a.nx + 127) / 128), (unsigned)a.ny, (unsigned)p.batch);
    xf_unpack_kernel<<<grid, 128, 0, stream>>>(dS, a.Sbuf[0], a.Sbuf[1], a.ny, a.nx, a.pitch, st);
    return 0;
}
