// xinv_api.cu -- host side of libxinv_b200.so: context, staging, the sweep
// driver (chunks of sweeps between host polls of the device-side active count)
// and the extern "C" entry points declared in include/xinv.h.
//
// The reference's per-slice Python loop (core.py:129-153) and the iteration
// loop of numbas.py:282-414 are replaced by: all slices resident in HBM, every
// sweep a batched kernel launch over all still-active slices, per-slice loop
// state (normPrev, loop, flags) updated on the device by the norm/decide
// kernel, and the host only polling "how many slices are still active".
#include <cuda_runtime.h>
#include <cuda.h>
#include <dlfcn.h>
#include <float.h>
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/xinv.h"
#include "xinv_device.cuh"
#include "xinv_colour_engine.cuh"
#include "xinv_march2d.cuh"
#include "xinv_march3d.cuh"
#include "xinv_lex_engine.cuh"
#include "xinv_flow.cuh"
#include "xinv_resident.cuh"
#include "xinv_cluster2d.cuh"

// ---------------------------------------------------------------------------
// errors
// ---------------------------------------------------------------------------
static thread_local char g_err[512] = "";

static int set_err(int code, const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
    return code;
}

#define CK(call)                                                                   \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess)                                                     \
            return set_err(XINV_E_CUDA, "%s failed: %s (%s:%d)", #call,            \
                           cudaGetErrorString(e_), __FILE__, __LINE__);            \
    } while (0)

// ---------------------------------------------------------------------------
// context
// ---------------------------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t bytes = 0;
};

struct Problem {
    bool open = false;
    int kind = 0;
    bool hasB = false;
    int zero_exit = 0;
    i64 batch = 0;
    XdGeom g{};
    XdCoef q{};
    double *dS = nullptr;        // device S [batch][N]
    double *dS2 = nullptr;       // ping-pong partner (fused engine)
    double *userS = nullptr;     // caller's S (host or device)
    double *userFlags = nullptr; // host flags [batch][3]
    int mem_space = 0;
    int ordering = 0;
    int engine = 0;
    int check_every = 0;
    int profile = 0;
    double tol = 0;
    i64 mxLoop = 0;
    i64 sweeps_launched = 0;
    int nblk_norm = 0;
    int h_nactive = 0;
    FusedPlan fused{};
    Fused3Plan fused3{};         // 3-D standard form (xinv_march3d.cuh)
    ResidentPlan resident{};     // small 2-D slices (xinv_resident.cuh)
    ClusterPlan cluster{};       // small / medium 2-D slices with row coefficients, on top of `fused` (xinv_cluster2d.cuh)
    bool front = false;          // xinv_std2d_rows: S is output only, de-masked on the device
    bool maskfront = false;      // xinv_std2d_front: dense coefficients, the user's forcing masked / the result de-masked on the device
    const double *dFmask = nullptr;   // ... its masked forcing (device)
    double out_undef = 0.0;
    int io_f32 = 0;              // XINV_IO_F32_* (front ends, host pointers)
    int accel = 0;               // XINV_ACCEL_*
    double rho2 = 0.0;           // Chebyshev: squared Jacobi spectral radius implied by optArg
    double optArg0 = 0.0;        // the caller's optArg (q.optArg is overwritten per half sweep on the colour engine)
    double omega = 1.0;          // colour engine: factor of the next half sweep (all active slices are at the same sweep)
    bool omega_first = true;
};

struct xinv_ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    bool own_stream = false;
    int sm_count = 148;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    cudaEvent_t tm0 = nullptr, tm1 = nullptr;   // xinv_timer_*
    std::vector<cudaEvent_t> prof_ev;   // pairs (begin, end) around the dominant kernels
    size_t prof_used = 0;
    // workspace (grown on demand, reused across calls)
    DevBuf stage[14];            // staged S, S2 and up to 12 coefficient arrays
    DevBuf f32buf;               // float32 side of the front ends' I/O (xinv_opts.io_f32)
    DevBuf state, psum, pcnt, ticket, nactive, flags_in;
    XmWork xm_work;              // padded operand copies of the fused engine
    int *h_nactive_pinned = nullptr;
    Problem pb;
    xinv_stats stats{};
    // NCCL (loaded lazily with dlopen)
    void *nccl_lib = nullptr;
    void *nccl_comm = nullptr;
    DevBuf nccl_buf;
    int nccl_rank = 0, nccl_world = 1;
};

static int ensure(DevBuf &b, size_t bytes)
{
    if (b.bytes >= bytes && b.p) return XINV_OK;
    if (b.p) { cudaFree(b.p); b.p = nullptr; b.bytes = 0; }
    if (bytes == 0) bytes = 16;
    cudaError_t e = cudaMalloc(&b.p, bytes);
    if (e != cudaSuccess)
        return set_err(XINV_E_NOMEM, "cudaMalloc(%zu) failed: %s", bytes, cudaGetErrorString(e));
    b.bytes = bytes;
    return XINV_OK;
}

static void release(DevBuf &b)
{
    if (b.p) cudaFree(b.p);
    b.p = nullptr;
    b.bytes = 0;
}

extern "C" const char *xinv_last_error(void) { return g_err; }
extern "C" int xinv_version(void) { return XINV_VERSION; }

extern "C" int xinv_device_count(int *out)
{
    if (!out) return set_err(XINV_E_ARG, "out is NULL");
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *out = 0; return set_err(XINV_E_CUDA, "cudaGetDeviceCount: %s", cudaGetErrorString(e)); }
    *out = n;
    return XINV_OK;
}

static int create_common(xinv_ctx **out, int device, void *stream, bool own)
{
    if (!out) return set_err(XINV_E_ARG, "out is NULL");
    *out = nullptr;
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return set_err(XINV_E_ARG, "device %d out of range (0..%d)", device, n - 1);
    CK(cudaSetDevice(device));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, device));
    if (prop.major < 10)
        return set_err(XINV_E_UNSUPPORTED, "device %d is sm_%d%d; libxinv_b200 is built for sm_100a only",
                       device, prop.major, prop.minor);
    xinv_ctx *c = new xinv_ctx();
    c->device = device;
    c->sm_count = prop.multiProcessorCount;
    if (own) {
        CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
        c->own_stream = true;
    } else {
        c->stream = (cudaStream_t)stream;
    }
    CK(cudaEventCreate(&c->ev0));
    CK(cudaEventCreate(&c->ev1));
    CK(cudaEventCreate(&c->tm0));
    CK(cudaEventCreate(&c->tm1));
    CK(cudaHostAlloc((void **)&c->h_nactive_pinned, 64, cudaHostAllocDefault));
    *out = c;
    return XINV_OK;
}

extern "C" int xinv_create(xinv_ctx **out, int device) { return create_common(out, device, nullptr, true); }
extern "C" int xinv_create_on_stream(xinv_ctx **out, int device, void *s) { return create_common(out, device, s, false); }

extern "C" int xinv_nccl_finalize(xinv_ctx *ctx);

extern "C" void xinv_destroy(xinv_ctx *c)
{
    if (!c) return;
    cudaSetDevice(c->device);
    cudaStreamSynchronize(c->stream);
    xinv_nccl_finalize(c);
    for (auto &b : c->stage) release(b);
    release(c->f32buf);
    release(c->state); release(c->psum); release(c->pcnt); release(c->ticket); release(c->nactive); release(c->flags_in);
    release(c->nccl_buf);
    fused_plan_release(c->pb.fused);
    fused3_plan_release(c->pb.fused3);
    xm_work_release(c->xm_work);
    if (c->h_nactive_pinned) cudaFreeHost(c->h_nactive_pinned);
    for (auto e : c->prof_ev) cudaEventDestroy(e);
    if (c->ev0) cudaEventDestroy(c->ev0);
    if (c->ev1) cudaEventDestroy(c->ev1);
    if (c->tm0) cudaEventDestroy(c->tm0);
    if (c->tm1) cudaEventDestroy(c->tm1);
    if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
    delete c;
}

extern "C" int xinv_get_stats(const xinv_ctx *c, xinv_stats *out)
{
    if (!c || !out) return set_err(XINV_E_ARG, "NULL argument");
    *out = c->stats;
    return XINV_OK;
}

extern "C" int xinv_synchronize(xinv_ctx *c)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaStreamSynchronize(c->stream));
    return XINV_OK;
}

extern "C" int xinv_timer_start(xinv_ctx *c)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->tm0, c->stream));
    return XINV_OK;
}

extern "C" int xinv_timer_stop(xinv_ctx *c, double *ms_out)
{
    if (!c || !ms_out) return set_err(XINV_E_ARG, "NULL argument");
    CK(cudaSetDevice(c->device));
    CK(cudaEventRecord(c->tm1, c->stream));
    CK(cudaEventSynchronize(c->tm1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->tm0, c->tm1));
    *ms_out = ms;
    return XINV_OK;
}

extern "C" int xinv_host_alloc(void **out, int64_t bytes)
{
    if (!out || bytes < 0) return set_err(XINV_E_ARG, "bad argument");
    CK(cudaHostAlloc(out, (size_t)(bytes ? bytes : 16), cudaHostAllocDefault));
    return XINV_OK;
}
extern "C" int xinv_host_free(void *p)
{
    if (p) CK(cudaFreeHost(p));
    return XINV_OK;
}
extern "C" int xinv_host_is_pinned(const void *p, int *out)
{
    if (!out) return set_err(XINV_E_ARG, "out is NULL");
    cudaPointerAttributes at;
    cudaError_t e = cudaPointerGetAttributes(&at, p);
    if (e != cudaSuccess) { (void)cudaGetLastError(); *out = 0; return XINV_OK; }
    *out = (at.type == cudaMemoryTypeHost) ? 1 : 0;
    return XINV_OK;
}
extern "C" int xinv_dev_alloc(xinv_ctx *c, void **out, int64_t bytes)
{
    if (!c || !out || bytes < 0) return set_err(XINV_E_ARG, "bad argument");
    CK(cudaSetDevice(c->device));
    CK(cudaMalloc(out, (size_t)(bytes ? bytes : 16)));
    return XINV_OK;
}
extern "C" int xinv_dev_free(xinv_ctx *c, void *p)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    if (p) CK(cudaFree(p));
    return XINV_OK;
}
extern "C" int xinv_memcpy_h2d(xinv_ctx *c, void *dst, const void *src, int64_t bytes)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyHostToDevice, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return XINV_OK;
}
extern "C" int xinv_memcpy_d2h(xinv_ctx *c, void *dst, const void *src, int64_t bytes)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    CK(cudaSetDevice(c->device));
    CK(cudaMemcpyAsync(dst, src, (size_t)bytes, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return XINV_OK;
}

// Dense front end (xinv_std2d_front: invert_Eliassen): what apps.__mask_FS and the de-masking of apps.__template do on
// full-size host arrays (apps.py:2112-2159, :1386-1392), on the device -- land (F == user_undef, any NaN when that is NaN)
// becomes the internal undef in the forcing, the initial guess is zero; afterwards land becomes out_undef in the result.
// flag[1] |= 1 when an unmasked forcing value is not finite (the caller falls back to the host-built path then).
__global__ void xd_front_mask_kernel(double *__restrict__ F, double *__restrict__ S, i64 n, double user_undef, double undef,
                                     int *__restrict__ flag)
{
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n) return;
    const double f = F[p];
    const bool land = ((user_undef != user_undef) ? (f != f) : (f == user_undef)) | (f == undef);
    if (land) F[p] = undef;
    else if (!isfinite(f)) flag[1] = 1;
    S[p] = 0.0;
}
__global__ void xd_front_demask_kernel(double *__restrict__ S, const double *__restrict__ F, i64 n, double undef, double out_undef)
{
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n && F[p] == undef) S[p] = out_undef;
}

// float32 I/O of the front ends (xinv_opts.io_f32): widen after the H2D copy, narrow before the D2H copy
__global__ void xd_widen_kernel(double *__restrict__ dst, const float *__restrict__ src, i64 n)
{
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = (double)src[p];
}
__global__ void xd_narrow_kernel(float *__restrict__ dst, const double *__restrict__ src, i64 n)
{
    const i64 p = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    if (p < n) dst[p] = __double2float_rn(src[p]);
}

// ---------------------------------------------------------------------------
// begin: validate, stage, initialise per-slice state
// ---------------------------------------------------------------------------
__global__ void xd_init_state_kernel(XdSliceState *st, const double *flags_in, int batch, int *nactive, int nit0)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b == 0) { nactive[0] = batch; nactive[1] = 0; }
    if (b >= batch) return;
    XdSliceState s;
    s.normPrev = DBL_MAX;                  // numbas.py:280 np.finfo(np.float64).max
    s.flags[0] = flags_in[3 * b + 0];
    s.flags[1] = flags_in[3 * b + 1];
    s.flags[2] = flags_in[3 * b + 2];
    s.loop = 0;
    s.active = 1;
    s.sweeps_done = 0;
    s.cur = 0;
    s.nit = nit0;
    s.redo = 0;
    s.pad_ = 0;
    s.omega = 1.0;
    st[b] = s;
}

static int valid_bc(int bc) { return bc == XINV_BC_FIXED || bc == XINV_BC_EXTEND || bc == XINV_BC_PERIODIC; }

struct BeginArgs {
    int kind;
    double *S;
    const double *coef[12];
    int ncoef;
    int b_index;            // index of the optional B array in coef[] (or -1)
    i64 batch, nz, ny, nx;
    int bcy, bcx;
    double p[8];
    double optArg, undef;
    double *flags;
    i64 mxLoop;
    double tol;
    const xinv_opts *opts;
    // xinv_std2d_rows (front end): coef[0] = A rows [ny], coef[2] = C rows [ny], coef[3] = user forcing
    bool front = false;
    bool maskfront = false;  // xinv_std2d_front: dense A, B, C; coef[3] = the user's forcing (S is output only)
    const double *f_scale = nullptr;
    double user_undef = 0.0, out_undef = 0.0;
    // xinv_gen2d_rows: coef[0] = rows [5][ny], coef[6] = user forcing
    int g_mode = 0;
    double g_p1 = 1.0, g_p2 = 1.0;
    // xinv_std3d_rows: coef[0] = rows [4][ny], coef[1] = N2 (n2_count elements, strides n2_strides), coef[3] = user forcing
    i64 n2_strides[4] = {0, 0, 0, 0};
    i64 n2_count = 0;
};

static int problem_begin(xinv_ctx *c, const BeginArgs &a)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    if (c->pb.open) return set_err(XINV_E_STATE, "a problem is already open on this ctx (call xinv_end)");
    if (!a.S || !a.flags) return set_err(XINV_E_ARG, "S and flags must not be NULL");
    if (a.batch < 0 || a.nz < 1 || a.ny < 1 || a.nx < 1) return set_err(XINV_E_ARG, "bad sizes batch=%lld nz=%lld ny=%lld nx=%lld", a.batch, a.nz, a.ny, a.nx);
    if (a.batch > 65535) return set_err(XINV_E_ARG, "batch %lld > 65535: split the call", a.batch);
    if (a.nx > 0x3fffffff || a.ny > 0x3fffffff) return set_err(XINV_E_ARG, "grid too large");
    if (!valid_bc(a.bcy) || !valid_bc(a.bcx)) return set_err(XINV_E_ARG, "bad boundary condition code");
    if (a.bcy == XINV_BC_EXTEND && a.ny < 2) return set_err(XINV_E_ARG, "'extend' needs ny >= 2");
    if (a.mxLoop < 0) return set_err(XINV_E_ARG, "mxLoop < 0");
    for (int m = 0; m < a.ncoef; ++m)
        if (!a.coef[m] && m != a.b_index) return set_err(XINV_E_ARG, "coefficient array %d is NULL", m);

    xinv_opts o;
    memset(&o, 0, sizeof o);
    for (int m = 0; m < 8; ++m) o.coef_stride[m] = -1;
    if (a.opts) {
        if (a.opts->struct_size != (int32_t)sizeof(xinv_opts))
            return set_err(XINV_E_ARG, "opts.struct_size %d != %zu", a.opts->struct_size, sizeof(xinv_opts));
        o = *a.opts;
    }
    if (o.ordering != XINV_ORDER_COLOUR && o.ordering != XINV_ORDER_LEX) return set_err(XINV_E_ARG, "bad ordering");
    if (o.mem_space != XINV_MEM_HOST && o.mem_space != XINV_MEM_DEVICE) return set_err(XINV_E_ARG, "bad mem_space");
    if (o.accel != XINV_ACCEL_NONE && o.accel != XINV_ACCEL_CHEBYSHEV) return set_err(XINV_E_ARG, "bad accel");
    if (o.io_f32 & ~(XINV_IO_F32_IN | XINV_IO_F32_OUT)) return set_err(XINV_E_ARG, "bad io_f32");
    if (o.io_f32 && (!a.front || o.mem_space != XINV_MEM_HOST))
        return set_err(XINV_E_UNSUPPORTED, "io_f32 is offered by the device front ends (xinv_*_rows) with host pointers");
    if (o.accel && o.ordering != XINV_ORDER_COLOUR) return set_err(XINV_E_UNSUPPORTED, "accel needs the colour ordering");
    if (o.accel && (o.engine == XINV_ENGINE_FUSED)) return set_err(XINV_E_UNSUPPORTED, "accel runs on the cluster, resident and colour engines");

    CK(cudaSetDevice(c->device));
    Problem &pb = c->pb;
    fused_plan_release(pb.fused);
    fused3_plan_release(pb.fused3);
    resident_plan_release(pb.resident);
    cluster_plan_release(pb.cluster);
    pb = Problem();
    pb.kind = a.kind;
    pb.batch = a.batch;
    pb.hasB = (a.b_index >= 0 && a.coef[a.b_index] != nullptr);
    pb.zero_exit = (a.kind == XD_STD2D || a.kind == XD_STD2DT || a.kind == XD_STD1D);   // kernels with the norm == 0 exit (numbas.py:410, :621, :735)
    pb.userS = a.S;
    pb.userFlags = a.flags;
    pb.mem_space = o.mem_space;
    pb.ordering = o.ordering;
    pb.check_every = o.check_every;
    pb.profile = o.profile;
    pb.tol = a.tol;
    pb.mxLoop = a.mxLoop;
    XdGeom &g = pb.g;
    g.nz = a.nz; g.ny = a.ny; g.nx = a.nx;
    g.N = a.nz * a.ny * a.nx;
    g.bcy = a.bcy; g.bcx = a.bcx;
    g.i0 = (a.bcx == XINV_BC_PERIODIC) ? 0 : 1;
    g.i1 = (a.bcx == XINV_BC_PERIODIC) ? (int)a.nx : (int)a.nx - 1;
    g.scheme = (a.kind == XD_STD2DT || (pb.hasB && (a.kind == XD_STD2D || a.kind == XD_GEN2D))) ? 4 : 2;
    g.wrapfix = (a.bcx == XINV_BC_PERIODIC) && (a.nx & 1);
    g.ncol = xd_num_colours(g.scheme, g.wrapfix);
    if (a.kind == XD_BIH2D) {                    // 13-point stencil: nine colours, columns 2 .. nx-3 unless periodic
        g.scheme = 9;
        g.wrapfix = (a.bcx == XINV_BC_PERIODIC) && (a.nx % 3 != 0);
        g.ncol = g.wrapfix ? 15 : 9;
        g.i0 = (a.bcx == XINV_BC_PERIODIC) ? 0 : 2;
        g.i1 = (a.bcx == XINV_BC_PERIODIC) ? (int)a.nx : (int)a.nx - 2;
    }
    for (int m = 0; m < 8; ++m) pb.q.p[m] = a.p[m];
    pb.q.optArg = a.optArg;
    pb.q.undef = a.undef;
    pb.accel = o.accel;
    pb.io_f32 = o.io_f32;
    pb.optArg0 = a.optArg;
    pb.omega = 1.0;
    pb.omega_first = true;
    {   // optArg = 2 / (1 + sqrt(1 - rho^2))  <=>  rho^2 = 1 - (2 / optArg - 1)^2
        double t = 2.0 / a.optArg - 1.0;
        t = 1.0 - t * t;
        pb.rho2 = (t > 0.0 && t < 1.0) ? t : 0.0;       // optArg outside (1, 2): no acceleration (omega stays 1 -> optArg is never reached)
        if (o.accel && !(a.optArg > 1.0 && a.optArg < 2.0)) return set_err(XINV_E_ARG, "accel needs 1 < optArg < 2");
    }

    memset(&c->stats, 0, sizeof c->stats);
    c->stats.ncolours = g.ncol;
    if (a.batch == 0) { pb.open = true; pb.h_nactive = 0; return XINV_OK; }

    if (pb.ordering == XINV_ORDER_LEX) {
        if (a.kind > XD_STD3D)
            return set_err(XINV_E_UNSUPPORTED, "lexicographic ordering is offered for the three hot-path kernels only");
        if (pb.hasB && a.bcx == XINV_BC_PERIODIC)
            return set_err(XINV_E_UNSUPPORTED, "lexicographic ordering with a 9-point stencil and periodic-x "
                                               "serialises completely; not offered on the GPU");
    }

    // ---- staging --------------------------------------------------------
    const size_t slice_bytes = (size_t)g.N * sizeof(double);
    cudaEvent_t e0 = c->ev0, e1 = c->ev1;
    XmFront front;
    X3Front front3;
    pb.front = a.front;
    // the user's forcing -> device (float64, or float32 widened on the device)
    auto stage_forcing = [&](void *dst, const void *src, size_t nelem) -> int {
        if (!(o.io_f32 & XINV_IO_F32_IN)) {
            CK(cudaMemcpyAsync(dst, src, nelem * sizeof(double), cudaMemcpyHostToDevice, c->stream));
            c->stats.h2d_bytes += (i64)(nelem * sizeof(double));
            return XINV_OK;
        }
        int rc_ = ensure(c->f32buf, nelem * sizeof(float));
        if (rc_) return rc_;
        CK(cudaMemcpyAsync(c->f32buf.p, src, nelem * sizeof(float), cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes += (i64)(nelem * sizeof(float));
        xd_widen_kernel<<<(unsigned)((nelem + 255) / 256), 256, 0, c->stream>>>((double *)dst, (const float *)c->f32buf.p, (i64)nelem);
        c->stats.kernel_launches++;
        return XINV_OK;
    };
    if (a.front) {
        // the front end hands over A rows, C rows, the user's forcing and the row scale; S is output only
        if (o.ordering != XINV_ORDER_COLOUR || o.engine == XINV_ENGINE_COLOUR)
            return set_err(XINV_E_UNSUPPORTED, "xinv_std2d_rows needs the fused engine (colour ordering)");
        std::string why;
        if (a.kind == XD_STD3D ? !fused3_plan_supported(g, why) : !fused_plan_supported(pb.kind, false, g, why))
            return set_err(XINV_E_UNSUPPORTED, "the device front end needs the fused engine: %s", why.c_str());
        front.on = true;
        front.user_undef = a.user_undef;
        front.out_undef = a.out_undef;
        const size_t row_bytes = (size_t)a.ny * sizeof(double);
        const bool gen = (a.kind == XD_GEN2D);
        front.g_mode = a.g_mode; front.g_p1 = a.g_p1; front.g_p2 = a.g_p2;
        if (a.kind == XD_STD3D) {
            front3.on = true;
            front3.user_undef = a.user_undef; front3.out_undef = a.out_undef;
            for (int m = 0; m < 4; ++m) front3.ns[m] = a.n2_strides[m];
            if (pb.mem_space == XINV_MEM_HOST) {
                CK(cudaEventRecord(e0, c->stream));
                int rc = ensure(c->stage[0], slice_bytes * a.batch);            // S (device only until xinv_end)
                if (rc) return rc;
                if ((rc = ensure(c->stage[2 + 3], slice_bytes * a.batch))) return rc;       // user forcing
                if ((rc = ensure(c->stage[2 + 0], 4 * row_bytes))) return rc;               // the four row vectors
                if ((rc = ensure(c->stage[2 + 1], sizeof(double) * (size_t)a.n2_count))) return rc;   // N2
                if ((rc = stage_forcing(c->stage[5].p, a.coef[3], (size_t)g.N * a.batch))) return rc;
                CK(cudaMemcpyAsync(c->stage[2].p, a.coef[0], 4 * row_bytes, cudaMemcpyHostToDevice, c->stream));
                CK(cudaMemcpyAsync(c->stage[3].p, a.coef[1], sizeof(double) * (size_t)a.n2_count, cudaMemcpyHostToDevice, c->stream));
                c->stats.h2d_bytes += (i64)(4 * row_bytes + sizeof(double) * (size_t)a.n2_count);
                CK(cudaEventRecord(e1, c->stream));
                CK(cudaStreamSynchronize(c->stream));
                float ms = 0;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                c->stats.h2d_ms = ms;
                pb.dS = (double *)c->stage[0].p;
                front3.F = (const double *)c->stage[5].p;
                front3.rows = (const double *)c->stage[2].p;
                front3.N2 = (const double *)c->stage[3].p;
            } else {
                pb.dS = a.S;
                front3.F = a.coef[3]; front3.rows = a.coef[0]; front3.N2 = a.coef[1];
            }
        } else if (gen && pb.mem_space == XINV_MEM_HOST) {
            CK(cudaEventRecord(e0, c->stream));
            int rc = ensure(c->stage[0], slice_bytes * a.batch);            // S (device only until xinv_end)
            if (rc) return rc;
            if ((rc = ensure(c->stage[2 + 6], slice_bytes * a.batch))) return rc;   // user forcing
            if ((rc = ensure(c->stage[2 + 0], 5 * row_bytes))) return rc;           // A, C, D, E, F rows
            if ((rc = stage_forcing(c->stage[8].p, a.coef[6], (size_t)g.N * a.batch))) return rc;
            CK(cudaMemcpyAsync(c->stage[2].p, a.coef[0], 5 * row_bytes, cudaMemcpyHostToDevice, c->stream));
            c->stats.h2d_bytes += (i64)(5 * row_bytes);
            CK(cudaEventRecord(e1, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            c->stats.h2d_ms = ms;
            pb.dS = (double *)c->stage[0].p;
            front.F = (const double *)c->stage[8].p;
            front.rows5 = (const double *)c->stage[2].p;
        } else if (gen) {
            pb.dS = a.S;
            front.F = a.coef[6]; front.rows5 = a.coef[0];
        } else if (pb.mem_space == XINV_MEM_HOST) {
            CK(cudaEventRecord(e0, c->stream));
            int rc = ensure(c->stage[0], slice_bytes * a.batch);            // S (device only until xinv_end)
            if (rc) return rc;
            if ((rc = ensure(c->stage[2 + 3], slice_bytes * a.batch))) return rc;   // user forcing
            if ((rc = ensure(c->stage[2 + 0], 3 * row_bytes))) return rc;           // A rows | C rows | scale
            double *rows = (double *)c->stage[2].p;
            if ((rc = stage_forcing(c->stage[5].p, a.coef[3], (size_t)g.N * a.batch))) return rc;
            CK(cudaMemcpyAsync(rows, a.coef[0], row_bytes, cudaMemcpyHostToDevice, c->stream));
            CK(cudaMemcpyAsync(rows + a.ny, a.coef[2], row_bytes, cudaMemcpyHostToDevice, c->stream));
            if (a.f_scale) CK(cudaMemcpyAsync(rows + 2 * a.ny, a.f_scale, row_bytes, cudaMemcpyHostToDevice, c->stream));
            c->stats.h2d_bytes += (i64)((a.f_scale ? 3 : 2) * row_bytes);
            CK(cudaEventRecord(e1, c->stream));
            CK(cudaStreamSynchronize(c->stream));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            c->stats.h2d_ms = ms;
            pb.dS = (double *)c->stage[0].p;
            front.F = (const double *)c->stage[5].p;
            front.Arow = rows; front.Crow = rows + a.ny; front.scale = a.f_scale ? rows + 2 * a.ny : nullptr;
        } else {
            pb.dS = a.S;
            front.F = a.coef[3]; front.Arow = a.coef[0]; front.Crow = a.coef[2]; front.scale = a.f_scale;
        }
        for (int m = 0; m < 12; ++m) { pb.q.c[m] = nullptr; pb.q.cs[m] = 0; }
    } else if (pb.mem_space == XINV_MEM_HOST) {
        CK(cudaEventRecord(e0, c->stream));
        int rc = ensure(c->stage[0], slice_bytes * a.batch);
        if (rc) return rc;
        pb.dS = (double *)c->stage[0].p;
        if (!a.maskfront) {                                          // (S is output only there: zeroed by the mask kernel)
            CK(cudaMemcpyAsync(pb.dS, a.S, slice_bytes * a.batch, cudaMemcpyHostToDevice, c->stream));
            c->stats.h2d_bytes += (i64)(slice_bytes * a.batch);
        }
        for (int m = 0; m < a.ncoef; ++m) {
            if (!a.coef[m]) { pb.q.c[m] = nullptr; pb.q.cs[m] = 0; continue; }
            const i64 stride = (m >= 8 || o.coef_stride[m] < 0) ? g.N : o.coef_stride[m];      // (xinv_opts has eight strides)
            if (stride != 0 && stride != g.N) return set_err(XINV_E_ARG, "coef_stride[%d]=%lld must be -1, 0 or the slice size for host staging", m, stride);
            const size_t bytes = (stride == 0) ? slice_bytes : slice_bytes * a.batch;
            rc = ensure(c->stage[2 + m], bytes);
            if (rc) return rc;
            CK(cudaMemcpyAsync(c->stage[2 + m].p, a.coef[m], bytes, cudaMemcpyHostToDevice, c->stream));
            c->stats.h2d_bytes += (i64)bytes;
            pb.q.c[m] = (const double *)c->stage[2 + m].p;
            pb.q.cs[m] = stride;
        }
        CK(cudaEventRecord(e1, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        c->stats.h2d_ms = ms;
    } else {
        pb.dS = a.S;
        for (int m = 0; m < a.ncoef; ++m) {
            pb.q.c[m] = a.coef[m];
            pb.q.cs[m] = (!a.coef[m]) ? 0 : ((m >= 8 || o.coef_stride[m] < 0) ? g.N : o.coef_stride[m]);
        }
    }

    if (a.maskfront) {
        // the forcing is masked in a buffer of our own (never in the caller's device array), S starts from zero
        pb.maskfront = true;
        pb.out_undef = a.out_undef;
        if (pb.q.cs[3] != g.N) return set_err(XINV_E_ARG, "xinv_std2d_front: the forcing needs one slice per batch entry");
        const i64 n = g.N * a.batch;
        double *dF;
        if (pb.mem_space == XINV_MEM_HOST) dF = (double *)c->stage[2 + 3].p;
        else {
            int rc_ = ensure(c->stage[2 + 3], slice_bytes * a.batch);
            if (rc_) return rc_;
            dF = (double *)c->stage[2 + 3].p;
            CK(cudaMemcpyAsync(dF, a.coef[3], slice_bytes * a.batch, cudaMemcpyDeviceToDevice, c->stream));
        }
        int rc_ = xm_work_ensure(c->xm_work, 7, 16) == cudaSuccess ? XINV_OK : XINV_E_NOMEM;
        if (rc_) return set_err(rc_, "flag buffer");
        CK(cudaMemsetAsync(c->xm_work.p[7], 0, 8, c->stream));
        xd_front_mask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(dF, pb.dS, n, a.user_undef, a.undef, (int *)c->xm_work.p[7]);
        c->stats.kernel_launches++;
        pb.q.c[3] = dF;
        pb.dFmask = dF;
    }

    int rc;
    // norm partial layout: enough blocks to fill the machine, few enough that the
    // last-block pass stays trivial
    {
        i64 want = (g.N + (i64)XD_NORM_THREADS * 8 - 1) / ((i64)XD_NORM_THREADS * 8);
        i64 per_slice_max = (4LL * c->sm_count + a.batch - 1) / a.batch;
        if (per_slice_max < 1) per_slice_max = 1;
        if (want > per_slice_max) want = per_slice_max;
        if (want < 1) want = 1;
        if (want > 1024) want = 1024;
        pb.nblk_norm = (int)want;
    }

    // ---- engine choice ----------------------------------------------------
    pb.engine = XINV_ENGINE_COLOUR;
    if (o.engine == XINV_ENGINE_RESIDENT) {
        std::string why;
        if (pb.ordering != XINV_ORDER_COLOUR || a.front) why = "needs the colour ordering and dense operands";
        else if (resident_plan_build(pb.resident, c->sm_count, pb.kind, pb.hasB, g, pb.q, pb.batch, pb.dS, why) == 0)
            pb.engine = XINV_ENGINE_RESIDENT;
        if (pb.engine != XINV_ENGINE_RESIDENT)
            return set_err(XINV_E_UNSUPPORTED, "resident engine unavailable: %s", why.c_str());
    } else if (pb.ordering == XINV_ORDER_COLOUR && o.engine != XINV_ENGINE_COLOUR) {
        std::string why;
        if (o.engine == XINV_ENGINE_CLUSTER && pb.kind == XD_STD3D) return set_err(XINV_E_UNSUPPORTED, "cluster engine: 2-D problems only");
        if (pb.kind == XD_STD3D) {
            const char *e3 = getenv("XINV_FUSED3");
            if (e3 && atoi(e3) == 0) why = "disabled by XINV_FUSED3=0";
            else if (pb.accel) why = "accel runs on the colour engine for 3-D problems";
            else if (fused3_plan_supported(g, why) &&
                     fused3_plan_build(pb.fused3, c->xm_work, c->sm_count, g, pb.q, pb.batch, pb.dS, c->stream, why,
                                       a.front ? &front3 : nullptr) == 0)
                pb.engine = XINV_ENGINE_FUSED;
            if (pb.engine != XINV_ENGINE_FUSED && (o.engine == XINV_ENGINE_FUSED || a.front))
                return set_err(XINV_E_UNSUPPORTED, "fused engine unavailable: %s", why.c_str());
        } else if (fused_plan_supported(pb.kind, pb.hasB, g, why)) {
            rc = fused_plan_build(pb.fused, c->xm_work, c->sm_count, pb.kind, g, pb.q, pb.batch, pb.dS, pb.mxLoop, c->stream, why,
                                  a.front ? &front : nullptr);
            if (rc == 0) pb.engine = XINV_ENGINE_FUSED;
            else if (o.engine == XINV_ENGINE_FUSED || o.engine == XINV_ENGINE_CLUSTER || a.front)
                return set_err(XINV_E_UNSUPPORTED, "fused engine unavailable: %s", why.c_str());
            // slices that fit a thread-block cluster stay on the SMs for the whole solve (same operands, xinv_cluster2d.cuh)
            if (rc == 0 && (o.engine == XINV_ENGINE_AUTO || o.engine == XINV_ENGINE_CLUSTER)) {
                const char *ec = getenv("XINV_CLUSTER");
                std::string whyc = "disabled by XINV_CLUSTER=0";
                if (!(ec && atoi(ec) == 0 && o.engine == XINV_ENGINE_AUTO))
                    cluster_plan_build(pb.cluster, pb.fused, c->sm_count, g, pb.q, pb.batch, o.engine == XINV_ENGINE_CLUSTER || pb.accel, whyc);
                if (!pb.cluster.built && o.engine == XINV_ENGINE_CLUSTER)
                    return set_err(XINV_E_UNSUPPORTED, "cluster engine unavailable: %s", whyc.c_str());
            }
            if (rc == 0 && pb.accel && !pb.cluster.built) {          // the marching kernels have the factor baked in
                fused_plan_release(pb.fused);
                pb.engine = XINV_ENGINE_COLOUR;
                if (a.front) return set_err(XINV_E_UNSUPPORTED, "accel: the device front end needs the cluster engine here");
            }
        } else if (o.engine == XINV_ENGINE_FUSED || o.engine == XINV_ENGINE_CLUSTER) {
            return set_err(XINV_E_UNSUPPORTED, "fused engine unavailable: %s", why.c_str());
        }
        // what the fused engines do not take (9-point stencil, x-varying general form, odd nx with periodic-x):
        // small slices go to the resident engine instead of the launch-bound colour engine
        if (pb.engine == XINV_ENGINE_COLOUR && o.engine == XINV_ENGINE_AUTO && !a.front && pb.kind != XD_STD3D) {
            const char *er = getenv("XINV_RESIDENT");
            std::string why2;
            if (!(er && atoi(er) == 0) &&
                resident_plan_build(pb.resident, c->sm_count, pb.kind, pb.hasB, g, pb.q, pb.batch, pb.dS, why2) == 0)
                pb.engine = XINV_ENGINE_RESIDENT;
        }
    }
    {
        int np = pb.nblk_norm;
        if (pb.engine == XINV_ENGINE_FUSED && pb.fused.nblk_partials > np) np = pb.fused.nblk_partials;
        if (pb.engine == XINV_ENGINE_FUSED && pb.fused3.nblk_partials > np) np = pb.fused3.nblk_partials;
        if ((rc = ensure(c->psum, sizeof(double) * a.batch * np))) return rc;
        if ((rc = ensure(c->pcnt, sizeof(i64) * a.batch * np))) return rc;
    }
    // ---- per-slice state ----------------------------------------------------
    if ((rc = ensure(c->state, sizeof(XdSliceState) * a.batch))) return rc;
    if ((rc = ensure(c->nactive, 64))) return rc;
    if ((rc = ensure(c->ticket, sizeof(unsigned) * a.batch))) return rc;
    // flags in -> device
    DevBuf &ftmp = c->flags_in;
    if ((rc = ensure(ftmp, sizeof(double) * 3 * a.batch))) return rc;
    CK(cudaMemcpyAsync(ftmp.p, a.flags, sizeof(double) * 3 * a.batch, cudaMemcpyHostToDevice, c->stream));
    CK(cudaMemsetAsync(c->ticket.p, 0, sizeof(unsigned) * a.batch, c->stream));
    {
        // iterations of the first pass: T, but never more sweeps than mxLoop allows (loop = 0 .. mxLoop)
        const int nit0 = (pb.engine == XINV_ENGINE_FUSED && pb.fused.built) ? (int)((a.mxLoop + 1 < (i64)pb.fused.T) ? a.mxLoop + 1 : (i64)pb.fused.T) : 1;
        xd_init_state_kernel<<<(unsigned)((a.batch + 127) / 128), 128, 0, c->stream>>>(
            (XdSliceState *)c->state.p, (const double *)ftmp.p, (int)a.batch, (int *)c->nactive.p, nit0);
    }
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(c->stream));       // a.flags (caller's host memory) has been read
    if (a.front || a.maskfront) {               // did the front end meet a non-finite unmasked forcing value?
        int fl[2] = {0, 0};
        CK(cudaMemcpy(fl, c->xm_work.p[7], sizeof fl, cudaMemcpyDeviceToHost));
        if (fl[1]) {
            fused_plan_release(pb.fused);
            fused3_plan_release(pb.fused3);
            return set_err(XINV_E_UNSUPPORTED, "device front end: an unmasked forcing value is not finite (use the full-array entry)");
        }
    }
    pb.h_nactive = (int)a.batch;

    c->stats.engine = pb.cluster.built ? XINV_ENGINE_CLUSTER : pb.engine;
    c->stats.iters_per_pass = (pb.engine == XINV_ENGINE_FUSED && pb.fused.built && !pb.cluster.built) ? pb.fused.T : 1;
    c->stats.row_coeffs = (pb.engine == XINV_ENGINE_FUSED && ((pb.fused.built && pb.fused.rc) || (pb.fused3.built && pb.fused3.arow))) ? 1 : 0;
    if (pb.engine == XINV_ENGINE_FUSED && pb.fused3.built && pb.fused3.rowsmode) c->stats.row_coeffs = 2;
    pb.open = true;
    return XINV_OK;
}

// ---------------------------------------------------------------------------
// one sweep (all colours + norm/decide) on every active slice
// ---------------------------------------------------------------------------
template <int KIND>
static void launch_colour(xinv_ctx *c, Problem &pb, int colour, dim3 grid, int nxblk)
{
    XdSliceState *st = (XdSliceState *)c->state.p;
    if (pb.hasB && (KIND == XD_STD2D || KIND == XD_GEN2D))
        xd_sweep_colour_kernel<KIND, true><<<grid, XD_SWEEP_THREADS, 0, c->stream>>>(pb.dS, pb.q, pb.g, colour, nxblk, st);
    else
        xd_sweep_colour_kernel<KIND, false><<<grid, XD_SWEEP_THREADS, 0, c->stream>>>(pb.dS, pb.q, pb.g, colour, nxblk, st);
}

static void prof_mark(xinv_ctx *c, const Problem &pb)
{
    if (!pb.profile) return;
    if (c->prof_used == c->prof_ev.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) return;
        c->prof_ev.push_back(e);
    }
    cudaEventRecord(c->prof_ev[c->prof_used++], c->stream);
}

static void prof_collect(xinv_ctx *c, int launches_per_pair)
{
    for (size_t i = 0; i + 1 < c->prof_used; i += 2) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, c->prof_ev[i], c->prof_ev[i + 1]) == cudaSuccess) {
            c->stats.dom_ms += ms;
            c->stats.dom_launches += launches_per_pair;
        }
    }
    c->prof_used = 0;
}

static int sweep_colour_engine(xinv_ctx *c, Problem &pb)
{
    const XdGeom &g = pb.g;
    XdSliceState *st = (XdSliceState *)c->state.p;
    if (pb.kind == XD_BIH2D) {
        if (g.bcy == XINV_BC_EXTEND) {
            dim3 grid((unsigned)((g.nx + 127) / 128), (unsigned)pb.batch);
            xd_extend_bih_kernel<<<grid, 128, 0, c->stream>>>(pb.dS, g, pb.q.undef, st);
            c->stats.kernel_launches++;
        }
        const i64 rows4 = g.ny - 4;
        if (rows4 > 0 && g.i1 > g.i0) {
            prof_mark(c, pb);
            for (int col = 0; col < g.ncol; ++col) {
                const int nxblk = (col >= 9) ? 1 : (int)(((g.nx + 2) / 3 + XD_SWEEP_THREADS - 1) / XD_SWEEP_THREADS);
                dim3 grid((unsigned)(rows4 * nxblk), (unsigned)pb.batch, 1);
                xd_sweep_bih_kernel<<<grid, XD_SWEEP_THREADS, 0, c->stream>>>(pb.dS, pb.q, g, col, nxblk, st);
                c->stats.kernel_launches++;
            }
            prof_mark(c, pb);
        }
        dim3 ngrid((unsigned)pb.nblk_norm, (unsigned)pb.batch, 1);
        xd_norm_decide_kernel<<<ngrid, XD_NORM_THREADS, 0, c->stream>>>(
            pb.dS, g.N, pb.q.undef, pb.nblk_norm, (double *)c->psum.p, (i64 *)c->pcnt.p,
            (unsigned *)c->ticket.p, st, (int *)c->nactive.p, pb.tol, pb.mxLoop, pb.zero_exit);
        c->stats.kernel_launches++;
        return XINV_OK;
    }
    if (pb.kind == XD_STD1D) {
        if (g.bcx == XINV_BC_EXTEND) {
            xd_extend1d_kernel<<<(unsigned)((pb.batch + 127) / 128), 128, 0, c->stream>>>(pb.dS, g.nx, (int)pb.batch, pb.q.undef, st);
            c->stats.kernel_launches++;
        }
    } else if (g.bcy == XINV_BC_EXTEND) {
        const i64 levels = (g.nz > 1) ? g.nz - 2 : 1;
        if (levels > 0) {
            dim3 grid((unsigned)((g.nx + 127) / 128), (unsigned)levels, (unsigned)pb.batch);
            xd_extend_kernel<<<grid, 128, 0, c->stream>>>(pb.dS, g, pb.q.undef, st);
            c->stats.kernel_launches++;
        }
    }
    const i64 rows = (pb.kind == XD_STD1D) ? 1 : (g.ny >= 3 ? g.ny - 2 : 0) * (XD_IS3D(pb.kind) ? (g.nz >= 3 ? g.nz - 2 : 0) : 1);
    const int base = (g.scheme == 4) ? 4 : 2;
    if (rows > 0 && g.i1 > g.i0) {
        prof_mark(c, pb);
        double w1 = pb.omega;
        if (pb.accel) w1 = xd_cheb_next(pb.omega, pb.rho2, pb.omega_first);
        for (int col = 0; col < g.ncol; ++col) {
            // Chebyshev: the first half of the colours with omega_h, the second half (and the wrap-fix colours) with omega_h+1
            if (pb.accel) pb.q.optArg = (col < base / 2) ? pb.omega : w1;
            int nxblk;
            if (col >= base) nxblk = 1;
            else nxblk = (int)(((g.nx + 1) / 2 + XD_SWEEP_THREADS - 1) / XD_SWEEP_THREADS);
            dim3 grid((unsigned)(rows * nxblk), (unsigned)pb.batch, 1);
            if (pb.kind == XD_STD2D) launch_colour<XD_STD2D>(c, pb, col, grid, nxblk);
            else if (pb.kind == XD_GEN2D) launch_colour<XD_GEN2D>(c, pb, col, grid, nxblk);
            else if (pb.kind == XD_STD3D) launch_colour<XD_STD3D>(c, pb, col, grid, nxblk);
            else if (pb.kind == XD_STD2DT) launch_colour<XD_STD2DT>(c, pb, col, grid, nxblk);
            else if (pb.kind == XD_GEN3D) launch_colour<XD_GEN3D>(c, pb, col, grid, nxblk);
            else launch_colour<XD_STD1D>(c, pb, col, grid, nxblk);
            c->stats.kernel_launches++;
        }
        prof_mark(c, pb);
        if (pb.accel) { pb.omega = xd_cheb_next(w1, pb.rho2, false); pb.omega_first = false; }
    }
    dim3 ngrid((unsigned)pb.nblk_norm, (unsigned)pb.batch, 1);
    xd_norm_decide_kernel<<<ngrid, XD_NORM_THREADS, 0, c->stream>>>(
        pb.dS, g.N, pb.q.undef, pb.nblk_norm, (double *)c->psum.p, (i64 *)c->pcnt.p,
        (unsigned *)c->ticket.p, st, (int *)c->nactive.p, pb.tol, pb.mxLoop, pb.zero_exit);
    c->stats.kernel_launches++;
    return XINV_OK;
}

static int auto_check_every(const xinv_ctx *c, const Problem &pb)
{
    if (pb.check_every > 0) return pb.check_every;
    // resident engine: a launch iterates inside the SM until its slices stop; the budget only bounds
    // how long a launch may run (operands are staged again by the next one)
    if (pb.engine == XINV_ENGINE_RESIDENT || pb.cluster.built) return 4096;
    // Aim at ~6 ms of device work between host polls of the active count.  Passes launched
    // after every slice has stopped find nothing to do (a few microseconds each), so polling
    // rarely costs little; polling often costs a stream synchronisation per poll (measured on
    // the 32-slice C5 workload: 5 % of the step at ~1.3 ms between polls).
    const double bytes_per_pass = (double)pb.g.N * (double)pb.batch * (pb.engine == XINV_ENGINE_FUSED ? 40.0 : 80.0);
    const double est_us = 5.0 + bytes_per_pass / 4.0e6;             // ~4 TB/s = 4e6 B/us
    int k = (int)(6000.0 / est_us);
    if (k < 8) k = 8;
    if (k > 256) k = 256;
    (void)c;
    return k;
}

extern "C" int xinv_step(xinv_ctx *c, int64_t sweeps, int64_t *n_active_out)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    Problem &pb = c->pb;
    if (!pb.open) return set_err(XINV_E_STATE, "no open problem (call xinv_*_begin first)");
    CK(cudaSetDevice(c->device));
    if (pb.batch == 0 || pb.h_nactive == 0) { if (n_active_out) *n_active_out = 0; return XINV_OK; }
    // a pass (launch) performs >= 1 sweep on every active slice, except for at most one
    // "redo" pass per slice (fused engine, T > 1): at most mxLoop + 2 passes (numbas.py:410)
    const bool f3 = (pb.engine == XINV_ENGINE_FUSED && pb.fused3.built);
    const i64 max_passes = pb.mxLoop + 1 + ((pb.engine == XINV_ENGINE_FUSED && !f3 && !pb.cluster.built && pb.fused.T > 1) ? 1 : 0);
    const i64 remaining = max_passes - pb.sweeps_launched;
    if (sweeps <= 0) sweeps = auto_check_every(c, pb);
    if (sweeps > remaining) sweeps = remaining;
    CK(cudaEventRecord(c->ev0, c->stream));
    for (i64 it = 0; it < sweeps;) {
        int rc;
        i64 did = 1;
        if (pb.ordering == XINV_ORDER_LEX)
            rc = lex_sweep(c->stream, pb.kind, pb.hasB, pb.g, pb.q, pb.batch, pb.dS, (XdSliceState *)c->state.p,
                           pb.nblk_norm, (double *)c->psum.p, (i64 *)c->pcnt.p, (unsigned *)c->ticket.p,
                           (int *)c->nactive.p, pb.tol, pb.mxLoop, pb.zero_exit, &c->stats.kernel_launches);
        else if (pb.engine == XINV_ENGINE_RESIDENT) {
            did = sweeps - it;                                                   // the whole chunk in one launch
            rc = resident_sweep(pb.resident, c->stream, (XdSliceState *)c->state.p, (int *)c->nactive.p, pb.tol, pb.mxLoop,
                                pb.zero_exit, (int)did, pb.accel, pb.rho2, &c->stats.kernel_launches);
            if (rc) return set_err(XINV_E_CUDA, "resident engine launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else if (pb.cluster.built) {
            did = sweeps - it;                                                   // the whole chunk in one launch
            rc = cluster_sweep(pb.cluster, c->stream, (XdSliceState *)c->state.p, (int *)c->nactive.p, pb.tol, pb.mxLoop,
                               pb.zero_exit, (int)did, pb.accel, pb.rho2, &c->stats.kernel_launches);
            if (rc) return set_err(XINV_E_CUDA, "cluster engine launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else if (f3) {
            did = sweeps - it < pb.fused3.ppl ? sweeps - it : pb.fused3.ppl;     // passes in this launch
            rc = fused3_sweep(pb.fused3, c->stream, (XdSliceState *)c->state.p, (double *)c->psum.p, (i64 *)c->pcnt.p,
                              (unsigned *)c->ticket.p, (int *)c->nactive.p, pb.tol, pb.mxLoop, (int)did,
                              &c->stats.kernel_launches);
            if (rc) return set_err(XINV_E_CUDA, "3-D fused engine launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else if (pb.engine == XINV_ENGINE_FUSED) {
            did = sweeps - it < pb.fused.ppl ? sweeps - it : pb.fused.ppl;       // passes in this launch
            rc = fused_sweep(pb.fused, c->stream, (XdSliceState *)c->state.p, (double *)c->psum.p, (i64 *)c->pcnt.p,
                             (unsigned *)c->ticket.p, (int *)c->nactive.p, pb.tol, pb.mxLoop, pb.zero_exit, (int)did,
                             &c->stats.kernel_launches);
            if (rc) return set_err(XINV_E_CUDA, "fused engine launch failed: %s", cudaGetErrorString(cudaGetLastError()));
        } else
            rc = sweep_colour_engine(c, pb);
        if (rc) return rc;
        it += did;
    }
    pb.sweeps_launched += sweeps;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));     // [ev0, ev1] = the kernels of this chunk, back to back
    CK(cudaMemcpyAsync(c->h_nactive_pinned, c->nactive.p, 2 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.solve_ms += ms;
    if (pb.profile) {
        if (pb.engine == XINV_ENGINE_FUSED && pb.ordering != XINV_ORDER_LEX) {
            // the chunk holds nothing but fused passes: its event bracket is their summed duration
            // (inter-launch gaps included), without an event pair around every launch
            c->stats.dom_ms += ms;
            c->stats.dom_launches += sweeps;
        } else
            prof_collect(c, pb.g.ncol);
    }
    c->stats.sweeps_launched = pb.sweeps_launched;
    pb.h_nactive = c->h_nactive_pinned[0];
    if (pb.h_nactive != 0 && pb.sweeps_launched >= max_passes)
        return set_err(XINV_E_STATE, "internal error: %d slices still active after %lld passes", pb.h_nactive,
                       (long long)max_passes);
    if (n_active_out) *n_active_out = pb.h_nactive;
    return XINV_OK;
}

extern "C" int xinv_end(xinv_ctx *c)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    Problem &pb = c->pb;
    if (!pb.open) return set_err(XINV_E_STATE, "no open problem");
    CK(cudaSetDevice(c->device));
    pb.open = false;
    if (pb.batch == 0) return XINV_OK;
    const XdGeom &g = pb.g;
    if (pb.engine == XINV_ENGINE_FUSED && pb.fused3.built) {
        fused3_unpack(pb.fused3, pb.dS, (const XdSliceState *)c->state.p, c->stream);
        c->stats.kernel_launches++;
        CK(cudaGetLastError());
    } else if (pb.engine == XINV_ENGINE_FUSED) {
        fused_unpack(pb.fused, pb.dS, (const XdSliceState *)c->state.p, c->stream);
        c->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    if (pb.maskfront) {
        const i64 n = g.N * pb.batch;
        xd_front_demask_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>(pb.dS, pb.dFmask, n, pb.q.undef, pb.out_undef);
        c->stats.kernel_launches++;
    }
    std::vector<XdSliceState> hs((size_t)pb.batch);
    CK(cudaEventRecord(c->ev0, c->stream));
    if (pb.mem_space == XINV_MEM_HOST && (pb.io_f32 & XINV_IO_F32_OUT)) {
        const i64 n = g.N * pb.batch;
        int rc_ = ensure(c->f32buf, (size_t)n * sizeof(float));
        if (rc_) return rc_;
        xd_narrow_kernel<<<(unsigned)((n + 255) / 256), 256, 0, c->stream>>>((float *)c->f32buf.p, pb.dS, n);
        c->stats.kernel_launches++;
        CK(cudaMemcpyAsync(pb.userS, c->f32buf.p, sizeof(float) * n, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (i64)sizeof(float) * n;
    } else if (pb.mem_space == XINV_MEM_HOST) {
        CK(cudaMemcpyAsync(pb.userS, pb.dS, sizeof(double) * g.N * pb.batch, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes += (i64)sizeof(double) * g.N * pb.batch;
    }
    CK(cudaMemcpyAsync(hs.data(), c->state.p, sizeof(XdSliceState) * pb.batch, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.d2h_ms = ms;
    i64 updates = 0;
    int max_done = 0;
    for (i64 b = 0; b < pb.batch; ++b) {
        if (hs[b].sweeps_done > max_done) max_done = hs[b].sweeps_done;
        pb.userFlags[3 * b + 0] = hs[b].flags[0];
        pb.userFlags[3 * b + 1] = hs[b].flags[1];
        pb.userFlags[3 * b + 2] = hs[b].flags[2];
        updates += (i64)hs[b].sweeps_done * g.N;
    }
    c->stats.cell_updates = updates;
    c->stats.sweep_ms = max_done ? c->stats.solve_ms / (double)max_done : 0.0;
    if (pb.profile && pb.engine == XINV_ENGINE_FUSED && pb.ordering != XINV_ORDER_LEX && pb.fused3.built) {
        c->stats.dom_launches = max_done;
    } else if (pb.profile && pb.engine == XINV_ENGINE_FUSED && pb.ordering != XINV_ORDER_LEX && pb.fused.T > 0) {
        // passes that did work: the chunks were timed as a whole and may end with passes that found every
        // slice stopped (a few microseconds each, left in dom_ms); count only the real ones
        c->stats.dom_launches = pb.cluster.built ? max_done : (max_done + pb.fused.T - 1) / pb.fused.T;   // cluster engine: per sweep
    }
    fused_plan_release(pb.fused);
    fused3_plan_release(pb.fused3);
    resident_plan_release(pb.resident);
    cluster_plan_release(pb.cluster);
    return XINV_OK;
}

// One sweep loop at a time per device.  Several contexts of one GPU exist so that the copies of one chunk of a batch
// run under the solve of another (solvers._execute); their sweep loops must not interleave: a loop is a sequence of
// (cooperative, whole-GPU) launches with a host poll in between, and two of them taking turns launch by launch both
// finish late -- the chunk whose result should be on its way to the host is still iterating (measured on C5,
// 3 chunks over 2 contexts: 25.2 ms per step interleaved against 21 ms with the loops one after the other).
static std::mutex g_solve_mutex[64];

static int run_to_completion(xinv_ctx *c)
{
    int64_t na = 1;
    int rc = XINV_OK;
    {
        std::lock_guard<std::mutex> turn(g_solve_mutex[(unsigned)c->device % 64u]);
        while (na > 0) {
            rc = xinv_step(c, 0, &na);
            if (rc) break;
        }
    }
    int rc2 = xinv_end(c);
    return rc ? rc : rc2;
}

// ---------------------------------------------------------------------------
// extern "C" solvers
// ---------------------------------------------------------------------------
extern "C" int xinv_std2d_begin(xinv_ctx *ctx, double *S, const double *A, const double *B,
                                const double *C, const double *F,
                                int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                                double delxSqr, double ratioQtr, double ratioSqr,
                                double optArg, double undef, double *flags,
                                int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    BeginArgs a{};
    a.kind = XD_STD2D; a.S = S;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = C; a.coef[3] = F; a.ncoef = 4; a.b_index = 1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSqr; a.p[1] = ratioQtr; a.p[2] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    return problem_begin(ctx, a);
}

extern "C" int xinv_std2d_rows(xinv_ctx *ctx, double *S_out, const double *A_rows, const double *C_rows,
                               const double *F_user, const double *F_row_scale, double user_undef, double out_undef,
                               int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                               double delxSqr, double ratioQtr, double ratioSqr,
                               double optArg, double undef, double *flags,
                               int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    BeginArgs a{};
    a.kind = XD_STD2D; a.S = S_out;
    a.coef[0] = A_rows; a.coef[1] = nullptr; a.coef[2] = C_rows; a.coef[3] = F_user; a.ncoef = 4; a.b_index = 1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSqr; a.p[1] = ratioQtr; a.p[2] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    a.front = true; a.f_scale = F_row_scale; a.user_undef = user_undef; a.out_undef = out_undef;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_gen2d_rows(xinv_ctx *ctx, double *S_out, const double *rows, const double *G_user,
                               int g_mode, double g_p1, double g_p2, double user_undef, double out_undef,
                               int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                               double delx, double delxSqr, double ratio, double ratioQtr, double ratioSqr,
                               double optArg, double undef, double *flags,
                               int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    BeginArgs a{};
    a.kind = XD_GEN2D; a.S = S_out;
    for (int m = 0; m < 7; ++m) a.coef[m] = rows;          // only [0] (the row block) and [6] (forcing) are used
    a.coef[1] = nullptr; a.coef[6] = G_user; a.ncoef = 7; a.b_index = 1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delx; a.p[1] = delxSqr; a.p[2] = ratio; a.p[3] = ratioQtr; a.p[4] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    a.front = true; a.user_undef = user_undef; a.out_undef = out_undef;
    a.g_mode = g_mode; a.g_p1 = g_p1; a.g_p2 = g_p2;
    if (!rows || !G_user) return set_err(XINV_E_ARG, "rows and G_user must not be NULL");
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_gen2d_begin(xinv_ctx *ctx, double *S, const double *A, const double *B,
                                const double *C, const double *D, const double *E,
                                const double *F, const double *G,
                                int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                                double delx, double delxSqr, double ratio, double ratioQtr,
                                double ratioSqr, double optArg, double undef, double *flags,
                                int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    BeginArgs a{};
    a.kind = XD_GEN2D; a.S = S;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = C; a.coef[3] = D; a.coef[4] = E; a.coef[5] = F; a.coef[6] = G;
    a.ncoef = 7; a.b_index = 1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delx; a.p[1] = delxSqr; a.p[2] = ratio; a.p[3] = ratioQtr; a.p[4] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    return problem_begin(ctx, a);
}

extern "C" int xinv_std3d_begin(xinv_ctx *ctx, double *S, const double *A, const double *B,
                                const double *C, const double *F,
                                int64_t batch, int64_t nz, int64_t ny, int64_t nx,
                                int bcz, int bcy, int bcx,
                                double delxSqr, double ratio2Sqr, double ratio1Sqr,
                                double optArg, double undef, double *flags,
                                int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    if (!valid_bc(bcz)) return set_err(XINV_E_ARG, "bad boundary condition code");
    BeginArgs a{};
    a.kind = XD_STD3D; a.S = S;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = C; a.coef[3] = F; a.ncoef = 4; a.b_index = -1;
    a.batch = batch; a.nz = nz; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSqr; a.p[1] = ratio2Sqr; a.p[2] = ratio1Sqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    return problem_begin(ctx, a);
}

// invert_Eliassen's front end: dense (or batch-shared) A, B, C and the user's forcing; masks, zero initial guess and
// de-masking on the device.  Any engine the dense entry would use (resident for small sections, colour, marching).
extern "C" int xinv_std2d_front(xinv_ctx *ctx, double *S_out, const double *A, const double *B, const double *C,
                                const double *F_user, double user_undef, double out_undef,
                                int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                                double delxSqr, double ratioQtr, double ratioSqr, double optArg, double undef,
                                double *flags, int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    BeginArgs a{};
    a.kind = XD_STD2D; a.S = S_out;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = C; a.coef[3] = F_user; a.ncoef = 4; a.b_index = 1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSqr; a.p[1] = ratioQtr; a.p[2] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    a.maskfront = true; a.user_undef = user_undef; a.out_undef = out_undef;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

// ---- the remaining kernels of numbas.py (colour engine) ----
extern "C" int xinv_std2d_test(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *C,
                               const double *D, const double *E, const double *F,
                               int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                               double delxSqr, double ratioQtr, double ratioSqr, double optArg, double undef,
                               double *flags, int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    BeginArgs a{};
    a.kind = XD_STD2DT; a.S = S;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = C; a.coef[3] = D; a.coef[4] = E; a.coef[5] = F; a.ncoef = 6; a.b_index = -1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSqr; a.p[1] = ratioQtr; a.p[2] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_gen3d(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *C,
                          const double *D, const double *E, const double *F, const double *G, const double *H,
                          int64_t batch, int64_t nz, int64_t ny, int64_t nx, int bcz, int bcy, int bcx,
                          double delx, double delxSqr, double ratio2, double ratio1, double ratio2Sqr, double ratio1Sqr,
                          double optArg, double undef, double *flags, int64_t mxLoop, double tolerance,
                          const xinv_opts *opts)
{
    if (!valid_bc(bcz)) return set_err(XINV_E_ARG, "bad boundary condition code");
    BeginArgs a{};
    a.kind = XD_GEN3D; a.S = S;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = C; a.coef[3] = D; a.coef[4] = E; a.coef[5] = F; a.coef[6] = G; a.coef[7] = H;
    a.ncoef = 8; a.b_index = -1;
    a.batch = batch; a.nz = nz; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delx; a.p[1] = delxSqr; a.p[2] = ratio2; a.p[3] = ratio1; a.p[4] = ratio2Sqr; a.p[5] = ratio1Sqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_bih2d(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *C, const double *D,
                          const double *E, const double *F, const double *G, const double *H, const double *I,
                          const double *J, int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                          double delxSSr, double delxTr, double delxSqr, double ratio, double ratioSSr, double ratioQtr,
                          double ratioSqr, double optArg, double undef, double *flags, int64_t mxLoop, double tolerance,
                          const xinv_opts *opts)
{
    if (ny < 5 || nx < 5) return set_err(XINV_E_ARG, "the 13-point stencil needs ny >= 5 and nx >= 5");
    BeginArgs a{};
    a.kind = XD_BIH2D; a.S = S;
    const double *arr[10] = {A, B, C, D, E, F, G, H, I, J};
    for (int m = 0; m < 10; ++m) a.coef[m] = arr[m];
    a.ncoef = 10; a.b_index = -1;
    a.batch = batch; a.nz = 1; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSSr; a.p[1] = delxTr; a.p[2] = delxSqr; a.p[3] = ratio; a.p[4] = ratioSSr; a.p[5] = ratioQtr; a.p[6] = ratioSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_std1d(xinv_ctx *ctx, double *S, const double *A, const double *B, const double *F,
                          int64_t batch, int64_t nx, int bcx, double delxSqr, double optArg, double undef,
                          double *flags, int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    if (nx < 3) return set_err(XINV_E_ARG, "nx < 3");
    BeginArgs a{};
    a.kind = XD_STD1D; a.S = S;
    a.coef[0] = A; a.coef[1] = B; a.coef[2] = F; a.ncoef = 3; a.b_index = -1;
    a.batch = batch; a.nz = 1; a.ny = 1; a.nx = nx; a.bcy = XINV_BC_FIXED; a.bcx = bcx;
    a.p[0] = delxSqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_std3d_rows(xinv_ctx *ctx, double *S_out, const double *rows, const double *N2,
                               const int64_t *n2_strides, int64_t n2_count, const double *F_user,
                               double user_undef, double out_undef,
                               int64_t batch, int64_t nz, int64_t ny, int64_t nx, int bcz, int bcy, int bcx,
                               double delxSqr, double ratio2Sqr, double ratio1Sqr,
                               double optArg, double undef, double *flags,
                               int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    if (!valid_bc(bcz)) return set_err(XINV_E_ARG, "bad boundary condition code");
    if (!rows || !N2 || !n2_strides || !F_user) return set_err(XINV_E_ARG, "rows, N2, n2_strides and F_user must not be NULL");
    if (n2_count < 1) return set_err(XINV_E_ARG, "n2_count < 1");
    {   // the largest element N2 is read at must lie inside the buffer
        const int64_t ext[4] = {batch, nz, ny, nx};
        int64_t last = 0;
        for (int m = 0; m < 4; ++m) {
            if (n2_strides[m] < 0) return set_err(XINV_E_ARG, "n2_strides must be >= 0");
            if (ext[m] > 0) last += n2_strides[m] * (ext[m] - 1);
        }
        if (last >= n2_count) return set_err(XINV_E_ARG, "n2_strides reach element %lld of an N2 buffer of %lld", (long long)last, (long long)n2_count);
    }
    BeginArgs a{};
    a.kind = XD_STD3D; a.S = S_out;
    a.coef[0] = rows; a.coef[1] = N2; a.coef[2] = rows; a.coef[3] = F_user; a.ncoef = 4; a.b_index = -1;
    a.batch = batch; a.nz = nz; a.ny = ny; a.nx = nx; a.bcy = bcy; a.bcx = bcx;
    a.p[0] = delxSqr; a.p[1] = ratio2Sqr; a.p[2] = ratio1Sqr;
    a.optArg = optArg; a.undef = undef; a.flags = flags; a.mxLoop = mxLoop; a.tol = tolerance; a.opts = opts;
    a.front = true; a.user_undef = user_undef; a.out_undef = out_undef;
    for (int m = 0; m < 4; ++m) a.n2_strides[m] = n2_strides[m];
    a.n2_count = n2_count;
    int rc = problem_begin(ctx, a);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_std2d(xinv_ctx *ctx, double *S, const double *A, const double *B,
                          const double *C, const double *F,
                          int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                          double delxSqr, double ratioQtr, double ratioSqr,
                          double optArg, double undef, double *flags,
                          int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    int rc = xinv_std2d_begin(ctx, S, A, B, C, F, batch, ny, nx, bcy, bcx, delxSqr, ratioQtr, ratioSqr,
                              optArg, undef, flags, mxLoop, tolerance, opts);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_gen2d(xinv_ctx *ctx, double *S, const double *A, const double *B,
                          const double *C, const double *D, const double *E,
                          const double *F, const double *G,
                          int64_t batch, int64_t ny, int64_t nx, int bcy, int bcx,
                          double delx, double delxSqr, double ratio, double ratioQtr,
                          double ratioSqr, double optArg, double undef, double *flags,
                          int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    int rc = xinv_gen2d_begin(ctx, S, A, B, C, D, E, F, G, batch, ny, nx, bcy, bcx, delx, delxSqr, ratio,
                              ratioQtr, ratioSqr, optArg, undef, flags, mxLoop, tolerance, opts);
    if (rc) return rc;
    return run_to_completion(ctx);
}

extern "C" int xinv_std3d(xinv_ctx *ctx, double *S, const double *A, const double *B,
                          const double *C, const double *F,
                          int64_t batch, int64_t nz, int64_t ny, int64_t nx,
                          int bcz, int bcy, int bcx,
                          double delxSqr, double ratio2Sqr, double ratio1Sqr,
                          double optArg, double undef, double *flags,
                          int64_t mxLoop, double tolerance, const xinv_opts *opts)
{
    int rc = xinv_std3d_begin(ctx, S, A, B, C, F, batch, nz, ny, nx, bcz, bcy, bcx, delxSqr, ratio2Sqr,
                              ratio1Sqr, optArg, undef, flags, mxLoop, tolerance, opts);
    if (rc) return rc;
    return run_to_completion(ctx);
}

// ---------------------------------------------------------------------------
// cal_flow epilogue
// ---------------------------------------------------------------------------
extern "C" int xinv_flow2d(xinv_ctx *c, double *out1, double *out2, const double *S,
                           int64_t batch, int64_t ny, int64_t nx, const xinv_flow_desc *d, const xinv_opts *opts)
{
    if (!c || !out1 || !out2 || !S || !d) return set_err(XINV_E_ARG, "NULL argument");
    if (d->struct_size != (int32_t)sizeof(xinv_flow_desc)) return set_err(XINV_E_ARG, "desc.struct_size %d != %zu", d->struct_size, sizeof(xinv_flow_desc));
    if (batch < 0 || batch > 65535 || ny < 2 || nx < 2) return set_err(XINV_E_ARG, "bad sizes batch=%lld ny=%lld nx=%lld", (long long)batch, (long long)ny, (long long)nx);
    if (d->comb < XINV_FLOW_GRAD || d->comb > XINV_FLOW_GM_CART) return set_err(XINV_E_ARG, "bad comb");
    const int need_rows = (d->comb == XINV_FLOW_GM_LL) ? 3 : 2;
    if (!d->rows || d->nrows != need_rows) return set_err(XINV_E_ARG, "rows: %d vectors needed", need_rows);
    const xinv_flow_axis *axs[2] = {&d->y, &d->x};
    const int64_t len[2] = {ny, nx};
    for (int m = 0; m < 2; ++m) {
        if (axs[m]->edge < XINV_EDGE_ONESIDED || axs[m]->edge > XINV_EDGE_PERIODIC) return set_err(XINV_E_ARG, "bad edge code");
        if (!axs[m]->uniform && !axs[m]->w) return set_err(XINV_E_ARG, "non-uniform axis without weights");
    }
    int mem = XINV_MEM_HOST;
    if (opts) {
        if (opts->struct_size != (int32_t)sizeof(xinv_opts)) return set_err(XINV_E_ARG, "opts.struct_size");
        mem = opts->mem_space;
    }
    CK(cudaSetDevice(c->device));
    memset(&c->stats, 0, sizeof c->stats);
    if (batch == 0) return XINV_OK;
    const size_t bytes = sizeof(double) * (size_t)batch * ny * nx;
    XfArgs a{};
    a.ny = ny; a.nx = nx; a.comb = d->comb; a.swap = d->swap; a.s1 = d->s1; a.s2 = d->s2; a.deg2m = d->deg2m;
    const double *dS = S;
    double *d1 = out1, *d2 = out2;
    int rc;
    // small operands: rows | y weights | x weights in one staging buffer
    const size_t nsmall = (size_t)need_rows * ny + 3 * (size_t)ny + 3 * (size_t)nx;
    if ((rc = ensure(c->stage[3], sizeof(double) * nsmall))) return rc;
    double *sm = (double *)c->stage[3].p;
    const cudaMemcpyKind kin = (mem == XINV_MEM_HOST) ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToDevice;
    CK(cudaMemcpyAsync(sm, d->rows, sizeof(double) * need_rows * ny, kin, c->stream));
    a.rows = sm;
    double *wp = sm + (size_t)need_rows * ny;
    XfAxis *dst[2] = {&a.y, &a.x};
    for (int m = 0; m < 2; ++m) {
        dst[m]->uniform = axs[m]->uniform; dst[m]->edge = axs[m]->edge; dst[m]->den = axs[m]->den;
        dst[m]->lo = axs[m]->lo; dst[m]->hi = axs[m]->hi; dst[m]->w = nullptr;
        if (!axs[m]->uniform) {
            CK(cudaMemcpyAsync(wp, axs[m]->w, sizeof(double) * 3 * len[m], kin, c->stream));
            dst[m]->w = wp;
        }
        wp += 3 * len[m];
    }
    if (mem == XINV_MEM_HOST) {
        if ((rc = ensure(c->stage[0], bytes))) return rc;
        if ((rc = ensure(c->stage[1], bytes))) return rc;
        if ((rc = ensure(c->stage[2], bytes))) return rc;
        CK(cudaMemcpyAsync(c->stage[0].p, S, bytes, cudaMemcpyHostToDevice, c->stream));
        c->stats.h2d_bytes = (i64)bytes;
        dS = (const double *)c->stage[0].p; d1 = (double *)c->stage[1].p; d2 = (double *)c->stage[2].p;
    }
    if (ny > 65535) return set_err(XINV_E_ARG, "ny > 65535");
    dim3 grid((unsigned)((nx + 127) / 128), (unsigned)ny, (unsigned)batch);
    CK(cudaEventRecord(c->ev0, c->stream));
    xf_flow2d_kernel<<<grid, 128, 0, c->stream>>>(d1, d2, dS, a);
    c->stats.kernel_launches = 1;
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, c->stream));
    if (mem == XINV_MEM_HOST) {
        CK(cudaMemcpyAsync(out1, d1, bytes, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaMemcpyAsync(out2, d2, bytes, cudaMemcpyDeviceToHost, c->stream));
        c->stats.d2h_bytes = (i64)(2 * bytes);
    }
    CK(cudaStreamSynchronize(c->stream));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    c->stats.solve_ms = ms;
    return XINV_OK;
}

// ---------------------------------------------------------------------------
// NCCL (dlopen'ed so the library loads on machines without it)
// ---------------------------------------------------------------------------
typedef struct { char internal[128]; } xnccl_uid;
typedef int (*nccl_get_uid_t)(xnccl_uid *);
typedef int (*nccl_init_rank_t)(void **, int, xnccl_uid, int);
typedef int (*nccl_allreduce_t)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_destroy_t)(void *);
typedef const char *(*nccl_errstr_t)(int);

static void *g_nccl = nullptr;
static void *nccl_handle()
{
    if (g_nccl) return g_nccl;
    const char *names[] = {"libnccl.so.2", "libnccl.so", nullptr};
    const char *env = getenv("XINV_NCCL_LIB");
    if (env) g_nccl = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
    for (int i = 0; !g_nccl && names[i]; ++i) g_nccl = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
    return g_nccl;
}

extern "C" int xinv_nccl_unique_id(void *id128)
{
    if (!id128) return set_err(XINV_E_ARG, "id buffer is NULL");
    void *h = nccl_handle();
    if (!h) return set_err(XINV_E_NCCL, "libnccl.so.2 not found (set XINV_NCCL_LIB): %s", dlerror());
    nccl_get_uid_t f = (nccl_get_uid_t)dlsym(h, "ncclGetUniqueId");
    if (!f) return set_err(XINV_E_NCCL, "ncclGetUniqueId missing");
    xnccl_uid id;
    int r = f(&id);
    if (r) return set_err(XINV_E_NCCL, "ncclGetUniqueId -> %d", r);
    memcpy(id128, &id, 128);
    return XINV_OK;
}

extern "C" int xinv_nccl_init(xinv_ctx *c, const void *id128, int rank, int world)
{
    if (!c || !id128) return set_err(XINV_E_ARG, "NULL argument");
    void *h = nccl_handle();
    if (!h) return set_err(XINV_E_NCCL, "libnccl.so.2 not found (set XINV_NCCL_LIB)");
    nccl_init_rank_t f = (nccl_init_rank_t)dlsym(h, "ncclCommInitRank");
    if (!f) return set_err(XINV_E_NCCL, "ncclCommInitRank missing");
    CK(cudaSetDevice(c->device));
    xnccl_uid id;
    memcpy(&id, id128, 128);
    void *comm = nullptr;
    int r = f(&comm, world, id, rank);
    if (r) return set_err(XINV_E_NCCL, "ncclCommInitRank -> %d", r);
    c->nccl_comm = comm;
    c->nccl_rank = rank;
    c->nccl_world = world;
    int rc = ensure(c->nccl_buf, 64);
    return rc;
}

extern "C" int xinv_nccl_allreduce_active(xinv_ctx *c, int64_t local, int64_t *global_out)
{
    if (!c || !global_out) return set_err(XINV_E_ARG, "NULL argument");
    if (!c->nccl_comm) return set_err(XINV_E_STATE, "xinv_nccl_init not called");
    nccl_allreduce_t f = (nccl_allreduce_t)dlsym(nccl_handle(), "ncclAllReduce");
    if (!f) return set_err(XINV_E_NCCL, "ncclAllReduce missing");
    CK(cudaSetDevice(c->device));
    i64 *d = (i64 *)c->nccl_buf.p;
    CK(cudaMemcpyAsync(d, &local, sizeof(i64), cudaMemcpyHostToDevice, c->stream));
    const int ncclInt64 = 4, ncclSum = 0;
    int r = f(d, d + 1, 1, ncclInt64, ncclSum, c->nccl_comm, c->stream);
    if (r) return set_err(XINV_E_NCCL, "ncclAllReduce -> %d", r);
    i64 out = 0;
    CK(cudaMemcpyAsync(&out, d + 1, sizeof(i64), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    *global_out = out;
    return XINV_OK;
}

extern "C" int xinv_nccl_finalize(xinv_ctx *c)
{
    if (!c) return set_err(XINV_E_ARG, "ctx is NULL");
    if (c->nccl_comm) {
        nccl_destroy_t f = (nccl_destroy_t)dlsym(nccl_handle(), "ncclCommDestroy");
        if (f) f(c->nccl_comm);
        c->nccl_comm = nullptr;
    }
    return XINV_OK;
}
