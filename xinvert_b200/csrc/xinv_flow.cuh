// xinv_flow.cuh -- the epilogue of the inversion: flow components from the inverted field
// (apps.cal_flow, apps.py:1181-1317), one pass over psi on the device.
//
// Both branches of the reference are centred differences of the field along the two core
// dimensions followed by a per-row combination:
//   * 'GillMatsuno' (apps.py:1277-1317): DataArray.differentiate == numpy.gradient with
//     edge_order 1 (one-sided differences at the two ends), then
//         u = -coef1 dphi/dx [/ deg2m / cosLat] - coef2 dphi/dy [/ deg2m]
//         v = -coef1 dphi/dy [/ deg2m]          + coef2 dphi/dx [/ deg2m / cosLat]
//   * 'streamfunction' / 'velocitypotential' (apps.py:1207-1271; finitediffs.py:151-207,
//     :548-659): the field is padded with its boundary condition (fixed / extend / reflect /
//     periodic), differentiated, trimmed and divided by the metric of the dimension.
// numpy.gradient uses (f[i+1] - f[i-1]) / (2 dx) when the coordinate is exactly uniform and
// a f[i-1] + b f[i] + c f[i+1] otherwise; the host decides which (exactly as numpy does, on
// the same coordinate values) and hands over 2 dx or the three weight vectors, so the
// device performs the same IEEE operations in the same order as the numpy expression.
#pragma once
#include "xinv_device.cuh"

#define XF_EDGE_ONESIDED 0   // no padding: one-sided differences at the ends (numpy.gradient, edge_order = 1)
#define XF_EDGE_FIXED    1   // padded with a fill value
#define XF_EDGE_EXTEND   2   // padded with the edge value
#define XF_EDGE_REFLECT  3   // padded with the first inner value
#define XF_EDGE_PERIODIC 4   // padded with the value from the other end

#define XF_COMB_GRAD     0   // out1 = s1 * (dS/dy / my[j]),  out2 = s2 * (dS/dx / mx[j]) ; swap selects (x, y) order
#define XF_COMB_GM_LL    1   // Gill-Matsuno, lat-lon
#define XF_COMB_GM_CART  2   // Gill-Matsuno, cartesian

struct XfAxis {
    int uniform, edge;
    double den;              // uniform: 2 dx
    double lo, hi;           // ONESIDED: spacing at the two ends; FIXED: the two fill values
    const double *w;         // non-uniform: [3][n] = a, b, c of every output index
};

struct XfArgs {
    XfAxis y, x;
    int comb, swap;
    double s1, s2;           // GRAD: signs (+1 / -1)
    double deg2m;            // GM_LL
    const double *rows;      // GRAD: [2][ny] = metric of y, metric of x per row; GM: [3][ny] = coef1, coef2, cosLat
    i64 ny, nx;
};

// centred difference of numpy.gradient at index i of a line of n values read through get(i)
template <typename Get>
__device__ __forceinline__ double xf_diff(const XfAxis &ax, i64 i, i64 n, Get get)
{
    if (ax.edge == XF_EDGE_ONESIDED) {
        if (i == 0) return (get(1) - get(0)) / ax.lo;
        if (i == n - 1) return (get(n - 1) - get(n - 2)) / ax.hi;
    }
    double fm, fp;
    if (i > 0) fm = get(i - 1);
    else fm = (ax.edge == XF_EDGE_FIXED) ? ax.lo : (ax.edge == XF_EDGE_EXTEND) ? get(0) : (ax.edge == XF_EDGE_REFLECT) ? get(1) : get(n - 1);
    if (i < n - 1) fp = get(i + 1);
    else fp = (ax.edge == XF_EDGE_FIXED) ? ax.hi : (ax.edge == XF_EDGE_EXTEND) ? get(n - 1) : (ax.edge == XF_EDGE_REFLECT) ? get(n - 2) : get(0);
    if (ax.uniform) return (fp - fm) / ax.den;
    return (ax.w[i] * fm + ax.w[n + i] * get(i)) + ax.w[2 * n + i] * fp;
}

__global__ void xf_flow2d_kernel(double *__restrict__ o1, double *__restrict__ o2, const double *__restrict__ S, XfArgs a)
{
    const i64 i = (i64)blockIdx.x * blockDim.x + threadIdx.x;
    const i64 j = blockIdx.y;
    if (i >= a.nx) return;
    const i64 nx = a.nx, ny = a.ny;
    {
        const i64 b = blockIdx.z;
        const double *P = S + b * ny * nx;
        const double dy = xf_diff(a.y, j, ny, [&](i64 q) { return P[q * nx + i]; });
        const double dx = xf_diff(a.x, i, nx, [&](i64 q) { return P[j * nx + q]; });
        double r1, r2;
        if (a.comb == XF_COMB_GRAD) {
            const double gy = dy / a.rows[j], gx = dx / a.rows[ny + j];
            r1 = a.swap ? gx : gy;
            r2 = a.swap ? gy : gx;
            if (a.s1 < 0) r1 = -r1;
            if (a.s2 < 0) r2 = -r2;
        } else {
            const double c1 = a.rows[j], c2 = a.rows[ny + j];
            if (a.comb == XF_COMB_GM_LL) {
                const double cl = a.rows[2 * ny + j];
                r1 = (((-c1) * dx) / a.deg2m) / cl - (c2 * dy) / a.deg2m;
                r2 = ((-c1) * dy) / a.deg2m + ((c2 * dx) / a.deg2m) / cl;
            } else {
                r1 = (-c1) * dx - c2 * dy;
                r2 = (-c1) * dy + c2 * dx;
            }
        }
        o1[b * ny * nx + j * nx + i] = r1;
        o2[b * ny * nx + j * nx + i] = r2;
    }
}
