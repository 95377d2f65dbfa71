// xinv_fused2d.cuh -- TMA-staged fused red+black iteration engine (2-D, B == 0).
// (stub in this revision: the colour engine runs everything)
#pragma once
#include <string>
#include "xinv_device.cuh"

struct FusedPlan {
    bool built = false;
    int nblk_partials = 0;
};

static inline void fused_plan_release(FusedPlan &p) { p = FusedPlan(); }

static inline bool fused_plan_supported(int, bool, const XdGeom &, std::string &why)
{
    why = "fused engine not built in this revision";
    return false;
}

static inline int fused_plan_build(FusedPlan &, int, int, const XdGeom &, const XdCoef &, i64, double *, double *,
                                   cudaStream_t, std::string &why)
{
    why = "fused engine not built in this revision";
    return -1;
}

static inline int fused_sweep(FusedPlan &, cudaStream_t, XdSliceState *, double *, i64 *, int *, double, i64, int,
                              int64_t *)
{
    return -1;
}
