// devmath.cuh -- the reference's point formulas as device functions, shared by
// kernels.cu and tail.cu.  Explicitly rounded intrinsics in the reference's
// operation order: no FMA contraction whatever the compiler flags.
#pragma once
#include "geom.h"

namespace mgb {

// mg_3d.h:437-442: multFact*(v[p-NN]+v[p+NN]+v[p-N]+v[p+N]+v[p-1]+v[p+1]-hSq*d[p])
__device__ __forceinline__ double gs_point(double im, double ip, double jm,
                                           double jp, double km, double kp,
                                           double hSq, double d, double sixth)
{
    double s = __dadd_rn(im, ip);
    s = __dadd_rn(s, jm);
    s = __dadd_rn(s, jp);
    s = __dadd_rn(s, km);
    s = __dadd_rn(s, kp);
    s = __dsub_rn(s, __dmul_rn(hSq, d));
    return __dmul_rn(sixth, s);
}

// mg_3d.h:818-820: d[p] - invHsq*(v[p-NN]+v[p+NN]+v[p-N]+v[p+N]+v[p-1]+v[p+1]-6*v[p])
__device__ __forceinline__ double res_point(double im, double ip, double jm,
                                            double jp, double km, double kp,
                                            double vc, double d, double invHsq)
{
    double s = __dadd_rn(im, ip);
    s = __dadd_rn(s, jm);
    s = __dadd_rn(s, jp);
    s = __dadd_rn(s, km);
    s = __dadd_rn(s, kp);
    s = __dsub_rn(s, __dmul_rn(6.0, vc));
    return __dsub_rn(d, __dmul_rn(invHsq, s));
}

// value of a colour-split array at local plane il, row j, column k
__device__ __forceinline__ double rd_split(const Geo &g, const double *a, int il,
                                           int j, int k)
{
    const int c = (g.i0 + il + j + k) & 1;
    return a[c * g.cs + ((long long)il * g.nj + j) * g.kh + (k >> 1)];
}

__device__ __forceinline__ double add0(double x) { return __dadd_rn(0., x); }

// fine point with even k on coarse column x (ok = 0)
__device__ __forceinline__ double pc_even(int oi, int oj, double a0, double a1, double b0,
                                          double b1)
{
    if (!oi && !oj)
        return a0;  // 1137-1138
    if (oi && !oj)
        return __dmul_rn(__dadd_rn(add0(a0), b0), 0.5);  // 1105-1111
    if (!oi)
        return __dmul_rn(__dadd_rn(add0(a0), a1), 0.5);  // 1112-1118
    // 1080-1089: j fastest, then i
    double t = __dadd_rn(add0(a0), a1);
    t = __dadd_rn(t, b0);
    t = __dadd_rn(t, b1);
    return __dmul_rn(t, 0.25);
}

// fine point with odd k between coarse columns x (low) and y (high) (ok = 1)
__device__ __forceinline__ double pc_odd(int oi, int oj, double a0x, double a0y, double a1x,
                                         double a1y, double b0x, double b0y, double b1x,
                                         double b1y)
{
    if (!oi && !oj)
        return __dmul_rn(__dadd_rn(add0(a0x), a0y), 0.5);  // 1119-1125
    if (oi && !oj) {  // 1070-1079: i fastest, then k
        double t = __dadd_rn(add0(a0x), b0x);
        t = __dadd_rn(t, a0y);
        t = __dadd_rn(t, b0y);
        return __dmul_rn(t, 0.25);
    }
    if (!oi) {  // 1059-1068: j fastest, then k
        double t = __dadd_rn(add0(a0x), a1x);
        t = __dadd_rn(t, a0y);
        t = __dadd_rn(t, a1y);
        return __dmul_rn(t, 0.25);
    }
    // 1023-1049: i-major, then j, then k
    double t = __dadd_rn(add0(a0x), a0y);
    t = __dadd_rn(t, a1x);
    t = __dadd_rn(t, a1y);
    t = __dadd_rn(t, b0x);
    t = __dadd_rn(t, b0y);
    t = __dadd_rn(t, b1x);
    t = __dadd_rn(t, b1y);
    return __dmul_rn(t, 0.125);
}


}  // namespace mgb
