// lu_band.cuh -- solveWithLU (gauss_elim.h:31-60) on the factorised coarse
// operator kept in BAND form, as a warp-cooperative device routine shared by
// the stand-alone solve kernel (lu.cu) and the one-kernel coarse tail (tail.cu).
//
// Why a band.  The coarse operator couples p with p +- nj*nk, p +- nk, p +- 1
// (mg_3d.h:257-268), so A -- and, without pivoting, L and U -- have half
// bandwidth bw = nj*nk.  The reference's dense loops add LU[i][j]*x[j] for
// every j; outside the band that product is (+-0)*x[j] = +-0, and a running sum
// that started at +0. can never be -0 (an exact cancellation gives +0 in
// round-to-nearest), so adding +-0 never changes it: skipping the out-of-band
// terms is bit-exact.
//
// Bit-exactness contract kept inside the band: row i's forward sum runs over
// ASCENDING j, its backward sum over DESCENDING j, each starting from 0.;
// x[i] = b[i] - sum, then x[i] = (x[i] - sum) / U[i][i] with a correctly
// rounded division.
//
// Parallel structure.  Rows are cut into blocks of 32; lane l of the warp that
// owns block R holds row i = 32R + l.  Everything a row needs from columns of
// EARLIER blocks (the "rectangle") is accumulated with a uniform distance
// d = i - j per step, so every load of the band arrays is coalesced; columns
// inside the block (the "triangle") are the sequential part: 32 steps of
// finish x_j -> shuffle -> one multiply-add per lane.  W warps take the blocks
// round-robin and hand finished blocks over through a counter in shared
// memory (release / acquire at CTA scope), so the rectangle work and the band
// loads of the next blocks overlap the triangle of the current one: the
// critical path is one triangle + the last 63 rectangle steps per block.
#pragma once
#include <cstddef>

namespace mgb {

struct LuBand {
    double *lb;  // lb[(d-1)*n + i] = L[i][i-d], d = 1..bw (0 where i-d < 0)
    double *ub;  // ub[(d-1)*n + i] = U[i][i+d], d = 1..bw (0 where i+d >= n)
    double *ud;  // ud[i] = U[i][i]
    int n, bw;
};

constexpr int kLuWarps = 4;                  // warps that take part in a solve
constexpr int kLuTriDoubles = 32 * 33;       // one triangle tile per warp
// shared memory a solve needs, in doubles: xs[n rounded up to 32] + tiles + flags
__host__ __device__ inline size_t lu_solve_smem_doubles(int n)
{
    return (size_t)((n + 31) & ~31) + (size_t)kLuWarps * kLuTriDoubles + 2;
}

__device__ __forceinline__ int lu_ld_acquire(const int *p)
{
    int v;
    asm volatile("ld.acquire.cta.shared.b32 %0, [%1];"
                 : "=r"(v)
                 : "r"((unsigned)__cvta_generic_to_shared(p))
                 : "memory");
    return v;
}
__device__ __forceinline__ void lu_st_release(int *p, int v)
{
    asm volatile("st.release.cta.shared.b32 [%0], %1;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(p)),
                 "r"(v)
                 : "memory");
}
__device__ __forceinline__ void lu_wait(const int *flag, int need)
{
    while (lu_ld_acquire(flag) < need) {
    }
}

// Called by the first kLuWarps warps of a block (all 32 lanes each), w = warp
// index.  On entry xs[0..n) holds b, xs[n..npad) = 0, flags[0] = flags[1] = 0,
// and a __syncthreads() has made that visible; on exit (after the caller's next
// __syncthreads()) xs[0..n) holds x.
__device__ __forceinline__ void lu_band_solve(const LuBand &B, double *xs, double *tri_all,
                                              int *flags, int w, int lane)
{
    constexpr int W = kLuWarps;
    constexpr unsigned FULL = 0xffffffffu;
    const int n = B.n, bw = B.bw;
    const int NB = (n + 31) >> 5;
    double *tri = tri_all + w * kLuTriDoubles + lane * 33;  // this lane's tile row
    int *done_f = flags, *done_b = flags + 1;

    // ---------------- forward: L z = b, unit lower triangle ----------------
    for (int R = w; R < NB; R += W) {
        const int i0 = R << 5, i = i0 + lane;
        const bool valid = i < n;
        // triangle tile: tri[jj] = L[i][i0+jj] for jj < lane (distance lane-jj)
#pragma unroll 4
        for (int d = 1; d < 32; d++)
            if (d <= lane)
                tri[lane - d] = (valid && d <= bw) ? B.lb[(size_t)(d - 1) * n + i] : 0.;
        double sum = 0.;
        // rectangle: columns j = i - d < i0, ascending j = descending d
        const int dmax = min(bw, i0 + 31);
        for (int dc = dmax; dc >= 1; dc -= 16) {
            double l[16];
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const int d = dc - t;
                const bool act = valid && d >= 1 && d > lane && d <= i;
                l[t] = act ? B.lb[(size_t)(d - 1) * n + i] : 0.;
            }
            // the newest column this chunk touches
            const int dlow = dc - 15 > 1 ? dc - 15 : 1;
            int jmax = i0 + 31 - dlow;
            if (jmax > i0 - 1)
                jmax = i0 - 1;
            if (jmax >= 0)
                lu_wait(done_f, (jmax >> 5) + 1);
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const int d = dc - t;
                const bool act = valid && d >= 1 && d > lane && d <= i;
                if (act)
                    sum = __dadd_rn(sum, __dmul_rn(l[t], xs[i - d]));
            }
        }
        if (R > 0)
            lu_wait(done_f, R);  // blocks are published in order
        __syncwarp();
        const double bi = valid ? xs[i] : 0.;
        double mine = 0.;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) {
            const double z = __shfl_sync(FULL, __dsub_rn(bi, sum), jj);
            if (lane == jj)
                mine = z;
            if (lane > jj)
                sum = __dadd_rn(sum, __dmul_rn(tri[jj], z));
        }
        if (valid)
            xs[i] = mine;
        __syncwarp();
        if (lane == 0)
            lu_st_release(done_f, R + 1);
    }

    // ---------------- backward: U x = z ----------------
    lu_wait(done_f, NB);
    for (int Rr = w; Rr < NB; Rr += W) {
        const int R = NB - 1 - Rr;  // blocks from the end
        const int i0 = R << 5, i = i0 + lane;
        const bool valid = i < n;
        __syncwarp();
        // triangle tile: tri[jj] = U[i][i0+jj] for jj > lane (distance jj-lane)
#pragma unroll 4
        for (int d = 1; d < 32; d++)
            if (lane + d < 32)
                tri[lane + d] = (valid && d <= bw && i + d < n) ? B.ub[(size_t)(d - 1) * n + i] : 0.;
        double sum = 0.;
        // rectangle: columns j = i + d > i0 + 31, descending j = descending d
        const int dmax = min(bw, n - 1 - i0);
        for (int dc = dmax; dc >= 1; dc -= 16) {
            double l[16];
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const int d = dc - t;
                const bool act = valid && d >= 1 && d > 31 - lane && i + d < n;
                l[t] = act ? B.ub[(size_t)(d - 1) * n + i] : 0.;
            }
            // the oldest (lowest) column this chunk touches: i0 + max(dlow, 32)
            const int dlow = dc - 15 > 1 ? dc - 15 : 1;
            const int jmin = i0 + (dlow > 32 ? dlow : 32);
            if (jmin < n)
                lu_wait(done_b, NB - (jmin >> 5));
#pragma unroll
            for (int t = 0; t < 16; t++) {
                const int d = dc - t;
                const bool act = valid && d >= 1 && d > 31 - lane && i + d < n;
                if (act)
                    sum = __dadd_rn(sum, __dmul_rn(l[t], xs[i + d]));
            }
        }
        if (Rr > 0)
            lu_wait(done_b, Rr);
        __syncwarp();
        const double zi = valid ? xs[i] : 0.;
        const double di = valid ? B.ud[i] : 1.;
        double mine = 0.;
#pragma unroll
        for (int jj = 31; jj >= 0; jj--) {
            const double x = __shfl_sync(FULL, __ddiv_rn(__dsub_rn(zi, sum), di), jj);
            if (lane == jj)
                mine = x;
            if (lane < jj)
                sum = __dadd_rn(sum, __dmul_rn(tri[jj], x));
        }
        if (valid)
            xs[i] = mine;
        __syncwarp();
        if (lane == 0)
            lu_st_release(done_b, Rr + 1);
    }
}

}  // namespace mgb
