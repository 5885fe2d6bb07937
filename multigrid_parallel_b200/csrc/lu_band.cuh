// lu_band.cuh -- solveWithLU (gauss_elim.h:31-60) on the factorised coarse
// operator, restricted to its BAND and stored in 32 x 32 tiles, as a
// warp-cooperative device routine shared by the stand-alone solve kernel (lu.cu)
// and the one-kernel coarse tail (tail.cu).
//
// Why a band.  The coarse operator couples p with p +- nj*nk, p +- nk, p +- 1
// (mg_3d.h:257-268), so A -- and, without pivoting, L and U -- have half
// bandwidth bw = nj*nk.  The reference's dense loops add LU[i][j]*x[j] for
// every j; outside the band that product is (+-0)*x[j] = +-0, and a running sum
// that started at +0. can never be -0 (an exact cancellation gives +0 in
// round-to-nearest), so adding +-0 never changes it: skipping the out-of-band
// terms -- or adding a few of them, as the tiles do -- is bit-exact.
//
// Bit-exactness contract kept inside the band: row i's forward sum runs over
// ASCENDING j, its backward sum over DESCENDING j, each starting from 0.;
// x[i] = b[i] - sum, then x[i] = (x[i] - sum) / U[i][i], correctly rounded.
//
// Parallel structure.  Rows are cut into blocks of 32; lane l of the warp that
// owns block R holds row i = 32R + l.  The factor is stored per block row as
// NT+1 tiles (NT = ceil(bw/32) column blocks to the left -- for U: to the
// right -- plus the diagonal block), each tile TRANSPOSED: entry (jj, l) =
// coefficient of column jj of that column block in row l.  One step of any
// tile is then "all lanes take column jj": one coalesced 256-byte load, the
// coefficient lives in a register with a compile-time index, x_j is a
// shared-memory broadcast.  The diagonal tile is the sequential part: finish
// x_j (lane j), shuffle it to everybody, one multiply-add per lane.  kLuWarps
// warps take the blocks round-robin and hand finished blocks over through a
// counter in shared memory (release / acquire at CTA scope); a warp has its
// last off-diagonal tile and its diagonal tile in registers BEFORE it waits for
// the previous block, so the critical path per block is 32 multiply-adds plus
// the 32 steps of the triangle.
//
// The division of the backward sweep sits on that critical path.  The fast form
// multiplies by the correctly rounded reciprocal and corrects once with the
// exact remainder (q0 = t*y, q = fma(fma(-d, q0, t), y, q0)): that is RN(t/d)
// unless t/d lies within ~2^-105 (relative) of a rounding boundary.  EVERY
// quotient is checked afterwards, off the critical path, against __ddiv_rn; a
// single mismatch makes the caller repeat the whole solve with __ddiv_rn on the
// path.  The result is therefore always the correctly rounded one.
#pragma once
#include <cstddef>

namespace mgb {

struct LuBand {
    double *lt;  // L tiles: lt[((R*(nt+1) + t)*32 + jj)*32 + l] = L[32R+l][32*(R-nt+t)+jj]
    double *ut;  // U tiles: ut[((R*(nt+1) + t)*32 + jj)*32 + l] = U[32R+l][32*(R+nt-t)+jj]
                 //   (t = nt: the diagonal block, strictly lower / strictly upper part)
    double *ud;  // ud[i] = U[i][i]
    double *rd;  // rd[i] = RN(1 / U[i][i])
    int n, bw, nt;
};

constexpr int kLuWarps = 8;  // warps that take part in a solve
__host__ __device__ inline int lu_num_tiles(int bw) { return (bw + 31) / 32; }
// doubles in each of the two tile arrays
__host__ __device__ inline size_t lu_tile_doubles(int n, int bw)
{
    return (size_t)((n + 31) / 32) * (size_t)(lu_num_tiles(bw) + 1) * 1024;
}
// shared memory a solve needs, in doubles: xs[n rounded up to 32] + flags
__host__ __device__ inline size_t lu_solve_smem_doubles(int n)
{
    return (size_t)((n + 31) & ~31) + 2;
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

// one off-diagonal tile: sum += c[jj] * x[jj] for the 32 columns of a finished
// block, ascending (forward) or descending (backward) jj
template <bool DESCENDING>
__device__ __forceinline__ double lu_tile_apply(const double (&c)[32], const double *xp, double sum)
{
#pragma unroll
    for (int s = 0; s < 32; s++) {
        const int jj = DESCENDING ? 31 - s : s;
        sum = __dadd_rn(sum, __dmul_rn(c[jj], xp[jj]));
    }
    return sum;
}
__device__ __forceinline__ void lu_tile_load(double (&c)[32], const double *tile)
{
#pragma unroll
    for (int jj = 0; jj < 32; jj++)
        c[jj] = tile[jj * 32];
}

// Called by the first kLuWarps warps of a block (all 32 lanes each), w = warp
// index.  On entry xs[0..n) holds b, xs[n..npad) = 0, flags[0..3] = 0, and a
// __syncthreads() has made that visible; on exit (after the caller's next
// __syncthreads()) xs[0..n) holds x.  Returns false on the lanes that saw a
// fast quotient differ from __ddiv_rn (EXACT = false only): the caller then
// repeats the solve with EXACT = true.
template <bool EXACT>
__device__ __forceinline__ bool lu_band_solve(const LuBand &B, double *xs, int *flags, int w,
                                              int lane)
{
    constexpr int W = kLuWarps;
    constexpr unsigned FULL = 0xffffffffu;
    const int n = B.n, NT = B.nt;
    const int NB = (n + 31) >> 5;
    int *done_f = flags, *done_b = flags + 1;
    bool ok = true;

    // ---------------- forward: L z = b, unit lower triangle ----------------
    for (int R = w; R < NB; R += W) {
        const int i = (R << 5) + lane;
        const bool valid = i < n;
        const double *tiles = B.lt + (size_t)R * (NT + 1) * 1024 + lane;
        double tri[32], sq[32];
        lu_tile_load(tri, tiles + (size_t)NT * 1024);
        const bool has_sq = NT >= 1 && R >= 1;
        if (has_sq)
            lu_tile_load(sq, tiles + (size_t)(NT - 1) * 1024);
        double sum = 0.;
        for (int t = 0; t < NT - 1; t++) {  // the far column blocks, oldest first
            const int P = R - NT + t;
            if (P < 0)
                continue;
            double c[32];
            lu_tile_load(c, tiles + (size_t)t * 1024);
            lu_wait(done_f, P + 1);
            sum = lu_tile_apply<false>(c, xs + (P << 5), sum);
        }
        if (R > 0)
            lu_wait(done_f, R);  // block R-1 is there (and blocks are published in order)
        if (has_sq)
            sum = lu_tile_apply<false>(sq, xs + ((R - 1) << 5), sum);
        const double bi = valid ? xs[i] : 0.;
        double mine = 0.;
#pragma unroll
        for (int jj = 0; jj < 32; jj++) {
            // lane jj has all of its row's terms: z_jj = b - sum; the tile holds +0. on and
            // above the diagonal, so the lanes <= jj add (+0.)*z: nothing
            const double z = __shfl_sync(FULL, __dsub_rn(bi, sum), jj);
            if (lane == jj)
                mine = z;
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
        const int i = (R << 5) + lane;
        const bool valid = i < n;
        const double *tiles = B.ut + (size_t)R * (NT + 1) * 1024 + lane;
        double tri[32], sq[32];
        lu_tile_load(tri, tiles + (size_t)NT * 1024);
        const bool has_sq = NT >= 1 && R + 1 < NB;
        if (has_sq)
            lu_tile_load(sq, tiles + (size_t)(NT - 1) * 1024);
        const double di = valid ? B.ud[i] : 1.;
        const double yi = valid ? B.rd[i] : 1.;
        double sum = 0.;
        for (int t = 0; t < NT - 1; t++) {  // the far column blocks, highest columns first
            const int P = R + NT - t;
            if (P >= NB)
                continue;
            double c[32];
            lu_tile_load(c, tiles + (size_t)t * 1024);
            lu_wait(done_b, NB - P);
            sum = lu_tile_apply<true>(c, xs + (P << 5), sum);
        }
        if (Rr > 0)
            lu_wait(done_b, Rr);
        if (has_sq)
            sum = lu_tile_apply<true>(sq, xs + ((R + 1) << 5), sum);
        const double zi = valid ? xs[i] : 0.;
        double mine = 0., tm = 0.;
#pragma unroll
        for (int jj = 31; jj >= 0; jj--) {
            const double t = __dsub_rn(zi, sum);
            double q;
            if (EXACT) {
                q = __ddiv_rn(t, di);
            } else {
                const double q0 = __dmul_rn(t, yi);
                q = __fma_rn(__fma_rn(-di, q0, t), yi, q0);
            }
            const double x = __shfl_sync(FULL, q, jj);
            if (lane == jj) {
                mine = x;
                tm = t;
            }
            sum = __dadd_rn(sum, __dmul_rn(tri[jj], x));
        }
        if (valid)
            xs[i] = mine;
        __syncwarp();
        if (lane == 0)
            lu_st_release(done_b, Rr + 1);
        if (!EXACT && __double_as_longlong(__ddiv_rn(tm, di)) != __double_as_longlong(mine))
            ok = false;
    }
    return ok;
}

}  // namespace mgb
