// tail.cu -- the deep coarse levels of the V-cycle as ONE kernel, held in shared memory.
//
// Below ~17^3 a level is well under a microsecond of arithmetic, and the cycle spends
// its time on launch latency (12 dependent kernels per level, ~3.4 us per graph node
// measured) and -- round 1's one-block tail -- on L2 round trips: every stage was
// load -> compute -> store -> barrier against global memory, ~2 us per stage on one SM,
// 89 us for the four levels 17^3 .. 3^3 including the LU solve's own launch.
//
// k_coarse_tail runs the whole sub-cycle of levels T .. 0 .. T (mg_3d.h:1242-1362: zero
// guess, pre-smoothing, residual, restriction, down to the LU solve, then prolongation +
// correction and post-smoothing back up) in a single block whose working set -- u and d of
// every level plus one residual buffer, 192 KB for 17^3 + 9^3 + 5^3 + 3^3 -- lives in
// SHARED MEMORY: the rhs of level T comes in once, the levels' u and d go back to HBM once
// at the end (the prolongation above reads u[T]; tests and downloads see every level).
// A stage is a strided loop over the level's slots with 32-bit index arithmetic and a
// __syncthreads() in place of the kernel boundary.  The coarsest solve runs in the same
// kernel when it fits one 32 x 32 tile (n <= 32: the reference's 3^3 coarse grid): one
// warp, the column sweep of lu_band.cuh on the diagonal tile.  Larger coarse systems keep
// the stand-alone solve kernel between a down-leg and an up-leg launch; tails whose levels
// do not fit shared memory run the same code on the global arrays.
// Same point formulas, same order of operations as the per-stage kernels (devmath.cuh):
// bit-identical results.
#include <cstdio>
#include <cstdlib>

#include "devmath.cuh"
#include "kernels.h"
#include "launch.h"
#include "lu_band.cuh"

namespace mgb {

long long *launch_counter();  // kernels.cu

namespace {

constexpr int kTailThreads = 512;  // 128 registers per thread: the in-kernel solve keeps a tile

// one level as the kernel sees it: int geometry (a tail level has < 2^31 slots by far)
// and pointers into shared or global memory
struct SL {
    int ni, nj, nk, kh, pj, cs;
    double *u, *d;
    double hSq, invHsq;
};

__device__ __forceinline__ double s_rd(const SL &L, const double *a, int i, int j, int k)
{
    const int c = (i + j + k) & 1;
    return a[c * L.cs + i * L.pj + j * L.kh + (k >> 1)];
}

// one colour of the smoother over the interior (mg_3d.h:432-443, 658-702)
__device__ void t_half_sweep(const SL &L, int colour)
{
    const int nr = L.nj - 2;
    const int total = (L.ni - 2) * nr * L.kh;
    const double *vo = L.u + (colour ^ 1) * L.cs;
    double *vc = L.u + colour * L.cs;
    const double *dc = L.d + colour * L.cs;
    const double sixth = 1. / 6;
    for (int t = threadIdx.x; t < total; t += kTailThreads) {
        const int m = t % L.kh;
        const int row = t / L.kh;
        const int j = 1 + row % nr;
        const int il = 1 + row / nr;
        const int kp = (colour ^ (il + j)) & 1;
        const int k = 2 * m + kp;
        if (k < 1 || k > L.nk - 2)
            continue;
        const int idx = il * L.pj + j * L.kh + m;
        vc[idx] = gs_point(vo[idx - L.pj], vo[idx + L.pj], vo[idx - L.kh], vo[idx + L.kh],
                           vo[idx + kp - 1], vo[idx + kp], L.hSq, dc[idx], sixth);
    }
}

// r = d - A v on the interior (mg_3d.h:794-842); the faces of r are never read below
__device__ void t_residual(const SL &L, double *r)
{
    const int nr = L.nj - 2;
    const int per = (L.ni - 2) * nr * L.kh;
    for (int t = threadIdx.x; t < 2 * per; t += kTailThreads) {
        const int c = t >= per;
        const int e = t - c * per;
        const int m = e % L.kh;
        const int row = e / L.kh;
        const int j = 1 + row % nr;
        const int il = 1 + row / nr;
        const int kp = (c ^ (il + j)) & 1;
        const int k = 2 * m + kp;
        if (k < 1 || k > L.nk - 2)
            continue;
        const double *vo = L.u + (c ^ 1) * L.cs;
        const int idx = il * L.pj + j * L.kh + m;
        r[c * L.cs + idx] =
            res_point(vo[idx - L.pj], vo[idx + L.pj], vo[idx - L.kh], vo[idx + L.kh],
                      vo[idx + kp - 1], vo[idx + kp], L.u[c * L.cs + idx], L.d[c * L.cs + idx],
                      L.invHsq);
    }
}

// full-weighting restriction r(fine) -> d(coarse) (mg_3d.h:844-998); coarse faces get the
// injected face residual, which is 0
__device__ void t_restrict(const SL &F, const double *r, const SL &C)
{
    const int per = C.ni * C.pj;
    for (int t = threadIdx.x; t < 2 * per; t += kTailThreads) {
        const int cc = t >= per;
        const int e = t - cc * per;
        const int M = e % C.kh;
        const int row = e / C.kh;
        const int J = row % C.nj;
        const int I = row / C.nj;
        const int K = 2 * M + ((cc ^ (I + J)) & 1);
        if (K >= C.nk)
            continue;
        double val = 0.;
        if (!(I == 0 || I == C.ni - 1 || J == 0 || J == C.nj - 1 || K == 0 || K == C.nk - 1)) {
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const int e3 = (a != 1) + (b != 1) + (c != 1);
                        const double w = 1.0 / (double)(8 << e3);
                        val = __dadd_rn(val, __dmul_rn(s_rd(F, r, 2 * I - 1 + a, 2 * J - 1 + b,
                                                            2 * K - 1 + c), w));
                    }
        }
        C.d[cc * C.cs + e] = val;
    }
}

// ef += P ec over ALL fine points (mg_3d.h:1000-1145)
__device__ void t_prolong(const SL &C, const SL &F)
{
    const int per = F.ni * F.pj;
    for (int t = threadIdx.x; t < 2 * per; t += kTailThreads) {
        const int c = t >= per;
        const int e = t - c * per;
        const int m = e % F.kh;
        const int row = e / F.kh;
        const int j = row % F.nj;
        const int i = row / F.nj;
        const int k = 2 * m + ((c ^ (i + j)) & 1);
        if (k >= F.nk)
            continue;
        const int oi = i & 1, oj = j & 1, ok = k & 1;
        const int I = i >> 1, J = j >> 1, K = k >> 1;
        double corr;
        if (!ok) {
            const double a0 = s_rd(C, C.u, I, J, K);
            const double a1 = oj ? s_rd(C, C.u, I, J + 1, K) : a0;
            const double b0 = oi ? s_rd(C, C.u, I + 1, J, K) : a0;
            const double b1 = (oi && oj) ? s_rd(C, C.u, I + 1, J + 1, K) : a0;
            corr = pc_even(oi, oj, a0, a1, b0, b1);
        } else {
            const double a0x = s_rd(C, C.u, I, J, K), a0y = s_rd(C, C.u, I, J, K + 1);
            double a1x = a0x, a1y = a0y, b0x = a0x, b0y = a0y, b1x = a0x, b1y = a0y;
            if (oj) {
                a1x = s_rd(C, C.u, I, J + 1, K);
                a1y = s_rd(C, C.u, I, J + 1, K + 1);
            }
            if (oi) {
                b0x = s_rd(C, C.u, I + 1, J, K);
                b0y = s_rd(C, C.u, I + 1, J, K + 1);
            }
            if (oi && oj) {
                b1x = s_rd(C, C.u, I + 1, J + 1, K);
                b1y = s_rd(C, C.u, I + 1, J + 1, K + 1);
            }
            corr = pc_odd(oi, oj, a0x, a0y, a1x, a1y, b0x, b0y, b1x, b1y);
        }
        double *p = F.u + c * F.cs + e;
        *p = __dadd_rn(*p, corr);
    }
}

__device__ void t_fill0(double *a, int n)
{
    for (int t = threadIdx.x; t < n; t += kTailThreads)
        a[t] = 0.;
}

// both colours of a level array, global <-> working copy (plane pitch kept, colour stride
// g.cs on the global side, L.cs in the working copy)
__device__ void t_copy_in(const SL &L, double *dst, const double *src, long long gcs)
{
    for (int t = threadIdx.x; t < 2 * L.cs; t += kTailThreads) {
        const int c = t >= L.cs;
        dst[t] = src[c * gcs + (t - c * L.cs)];
    }
}
__device__ void t_copy_out(const SL &L, double *dst, const double *src, long long gcs)
{
    for (int t = threadIdx.x; t < 2 * L.cs; t += kTailThreads) {
        const int c = t >= L.cs;
        dst[c * gcs + (t - c * L.cs)] = src[t];
    }
}

// solveWithLU (gauss_elim.h:31-60) for n <= 32 by ONE warp: the column sweep of
// lu_band.cuh on the single diagonal tile of the tile form (block row 0, tile index nt).
// Row i's forward sum takes its terms in ascending j, its backward sum in descending j,
// both from 0.; above / on the diagonal the L tile holds +0. (below / on: the U tile), and
// adding (+0.)*x to a sum that started at +0. changes nothing.  Quotients by __ddiv_rn.
__device__ void t_lu_small(const LuBand &B, const SL &L0, int lane)
{
    constexpr unsigned FULL = 0xffffffffu;
    const int n = B.n;
    const bool valid = lane < n;
    int il = 0, j = 0, k = 0;
    if (valid) {
        k = lane % L0.nk;
        j = (lane / L0.nk) % L0.nj;
        il = lane / (L0.nk * L0.nj);
    }
    double tri[32];
    lu_tile_load(tri, B.lt + (size_t)B.nt * 1024 + lane);
    const double bi = valid ? s_rd(L0, L0.d, il, j, k) : 0.;
    double sum = 0., z_mine = 0.;
#pragma unroll
    for (int jj = 0; jj < 32; jj++) {
        const double z = __shfl_sync(FULL, __dsub_rn(bi, sum), jj);
        if (lane == jj)
            z_mine = z;
        sum = __dadd_rn(sum, __dmul_rn(tri[jj], z));
    }
    lu_tile_load(tri, B.ut + (size_t)B.nt * 1024 + lane);
    const double di = valid ? B.ud[lane] : 1.;
    double x_mine = 0.;
    sum = 0.;
#pragma unroll
    for (int jj = 31; jj >= 0; jj--) {
        const double x = __shfl_sync(FULL, __ddiv_rn(__dsub_rn(z_mine, sum), di), jj);
        if (lane == jj)
            x_mine = x;
        sum = __dadd_rn(sum, __dmul_rn(tri[jj], x));
    }
    if (valid) {
        const int c = (il + j + k) & 1;
        L0.u[c * L0.cs + il * L0.pj + j * L0.kh + (k >> 1)] = x_mine;
    }
}

// phase 0: the whole sub-cycle (solve included, n <= 32); phase 1: the down leg (levels
// top .. 1) and the zero guess of level 0; phase 2: the up leg (levels 1 .. top), with the
// stand-alone solve kernel (lu.cu) launched in between.
// (min blocks = 1 spelled out: without it ptxas stops at 64 registers and spills the solve's tile)
__global__ void __launch_bounds__(kTailThreads, 1) k_coarse_tail(const TailP P)
{
    pdl_enter();
    extern __shared__ double tail_sh[];
    __shared__ SL lv[8];
    __shared__ double *r_buf;
    if (threadIdx.x == 0) {
        int off = 0;
        for (int q = 0; q <= P.top; q++) {
            const Geo &g = P.lv[q].g;
            SL &L = lv[q];
            L.ni = g.ni, L.nj = g.nj, L.nk = g.nk, L.kh = g.kh, L.pj = (int)g.pj;
            L.hSq = P.lv[q].hSq, L.invHsq = P.lv[q].invHsq;
            if (P.smem) {
                L.cs = g.li * (int)g.pj;
                L.u = tail_sh + off;
                L.d = tail_sh + off + 2 * L.cs;
                off += 4 * L.cs;
            } else {
                L.cs = (int)g.cs;
                L.u = P.lv[q].u;
                L.d = P.lv[q].d;
            }
        }
        r_buf = P.smem ? tail_sh + off : P.lv[P.top].r;
    }
    __syncthreads();
    const int top = P.top;
    const bool down = P.phase != 2, up = P.phase != 1;

    // ---- bring the working set in ----
    if (P.smem) {
        if (down) {
            // everything starts at 0 (zero guesses; pads and faces of every d), then the
            // rhs of the top level -- and its u when it does not start from a zero guess
            int words = 0;
            for (int q = 0; q <= top; q++)
                words += 4 * lv[q].cs;
            t_fill0(tail_sh, words);
            __syncthreads();
            t_copy_in(lv[top], lv[top].d, P.lv[top].d, P.lv[top].g.cs);
            if (!P.zero_top)
                t_copy_in(lv[top], lv[top].u, P.lv[top].u, P.lv[top].g.cs);
        } else {
            for (int q = 0; q <= top; q++) {
                t_copy_in(lv[q], lv[q].u, P.lv[q].u, P.lv[q].g.cs);
                if (q >= 1)
                    t_copy_in(lv[q], lv[q].d, P.lv[q].d, P.lv[q].g.cs);
            }
        }
        __syncthreads();
    }

    if (down) {
        // down (mg_3d.h:1254-1318)
        for (int q = top; q >= 1; q--) {
            const SL &L = lv[q];
            if (!P.smem && (q < top || P.zero_top)) {  // coarse levels start from a zero guess
                t_fill0(L.u, 2 * L.cs);
                __syncthreads();
            }
            for (int it = 0; it < P.gs; it++) {  // preSmoother: RED then BLACK
                t_half_sweep(L, 1);
                __syncthreads();
                t_half_sweep(L, 0);
                __syncthreads();
            }
            t_residual(L, r_buf);
            __syncthreads();
            t_restrict(L, r_buf, lv[q - 1]);
            __syncthreads();
        }
        // level 0 (1262-1277): the solve overwrites every entry of u[0]; pads stay 0
        if (!P.smem && (top >= 1 || P.zero_top)) {
            t_fill0(lv[0].u, 2 * lv[0].cs);
            __syncthreads();
        }
    }
    if (P.phase == 0) {
        if (threadIdx.x < 32)
            t_lu_small(P.lu, lv[0], threadIdx.x);
        __syncthreads();
    }
    if (up) {
        // up (1331-1351)
        for (int q = 1; q <= top; q++) {
            const SL &L = lv[q];
            t_prolong(lv[q - 1], L);
            __syncthreads();
            for (int it = 0; it < P.gs; it++) {  // postSmoother: BLACK then RED
                t_half_sweep(L, 0);
                __syncthreads();
                t_half_sweep(L, 1);
                __syncthreads();
            }
        }
    }

    // ---- and back out: every level's u; the d this kernel produced ----
    if (P.smem) {
        for (int q = 0; q <= top; q++) {
            t_copy_out(lv[q], P.lv[q].u, lv[q].u, P.lv[q].g.cs);
            if (down && q < top)
                t_copy_out(lv[q], P.lv[q].d, lv[q].d, P.lv[q].g.cs);
        }
    }
}

// bytes of shared memory the working set of levels 0..top needs (u, d per level + one
// residual buffer of the top level's size)
size_t tail_smem_bytes(const TailP &p)
{
    size_t words = 0;
    for (int q = 0; q <= p.top; q++)
        words += 4 * (size_t)p.lv[q].g.li * (size_t)p.lv[q].g.pj;
    words += 2 * (size_t)p.lv[p.top].g.li * (size_t)p.lv[p.top].g.pj;
    return words * sizeof(double);
}

}  // namespace

void launch_coarse_tail(const TailP &p, cudaStream_t st)
{
    static const bool no_smem = getenv("MGB_TAIL_SMEM") && atoi(getenv("MGB_TAIL_SMEM")) == 0;
    static int smem_limit = -1;
    if (smem_limit < 0) {
        int dev = 0, lim = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&lim, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
        smem_limit = lim - 1024;  // the kernel's own static shared memory
    }
    TailP q = p;
    bool whole = true;  // every level stored whole and unshifted on this rank
    for (int l = 0; l <= p.top; l++)
        whole = whole && p.lv[l].g.i0 == 0 && p.lv[l].g.li == p.lv[l].g.ni;
    const size_t need = tail_smem_bytes(p);
    q.smem = whole && !no_smem && need <= (size_t)smem_limit;
    const size_t sh = q.smem ? need : 0;
    static size_t allowed[64] = {};
    if (grows_on_device(allowed, sh))
        cudaFuncSetAttribute(k_coarse_tail, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    if (p.lu.n <= 32 && whole) {
        q.phase = 0;
        launch_k(k_coarse_tail, 1, kTailThreads, sh, st, q);
        ++*launch_counter();
        return;
    }
    q.phase = 1;
    launch_k(k_coarse_tail, 1, kTailThreads, sh, st, q);
    ++*launch_counter();
    launch_lu_solve_level(p.lu, p.lv[0].g, p.lv[0].d, p.lv[0].u, st);
    if (p.top >= 1) {
        q.phase = 2;
        launch_k(k_coarse_tail, 1, kTailThreads, sh, st, q);
        ++*launch_counter();
    }
}

}  // namespace mgb
