// tail.cu -- the deep coarse levels of the V-cycle as ONE kernel.
//
// Below ~17^3 a level is a microsecond of work, and the cycle spends its time
// on launch latency: 12 dependent kernels per level even inside a CUDA graph
// (~3.4 us per node measured).  k_coarse_tail runs the whole sub-cycle of levels
// T .. 0 .. T
// (mg_3d.h:1242-1362: zero guess, pre-smoothing, residual, restriction, down to
// the LU solve, then prolongation + correction and post-smoothing back up) in a
// single 1024-thread block: every stage is a strided loop over the level's
// points and a __syncthreads() takes the place of the kernel boundary.  The
// data (a few hundred KB) stays in L1/L2.  Same point formulas, same order of
// operations as the per-stage kernels (devmath.cuh): bit-identical results.
// Measured on B200: worth it up to 17^3 per level (5-level cycle 154 -> 137 us);
// at 33^3 one SM cannot hide the L2 latency any more and the per-stage kernels
// spread over the GPU win, so the default threshold is 5000 points.
#include <cstdio>
#include <cstdlib>

#include "devmath.cuh"
#include "kernels.h"

namespace mgb {

long long *launch_counter();  // kernels.cu

namespace {

constexpr int kTailThreads = 1024;

// one colour of the smoother over the interior (mg_3d.h:432-443, 658-702)
__device__ void t_half_sweep(const TailLevel &L, int colour)
{
    const Geo &g = L.g;
    const int nr = g.nj - 2;
    const long long total = (long long)(g.ni - 2) * nr * g.kh;
    const double *vo = L.u + (long long)(colour ^ 1) * g.cs;
    double *vc = L.u + (long long)colour * g.cs;
    const double *dc = L.d + (long long)colour * g.cs;
    const double sixth = 1. / 6;
    for (long long t = threadIdx.x; t < total; t += kTailThreads) {
        const int m = (int)(t % g.kh);
        const long long row = t / g.kh;
        const int j = 1 + (int)(row % nr);
        const int il = 1 + (int)(row / nr);
        const int kp = (colour ^ (il + j)) & 1;
        const int k = 2 * m + kp;
        if (k < 1 || k > g.nk - 2)
            continue;
        const long long idx = (long long)il * g.pj + (long long)j * g.kh + m;
        vc[idx] = gs_point(vo[idx - g.pj], vo[idx + g.pj], vo[idx - g.kh], vo[idx + g.kh],
                           vo[idx + kp - 1], vo[idx + kp], L.hSq, dc[idx], sixth);
    }
}

// r = d - A v on the interior (mg_3d.h:794-842); the faces of r stay as they are
__device__ void t_residual(const TailLevel &L)
{
    const Geo &g = L.g;
    const int nr = g.nj - 2;
    const long long per = (long long)(g.ni - 2) * nr * g.kh;
    for (long long t = threadIdx.x; t < 2 * per; t += kTailThreads) {
        const int c = t >= per;
        const long long e = t - c * per;
        const int m = (int)(e % g.kh);
        const long long row = e / g.kh;
        const int j = 1 + (int)(row % nr);
        const int il = 1 + (int)(row / nr);
        const int kp = (c ^ (il + j)) & 1;
        const int k = 2 * m + kp;
        if (k < 1 || k > g.nk - 2)
            continue;
        const double *vo = L.u + (long long)(c ^ 1) * g.cs;
        const long long idx = (long long)il * g.pj + (long long)j * g.kh + m;
        L.r[(long long)c * g.cs + idx] =
            res_point(vo[idx - g.pj], vo[idx + g.pj], vo[idx - g.kh], vo[idx + g.kh],
                      vo[idx + kp - 1], vo[idx + kp], L.u[(long long)c * g.cs + idx],
                      L.d[(long long)c * g.cs + idx], L.invHsq);
    }
}

// full-weighting restriction r(fine) -> d(coarse) (mg_3d.h:844-998); coarse
// faces get the injected face residual, which is 0
__device__ void t_restrict(const TailLevel &F, const TailLevel &C)
{
    const Geo &gf = F.g, &gc = C.g;
    const long long per = (long long)gc.ni * gc.pj;
    for (long long t = threadIdx.x; t < 2 * per; t += kTailThreads) {
        const int cc = t >= per;
        const long long e = t - cc * per;
        const int M = (int)(e % gc.kh);
        const long long row = e / gc.kh;
        const int J = (int)(row % gc.nj);
        const int I = (int)(row / gc.nj);
        const int K = 2 * M + ((cc ^ (I + J)) & 1);
        if (K >= gc.nk)
            continue;
        double val = 0.;
        if (!(I == 0 || I == gc.ni - 1 || J == 0 || J == gc.nj - 1 || K == 0 || K == gc.nk - 1)) {
#pragma unroll
            for (int a = 0; a < 3; a++)
#pragma unroll
                for (int b = 0; b < 3; b++)
#pragma unroll
                    for (int c = 0; c < 3; c++) {
                        const int e3 = (a != 1) + (b != 1) + (c != 1);
                        const double w = 1.0 / (double)(8 << e3);
                        val = __dadd_rn(val, __dmul_rn(rd_split(gf, F.r, 2 * I - 1 + a, 2 * J - 1 + b,
                                                               2 * K - 1 + c), w));
                    }
        }
        C.d[(long long)cc * gc.cs + e] = val;
    }
}

// ef += P ec over ALL fine points (mg_3d.h:1000-1145)
__device__ void t_prolong(const TailLevel &C, const TailLevel &F)
{
    const Geo &gc = C.g, &gf = F.g;
    const long long per = (long long)gf.ni * gf.pj;
    for (long long t = threadIdx.x; t < 2 * per; t += kTailThreads) {
        const int c = t >= per;
        const long long e = t - c * per;
        const int m = (int)(e % gf.kh);
        const long long row = e / gf.kh;
        const int j = (int)(row % gf.nj);
        const int i = (int)(row / gf.nj);
        const int k = 2 * m + ((c ^ (i + j)) & 1);
        if (k >= gf.nk)
            continue;
        const int oi = i & 1, oj = j & 1, ok = k & 1;
        const int I = i >> 1, J = j >> 1, K = k >> 1;
        double corr;
        if (!ok) {
            const double a0 = rd_split(gc, C.u, I, J, K);
            const double a1 = oj ? rd_split(gc, C.u, I, J + 1, K) : a0;
            const double b0 = oi ? rd_split(gc, C.u, I + 1, J, K) : a0;
            const double b1 = (oi && oj) ? rd_split(gc, C.u, I + 1, J + 1, K) : a0;
            corr = pc_even(oi, oj, a0, a1, b0, b1);
        } else {
            const double a0x = rd_split(gc, C.u, I, J, K), a0y = rd_split(gc, C.u, I, J, K + 1);
            double a1x = a0x, a1y = a0y, b0x = a0x, b0y = a0y, b1x = a0x, b1y = a0y;
            if (oj) {
                a1x = rd_split(gc, C.u, I, J + 1, K);
                a1y = rd_split(gc, C.u, I, J + 1, K + 1);
            }
            if (oi) {
                b0x = rd_split(gc, C.u, I + 1, J, K);
                b0y = rd_split(gc, C.u, I + 1, J, K + 1);
            }
            if (oi && oj) {
                b1x = rd_split(gc, C.u, I + 1, J + 1, K);
                b1y = rd_split(gc, C.u, I + 1, J + 1, K + 1);
            }
            corr = pc_odd(oi, oj, a0x, a0y, a1x, a1y, b0x, b0y, b1x, b1y);
        }
        double *p = F.u + (long long)c * gf.cs + e;
        *p = __dadd_rn(*p, corr);
    }
}

__device__ void t_zero(const TailLevel &L)
{
    const long long n = 2 * L.g.cs;
    for (long long t = threadIdx.x; t < n; t += kTailThreads)
        L.u[t] = 0.;
}

// phase 1: the down leg (levels top .. 1) and the zero guess of level 0;
// phase 2: the up leg (levels 1 .. top).  In between the caller launches the
// coarsest solve (lu.cu): its warps keep three 32-entry tiles in registers,
// which the 64 registers per thread of a 1024-thread block cannot hold.
__global__ void __launch_bounds__(kTailThreads) k_coarse_tail(const TailP P)
{
    if (P.phase == 1) {
        // down (mg_3d.h:1254-1318)
        for (int q = P.top; q >= 1; q--) {
            const TailLevel &L = P.lv[q];
            if (q < P.top || P.zero_top) {  // coarse levels start from a zero guess
                t_zero(L);
                __syncthreads();
            }
            for (int it = 0; it < P.gs; it++) {  // preSmoother: RED then BLACK
                t_half_sweep(L, 1);
                __syncthreads();
                t_half_sweep(L, 0);
                __syncthreads();
            }
            t_residual(L);
            __syncthreads();
            t_restrict(L, P.lv[q - 1]);
            __syncthreads();
        }
        // level 0 (1262-1277): the solve overwrites every entry of u[0]; pads stay 0
        if (P.top >= 1 || P.zero_top)
            t_zero(P.lv[0]);
        return;
    }
    // up (1331-1351)
    for (int q = 1; q <= P.top; q++) {
        const TailLevel &L = P.lv[q];
        t_prolong(P.lv[q - 1], L);
        __syncthreads();
        for (int it = 0; it < P.gs; it++) {  // postSmoother: BLACK then RED
            t_half_sweep(L, 0);
            __syncthreads();
            t_half_sweep(L, 1);
            __syncthreads();
        }
    }
}

}  // namespace

void launch_coarse_tail(const TailP &p, cudaStream_t st)
{
    TailP q = p;
    q.phase = 1;
    k_coarse_tail<<<1, kTailThreads, 0, st>>>(q);
    ++*launch_counter();
    launch_lu_solve_level(p.lu, p.lv[0].g, p.lv[0].d, p.lv[0].u, st);
    if (p.top >= 1) {
        q.phase = 2;
        k_coarse_tail<<<1, kTailThreads, 0, st>>>(q);
        ++*launch_counter();
    }
}

}  // namespace mgb
