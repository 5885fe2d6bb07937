// gslex.cu -- GaussSeidelSmoother (mg_3d.h:546-637): the LEXICOGRAPHIC Gauss-Seidel sweep
// on the GPU, bit-identical to the reference's serial triple loop.
//
// The serial sweep visits (i,j,k) in ascending order, so the update of a point reads the
// NEW values of (i-1,j,k), (i,j-1,k), (i,j,k-1) and the OLD values of the three "+1"
// neighbours.  All points of a hyperplane i+j+k = h depend only on hyperplanes h-1 (new)
// and h+1 (old) and not on each other, so sweeping the hyperplanes in ascending h with all
// points of a hyperplane in parallel reproduces the serial result exactly: each point sees
// precisely the values the serial loop would have shown it, and evaluates the same
// expression in the same order (gs_point, devmath.cuh).  The same argument holds one level
// up, for TILES: tile (a,b,c) needs the tiles (a-1,b,c), (a,b-1,c), (a,b,c-1) of the same
// sweep finished and the tiles (a+1,b,c), (a,b+1,c), (a,b,c+1) still at the previous sweep.
//
// Two kernels:
//
// k_gs_lex_tile (default).  A block takes a tile of TI x TJ x TK interior points, brings it
//   and its one-point halo into SHARED memory (cp.async after an acquire of the producers'
//   flags: the halo was written by other blocks of the same launch), the right-hand side
//   pre-multiplied by h^2 next to it, and runs the hyperplane wavefront INSIDE the tile with
//   __syncthreads() between hyperplanes (thread (i,j) owns the line of TK points along k and
//   is active on TK consecutive steps); then it stores the tile and publishes "tile done,
//   sweep s" with a release store.
//   Tiles are handed out through a ticket counter in the order of a + b + c + 2 s, a
//   topological order of the dependences above (sweep s+1 follows sweep s two tile
//   hyperplanes behind), so a block only ever waits for tiles that are already running:
//   no co-residency requirement, no grid-wide barrier, and `iters` sweeps pipeline through
//   one launch.  Global traffic: every value is read ~1.4 times and written once per sweep.
//   513^3: 2.6 ms per sweep (the serial reference loop: ~0.5 s); bound by the latency of a
//   tile (~20 us: flags, load, 62 hyperplanes, store) along the ~80 tile hyperplanes of the
//   grid and by the same latency times 16 k tiles over 148 SMs, not by HBM.
//
// k_gs_lex (small grids, MGB_GSLEX_TILE=0).  Global hyperplanes: one persistent cooperative
//   kernel, one grid-wide barrier per hyperplane, the colour-split arrays read through L2.
//   Latency-bound (3N dependent steps of ~4-6 us) and with poor locality (a hyperplane
//   touches one 32-byte sector per point): 9.6 ms per sweep at 513^3 -- what the tile kernel
//   replaces.
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "devmath.cuh"
#include "kernels.h"
#include "launch.h"

namespace mgb {

long long *launch_counter();  // kernels.cu

namespace {

struct LexP {
    Geo g;
    double *v;
    const double *d;
    double hSq;
    int iters;
    unsigned int *bar;  // zeroed arrival counter
};

__device__ __forceinline__ void lex_grid_barrier(unsigned int *bar, unsigned int &target)
{
    __syncthreads();
    if (gridDim.x > 1 && threadIdx.x == 0) {
        target += gridDim.x;
        __threadfence();  // this block's stores (ordered before by the CTA barrier) first
        atomicAdd(bar, 1u);
        unsigned int seen;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(bar) : "memory");
        } while (seen < target);
    }
    __syncthreads();
}

__global__ void __launch_bounds__(1024) k_gs_lex(const LexP P)
{
    const Geo &g = P.g;
    const int ni = g.ni, nj = g.nj, nk = g.nk, kh = g.kh;
    const int hmin = 3, hmax = ni + nj + nk - 6;
    const int steps = hmax - hmin + 1 + 2 * (P.iters - 1);
    const int nrj = nj - 2;
    const int gt = blockIdx.x * blockDim.x + threadIdx.x, gn = gridDim.x * blockDim.x;
    const double sixth = 1. / 6;
    unsigned int target = 0;
    for (int tau = 0; tau < steps; tau++) {
        for (int s = 0; s < P.iters; s++) {
            const int h = hmin + tau - 2 * s;
            if (h < hmin || h > hmax)
                continue;
            // planes i that hold interior points of this hyperplane
            const int i_lo = max(1, h - (nj - 2) - (nk - 2)), i_hi = min(ni - 2, h - 2);
            const int cnt = (i_hi - i_lo + 1) * nrj;
            const int c = h & 1;
            double *vc = P.v + (long long)c * g.cs;
            const double *vo = P.v + (long long)(c ^ 1) * g.cs;
            const double *dc = P.d + (long long)c * g.cs;
            for (int t = gt; t < cnt; t += gn) {
                const int i = i_lo + t / nrj, j = 1 + t % nrj, k = h - i - j;
                if (k < 1 || k > nk - 2)
                    continue;
                const long long row = ((long long)i * nj + j) * kh;
                const long long idx = row + (k >> 1);
                const double im = __ldcg(vo + idx - g.pj), ip = __ldcg(vo + idx + g.pj);
                const double jm = __ldcg(vo + idx - kh), jp = __ldcg(vo + idx + kh);
                const double km = __ldcg(vo + row + ((k - 1) >> 1));
                const double kp = __ldcg(vo + row + ((k + 1) >> 1));
                __stcg(vc + idx, gs_point(im, ip, jm, jp, km, kp, P.hSq, __ldg(dc + idx), sixth));
            }
        }
        lex_grid_barrier(P.bar, target);
    }
}

// ---------------------------------------------------------------------------------------
// the tile wavefront
// ---------------------------------------------------------------------------------------
struct LexTileP {
    Geo g;
    double *v;
    const double *d;
    double hSq;
    int nta, ntb, ntc;        // tiles along i, j, k
    int nitems;               // nta * ntb * ntc * iters work items ...
    const int4 *items;        // ... (sweep, a, b, c) sorted by a + b + c + 2 * sweep
    unsigned int *done;       // [tile]: sweeps completed (zeroed before the launch)
    unsigned int *ticket;     // next work item (zeroed before the launch)
};

__device__ __forceinline__ unsigned int lex_ld_acquire(const unsigned int *p)
{
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

__device__ __forceinline__ void lex_cp_async8(double *smem_dst, const double *gsrc)
{
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(
                     (unsigned)__cvta_generic_to_shared(smem_dst)),
                 "l"(gsrc)
                 : "memory");
}

// Threads = TI x TJ: thread (x, jl) owns the LINE of TK points along k at (i0+x, j0+jl) and
// does point kl = h - x - jl on hyperplane h of the tile, i.e. it is busy on TK consecutive
// steps out of TI+TJ+TK-2 and the lanes of a warp (consecutive jl) are active together --
// the first version mapped threads to (jl, kl) with the SHORT axis in time: every warp had a
// quarter of its lanes active on every step and a 513^3 sweep took 4.1 ms.
template <int TI, int TJ, int TK>
__global__ void __launch_bounds__(TI * TJ) k_gs_lex_tile(const LexTileP P)
{
    constexpr int PK = TK + 2;              // even: lanes (jl) step by PK - 1 (odd) words on a
    constexpr int PJ = (TJ + 2) * PK;       // hyperplane -> conflict-free
    constexpr int DK = TK + 2, DJ = TJ * DK;
    constexpr int NT = TI * TJ;
    extern __shared__ double lex_sh[];
    double *us = lex_sh;                     // (TI+2) x (TJ+2) x PK: the tile and its halo
    double *ds = lex_sh + (TI + 2) * PJ;     // TI x TJ x DK: h^2 * d
    __shared__ int s_item;
    const Geo &g = P.g;
    const int tid = threadIdx.x;
    const int jl = tid % TJ, x = tid / TJ;
    const double sixth = 1. / 6;
    for (;;) {
        if (tid == 0)
            s_item = (int)atomicAdd(P.ticket, 1u);
        __syncthreads();
        const int it = s_item;
        __syncthreads();
        if (it >= P.nitems)
            return;
        const int4 w = P.items[it];
        const int sweep = w.x, a = w.y, b = w.z, c = w.w;
        const int i0 = 1 + a * TI, j0 = 1 + b * TJ, k0 = 1 + c * TK;
        const int ti = min(TI, g.ni - 1 - i0), tj = min(TJ, g.nj - 1 - j0), tk = min(TK, g.nk - 1 - k0);
        // dependences: the three "minus" tiles at this sweep, the three "plus" tiles at the
        // previous one (and the tile itself at the previous sweep: implied by a "plus" tile
        // where there is one, not for the last tile of an axis).  The acquire + the barrier
        // order every thread's loads below after the producers' stores.
        if (tid < 7) {
            const int axis = tid >> 1, plus = tid & 1;
            const int na = a + (axis == 0 ? (plus ? 1 : -1) : 0);
            const int nb = b + (axis == 1 ? (plus ? 1 : -1) : 0);
            const int nc = c + (axis == 2 ? (plus ? 1 : -1) : 0);
            const unsigned int need = (plus || tid == 6) ? (unsigned int)sweep : (unsigned int)sweep + 1u;
            if (need && na >= 0 && na < P.nta && nb >= 0 && nb < P.ntb && nc >= 0 && nc < P.ntc) {
                const unsigned int *f = P.done + ((long long)na * P.ntb + nb) * P.ntc + nc;
                while (lex_ld_acquire(f) < need) {
                }
            }
        }
        __syncthreads();
        // the tile with its halo (every halo point is a grid point: faces are Dirichlet data)
        // and the right-hand side, as asynchronous 8-byte copies: all of a thread's ~50 loads
        // are in flight together
        {
            constexpr int EK = TK + 2, EJ = TJ + 2, TOTAL = (TI + 2) * EJ * EK;
            for (int f = tid; f < TOTAL; f += NT) {
                const int kk = f % EK, r = f / EK;
                const int jj = r % EJ, ii = r / EJ;
                const int i = i0 - 1 + ii, j = j0 - 1 + jj, k = k0 - 1 + kk;
                if (i < g.ni && j < g.nj && k < g.nk) {
                    const int col = (i + j + k) & 1;
                    lex_cp_async8(us + ii * PJ + jj * PK + kk,
                                  P.v + (long long)col * g.cs + ((long long)i * g.nj + j) * g.kh + (k >> 1));
                }
            }
            constexpr int DTOTAL = TI * TJ * TK;
            for (int f = tid; f < DTOTAL; f += NT) {
                const int kk = f % TK, r = f / TK;
                const int jj = r % TJ, ii = r / TJ;
                const int i = i0 + ii, j = j0 + jj, k = k0 + kk;
                if (ii < ti && jj < tj && kk < tk) {
                    const int col = (i + j + k) & 1;
                    lex_cp_async8(ds + ii * DJ + jj * DK + kk,
                                  P.d + (long long)col * g.cs + ((long long)i * g.nj + j) * g.kh + (k >> 1));
                }
            }
            asm volatile("cp.async.wait_all;" ::: "memory");
        }
        __syncthreads();
        const bool mine = x < ti && jl < tj;
        double *up = us + (x + 1) * PJ + (jl + 1) * PK + 1;  // kl = 0
        double *dp = ds + x * DJ + jl * DK;
        if (mine)
            for (int kl = 0; kl < tk; kl++)
                dp[kl] = __dmul_rn(P.hSq, dp[kl]);  // own line only: no barrier needed
        // hyperplanes of the tile
        const int steps = ti + tj + tk - 2;
        for (int h = 0; h < steps; h++) {
            const int kl = h - x - jl;
            if (mine && kl >= 0 && kl < tk) {
                double *q = up + kl;
                double sum = __dadd_rn(q[-PJ], q[PJ]);
                sum = __dadd_rn(sum, q[-PK]);
                sum = __dadd_rn(sum, q[PK]);
                sum = __dadd_rn(sum, q[-1]);
                sum = __dadd_rn(sum, q[1]);
                *q = __dmul_rn(sixth, __dsub_rn(sum, dp[kl]));
            }
            __syncthreads();
        }
        // back to the arrays, then "tile done"
        {
            constexpr int DTOTAL = TI * TJ * TK;
            for (int f = tid; f < DTOTAL; f += NT) {
                const int kk = f % TK, r = f / TK;
                const int jj = r % TJ, ii = r / TJ;
                if (ii < ti && jj < tj && kk < tk) {
                    const int i = i0 + ii, j = j0 + jj, k = k0 + kk;
                    const int col = (i + j + k) & 1;
                    __stcg(P.v + (long long)col * g.cs + ((long long)i * g.nj + j) * g.kh + (k >> 1),
                           us[(ii + 1) * PJ + (jj + 1) * PK + kk + 1]);
                }
            }
        }
        __syncthreads();
        if (tid == 0) {
            __threadfence();
            unsigned int *f = P.done + ((long long)a * P.ntb + b) * P.ntc + c;
            asm volatile("st.release.gpu.global.u32 [%0], %1;" ::"l"(f), "r"((unsigned int)sweep + 1u)
                         : "memory");
        }
    }
}

template <int TI, int TJ, int TK>
int launch_lex_tile(const Geo &g, double *v, const double *d, double hSq, int iters, cudaStream_t st)
{
    const int nta = (g.ni - 2 + TI - 1) / TI, ntb = (g.nj - 2 + TJ - 1) / TJ, ntc = (g.nk - 2 + TK - 1) / TK;
    const long long ntiles = (long long)nta * ntb * ntc;
    const long long nitems = ntiles * iters;
    if (nitems > (1LL << 30))
        return 1;
    // work items in the order of a + b + c + 2 * sweep (counting sort)
    const int nkeys = nta + ntb + ntc - 2 + 2 * (iters - 1);
    std::vector<long long> start(nkeys + 1, 0);
    for (int s = 0; s < iters; s++)
        for (int a = 0; a < nta; a++)
            for (int b = 0; b < ntb; b++)
                for (int c = 0; c < ntc; c++)
                    start[a + b + c + 2 * s + 1]++;
    for (int k = 0; k < nkeys; k++)
        start[k + 1] += start[k];
    std::vector<int4> items((size_t)nitems);
    for (int s = 0; s < iters; s++)
        for (int a = 0; a < nta; a++)
            for (int b = 0; b < ntb; b++)
                for (int c = 0; c < ntc; c++)
                    items[(size_t)start[a + b + c + 2 * s]++] = make_int4(s, a, b, c);
    // scratch: items, done flags, ticket (freed stream-ordered)
    char *scratch = nullptr;
    const size_t items_b = sizeof(int4) * (size_t)nitems;
    const size_t flags_b = sizeof(unsigned int) * (size_t)(ntiles + 1);
    if (cudaMallocAsync((void **)&scratch, items_b + flags_b, st) != cudaSuccess)
        return 1;
    cudaMemcpyAsync(scratch, items.data(), items_b, cudaMemcpyHostToDevice, st);
    cudaMemsetAsync(scratch + items_b, 0, flags_b, st);
    cudaStreamSynchronize(st);  // `items` is a pageable host vector about to go away
    LexTileP p{g, v, d, hSq, nta, ntb, ntc, (int)nitems, reinterpret_cast<const int4 *>(scratch),
               reinterpret_cast<unsigned int *>(scratch + items_b),
               reinterpret_cast<unsigned int *>(scratch + items_b) + ntiles};
    constexpr size_t sh = sizeof(double) * ((size_t)(TI + 2) * (TJ + 2) * (TK + 2) + (size_t)TI * TJ * (TK + 2));
    static unsigned long long attr_seen = 0;
    if (first_on_device(attr_seen))
        cudaFuncSetAttribute(k_gs_lex_tile<TI, TJ, TK>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    static int sms = 0, per_sm = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gs_lex_tile<TI, TJ, TK>, TI * TJ, sh);
        if (per_sm < 1)
            per_sm = 1;
    }
    long long blocks = (long long)sms * per_sm;
    if (blocks > nitems)
        blocks = nitems;
    k_gs_lex_tile<TI, TJ, TK><<<(unsigned)blocks, TI * TJ, sh, st>>>(p);
    const cudaError_t e = cudaGetLastError();
    ++*launch_counter();
    cudaFreeAsync(scratch, st);
    return e != cudaSuccess;
}

int &lex_mode()
{
    static int mode = getenv("MGB_GSLEX_TILE") ? atoi(getenv("MGB_GSLEX_TILE")) : 1;
    return mode;
}

}  // namespace

void gs_lex_set_mode(int mode) { lex_mode() = mode; }

// `iters` lexicographic sweeps over the interior of a whole (unpartitioned) level;
// `bar`: one unsigned int of device scratch.  Returns non-zero when the launch failed.
int launch_gs_lex(const Geo &g, double *v, const double *d, double hSq, int iters,
                  unsigned int *bar, cudaStream_t st)
{
    if (iters < 1)
        return 0;
    // tiles once a level is big enough to keep the SMs busy with them
    const int use_tile = lex_mode();
    const long long interior = (long long)(g.ni - 2) * (g.nj - 2) * (g.nk - 2);
    if (use_tile && g.ni >= 3 && g.nj >= 3 && g.nk >= 3) {
        if (use_tile == 4)
            return launch_lex_tile<8, 32, 32>(g, v, d, hSq, iters, st);
        // measured on B200 (tools/bench_gslex.py, ms per sweep at 65^3 / 129^3 / 257^3 / 513^3):
        // 16x16x32 0.21 / 0.39 / 0.78 / 2.60, 8x16x32 0.38 / 0.97 / 0.97 / 2.56, 8x8x32 0.30 /
        // 0.61 / 1.27 / 3.03, 8x32x32 0.60 / 0.49 / 1.12 / 2.79; global hyperplanes 0.69 / 1.60 /
        // 3.58 / 9.59
        if (use_tile == 5 || (use_tile == 1 && interior >= 24LL * 24 * 24))
            return launch_lex_tile<16, 16, 32>(g, v, d, hSq, iters, st);
        if (use_tile == 2)
            return launch_lex_tile<8, 16, 32>(g, v, d, hSq, iters, st);
        if (use_tile == 3)
            return launch_lex_tile<8, 8, 32>(g, v, d, hSq, iters, st);
    }
    static int sms = 0, per_sm = 0;
    if (!sms) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_gs_lex, 1024, 0);
        if (per_sm < 1)
            per_sm = 1;
    }
    // the largest hyperplane has at most (ni-2)*(nj-2) points; one block per SM keeps the
    // barrier short
    const long long most = (long long)(g.ni - 2) * (g.nj - 2);
    long long blocks = (most + 1023) / 1024;
    if (blocks > sms)
        blocks = sms;
    if (blocks < 1)
        blocks = 1;
    cudaMemsetAsync(bar, 0, sizeof(unsigned int), st);
    LexP p{g, v, d, hSq, iters, bar};
    void *args[] = {&p};
    const cudaError_t e = cudaLaunchCooperativeKernel((const void *)k_gs_lex, dim3((unsigned)blocks),
                                                      dim3(1024), args, 0, st);
    ++*launch_counter();
    return e != cudaSuccess;
}

}  // namespace mgb
