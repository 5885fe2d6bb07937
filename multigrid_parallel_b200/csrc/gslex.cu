// gslex.cu -- GaussSeidelSmoother (mg_3d.h:546-637): the LEXICOGRAPHIC Gauss-Seidel sweep
// on the GPU, bit-identical to the reference's serial triple loop.
//
// The serial sweep visits (i,j,k) in ascending order, so the update of a point reads the
// NEW values of (i-1,j,k), (i,j-1,k), (i,j,k-1) and the OLD values of the three "+1"
// neighbours.  All points of a hyperplane i+j+k = h depend only on hyperplanes h-1 (new)
// and h+1 (old) and not on each other, so sweeping the hyperplanes in ascending h with all
// points of a hyperplane in parallel reproduces the serial result exactly: each point sees
// precisely the values the serial loop would have shown it, and evaluates the same
// expression in the same order (gs_point, devmath.cuh).  Two more facts make it cheap:
//   * a hyperplane has ONE colour, (i+j+k)&1 = h&1: in the colour-split layout a step
//     writes one colour array and reads only the other one;
//   * sweep s+1 may follow sweep s two hyperplanes behind (its "+1" neighbours on h+1 were
//     finished by sweep s one step earlier, its "-1" neighbours on h-1 are its own previous
//     step), so `iters` sweeps are pipelined through ONE pass of ni+nj+nk-8 + 2(iters-1)
//     steps: further smoothing iterations are almost free.
// One persistent kernel (cooperative launch: all blocks co-resident), one grid-wide barrier
// per step (arrival counter + ld.acquire spin); values another block wrote are read with
// ld.global.cg (L2), never through L1.  The sweep is latency-bound (3N dependent steps of
// ~2-3 us), not bandwidth-bound: 513^3 costs ~1500 steps whatever the HBM rate -- still two
// orders of magnitude faster than the serial CPU sweep it replaces, which is the only
// alternative that gives the same bits.
#include <cstdio>
#include <cstdlib>

#include "devmath.cuh"
#include "kernels.h"

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

}  // namespace

// `iters` lexicographic sweeps over the interior of a whole (unpartitioned) level;
// `bar`: one unsigned int of device scratch.  Returns non-zero when the launch failed.
int launch_gs_lex(const Geo &g, double *v, const double *d, double hSq, int iters,
                  unsigned int *bar, cudaStream_t st)
{
    if (iters < 1)
        return 0;
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
