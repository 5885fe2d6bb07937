// kernels.cu -- hand-written sm_100a kernels of the multigrid V-cycle hot path.
//
// All arithmetic is IEEE fp64 in the reference's exact operation order, written
// with the explicitly rounded intrinsics (__dadd_rn/__dmul_rn/__dsub_rn) so no
// FMA contraction can happen whatever the compiler flags: the iterates are
// bit-identical to the CPU reference.
//
// Every stage is HBM-bound (0.3 flop/B); no tensor cores on purpose.
#include <cstdint>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"
#include "launch.h"
#include "devmath.cuh"

namespace mgb {

static long long g_launches = 0;
long long launches_issued() { return g_launches; }
long long *launch_counter() { return &g_launches; }
#define COUNT_LAUNCH() (++g_launches)

// ----------------------------------------------------------------------------
// device helpers
// ----------------------------------------------------------------------------
__device__ __forceinline__ double2 ld2(const double *p)
{
    return *reinterpret_cast<const double2 *>(p);
}
__device__ __forceinline__ void st2(double *p, double a, double b)
{
    *reinterpret_cast<double2 *>(p) = make_double2(a, b);
}

// mg_3d.h:89-90: x*x - 2*y*y + z*z
__device__ __forceinline__ double bc_func(double x, double y, double z)
{
    double a = __dmul_rn(x, x);
    double b = __dmul_rn(__dmul_rn(2.0, y), y);
    double c = __dmul_rn(z, z);
    return __dadd_rn(__dsub_rn(a, b), c);
}

// deterministic block sum (blockDim.x multiple of 32, <= 1024); result valid
// in thread 0
__device__ __forceinline__ double block_sum(double x)
{
    __shared__ double warp_part[32];
    for (int o = 16; o > 0; o >>= 1)
        x = __dadd_rn(x, __shfl_down_sync(0xffffffffu, x, o));
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    if (lane == 0)
        warp_part[w] = x;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double y = 0.;
    if (w == 0) {
        y = lane < nw ? warp_part[lane] : 0.;
        for (int o = 16; o > 0; o >>= 1)
            y = __dadd_rn(y, __shfl_down_sync(0xffffffffu, y, o));
    }
    return y;
}

// second stage of every reduction: one block adds the per-block partials in a
// fixed order
__global__ void __launch_bounds__(1024) k_finish_sum(const double *partials, int n,
                                                     double *out)
{
    pdl_enter();
    double acc = 0.;
    for (int t = threadIdx.x; t < n; t += blockDim.x)
        acc = __dadd_rn(acc, partials[t]);
    acc = block_sum(acc);
    if (threadIdx.x == 0)
        *out = acc;
}

// ----------------------------------------------------------------------------
// natural <-> colour-split
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(Geo g, const double *__restrict__ nat,
                                              double *__restrict__ split)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.li * g.pj;
    if (t >= total)
        return;
    const int m = (int)(t % g.kh);
    const long long row = t / g.kh;  // il*nj + j
    const int j = (int)(row % g.nj);
    const int il = (int)(row / g.nj);
    const int s = (g.i0 + il + j) & 1;
    const double *src = nat + row * g.nk;
    const int ke = 2 * m, ko = 2 * m + 1;
    const double ve = ke < g.nk ? src[ke] : 0.;
    const double vo = ko < g.nk ? src[ko] : 0.;
    split[(long long)s * g.cs + t] = ve;        // even k has colour s
    split[(long long)(s ^ 1) * g.cs + t] = vo;  // odd k the other one
}

__global__ void __launch_bounds__(256) k_unpack(Geo g, const double *__restrict__ split,
                                                double *__restrict__ nat)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.li * g.pj;
    if (t >= total)
        return;
    const int m = (int)(t % g.kh);
    const long long row = t / g.kh;
    const int j = (int)(row % g.nj);
    const int il = (int)(row / g.nj);
    const int s = (g.i0 + il + j) & 1;
    double *dst = nat + row * g.nk;
    const int ke = 2 * m, ko = 2 * m + 1;
    if (ke < g.nk)
        dst[ke] = split[(long long)s * g.cs + t];
    if (ko < g.nk)
        dst[ko] = split[(long long)(s ^ 1) * g.cs + t];
}

// a contiguous run [first, first+n) of the natural layout (local planes) <-> split
__global__ void __launch_bounds__(256) k_pack_range(Geo g, const double *__restrict__ nat,
                                                    double *__restrict__ split, long long first,
                                                    long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    const long long p = first + t;
    const int k = (int)(p % g.nk);
    const long long row = p / g.nk;
    const int j = (int)(row % g.nj), il = (int)(row / g.nj);
    const int c = (g.i0 + il + j + k) & 1;
    split[(long long)c * g.cs + row * g.kh + (k >> 1)] = nat[t];
}

__global__ void __launch_bounds__(256) k_unpack_range(Geo g, const double *__restrict__ split,
                                                      double *__restrict__ nat, long long first,
                                                      long long n)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= n)
        return;
    const long long p = first + t;
    const int k = (int)(p % g.nk);
    const long long row = p / g.nk;
    const int j = (int)(row % g.nj), il = (int)(row / g.nj);
    const int c = (g.i0 + il + j + k) & 1;
    nat[t] = split[(long long)c * g.cs + row * g.kh + (k >> 1)];
}

void launch_pack_range(const Geo &g, const double *nat, double *split, long long first,
                       long long n, cudaStream_t st)
{
    k_pack_range<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, nat, split, first, n);
    COUNT_LAUNCH();
}
void launch_unpack_range(const Geo &g, const double *split, double *nat, long long first,
                         long long n, cudaStream_t st)
{
    k_unpack_range<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(g, split, nat, first, n);
    COUNT_LAUNCH();
}

void launch_finish_sum(const double *partials, int n, double *out, cudaStream_t st)
{
    launch_k(k_finish_sum, 1, 1024, 0, st, partials, n, out);
    COUNT_LAUNCH();
}

void launch_pack(const Geo &g, const double *nat, double *split, cudaStream_t st)
{
    const long long total = (long long)g.li * g.pj;
    k_pack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, nat, split);
    COUNT_LAUNCH();
}
void launch_unpack(const Geo &g, const double *split, double *nat, cudaStream_t st)
{
    const long long total = (long long)g.li * g.pj;
    k_unpack<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, split, nat);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// Dirichlet faces
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_set_dirichlet(Geo g, double *__restrict__ a,
                                                       double h)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)g.li * g.pj;
    if (t >= total)
        return;
    const int m = (int)(t % g.kh);
    const long long row = t / g.kh;
    const int j = (int)(row % g.nj);
    const int il = (int)(row / g.nj);
    const int ig = g.i0 + il;
    const int s = (ig + j) & 1;
#pragma unroll
    for (int c = 0; c < 2; c++) {
        const int k = 2 * m + (c ^ s);
        if (k >= g.nk)
            continue;
        if (ig == 0 || ig == g.ni - 1 || j == 0 || j == g.nj - 1 || k == 0 ||
            k == g.nk - 1)
            a[(long long)c * g.cs + t] =
                bc_func(__dmul_rn((double)ig, h), __dmul_rn((double)j, h),
                        __dmul_rn((double)k, h));
    }
}

void launch_set_dirichlet(const Geo &g, double *a, double h, cudaStream_t st)
{
    const long long total = (long long)g.li * g.pj;
    k_set_dirichlet<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, a, h);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// launch geometry shared by the plane-marching kernels: x = pairs of one
// plane, y = chunks of planes; each thread marches `chunk` planes keeping the
// i-1 / i / i+1 values of its own column in registers
// ----------------------------------------------------------------------------
struct MarchCfg {
    dim3 grid, block;
    int chunk;
};

// number of plane chunks (grid.y) when one chunk is bx blocks and `resident`
// blocks fit on an SM.  Small grids: one plane per thread (all the parallelism
// there is; they are latency-bound).  Large grids: chunks of >= 8 planes (the
// i-1/i planes are re-read at every chunk start) and a block count that ends
// close to a whole number of waves over the 148 SMs.
static int pick_chunks(long long bx, int nplanes, int resident)
{
    const long long cap = 148LL * resident;
    if (bx * nplanes <= 3 * cap || nplanes < 16)
        return nplanes;
    int best = 1;
    double best_score = -1.;
    // chunks of >= 4 planes: 129^3 (127 planes) 8.3 -> 6.4 us per half-sweep against the
    // 8-plane minimum, which left the launch at 46 % of the resident slots (measured)
    const int max_by = nplanes / 4;
    for (int by = 1; by <= max_by; by++) {
        const int chunk = (nplanes + by - 1) / by;
        const int real_by = (nplanes + chunk - 1) / chunk;
        const long long total = bx * real_by;
        if (total < 2 * cap && by < max_by)
            continue;  // want at least two waves
        const long long waves = (total + cap - 1) / cap;
        double eff = (double)total / (double)(waves * cap);  // tail efficiency
        eff *= (double)chunk / (chunk + 2.0);                // chunk start-up cost
        if (eff > best_score + 1e-9) {
            best_score = eff;
            best = real_by;
        }
    }
    return best;
}

// resident blocks per SM of a kernel, from the occupancy calculator (cached)
template <typename K>
static int resident_blocks(K kernel, int threads, size_t smem)
{
    int n = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kernel, threads, smem) != cudaSuccess ||
        n < 1)
        n = 1;
    return n;
}

static MarchCfg march_cfg(const Geo &g, int nplanes, int threads, int sm_blocks)
{
    MarchCfg c;
    const long long pairs = (long long)g.pj / 2;
    const unsigned bx = (unsigned)((pairs + threads - 1) / threads);
    static const int chunk_env = getenv("MGB_MARCH_CHUNK") ? atoi(getenv("MGB_MARCH_CHUNK")) : 0;
    const int nch = chunk_env > 0 ? (nplanes + chunk_env - 1) / chunk_env
                                  : pick_chunks(bx, nplanes, sm_blocks);
    c.chunk = (nplanes + nch - 1) / nch;
    const unsigned by = (unsigned)((nplanes + c.chunk - 1) / c.chunk);
    c.grid = dim3(bx, by, 1);
    c.block = dim3(threads, 1, 1);
    return c;
}

// ----------------------------------------------------------------------------
// RB-GS half-sweep (mg_3d.h:432-443, 658-702 / 729-773)
//   reads : other colour of v (i+-1, j+-1 rows at the same m; own row at
//           m+kp-1, m+kp), this colour of d
//   writes: this colour of v              -> 12 B/DOF, 128-bit accesses
// ----------------------------------------------------------------------------
template <int COLOUR>
__global__ void __launch_bounds__(256)
k_half_sweep(Geo g, const double *__restrict__ vo, double *__restrict__ vc,
             const double *__restrict__ dc, double hSq, int il_lo, int il_hi,
             int chunk)
{
    pdl_enter();
    const int npair = g.kh >> 1;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (g.pj >> 1))
        return;
    const int j = (int)(q / npair);
    if (j < 1 || j > g.nj - 2)
        return;
    const int mp = (int)(q - (long long)j * npair);
    const int ia = il_lo + blockIdx.y * chunk;
    const int ib = min(ia + chunk, il_hi);
    if (ia >= ib)
        return;
    const long long off = 2 * q;
    const double sixth = 1. / 6;
    const int kh = g.kh;
    const int kmax = g.nk - 2;

    const double *po = vo + (long long)ia * g.pj + off;
    double2 bot = ld2(po - g.pj);
    double2 mid = ld2(po);
    for (int il = ia; il < ib; il++, po += g.pj) {
        const double2 top = ld2(po + g.pj);
        const double2 jm = ld2(po - kh);
        const double2 jp = ld2(po + kh);
        const long long idx = (long long)il * g.pj + off;
        const double2 dd = ld2(dc + idx);
        const int kp = (COLOUR ^ (g.i0 + il + j)) & 1;
        // same-row neighbours of the pair (m, m+1): other-colour entries
        // m+kp-1, m+kp, m+kp+1
        double a0, a1, a2;
        if (kp) {
            a0 = mid.x; a1 = mid.y; a2 = po[2];
        } else {
            a0 = po[-1]; a1 = mid.x; a2 = mid.y;
        }
        const double r0 = gs_point(bot.x, top.x, jm.x, jp.x, a0, a1, hSq, dd.x, sixth);
        const double r1 = gs_point(bot.y, top.y, jm.y, jp.y, a1, a2, hSq, dd.y, sixth);
        const int k0 = 4 * mp + kp, k1 = k0 + 2;
        const bool ok0 = k0 >= 1 && k0 <= kmax;
        const bool ok1 = k1 <= kmax;  // k1 >= 2 always
        if (ok0 && ok1)
            st2(vc + idx, r0, r1);
        else if (ok0)
            vc[idx] = r0;
        else if (ok1)
            vc[idx + 1] = r1;
        bot = mid;
        mid = top;
    }
}

// The same half-sweep, software-pipelined: the two DRAM streams of plane il+1
// (the thread's own pair of the other colour and of the rhs) are requested
// before plane il is computed, into a second register set (the plane loop is
// unrolled by two so that no in-flight value is ever copied).  Twice the bytes
// in flight per thread; the j+-1 rows and the k neighbour still come through
// L1/L2 at the point of use.
// HALO = false: the single-GPU instantiation carries none of the exchange code
template <int COLOUR, bool HALO>
__global__ void __launch_bounds__(256, 4)
k_half_sweep_pipe(Geo g, const double *__restrict__ vo, double *__restrict__ vc,
                  const double *__restrict__ dc, double hSq, int il_lo, int il_hi, int chunk,
                  const HaloCtl h)
{
    pdl_enter();
    const int npair = g.kh >> 1;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = il_lo + blockIdx.y * chunk;
    const int ib = min(ia + chunk, il_hi);
    if (ia >= ib)
        return;  // whole block
    // partitioned level: the first / last chunk of the slab reads a halo plane
    if (HALO)
        halo_wait_cta(h, ia == il_lo, ib == il_hi);
    const int j = (int)(q / npair);
    const bool active = q < (g.pj >> 1) && j >= 1 && j <= g.nj - 2;
    const bool pushes = HALO && h.epoch && (halo_takes_part(h.push_up, ia, ib) ||
                                            halo_takes_part(h.push_low, ia, ib));
    if (active) {
        const int mp = (int)(q - (long long)j * npair);
        const long long off = 2 * q;
        const double sixth = 1. / 6;
        const int kh = g.kh;
        const int kmax = g.nk - 2;
        const long long pj = g.pj;

        const double *po = vo + (long long)ia * pj + off;  // own pair, other colour, plane il
        const double *pd = dc + (long long)ia * pj + off;  // rhs of this colour, plane il
        double *pc = vc + (long long)ia * pj + off;
        double2 bot = ld2(po - pj);
        double2 mid = ld2(po);
        struct Pre { double2 top, dd; };
        Pre A, B;
        A.top = ld2(po + pj);
        A.dd = ld2(pd);
        B = A;
        int kp = (COLOUR ^ (g.i0 + ia + j)) & 1;
        int il = ia;
        auto step = [&](const Pre &cur, Pre &nxt) {
            if (il + 1 < ib) {
                nxt.top = ld2(po + 2 * pj);
                nxt.dd = ld2(pd + pj);
            }
            const double2 jm = ld2(po - kh);
            const double2 jp = ld2(po + kh);
            double a0, a1, a2;
            if (kp) {
                a0 = mid.x; a1 = mid.y; a2 = po[2];
            } else {
                a0 = po[-1]; a1 = mid.x; a2 = mid.y;
            }
            const double r0 = gs_point(bot.x, cur.top.x, jm.x, jp.x, a0, a1, hSq, cur.dd.x, sixth);
            const double r1 = gs_point(bot.y, cur.top.y, jm.y, jp.y, a1, a2, hSq, cur.dd.y, sixth);
            const int k0 = 4 * mp + kp, k1 = k0 + 2;
            const bool ok0 = k0 >= 1 && k0 <= kmax;
            const bool ok1 = k1 <= kmax;  // k1 >= 2 always
            if (ok0 && ok1)
                st2(pc, r0, r1);
            else if (ok0)
                pc[0] = r0;
            else if (ok1)
                pc[1] = r1;
            if (pushes)  // boundary plane of the slab: the same store, into the peer's halo
                halo_mirror_pair(h, il, off, ok0, ok1, r0, r1);
            bot = mid;
            mid = cur.top;  // already consumed above: the copy does not wait
            po += pj; pd += pj; pc += pj;
            kp ^= 1;
            il++;
        };
        while (il < ib) {
            step(A, B);
            if (il < ib)
                step(B, A);
        }
    }
    if (pushes) {
        if (halo_takes_part(h.push_up, ia, ib))
            halo_signal_cta(h, h.push_up);
        if (halo_takes_part(h.push_low, ia, ib))
            halo_signal_cta(h, h.push_low);
    }
}

// First half-sweep of a level whose guess is identically zero (mg_3d.h:1258-1259
// zeroes every coarse level before pre-smoothing): the six neighbours are 0, so
// the update needs the rhs only -- same operations as gs_point with zeros, hence
// the same bits -- and the level does not have to be zeroed first: its interior
// is overwritten by this and the next colour's sweep, its faces and pads stay 0.
template <int COLOUR>
__global__ void __launch_bounds__(256)
k_first_sweep_zero(Geo g, double *__restrict__ vc, const double *__restrict__ dc, double hSq,
                   int il_lo, int il_hi, const HaloCtl h)
{
    pdl_enter();
    const int npair = g.kh >> 1;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int il = il_lo + blockIdx.y;
    if (il >= il_hi)
        return;  // whole block
    const int j = (int)(q / npair);
    if (q < (g.pj >> 1) && j >= 1 && j <= g.nj - 2) {
        const int mp = (int)(q - (long long)j * npair);
        const long long idx = (long long)il * g.pj + 2 * q;
        const double2 dd = ld2(dc + idx);
        const double sixth = 1. / 6;
        const double r0 = gs_point(0., 0., 0., 0., 0., 0., hSq, dd.x, sixth);
        const double r1 = gs_point(0., 0., 0., 0., 0., 0., hSq, dd.y, sixth);
        const int kp = (COLOUR ^ (g.i0 + il + j)) & 1;
        const int k0 = 4 * mp + kp, k1 = k0 + 2, kmax = g.nk - 2;
        const bool ok0 = k0 >= 1 && k0 <= kmax, ok1 = k1 <= kmax;
        if (ok0 && ok1)
            st2(vc + idx, r0, r1);
        else if (ok0)
            vc[idx] = r0;
        else if (ok1)
            vc[idx + 1] = r1;
        if (h.epoch)  // reads the rhs only: nothing to wait for, but the result travels
            halo_mirror_pair(h, il, 2 * q, ok0, ok1, r0, r1);
    }
    if (h.epoch) {
        if (halo_takes_part(h.push_up, il, il + 1))
            halo_signal_cta(h, h.push_up);
        if (halo_takes_part(h.push_low, il, il + 1))
            halo_signal_cta(h, h.push_low);
    }
}

void launch_first_sweep_zero(const Geo &g, double *v, const double *d, double hSq, int colour,
                             int il_lo, int il_hi, cudaStream_t st, const HaloCtl *hp)
{
    if (il_hi <= il_lo)
        return;
    const long long pairs = (long long)g.pj / 2;
    const dim3 grid((unsigned)((pairs + 255) / 256), (unsigned)(il_hi - il_lo));
    HaloCtl h = hp ? *hp : HaloCtl{};
    h.push_up.nblocks = halo_chunks_with(h.push_up, il_lo, il_hi, 1) * grid.x;
    h.push_low.nblocks = halo_chunks_with(h.push_low, il_lo, il_hi, 1) * grid.x;
    if (colour)
        launch_k(k_first_sweep_zero<1>, grid, 256, 0, st, g, v + g.cs, d + g.cs, hSq, il_lo, il_hi, h);
    else
        launch_k(k_first_sweep_zero<0>, grid, 256, 0, st, g, v, d, hSq, il_lo, il_hi, h);
    COUNT_LAUNCH();
}

void launch_half_sweep(const Geo &g, double *v, const double *d, double hSq,
                       int colour, int il_lo, int il_hi, cudaStream_t st, const HaloCtl *hp)
{
    if (il_hi <= il_lo)
        return;
    if (launch_tile_half_sweep(g, v, d, hSq, colour, il_lo, il_hi, st, hp))
        return;
    static const int pipe = getenv("MGB_SWEEP_PIPE") ? atoi(getenv("MGB_SWEEP_PIPE")) : 1;
    if (pipe || hp) {
        static const int occp = resident_blocks(k_half_sweep_pipe<1, false>, 256, 0);
        static const int occh = resident_blocks(k_half_sweep_pipe<1, true>, 256, 0);
        const MarchCfg c = march_cfg(g, il_hi - il_lo, 256, hp ? occh : occp);
        HaloCtl h = hp ? *hp : HaloCtl{};
        h.push_up.nblocks = halo_chunks_with(h.push_up, il_lo, il_hi, c.chunk) * c.grid.x;
        h.push_low.nblocks = halo_chunks_with(h.push_low, il_lo, il_hi, c.chunk) * c.grid.x;
        if (hp) {
            if (colour)
                launch_k(k_half_sweep_pipe<1, true>, c.grid, c.block, 0, st, g, v, v + g.cs, d + g.cs, hSq,
                                                                       il_lo, il_hi, c.chunk, h);
            else
                launch_k(k_half_sweep_pipe<0, true>, c.grid, c.block, 0, st, g, v + g.cs, v, d, hSq, il_lo,
                                                                       il_hi, c.chunk, h);
        } else {
            if (colour)
                launch_k(k_half_sweep_pipe<1, false>, c.grid, c.block, 0, st, g, v, v + g.cs, d + g.cs, hSq,
                                                                        il_lo, il_hi, c.chunk, h);
            else
                launch_k(k_half_sweep_pipe<0, false>, c.grid, c.block, 0, st, g, v + g.cs, v, d, hSq, il_lo,
                                                                        il_hi, c.chunk, h);
        }
        COUNT_LAUNCH();
        return;
    }
    static const int occ = resident_blocks(k_half_sweep<1>, 256, 0);
    const MarchCfg c = march_cfg(g, il_hi - il_lo, 256, occ);
    if (colour)
        launch_k(k_half_sweep<1>, c.grid, c.block, 0, st, g, v, v + g.cs, d + g.cs, hSq,
                                                    il_lo, il_hi, c.chunk);
    else
        launch_k(k_half_sweep<0>, c.grid, c.block, 0, st, g, v + g.cs, v, d, hSq, il_lo,
                                                    il_hi, c.chunk);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// residual (mg_3d.h:794-842): both colours per thread, optional store, sum of
// squares by warp shuffles -> per-block partial -> k_finish_sum
// ----------------------------------------------------------------------------
template <bool STORE>
__global__ void __launch_bounds__(256)
k_residual(Geo g, const double *__restrict__ v, const double *__restrict__ d,
           double *__restrict__ r, double invHsq, int il_lo, int il_hi, int chunk,
           double *__restrict__ partials, const HaloCtl h)
{
    pdl_enter();
    const int npair = g.kh >> 1;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.;
    const int j = (int)(q / npair);
    const int ia = il_lo + blockIdx.y * chunk;
    const int ib = min(ia + chunk, il_hi);
    halo_wait_cta(h, ia < ib && ia == il_lo, ia < ib && ib == il_hi);
    if (q < (g.pj >> 1) && j >= 1 && j <= g.nj - 2 && ia < ib) {
        const int mp = (int)(q - (long long)j * npair);
        const long long off = 2 * q;
        const int kh = g.kh;
        const int kmax = g.nk - 2;
        const double *p0 = v + (long long)ia * g.pj + off;  // colour 0
        const double *p1 = p0 + g.cs;                        // colour 1
        double2 bot0 = ld2(p0 - g.pj), mid0 = ld2(p0);
        double2 bot1 = ld2(p1 - g.pj), mid1 = ld2(p1);
        for (int il = ia; il < ib; il++, p0 += g.pj, p1 += g.pj) {
            const double2 top0 = ld2(p0 + g.pj), top1 = ld2(p1 + g.pj);
            const double2 jm0 = ld2(p0 - kh), jp0 = ld2(p0 + kh);
            const double2 jm1 = ld2(p1 - kh), jp1 = ld2(p1 + kh);
            const long long idx = (long long)il * g.pj + off;
            const double2 d0 = ld2(d + idx), d1 = ld2(d + g.cs + idx);
            const int s = (g.i0 + il + j) & 1;
            // colour 0 points: kp = s, neighbours in colour 1
            // colour 1 points: kp = s^1, neighbours in colour 0
            double b0, b1, b2;  // colour-1 entries m+s-1, m+s, m+s+1
            double c0, c1, c2;  // colour-0 entries m+(s^1)-1, ...
            if (s) {
                b0 = mid1.x; b1 = mid1.y; b2 = p1[2];
                c0 = p0[-1]; c1 = mid0.x; c2 = mid0.y;
            } else {
                b0 = p1[-1]; b1 = mid1.x; b2 = mid1.y;
                c0 = mid0.x; c1 = mid0.y; c2 = p0[2];
            }
            const double rb0 = res_point(bot1.x, top1.x, jm1.x, jp1.x, b0, b1, mid0.x, d0.x, invHsq);
            const double rb1 = res_point(bot1.y, top1.y, jm1.y, jp1.y, b1, b2, mid0.y, d0.y, invHsq);
            const double rr0 = res_point(bot0.x, top0.x, jm0.x, jp0.x, c0, c1, mid1.x, d1.x, invHsq);
            const double rr1 = res_point(bot0.y, top0.y, jm0.y, jp0.y, c1, c2, mid1.y, d1.y, invHsq);
            // colour 0: k = 4mp + s (+2); colour 1: k = 4mp + (s^1) (+2)
            const int kb0 = 4 * mp + s, kb1 = kb0 + 2;
            const int kr0 = 4 * mp + (s ^ 1), kr1 = kr0 + 2;
            const bool okb0 = kb0 >= 1 && kb0 <= kmax, okb1 = kb1 <= kmax;
            const bool okr0 = kr0 >= 1 && kr0 <= kmax, okr1 = kr1 <= kmax;
            if (okb0) acc = __dadd_rn(acc, __dmul_rn(rb0, rb0));
            if (okr0) acc = __dadd_rn(acc, __dmul_rn(rr0, rr0));
            if (okb1) acc = __dadd_rn(acc, __dmul_rn(rb1, rb1));
            if (okr1) acc = __dadd_rn(acc, __dmul_rn(rr1, rr1));
            if (STORE) {
                if (okb0 && okb1) st2(r + idx, rb0, rb1);
                else if (okb0) r[idx] = rb0;
                else if (okb1) r[idx + 1] = rb1;
                if (okr0 && okr1) st2(r + g.cs + idx, rr0, rr1);
                else if (okr0) r[g.cs + idx] = rr0;
                else if (okr1) r[g.cs + idx + 1] = rr1;
            }
            bot0 = mid0; mid0 = top0;
            bot1 = mid1; mid1 = top1;
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0)
        partials[blockIdx.y * gridDim.x + blockIdx.x] = acc;
}

void launch_residual(const Geo &g, const double *v, const double *d, double *r,
                     double invHsq, int il_lo, int il_hi, double *partials,
                     double *out_sumsq, cudaStream_t st, const HaloCtl *hp)
{
    const HaloCtl h = hp ? *hp : HaloCtl{};
    if (il_hi <= il_lo) {
        cudaMemsetAsync(out_sumsq, 0, sizeof(double), st);
        return;
    }
    static const int occ = resident_blocks(k_residual<true>, 256, 0);
    MarchCfg c = march_cfg(g, il_hi - il_lo, 256, occ);
    while ((long long)c.grid.x * c.grid.y > kMaxPartials) {  // coarser chunks
        c.chunk *= 2;
        c.grid.y = (il_hi - il_lo + c.chunk - 1) / c.chunk;
    }
    if (r)
        launch_k(k_residual<true>, c.grid, c.block, 0, st, g, v, d, r, invHsq, il_lo, il_hi,
                                                     c.chunk, partials, h);
    else
        launch_k(k_residual<false>, c.grid, c.block, 0, st, g, v, d, nullptr, invHsq, il_lo,
                                                      il_hi, c.chunk, partials, h);
    COUNT_LAUNCH();
    launch_k(k_finish_sum, 1, 1024, 0, st, partials, (int)(c.grid.x * c.grid.y), out_sumsq);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// full-weighting restriction (mg_3d.h:844-998): one thread per coarse entry
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_restrict(Geo gf, const double *__restrict__ rf, Geo gc, double *__restrict__ dc,
           int Il_lo, int Il_hi)
{
    pdl_enter();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_colour = (long long)(Il_hi - Il_lo) * gc.pj;
    if (t >= 2 * per_colour)
        return;
    const int cc = t >= per_colour;
    const long long e = t - cc * per_colour;
    const int M = (int)(e % gc.kh);
    const long long row = e / gc.kh;
    const int J = (int)(row % gc.nj);
    const int Il = Il_lo + (int)(row / gc.nj);
    const int Ig = gc.i0 + Il;
    const int K = 2 * M + ((cc ^ (Ig + J)) & 1);
    if (K >= gc.nk)
        return;
    const int fil = 2 * Ig - gf.i0, fj = 2 * J, fk = 2 * K;
    double val;
    if (Ig == 0 || Ig == gc.ni - 1 || J == 0 || J == gc.nj - 1 || K == 0 ||
        K == gc.nk - 1) {
        val = rd_split(gf, rf, fil, fj, fk);  // injection (881-957)
    } else {
        val = 0.;
#pragma unroll
        for (int a = 0; a < 3; a++)
#pragma unroll
            for (int b = 0; b < 3; b++)
#pragma unroll
                for (int c = 0; c < 3; c++) {
                    const int e3 = (a != 1) + (b != 1) + (c != 1);
                    const double w = 1.0 / (double)(8 << e3);
                    val = __dadd_rn(val, __dmul_rn(rd_split(gf, rf, fil - 1 + a,
                                                           fj - 1 + b, fk - 1 + c), w));
                }
    }
    dc[(long long)cc * gc.cs + ((long long)Il * gc.nj + J) * gc.kh + M] = val;
}

void launch_restrict(const Geo &gf, const double *rf, const Geo &gc, double *dc,
                     int Il_lo, int Il_hi, cudaStream_t st)
{
    if (Il_hi <= Il_lo)
        return;
    const long long total = 2LL * (Il_hi - Il_lo) * gc.pj;
    launch_k(k_restrict, (unsigned)((total + 255) / 256), 256, 0, st, gf, rf, gc, dc, Il_lo,
                                                               Il_hi);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// fused residual + full-weighting restriction (mg_3d.h:794-842 + 844-998)
//
// The fine residual is never written to HBM (17 B/DOF instead of 24 + 9).
// A block owns TY coarse rows (all K) and marches over fine planes.  Per fine
// plane f every thread computes the residual of its four fine points
// (row j, k = 4mp..4mp+3, both colours; i-neighbours carried in registers like
// k_residual) and drops them into a shared plane buffer in natural k order;
// after one barrier the threads sitting on even fine rows (j = 2J) add the
// 3x3 neighbourhood of that plane into the accumulators of their two coarse
// points (J, 2mp) and (J, 2mp+1).  A coarse value is the sum over the planes
// 2I-1, 2I, 2I+1 in that order, nine terms each in (tj,tk) order, starting from
// 0. -- exactly the (ti,tj,tk) loop nest of mg_3d.h:980-989 -- so an odd plane
// closes coarse plane I and opens I+1 with the same nine products.
// The block covers fine rows 2Ja-1 .. 2(Ja+TY)-1: one odd row is recomputed by
// the neighbouring block (1/(2TY+1) redundancy, L2 hits).
// Boundary coarse entries are the injected boundary residual, which is 0
// (mg_3d.h:881-957 with r's faces never written).
// ----------------------------------------------------------------------------
template <int MAXT>
__global__ void __launch_bounds__(MAXT)
k_residual_restrict(Geo gf, const double *__restrict__ v, const double *__restrict__ d,
                    double invHsq, Geo gc, double *__restrict__ dc, int Il_lo, int Il_hi,
                    int chunk, int TY, const HaloCtl h)
{
    pdl_enter();
    extern __shared__ double rs_all[];
    const int npair = gf.kh >> 1;
    const int R = 2 * TY + 1;
    const int SP = 4 * npair + 2;  // row pitch; data starts at +2 (k = -1 lands on +1)
    const int t = threadIdx.x;
    const int jl = t / npair;
    const int mp = t - jl * npair;
    const bool active = jl < R;
    const int Ja = blockIdx.x * TY;
    const int j = 2 * Ja - 1 + jl;
    const bool row_in = active && j >= 1 && j <= gf.nj - 2;  // interior fine row
    // coarse role: even fine rows j = 2J
    const int J = Ja + ((jl - 1) >> 1);
    const bool coarse_thr = active && (jl & 1) && jl <= R - 2 && J < gc.nj;
    const int K0 = 2 * mp, K1 = 2 * mp + 1;
    const bool K0in = K0 < gc.nk, K1in = K1 < gc.nk;
    const bool Jint = J >= 1 && J <= gc.nj - 2;
    const bool K0int = Jint && K0 >= 1 && K0 <= gc.nk - 2;
    const bool K1int = Jint && K1 >= 1 && K1 <= gc.nk - 2;

    const int Ia = Il_lo + blockIdx.y * chunk;
    const int Ib = min(Ia + chunk, Il_hi);
    // interior coarse planes of this chunk (global index 1 .. ni-2)
    const int Im0 = max(Ia, 1 - gc.i0);
    const int Im1 = min(Ib, gc.ni - 1 - gc.i0);
    // partitioned level: the chunks at the ends of the slab read fine halo planes
    halo_wait_cta(h, Ia < Ib && Ia == Il_lo, Ia < Ib && Ib == Il_hi);

    // coarse boundary planes inside the chunk: zeros
    if (coarse_thr) {
        for (int Il = Ia; Il < Ib; Il++) {
            const int Ig = gc.i0 + Il;
            if (Ig != 0 && Ig != gc.ni - 1)
                continue;
            const int S = (Ig + J) & 1;
            const long long row = ((long long)Il * gc.nj + J) * gc.kh;
            if (K0in) dc[(long long)S * gc.cs + row + mp] = 0.;
            if (K1in) dc[(long long)(S ^ 1) * gc.cs + row + mp] = 0.;
        }
    }
    if (Im0 >= Im1)
        return;

    const int f0 = 2 * (gc.i0 + Im0) - 1 - gf.i0;      // first fine local plane
    const int f1 = 2 * (gc.i0 + Im1 - 1) + 1 - gf.i0;  // last fine local plane
    const long long off = ((long long)j * npair + mp) * 2;
    const int kh = gf.kh;
    const int kmax = gf.nk - 2;
    const double *p0 = v + (long long)f0 * gf.pj + off;
    const double *p1 = p0 + gf.cs;
    double2 bot0, mid0, bot1, mid1;
    bot0 = mid0 = bot1 = mid1 = make_double2(0., 0.);
    if (row_in) {
        bot0 = ld2(p0 - gf.pj); mid0 = ld2(p0);
        bot1 = ld2(p1 - gf.pj); mid1 = ld2(p1);
    }
    double acc0 = 0., acc1 = 0.;
    bool have_cur = false;
    int buf = 0;
    for (int il = f0; il <= f1; il++, p0 += gf.pj, p1 += gf.pj) {
        double *rs = rs_all + (size_t)buf * R * SP;
        const int ig = gf.i0 + il;
        if (active) {
            double n0 = 0., n1 = 0., n2 = 0., n3 = 0.;  // residuals at k = 4mp..4mp+3
            if (row_in) {
                const double2 top0 = ld2(p0 + gf.pj), top1 = ld2(p1 + gf.pj);
                const double2 jm0 = ld2(p0 - kh), jp0 = ld2(p0 + kh);
                const double2 jm1 = ld2(p1 - kh), jp1 = ld2(p1 + kh);
                const long long idx = (long long)il * gf.pj + off;
                const double2 d0 = ld2(d + idx), d1 = ld2(d + gf.cs + idx);
                const int s = (ig + j) & 1;
                double b0, b1, b2, c0, c1, c2;
                if (s) {
                    b0 = mid1.x; b1 = mid1.y; b2 = p1[2];
                    c0 = p0[-1]; c1 = mid0.x; c2 = mid0.y;
                } else {
                    b0 = p1[-1]; b1 = mid1.x; b2 = mid1.y;
                    c0 = mid0.x; c1 = mid0.y; c2 = p0[2];
                }
                double rb0 = res_point(bot1.x, top1.x, jm1.x, jp1.x, b0, b1, mid0.x, d0.x, invHsq);
                double rb1 = res_point(bot1.y, top1.y, jm1.y, jp1.y, b1, b2, mid0.y, d0.y, invHsq);
                double rr0 = res_point(bot0.x, top0.x, jm0.x, jp0.x, c0, c1, mid1.x, d1.x, invHsq);
                double rr1 = res_point(bot0.y, top0.y, jm0.y, jp0.y, c1, c2, mid1.y, d1.y, invHsq);
                const int kb0 = 4 * mp + s, kr0 = 4 * mp + (s ^ 1);
                if (!(kb0 >= 1 && kb0 <= kmax)) rb0 = 0.;
                if (!(kb0 + 2 <= kmax)) rb1 = 0.;
                if (!(kr0 >= 1 && kr0 <= kmax)) rr0 = 0.;
                if (!(kr0 + 2 <= kmax)) rr1 = 0.;
                if (s) { n0 = rr0; n1 = rb0; n2 = rr1; n3 = rb1; }
                else   { n0 = rb0; n1 = rr0; n2 = rb1; n3 = rr1; }
                bot0 = mid0; mid0 = top0;
                bot1 = mid1; mid1 = top1;
            }
            double *w = rs + (size_t)jl * SP + 2 + 4 * mp;
            st2(w, n0, n1);
            st2(w + 2, n2, n3);
        }
        __syncthreads();
        if (coarse_thr) {
            const bool odd = ig & 1;
            const double wf = odd ? 0.5 : 1.0;  // ti = 0/2 vs ti = 1
            const double *r0 = rs + (size_t)(jl - 1) * SP + 2 + 4 * mp - 1;
            double s0 = 0., s1 = 0.;  // fresh sums (for the plane that opens I+1)
            double a0 = acc0, a1 = acc1;
#pragma unroll
            for (int tj = 0; tj < 3; tj++) {
                const double *rr = r0 + (size_t)tj * SP;
                const double x0 = rr[0], x1 = rr[1], x2 = rr[2], x3 = rr[3], x4 = rr[4];
                const double wc = (tj == 1 ? 0.125 : 0.0625) * wf;  // tk = 1
                const double we = 0.5 * wc;                          // tk = 0, 2
                const double p00 = __dmul_rn(x0, we), p01 = __dmul_rn(x1, wc),
                             p02 = __dmul_rn(x2, we);
                const double p10 = p02, p11 = __dmul_rn(x3, wc), p12 = __dmul_rn(x4, we);
                a0 = __dadd_rn(__dadd_rn(__dadd_rn(a0, p00), p01), p02);
                a1 = __dadd_rn(__dadd_rn(__dadd_rn(a1, p10), p11), p12);
                s0 = __dadd_rn(__dadd_rn(__dadd_rn(s0, p00), p01), p02);
                s1 = __dadd_rn(__dadd_rn(__dadd_rn(s1, p10), p11), p12);
            }
            if (odd) {
                if (have_cur) {  // plane 2I+1 closes coarse plane I
                    const int Il = ((ig - 1) >> 1) - gc.i0;
                    const int S = (gc.i0 + Il + J) & 1;
                    const long long row = ((long long)Il * gc.nj + J) * gc.kh;
                    if (K0in) dc[(long long)S * gc.cs + row + mp] = K0int ? a0 : 0.;
                    if (K1in) dc[(long long)(S ^ 1) * gc.cs + row + mp] = K1int ? a1 : 0.;
                    // my last coarse plane is the upper neighbour's coarse-rhs halo
                    if (h.push_up.peer_flag && Il == h.push_up.plane[0]) {
                        const long long prow = (long long)J * gc.kh + mp;
                        if (K0in) h.push_up.dst[S][prow] = K0int ? a0 : 0.;
                        if (K1in) h.push_up.dst[S ^ 1][prow] = K1int ? a1 : 0.;
                    }
                }
                acc0 = s0;  // ... and opens I+1
                acc1 = s1;
                have_cur = true;
            } else {
                acc0 = a0;
                acc1 = a1;
            }
        }
        buf ^= 1;
    }
    if (halo_takes_part(h.push_up, Ia, Ib))
        halo_signal_cta(h, h.push_up);
}

void launch_residual_restrict(const Geo &gf, const double *vf, const double *df,
                              double invHsq, const Geo &gc, double *dc, int Il_lo,
                              int Il_hi, cudaStream_t st, const HaloCtl *hp)
{
    if (Il_hi <= Il_lo)
        return;
    const int npair = gf.kh >> 1;
    static const int maxt_env = getenv("MGB_RR_MAXT") ? atoi(getenv("MGB_RR_MAXT")) : 0;
    const int maxt = maxt_env >= 96 && maxt_env <= 1024 ? maxt_env : 1024;
    int R = maxt / npair;
    if (R < 3) {
        fprintf(stderr, "mgb: k_residual_restrict: rows of %d points are too long\n", gf.nk);
        return;
    }
    if (!(R & 1))
        R--;
    int TY = (R - 1) / 2;
    if (TY > gc.nj)
        TY = gc.nj;
    R = 2 * TY + 1;
    const int threads = ((R * npair + 31) / 32) * 32;
    const unsigned bx = (gc.nj + TY - 1) / TY;
    const int nplanes = Il_hi - Il_lo;
    const size_t sh = sizeof(double) * 2 * R * (4 * npair + 2);
    static unsigned long long attr_seen = 0;
    if (first_on_device(attr_seen))
        cudaFuncSetAttribute(k_residual_restrict<1024>,
                             cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024);
    const int occ = resident_blocks(k_residual_restrict<1024>, threads, sh);
    const int nch = pick_chunks(bx, nplanes, occ);
    const int chunk = (nplanes + nch - 1) / nch;
    const unsigned by = (nplanes + chunk - 1) / chunk;
    HaloCtl h = hp ? *hp : HaloCtl{};
    h.push_up.nblocks = halo_chunks_with(h.push_up, Il_lo, Il_hi, chunk) * bx;
    launch_k(k_residual_restrict<1024>, dim3(bx, by), threads, sh, st, gf, vf, df, invHsq, gc, dc,
                                                               Il_lo, Il_hi, chunk, TY, h);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// prolongation + correction (mg_3d.h:1000-1145)
//
// A thread owns the four fine points k = 4mp..4mp+3 of fine row (i,j) -- the
// pair (m, m+1) = (2mp, 2mp+1) of BOTH colours -- and marches over fine planes.
// Everything it needs from the coarse grid is three consecutive entries
// K = 2mp, 2mp+1, 2mp+2 of at most four coarse rows (I or I+1, J or J+1); the
// two coarse planes are carried in registers, so each coarse row is fetched
// once per two fine planes.  Fine traffic: one 128-bit load + store per colour.
// The sums keep the reference's corner order and start from 0. (1018).
// ----------------------------------------------------------------------------
struct C3 {
    double x, y, z;  // coarse entries K = 2mp, 2mp+1, 2mp+2
};

__device__ __forceinline__ C3 ld_c3(const Geo &gc, const double *__restrict__ ec, int Il,
                                    int J, int mp)
{
    const int S = (gc.i0 + Il + J) & 1;  // colour of the even K in this coarse row
    const long long row = ((long long)Il * gc.nj + J) * gc.kh;
    const double *e0 = ec + (long long)S * gc.cs + row;
    const double *e1 = ec + (long long)(S ^ 1) * gc.cs + row;
    C3 r;
    r.x = e0[mp];
    r.y = e1[mp];
    r.z = e0[mp + 1];
    return r;
}

__global__ void __launch_bounds__(256)
k_prolong_correct(Geo gc, const double *__restrict__ ec, Geo gf,
                  double *__restrict__ ef, int il_lo, int il_hi, int chunk)
{
    pdl_enter();
    const int npair = gf.kh >> 1;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= (gf.pj >> 1))
        return;
    const int j = (int)(q / npair);
    const int mp = (int)(q - (long long)j * npair);
    const int ia = il_lo + blockIdx.y * chunk;
    const int ib = min(ia + chunk, il_hi);
    if (ia >= ib)
        return;
    const int oj = j & 1, J0 = j >> 1;
    const int k0 = 4 * mp;
    if (k0 >= gf.nk)
        return;  // pad columns only
    const bool v1 = k0 + 1 < gf.nk, v2 = k0 + 2 < gf.nk, v3 = k0 + 3 < gf.nk;
    const long long off = 2 * q;

    C3 A0, A1, B0, B1;
    {
        const int I = ((gf.i0 + ia) >> 1) - gc.i0;
        A0 = ld_c3(gc, ec, I, J0, mp);
        A1 = oj ? ld_c3(gc, ec, I, J0 + 1, mp) : A0;
    }
    B0 = A0;
    B1 = A1;
    for (int il = ia; il < ib; il++) {
        const int ig = gf.i0 + il;
        const int oi = ig & 1;
        if (oi) {
            const int I1 = (ig >> 1) + 1 - gc.i0;
            B0 = ld_c3(gc, ec, I1, J0, mp);
            B1 = oj ? ld_c3(gc, ec, I1, J0 + 1, mp) : B0;
        }
        const int s = (ig + j) & 1;  // colour holding the even k of this row
        double *pe = ef + (long long)s * gf.cs + (long long)il * gf.pj + off;
        double *po = ef + (long long)(s ^ 1) * gf.cs + (long long)il * gf.pj + off;
        // even k: 4mp (coarse column x), 4mp+2 (column y)
        const double e0 = pc_even(oi, oj, A0.x, A1.x, B0.x, B1.x);
        const double e1 = pc_even(oi, oj, A0.y, A1.y, B0.y, B1.y);
        // odd k: 4mp+1 (between x and y), 4mp+3 (between y and z)
        const double o0 = pc_odd(oi, oj, A0.x, A0.y, A1.x, A1.y, B0.x, B0.y, B1.x, B1.y);
        const double o1 = pc_odd(oi, oj, A0.y, A0.z, A1.y, A1.z, B0.y, B0.z, B1.y, B1.z);
        if (v2) {
            const double2 f = ld2(pe);
            st2(pe, __dadd_rn(f.x, e0), __dadd_rn(f.y, e1));
        } else {
            pe[0] = __dadd_rn(pe[0], e0);
        }
        if (v3) {
            const double2 f = ld2(po);
            st2(po, __dadd_rn(f.x, o0), __dadd_rn(f.y, o1));
        } else if (v1) {
            po[0] = __dadd_rn(po[0], o0);
        }
        if (oi) {
            A0 = B0;
            A1 = B1;
        }
    }
}

// ----------------------------------------------------------------------------
// prolongation + correction, wide form: a thread owns EIGHT consecutive fine
// points k = 8q..8q+7 of one fine row (two 16-byte pairs of each colour) and
// marches over fine planes in (even, odd) pairs.  The five coarse entries
// K = 4q..4q+4 it interpolates from come as two 128-bit loads + one scalar per
// coarse row; coarse plane I+1 is fetched at the top of a pair and becomes
// plane I of the next one.  The fine values of the NEXT pair are loaded before
// the current pair is stored, into a second register set (the pair loop is
// unrolled by two, so no in-flight value is ever copied): every thread keeps
// 256 B of reads in flight, which is what the read-modify-write stream needs.
// ----------------------------------------------------------------------------
struct C5 {
    double v[5];  // coarse entries K = 4q .. 4q+4
};

__device__ __forceinline__ C5 ld_c5(const Geo &gc, const double *__restrict__ ec, int Il, int J,
                                    int q)
{
    const int S = (gc.i0 + Il + J) & 1;  // colour of the even K in this coarse row
    const long long row = ((long long)Il * gc.nj + J) * gc.kh + 2 * q;
    const double *e0 = ec + (long long)S * gc.cs + row;
    const double *e1 = ec + (long long)(S ^ 1) * gc.cs + row;
    const double2 a = ld2(e0), b = ld2(e1);
    C5 r;
    r.v[0] = a.x; r.v[1] = b.x; r.v[2] = a.y; r.v[3] = b.y;
    r.v[4] = e0[2];
    return r;
}

struct F8 {
    double2 e0, e1, o0, o1;  // even-k colour entries 4q..4q+3, odd-k colour entries
};

__global__ void __launch_bounds__(256, 2)
k_prolong_correct8(Geo gc, const double *__restrict__ ec, Geo gf, double *ef, int il_lo,
                   int il_hi, int chunk, int cmask, const HaloCtl h)
{
    pdl_enter();
    const int noct = gf.kh >> 2;
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const int ia = il_lo + blockIdx.y * chunk;
    const int ib = min(ia + chunk, il_hi);
    // partitioned level: the coarse rows the ends of the slab read are halo planes
    halo_wait_cta(h, ia < ib && ia == il_lo, ia < ib && ib == il_hi);
    if (t >= (long long)gf.nj * noct)
        return;
    const int j = (int)(t / noct);
    const int q = (int)(t - (long long)j * noct);
    const int k0 = 8 * q;
    if (ia >= ib || k0 >= gf.nk)
        return;  // pad columns only
    const int oj = j & 1, J0 = j >> 1;
    const int nvalid = min(8, gf.nk - k0);  // fine points k0 .. k0+nvalid-1 exist
    const long long off = (long long)j * gf.kh + 4 * q;

    // cmask bit c: colour c is corrected (the other one is left alone: inside the
    // V-cycle the post-smoother's first half-sweep overwrites it without reading it)
    auto load_f = [&](int il, F8 &f) {
        const int s = (gf.i0 + il + j) & 1;  // colour holding the even k of this row
        const double *pe = ef + (long long)s * gf.cs + (long long)il * gf.pj + off;
        const double *po = ef + (long long)(s ^ 1) * gf.cs + (long long)il * gf.pj + off;
        if ((cmask >> s) & 1) {
            f.e0 = ld2(pe); f.e1 = ld2(pe + 2);
        }
        if ((cmask >> (s ^ 1)) & 1) {
            f.o0 = ld2(po); f.o1 = ld2(po + 2);
        }
    };
    // add the interpolated correction of one fine plane and store it
    auto plane = [&](int il, int oi, const C5 &A0, const C5 &A1, const C5 &B0, const C5 &B1,
                     const F8 &f) {
        double ev[4], od[4];
#pragma unroll
        for (int e = 0; e < 4; e++) {
            ev[e] = pc_even(oi, oj, A0.v[e], A1.v[e], B0.v[e], B1.v[e]);
            od[e] = pc_odd(oi, oj, A0.v[e], A0.v[e + 1], A1.v[e], A1.v[e + 1], B0.v[e],
                           B0.v[e + 1], B1.v[e], B1.v[e + 1]);
        }
        const int s = (gf.i0 + il + j) & 1;
        const bool do_e = (cmask >> s) & 1, do_o = (cmask >> (s ^ 1)) & 1;
        double *pe = ef + (long long)s * gf.cs + (long long)il * gf.pj + off;
        double *po = ef + (long long)(s ^ 1) * gf.cs + (long long)il * gf.pj + off;
        const double r0 = __dadd_rn(f.e0.x, ev[0]), r1 = __dadd_rn(f.e0.y, ev[1]);
        const double r2 = __dadd_rn(f.e1.x, ev[2]), r3 = __dadd_rn(f.e1.y, ev[3]);
        const double w0 = __dadd_rn(f.o0.x, od[0]), w1 = __dadd_rn(f.o0.y, od[1]);
        const double w2 = __dadd_rn(f.o1.x, od[2]), w3 = __dadd_rn(f.o1.y, od[3]);
        if (nvalid == 8) {
            if (do_e) { st2(pe, r0, r1); st2(pe + 2, r2, r3); }
            if (do_o) { st2(po, w0, w1); st2(po + 2, w2, w3); }
        } else {  // last octet of the row: even k = k0+2e, odd k = k0+2e+1
            if (do_e) {
                if (nvalid > 0) pe[0] = r0;
                if (nvalid > 2) pe[1] = r1;
                if (nvalid > 4) pe[2] = r2;
                if (nvalid > 6) pe[3] = r3;
            }
            if (do_o) {
                if (nvalid > 1) po[0] = w0;
                if (nvalid > 3) po[1] = w1;
                if (nvalid > 5) po[2] = w2;
            }
        }
    };

    int il = ia;
    int I = ((gf.i0 + il) >> 1) - gc.i0;  // coarse plane at or below fine plane il
    C5 A0 = ld_c5(gc, ec, I, J0, q);
    C5 A1 = oj ? ld_c5(gc, ec, I, J0 + 1, q) : A0;
    if ((gf.i0 + il) & 1) {  // chunk starts on an odd plane: do it on its own
        const C5 B0 = ld_c5(gc, ec, I + 1, J0, q);
        const C5 B1 = oj ? ld_c5(gc, ec, I + 1, J0 + 1, q) : B0;
        F8 f;
        load_f(il, f);
        plane(il, 1, A0, A1, B0, B1, f);
        A0 = B0; A1 = B1;
        I++; il++;
    }
    if (il >= ib)
        return;
    // from here il is even: pairs (il, il+1)
    F8 Pe = F8{}, Po = F8{}, Qe = F8{}, Qo = F8{};
    load_f(il, Pe);
    if (il + 1 < ib)
        load_f(il + 1, Po);
    auto pair = [&](F8 &ce, F8 &co, F8 &ne, F8 &no) {
        const bool has_odd = il + 1 < ib;
        C5 B0 = A0, B1 = A1;
        if (has_odd) {  // coarse plane I+1: used by the odd plane, then becomes A
            B0 = ld_c5(gc, ec, I + 1, J0, q);
            B1 = oj ? ld_c5(gc, ec, I + 1, J0 + 1, q) : B0;
        }
        if (il + 2 < ib)
            load_f(il + 2, ne);
        if (il + 3 < ib)
            load_f(il + 3, no);
        plane(il, 0, A0, A1, A0, A1, ce);
        if (has_odd)
            plane(il + 1, 1, A0, A1, B0, B1, co);
        A0 = B0; A1 = B1;
        I++; il += 2;
    };
    while (il < ib) {
        pair(Pe, Po, Qe, Qo);
        if (il < ib)
            pair(Qe, Qo, Pe, Po);
    }
}

// ef[p] = ef[p] + 0. on the face points of one colour (see kernels.h).  Three
// sections (blockIdx.y): 0 = the two i-faces, whole colour-planes, coalesced;
// 1 = the rows j = 0 and j = nj-1 of every other plane; 2 = the k = 0 / k = nk-1
// entries of every remaining row.  Every face point is touched exactly once.
__global__ void __launch_bounds__(256)
k_add_zero_faces(Geo g, double *__restrict__ a, int colour, int il_lo, int il_hi)
{
    pdl_enter();
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double *base = a + (long long)colour * g.cs;
    const int lo_face = -g.i0, hi_face = g.ni - 1 - g.i0;  // local indices of the i-faces
    if (blockIdx.y == 0) {
        // i-faces inside [il_lo, il_hi): entry t of plane 0 / plane 1
        const int which = (int)(t / g.pj);
        if (which > 1)
            return;
        const int il = which == 0 ? lo_face : hi_face;
        if (il < il_lo || il >= il_hi)
            return;
        const long long e = t - (long long)which * g.pj;
        const int j = (int)(e / g.kh), m = (int)(e - (long long)j * g.kh);
        const int kp = (colour ^ (g.i0 + il + j)) & 1;
        if (2 * m + kp < g.nk) {
            double *p = base + (long long)il * g.pj + e;
            *p = __dadd_rn(*p, 0.);
        }
    } else if (blockIdx.y == 1) {
        // rows j = 0 and j = nj-1 of the planes that are not i-faces
        const long long per_plane = 2LL * g.kh;
        const int il = il_lo + (int)(t / per_plane);
        if (il >= il_hi || il == lo_face || il == hi_face)
            return;
        const long long e = t % per_plane;
        const int j = e < g.kh ? 0 : g.nj - 1;
        const int m = (int)(e < g.kh ? e : e - g.kh);
        const int kp = (colour ^ (g.i0 + il + j)) & 1;
        if (2 * m + kp < g.nk) {
            double *p = base + (long long)il * g.pj + (long long)j * g.kh + m;
            *p = __dadd_rn(*p, 0.);
        }
    } else {
        // k = 0 and k = nk-1 of the rows that are on no other face
        const int il = il_lo + (int)(t / g.nj);
        const int j = (int)(t % g.nj);
        if (il >= il_hi || il == lo_face || il == hi_face || j == 0 || j == g.nj - 1)
            return;
        const int kp = (colour ^ (g.i0 + il + j)) & 1;
        double *row = base + (long long)il * g.pj + (long long)j * g.kh;
        if (kp == 0)
            row[0] = __dadd_rn(row[0], 0.);  // k = 0
        if (((g.nk - 1 - kp) & 1) == 0) {    // k = nk-1 has this colour
            const int m = (g.nk - 1 - kp) >> 1;
            row[m] = __dadd_rn(row[m], 0.);
        }
    }
}

void launch_add_zero_faces(const Geo &g, double *a, int colour, int il_lo, int il_hi,
                           cudaStream_t st)
{
    if (il_hi <= il_lo)
        return;
    const long long n0 = 2 * g.pj, n1 = (long long)(il_hi - il_lo) * 2 * g.kh,
                    n2 = (long long)(il_hi - il_lo) * g.nj;
    const long long most = n0 > n1 ? (n0 > n2 ? n0 : n2) : (n1 > n2 ? n1 : n2);
    launch_k(k_add_zero_faces, dim3((unsigned)((most + 255) / 256), 3), 256, 0, st, g, a, colour, il_lo,
                                                                            il_hi);
    COUNT_LAUNCH();
}

void launch_prolong_correct(const Geo &gc, const double *ec, const Geo &gf,
                            double *ef, int il_lo, int il_hi, cudaStream_t st, int cmask,
                            const HaloCtl *hp)
{
    if (il_hi <= il_lo)
        return;
    // the TMA ring kernel marches (even, odd) plane pairs: a range that starts on an odd
    // global plane (a slab's lower halo plane) gets that one plane from the marching kernel
    if (((gf.i0 + il_lo) & 1) && il_hi - il_lo > 8) {
        launch_prolong_correct(gc, ec, gf, ef, il_lo, il_lo + 1, st, cmask, hp);
        il_lo++;
    }
    if (launch_tile_prolong(gc, ec, gf, ef, il_lo, il_hi, cmask, st, hp))
        return;
    static const int wide = getenv("MGB_PROLONG_WIDE") ? atoi(getenv("MGB_PROLONG_WIDE")) : 1;
    if (wide || hp) {
        static const int occ8 = resident_blocks(k_prolong_correct8, 256, 0);
        const long long items = (long long)gf.nj * (gf.kh >> 2);
        const unsigned bx = (unsigned)((items + 255) / 256);
        const int nplanes = il_hi - il_lo;
        int nch = pick_chunks(bx, nplanes, occ8);
        int chunk = (nplanes + nch - 1) / nch;
        chunk += chunk & 1;  // whole (even, odd) pairs per chunk
        const unsigned by = (unsigned)((nplanes + chunk - 1) / chunk);
        launch_k(k_prolong_correct8, dim3(bx, by), 256, 0, st, gc, ec, gf, ef, il_lo, il_hi, chunk,
                                                         cmask, hp ? *hp : HaloCtl{});
        COUNT_LAUNCH();
        return;
    }
    static const int occ = resident_blocks(k_prolong_correct, 256, 0);
    const MarchCfg c = march_cfg(gf, il_hi - il_lo, 256, occ);
    launch_k(k_prolong_correct, c.grid, c.block, 0, st, gc, ec, gf, ef, il_lo, il_hi, c.chunk);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// Multi-GPU plumbing that is NOT fused into a compute kernel (halo.cuh has the
// fused part): explicit halo steps (uploads, fences, the odd stand-alone call),
// the all-gather of the first replicated level, the exchange of norm partials,
// and the kernel that closes an epoch.  One process per GPU; peers' memory is
// mapped through CUDA IPC and written with plain stores over NVLink.
// ----------------------------------------------------------------------------
struct HaloDir {  // one direction of an explicit halo step: up to four runs + the flag
    const double *src[4];
    double *dst[4];
    long long n[4];
    unsigned long long *peer_flag;  // nullptr: nothing goes this way
    unsigned int *count;            // local arrival counter
    unsigned long long off;         // sequence offset within the epoch
};

__global__ void __launch_bounds__(256)
k_halo_push(HaloDir up, HaloDir low, const unsigned long long *epoch)
{
    pdl_enter();
    // the first gridDim.x/2 blocks (or all, if only one direction is active)
    // serve `up`, the rest `low`
    const bool both = up.peer_flag && low.peer_flag;
    const unsigned int half = both ? gridDim.x / 2 : gridDim.x;
    const bool mine_up = up.peer_flag && (!both || blockIdx.x < half);
    const HaloDir &d = mine_up ? up : low;
    const unsigned int nb = both ? (mine_up ? half : gridDim.x - half) : gridDim.x;
    const unsigned int b = both && !mine_up ? blockIdx.x - half : blockIdx.x;
    const long long stride = (long long)nb * blockDim.x;
    const long long t0 = (long long)b * blockDim.x + threadIdx.x;
    // runs start on 16-byte boundaries and have even lengths (colour planes)
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const double2 *s = reinterpret_cast<const double2 *>(d.src[k]);
        double2 *o = reinterpret_cast<double2 *>(d.dst[k]);
        for (long long t = t0; 2 * t < d.n[k]; t += stride)
            o[t] = s[t];
    }
    __threadfence_system();  // this thread's peer stores are visible system-wide ...
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned int prev = atomicAdd(d.count, 1u);
        if (prev == nb - 1) {  // ... and the last block to get here has seen all of them
            *d.count = 0;
            __threadfence_system();
            halo_st_release(d.peer_flag, *epoch * kHaloEpochStride + d.off);
        }
    }
}

// thread 0 waits for the lower side, thread 1 for the upper one
__global__ void k_halo_wait(HaloCtl h)
{
    pdl_enter();
    if (threadIdx.x == 0)
        halo_wait_one(h, h.wait_low, 1u);
    else if (threadIdx.x == 1)
        halo_wait_one(h, h.wait_up, 2u);
}

// end of a collective operation: remember the last sends of this epoch (a wait issued
// before the first send of a later epoch refers to them), open the next epoch
__global__ void k_epoch_close(unsigned long long *xf, unsigned long long n_up,
                              unsigned long long n_low)
{
    pdl_enter();
    const unsigned long long e = xf[XF_EPOCH];
    if (n_up)
        xf[XF_PREV_UP] = e * kHaloEpochStride + n_up;
    if (n_low)
        xf[XF_PREV_LOW] = e * kHaloEpochStride + n_low;
    xf[XF_EPOCH] = e + 1;
}

// All-gather of the first replicated level's right-hand side: every rank has computed
// its slab of coarse planes (two runs: one per colour) and needs everybody else's.
// Three small kernels: "I am done reading the old contents" to every peer + wait for the
// same from every peer; copy my slab into every peer's array (many blocks per peer, the
// last one of a group tells the peer its copy is complete); wait for every peer's slab.
// Signals come before waits on every rank, so the handshakes cannot deadlock.
struct GatherArg {
    int me, nranks;
    const double *src[2];
    long long n[2];                       // doubles per run (even, 16-byte aligned)
    double *dst[kMaxRanks][2];            // [peer][run]
    unsigned long long *peer_xf[kMaxRanks];
    unsigned long long *my_xf;
    unsigned long long off;               // sequence offset of this gather within the epoch
    unsigned long long timeout_ns;
    unsigned int *err;
};

// phase 0 (one block, thread p <-> peer p): "I am done reading the old contents" to every
// peer, then wait for the same from every peer
__global__ void k_gather_ready(const GatherArg a)
{
    pdl_enter();
    const int p = threadIdx.x;
    if (p >= a.nranks || p == a.me)
        return;
    const unsigned long long v = a.my_xf[XF_EPOCH] * kHaloEpochStride + a.off;
    halo_st_release(a.peer_xf[p] + XF_READY + a.me, v);
    halo_spin(a.my_xf + XF_READY + p, v, a.timeout_ns, a.err, 4u);
}

// phase 1 (blockIdx.y <-> peer, gridDim.x blocks each): copy my slab into the peer's
// array; the last block of a peer's group to finish tells the peer its copy is complete
__global__ void __launch_bounds__(256) k_gather_copy(const GatherArg a)
{
    pdl_enter();
    int peer = blockIdx.y;
    if (peer >= a.me)
        peer++;
    const long long stride = (long long)gridDim.x * blockDim.x;
    const long long t0 = (long long)blockIdx.x * blockDim.x + threadIdx.x;
#pragma unroll
    for (int k = 0; k < 2; k++) {
        const double2 *s = reinterpret_cast<const double2 *>(a.src[k]);
        double2 *o = reinterpret_cast<double2 *>(a.dst[peer][k]);
        for (long long t = t0; 2 * t < a.n[k]; t += stride)
            o[t] = s[t];
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        unsigned int *cnt = reinterpret_cast<unsigned int *>(a.my_xf + XF_GATHER_CNT + peer);
        if (atomicAdd(cnt, 1u) == gridDim.x - 1) {
            *cnt = 0;
            __threadfence_system();
            halo_st_release(a.peer_xf[peer] + XF_DATA + a.me,
                            a.my_xf[XF_EPOCH] * kHaloEpochStride + a.off);
        }
    }
}

// phase 2 (one block, thread p <-> peer p): wait for every peer's slab
__global__ void k_gather_wait(const GatherArg a)
{
    pdl_enter();
    const int p = threadIdx.x;
    if (p >= a.nranks || p == a.me)
        return;
    halo_spin(a.my_xf + XF_DATA + p, a.my_xf[XF_EPOCH] * kHaloEpochStride + a.off, a.timeout_ns,
              a.err, 4u);
}

// Sum of one device scalar over the ranks, in RANK ORDER on every rank (bitwise the same
// everywhere and from run to run): thread p hands my partial to peer p and waits for its.
struct NormArg {
    int me, nranks;
    double *scalar;                        // in: my partial, out: the sum
    unsigned long long *peer_xf[kMaxRanks];
    unsigned long long *my_xf;
    unsigned long long off;
    unsigned long long timeout_ns;
    unsigned int *err;
};

__global__ void k_norm_exchange(const NormArg a)
{
    pdl_enter();
    const int p = threadIdx.x;
    const unsigned long long e = a.my_xf[XF_EPOCH];
    const unsigned long long v = e * kHaloEpochStride + a.off;
    const int slot = XF_NORMPART + (int)((e + a.off) & 1) * kMaxRanks;  // alternate buffers
    if (p < a.nranks && p != a.me) {
        double *parts = reinterpret_cast<double *>(a.peer_xf[p] + slot);
        parts[a.me] = *a.scalar;
        __threadfence_system();
        halo_st_release(a.peer_xf[p] + XF_NORMFLAG + a.me, v);
        halo_spin(a.my_xf + XF_NORMFLAG + p, v, a.timeout_ns, a.err, 8u);
    }
    __syncthreads();
    if (p == 0) {
        const double *parts = reinterpret_cast<const double *>(a.my_xf + slot);
        double sum = 0.;
        for (int r = 0; r < a.nranks; r++)
            sum = __dadd_rn(sum, r == a.me ? *a.scalar : parts[r]);
        *a.scalar = sum;
    }
}

void launch_halo_push(const HaloRun &up, const HaloRun &low, const unsigned long long *epoch,
                      cudaStream_t st)
{
    if (!up.peer_flag && !low.peer_flag)
        return;
    HaloDir a{}, b{};
    long long most = 0;
    const HaloRun *in[2] = {&up, &low};
    HaloDir *out[2] = {&a, &b};
    for (int i = 0; i < 2; i++) {
        for (int k = 0; k < 4; k++) {
            out[i]->src[k] = in[i]->src[k];
            out[i]->dst[k] = in[i]->dst[k];
            out[i]->n[k] = in[i]->n[k];
            if (in[i]->peer_flag && in[i]->n[k] > most)
                most = in[i]->n[k];
        }
        out[i]->peer_flag = in[i]->peer_flag;
        out[i]->count = in[i]->count;
        out[i]->off = in[i]->off;
    }
    long long nb = (most / 2 + 256 * 8 - 1) / (256 * 8);  // ~8 x 16 B per thread and direction
    if (nb > 64)
        nb = 64;
    if (nb < 1)
        nb = 1;
    if (up.peer_flag && low.peer_flag)
        nb *= 2;
    launch_k(k_halo_push, (unsigned)nb, 256, 0, st, a, b, epoch);
    COUNT_LAUNCH();
}

void launch_halo_wait(const HaloCtl &h, cudaStream_t st)
{
    if (!h.epoch || (!h.wait_low.flag && !h.wait_up.flag))
        return;
    launch_k(k_halo_wait, 1, 2, 0, st, h);
    COUNT_LAUNCH();
}

void launch_epoch_close(unsigned long long *xf, unsigned long long n_up, unsigned long long n_low,
                        cudaStream_t st)
{
    launch_k(k_epoch_close, 1, 1, 0, st, xf, n_up, n_low);
    COUNT_LAUNCH();
}

void launch_gather(const GatherHost &g, cudaStream_t st)
{
    GatherArg a{};
    a.me = g.me;
    a.nranks = g.nranks;
    for (int k = 0; k < 2; k++) {
        a.src[k] = g.src[k];
        a.n[k] = g.n[k];
    }
    for (int r = 0; r < g.nranks && r < kMaxRanks; r++) {
        a.dst[r][0] = g.dst[r][0];
        a.dst[r][1] = g.dst[r][1];
        a.peer_xf[r] = g.peer_xf[r];
    }
    a.my_xf = g.my_xf;
    a.off = g.off;
    a.timeout_ns = g.timeout_ns;
    a.err = g.err;
    long long most = a.n[0] > a.n[1] ? a.n[0] : a.n[1];
    long long nb = (most / 2 + 256 * 4 - 1) / (256 * 4);  // ~4 x 16 B per thread and run
    if (nb > 32)
        nb = 32;
    if (nb < 1)
        nb = 1;
    launch_k(k_gather_ready, 1, 32, 0, st, a);
    launch_k(k_gather_copy, dim3((unsigned)nb, g.nranks - 1), 256, 0, st, a);
    launch_k(k_gather_wait, 1, 32, 0, st, a);
    g_launches += 3;
}

void launch_norm_exchange(const NormHost &g, cudaStream_t st)
{
    NormArg a{};
    a.me = g.me;
    a.nranks = g.nranks;
    a.scalar = g.scalar;
    for (int r = 0; r < g.nranks && r < kMaxRanks; r++)
        a.peer_xf[r] = g.peer_xf[r];
    a.my_xf = g.my_xf;
    a.off = g.off;
    a.timeout_ns = g.timeout_ns;
    a.err = g.err;
    launch_k(k_norm_exchange, 1, 32, 0, st, a);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// updateEdgeValues (mg_3d.h:304-430): every inner point of the 12 edges becomes
// the mean of its two inward (face) neighbours, then every corner the mean of its
// three edge neighbours (added in k, j, i order).  Edge updates only read face
// points, corners read the updated edges: two launches.  Single-GPU levels.
// ----------------------------------------------------------------------------
__device__ __forceinline__ long long split_at(const Geo &g, int il, int j, int k)
{
    const int c = (g.i0 + il + j + k) & 1;
    return (long long)c * g.cs + ((long long)il * g.nj + j) * g.kh + (k >> 1);
}

__global__ void __launch_bounds__(128) k_edge_values(Geo g, double *__restrict__ a)
{
    // blockIdx.y: edge 0..11 = axis (0: along i, 1: along j, 2: along k) x the four
    // (end, end) combinations of the other two axes
    const int e = blockIdx.y, axis = e >> 2, eb = (e >> 1) & 1, ec = e & 1;
    const int n[3] = {g.ni, g.nj, g.nk};
    const int t = 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (t > n[axis] - 2)
        return;
    const int b = (axis + 1) % 3, c = (axis + 2) % 3;
    int p[3], q1[3], q2[3];
    p[axis] = t;
    p[b] = eb ? n[b] - 1 : 0;
    p[c] = ec ? n[c] - 1 : 0;
    for (int x = 0; x < 3; x++)
        q1[x] = q2[x] = p[x];
    q1[b] += eb ? -1 : 1;  // inward along b
    q2[c] += ec ? -1 : 1;  // inward along c
    // 0.5 * (u[inward neighbour 1] + u[inward neighbour 2]) (mg_3d.h:311-400); a sum of
    // two terms does not depend on their order, bit for bit
    const double s = __dadd_rn(a[split_at(g, q1[0], q1[1], q1[2])],
                               a[split_at(g, q2[0], q2[1], q2[2])]);
    a[split_at(g, p[0], p[1], p[2])] = __dmul_rn(0.5, s);
}

__global__ void k_corner_values(Geo g, double *__restrict__ a)
{
    const int t = threadIdx.x;  // 8 corners
    if (t >= 8)
        return;
    const int i = (t & 4) ? g.ni - 1 : 0, j = (t & 2) ? g.nj - 1 : 0, k = (t & 1) ? g.nk - 1 : 0;
    const int ii = i ? i - 1 : 1, jj = j ? j - 1 : 1, kk = k ? k - 1 : 1;
    // (1./3) * (u[pos+-1] + u[pos+-N] + u[pos+-NN]), left to right (mg_3d.h:405-429)
    double s = __dadd_rn(a[split_at(g, i, j, kk)], a[split_at(g, i, jj, k)]);
    s = __dadd_rn(s, a[split_at(g, ii, j, k)]);
    a[split_at(g, i, j, k)] = __dmul_rn(1. / 3, s);
}

void launch_edge_values(const Geo &g, double *a, cudaStream_t st)
{
    int most = g.ni > g.nj ? g.ni : g.nj;
    if (g.nk > most)
        most = g.nk;
    k_edge_values<<<dim3((most + 127) / 128, 12), 128, 0, st>>>(g, a);
    k_corner_values<<<1, 32, 0, st>>>(g, a);
    g_launches += 2;
}

// ----------------------------------------------------------------------------
// reductions over whole arrays
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_sumsq(const double *__restrict__ a, long long n, double *__restrict__ partials)
{
    double acc = 0.;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n;
         t += stride) {
        const double x = a[t];
        acc = __dadd_rn(acc, __dmul_rn(x, x));
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0)
        partials[blockIdx.x] = acc;
}

void launch_sumsq(const double *a0, long long n0, const double *a1, long long n1,
                  double *partials, double *out, cudaStream_t st)
{
    long long nb = (n0 + 255) / 256;
    if (nb > 148 * 16)
        nb = 148 * 16;
    if (nb < 1)
        nb = 1;
    k_sumsq<<<(unsigned)nb, 256, 0, st>>>(a0, n0, partials);
    COUNT_LAUNCH();
    int total = (int)nb;
    if (a1 && n1 > 0) {
        k_sumsq<<<(unsigned)nb, 256, 0, st>>>(a1, n1, partials + nb);
        COUNT_LAUNCH();
        total += (int)nb;
    }
    launch_k(k_finish_sum, 1, 1024, 0, st, partials, total, out);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256)
k_error_sumsq(Geo g, const double *__restrict__ u, double h, int il_lo, int il_hi,
              double *__restrict__ partials)
{
    double acc = 0.;
    const long long first = (long long)il_lo * g.pj, total = (long long)il_hi * g.pj;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = first + (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += stride) {
        const int m = (int)(t % g.kh);
        const long long row = t / g.kh;
        const int j = (int)(row % g.nj);
        const int il = (int)(row / g.nj);
        const int ig = g.i0 + il;
        const int s = (ig + j) & 1;
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int k = 2 * m + (c ^ s);
            if (k < g.nk) {
                const double ex = bc_func(__dmul_rn((double)ig, h), __dmul_rn((double)j, h),
                                          __dmul_rn((double)k, h));
                const double df = __dsub_rn(u[(long long)c * g.cs + t], ex);
                acc = __dadd_rn(acc, __dmul_rn(df, df));
            }
        }
    }
    acc = block_sum(acc);
    if (threadIdx.x == 0)
        partials[blockIdx.x] = acc;
}

void launch_error_sumsq(const Geo &g, const double *u, double h, int il_lo, int il_hi,
                        double *partials, double *out, cudaStream_t st)
{
    const long long total = (long long)(il_hi - il_lo) * g.pj;
    long long nb = (total + 255) / 256;
    if (nb > 148 * 16)
        nb = 148 * 16;
    if (nb < 1)
        nb = 1;
    k_error_sumsq<<<(unsigned)nb, 256, 0, st>>>(g, u, h, il_lo, il_hi, partials);
    COUNT_LAUNCH();
    launch_k(k_finish_sum, 1, 1024, 0, st, partials, (int)nb, out);
    COUNT_LAUNCH();
}

}  // namespace mgb
