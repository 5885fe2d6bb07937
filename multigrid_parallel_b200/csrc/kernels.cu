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

#include "kernels.h"

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

// mg_3d.h:89-90: x*x - 2*y*y + z*z
__device__ __forceinline__ double bc_func(double x, double y, double z)
{
    double a = __dmul_rn(x, x);
    double b = __dmul_rn(__dmul_rn(2.0, y), y);
    double c = __dmul_rn(z, z);
    return __dadd_rn(__dsub_rn(a, b), c);
}

// value of a colour-split array at local plane il, row j, column k
__device__ __forceinline__ double rd_split(const Geo &g, const double *a, int il,
                                           int j, int k)
{
    const int c = (g.i0 + il + j + k) & 1;
    return a[c * g.cs + ((long long)il * g.nj + j) * g.kh + (k >> 1)];
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

static MarchCfg march_cfg(const Geo &g, int nplanes, int threads, int sm_blocks)
{
    MarchCfg c;
    const long long pairs = (long long)g.pj / 2;
    const unsigned bx = (unsigned)((pairs + threads - 1) / threads);
    // aim at ~2 waves of resident blocks over the 148 SMs
    long long want = 2LL * 148 * sm_blocks;
    long long nch = (want + bx - 1) / bx;
    if (nch < 1)
        nch = 1;
    if (nch > nplanes)
        nch = nplanes;
    c.chunk = (int)((nplanes + nch - 1) / nch);
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

void launch_half_sweep(const Geo &g, double *v, const double *d, double hSq,
                       int colour, int il_lo, int il_hi, cudaStream_t st)
{
    if (il_hi <= il_lo)
        return;
    const MarchCfg c = march_cfg(g, il_hi - il_lo, 256, 8);
    if (colour)
        k_half_sweep<1><<<c.grid, c.block, 0, st>>>(g, v, v + g.cs, d + g.cs, hSq,
                                                    il_lo, il_hi, c.chunk);
    else
        k_half_sweep<0><<<c.grid, c.block, 0, st>>>(g, v + g.cs, v, d, hSq, il_lo,
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
           double *__restrict__ partials)
{
    const int npair = g.kh >> 1;
    const long long q = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    double acc = 0.;
    const int j = (int)(q / npair);
    const int ia = il_lo + blockIdx.y * chunk;
    const int ib = min(ia + chunk, il_hi);
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
                     double *out_sumsq, cudaStream_t st)
{
    if (il_hi <= il_lo) {
        cudaMemsetAsync(out_sumsq, 0, sizeof(double), st);
        return;
    }
    MarchCfg c = march_cfg(g, il_hi - il_lo, 256, 4);
    while ((long long)c.grid.x * c.grid.y > kMaxPartials) {  // coarser chunks
        c.chunk *= 2;
        c.grid.y = (il_hi - il_lo + c.chunk - 1) / c.chunk;
    }
    if (r)
        k_residual<true><<<c.grid, c.block, 0, st>>>(g, v, d, r, invHsq, il_lo, il_hi,
                                                     c.chunk, partials);
    else
        k_residual<false><<<c.grid, c.block, 0, st>>>(g, v, d, nullptr, invHsq, il_lo,
                                                      il_hi, c.chunk, partials);
    COUNT_LAUNCH();
    k_finish_sum<<<1, 1024, 0, st>>>(partials, (int)(c.grid.x * c.grid.y), out_sumsq);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// full-weighting restriction (mg_3d.h:844-998): one thread per coarse entry
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_restrict(Geo gf, const double *__restrict__ rf, Geo gc, double *__restrict__ dc,
           int Il_lo, int Il_hi)
{
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
    k_restrict<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gf, rf, gc, dc, Il_lo,
                                                               Il_hi);
    COUNT_LAUNCH();
}

// ----------------------------------------------------------------------------
// prolongation + correction (mg_3d.h:1000-1145): one thread per fine entry
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
k_prolong_correct(Geo gc, const double *__restrict__ ec, Geo gf,
                  double *__restrict__ ef, int il_lo, int il_hi)
{
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long per_colour = (long long)(il_hi - il_lo) * gf.pj;
    if (t >= 2 * per_colour)
        return;
    const int c = t >= per_colour;
    const long long e = t - c * per_colour;
    const int m = (int)(e % gf.kh);
    const long long row = e / gf.kh;
    const int j = (int)(row % gf.nj);
    const int il = il_lo + (int)(row / gf.nj);
    const int ig = gf.i0 + il;
    const int k = 2 * m + ((c ^ (ig + j)) & 1);
    if (k >= gf.nk)
        return;
    const int oi = ig & 1, oj = j & 1, ok = k & 1;
    const int I = (ig >> 1) - gc.i0, J = j >> 1, K = k >> 1;
#define EC(a, b, cc_) rd_split(gc, ec, I + (a), J + (b), K + (cc_))
    double add;
    const int nodd = oi + oj + ok;
    if (nodd == 3) {  // 1023-1049: i-major, then j, then k
        add = __dadd_rn(0., EC(0, 0, 0));
        add = __dadd_rn(add, EC(0, 0, 1));
        add = __dadd_rn(add, EC(0, 1, 0));
        add = __dadd_rn(add, EC(0, 1, 1));
        add = __dadd_rn(add, EC(1, 0, 0));
        add = __dadd_rn(add, EC(1, 0, 1));
        add = __dadd_rn(add, EC(1, 1, 0));
        add = __dadd_rn(add, EC(1, 1, 1));
        add = __dmul_rn(add, 0.125);
    } else if (nodd == 2) {
        if (!oi) {  // 1059-1068: j fastest, then k
            add = __dadd_rn(0., EC(0, 0, 0));
            add = __dadd_rn(add, EC(0, 1, 0));
            add = __dadd_rn(add, EC(0, 0, 1));
            add = __dadd_rn(add, EC(0, 1, 1));
        } else if (!oj) {  // 1070-1079: i fastest, then k
            add = __dadd_rn(0., EC(0, 0, 0));
            add = __dadd_rn(add, EC(1, 0, 0));
            add = __dadd_rn(add, EC(0, 0, 1));
            add = __dadd_rn(add, EC(1, 0, 1));
        } else {  // 1080-1089: j fastest, then i
            add = __dadd_rn(0., EC(0, 0, 0));
            add = __dadd_rn(add, EC(0, 1, 0));
            add = __dadd_rn(add, EC(1, 0, 0));
            add = __dadd_rn(add, EC(1, 1, 0));
        }
        add = __dmul_rn(add, 0.25);
    } else if (nodd == 1) {  // 1101-1134: low end, high end
        add = __dadd_rn(0., EC(0, 0, 0));
        add = __dadd_rn(add, EC(oi, oj, ok));
        add = __dmul_rn(add, 0.5);
    } else {  // 1137-1138
        add = EC(0, 0, 0);
    }
#undef EC
    const long long idx = (long long)c * gf.cs + ((long long)il * gf.nj + j) * gf.kh + m;
    ef[idx] = __dadd_rn(ef[idx], add);
}

void launch_prolong_correct(const Geo &gc, const double *ec, const Geo &gf,
                            double *ef, int il_lo, int il_hi, cudaStream_t st)
{
    if (il_hi <= il_lo)
        return;
    const long long total = 2LL * (il_hi - il_lo) * gf.pj;
    k_prolong_correct<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(gc, ec, gf, ef,
                                                                      il_lo, il_hi);
    COUNT_LAUNCH();
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

void launch_sumsq(const double *a, long long n, double *partials, double *out,
                  cudaStream_t st)
{
    long long nb = (n + 255) / 256;
    if (nb > 148 * 16)
        nb = 148 * 16;
    if (nb < 1)
        nb = 1;
    k_sumsq<<<(unsigned)nb, 256, 0, st>>>(a, n, partials);
    COUNT_LAUNCH();
    k_finish_sum<<<1, 1024, 0, st>>>(partials, (int)nb, out);
    COUNT_LAUNCH();
}

__global__ void __launch_bounds__(256)
k_error_sumsq(Geo g, const double *__restrict__ u, double h,
              double *__restrict__ partials)
{
    double acc = 0.;
    const long long total = (long long)g.li * g.pj;
    const long long stride = (long long)gridDim.x * blockDim.x;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
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

void launch_error_sumsq(const Geo &g, const double *u, double h, double *partials,
                        double *out, cudaStream_t st)
{
    const long long total = (long long)g.li * g.pj;
    long long nb = (total + 255) / 256;
    if (nb > 148 * 16)
        nb = 148 * 16;
    k_error_sumsq<<<(unsigned)nb, 256, 0, st>>>(g, u, h, partials);
    COUNT_LAUNCH();
    k_finish_sum<<<1, 1024, 0, st>>>(partials, (int)nb, out);
    COUNT_LAUNCH();
}

}  // namespace mgb
