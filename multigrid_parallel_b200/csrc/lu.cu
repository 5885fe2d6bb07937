// lu.cu -- coarsest-level operator on the GPU: build, factor once (Doolittle,
// no pivoting), solve once per cycle -- all restricted to the BAND of the
// operator (half bandwidth nj*nk), bit-identical to the reference's dense loops.
//
// Bit-exactness contract (gauss_elim.h:9-60): the factorisation updates every
// entry with the pivots in ascending order, multiplier = a_ki * (1/a_ii); the
// forward sum of row i runs over ascending j, the backward sum over DESCENDING
// j, each starting from 0.
//
// Why the band is enough.  Without pivoting the fill of L and U stays inside
// the band of A.  Outside it the reference computes z = (+0)*(1/a_ii) = +-0 and
// a[k][j] -= (+-0)*a[i][j]; an entry of the trailing matrix is never -0 (it is
// +0 from the memset or the result of a subtraction, and x - x = +0), so that
// update changes nothing, bit for bit.  The multipliers themselves are stored:
// out of the band they are copysign(0, 1/a_ii), which a fix-up pass writes so
// that even the sign of zero of the dense factor (mgb_coarse_lu_download, the
// drop-in's `A`) equals the reference's.  The solve: see lu_band.cuh.
#include <cstdio>

#include "kernels.h"
#include "launch.h"
#include "lu_band.cuh"
#include "devmath.cuh"

namespace mgb {

extern long long *launch_counter();

// mg_3d.h:147-273: identity rows on the boundary, 7-point/h^2 rows inside
__global__ void __launch_bounds__(256)
k_coarse_matrix(double *__restrict__ A, int ni, int nj, int nk, double one, double six)
{
    const long long n = (long long)ni * nj * nk;
    const long long p = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n)
        return;
    const int k = (int)(p % nk);
    const int j = (int)((p / nk) % nj);
    const int i = (int)(p / ((long long)nk * nj));
    double *row = A + p * n;
    if (i == 0 || i == ni - 1 || j == 0 || j == nj - 1 || k == 0 || k == nk - 1) {
        row[p] = 1.;
    } else {
        const long long sj = nk, si = (long long)nj * nk;
        row[p - si] = one; row[p + si] = one;
        row[p - sj] = one; row[p + sj] = one;
        row[p - 1] = one;  row[p + 1] = one;
        row[p] = -six;
    }
}

void launch_coarse_matrix(double *A, int ni, int nj, int nk, double h, cudaStream_t st)
{
    const long long n = (long long)ni * nj * nk;
    cudaMemsetAsync(A, 0, sizeof(double) * n * n, st);
    const double invHsq = 1. / (h * h);  // mg_3d.h:155-159
    k_coarse_matrix<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(A, ni, nj, nk,
                                                                 1. * invHsq, 6. * invHsq);
    ++*launch_counter();
}

// ----------------------------------------------------------------------------
// factorisation, band-limited, ONE launch (gauss_elim.h:9-29)
//
// One block; per pivot p the window rows/columns p+1 .. min(p+bw, n-1) get
// a[r][c] -= z_r * a[p][c] with z_r = a[r][p] * (1/a[p][p]) recomputed by every
// thread that needs it (the same product, hence the same bits); one barrier per
// pivot.  Column p is left UNSCALED -- nothing reads it again during the
// factorisation -- and k_lu_finish_lower turns it into the stored multipliers.
// ----------------------------------------------------------------------------
__global__ void __launch_bounds__(1024) k_lu_factor_band(double *a, int n, int bw)
{
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int p = 0; p < n - 1; p++) {
        const int w = min(bw, n - 1 - p);
        const double pinv = __drcp_rn(a[(size_t)p * n + p]);  // 1./a[ni+i]
        const double *prow = a + (size_t)p * n + p + 1;
        for (int rr = ty; rr < w; rr += 32) {
            double *row = a + (size_t)(p + 1 + rr) * n + p;
            const double z = __dmul_rn(row[0], pinv);
            for (int cc = tx; cc < w; cc += 32)
                row[1 + cc] = __dsub_rn(row[1 + cc], __dmul_rn(z, prow[cc]));
        }
        __syncthreads();
    }
}

// strictly lower triangle: multipliers inside the band, signed zeros outside
__global__ void __launch_bounds__(256) k_lu_finish_lower(double *a, int n, int bw)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    if (c >= r || c >= n)
        return;
    const double pinv = __drcp_rn(a[(size_t)c * n + c]);
    double *e = a + (size_t)r * n + c;
    *e = r - c <= bw ? __dmul_rn(*e, pinv) : __dmul_rn(0., pinv);
}

void launch_lu_factor_band(double *a, int n, int bw, cudaStream_t st)
{
    if (n < 2)
        return;
    if (bw > n - 1)
        bw = n - 1;
    if (bw < 0)
        bw = 0;
    k_lu_factor_band<<<1, 1024, 0, st>>>(a, n, bw);
    k_lu_finish_lower<<<dim3((n + 255) / 256, n), 256, 0, st>>>(a, n, bw);
    *launch_counter() += 2;
}

// half bandwidth of a dense matrix: max |r - c| over its non-zero entries
__global__ void __launch_bounds__(256) k_lu_bandwidth(const double *__restrict__ a, int n, int *out)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    const int r = blockIdx.y;
    int d = 0;
    if (c < n && a[(size_t)r * n + c] != 0.)
        d = r > c ? r - c : c - r;
    for (int o = 16; o > 0; o >>= 1)
        d = max(d, __shfl_down_sync(0xffffffffu, d, o));
    if ((threadIdx.x & 31) == 0 && d > 0)
        atomicMax(out, d);
}

int lu_bandwidth(const double *a, int n, cudaStream_t st)
{
    int *d_bw = nullptr, bw = n - 1;
    if (cudaMalloc(&d_bw, sizeof(int)) != cudaSuccess)
        return bw;
    cudaMemsetAsync(d_bw, 0, sizeof(int), st);
    k_lu_bandwidth<<<dim3((n + 255) / 256, n), 256, 0, st>>>(a, n, d_bw);
    ++*launch_counter();
    cudaMemcpyAsync(&bw, d_bw, sizeof(int), cudaMemcpyDeviceToHost, st);
    cudaStreamSynchronize(st);
    cudaFree(d_bw);
    return bw;
}

// dense factor -> 32 x 32 transposed tiles of lu_band.cuh (+ diagonal and its reciprocal)
__global__ void __launch_bounds__(256) k_lu_extract_tiles(const double *__restrict__ a, LuBand B)
{
    const int n = B.n, NT = B.nt, NB = (n + 31) >> 5;
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)NB * (NT + 1) * 1024;
    if (e < n) {
        const double dg = a[(size_t)e * n + e];
        B.ud[e] = dg;
        B.rd[e] = __drcp_rn(dg);
    }
    if (e >= total)
        return;
    const int l = (int)(e & 31), jj = (int)((e >> 5) & 31);
    const int t = (int)((e >> 10) % (NT + 1)), R = (int)((e >> 10) / (NT + 1));
    const int row = 32 * R + l;
    // L: column block R-NT+t (t = NT: the diagonal block, strictly lower part)
    {
        const int col = 32 * (R - NT + t) + jj;
        double v = 0.;
        if (row < n && col >= 0 && col < row)
            v = a[(size_t)row * n + col];
        B.lt[e] = v;
    }
    // U: column block R+NT-t (t = NT: the diagonal block, strictly upper part)
    {
        const int col = 32 * (R + NT - t) + jj;
        double v = 0.;
        if (row < n && col < n && col > row)
            v = a[(size_t)row * n + col];
        B.ut[e] = v;
    }
}

void launch_lu_extract_band(const double *a, const LuBand &B, cudaStream_t st)
{
    long long total = (long long)lu_tile_doubles(B.n, B.bw);
    if (total < B.n)
        total = B.n;
    k_lu_extract_tiles<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(a, B);
    ++*launch_counter();
}

// ----------------------------------------------------------------------------
// solve (gauss_elim.h:31-60), one launch: LEVEL form reads b from / writes x to
// the colour-split arrays of level 0 directly (the dense vectors of
// solveWithLU(LU, n, d[0], u[0]) are their natural-layout views, mg_3d.h:1270)
// ----------------------------------------------------------------------------
template <bool LEVEL>
__global__ void __launch_bounds__(32 * kLuWarps)
k_lu_band_solve(const LuBand B, const Geo g, const double *__restrict__ rhs, double *__restrict__ x)
{
    pdl_enter();
    extern __shared__ double lu_sh[];
    const int n = B.n, npad = (n + 31) & ~31;
    double *xs = lu_sh;
    int *flags = reinterpret_cast<int *>(xs + npad);  // done_f, done_b, redo
    const int t = threadIdx.x;
    auto load_b = [&]() {
        for (int i = t; i < npad; i += blockDim.x) {
            double b = 0.;
            if (i < n) {
                if (LEVEL) {
                    const int k = i % g.nk, j = (i / g.nk) % g.nj, il = i / (g.nk * g.nj);
                    b = rd_split(g, rhs, il, j, k);
                } else {
                    b = rhs[i];
                }
            }
            xs[i] = b;
        }
        if (t < 2)
            flags[t] = 0;
    };
    load_b();
    if (t == 2)
        flags[2] = 0;
    __syncthreads();
    if (!lu_band_solve<false>(B, xs, flags, t >> 5, t & 31))
        flags[2] = 1;
    __syncthreads();
    if (flags[2]) {  // a fast quotient was not the correctly rounded one: exact divisions
        __syncthreads();
        load_b();
        __syncthreads();
        lu_band_solve<true>(B, xs, flags, t >> 5, t & 31);
        __syncthreads();
    }
    for (int i = t; i < n; i += blockDim.x) {
        if (LEVEL) {
            const int k = i % g.nk, j = (i / g.nk) % g.nj, il = i / (g.nk * g.nj);
            const int c = (g.i0 + il + j + k) & 1;
            x[(long long)c * g.cs + ((long long)il * g.nj + j) * g.kh + (k >> 1)] = xs[i];
        } else {
            x[i] = xs[i];
        }
    }
}

template <bool LEVEL>
static void launch_band_solve(const LuBand &B, const Geo &g, const double *rhs, double *x,
                              cudaStream_t st)
{
    const size_t sh = sizeof(double) * lu_solve_smem_doubles(B.n) + 16;
    static size_t allowed[64] = {};
    if (grows_on_device(allowed, sh))
        cudaFuncSetAttribute(k_lu_band_solve<LEVEL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)sh);
    launch_k(k_lu_band_solve<LEVEL>, 1, 32 * kLuWarps, sh, st, B, g, rhs, x);
    ++*launch_counter();
}

void launch_lu_solve_dense(const LuBand &B, const double *b, double *x, cudaStream_t st)
{
    launch_band_solve<false>(B, Geo{}, b, x, st);
}

void launch_lu_solve_level(const LuBand &B, const Geo &g, const double *d0, double *u0,
                           cudaStream_t st)
{
    launch_band_solve<true>(B, g, d0, u0, st);
}

// storage of the tile form
int lu_band_alloc(LuBand *B, int n, int bw)
{
    B->n = n;
    B->bw = bw;
    B->nt = lu_num_tiles(bw);
    const size_t td = lu_tile_doubles(n, bw);
    if (cudaMalloc(&B->lt, sizeof(double) * td) != cudaSuccess ||
        cudaMalloc(&B->ut, sizeof(double) * td) != cudaSuccess ||
        cudaMalloc(&B->ud, sizeof(double) * n) != cudaSuccess ||
        cudaMalloc(&B->rd, sizeof(double) * n) != cudaSuccess)
        return 1;
    return 0;
}

void lu_band_free(LuBand *B)
{
    if (B->lt) cudaFree(B->lt);
    if (B->ut) cudaFree(B->ut);
    if (B->ud) cudaFree(B->ud);
    if (B->rd) cudaFree(B->rd);
    B->lt = B->ut = B->ud = B->rd = nullptr;
}

}  // namespace mgb
