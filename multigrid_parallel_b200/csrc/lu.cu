// lu.cu -- dense coarsest-level operator on the GPU: build, factor once
// (Doolittle, no pivoting), solve by column sweeps.
//
// Bit-exactness contract (gauss_elim.h:9-60): the factorisation updates every
// entry with the pivots in ascending order, multiplier = a_ki * (1/a_ii); the
// forward sum of row i runs over ascending j, the backward sum over DESCENDING
// j, each starting from 0.  A column sweep (finish x_j, then add column j into
// all pending row sums) performs exactly that sequence per row while exposing
// n-way parallelism per step.
#include "kernels.h"

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

// pivot p, step 1: multipliers  a[r][p] *= 1/a[p][p]   (r > p)
__global__ void __launch_bounds__(256) k_lu_scale(double *__restrict__ a, int n, int p)
{
    const int r = p + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= n)
        return;
    const double pinv = __drcp_rn(a[(long long)p * n + p]);
    a[(long long)r * n + p] = __dmul_rn(a[(long long)r * n + p], pinv);
}

// pivot p, step 2: a[r][c] -= a[r][p]*a[p][c]   (r > p, c > p)
__global__ void __launch_bounds__(256) k_lu_update(double *__restrict__ a, int n, int p)
{
    const int c = p + 1 + blockIdx.x * blockDim.x + threadIdx.x;
    const int r = p + 1 + blockIdx.y;
    if (c >= n)
        return;
    const double z = a[(long long)r * n + p];
    const double t = __dmul_rn(z, a[(long long)p * n + c]);
    a[(long long)r * n + c] = __dsub_rn(a[(long long)r * n + c], t);
}

void launch_lu_factor(double *a, int n, cudaStream_t st)
{
    for (int p = 0; p < n - 1; p++) {
        const int rem = n - 1 - p;
        k_lu_scale<<<(rem + 255) / 256, 256, 0, st>>>(a, n, p);
        dim3 grid((rem + 255) / 256, rem);
        k_lu_update<<<grid, 256, 0, st>>>(a, n, p);
        *launch_counter() += 2;
    }
}

__global__ void k_transpose(const double *__restrict__ a, double *__restrict__ at, int n)
{
    __shared__ double tile[32][33];
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 32 + threadIdx.y;
    if (x < n && y < n)
        tile[threadIdx.y][threadIdx.x] = a[(long long)y * n + x];
    __syncthreads();
    x = blockIdx.y * 32 + threadIdx.x;
    y = blockIdx.x * 32 + threadIdx.y;
    if (x < n && y < n)
        at[(long long)y * n + x] = tile[threadIdx.x][threadIdx.y];
}

void launch_transpose(const double *a, double *at, int n, cudaStream_t st)
{
    dim3 grid((n + 31) / 32, (n + 31) / 32), block(32, 32);
    k_transpose<<<grid, block, 0, st>>>(a, at, n);
    ++*launch_counter();
}

// One block.  Thread t owns rows t, t+B, t+2B, ... (at most RPT of them) and
// keeps their running sums in registers; xs[] (shared) carries z, then x.
// lut[c*n + r] = lu[r*n + c], so step c reads one contiguous row of lut.
template <int RPT>
__global__ void __launch_bounds__(1024)
k_lu_solve(const double *__restrict__ lu, const double *__restrict__ lut, int n,
           const double *__restrict__ b, double *__restrict__ x)
{
    extern __shared__ double xs[];
    const int t = threadIdx.x, B = blockDim.x;
    double sum[RPT];
#pragma unroll
    for (int s = 0; s < RPT; s++)
        sum[s] = 0.;
    // forward: L z = b, unit lower triangle, ascending columns
    for (int c = 0; c < n; c++) {
        if (c % B == t) {
            const int sl = c / B;
            double sv = 0.;
#pragma unroll
            for (int s = 0; s < RPT; s++)
                if (s == sl)
                    sv = sum[s];
            xs[c] = __dsub_rn(b[c], sv);
        }
        __syncthreads();
        const double xc = xs[c];
        const double *col = lut + (long long)c * n;
#pragma unroll
        for (int s = 0; s < RPT; s++) {
            const int r = t + s * B;
            if (r > c && r < n)
                sum[s] = __dadd_rn(sum[s], __dmul_rn(col[r], xc));
        }
    }
#pragma unroll
    for (int s = 0; s < RPT; s++)
        sum[s] = 0.;
    __syncthreads();
    // backward: U x = z, descending columns
    for (int c = n - 1; c >= 0; c--) {
        if (c % B == t) {
            const int sl = c / B;
            double sv = 0.;
#pragma unroll
            for (int s = 0; s < RPT; s++)
                if (s == sl)
                    sv = sum[s];
            xs[c] = __ddiv_rn(__dsub_rn(xs[c], sv), lu[(long long)c * n + c]);
        }
        __syncthreads();
        const double xc = xs[c];
        const double *col = lut + (long long)c * n;
#pragma unroll
        for (int s = 0; s < RPT; s++) {
            const int r = t + s * B;
            if (r < c)
                sum[s] = __dadd_rn(sum[s], __dmul_rn(col[r], xc));
        }
    }
    __syncthreads();
    for (int r = t; r < n; r += B)
        x[r] = xs[r];
}

void launch_lu_solve(const double *lu, const double *lut, int n, const double *b,
                     double *x, cudaStream_t st)
{
    int B = 32;
    while (B < n && B < 1024)
        B *= 2;
    const int rpt = (n + B - 1) / B;
    const size_t sh = sizeof(double) * n;
    if (sh > 48 * 1024) {  // opt in to large dynamic shared memory (n <= 8192)
        cudaFuncSetAttribute(k_lu_solve<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        cudaFuncSetAttribute(k_lu_solve<4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
        cudaFuncSetAttribute(k_lu_solve<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sh);
    }
    if (rpt <= 1)
        k_lu_solve<1><<<1, B, sh, st>>>(lu, lut, n, b, x);
    else if (rpt <= 2)
        k_lu_solve<2><<<1, B, sh, st>>>(lu, lut, n, b, x);
    else if (rpt <= 4)
        k_lu_solve<4><<<1, B, sh, st>>>(lu, lut, n, b, x);
    else
        k_lu_solve<8><<<1, B, sh, st>>>(lu, lut, n, b, x);
    ++*launch_counter();
}

}  // namespace mgb
