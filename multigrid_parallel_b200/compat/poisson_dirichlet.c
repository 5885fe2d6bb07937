/*
 * poisson_dirichlet.c -- repo-owned example driver for the drop-in mg_3d.h.
 *
 * Same flow as the reference's test_mg_3d.c (set-up, OpenMP-parallel V-cycle
 * loop to 1e-8*||rhs||, timing table, error against the analytic solution),
 * printed with full precision, plus the parts of the API that driver does not
 * reach: SolverGetResidual, SolverResetTimingInfo, the raw-pointer smoother on
 * caller-owned arrays, and a host write into `grid` between two solves (the
 * page-protection coherence of the host mirrors).
 *
 *   poisson_dirichlet <coarse> <levels> <gs>
 */
#include <stdio.h>
#include <string.h>

#define GRID_LENGTH (1.)
#include "mg_3d.h"
#include "postprocess.h"

static int solve_to(double cmpNorm, double *threadNorm, int numThreads, int print)
{
    double norm = 1e9;
    int cycles = 0;
#pragma omp parallel
    {
        const int tid = omp_get_thread_num();
        while (norm > cmpNorm) {
            threadNorm[tid] = SolverLinSolve();
#pragma omp barrier
#pragma omp single
            {
                double s = 0;
                for (int t = 0; t < numThreads; t++)
                    s += threadNorm[t] * threadNorm[t];
                norm = sqrt(s);
                cycles++;
                if (print)
                    printf("cycle %d norm %.17g\n", cycles, norm);
            }
        }
    }
    return cycles;
}

int main(int argc, char **argv)
{
    SolverInitialize(argc, argv);
    double *grid = NULL, *rhs = NULL, h;
    const int N = SolverGetDetails(&grid, &rhs, &h);
    SolverSetupBoundaryConditions();
    const int numThreads = omp_get_max_threads();
    double *threadNorm = calloc(numThreads, sizeof(double));
    const double initResidual = SolverGetInitialResidual();
    setupBoundaryConditions(grid, N, h);
    printf("N %d h %.17g init %.17g\n", N, h, initResidual);
    printf("true_init %.17g\n", SolverGetResidual());

    const int cycles = solve_to(initResidual * 1e-8, threadNorm, numThreads, 1);
    printf("cycles %d\n", cycles);
    printf("residual_after %.17g\n", SolverGetResidual());
    SolverPrintTimingInfo();

    /* error against the analytic solution (test_mg_3d.c:78-97), grid untouched */
    double err = 0., sumsq = 0.;
    for (int i = 0; i < N; i++)
        for (int j = 0; j < N; j++)
            for (int k = 0; k < N; k++) {
                const double v = grid[((size_t)i * N + j) * N + k];
                const double diff = v - BCFunc(i * h, j * h, k * h);
                err += diff * diff;
                sumsq += v * v;
            }
    printf("errnorm %.17g sumsq %.17g probe %.17g\n", sqrt(err), sumsq,
           grid[((size_t)1 * N + 2) * N + 3]);

    /* perturb the solution on the host, solve again: the write must reach the GPU */
    const size_t mid = ((size_t)(N / 2) * N + N / 2) * N + N / 2;
    grid[mid] += 1.0;
    printf("perturbed_residual %.17g\n", SolverGetResidual());
    SolverResetTimingInfo();
    const int cycles2 = solve_to(initResidual * 1e-8, threadNorm, numThreads, 0);
    printf("cycles_after_perturbation %d calls_finest_smoother1 %d\n", cycles2,
           tInfo[numLevels - 1]->numCalls[0]);
    printf("restored %.3e\n", fabs(grid[mid] - BCFunc((N / 2) * h, (N / 2) * h, (N / 2) * h)));

    /* raw-pointer smoother + residual on caller-owned arrays (test_rb_gs_3d.c flow) */
    const int M = 17;
    const double hm = GRID_LENGTH / (M - 1);
    double *uu = calloc((size_t)M * M * M, sizeof(double));
    double *dd = calloc((size_t)M * M * M, sizeof(double));
    setupBoundaryConditions(uu, M, hm);
    double part[64] = {0}, nrm = 0;
    const double init2 = calculateResidual(uu, dd, M, hm, NULL);
#pragma omp parallel
    {
        preSmoother(uu, dd, M, hm, 1);
        postSmoother(uu, dd, M, hm, 1);
        part[omp_get_thread_num() % 64] = calculateResidual(uu, dd, M, hm, NULL);
    }
    for (int t = 0; t < 64; t++)
        nrm += part[t] * part[t];
    printf("rbgs17 init %.17g after %.17g\n", init2, sqrt(nrm));
    free(uu);
    free(dd);

    if (getenv("MGB_WRITE_VTK"))
        writeOutputData(getenv("MGB_WRITE_VTK"), grid, h, N);
    SolverFinalize();
    free(threadNorm);
    return 0;
}
