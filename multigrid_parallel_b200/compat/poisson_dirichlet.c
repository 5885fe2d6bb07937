/*
 * poisson_dirichlet.c -- repo-owned example driver for the drop-in mg_3d.h.
 *
 * Same flow as the reference's test_mg_3d.c (set-up, OpenMP-parallel V-cycle
 * loop to 1e-8*||rhs||, timing table, error against the analytic solution),
 * printed with full precision, plus the parts of the API that driver does not
 * reach: SolverGetResidual, SolverResetTimingInfo, SolverSmoothenEdgeValues,
 * SolverFMGInitialize (MGB_USE_FMG=1), the raw-pointer smoother on caller-owned
 * arrays (small: staged per call; large: device-resident sessions), and host
 * writes into `grid` / the caller's arrays between device calls (the
 * page-protection coherence of mgb_coherence.h).
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

    if (getenv("MGB_USE_FMG")) { /* mg_dirichlet_analytic.c:984-989 */
#pragma omp parallel
        SolverFMGInitialize();
        printf("fmg_residual %.17g\n", SolverGetResidual());
    }
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

    /* edge / corner averaging on the device against the host routine on a copy */
    {
        double *copy = malloc(sizeof(double) * (size_t)N * N * N);
        memcpy(copy, grid, sizeof(double) * (size_t)N * N * N);
        updateEdgeValues(copy, N);
        SolverSmoothenEdgeValues();
        printf("edges_equal %d corner %.17g\n",
               memcmp(copy, grid, sizeof(double) * (size_t)N * N * N) == 0, grid[0]);
        free(copy);
    }

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

    /* the same flow on arrays large enough for a device-resident session (host pointer
     * -> device registry): many calls, a host write in between, then the result is read
     * straight from memory like test_rb_gs_3d.c:117-131 does; twice, the second time on
     * freshly allocated arrays (which may get the same addresses) */
    for (int round = 0; round < 2; round++) {
        const int P = 65;
        const size_t n3 = (size_t)P * P * P;
        const double hp = GRID_LENGTH / (P - 1) * (round ? 0.5 : 1.0);
        double *v = calloc(n3, sizeof(double)), *f = calloc(n3, sizeof(double));
        setupBoundaryConditions(v, P, hp);
        for (size_t t = 0; t < n3; t += 7)
            f[t] = 1e-3 * (double)(t % 13) * (round + 1);
        double last = 0.;
#pragma omp parallel
        {
            for (int it = 0; it < 6; it++) {
                preSmoother(v, f, P, hp, 1);
                postSmoother(v, f, P, hp, 1);
                const double pr = calculateResidual(v, f, P, hp, NULL);
                if (pr != 0.) { /* exactly one thread of the team gets the norm */
#pragma omp critical
                    last = pr;
                }
#pragma omp barrier
#pragma omp single
                {
                    if (it == 2) /* host write between two device calls (interior page) */
                        v[((size_t)(P / 2) * P + P / 2) * P + P / 2] += 0.25;
                    if (it == 3) /* ... and one into the unprotectable first page: the face
                                    point (0,1,1), a neighbour of the interior point (1,1,1) */
                        v[P + 1] -= 0.125;
                }
            }
        }
        double ss = 0.;
        for (size_t t = 0; t < n3; t++)
            ss += v[t] * v[t];
        printf("session%d residual %.17g sumsq %.17g probe %.17g first %.17g\n", round, last, ss,
               v[((size_t)3 * P + 5) * P + 7], v[P + 1]);
        if (getenv("MGB_DUMP_DIR")) {
            char path[512];
            snprintf(path, sizeof path, "%s/session%d.bin", getenv("MGB_DUMP_DIR"), round);
            FILE *fp = fopen(path, "wb");
            fwrite(v, sizeof(double), n3, fp);
            fclose(fp);
        }
        free(v);
        free(f);
    }

    if (getenv("MGB_WRITE_VTK"))
        writeOutputData(getenv("MGB_WRITE_VTK"), grid, h, N);
    SolverFinalize();
    free(threadNorm);
    return 0;
}
