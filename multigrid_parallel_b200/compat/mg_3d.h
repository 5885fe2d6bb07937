/*
 * mg_3d.h -- drop-in replacement for knram06/multigrid_parallel's mg_3d.h.
 *
 * Same globals, same function names and signatures, same calling convention
 * (include once after `#define GRID_LENGTH`, call every compute routine from
 * ALL threads of the enclosing `#pragma omp parallel`, square-sum the returned
 * per-thread norm partials), so the reference's own drivers -- test_mg_3d.c
 * first of all -- compile against it unchanged.  Underneath, every operator of
 * the V-cycle runs as sm_100a CUDA through the C ABI of include/mgb.h; nothing
 * on the hot path is computed on the host and there is no CPU fallback (a
 * failing libmgb call aborts, like the reference's asserts).
 *
 * Host-pointer semantics.  The reference hands out raw pointers into solver
 * storage (SolverGetDetails, reference mg_3d.h:275-293) and the drivers read
 * and write them with no API call in between (test_mg_3d.c:29, 81-95).  The
 * finest-level `grid`/`rhs` therefore stay ordinary host arrays here, kept
 * coherent with the device copy by page protection: after the GPU has changed
 * the solution the host pages are PROT_NONE and the first touch downloads
 * them; after a download they are read-only and the first host write marks
 * them dirty, so the next GPU call uploads them.  MGB_LAZY_SYNC=0 turns this
 * off and falls back to explicit points (upload before the first GPU call
 * after SolverGetDetails/setupBoundaryConditions, download in
 * SolverPrintTimingInfo/SolverGetResidual/SolverSmoothenEdgeValues/
 * SolverFinalize).
 *
 * Environment: MGB_DEVICE (default 0), MGB_PROFILE (default 1: per-stage CUDA
 * event timing into tInfo, eager launches; 0: one CUDA graph per V-cycle),
 * MGB_LAZY_SYNC (default 1).
 */
#ifndef MG_3D_H
#define MG_3D_H

#include <assert.h>
#include <limits.h>
#include <math.h>
#include <signal.h>
#include <stdbool.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <unistd.h>

#include <omp.h>

#include "mgb.h"

#include "gauss_elim.h"
#include "timing_info.h"

/* ---- the reference's globals (reference mg_3d.h:19-28) ---- */
TimingInfo **tInfo = NULL;
int coarseGridNum;
int finestOneSideNum;
int numLevels;
int gsIterNum;
double **u, **d, **r; /* per-level host arrays, natural layout */
double *A;            /* coarsest-level matrix (LU-factorised)  */
double spacing;

/* ---- private state of the GPU backend ---- */
static mgb_solver *mgGpu = NULL;
static int mgLazySync = 1;

enum { MG_CLEAN = 0, MG_HOST_NEWER = 1, MG_DEV_NEWER = 2 };
typedef struct {
    double *host;
    size_t bytes;  /* page-rounded */
    int which;     /* MGB_U / MGB_D */
    volatile int state;
    volatile int prot; /* current PROT_* of the host pages */
} MgMirror;
static MgMirror mgMirror[2];
static volatile int mgMirrorLock = 0;
static struct sigaction mgOldSegv;
static int mgSegvInstalled = 0;

#define MGB_OK(call) mgb_compat_check((call), #call)

static size_t mgPageRound(size_t n)
{
    const size_t pg = (size_t)sysconf(_SC_PAGESIZE);
    return (n + pg - 1) / pg * pg;
}

static void mgProtect(MgMirror *m, int prot)
{
    if (!mgLazySync || !m->host || m->prot == prot)
        return;
    mprotect(m->host, m->bytes, prot);
    m->prot = prot;
}

/* bring the host copy up to date (device -> host) */
static void mgPull(MgMirror *m)
{
    if (m->state != MG_DEV_NEWER)
        return;
    mgProtect(m, PROT_READ | PROT_WRITE);
    MGB_OK(mgb_download(mgGpu, numLevels - 1, m->which, m->host));
    /* lazy mode: a later host write faults once and marks the pages dirty;
     * explicit mode cannot see host writes, so assume one happens */
    m->state = mgLazySync ? MG_CLEAN : MG_HOST_NEWER;
    mgProtect(m, PROT_READ);
}

/* bring the device copy up to date (host -> device) */
static void mgPush(MgMirror *m)
{
    if (m->state != MG_HOST_NEWER)
        return;
    MGB_OK(mgb_upload(mgGpu, numLevels - 1, m->which, m->host));
    m->state = MG_CLEAN;
    mgProtect(m, PROT_READ);
}

static void mgSegvHandler(int sig, siginfo_t *si, void *ctx)
{
    const uintptr_t a = (uintptr_t)si->si_addr;
    for (int t = 0; t < 2; t++) {
        MgMirror *m = &mgMirror[t];
        if (!m->host || a < (uintptr_t)m->host || a >= (uintptr_t)m->host + m->bytes)
            continue;
        while (__atomic_test_and_set(&mgMirrorLock, __ATOMIC_ACQUIRE))
            ;
        if (m->prot == PROT_NONE) {
            mgPull(m); /* first touch after GPU work */
        } else if (m->prot == PROT_READ) {
            m->state = MG_HOST_NEWER; /* first host write since the last sync */
            mgProtect(m, PROT_READ | PROT_WRITE);
        }
        __atomic_clear(&mgMirrorLock, __ATOMIC_RELEASE);
        return; /* retry the faulting access */
    }
    /* not ours: hand over to whoever was installed before */
    sigaction(SIGSEGV, &mgOldSegv, NULL);
    (void)sig;
    (void)ctx;
}

static void mgInstallSegv(void)
{
    if (!mgLazySync || mgSegvInstalled)
        return;
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = mgSegvHandler;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    sigaction(SIGSEGV, &sa, &mgOldSegv);
    mgSegvInstalled = 1;
}

static MgMirror *mgMirrorOf(const double *p)
{
    for (int t = 0; t < 2; t++)
        if (mgMirror[t].host && mgMirror[t].host == p)
            return &mgMirror[t];
    return NULL;
}

/* explicit-mode bookkeeping: the host may have written p */
static void mgTouchedByHost(const double *p)
{
    MgMirror *m = mgMirrorOf(p);
    if (m && !mgLazySync)
        m->state = MG_HOST_NEWER;
}

static void mgSyncToDevice(void)
{
    mgPush(&mgMirror[0]);
    mgPush(&mgMirror[1]);
}

static void mgDeviceChangedSolution(void)
{
    MgMirror *m = &mgMirror[0];
    m->state = MG_DEV_NEWER;
    mgProtect(m, PROT_NONE);
}

static void mgPullTimings(void)
{
    for (int l = 0; l < numLevels; l++)
        for (int s = 0; s < tInfo[l]->numStages && s < MGB_NUM_STAGES; s++)
            MGB_OK(mgb_timing(mgGpu, l, s, &tInfo[l]->numCalls[s], &tInfo[l]->timeTaken[s]));
}

/* ---- level storage (reference mg_3d.h:30-48, 295-302) ---- */
static size_t mgLevelBytes(int level, int coarseN)
{
    const size_t n = (size_t)(coarseN - 1) * ((size_t)1 << level) + 1;
    return mgPageRound(n * n * n * sizeof(double));
}

void allocGridLevels(double ***arr, const int numLevels, const int N)
{
    *arr = (double **)malloc(sizeof(double *) * numLevels);
    assert(*arr);
    for (int l = 0; l < numLevels; l++) {
        /* zero-filled, page-aligned, committed lazily */
        void *p = mmap(NULL, mgLevelBytes(l, N), PROT_READ | PROT_WRITE,
                       MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
        assert(p != MAP_FAILED);
        (*arr)[l] = (double *)p;
    }
}

void deAllocGridLevels(double ***arr, const int numLevels)
{
    for (int l = 0; l < numLevels; l++)
        munmap((*arr)[l], mgLevelBytes(l, coarseGridNum));
    free(*arr);
}

/* ---- debugging printers (reference mg_3d.h:51-87) ---- */
void printGrid3D(const double *grid, const int oneSideN)
{
    const int N = oneSideN;
    for (int i = 0; i < N; i++) {
        printf("LEVEL %d\n", i);
        for (int k = N - 1; k >= 0; k--) {
            for (int j = 0; j < N; j++)
                printf("%10.5g ", grid[((size_t)i * N + j) * N + k]);
            printf("\n");
        }
        printf("\n");
    }
}

void printMatrix(const double *mat, const int oneSideDim)
{
    for (int i = 0; i < oneSideDim; i++) {
        for (int j = 0; j < oneSideDim; j++)
            printf("%10.5lf ", mat[(size_t)i * oneSideDim + j]);
        printf("\n");
    }
}

/* analytic Dirichlet data (reference mg_3d.h:89-90) */
double BCFunc(double x, double y, double z) { return x * x - 2 * y * y + z * z; }

bool isPowerOfTwo(int x) { return (x & (x - 1)) == 0; }

/* ---- set-up (reference mg_3d.h:107-144) ---- */
void SolverInitialize(int argc, char **argv)
{
    if (argc != 4) {
        printf("Usage: %s <coarse grid points on one side> <number of levels> <gauss seidel iterations>\n",
               argv[0]);
        exit(1);
    }
    coarseGridNum = atoi(argv[1]);
    numLevels = atoi(argv[2]);
    gsIterNum = atoi(argv[3]);
    assert(isPowerOfTwo(coarseGridNum - 1));
    finestOneSideNum = (coarseGridNum - 1) * (1 << (numLevels - 1)) + 1;

    const char *e;
    mgLazySync = (e = getenv("MGB_LAZY_SYNC")) ? atoi(e) != 0 : 1;
    const int device = (e = getenv("MGB_DEVICE")) ? atoi(e) : 0;
    const int profile = (e = getenv("MGB_PROFILE")) ? atoi(e) != 0 : 1;

    u = NULL; d = NULL; r = NULL;
    allocGridLevels(&u, numLevels, coarseGridNum);
    allocGridLevels(&d, numLevels, coarseGridNum);
    allocGridLevels(&r, numLevels, coarseGridNum);

    tInfo = (TimingInfo **)malloc(sizeof(TimingInfo *) * numLevels);
    char *stageNames[7] = {"Smoother1", "CalcResidual1", "Restrict Residual",
                           "Recurse, Direct Solve", "Prolongate&Correct", "Smoother2",
                           "CalcResidual2"};
    for (int l = 0; l < numLevels; l++)
        allocTimingInfo(&tInfo[l], stageNames, 7);

    spacing = GRID_LENGTH / (finestOneSideNum - 1);

    /* the GPU hierarchy: levels in HBM, coarse operator built + factorised there */
    MGB_OK(mgb_create(&mgGpu, coarseGridNum, coarseGridNum, coarseGridNum, numLevels,
                      gsIterNum, device));
    MGB_OK(mgb_set_option(mgGpu, MGB_OPT_PROFILE, profile));
    const int top = numLevels - 1;
    const size_t bytes = mgLevelBytes(top, coarseGridNum);
    mgMirror[0] = (MgMirror){u[top], bytes, MGB_U, MG_CLEAN, PROT_READ | PROT_WRITE};
    mgMirror[1] = (MgMirror){d[top], bytes, MGB_D, MG_CLEAN, PROT_READ | PROT_WRITE};
    mgInstallSegv();
    /* device arrays start zeroed like the calloc'd host arrays: in sync; from
     * now on the first host write to either array is noticed */
    mgProtect(&mgMirror[0], PROT_READ);
    mgProtect(&mgMirror[1], PROT_READ);
}

/* coarsest operator (reference mg_3d.h:147-273), assembled on the GPU */
void constructCoarseMatrixA(double *A, int N, const double h)
{
    assert((long long)N * N * N * N * N * N < INT_MAX);
    MGB_OK(mgb_host_coarse_matrix(A, N, N, N, h));
}

/* reference mg_3d.h:275-293 */
int SolverGetDetails(double **grid, double **rhs, double *h)
{
    (*grid) = u[numLevels - 1];
    (*rhs) = d[numLevels - 1];
    /* A = the factorised coarse operator; mgb_create already built it from the
     * coarse spacing spacing*2^(L-1) and LU-factorised it on the device */
    const size_t matDim = (size_t)coarseGridNum * coarseGridNum * coarseGridNum;
    A = (double *)calloc(matDim * matDim, sizeof(double));
    assert(A);
    MGB_OK(mgb_coarse_lu_download(mgGpu, A));
    if (!mgLazySync)
        mgMirror[0].state = mgMirror[1].state = MG_HOST_NEWER;
    *h = spacing;
    return finestOneSideNum;
}

/* Dirichlet faces, host side, O(N^2) set-up (reference mg_3d.h:1147-1239) */
void setupBoundaryConditions(double *v, int levelN, double spacing)
{
    const int N = levelN;
    const size_t NN = (size_t)N * N;
    for (int a = 0; a < N; a++)
        for (int b = 0; b < N; b++) {
            const double ah = a * spacing, bh = b * spacing, end = (N - 1) * spacing;
            v[NN * a + (size_t)N * 0 + b] = BCFunc(ah, 0 * spacing, bh);       /* j = 0   */
            v[NN * a + (size_t)N * (N - 1) + b] = BCFunc(ah, end, bh);         /* j = N-1 */
            v[NN * a + (size_t)N * b + 0] = BCFunc(ah, bh, 0 * spacing);       /* k = 0   */
            v[NN * a + (size_t)N * b + (N - 1)] = BCFunc(ah, bh, end);         /* k = N-1 */
            v[NN * 0 + (size_t)N * a + b] = BCFunc(0 * spacing, ah, bh);       /* i = 0   */
            v[NN * (N - 1) + (size_t)N * a + b] = BCFunc(end, ah, bh);         /* i = N-1 */
        }
    mgTouchedByHost(v);
}

void SolverSetupBoundaryConditions()
{
    setupBoundaryConditions(d[numLevels - 1], finestOneSideNum, spacing);
}

/* ---- raw-pointer operators: collective over the OpenMP team, run once ----
 * The arrays of the solver's finest level are operated on in place on the
 * device; any other pointers go through the stateless staging entry points. */
static int mgIsFinest(const double *v, const double *dd, int N)
{
    return mgGpu && N == finestOneSideNum && v == u[numLevels - 1] && dd == d[numLevels - 1];
}

static void mgSmooth(double *v, const double *dd, int N, double h, int iters, int firstRed)
{
#pragma omp single
    {
        if (mgIsFinest(v, dd, N) && h == spacing) {
            mgSyncToDevice();
            MGB_OK(mgb_smooth(mgGpu, numLevels - 1, iters, firstRed));
            MGB_OK(mgb_sync(mgGpu));
            mgDeviceChangedSolution();
        } else {
            MGB_OK(mgb_host_smooth(v, dd, N, N, N, h, iters, firstRed));
        }
    }
}

/* red then black (reference mg_3d.h:640-709) */
void preSmoother(double *__restrict__ v, const double *__restrict__ d, const int N,
                 const double h, const int smootherIter)
{
    mgSmooth(v, d, N, h, smootherIter, 1);
}

/* black then red (reference mg_3d.h:711-781) */
void postSmoother(double *__restrict__ v, const double *__restrict__ d, const int N,
                  const double h, const int smootherIter)
{
    mgSmooth(v, d, N, h, smootherIter, 0);
}

static double mgSingleResult;
static int mgSingleOwner;

/* reference mg_3d.h:794-842: exactly one thread of the team returns the norm,
 * the others 0, so sqrt(sum of squares of the partials) is unchanged */
double calculateResidual(const double *__restrict__ v, const double *__restrict__ d,
                         const int N, const double h, double *res)
{
#pragma omp single
    {
        double ss = 0.;
        if (mgIsFinest(v, d, N) && h == spacing && (res == NULL || res == r[numLevels - 1])) {
            mgSyncToDevice();
            MGB_OK(mgb_residual(mgGpu, numLevels - 1, res != NULL, &ss));
            if (res)
                MGB_OK(mgb_download(mgGpu, numLevels - 1, MGB_R, res));
        } else {
            MGB_OK(mgb_host_residual(v, d, N, N, N, h, res, &ss));
        }
        mgSingleResult = sqrt(ss);
        mgSingleOwner = omp_get_thread_num();
    }
    return omp_get_thread_num() == mgSingleOwner ? mgSingleResult : 0.;
}

double GetL2NormOfVector(const double *v, const int n)
{
    /* reference mg_3d.h:783-792: serial host sum, defines the stopping
     * threshold (test_mg_3d.c:26-31); kept on the host in the same order */
    double ret = 0.;
    for (int i = 0; i < n; i++)
        ret += v[i] * v[i];
    return sqrt(ret);
}

/* reference mg_3d.h:844-998 */
void restrictResidual(const double *__restrict__ rf, const int Nf, double *__restrict__ dc,
                      const int Nc)
{
#pragma omp single
    MGB_OK(mgb_host_restrict(rf, Nf, Nf, Nf, dc, Nc, Nc, Nc));
}

/* reference mg_3d.h:1000-1145 */
void prolongateAndCorrectError(const double *__restrict__ ec, const int Nc,
                               double *__restrict__ ef, const int Nf)
{
#pragma omp single
    MGB_OK(mgb_host_prolong_correct(ec, Nc, Nc, Nc, ef, Nf, Nf, Nf));
}

/* ---- boundary edge/corner averaging and lexicographic Gauss-Seidel ----
 * Not called by the V-cycle (reference call sites are commented out:
 * mg_3d.h:707,779,1281,1340); kept as plain host C so the API is complete. */
void updateEdgeValues(double *__restrict__ v, const int N)
{
    /* reference mg_3d.h:304-430: each of the 12 edges' inner points becomes the
     * mean of its two inward neighbours; then each corner the mean of its
     * three (added in k, j, i order) */
    const long long s[3] = {(long long)N * N, N, 1};
    const int end[2] = {0, N - 1};
    for (int fa = 0; fa < 3; fa++) { /* the axis running along the edge */
        const int b = (fa + 1) % 3, c = (fa + 2) % 3;
        for (int eb = 0; eb < 2; eb++)
            for (int ec = 0; ec < 2; ec++) {
                const long long inb = eb ? -s[b] : s[b], inc = ec ? -s[c] : s[c];
                for (int t = 1; t < N - 1; t++) {
                    const long long p = t * s[fa] + end[eb] * s[b] + end[ec] * s[c];
                    v[p] = 0.5 * (v[p + inb] + v[p + inc]);
                }
            }
    }
    for (int ei = 0; ei < 2; ei++)
        for (int ej = 0; ej < 2; ej++)
            for (int ek = 0; ek < 2; ek++) {
                const long long p = end[ei] * s[0] + end[ej] * s[1] + end[ek] * s[2];
                const long long ini = ei ? -s[0] : s[0], inj = ej ? -s[1] : s[1],
                                ink = ek ? -s[2] : s[2];
                v[p] = (1. / 3) * (v[p + ink] + v[p + inj] + v[p + ini]);
            }
}

void GaussSeidelSmoother(double *__restrict__ v, const double *__restrict__ d, const int N,
                         const double h, const int smootherIter)
{
    /* reference mg_3d.h:546-637: lexicographic sweep (inherently serial) */
    const double hSq = h * h, sixth = 1. / 6;
    const long long NN = (long long)N * N;
    for (int s = 0; s < smootherIter; s++)
        for (int i = 1; i < N - 1; i++)
            for (int j = 1; j < N - 1; j++)
                for (int k = 1; k < N - 1; k++) {
                    const long long p = NN * i + (long long)N * j + k;
                    v[p] = sixth * (v[p - NN] + v[p + NN] + v[p - N] + v[p + N] + v[p - 1] +
                                    v[p + 1] - hSq * d[p]);
                }
    updateEdgeValues(v, N);
    mgTouchedByHost(v);
}

/* ---- the V-cycle (reference mg_3d.h:1242-1362) ---- */
static double mgDeviceCycle(void)
{
    double ss = 0.;
    mgSyncToDevice();
    MGB_OK(mgb_vcycle(mgGpu, &ss));
    mgDeviceChangedSolution();
    mgPullTimings();
    return sqrt(ss);
}

double vcycle(double **uu, double **ff, double **rr, double h, int q, const int nLevels,
              const int smootherIter, int N, double *LU)
{
    /* the solver's own hierarchy from the top: one device-resident cycle */
    if (uu == u && ff == d && q == numLevels - 1 && nLevels == numLevels &&
        smootherIter == gsIterNum && h == spacing) {
#pragma omp single
        {
            mgSingleResult = mgDeviceCycle();
            mgSingleOwner = omp_get_thread_num();
        }
        return omp_get_thread_num() == mgSingleOwner ? mgSingleResult : 0.;
    }
    /* any other arrays / entry level: the reference's recursion, operator by
     * operator on the GPU through the staging entry points */
    double *v = uu[q], *f = ff[q], *res = rr[q];
#pragma omp single
    {
        if (q < nLevels - 1)
            memset(v, 0, sizeof(double) * (size_t)N * N * N);
    }
    if (q == 0) {
#pragma omp single
        solveWithLU(LU, N * N * N, f, v);
        return 0.;
    }
    preSmoother(v, f, N, h, smootherIter);
    calculateResidual(v, f, N, h, res);
    const int Nc = (N + 1) / 2;
    restrictResidual(res, N, ff[q - 1], Nc);
    vcycle(uu, ff, rr, 2 * h, q - 1, nLevels, smootherIter, Nc, LU);
    prolongateAndCorrectError(uu[q - 1], Nc, v, N);
    postSmoother(v, f, N, h, smootherIter);
    return calculateResidual(v, f, N, h, NULL);
}

/* ---- Solver* wrappers (reference mg_3d.h:1412-1467) ---- */
double SolverLinSolve()
{
    return vcycle(u, d, r, spacing, numLevels - 1, numLevels, gsIterNum, finestOneSideNum, A);
}

void SolverSmoothenEdgeValues()
{
    mgPull(&mgMirror[0]);
    updateEdgeValues(u[numLevels - 1], finestOneSideNum);
    mgTouchedByHost(u[numLevels - 1]);
}

double SolverGetResidual()
{
    return calculateResidual(u[numLevels - 1], d[numLevels - 1], finestOneSideNum, spacing, NULL);
}

double SolverGetInitialResidual()
{
    return GetL2NormOfVector(d[numLevels - 1],
                             finestOneSideNum * finestOneSideNum * finestOneSideNum);
}

void SolverResetTimingInfo()
{
    MGB_OK(mgb_timing_reset(mgGpu));
    for (int l = 0; l < numLevels; l++)
        resetTimingInfo(tInfo[l]);
}

void SolverPrintTimingInfo()
{
    if (!mgLazySync)
        mgPull(&mgMirror[0]); /* explicit mode: the driver reads grid next */
    for (int l = 0; l < numLevels; l++) {
        printf("LEVEL %d\n", l);
        printTimingInfo(tInfo[l]);
    }
}

void SolverFinalize()
{
    mgProtect(&mgMirror[0], PROT_READ | PROT_WRITE);
    mgProtect(&mgMirror[1], PROT_READ | PROT_WRITE);
    mgMirror[0].host = mgMirror[1].host = NULL;
    for (int l = 0; l < numLevels; l++)
        deAllocTimingInfo(&tInfo[l]);
    free(tInfo);
    free(A);
    deAllocGridLevels(&u, numLevels);
    deAllocGridLevels(&d, numLevels);
    deAllocGridLevels(&r, numLevels);
    MGB_OK(mgb_destroy(mgGpu));
    mgGpu = NULL;
}

#endif /* MG_3D_H */
