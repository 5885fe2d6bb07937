/*
 * postprocess.h -- drop-in for the reference's ASCII legacy-VTK writer
 * (reference postprocess.h:5-47).  Output is byte-identical.  At 513^3 the
 * reference's fprintf-per-line loop costs minutes (8 GB of text), so the text is
 * produced on the GPU (libmgb: mgb_vtk_open/_next/_close, csrc/vtk.cu -- exact
 * decimal conversion on the device, the C library's snprintf for the few values
 * whose rounding the device cannot decide) and this routine only fwrite()s the
 * chunks.  MGB_VTK_GPU=0 (or a build without mgb.h) selects the host formatter: lines
 * formatted in parallel by the OpenMP team into per-chunk buffers; a failing GPU path
 * aborts with the library's message, it never falls back by itself.
 */
#ifndef POSTPROCESS_H
#define POSTPROCESS_H

#include <stdio.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#if defined(__has_include)
#if __has_include("mgb.h")
#include "mgb.h"
#define MG_VTK_HAVE_GPU 1
#endif
#endif

/* formats lines [lo,hi) of one section into buf, returns bytes written */
static size_t mgVtkFormat(char *buf, int section, long long lo, long long hi,
                          const double *grid, double h, int N)
{
    size_t n = 0;
    if (section == 0) { /* "%10.8e %10.8e %10.8e\n" : x y z with k fastest */
        for (long long p = lo; p < hi; p++) {
            const int k = (int)(p % N), j = (int)((p / N) % N), i = (int)(p / ((long long)N * N));
            n += (size_t)sprintf(buf + n, "%10.8e %10.8e %10.8e\n", h * i, h * j, h * k);
        }
    } else { /* "%10.8e\n" */
        for (long long p = lo; p < hi; p++)
            n += (size_t)sprintf(buf + n, "%10.8e\n", grid[p]);
    }
    return n;
}

#ifdef MG_VTK_HAVE_GPU
/* the file through the device formatter; returns 0 when it was written completely */
static int mgVtkWriteGpu(FILE *f, const double *grid, double h, int N, long long *hostChunks)
{
    mgb_vtk *w = NULL;
    const double *src = grid;
#ifdef MGB_COHERENCE_H
    /* an array under page protection (grid of SolverGetDetails): bring it up to date and
     * hand libmgb the always-accessible mapping */
    MG_LOCK();
    src = mgReadable(grid);
    MG_UNLOCK();
#endif
    const char *e = getenv("MGB_DEVICE");
    if (mgb_vtk_open(&w, src, N, N, N, h, e ? atoi(e) : 0) != 0)
        return 1;
    int rc = 0;
    for (;;) {
        const char *p = NULL;
        long long n = 0;
        if (mgb_vtk_next(w, &p, &n) != 0) {
            rc = 1;
            break;
        }
        if (n == 0)
            break;
        if (fwrite(p, 1, (size_t)n, f) != (size_t)n) {
            rc = 1;
            break;
        }
    }
    mgb_vtk_host_chunks(w, hostChunks);
    mgb_vtk_close(w);
    return rc;
}
#endif

void writeOutputData(const char *fileName, const double *grid, const double h, const int N)
{
#ifdef _OPENMP
    const double mgVtkT0 = omp_get_wtime();
#endif
    FILE *f = fopen(fileName, "w");
    if (!f) {
        perror(fileName);
        return;
    }
    const long long total = (long long)N * N * N;
#ifdef MG_VTK_HAVE_GPU
    {
        const char *e = getenv("MGB_VTK_GPU");
        if (!e || atoi(e) != 0) {
            long long hostChunks = 0;
            if (mgVtkWriteGpu(f, grid, h, N, &hostChunks) == 0) {
                fclose(f);
#ifdef _OPENMP
                if (getenv("MGB_VTK_TIMING"))
                    fprintf(stderr, "mgb: writeOutputData %s: %d^3 points, formatted on the GPU "
                            "(%lld chunks by the host's snprintf), %.3f s\n", fileName, N,
                            hostChunks, omp_get_wtime() - mgVtkT0);
#endif
                return;
            }
            /* no silent fall-back: the host formatter runs only when asked for */
            fprintf(stderr, "mgb: writeOutputData %s failed: %s (MGB_VTK_GPU=0 selects the host "
                    "formatter)\n", fileName, mgb_last_error());
            abort();
        }
    }
#endif
    /* header, reference postprocess.h:13-19 */
    fprintf(f,
            "# vtk DataFile Version 2.0\n"
            "Potential data\n"
            "ASCII\n"
            "DATASET STRUCTURED_GRID\n"
            "DIMENSIONS %d %d %d\n"
            "POINTS %d float\n",
            N, N, N, (int)total);
    const long long chunk = 1 << 16; /* lines per task */
#ifdef _OPENMP
    const int nt = omp_get_max_threads();
#else
    const int nt = 1;
#endif
    const size_t per_line[2] = {3 * 24 + 4, 24 + 2};
    for (int section = 0; section < 2; section++) {
        if (section == 1) /* reference postprocess.h:37-41 */
            fprintf(f,
                    "\n"
                    "POINT_DATA %d\n"
                    "SCALARS data float 1\n"
                    "LOOKUP_TABLE default\n",
                    (int)total);
        char **bufs = (char **)malloc(sizeof(char *) * nt);
        size_t *lens = (size_t *)malloc(sizeof(size_t) * nt);
        for (int t = 0; t < nt; t++)
            bufs[t] = (char *)malloc(per_line[section] * (size_t)chunk);
        for (long long base = 0; base < total; base += chunk * nt) {
#pragma omp parallel for schedule(static, 1)
            for (int t = 0; t < nt; t++) {
                long long lo = base + (long long)t * chunk, hi = lo + chunk;
                if (hi > total) hi = total;
                lens[t] = lo < hi ? mgVtkFormat(bufs[t], section, lo, hi, grid, h, N) : 0;
            }
            for (int t = 0; t < nt; t++)
                if (lens[t])
                    fwrite(bufs[t], 1, lens[t], f);
        }
        for (int t = 0; t < nt; t++)
            free(bufs[t]);
        free(bufs);
        free(lens);
    }
    fclose(f);
#ifdef _OPENMP
    if (getenv("MGB_VTK_TIMING")) /* stderr: stdout stays what the reference prints */
        fprintf(stderr, "mgb: writeOutputData %s: %d^3 points, %d threads, %.3f s\n", fileName, N,
                nt, omp_get_wtime() - mgVtkT0);
#endif
}

#endif
