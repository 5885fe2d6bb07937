/*
 * postprocess.h -- drop-in for the reference's ASCII legacy-VTK writer
 * (reference postprocess.h:5-47).  Output is byte-identical; the lines are
 * formatted in parallel into per-chunk buffers and written in order, because
 * at 513^3 the reference's fprintf-per-line loop costs minutes.
 */
#ifndef POSTPROCESS_H
#define POSTPROCESS_H

#include <stdio.h>
#include <stdlib.h>

#ifdef _OPENMP
#include <omp.h>
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
