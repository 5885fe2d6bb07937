/*
 * timing_info.h -- drop-in for the reference's timing_info.h (same struct,
 * same four functions, same print format; reference timing_info.h:6-80).
 * In this build the numbers are CUDA-event times per level and stage, fed by
 * mgb_timing() from mg_3d.h.
 */
#ifndef TIMING_INFO_H
#define TIMING_INFO_H

#include <assert.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

/* layout is part of the API: drivers touch the fields (test_rb_gs_3d.c:22-24) */
typedef struct __time_t
{
    int numStages;
    char **stageNames;
    int *numCalls;
    double *timeTaken;
} TimingInfo;

/* reference timing_info.h:14-32 */
void allocTimingInfo(TimingInfo **tInfo, char **stageNames, const int numStages)
{
    TimingInfo *t = (TimingInfo *)malloc(sizeof *t);
    assert(t);
    t->numStages = numStages;
    t->stageNames = (char **)malloc(sizeof(char *) * numStages);
    t->numCalls = (int *)calloc(numStages, sizeof(int));
    t->timeTaken = (double *)calloc(numStages, sizeof(double));
    assert(t->stageNames && t->numCalls && t->timeTaken);
    for (int s = 0; s < numStages; s++) {
        size_t len = strlen(stageNames[s]) + 1;
        t->stageNames[s] = (char *)malloc(len);
        memcpy(t->stageNames[s], stageNames[s], len);
    }
    *tInfo = t;
}

/* reference timing_info.h:34-38 */
void resetTimingInfo(TimingInfo *tInfo)
{
    for (int s = 0; s < tInfo->numStages; s++) {
        tInfo->numCalls[s] = 0;
        tInfo->timeTaken[s] = 0.;
    }
}

/* reference timing_info.h:40-47: header "%20s %20s %20s", rows "%20.20s %20d %20lf" */
void printTimingInfo(TimingInfo *tInfo)
{
    printf("%20s %20s %20s\n", "", "numCalls", "timeTaken");
    for (int s = 0; s < tInfo->numStages; s++)
        printf("%20.20s %20d %20lf\n", tInfo->stageNames[s], tInfo->numCalls[s],
               tInfo->timeTaken[s]);
}

/* reference timing_info.h:69-80 */
void deAllocTimingInfo(TimingInfo **tInfo)
{
    TimingInfo *t = *tInfo;
    for (int s = 0; s < t->numStages; s++)
        free(t->stageNames[s]);
    free(t->stageNames);
    free(t->numCalls);
    free(t->timeTaken);
    free(t);
}

#endif
