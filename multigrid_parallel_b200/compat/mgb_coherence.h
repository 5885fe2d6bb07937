/*
 * mgb_coherence.h -- keeps host arrays the CALLER holds raw pointers to coherent
 * with their device copies inside libmgb.  Part of the drop-in for
 * knram06/multigrid_parallel's mg_3d.h; included by compat/mg_3d.h only.
 *
 * Why.  The reference hands out raw pointers into solver storage
 * (SolverGetDetails, reference mg_3d.h:275-293) and its drivers read and write
 * them with no API call in between (test_mg_3d.c:29, 81-95); the raw-pointer
 * routines (preSmoother / postSmoother / calculateResidual, test_rb_gs_3d.c:
 * 56-101) are called over and over on the caller's own arrays and the result is
 * then read straight from memory (test_rb_gs_3d.c:117-131).  The device copy is
 * what the kernels work on, so somebody has to notice when the host side is
 * read or written.
 *
 * How (MGB_LAZY_SYNC=1, the default).  The page-aligned part of such an array
 * is kept under page protection: PROT_NONE while the device copy is newer,
 * PROT_READ while both agree, read-write once the host has written.  The
 * SIGSEGV handler does NO CUDA work and takes no lock:
 *   - fault on a PROT_NONE array: post a request to the SERVICE THREAD (sem_post,
 *     async-signal-safe) and wait until the protection has changed, then return
 *     and let the access run again.  The service thread -- ordinary thread
 *     context -- downloads the array and only THEN opens the pages, so no thread
 *     ever sees a half-written array: the solver's own arrays are mapped twice
 *     (memfd), the download goes through the second, always-writable mapping
 *     while the caller's view stays PROT_NONE;
 *   - write fault on a PROT_READ array: mark it host-newer, make it writable;
 *     a READ can only fault under PROT_NONE, so a thread that waited for a pull
 *     is never mistaken for a writer (x86-64: the page-fault error code says
 *     which it was);
 *   - any other address: the handler that was installed before ours is called
 *     (SA_SIGINFO or plain), or the default action is restored.
 * Partial pages at either end of a caller-owned array cannot be protected (they
 * may hold other data: a calloc'd array starts 16 bytes into its first page); these
 * two fragments (< one page each) are copied eagerly: from the device after every
 * kernel that wrote the array, to the device before every kernel if they changed.
 * Caller-owned arrays have no second mapping: their pages are opened for the
 * duration of the copy (the reference's drivers read them from serial code).
 *
 * MGB_LAZY_SYNC=0: no page protection, no handler; arrays are uploaded before
 * the first device call after the host may have written them and downloaded at
 * the documented points (see mg_3d.h).
 *
 * All libmgb calls of the drop-in are serialised by mgGpuLock (the service
 * thread is the one caller that is not inside the driver's `omp single`).
 */
#ifndef MGB_COHERENCE_H
#define MGB_COHERENCE_H

#include <errno.h>
#include <pthread.h>
#include <sched.h>
#include <semaphore.h>
#include <signal.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/syscall.h>
#include <ucontext.h>
#include <unistd.h>

#include "mgb.h"

enum { MG_CLEAN = 0, MG_HOST_NEWER = 1, MG_DEV_NEWER = 2 };
enum { MG_MAX_ARRAYS = 16 };

typedef struct MgArr {
    double *host;       /* the caller's view */
    size_t bytes;       /* of the array */
    char *alias;        /* second, always-writable mapping of the same pages, or NULL */
    char *plo;          /* protected part of the caller's view: [plo, plo+plen) */
    size_t plen;
    mgb_solver *gpu;    /* device home */
    int level, which;
    volatile int state; /* MG_* */
    volatile int prot;  /* current PROT_* of [plo, plo+plen) */
    volatile int want_pull;
    int lazy;           /* page protection in use */
    int pinned;
    int live;
    /* caller-owned arrays: the unprotectable fragments in front of / behind the pages */
    size_t head_n, tail_first, tail_n; /* in doubles: [0, head_n) and [tail_first, +tail_n) */
    double *frag;                      /* what the device has for them (head, then tail) */
} MgArr;

static MgArr mgArrs[MG_MAX_ARRAYS];
static int mgLazySync = 1;
static pthread_mutex_t mgGpuLock = PTHREAD_MUTEX_INITIALIZER;
static sem_t mgSvcReq;
static pthread_t mgSvcThread;
static volatile int mgSvcRunning = 0;
static struct sigaction mgOldSegv;
static int mgSegvInstalled = 0;
static int mgProbePipe[2] = {-1, -1};

#define MG_LOCK() pthread_mutex_lock(&mgGpuLock)
#define MG_UNLOCK() pthread_mutex_unlock(&mgGpuLock)

#ifndef MGB_COMPAT_CHECK_DEFINED
#define MGB_COMPAT_CHECK_DEFINED
static void mgb_compat_check(int rc, const char *what)
{
    if (rc) {
        fprintf(stderr, "mg_3d.h (libmgb): %s failed: %s\n", what, mgb_last_error());
        abort(); /* the reference asserts; there is no CPU fallback to continue on */
    }
}
#endif
#define MGB_OK(call) mgb_compat_check((call), #call)

static size_t mgPageSize(void) { return (size_t)sysconf(_SC_PAGESIZE); }
static size_t mgPageRound(size_t n)
{
    const size_t pg = mgPageSize();
    return (n + pg - 1) / pg * pg;
}

static void mgSetProt(MgArr *a, int prot)
{
    if (!a->lazy || !a->plen || a->prot == prot)
        return;
    mprotect(a->plo, a->plen, prot);
    a->prot = prot;
}

/* the pointer libmgb copies from / to: never a protected view */
static double *mgXfer(MgArr *a) { return a->alias ? (double *)a->alias : a->host; }

static void mgFragRemember(MgArr *a)
{
    if (!a->frag)
        return;
    memcpy(a->frag, a->host, a->head_n * sizeof(double));
    memcpy(a->frag + a->head_n, a->host + a->tail_first, a->tail_n * sizeof(double));
}

/* device -> host (normal thread context, mgGpuLock held) */
static void mgArrPull(MgArr *a)
{
    if (a->state != MG_DEV_NEWER)
        return;
    /* caller-owned array (no second mapping): open the pages for the copy, but keep
     * `prot` saying PROT_NONE so that threads already waiting in the handler stay there
     * until the data is complete */
    if (!a->alias && a->lazy && a->plen && a->prot == PROT_NONE)
        mprotect(a->plo, a->plen, PROT_READ | PROT_WRITE);
    MGB_OK(mgb_download(a->gpu, a->level, a->which, mgXfer(a)));
    /* lazy: a later host write faults once and marks the array dirty; explicit mode
     * cannot see host writes, so it assumes one */
    a->state = a->lazy ? MG_CLEAN : MG_HOST_NEWER;
    mgFragRemember(a);
    if (a->lazy && a->plen) {
        if (a->prot == PROT_NONE) {
            mprotect(a->plo, a->plen, PROT_READ);
            __atomic_store_n(&a->prot, PROT_READ, __ATOMIC_RELEASE);
        } else {
            mgSetProt(a, PROT_READ);
        }
    }
}

/* host -> device (normal thread context, mgGpuLock held) */
static void mgArrPush(MgArr *a)
{
    if (a->state != MG_HOST_NEWER)
        return;
    /* stop further silent writes first, then copy: a write that lands during the copy
     * faults and marks the array dirty again */
    if (a->lazy)
        mgSetProt(a, PROT_READ);
    a->state = MG_CLEAN;
    MGB_OK(mgb_upload(a->gpu, a->level, a->which, mgXfer(a)));
    mgFragRemember(a);
}

/* before a kernel reads the array (mgGpuLock held): whole array if the host wrote it,
 * otherwise just the fragments nobody can watch, if they changed */
static void mgArrToDevice(MgArr *a)
{
    if (a->state == MG_HOST_NEWER) {
        mgArrPush(a);
        return;
    }
    if (!a->frag)
        return;
    if (a->head_n && memcmp(a->host, a->frag, a->head_n * sizeof(double)) != 0) {
        MGB_OK(mgb_upload_range(a->gpu, a->level, a->which, 0, (long long)a->head_n, a->host));
        memcpy(a->frag, a->host, a->head_n * sizeof(double));
    }
    if (a->tail_n && memcmp(a->host + a->tail_first, a->frag + a->head_n,
                            a->tail_n * sizeof(double)) != 0) {
        MGB_OK(mgb_upload_range(a->gpu, a->level, a->which, (long long)a->tail_first,
                                (long long)a->tail_n, a->host + a->tail_first));
        memcpy(a->frag + a->head_n, a->host + a->tail_first, a->tail_n * sizeof(double));
    }
}

/* the device copy has just been changed by a kernel (mgGpuLock held) */
static void mgArrDeviceWrote(MgArr *a)
{
    a->state = MG_DEV_NEWER;
    if (a->lazy && a->plen)
        mgSetProt(a, PROT_NONE);
    if (a->frag) { /* the fragments cannot fault: fetch them now */
        if (a->head_n)
            MGB_OK(mgb_download_range(a->gpu, a->level, a->which, 0, (long long)a->head_n,
                                      a->host));
        if (a->tail_n)
            MGB_OK(mgb_download_range(a->gpu, a->level, a->which, (long long)a->tail_first,
                                      (long long)a->tail_n, a->host + a->tail_first));
        mgFragRemember(a);
    }
}

static void *mgServiceMain(void *arg)
{
    (void)arg;
    for (;;) {
        while (sem_wait(&mgSvcReq) != 0 && errno == EINTR)
            ;
        for (int t = 0; t < MG_MAX_ARRAYS; t++) {
            MgArr *a = &mgArrs[t];
            if (!a->live || !a->want_pull)
                continue;
            MG_LOCK();
            a->want_pull = 0;
            if (a->live)
                mgArrPull(a);
            MG_UNLOCK();
        }
    }
    return NULL;
}

static int mgFaultIsWrite(void *ctx)
{
#if defined(__x86_64__)
    /* page-fault error code, bit 1 = write access (gregs[19] = REG_ERR) */
    return (int)((((ucontext_t *)ctx)->uc_mcontext.gregs[19] >> 1) & 1);
#else
    (void)ctx;
    return -1; /* unknown */
#endif
}

static void mgSegvHandler(int sig, siginfo_t *si, void *ctx)
{
    const int saved_errno = errno;
    const uintptr_t addr = (uintptr_t)si->si_addr;
    for (int t = 0; t < MG_MAX_ARRAYS; t++) {
        MgArr *a = &mgArrs[t];
        if (!a->live || !a->lazy || !a->plen || addr < (uintptr_t)a->plo ||
            addr >= (uintptr_t)a->plo + a->plen)
            continue;
        const int prot = __atomic_load_n(&a->prot, __ATOMIC_ACQUIRE);
        if (prot == PROT_NONE) {
            /* first touch after device work: the service thread pulls, we wait */
            __atomic_store_n(&a->want_pull, 1, __ATOMIC_RELEASE);
            sem_post(&mgSvcReq);
            while (__atomic_load_n(&a->prot, __ATOMIC_ACQUIRE) == PROT_NONE)
                sched_yield();
        } else if (prot == PROT_READ) {
            const int w = mgFaultIsWrite(ctx);
            if (w != 0) { /* a write (or unknown: assume one, costs an upload at most) */
                a->state = MG_HOST_NEWER;
                mprotect(a->plo, a->plen, PROT_READ | PROT_WRITE);
                __atomic_store_n(&a->prot, PROT_READ | PROT_WRITE, __ATOMIC_RELEASE);
            }
            /* a read that faulted under PROT_NONE and got here late: just retry */
        }
        errno = saved_errno;
        return; /* run the faulting access again */
    }
    /* not ours: whoever was installed before us */
    errno = saved_errno;
    if (mgOldSegv.sa_flags & SA_SIGINFO) {
        if (mgOldSegv.sa_sigaction) {
            mgOldSegv.sa_sigaction(sig, si, ctx);
            return;
        }
    } else if (mgOldSegv.sa_handler == SIG_IGN) {
        return;
    } else if (mgOldSegv.sa_handler != SIG_DFL) {
        mgOldSegv.sa_handler(sig);
        return;
    }
    signal(SIGSEGV, SIG_DFL); /* default action on the re-executed access */
}

static void mgCoherenceInit(void)
{
    if (!mgLazySync || mgSegvInstalled)
        return;
    sem_init(&mgSvcReq, 0, 0);
    if (pthread_create(&mgSvcThread, NULL, mgServiceMain, NULL) != 0) {
        mgLazySync = 0; /* no service thread: explicit synchronisation */
        return;
    }
    pthread_detach(mgSvcThread);
    mgSvcRunning = 1;
    struct sigaction sa;
    memset(&sa, 0, sizeof sa);
    sa.sa_sigaction = mgSegvHandler;
    sa.sa_flags = SA_SIGINFO | SA_NODEFER;
    sigemptyset(&sa.sa_mask);
    sigaction(SIGSEGV, &sa, &mgOldSegv);
    mgSegvInstalled = 1;
    if (pipe(mgProbePipe) != 0)
        mgProbePipe[0] = mgProbePipe[1] = -1;
}

static MgArr *mgArrNew(void)
{
    for (int t = 0; t < MG_MAX_ARRAYS; t++)
        if (!mgArrs[t].live) {
            memset(&mgArrs[t], 0, sizeof(MgArr));
            return &mgArrs[t];
        }
    return NULL;
}

static MgArr *mgArrOf(const double *p)
{
    for (int t = 0; t < MG_MAX_ARRAYS; t++)
        if (mgArrs[t].live && mgArrs[t].host == p)
            return &mgArrs[t];
    return NULL;
}

/* zero-filled array of `bytes` (page-rounded) mapped TWICE: *view is what the caller
 * gets, *alias the always-writable second mapping (NULL if memfd is unavailable, then
 * the array is an ordinary anonymous mapping) */
static void mgMapTwice(size_t bytes, double **view, char **alias)
{
    *alias = NULL;
    int fd = -1;
#ifdef SYS_memfd_create
    fd = (int)syscall(SYS_memfd_create, "mgb-grid", 0u);
#endif
    if (fd >= 0 && ftruncate(fd, (off_t)bytes) == 0) {
        void *v = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        void *a = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
        close(fd);
        if (v != MAP_FAILED && a != MAP_FAILED) {
            *view = (double *)v;
            *alias = (char *)a;
            return;
        }
        if (v != MAP_FAILED) munmap(v, bytes);
        if (a != MAP_FAILED) munmap(a, bytes);
    } else if (fd >= 0) {
        close(fd);
    }
    void *p = mmap(NULL, bytes, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (p == MAP_FAILED) {
        fprintf(stderr, "mg_3d.h: cannot map %zu bytes\n", bytes);
        abort();
    }
    *view = (double *)p;
}

/* Is [plo, ..) still under the protection we left it in?  A caller may have freed the
 * array and got the same addresses back from a fresh mmap (then it is readable and
 * writable again).  write()/read() on a pipe make the kernel do the access check for us
 * without a signal: EFAULT = still protected. */
static int mgStillProtected(MgArr *a)
{
    if (!a->lazy || !a->plen || mgProbePipe[0] < 0)
        return 1;
    char b = 0;
    if (a->prot == PROT_NONE) {
        if (write(mgProbePipe[1], a->plo, 1) == 1) { /* readable: not ours any more */
            (void)!read(mgProbePipe[0], &b, 1);
            return 0;
        }
        return 1;
    }
    if (a->prot == PROT_READ) {
        b = *(volatile char *)a->plo; /* put the same byte back: non-destructive */
        if (write(mgProbePipe[1], &b, 1) != 1)
            return 1;
        if (read(mgProbePipe[0], a->plo, 1) == 1)
            return 0; /* writable: not ours any more */
        (void)!read(mgProbePipe[0], &b, 1);
        return 1;
    }
    return 1;
}

/* register an array; `own_pages`: the whole page-rounded range belongs to the array
 * (the solver's own level arrays); otherwise only fully covered pages are protected */
static MgArr *mgArrRegister(double *host, size_t bytes, char *alias, int own_pages,
                            mgb_solver *gpu, int level, int which, int state)
{
    MgArr *a = mgArrNew();
    if (!a)
        return NULL;
    a->host = host;
    a->bytes = bytes;
    a->alias = alias;
    a->gpu = gpu;
    a->level = level;
    a->which = which;
    a->state = state;
    a->prot = PROT_READ | PROT_WRITE;
    a->lazy = mgLazySync;
    const size_t pg = mgPageSize();
    if (own_pages) {
        a->plo = (char *)host;
        a->plen = mgPageRound(bytes);
    } else {
        const uintptr_t lo = ((uintptr_t)host + pg - 1) / pg * pg;
        const uintptr_t hi = ((uintptr_t)host + bytes) / pg * pg;
        a->plo = (char *)lo;
        a->plen = hi > lo ? (size_t)(hi - lo) : 0;
        if (a->lazy && a->plen) {
            a->head_n = (size_t)(lo - (uintptr_t)host) / sizeof(double);
            a->tail_first = (size_t)(hi - (uintptr_t)host) / sizeof(double);
            a->tail_n = bytes / sizeof(double) - a->tail_first;
            if (a->head_n + a->tail_n)
                a->frag = (double *)malloc((a->head_n + a->tail_n) * sizeof(double));
            mgFragRemember(a);
        } else {
            a->lazy = 0;
        }
    }
    /* pinned transfers (the bench's e2e figure is measured with pinned buffers) */
    a->pinned = mgb_pin_host(mgXfer(a), own_pages ? mgPageRound(bytes) : bytes) == 0;
    a->live = 1;
    return a;
}

static void mgArrRelease(MgArr *a, int sync_host)
{
    if (!a || !a->live)
        return;
    if (sync_host)
        mgArrPull(a);
    mgSetProt(a, PROT_READ | PROT_WRITE);
    if (a->pinned)
        mgb_unpin_host(mgXfer(a));
    free(a->frag);
    a->frag = NULL;
    a->live = 0;
}

#endif /* MGB_COHERENCE_H */
