// api.cu -- the extern "C" boundary (include/mgb.h) and the host-side V-cycle
// driver: level hierarchy in HBM, launch sequence of one V-cycle
// (mg_3d.h:1242-1362), CUDA-graph replay, CUDA-event stage timing.
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mgb.h"
#include "kernels.h"
#include "launch.h"
#include "nccl_dl.h"

using namespace mgb;

// ----------------------------------------------------------------------------
// errors
// ----------------------------------------------------------------------------
static thread_local std::string g_err;

static int fail(const char *fmt, ...)
{
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
    return 1;
}

// for the other translation units of the library (vtk.cu)
namespace mgb {
int set_error(const char *msg)
{
    g_err = msg;
    return 1;
}
}  // namespace mgb

#define CK(call)                                                                  \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess)                                                    \
            return fail("%s:%d: %s -> %s", __FILE__, __LINE__, #call,             \
                        cudaGetErrorString(e_));                                  \
    } while (0)

#define CKLAUNCH() CK(cudaGetLastError())

extern "C" const char *mgb_last_error(void) { return g_err.c_str(); }
extern "C" const char *mgb_version(void) { return "mgb 0.1 (sm_100a, colour-split fp64)"; }

extern "C" int mgb_device_count(int *count)
{
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) {
        *count = 0;
        return fail("cudaGetDeviceCount: %s", cudaGetErrorString(e));
    }
    *count = n;
    return 0;
}

// ----------------------------------------------------------------------------
// a colour-split array in device memory
// ----------------------------------------------------------------------------
struct DevArray {
    void *alloc = nullptr;
    double *base = nullptr;  // colour 0, local plane 0
    size_t bytes = 0;
};

static Geo make_geo(int ni, int nj, int nk, int li, int i0);
static Geo make_geo(int ni, int nj, int nk, int li, int i0)
{
    Geo g;
    g.ni = ni; g.nj = nj; g.nk = nk;
    g.li = li; g.i0 = i0;
    g.kh = mgb_half_pitch(nk);
    g.pj = (long long)nj * g.kh;
    g.cs = ((long long)li * g.pj + 15) & ~15LL;
    return g;
}

static int dev_alloc(const Geo &g, DevArray *a)
{
    a->bytes = sizeof(double) * (size_t)(2 * g.cs + 2 * MGB_GUARD);
    CK(cudaMalloc(&a->alloc, a->bytes));
    CK(cudaMemset(a->alloc, 0, a->bytes));
    a->base = (double *)a->alloc + MGB_GUARD;
    return 0;
}

static void dev_free(DevArray *a)
{
    if (a->alloc)
        cudaFree(a->alloc);
    a->alloc = nullptr;
    a->base = nullptr;
}

struct Level {
    Geo g;
    double h, hSq, invHsq;
    DevArray a[3];  // MGB_U, MGB_D, MGB_R
    // slab decomposition over i (multi-GPU): this rank owns the global planes
    // [own_lo, own_hi) and stores `lower` halo planes below and `upper` above
    bool dist = false;
    int own_lo = 0, own_hi = 0;
    int lower = 0, upper = 0;
    // local plane range of the owned INTERIOR planes (what sweeps update)
    int sweep_lo() const { return (own_lo > 1 ? own_lo : 1) - g.i0; }
    int sweep_hi() const { return (own_hi < g.ni - 1 ? own_hi : g.ni - 1) - g.i0; }
};

struct StageMark {
    int level, stage;
    cudaEvent_t a, b;
};

// the neighbouring rank on one side of the slab: its level arrays and flag
// block mapped into this process (CUDA IPC), and its local geometry
struct PeerSide {
    bool present = false;
    std::vector<double *> arr[2];  // [MGB_U / MGB_D][level]: colour 0, local plane 0
    std::vector<Geo> g;            // [level]
    unsigned long long *flags = nullptr;
    std::vector<void *> opened;    // cudaIpcOpenMemHandle results
};

struct mgb_solver {
    int device = 0;
    int L = 0, gs = 0;
    std::vector<Level> lv;
    cudaStream_t st = nullptr;
    double *partials = nullptr;
    double *d_scal = nullptr;  // device scalars
    double *h_scal = nullptr;  // pinned mirror
    double *stage = nullptr;   // natural-layout staging buffer
    size_t stage_n = 0;
    // coarsest operator
    int nc = 0;
    double *lu = nullptr;      // dense factor (what the reference calls A after LU)
    LuBand band{};             // the same factor in band form: what the solve reads
    double factor_secs = 0.;   // device time of build + factorisation (once, at create)
    // options
    int opt_graph = 1, opt_profile = 0, opt_fuse = 1;
    // levels 0..tail_top (<= ~33^3) run as ONE kernel (tail.cu); -1: none
    int opt_tail = 1, tail_top = -1;
    // coarse levels are not zeroed before pre-smoothing: their first half-sweep
    // takes the guess as 0 (launch_first_sweep_zero).  That needs the faces and
    // pads of the coarse u arrays to BE zero, which the cycle itself maintains;
    // API calls that can break it set coarse_dirty, and the next cycle zeroes once.
    int opt_zero_guess = 1;
    // inside the cycle the prolongation corrects only the colour the post-smoother
    // does not overwrite first (8 of its 17 B/DOF less)
    int opt_prolong_mask = 1;
    bool coarse_dirty = false;
    cudaGraphExec_t gexec = nullptr;
    long long graph_launches = 0;
    int eager_cycles = 0;  // partitioned solver: cycles run eagerly before the capture
    // bookkeeping
    long long launches = 0;
    std::vector<double> secs;  // L*7
    std::vector<int> calls;    // L*7
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_used = 0;
    std::vector<StageMark> marks;
    cudaEvent_t sw_a = nullptr, sw_b = nullptr;  // stopwatch
    // multi-GPU: one process per GPU, slabs over i, NCCL over NVLink
    int rank = 0, nranks = 1;
    int LD = 0;  // levels >= LD are slab-partitioned, levels < LD live on rank 0
    ncclComm_t comm = nullptr;
    long long nccl_calls = 0;
    // every NCCL call runs on its own stream, ordered against the kernel stream
    // by events, so that a halo exchange can overlap the interior of a sweep
    cudaStream_t st_comm = nullptr;
    cudaEvent_t ev_m2c = nullptr, ev_c2m = nullptr;
    // halo planes over NVLink peer memory instead of ncclSend/Recv, fused into the
    // compute kernels (halo.cuh).  xflags = this rank's flag block (XF_* slots), mapped
    // by the peers; n_* = sends issued so far in the open epoch (counted identically on
    // every rank, whether or not it has a neighbour on that side)
    int opt_p2p = 1;
    PeerSide low, up;
    unsigned long long *xflags = nullptr;
    unsigned long long n_up = 0, n_low = 0, n_gather = 0, n_norm = 0;
    // levels < LD: computed redundantly on EVERY rank from an all-gathered right-hand
    // side (P2P path) instead of on rank 0 with a gather + broadcast (NCCL path)
    bool replicate = false;
    struct PeerAny {
        unsigned long long *xf = nullptr;  // its flag block
        double *d_coarse = nullptr;        // its d array of level LD-1 (colour 0, plane 0)
    };
    std::vector<PeerAny> peers;  // [rank]
    std::vector<void *> peers_opened;
    // a halo wait that gives up (MGB_HALO_TIMEOUT_S, default 300 s, 0 = never) sets
    // bits here (host-mapped) instead of trapping; checked after every synchronisation
    unsigned int *h_halo_err = nullptr, *d_halo_err = nullptr;
    unsigned long long halo_timeout_ns = 300ULL * 1000000000ULL;
    bool is_dist() const { return nranks > 1; }
    // does this rank compute on level q?
    bool works_on(int q) const { return nranks == 1 || q >= LD || rank == 0 || replicate; }
};

struct LaunchScope {
    mgb_solver *s;
    long long start;
    explicit LaunchScope(mgb_solver *s_) : s(s_), start(launches_issued()) {}
    ~LaunchScope() { s->launches += launches_issued() - start; }
};

static int bind(const mgb_solver *s)
{
    if (!s)
        return fail("null solver");
    CK(cudaSetDevice(s->device));
    return 0;
}

// after a synchronisation: did a halo wait give up on a neighbour?
static int halo_status(mgb_solver *s)
{
    if (!s->h_halo_err || !*s->h_halo_err)
        return 0;
    const unsigned int e = *s->h_halo_err;
    *s->h_halo_err = 0;
    return fail("rank %d: halo exchange timed out after %.0f s waiting for the %s%s%s neighbour "
                "(MGB_HALO_TIMEOUT_S; all array-changing calls of a partitioned solver are "
                "collective over the ranks); the level arrays are no longer consistent",
                s->rank, 1e-9 * (double)s->halo_timeout_ns, (e & 1) ? "lower" : "",
                (e & 3) == 3 ? " and " : "", (e & 2) ? "upper" : "");
}

static int check_level(const mgb_solver *s, int level, int which)
{
    if (level < 0 || level >= s->L)
        return fail("level %d out of range [0,%d)", level, s->L);
    if (which < 0 || which > 2)
        return fail("array selector %d out of range", which);
    return 0;
}

// ----------------------------------------------------------------------------
// lifecycle
// ----------------------------------------------------------------------------
extern "C" int mgb_destroy(mgb_solver *s)
{
    if (!s)
        return 0;
    cudaSetDevice(s->device);
    if (s->st)
        cudaStreamSynchronize(s->st);
    if (s->comm && s->xflags) {
        // collective: the neighbours have my arrays mapped (CUDA IPC) and may still be
        // pushing into them; nobody frees anything before everybody has got here
        nccl().AllReduce(s->xflags + XF_SCRATCH, s->xflags + XF_SCRATCH, 1, ncclInt, ncclSum,
                         s->comm, s->st);
        cudaStreamSynchronize(s->st);
    }
    if (s->h_halo_err) cudaFreeHost(s->h_halo_err);
    if (s->gexec)
        cudaGraphExecDestroy(s->gexec);
    for (auto &l : s->lv)
        for (int w = 0; w < 3; w++)
            dev_free(&l.a[w]);
    for (auto e : s->ev_pool)
        cudaEventDestroy(e);
    if (s->sw_a) cudaEventDestroy(s->sw_a);
    if (s->sw_b) cudaEventDestroy(s->sw_b);
    if (s->partials) cudaFree(s->partials);
    if (s->d_scal) cudaFree(s->d_scal);
    if (s->h_scal) cudaFreeHost(s->h_scal);
    if (s->stage) cudaFree(s->stage);
    if (s->lu) cudaFree(s->lu);
    lu_band_free(&s->band);
    for (PeerSide *ps : {&s->low, &s->up})
        for (void *p : ps->opened)
            cudaIpcCloseMemHandle(p);
    for (void *p : s->peers_opened)
        cudaIpcCloseMemHandle(p);
    if (s->xflags) cudaFree(s->xflags);
    if (s->comm) nccl().CommDestroy(s->comm);
    if (s->st_comm) { cudaStreamSynchronize(s->st_comm); cudaStreamDestroy(s->st_comm); }
    if (s->ev_m2c) cudaEventDestroy(s->ev_m2c);
    if (s->ev_c2m) cudaEventDestroy(s->ev_c2m);
    if (s->st) cudaStreamDestroy(s->st);
    delete s;
    return 0;
}

// ----------------------------------------------------------------------------
// slab planning (pure host arithmetic; also exported for the CPU-side tests)
//
// Level l has ni-1 = (ci-1)*2^l intervals along i.  Rank p of P owns the planes
// [p*w, (p+1)*w) with w = (ni-1)/P, the last rank also the closing plane ni-1;
// because w doubles with every refinement the cuts of all distributed levels
// coincide (fine plane 2I belongs to the owner of coarse plane I).
// ----------------------------------------------------------------------------
extern "C" int mgb_plan_slab(int ni, int nranks, int rank, int *own_lo, int *own_hi)
{
    if (nranks < 1 || rank < 0 || rank >= nranks || ni < 3)
        return fail("mgb_plan_slab: bad arguments");
    if ((ni - 1) % nranks != 0)
        return fail("%d intervals do not divide over %d ranks", ni - 1, nranks);
    const int w = (ni - 1) / nranks;
    *own_lo = rank * w;
    *own_hi = rank == nranks - 1 ? ni : (rank + 1) * w;
    return 0;
}

// first (coarsest) slab-partitioned level: every rank must own an even number
// of planes (so the cut survives one more coarsening for the restriction
// target), at least `min_planes` of them and at least `min_points` grid points
// -- below that a level is latency-bound and cheaper on one GPU than split
// with halo messages; level 0 always stays on rank 0
extern "C" int mgb_plan_first_dist_level(int ci, int cj, int ck, int levels, int nranks,
                                         int min_planes, long long min_points)
{
    if (nranks <= 1)
        return 0;
    if (min_planes < 2)
        min_planes = 2;
    for (int l = 1; l < levels; l++) {
        const long long iv = (long long)(ci - 1) << l;
        if (iv % nranks)
            continue;
        const long long w = iv / nranks;
        const long long plane = (((long long)(cj - 1) << l) + 1) * (((long long)(ck - 1) << l) + 1);
        if (w >= min_planes && w % 2 == 0 && w * plane >= min_points)
            return l;
    }
    return levels;  // nothing can be partitioned
}

// ----------------------------------------------------------------------------
// peer mapping of the neighbours' arrays (CUDA IPC) for the P2P halo exchange
// ----------------------------------------------------------------------------
struct IpcRec {
    cudaIpcMemHandle_t h;
    unsigned long long offset;  // of the pointer inside its cudaMalloc block
    unsigned long long pad;
};

typedef int (*MemRangeFn)(unsigned long long *, size_t *, unsigned long long);

static bool ipc_record(MemRangeFn range, void *ptr, IpcRec *rec)
{
    unsigned long long base = 0;
    size_t size = 0;
    if (range(&base, &size, (unsigned long long)ptr) != 0)
        return false;
    memset(rec, 0, sizeof *rec);
    rec->offset = (unsigned long long)ptr - base;
    return cudaIpcGetMemHandle(&rec->h, (void *)base) == cudaSuccess;
}

static Geo make_geo(int ni, int nj, int nk, int li, int i0);

// local geometry of level `lv` on rank r (the same arithmetic create_impl uses)
static Geo rank_geo(const Geo &mine, int nranks, int r)
{
    int lo = 0, hi = 0;
    mgb_plan_slab(mine.ni, nranks, r, &lo, &hi);
    const int lower = r > 0 ? 2 : 0, upper = r < nranks - 1 ? 1 : 0;
    return make_geo(mine.ni, mine.nj, mine.nk, hi - lo + lower + upper, lo - lower);
}

// returns 0 when everything is mapped; any failure leaves the solver on the NCCL
// path (the caller makes the decision collective).  records per rank: [0] its flag
// block, [1 + 2*(l-LD) + w] array w (U, D) of partitioned level l, [count-1] the rhs
// array of the first replicated level LD-1
static int setup_p2p_local(mgb_solver *s, std::vector<IpcRec> &all, int count)
{
    // one cudaMalloc block is opened once, however many pointers into it are needed
    std::vector<std::pair<IpcRec, void *>> seen;
    auto map = [&](const IpcRec &rec) -> void * {
        for (auto &e : seen)
            if (!memcmp(&e.first.h, &rec.h, sizeof rec.h))
                return (char *)e.second + rec.offset;
        void *base = nullptr;
        if (cudaIpcOpenMemHandle(&base, rec.h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        s->peers_opened.push_back(base);
        seen.push_back({rec, base});
        return (char *)base + rec.offset;
    };
    for (int side = 0; side < 2; side++) {
        const int r = side == 0 ? s->rank - 1 : s->rank + 1;
        PeerSide &ps = side == 0 ? s->low : s->up;
        if (r < 0 || r >= s->nranks)
            continue;
        ps.arr[0].assign(s->L, nullptr);
        ps.arr[1].assign(s->L, nullptr);
        ps.g.assign(s->L, Geo{});
        const IpcRec *recs = &all[(size_t)r * count];
        ps.flags = (unsigned long long *)map(recs[0]);
        if (!ps.flags)
            return 1;
        for (int l = s->LD; l < s->L; l++)
            for (int w = 0; w < 2; w++) {
                void *p = map(recs[1 + 2 * (l - s->LD) + w]);
                if (!p)
                    return 1;
                ps.arr[w][l] = (double *)p + MGB_GUARD;
                ps.g[l] = rank_geo(s->lv[l].g, s->nranks, r);
            }
        ps.present = true;
    }
    // the all-gather in front of the replicated coarse levels and the exchange of norm
    // partials talk to ALL peers, not just to the neighbours
    s->peers.assign(s->nranks, mgb_solver::PeerAny{});
    for (int r = 0; r < s->nranks; r++) {
        if (r == s->rank)
            continue;
        const IpcRec *recs = &all[(size_t)r * count];
        void *xf = map(recs[0]), *dc = map(recs[count - 1]);
        if (!xf || !dc)
            return 1;
        s->peers[r].xf = (unsigned long long *)xf;
        s->peers[r].d_coarse = (double *)dc + MGB_GUARD;
    }
    return 0;
}

static int setup_p2p(mgb_solver *s)
{
    int ok = 1;
    MemRangeFn range = nullptr;
    {
        void *fn = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuMemGetAddressRange", &fn, cudaEnableDefault, &q) ==
                cudaSuccess && q == cudaDriverEntryPointSuccess)
            range = (MemRangeFn)fn;
        else
            ok = 0;
    }
    const int count = 2 + 2 * (s->L - s->LD);
    std::vector<IpcRec> mine(count), all((size_t)count * s->nranks);
    if (cudaMalloc(&s->xflags, 4096) != cudaSuccess || cudaMemset(s->xflags, 0, 4096) != cudaSuccess)
        ok = 0;
    if (ok) {  // epochs start at 1 (0 = "nothing sent yet" for every flag)
        const unsigned long long one = 1;
        ok = cudaMemcpy(s->xflags + XF_EPOCH, &one, sizeof one, cudaMemcpyHostToDevice) == cudaSuccess;
    }
    if (ok) {
        ok = ipc_record(range, s->xflags, &mine[0]);
        for (int l = s->LD; ok && l < s->L; l++)
            for (int w = 0; ok && w < 2; w++)
                ok = ipc_record(range, s->lv[l].a[w].alloc, &mine[1 + 2 * (l - s->LD) + w]);
        if (ok)
            ok = ipc_record(range, s->lv[s->LD - 1].a[MGB_D].alloc, &mine[count - 1]);
    }
    // everybody learns everybody's handles (the exchange itself is collective
    // even if this rank already failed)
    char *dsend = nullptr, *drecv = nullptr;
    const size_t bytes = sizeof(IpcRec) * count;
    CK(cudaMalloc(&dsend, bytes));
    CK(cudaMalloc(&drecv, bytes * s->nranks));
    CK(cudaMemcpy(dsend, mine.data(), bytes, cudaMemcpyHostToDevice));
    ncclResult_t r = nccl().AllGather(dsend, drecv, bytes, ncclChar, s->comm, s->st);
    if (r != ncclSuccess)
        return fail("ncclAllGather: %s", nccl().GetErrorString(r));
    CK(cudaStreamSynchronize(s->st));
    CK(cudaMemcpy(all.data(), drecv, bytes * s->nranks, cudaMemcpyDeviceToHost));
    if (ok && setup_p2p_local(s, all, count))
        ok = 0;
    // collective decision: P2P only if every rank mapped its neighbours
    int *dflag = (int *)dsend;
    CK(cudaMemcpy(dflag, &ok, sizeof ok, cudaMemcpyHostToDevice));
    r = nccl().AllReduce(dflag, dflag, 1, ncclInt, ncclMin, s->comm, s->st);
    if (r != ncclSuccess)
        return fail("ncclAllReduce: %s", nccl().GetErrorString(r));
    CK(cudaStreamSynchronize(s->st));
    CK(cudaMemcpy(&ok, dflag, sizeof ok, cudaMemcpyDeviceToHost));
    cudaFree(dsend);
    cudaFree(drecv);
    s->opt_p2p = ok;
    s->replicate = ok != 0;
    if (getenv("MGB_VERBOSE"))
        fprintf(stderr, "mgb: rank %d halo exchange over %s\n", s->rank,
                ok ? "NVLink peer memory (P2P stores + flags)" : "ncclSend/ncclRecv");
    return 0;
}

static bool pow2plus1(int n)
{
    return n >= 3 && ((n - 1) & (n - 2)) == 0;
}

static int create_impl(mgb_solver **out, int ci, int cj, int ck, int levels, int gs_iters,
                       int device, int rank, int nranks, const void *uid, int min_planes,
                       long long min_points)
{
    if (!out)
        return fail("out is null");
    *out = nullptr;
    if (levels < 1 || levels > 24)
        return fail("levels=%d out of range", levels);
    if (ci < 3 || cj < 3 || ck < 3)
        return fail("coarse extents must be >= 3 (got %d %d %d)", ci, cj, ck);
    // mg_3d.h:120-123: (coarse-1) must be a power of two (only enforced when a
    // hierarchy is actually built)
    if (levels > 1 && !(pow2plus1(ci) && pow2plus1(cj) && pow2plus1(ck)))
        return fail("coarse extents minus one must be powers of two (got %d %d %d)", ci,
                    cj, ck);
    if (gs_iters < 0)
        return fail("gs_iters < 0");
    int ndev = 0;
    if (mgb_device_count(&ndev))
        return 1;
    if (ndev < 1)
        return fail("no CUDA device: libmgb has no CPU fallback");
    if (device < 0 || device >= ndev)
        return fail("device %d out of range [0,%d)", device, ndev);
    const long long nfine_k = (long long)(ck - 1) * (1LL << (levels - 1)) + 1;
    const long long nfine_j = (long long)(cj - 1) * (1LL << (levels - 1)) + 1;
    const long long nfine_i = (long long)(ci - 1) * (1LL << (levels - 1)) + 1;
    if (nfine_k > 1 << 20 || nfine_j > 1 << 20 || nfine_i > 1 << 20)
        return fail("grid too large");

    mgb_solver *s = new mgb_solver();
    s->device = device;
    s->L = levels;
    s->gs = gs_iters;
    s->rank = rank;
    s->nranks = nranks;
    if (nranks > 1) {
        if (levels < 2) {
            delete s;
            return fail("a partitioned solver needs at least 2 levels");
        }
        pdl_dist_active() = true;  // launch.h: plain launches in a process with peers
        s->LD = mgb_plan_first_dist_level(ci, cj, ck, levels, nranks, min_planes, min_points);
        if (s->LD >= levels) {
            delete s;
            return fail("no level of this hierarchy can be split over %d ranks with >= %d "
                        "planes and >= %lld points each", nranks, min_planes, min_points);
        }
        s->opt_graph = 0;  // NCCL calls are issued eagerly between the kernels
    }
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) {
        delete s;
        return fail("cudaSetDevice(%d): %s", device, cudaGetErrorString(e));
    }
#define CKD(call)                                                                 \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            fail("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            mgb_destroy(s);                                                       \
            return 1;                                                             \
        }                                                                         \
    } while (0)
    CKD(cudaStreamCreateWithFlags(&s->st, cudaStreamNonBlocking));
    CKD(cudaMalloc(&s->partials, sizeof(double) * kMaxPartials));
    CKD(cudaMalloc(&s->d_scal, sizeof(double) * 16));
    CKD(cudaMemset(s->d_scal, 0, sizeof(double) * 16));
    CKD(cudaMallocHost(&s->h_scal, sizeof(double) * 16));

    s->lv.resize(levels);
    const double hfine = 1. / (double)(nfine_k - 1);  // mg_3d.h:143 with GRID_LENGTH=1
    for (int l = 0; l < levels; l++) {
        Level &lv = s->lv[l];
        const int ni = (ci - 1) * (1 << l) + 1;  // mg_3d.h:41
        const int nj = (cj - 1) * (1 << l) + 1;
        const int nk = (ck - 1) * (1 << l) + 1;
        lv.own_lo = 0;
        lv.own_hi = ni;
        if (nranks > 1 && l >= s->LD) {
            lv.dist = true;
            mgb_plan_slab(ni, nranks, rank, &lv.own_lo, &lv.own_hi);
            // two planes below (the fused residual+restriction evaluates the
            // residual on plane own_lo-1, which reads own_lo-2), one above
            lv.lower = rank > 0 ? 2 : 0;
            lv.upper = rank < nranks - 1 ? 1 : 0;
            lv.g = make_geo(ni, nj, nk, lv.own_hi - lv.own_lo + lv.lower + lv.upper,
                            lv.own_lo - lv.lower);
        } else {
            lv.g = make_geo(ni, nj, nk, ni, 0);
        }
        // h_l = h_fine * 2^(L-1-l), formed by repeated doubling like
        // mg_3d.h:1303 (exact in binary either way)
        double h = hfine;
        for (int t = levels - 1; t > l; t--)
            h = 2 * h;
        lv.h = h;
        lv.hSq = h * h;            // mg_3d.h:644
        lv.invHsq = 1. / (h * h);  // mg_3d.h:797
        for (int w = 0; w < 3; w++)
            if (dev_alloc(lv.g, &lv.a[w])) {
                mgb_destroy(s);
                return 1;
            }
    }
    {
        // deepest run of small, unpartitioned levels: one kernel does them all
        // (one SM hides L2 latency for <= 17^3 points per level; at 33^3 the
        // per-stage kernels spread over the GPU are faster again -- measured)
        const long long max_pts = getenv("MGB_TAIL_POINTS") ? atoll(getenv("MGB_TAIL_POINTS")) : 5000;
        if (getenv("MGB_TAIL"))
            s->opt_tail = atoi(getenv("MGB_TAIL")) != 0;
        if (getenv("MGB_ZERO_GUESS"))
            s->opt_zero_guess = atoi(getenv("MGB_ZERO_GUESS")) != 0;
        if (getenv("MGB_PROLONG_MASK"))
            s->opt_prolong_mask = atoi(getenv("MGB_PROLONG_MASK")) != 0;
        const int lim = nranks > 1 ? s->LD - 1 : levels - 2;  // strictly coarse, on one GPU
        for (int l = 0; l <= lim && l < 8; l++) {
            const Geo &g = s->lv[l].g;
            if ((long long)g.ni * g.nj * g.nk > max_pts)
                break;
            s->tail_top = l;
        }
        if ((long long)ci * cj * ck > 1024)
            s->tail_top = -1;
    }
    s->secs.assign((size_t)levels * MGB_NUM_STAGES, 0.);
    s->calls.assign((size_t)levels * MGB_NUM_STAGES, 0);

    // coarsest operator (mg_3d.h:281-288), built + factorised on the device
    const long long nc = (long long)ci * cj * ck;
    if (nc <= MGB_MAX_DENSE_N) {
        s->nc = (int)nc;
        // half bandwidth of the 7-point operator in the ordering p = (i*nj+j)*nk+k
        CKD(cudaMalloc(&s->lu, sizeof(double) * nc * nc));
        if (lu_band_alloc(&s->band, s->nc, cj * ck < s->nc - 1 ? cj * ck : s->nc - 1)) {
            fail("cudaMalloc of the coarse factor's tiles failed");
            mgb_destroy(s);
            return 1;
        }
        LaunchScope ls(s);
        cudaEvent_t e0, e1;
        CKD(cudaEventCreate(&e0));
        CKD(cudaEventCreate(&e1));
        CKD(cudaEventRecord(e0, s->st));
        launch_coarse_matrix(s->lu, ci, cj, ck, s->lv[0].h, s->st);
        launch_lu_factor_band(s->lu, s->nc, s->band.bw, s->st);
        launch_lu_extract_band(s->lu, s->band, s->st);
        CKD(cudaEventRecord(e1, s->st));
        CKD(cudaGetLastError());
        CKD(cudaStreamSynchronize(s->st));
        float ms = 0.f;
        CKD(cudaEventElapsedTime(&ms, e0, e1));
        s->factor_secs = 1e-3 * ms;
        cudaEventDestroy(e0);
        cudaEventDestroy(e1);
    } else if (levels > 1) {
        mgb_destroy(s);
        return fail("coarsest grid has %lld unknowns > MGB_MAX_DENSE_N", nc);
    }
    if (nranks > 1) {
        if (!nccl().load()) {
            fail("NCCL: %s", nccl().error);
            mgb_destroy(s);
            return 1;
        }
        CKD(cudaStreamCreateWithFlags(&s->st_comm, cudaStreamNonBlocking));
        CKD(cudaEventCreateWithFlags(&s->ev_m2c, cudaEventDisableTiming));
        CKD(cudaEventCreateWithFlags(&s->ev_c2m, cudaEventDisableTiming));
        ncclUniqueId id;
        memcpy(&id, uid, sizeof id);
        ncclResult_t r = nccl().CommInitRank(&s->comm, nranks, id, rank);
        if (r != ncclSuccess) {
            fail("ncclCommInitRank: %s", nccl().GetErrorString(r));
            s->comm = nullptr;
            mgb_destroy(s);
            return 1;
        }
        if (getenv("MGB_P2P"))
            s->opt_p2p = atoi(getenv("MGB_P2P")) != 0;
        if (getenv("MGB_HALO_TIMEOUT_S"))
            s->halo_timeout_ns = (unsigned long long)(atof(getenv("MGB_HALO_TIMEOUT_S")) * 1e9);
        CKD(cudaHostAlloc(&s->h_halo_err, sizeof(unsigned int), cudaHostAllocMapped));
        *s->h_halo_err = 0;
        CKD(cudaHostGetDevicePointer(&s->d_halo_err, s->h_halo_err, 0));
        if (s->opt_p2p && setup_p2p(s)) {
            mgb_destroy(s);
            return 1;
        }
        // with the halo exchange as plain kernels (device-side sequence numbers)
        // the whole partitioned cycle replays as one CUDA graph per rank; the
        // three NCCL calls left in it are capturable
        s->opt_graph = s->opt_p2p;
        if (getenv("MGB_DIST_GRAPH"))
            s->opt_graph = s->opt_p2p && atoi(getenv("MGB_DIST_GRAPH")) != 0;
    }
#undef CKD
    *out = s;
    return 0;
}

extern "C" int mgb_create(mgb_solver **out, int ci, int cj, int ck, int levels,
                          int gs_iters, int device)
{
    return create_impl(out, ci, cj, ck, levels, gs_iters, device, 0, 1, nullptr, 0, 0);
}

extern "C" int mgb_nccl_unique_id(void *out128)
{
    if (!out128)
        return fail("out128 is null");
    if (!nccl().load())
        return fail("NCCL: %s", nccl().error);
    ncclUniqueId id;
    ncclResult_t r = nccl().GetUniqueId(&id);
    if (r != ncclSuccess)
        return fail("ncclGetUniqueId: %s", nccl().GetErrorString(r));
    memcpy(out128, &id, sizeof id);
    return 0;
}

extern "C" int mgb_create_dist(mgb_solver **out, int ci, int cj, int ck, int levels,
                               int gs_iters, int device, int rank, int nranks,
                               const void *nccl_uid128, int min_planes_per_rank,
                               long long min_points_per_rank)
{
    if (nranks < 1 || rank < 0 || rank >= nranks)
        return fail("rank %d / nranks %d out of range", rank, nranks);
    if (nranks > 1 && !nccl_uid128)
        return fail("nccl_uid128 is null");
    if (nranks & (nranks - 1))
        return fail("nranks must be a power of two (got %d)", nranks);
    // defaults: a level is partitioned while every rank owns >= 2 planes (an even number)
    // and the level as a whole has >= 2^22 points (8 GPUs, round 2: 1025^3 strong 4.686 ms
    // with 129^3 partitioned, 4.663 replicated; weak 4.144 -> 4.127); the levels below are
    // replicated:
    // computed redundantly on every rank from an all-gathered right-hand side.  Measured on
    // 2 B200s: a half-sweep on a slab of a small level costs kernel + ~7 us of flag latency
    // (every step has to hear from the neighbour), the same kernel on the whole level ~4 us
    // while that level is latency-bound (up to ~65^3..129^3 points).  The threshold is on
    // the GLOBAL size because a replicated level costs every rank the whole level, which
    // grows with the number of GPUs in weak scaling.
    return create_impl(out, ci, cj, ck, levels, gs_iters, device, rank, nranks, nccl_uid128,
                       min_planes_per_rank > 0 ? min_planes_per_rank : 2,
                       min_points_per_rank >= 0 ? min_points_per_rank
                                                : (1LL << 22) / (nranks > 0 ? nranks : 1));
}

extern "C" int mgb_dist_info(const mgb_solver *s, int *rank, int *nranks, int *first_dist_level)
{
    if (!s)
        return fail("null solver");
    if (rank) *rank = s->rank;
    if (nranks) *nranks = s->nranks;
    if (first_dist_level) *first_dist_level = s->LD;
    return 0;
}

extern "C" int mgb_local_range(const mgb_solver *s, int level, int *i0, int *li, int *own_lo,
                               int *own_hi)
{
    if (!s)
        return fail("null solver");
    if (check_level(s, level, 0))
        return 1;
    const Level &lv = s->lv[level];
    if (i0) *i0 = lv.g.i0;
    if (li) *li = lv.g.li;
    if (own_lo) *own_lo = lv.own_lo;
    if (own_hi) *own_hi = lv.own_hi;
    return 0;
}

extern "C" int mgb_levels(const mgb_solver *s) { return s ? s->L : 0; }

extern "C" int mgb_dims(const mgb_solver *s, int level, int *ni, int *nj, int *nk)
{
    if (!s)
        return fail("null solver");
    if (check_level(s, level, 0))
        return 1;
    *ni = s->lv[level].g.ni;
    *nj = s->lv[level].g.nj;
    *nk = s->lv[level].g.nk;
    return 0;
}

extern "C" double mgb_spacing(const mgb_solver *s, int level)
{
    if (!s || level < 0 || level >= s->L)
        return 0.;
    return s->lv[level].h;
}

static void drop_graph(mgb_solver *s)
{
    if (s->gexec) {
        cudaGraphExecDestroy(s->gexec);
        s->gexec = nullptr;
    }
}

extern "C" int mgb_set_option(mgb_solver *s, int key, int value)
{
    if (bind(s))
        return 1;
    switch (key) {
    case MGB_OPT_GRAPH: s->opt_graph = value != 0 && (!s->is_dist() || s->opt_p2p); break;
    case MGB_OPT_PROFILE: s->opt_profile = value != 0; break;
    case MGB_OPT_FUSE:
        if (s->is_dist() && value <= 0)
            return fail("the partitioned solver only has the fused residual+restriction");
        s->opt_fuse = value < 0 ? 0 : (value > 2 ? 2 : value);
        drop_graph(s);
        break;
    case MGB_OPT_GRAPH_LEVELS: break;
    case MGB_OPT_TAIL:
        s->opt_tail = value != 0;
        drop_graph(s);
        break;
    case MGB_OPT_ZERO_GUESS:
        s->opt_zero_guess = value != 0;
        drop_graph(s);
        break;
    case MGB_OPT_PROLONG_MASK:
        s->opt_prolong_mask = value != 0;
        drop_graph(s);
        break;
    default: return fail("unknown option %d", key);
    }
    return 0;
}

extern "C" int mgb_set_global(int key, long long value)
{
    switch (key) {
    case MGB_G_TILE: tile_set(value != 0, -1); break;
    case MGB_G_TILE_MIN_PLANE: tile_set(-1, value < 0 ? 0 : value); break;
    case MGB_G_GSLEX_TILE: gs_lex_set_mode((int)value); break;
    default: return fail("unknown global option %d", key);
    }
    return 0;
}

extern "C" int mgb_tile_plan(int kind, int ni, int nj, int nk, long long *out12)
{
    if (kind < 0 || kind > 3 || ni < 3 || nj < 3 || nk < 3 || !out12)
        return fail("mgb_tile_plan: bad arguments");
    const Geo gf = make_geo(ni, nj, nk, ni, 0);
    if (kind == 1) {
        if (!(ni & 1) || !(nj & 1) || !(nk & 1))
            return fail("mgb_tile_plan: restriction needs odd extents");
        const Geo gc = make_geo((ni + 1) / 2, (nj + 1) / 2, (nk + 1) / 2, (ni + 1) / 2, 0);
        tile_plan_query(1, gf, &gc, 0, gc.ni, out12);
    } else if (kind == 3) {
        tile_plan_query(3, gf, nullptr, 0, ni, out12);
    } else {
        tile_plan_query(kind, gf, nullptr, 1, ni - 1, out12);
    }
    return 0;
}

extern "C" int mgb_sync(mgb_solver *s)
{
    if (bind(s))
        return 1;
    CK(cudaStreamSynchronize(s->st));
    return halo_status(s);
}

// ----------------------------------------------------------------------------
// level arrays across the boundary
// ----------------------------------------------------------------------------
static void halo_fence(mgb_solver *s, Level &lv);  // P2P halo exchange, below
static void close_epoch(mgb_solver *s);

static int need_stage(mgb_solver *s, size_t n)
{
    if (s->stage_n >= n)
        return 0;
    if (s->stage) {
        CK(cudaStreamSynchronize(s->st));
        CK(cudaFree(s->stage));
        s->stage = nullptr;
        s->stage_n = 0;
    }
    CK(cudaMalloc(&s->stage, sizeof(double) * n));
    s->stage_n = n;
    return 0;
}

extern "C" int mgb_upload(mgb_solver *s, int level, int which, const double *host)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    if (!host)
        return fail("host pointer is null");
    Level &lv = s->lv[level];
    if (level < s->L - 1)
        s->coarse_dirty = true;  // may put non-zero values on a coarse level's faces
    const size_t n = (size_t)lv.g.li * lv.g.nj * lv.g.nk;  // local planes [i0, i0+li)
    if (need_stage(s, n))
        return 1;
    LaunchScope ls(s);
    CK(cudaMemcpyAsync(s->stage, host, sizeof(double) * n, cudaMemcpyHostToDevice, s->st));
    launch_pack(lv.g, s->stage, lv.a[which].base, s->st);
    halo_fence(s, lv);
    close_epoch(s);
    CKLAUNCH();
    CK(cudaStreamSynchronize(s->st));
    return halo_status(s);
}

extern "C" int mgb_download(mgb_solver *s, int level, int which, double *host)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    if (!host)
        return fail("host pointer is null");
    Level &lv = s->lv[level];
    const size_t n = (size_t)lv.g.li * lv.g.nj * lv.g.nk;
    if (need_stage(s, n))
        return 1;
    LaunchScope ls(s);
    launch_unpack(lv.g, lv.a[which].base, s->stage, s->st);
    CKLAUNCH();
    CK(cudaMemcpyAsync(host, s->stage, sizeof(double) * n, cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return 0;
}

// a contiguous run of the natural layout: doubles [first, first+count) of the local
// planes (the unprotectable head / tail fragments of caller-owned arrays, compat/)
static int range_args(mgb_solver *s, int level, int which, long long first, long long count,
                      const void *host)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    if (!host)
        return fail("host pointer is null");
    const Level &lv = s->lv[level];
    const long long n = (long long)lv.g.li * lv.g.nj * lv.g.nk;
    if (first < 0 || count < 1 || first + count > n)
        return fail("range [%lld, %lld) outside the %lld local doubles", first, first + count, n);
    return 0;
}

extern "C" int mgb_upload_range(mgb_solver *s, int level, int which, long long first,
                                long long count, const double *host)
{
    if (range_args(s, level, which, first, count, host) || need_stage(s, (size_t)count))
        return 1;
    Level &lv = s->lv[level];
    if (level < s->L - 1)
        s->coarse_dirty = true;
    LaunchScope ls(s);
    CK(cudaMemcpyAsync(s->stage, host, sizeof(double) * count, cudaMemcpyHostToDevice, s->st));
    launch_pack_range(lv.g, s->stage, lv.a[which].base, first, count, s->st);
    halo_fence(s, lv);
    close_epoch(s);
    CKLAUNCH();
    CK(cudaStreamSynchronize(s->st));
    return halo_status(s);
}

extern "C" int mgb_download_range(mgb_solver *s, int level, int which, long long first,
                                  long long count, double *host)
{
    if (range_args(s, level, which, first, count, host) || need_stage(s, (size_t)count))
        return 1;
    Level &lv = s->lv[level];
    LaunchScope ls(s);
    launch_unpack_range(lv.g, lv.a[which].base, s->stage, first, count, s->st);
    CKLAUNCH();
    CK(cudaMemcpyAsync(host, s->stage, sizeof(double) * count, cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    return 0;
}

extern "C" int mgb_zero(mgb_solver *s, int level, int which)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    Level &lv = s->lv[level];
    CK(cudaMemsetAsync(lv.a[which].base, 0, sizeof(double) * 2 * lv.g.cs, s->st));
    halo_fence(s, lv);
    close_epoch(s);
    return 0;
}

extern "C" int mgb_set_dirichlet(mgb_solver *s, int level, int which)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    Level &lv = s->lv[level];
    if (level < s->L - 1)
        s->coarse_dirty = true;
    LaunchScope ls(s);
    launch_set_dirichlet(lv.g, lv.a[which].base, lv.h, s->st);
    halo_fence(s, lv);
    close_epoch(s);
    CKLAUNCH();
    return 0;
}

// ----------------------------------------------------------------------------
// NCCL plumbing of the partitioned solver
// ----------------------------------------------------------------------------
static bool g_nccl_failed = false;
#define NC(call)                                                       \
    do {                                                               \
        ncclResult_t r_ = (call);                                      \
        if (r_ != ncclSuccess && !g_nccl_failed) {                     \
            g_nccl_failed = true;                                      \
            fail("%s -> %s", #call, nccl().GetErrorString(r_));        \
        }                                                              \
    } while (0)

static int nccl_status()
{
    if (g_nccl_failed) {
        g_nccl_failed = false;
        return 1;
    }
    return 0;
}

// the communication stream picks up after everything enqueued on the kernel
// stream so far ...
// (with the P2P halo exchange the few remaining NCCL calls -- norm all-reduce,
// agglomeration gather / broadcast -- simply run on the kernel stream)
static cudaStream_t nccl_stream(mgb_solver *s) { return s->opt_p2p ? s->st : s->st_comm; }
static void comm_begin(mgb_solver *s)
{
    if (s->opt_p2p)
        return;
    cudaEventRecord(s->ev_m2c, s->st);
    cudaStreamWaitEvent(s->st_comm, s->ev_m2c, 0);
}
// ... and the kernel stream continues once the communication enqueued so far is done
static void comm_end(mgb_solver *s)
{
    if (s->opt_p2p)
        return;
    cudaEventRecord(s->ev_c2m, s->st_comm);
    cudaStreamWaitEvent(s->st, s->ev_c2m, 0);
}

// ---- the fused halo protocol (halo.cuh): host side ----------------------------
// What a kernel on a partitioned level gets: "wait for everything both neighbours were
// asked to send so far" (no pushes yet -- the caller adds those AFTER taking this, so
// that a kernel never waits for its own sends' counterparts)
static HaloCtl make_ctl(mgb_solver *s, const Level &lv)
{
    HaloCtl h{};
    h.push_up.plane[0] = h.push_up.plane[1] = -1;
    h.push_low.plane[0] = h.push_low.plane[1] = -1;
    if (!lv.dist || !s->opt_p2p)
        return h;  // epoch == nullptr: no halo work
    h.epoch = s->xflags + XF_EPOCH;
    // the lower neighbour's up-sends are counted by my own n_up (same program on every rank)
    if (s->rank > 0)
        h.wait_low = HaloWait{s->xflags + XF_FROM_LOW, s->xflags + XF_PREV_UP, s->n_up};
    if (s->rank < s->nranks - 1)
        h.wait_up = HaloWait{s->xflags + XF_FROM_UP, s->xflags + XF_PREV_LOW, s->n_low};
    h.timeout_ns = s->halo_timeout_ns;
    h.err = s->d_halo_err;
    return h;
}

// a half-sweep of `colour` on partitioned level q also feeds the neighbours' halos: its
// planes own_hi-2 and own_hi-1 are the upper neighbour's two lower halo planes, own_lo is
// the lower neighbour's upper halo plane
static void add_sweep_push(mgb_solver *s, int q, int colour, HaloCtl &h)
{
    Level &lv = s->lv[q];
    const Geo &g = lv.g;
    ++s->n_up;
    ++s->n_low;
    if (s->rank < s->nranks - 1) {
        const Geo &gn = s->up.g[q];
        HaloPush &p = h.push_up;
        for (int k = 0; k < 2; k++) {
            const int gp = lv.own_hi - 1 - k;  // global plane
            p.plane[k] = gp - g.i0;
            p.dst[k] = s->up.arr[MGB_U][q] + (long long)colour * gn.cs + (long long)(gp - gn.i0) * gn.pj;
        }
        p.peer_flag = s->up.flags + XF_FROM_LOW;
        p.count = (unsigned int *)(s->xflags + XF_CNT_UP);
        p.off = s->n_up;
    }
    if (s->rank > 0) {
        const Geo &gn = s->low.g[q];
        HaloPush &p = h.push_low;
        p.plane[0] = lv.own_lo - g.i0;
        p.plane[1] = -1;
        p.dst[0] = s->low.arr[MGB_U][q] + (long long)colour * gn.cs + (long long)(lv.own_lo - gn.i0) * gn.pj;
        p.dst[1] = nullptr;
        p.peer_flag = s->low.flags + XF_FROM_UP;
        p.count = (unsigned int *)(s->xflags + XF_CNT_LOW);
        p.off = s->n_low;
    }
}

// the fused residual+restriction of level q into partitioned level q-1: my last coarse
// plane is the upper neighbour's coarse-rhs halo (it feeds the NEXT restriction)
static void add_coarse_rhs_push(mgb_solver *s, int qc, HaloCtl &h)
{
    Level &c = s->lv[qc];
    ++s->n_up;
    if (s->rank < s->nranks - 1) {
        const Geo &gn = s->up.g[qc];
        HaloPush &p = h.push_up;
        const int gp = c.own_hi - 1;
        p.plane[0] = gp - c.g.i0;
        p.plane[1] = -1;
        for (int col = 0; col < 2; col++)
            p.dst[col] = s->up.arr[MGB_D][qc] + (long long)col * gn.cs + (long long)(gp - gn.i0) * gn.pj;
        p.peer_flag = s->up.flags + XF_FROM_LOW;
        p.count = (unsigned int *)(s->xflags + XF_CNT_UP);
        p.off = s->n_up;
    }
}

// end of a collective operation on the P2P path: remember the last sends, open the next
// epoch (one tiny kernel; inside a captured cycle it is the graph's last node)
static void close_epoch(mgb_solver *s)
{
    if (!s->is_dist() || !s->opt_p2p)
        return;
    launch_epoch_close(s->xflags, s->n_up, s->n_low, s->st);
    s->n_up = s->n_low = s->n_gather = s->n_norm = 0;
}
struct Collective {
    mgb_solver *s;
    explicit Collective(mgb_solver *s_) : s(s_) {}
    ~Collective() { close_epoch(s); }
};

// sum a device scalar over the ranks: P2P path in RANK ORDER on every rank (bitwise the
// same everywhere and from run to run), NCCL path in NCCL's order
static void allreduce_scalar(mgb_solver *s, int slot)
{
    if (!s->is_dist())
        return;
    if (s->opt_p2p) {
        NormHost g{};
        g.me = s->rank;
        g.nranks = s->nranks;
        g.scalar = s->d_scal + slot;
        for (int r = 0; r < s->nranks; r++)
            g.peer_xf[r] = s->peers[r].xf;
        g.my_xf = s->xflags;
        g.off = ++s->n_norm;
        g.timeout_ns = s->halo_timeout_ns;
        g.err = s->d_halo_err;
        launch_norm_exchange(g, s->st);
        s->nccl_calls++;
        return;
    }
    comm_begin(s);
    NC(nccl().AllReduce(s->d_scal + slot, s->d_scal + slot, 1, ncclDouble, ncclSum, s->comm,
                        nccl_stream(s)));
    comm_end(s);
    s->nccl_calls++;
}

// one EXPLICIT halo step on array `a` of a partitioned level, colours in `mask`
// (bit c = colour c): uploads, fences, stand-alone calls -- the cycle's own exchanges
// ride inside its kernels.  Planes are global indices, -1 = nothing:
//   send_up / send_up2 -> upper neighbour stores them as the same global planes
//   send_down          -> lower neighbour
// `dir`: bit 0 = something goes up, bit 1 = something goes down (the same on every rank,
// whether or not it has that neighbour: sends are counted identically everywhere).
// NCCL path: one grouped send/recv per step, ordered in the receiver's stream.
static void halo_step(mgb_solver *s, Level &lv, double *a, int mask, int send_up, int send_up2,
                      int send_down, int dir)
{
    const Geo &g = lv.g;
    const bool has_low = s->rank > 0, has_up = s->rank < s->nranks - 1;
    const int q = (int)(&lv - &s->lv[0]);
    const int w = a == lv.a[MGB_U].base ? MGB_U : MGB_D;
    if (s->opt_p2p) {
        HaloRun run[2];  // 0: to the upper neighbour, 1: to the lower
        if (dir & 1)
            ++s->n_up;
        if (dir & 2)
            ++s->n_low;
        for (int d = 0; d < 2; d++) {
            if (!(dir & (1 << d)) || !(d == 0 ? has_up : has_low))
                continue;
            const PeerSide &ps = d == 0 ? s->up : s->low;
            const Geo &gn = ps.g[q];
            const int planes[2] = {d == 0 ? send_up : send_down, d == 0 ? send_up2 : -1};
            int k = 0;
            for (int pi = 0; pi < 2; pi++) {
                if (planes[pi] < 0)
                    continue;
                for (int c = 0; c < 2; c++) {
                    if (!(mask & (1 << c)))
                        continue;
                    run[d].src[k] = a + (long long)c * g.cs + (long long)(planes[pi] - g.i0) * g.pj;
                    run[d].dst[k] = ps.arr[w][q] + (long long)c * gn.cs +
                                    (long long)(planes[pi] - gn.i0) * gn.pj;
                    run[d].n[k] = g.pj;
                    k++;
                }
            }
            run[d].peer_flag = ps.flags + (d == 0 ? XF_FROM_LOW : XF_FROM_UP);
            run[d].count = (unsigned int *)(s->xflags + (d == 0 ? XF_CNT_UP : XF_CNT_LOW));
            run[d].off = d == 0 ? s->n_up : s->n_low;
        }
        launch_halo_push(run[0], run[1], s->xflags + XF_EPOCH, s->st);
        launch_halo_wait(make_ctl(s, lv), s->st);
        s->nccl_calls++;
        return;
    }
    comm_begin(s);
    NC(nccl().GroupStart());
    for (int c = 0; c < 2; c++) {
        if (!(mask & (1 << c)))
            continue;
        double *base = a + (long long)c * g.cs;
        const size_t n = (size_t)g.pj;
        const int ups[2] = {send_up, send_up2};
        for (int pi = 0; pi < 2; pi++) {
            if (!(dir & 1) || ups[pi] < 0)
                continue;
            if (has_up)
                NC(nccl().Send(base + (long long)(ups[pi] - g.i0) * g.pj, n, ncclDouble, s->rank + 1,
                               s->comm, s->st_comm));
            if (has_low)  // the lower neighbour's same step: its plane at the same distance
                NC(nccl().Recv(base + (long long)(lv.own_lo - (lv.own_hi - ups[pi]) - g.i0) * g.pj, n,
                               ncclDouble, s->rank - 1, s->comm, s->st_comm));
        }
        if ((dir & 2) && send_down >= 0) {
            if (has_low)
                NC(nccl().Send(base + (long long)(send_down - g.i0) * g.pj, n, ncclDouble,
                               s->rank - 1, s->comm, s->st_comm));
            if (has_up)
                NC(nccl().Recv(base + (long long)(lv.own_hi + (send_down - lv.own_lo) - g.i0) * g.pj, n,
                               ncclDouble, s->rank + 1, s->comm, s->st_comm));
        }
    }
    NC(nccl().GroupEnd());
    comm_end(s);
    s->nccl_calls++;
}

// P2P path only.  A push lands in the neighbour's halo planes whenever the SENDER gets
// there; ncclRecv, by contrast, is ordered in the receiver's stream.  Wherever a rank
// writes its own halo planes with a local kernel (uploads, mgb_zero, the full-colour
// prolongation), it therefore tells both neighbours when that is done, and waits for the
// same from them, before anybody pushes into those planes again: a data-less halo step.
static void halo_fence(mgb_solver *s, Level &lv)
{
    if (!lv.dist || !s->opt_p2p)
        return;
    halo_step(s, lv, lv.a[MGB_U].base, 0, -1, -1, -1, 3);
}

// NCCL path / stand-alone calls: after a half-sweep of `colour` the freshly written
// boundary planes go to the neighbours' halos as an explicit step
static void halo_after_sweep(mgb_solver *s, Level &lv, int colour)
{
    if (!lv.dist)
        return;
    halo_step(s, lv, lv.a[MGB_U].base, 1 << colour, lv.own_hi - 1, lv.own_hi - 2, lv.own_lo, 3);
}

static int fetch_scalar(mgb_solver *s, int slot, double *out)
{
    CK(cudaMemcpyAsync(s->h_scal + slot, s->d_scal + slot, sizeof(double),
                       cudaMemcpyDeviceToHost, s->st));
    CK(cudaStreamSynchronize(s->st));
    *out = s->h_scal[slot];
    return halo_status(s);
}

extern "C" int mgb_edge_values(mgb_solver *s, int level, int which)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    Level &lv = s->lv[level];
    if (lv.dist)
        return fail("mgb_edge_values: not available on a partitioned level");
    if (level < s->L - 1)
        s->coarse_dirty = true;
    LaunchScope ls(s);
    launch_edge_values(lv.g, lv.a[which].base, s->st);
    CKLAUNCH();
    return 0;
}

extern "C" int mgb_set_spacing(mgb_solver *s, double h)
{
    if (bind(s))
        return 1;
    if (s->L != 1)
        return fail("mgb_set_spacing: only for single-grid sessions (levels == 1): the coarse "
                    "operator of a hierarchy is built from the spacing at mgb_create");
    if (!(h > 0.))
        return fail("spacing must be positive");
    Level &lv = s->lv[0];
    lv.h = h;
    lv.hSq = h * h;            // mg_3d.h:644
    lv.invHsq = 1. / (h * h);  // mg_3d.h:797
    drop_graph(s);
    return 0;
}

extern "C" int mgb_pin_host(void *p, unsigned long long bytes)
{
    if (!p || !bytes)
        return fail("mgb_pin_host: null range");
    int ndev = 0;
    if (mgb_device_count(&ndev) || ndev < 1)
        return fail("no CUDA device");
    cudaError_t e = cudaHostRegister(p, (size_t)bytes, cudaHostRegisterPortable);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail("cudaHostRegister(%p, %llu): %s", p, bytes, cudaGetErrorString(e));
    }
    return 0;
}

extern "C" int mgb_unpin_host(void *p)
{
    if (!p)
        return 0;
    cudaError_t e = cudaHostUnregister(p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return fail("cudaHostUnregister(%p): %s", p, cudaGetErrorString(e));
    }
    return 0;
}

extern "C" int mgb_sumsq(mgb_solver *s, int level, int which, double *sumsq)
{
    if (bind(s) || check_level(s, level, which))
        return 1;
    Level &lv = s->lv[level];
    LaunchScope ls(s);
    // owned planes only (halos belong to the neighbours)
    const long long first = (long long)(lv.own_lo - lv.g.i0) * lv.g.pj;
    const long long n = (long long)(lv.own_hi - lv.own_lo) * lv.g.pj;
    if (!s->works_on(level))
        CK(cudaMemsetAsync(s->d_scal + 1, 0, sizeof(double), s->st));
    else
        launch_sumsq(lv.a[which].base + first, n, lv.a[which].base + lv.g.cs + first, n,
                     s->partials, s->d_scal + 1, s->st);
    CKLAUNCH();
    if (lv.dist) {
        allreduce_scalar(s, 1);
        close_epoch(s);
    }
    if (nccl_status())
        return 1;
    return fetch_scalar(s, 1, sumsq);
}

extern "C" int mgb_error_sumsq(mgb_solver *s, double *sumsq)
{
    if (bind(s))
        return 1;
    Level &lv = s->lv[s->L - 1];
    LaunchScope ls(s);
    launch_error_sumsq(lv.g, lv.a[MGB_U].base, lv.h, lv.own_lo - lv.g.i0, lv.own_hi - lv.g.i0,
                       s->partials, s->d_scal + 2, s->st);
    CKLAUNCH();
    if (lv.dist) {
        allreduce_scalar(s, 2);
        close_epoch(s);
    }
    if (nccl_status())
        return 1;
    return fetch_scalar(s, 2, sumsq);
}

// ----------------------------------------------------------------------------
// operators (enqueue only; the caller of the C entry point synchronises)
// ----------------------------------------------------------------------------
static void q_half_sweep(mgb_solver *s, int q, int colour, bool zero_guess = false)
{
    if (!s->works_on(q))
        return;
    Level &lv = s->lv[q];
    const int lo = lv.sweep_lo(), hi = lv.sweep_hi();
    // partitioned level, P2P path: the kernel itself waits for the neighbours' last sends
    // and stores its boundary planes into their halos (halo.cuh)
    HaloCtl h = make_ctl(s, lv);
    if (h.epoch)
        add_sweep_push(s, q, colour, h);
    const HaloCtl *hp = h.epoch ? &h : nullptr;
    if (zero_guess)
        launch_first_sweep_zero(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base, lv.hSq, colour, lo, hi,
                                s->st, hp);
    else
        launch_half_sweep(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base, lv.hSq, colour, lo, hi, s->st,
                          hp);
    if (lv.dist && !s->opt_p2p)
        halo_after_sweep(s, lv, colour);  // NCCL path: an explicit exchange
}

static void q_smooth(mgb_solver *s, int q, int iters, int first_red, bool zero_guess = false)
{
    for (int it = 0; it < iters; it++) {
        q_half_sweep(s, q, first_red ? 1 : 0, zero_guess && it == 0);
        q_half_sweep(s, q, first_red ? 0 : 1);
    }
}

static void q_residual(mgb_solver *s, int q, bool store, int slot)
{
    Level &lv = s->lv[q];
    if (!s->works_on(q)) {
        cudaMemsetAsync(s->d_scal + slot, 0, sizeof(double), s->st);
        return;
    }
    const HaloCtl h = make_ctl(s, lv);  // reads the nearest halo plane on each side
    const HaloCtl *hp = h.epoch ? &h : nullptr;
    if (store ||
        !launch_tile_residual(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base, lv.hSq, lv.invHsq, -1,
                              lv.sweep_lo(), lv.sweep_hi(), s->partials, s->d_scal + slot, s->st,
                              hp))
        launch_residual(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base,
                        store ? lv.a[MGB_R].base : nullptr, lv.invHsq, lv.sweep_lo(),
                        lv.sweep_hi(), s->partials, s->d_scal + slot, s->st, hp);
    if (lv.dist)
        allreduce_scalar(s, slot);
}

// half-sweep of `colour` fused with the residual norm (tile.cu); falls back to
// the two plain kernels where the tile kernel does not apply
static void q_sweep_residual(mgb_solver *s, int q, int colour, int slot)
{
    Level &lv = s->lv[q];
    if (s->works_on(q) && !lv.dist &&
        launch_tile_residual(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base, lv.hSq, lv.invHsq, colour,
                             lv.sweep_lo(), lv.sweep_hi(), s->partials, s->d_scal + slot, s->st))
        return;
    q_half_sweep(s, q, colour);
    q_residual(s, q, false, slot);
}

static void q_restrict(mgb_solver *s, int q)
{
    if (!s->works_on(q))
        return;
    Level &f = s->lv[q], &c = s->lv[q - 1];
    launch_restrict(f.g, f.a[MGB_R].base, c.g, c.a[MGB_D].base, 0, c.g.li, s->st);
}

// all-gather of level qc's right-hand side (the first replicated level): every rank has
// computed the coarse planes [Ilo, Ihi) and needs everybody else's (P2P path)
static void q_gather_rhs(mgb_solver *s, int qc, int Ilo, int Ihi)
{
    Level &c = s->lv[qc];
    GatherHost g{};
    g.me = s->rank;
    g.nranks = s->nranks;
    for (int col = 0; col < 2; col++) {
        g.src[col] = c.a[MGB_D].base + (long long)col * c.g.cs + (long long)Ilo * c.g.pj;
        g.n[col] = (long long)(Ihi - Ilo) * c.g.pj;
    }
    for (int r = 0; r < s->nranks; r++) {
        if (r == s->rank)
            continue;
        for (int col = 0; col < 2; col++)  // level qc has the same (whole) geometry everywhere
            g.dst[r][col] = s->peers[r].d_coarse + (long long)col * c.g.cs + (long long)Ilo * c.g.pj;
        g.peer_xf[r] = s->peers[r].xf;
    }
    g.my_xf = s->xflags;
    g.off = ++s->n_gather;
    g.timeout_ns = s->halo_timeout_ns;
    g.err = s->d_halo_err;
    launch_gather(g, s->st);
    s->nccl_calls++;
}

// residual + restriction; colour >= 0: the half-sweep of that colour is fused in front
// (single-GPU levels; elsewhere it runs as its own kernel first).  `fresh_halos`: the
// caller knows that both colours of the two lower halo planes are current (inside the
// cycle every half-sweep pushes its colour of planes own_hi-2 and own_hi-1).
static void q_residual_restrict(mgb_solver *s, int q, int colour = -1, bool fresh_halos = false)
{
    if (!s->works_on(q))
        return;
    Level &f = s->lv[q], &c = s->lv[q - 1];
    if (!f.dist) {
        if (colour >= 0 &&
            launch_tile_residual_restrict(f.g, f.a[MGB_U].base, f.a[MGB_D].base, f.hSq, f.invHsq,
                                          colour, c.g, c.a[MGB_D].base, 0, c.g.li, s->st))
            return;
        if (colour >= 0)
            q_half_sweep(s, q, colour);
        if (!launch_tile_residual_restrict(f.g, f.a[MGB_U].base, f.a[MGB_D].base, f.hSq, f.invHsq,
                                           -1, c.g, c.a[MGB_D].base, 0, c.g.li, s->st))
            launch_residual_restrict(f.g, f.a[MGB_U].base, f.a[MGB_D].base, f.invHsq, c.g,
                                     c.a[MGB_D].base, 0, c.g.li, s->st);
        return;
    }
    if (colour >= 0)
        q_half_sweep(s, q, colour);
    // the residual on plane own_lo-1 (needed by my first coarse plane) reads the solution
    // on own_lo-2 .. own_lo and the rhs on own_lo-1
    if (!fresh_halos || !s->opt_p2p)
        halo_step(s, f, f.a[MGB_U].base, 3, f.own_hi - 1, f.own_hi - 2, f.own_lo, 3);
    // my share of the coarse planes: those whose fine plane 2I I own
    int Ilo, Ihi;
    mgb_plan_slab(c.g.ni, s->nranks, s->rank, &Ilo, &Ihi);
    HaloCtl h = make_ctl(s, f);
    if (h.epoch && c.dist)
        add_coarse_rhs_push(s, q - 1, h);
    const HaloCtl *hp = h.epoch ? &h : nullptr;
    if (!launch_tile_residual_restrict(f.g, f.a[MGB_U].base, f.a[MGB_D].base, f.hSq, f.invHsq, -1,
                                       c.g, c.a[MGB_D].base, Ilo - c.g.i0, Ihi - c.g.i0, s->st, hp))
        launch_residual_restrict(f.g, f.a[MGB_U].base, f.a[MGB_D].base, f.invHsq, c.g,
                                 c.a[MGB_D].base, Ilo - c.g.i0, Ihi - c.g.i0, s->st, hp);
    if (c.dist) {
        // the coarse rhs on plane own_lo-1 feeds the next restriction
        if (!s->opt_p2p)
            halo_step(s, c, c.a[MGB_D].base, 3, c.own_hi - 1, -1, -1, 1);
    } else if (s->replicate) {
        q_gather_rhs(s, q - 1, Ilo, Ihi);
    } else {
        // agglomeration (NCCL path): gather the coarse rhs slabs on rank 0
        comm_begin(s);
        NC(nccl().GroupStart());
        for (int col = 0; col < 2; col++) {
            double *base = c.a[MGB_D].base + (long long)col * c.g.cs;
            if (s->rank > 0) {
                NC(nccl().Send(base + (long long)Ilo * c.g.pj, (size_t)(Ihi - Ilo) * c.g.pj,
                               ncclDouble, 0, s->comm, nccl_stream(s)));
            } else {
                for (int r = 1; r < s->nranks; r++) {
                    int lo, hi;
                    mgb_plan_slab(c.g.ni, s->nranks, r, &lo, &hi);
                    NC(nccl().Recv(base + (long long)lo * c.g.pj, (size_t)(hi - lo) * c.g.pj,
                                   ncclDouble, r, s->comm, nccl_stream(s)));
                }
            }
        }
        NC(nccl().GroupEnd());
        comm_end(s);
        s->nccl_calls++;
    }
}

// red_only (inside the cycle, gs >= 1): the post-smoother starts with BLACK and
// overwrites every interior black point without reading it, so only the RED
// points need the correction; black face points still get the reference's
// `+= 0.` on the finest level (it turns a -0. boundary value into +0.; coarse
// faces are +0. already)
static void q_prolong(mgb_solver *s, int q, bool red_only = false)
{
    if (!s->works_on(q))
        return;
    Level &f = s->lv[q], &c = s->lv[q - 1];
    const int cmask = red_only ? 2 : 3;
    if (!f.dist) {
        launch_prolong_correct(c.g, c.a[MGB_U].base, f.g, f.a[MGB_U].base, 0, f.g.li, s->st,
                               cmask);
        if (red_only && q == s->L - 1)
            launch_add_zero_faces(f.g, f.a[MGB_U].base, 0, 0, f.g.li, s->st);
        return;
    }
    // owned planes plus the nearest halo plane on each side: the neighbours
    // compute the identical values, so no exchange is needed afterwards.  The coarse
    // planes read at the ends of the slab are halo planes of the coarse level (if that one
    // is partitioned): its last sweeps pushed them
    const int lo = f.own_lo - (s->rank > 0 ? 1 : 0);
    const int hi = f.own_hi + (s->rank < s->nranks - 1 ? 1 : 0);
    const HaloCtl h = make_ctl(s, c.dist ? c : f);
    launch_prolong_correct(c.g, c.a[MGB_U].base, f.g, f.a[MGB_U].base, lo - f.g.i0,
                           hi - f.g.i0, s->st, cmask, (h.epoch && c.dist) ? &h : nullptr);
    if (red_only) {
        // only red entries of my halo planes were written, and the neighbours' next
        // push into them is BLACK: no fence needed
        if (q == s->L - 1)
            launch_add_zero_faces(f.g, f.a[MGB_U].base, 0, f.own_lo - f.g.i0, f.own_hi - f.g.i0,
                                  s->st);
        return;
    }
    halo_fence(s, f);
}

// after rank 0 has finished the agglomerated levels: everybody gets the
// correction of level LD-1 (a small grid) for the prolongation
static void q_broadcast_agglomerated(mgb_solver *s)
{
    if (s->replicate)
        return;  // every rank has computed level LD-1 itself
    Level &c = s->lv[s->LD - 1];
    comm_begin(s);
    NC(nccl().Broadcast(c.a[MGB_U].base, c.a[MGB_U].base, (size_t)(2 * c.g.cs), ncclDouble, 0,
                        s->comm, nccl_stream(s)));
    comm_end(s);
    s->nccl_calls++;
}

static void q_coarse_solve(mgb_solver *s)
{
    // solveWithLU(LU, n, d[0], u[0]) (mg_3d.h:1270): the dense vectors are the
    // natural-layout views of level 0
    if (!s->works_on(0))
        return;
    Level &lv = s->lv[0];
    launch_lu_solve_level(s->band, lv.g, lv.a[MGB_D].base, lv.a[MGB_U].base, s->st);
}

#define OP_PROLOGUE(level_expr, min_level)                                        \
    if (bind(s))                                                                  \
        return 1;                                                                 \
    if ((level_expr) < (min_level) || (level_expr) >= s->L)                       \
        return fail("level %d out of range [%d,%d)", (level_expr), (min_level), s->L); \
    LaunchScope ls(s);                                                            \
    Collective epoch_scope(s)

extern "C" int mgb_half_sweep(mgb_solver *s, int level, int colour)
{
    OP_PROLOGUE(level, 0);
    q_half_sweep(s, level, colour ? 1 : 0);
    CKLAUNCH();
    return nccl_status();
}

// tuning aid (not part of the reference's API): one colour over the local planes
// [il_lo, il_hi) only -- what a rank of a partitioned solver runs on its slab
extern "C" int mgb_debug_half_sweep_range(mgb_solver *s, int level, int colour, int il_lo,
                                          int il_hi)
{
    OP_PROLOGUE(level, 0);
    Level &lv = s->lv[level];
    if (lv.dist || il_lo < 1 || il_hi > lv.g.li - 1 || il_lo >= il_hi)
        return fail("mgb_debug_half_sweep_range: bad range or partitioned level");
    launch_half_sweep(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base, lv.hSq, colour ? 1 : 0, il_lo, il_hi,
                      s->st);
    CKLAUNCH();
    return 0;
}

extern "C" int mgb_smooth(mgb_solver *s, int level, int iters, int first_red)
{
    OP_PROLOGUE(level, 0);
    q_smooth(s, level, iters, first_red);
    CKLAUNCH();
    return nccl_status();
}

// GaussSeidelSmoother's sweeps (mg_3d.h:559-634): `iters` LEXICOGRAPHIC Gauss-Seidel
// sweeps as a hyperplane wavefront (gslex.cu), bit-identical to the serial triple loop
extern "C" int mgb_gs_lex(mgb_solver *s, int level, int iters)
{
    OP_PROLOGUE(level, 0);
    Level &lv = s->lv[level];
    if (lv.dist)
        return fail("mgb_gs_lex: not available on a partitioned level");
    if (iters < 0)
        return fail("mgb_gs_lex: iters >= 0");
    if (level < s->L - 1)
        s->coarse_dirty = true;
    // d_scal[15] doubles as the wavefront's arrival counter
    if (launch_gs_lex(lv.g, lv.a[MGB_U].base, lv.a[MGB_D].base, lv.hSq, iters,
                      reinterpret_cast<unsigned int *>(s->d_scal + 15), s->st))
        return fail("mgb_gs_lex: cooperative launch failed: %s",
                    cudaGetErrorString(cudaGetLastError()));
    CKLAUNCH();
    return 0;
}

extern "C" int mgb_residual(mgb_solver *s, int level, int store_r, double *sumsq)
{
    OP_PROLOGUE(level, 0);
    if (store_r && s->is_dist())
        return fail("the partitioned solver does not store the fine residual");
    q_residual(s, level, store_r != 0, 0);
    CKLAUNCH();
    if (nccl_status())
        return 1;
    if (sumsq)
        return fetch_scalar(s, 0, sumsq);
    return 0;
}

extern "C" int mgb_restrict(mgb_solver *s, int level)
{
    OP_PROLOGUE(level, 1);
    s->coarse_dirty = true;  // injects whatever sits on r's faces
    if (s->is_dist())
        return fail("the partitioned solver only has the fused residual+restriction");
    q_restrict(s, level);
    CKLAUNCH();
    return 0;
}

extern "C" int mgb_residual_restrict(mgb_solver *s, int level)
{
    OP_PROLOGUE(level, 1);
    q_residual_restrict(s, level);
    CKLAUNCH();
    return nccl_status();
}

extern "C" int mgb_sweep_residual_restrict(mgb_solver *s, int level, int colour)
{
    OP_PROLOGUE(level, 1);
    q_residual_restrict(s, level, colour ? 1 : 0);
    CKLAUNCH();
    return nccl_status();
}

extern "C" int mgb_sweep_residual(mgb_solver *s, int level, int colour, double *sumsq)
{
    OP_PROLOGUE(level, 0);
    q_sweep_residual(s, level, colour ? 1 : 0, 0);
    CKLAUNCH();
    if (nccl_status())
        return 1;
    if (sumsq)
        return fetch_scalar(s, 0, sumsq);
    return 0;
}

extern "C" int mgb_prolong_correct(mgb_solver *s, int level)
{
    OP_PROLOGUE(level, 1);
    if (s->is_dist() && level == s->LD)
        q_broadcast_agglomerated(s);  // level LD-1 lives on rank 0
    q_prolong(s, level);
    CKLAUNCH();
    return nccl_status();
}

extern "C" int mgb_coarse_solve(mgb_solver *s)
{
    if (bind(s))
        return 1;
    if (!s->lu)
        return fail("no coarse operator (single-grid session)");
    s->coarse_dirty = true;  // u[0]'s faces become d[0]'s
    LaunchScope ls(s);
    q_coarse_solve(s);
    CKLAUNCH();
    return 0;
}

extern "C" int mgb_coarse_info(const mgb_solver *s, int *n, int *half_bandwidth,
                               double *factor_seconds)
{
    if (!s)
        return fail("null solver");
    if (n) *n = s->nc;
    if (half_bandwidth) *half_bandwidth = s->band.bw;
    if (factor_seconds) *factor_seconds = s->factor_secs;
    return 0;
}

extern "C" int mgb_coarse_lu_download(mgb_solver *s, double *host_lu)
{
    if (bind(s))
        return 1;
    if (!s->lu)
        return fail("no coarse operator (single-grid session)");
    CK(cudaStreamSynchronize(s->st));
    CK(cudaMemcpy(host_lu, s->lu, sizeof(double) * (size_t)s->nc * s->nc,
                  cudaMemcpyDeviceToHost));
    return 0;
}

// ----------------------------------------------------------------------------
// the V-cycle (mg_3d.h:1242-1362)
// ----------------------------------------------------------------------------
static cudaEvent_t take_event(mgb_solver *s)
{
    if (s->ev_used == s->ev_pool.size()) {
        cudaEvent_t e;
        cudaEventCreate(&e);
        s->ev_pool.push_back(e);
    }
    return s->ev_pool[s->ev_used++];
}

struct StageTimer {
    mgb_solver *s;
    bool on;
    int level, stage;
    cudaEvent_t a = nullptr;
    StageTimer(mgb_solver *s_, bool on_, int level_, int stage_)
        : s(s_), on(on_), level(level_), stage(stage_)
    {
        s->calls[(size_t)level * MGB_NUM_STAGES + stage]++;
        if (on) {
            a = take_event(s);
            cudaEventRecord(a, s->st);
        }
    }
    ~StageTimer()
    {
        if (on) {
            cudaEvent_t b = take_event(s);
            cudaEventRecord(b, s->st);
            s->marks.push_back({level, stage, a, b});
        }
    }
};

// enqueue one cycle from level q down and back up; `count` = bump call counts
// and (if timed) record stage events.  The residual norm of the finest level
// lands in d_scal[0].
// `entry` = the level the cycle was entered at (the finest one for a V-cycle;
// mgb_fmg_init enters at every level in turn): a coarse entry level holds real
// data (boundary values on its faces), so it is zeroed for real like
// mg_3d.h:1258-1259 does, not through the zero-guess short-cut.
static void enqueue_cycle(mgb_solver *s, int q, bool timed, int entry = -1)
{
    Level &lv = s->lv[q];
    if (!s->works_on(q))
        return;  // agglomerated levels run on rank 0 only
    if (q == s->tail_top && s->opt_tail && !timed && s->lu) {
        // levels q .. 0 .. q in one single-block kernel (tail.cu); the stage timers
        // of MGB_OPT_PROFILE keep the per-stage kernels
        TailP p{};
        p.top = q;
        p.gs = s->gs;
        p.zero_top = 1;
        p.lu = s->band;
        for (int l = 0; l <= q; l++) {
            Level &t = s->lv[l];
            p.lv[l] = TailLevel{t.g, t.a[MGB_U].base, t.a[MGB_D].base, t.a[MGB_R].base, t.hSq,
                                t.invHsq};
        }
        launch_coarse_tail(p, s->st);
        return;
    }
    // 1254-1260: coarse levels start from a zero guess.  With at least one
    // smoothing iteration the level is not zeroed: its first half-sweep treats
    // the guess as 0 and the second one overwrites the other colour
    const bool zero_guess = q > 0 && q < s->L - 1 && s->opt_zero_guess && s->gs >= 1 && q != entry;
    if (q < s->L - 1 && !zero_guess && q > 0) {
        cudaMemsetAsync(lv.a[MGB_U].base, 0, sizeof(double) * 2 * lv.g.cs, s->st);
        halo_fence(s, lv);
    }
    if (q == 0) {  // 1262-1277
        StageTimer t(s, timed, 0, MGB_ST_RECURSE);
        q_coarse_solve(s);
        return;
    }
    // opt_fuse >= 2: the last colour of each smoother leg rides inside the
    // residual kernel that follows it (tile.cu); same arithmetic, same bits
    const bool fuse_sweep = s->opt_fuse >= 2 && s->gs >= 1;
    {
        StageTimer t(s, timed, q, MGB_ST_SMOOTH1);  // 1282
        if (fuse_sweep) {
            q_smooth(s, q, s->gs - 1, 1, zero_guess);
            q_half_sweep(s, q, 1, zero_guess && s->gs == 1);
        } else {
            q_smooth(s, q, s->gs, 1, zero_guess);
        }
    }
    if (s->opt_fuse) {
        StageTimer t(s, timed, q, MGB_ST_RESID1);  // 1294 + 1310 in one pass
        q_residual_restrict(s, q, fuse_sweep ? 0 : -1, s->gs >= 1);
        s->calls[(size_t)q * MGB_NUM_STAGES + MGB_ST_RESTRICT]++;
    } else {
        {
            StageTimer t(s, timed, q, MGB_ST_RESID1);  // 1294
            q_residual(s, q, true, 3);
        }
        {
            StageTimer t(s, timed, q, MGB_ST_RESTRICT);  // 1310
            q_restrict(s, q);
        }
    }
    {
        StageTimer t(s, timed, q, MGB_ST_RECURSE);  // 1320
        enqueue_cycle(s, q - 1, timed, entry);
        if (s->is_dist() && q == s->LD)
            q_broadcast_agglomerated(s);
    }
    {
        StageTimer t(s, timed, q, MGB_ST_PROLONG);  // 1331
        q_prolong(s, q, s->opt_prolong_mask && s->gs >= 1);
    }
    const bool fuse_post = fuse_sweep && q == s->L - 1;
    {
        StageTimer t(s, timed, q, MGB_ST_SMOOTH2);  // 1341
        if (fuse_post) {
            q_smooth(s, q, s->gs - 1, 0);
            q_half_sweep(s, q, 0);
        } else {
            q_smooth(s, q, s->gs, 0);
        }
    }
    {
        // 1354: the reference evaluates the norm on every level but only the
        // finest one is ever used (test_mg_3d.c:45); coarser ones are elided
        StageTimer t(s, timed && q == s->L - 1, q, MGB_ST_RESID2);
        if (fuse_post)
            q_sweep_residual(s, q, 1, 0);
        else if (q == s->L - 1)
            q_residual(s, q, false, 0);
    }
}

static int build_graph(mgb_solver *s)
{
    cudaGraph_t graph = nullptr;
    std::vector<int> saved = s->calls;
    const long long before = launches_issued();
    CK(cudaStreamBeginCapture(s->st, cudaStreamCaptureModeThreadLocal));
    enqueue_cycle(s, s->L - 1, false);
    close_epoch(s);  // the graph's last node: every replay is one epoch of the halo protocol
    cudaError_t e = cudaStreamEndCapture(s->st, &graph);
    s->calls = saved;  // capture is not a cycle
    s->graph_launches = launches_issued() - before;
    if (e != cudaSuccess)
        return fail("graph capture failed: %s", cudaGetErrorString(e));
    e = cudaGraphInstantiate(&s->gexec, graph, 0);
    cudaGraphDestroy(graph);
    if (e != cudaSuccess) {
        s->gexec = nullptr;
        return fail("graph instantiate failed: %s", cudaGetErrorString(e));
    }
    return 0;
}

static void bump_calls_like_cycle(mgb_solver *s)
{
    for (int q = s->L - 1; q >= 1; q--) {
        int *c = &s->calls[(size_t)q * MGB_NUM_STAGES];
        c[MGB_ST_SMOOTH1]++; c[MGB_ST_RESID1]++; c[MGB_ST_RESTRICT]++;
        c[MGB_ST_RECURSE]++; c[MGB_ST_PROLONG]++; c[MGB_ST_SMOOTH2]++;
        c[MGB_ST_RESID2]++;
    }
    s->calls[MGB_ST_RECURSE]++;
}

// one cycle on the stream; `fetch`: read the norm back (synchronises)
static int vcycle_impl(mgb_solver *s, double *sumsq, bool fetch)
{
    if (bind(s))
        return 1;
    if (s->L > 1 && !s->lu)
        return fail("no coarse operator");
    if (s->L < 2)
        return fail("a V-cycle needs at least 2 levels");
    if (s->coarse_dirty) {
        // direct array / operator calls may have left non-zero values on the faces of
        // coarse u arrays; the cycle (which no longer zeroes them every time) needs 0
        Collective scope(s);
        for (int q = 0; q < s->L - 1; q++) {
            Level &lv = s->lv[q];
            if (!s->works_on(q))
                continue;
            CK(cudaMemsetAsync(lv.a[MGB_U].base, 0, sizeof(double) * 2 * lv.g.cs, s->st));
            halo_fence(s, lv);
        }
        s->coarse_dirty = false;
    }
    if (s->opt_profile) {
        // eager launches bracketed by CUDA events per level and stage
        const long long before = launches_issued();
        s->ev_used = 0;
        s->marks.clear();
        enqueue_cycle(s, s->L - 1, true);
        close_epoch(s);
        s->launches += launches_issued() - before;
        CKLAUNCH();
        if (nccl_status())
            return 1;
        double v;
        if (fetch_scalar(s, 0, &v))
            return 1;
        for (auto &m : s->marks) {
            float ms = 0.f;
            CK(cudaEventElapsedTime(&ms, m.a, m.b));
            s->secs[(size_t)m.level * MGB_NUM_STAGES + m.stage] += 1e-3 * ms;
        }
        if (sumsq)
            *sumsq = v;
        return 0;
    }
    // a partitioned solver runs its first cycle eagerly: NCCL sets up its
    // connections on first use, which must not happen inside a capture
    const bool warm_up = s->is_dist() && s->opt_graph && !s->gexec && s->eager_cycles++ < 1;
    if (s->opt_graph && !warm_up) {
        if (!s->gexec && build_graph(s))
            return 1;
        CK(cudaGraphLaunch(s->gexec, s->st));
        s->launches += s->graph_launches;
    } else {
        const long long before = launches_issued();
        std::vector<int> saved = s->calls;
        enqueue_cycle(s, s->L - 1, false);
        close_epoch(s);
        s->calls = saved;
        s->launches += launches_issued() - before;
        CKLAUNCH();
        if (nccl_status())
            return 1;
    }
    bump_calls_like_cycle(s);
    if (!fetch)
        return 0;
    double v;
    if (fetch_scalar(s, 0, &v))
        return 1;
    if (sumsq)
        *sumsq = v;
    return 0;
}

extern "C" int mgb_vcycle(mgb_solver *s, double *sumsq) { return vcycle_impl(s, sumsq, true); }

// n cycles back to back on the stream, the host reads only the last norm: a fixed number
// of cycles (a preconditioner application, or timing the device alone) does not pay the
// per-cycle round trip of the reference's `while (norm > cmpNorm)` loop
extern "C" int mgb_vcycles(mgb_solver *s, int n, double *sumsq)
{
    if (n < 1)
        return fail("mgb_vcycles: n >= 1");
    for (int c = 0; c < n; c++)
        if (vcycle_impl(s, sumsq, c == n - 1 || s->opt_profile))
            return 1;
    return 0;
}

// SolverFMGInitialize (mg_3d.h:1364-1404, commented out upstream; live against an
// older vcycle in mg_dirichlet_analytic.c:771-806), statement by statement on the
// device: boundary values into u[0], the LU solve of d[0] into u[0] (which
// overwrites them), then per level l = 1 .. L-1: u[l] += P u[l-1] (the full
// prolongation, all points), boundary values onto the faces of u[l], u[l-1] = 0,
// one V-cycle entered at level l (which zeroes u[l] again unless l is the finest
// level, 1254-1260).  Eager launches (a set-up step, not the steady state).
extern "C" int mgb_fmg_init(mgb_solver *s, double *sumsq)
{
    if (bind(s))
        return 1;
    if (s->L < 2 || !s->lu)
        return fail("FMG needs a hierarchy of at least 2 levels");
    LaunchScope ls(s);
    std::vector<int> saved = s->calls;
    {
        Level &l0 = s->lv[0];
        if (s->works_on(0)) {
            launch_set_dirichlet(l0.g, l0.a[MGB_U].base, l0.h, s->st);
            q_coarse_solve(s);
        }
    }
    for (int l = 1; l < s->L; l++) {
        Level &f = s->lv[l], &c = s->lv[l - 1];
        if (s->is_dist() && l == s->LD)
            q_broadcast_agglomerated(s);  // u[l-1] lives on rank 0
        q_prolong(s, l, false);
        if (s->works_on(l)) {
            launch_set_dirichlet(f.g, f.a[MGB_U].base, f.h, s->st);
            halo_fence(s, f);
        }
        if (s->works_on(l - 1)) {
            CK(cudaMemsetAsync(c.a[MGB_U].base, 0, sizeof(double) * 2 * c.g.cs, s->st));
            halo_fence(s, c);
        }
        enqueue_cycle(s, l, false, l);
    }
    s->calls = saved;
    s->coarse_dirty = false;  // every coarse u was zeroed on the way up; the cycles keep faces 0
    close_epoch(s);
    CKLAUNCH();
    if (nccl_status())
        return 1;
    double v;
    if (fetch_scalar(s, 0, &v))
        return 1;
    if (sumsq)
        *sumsq = v;
    return 0;
}

extern "C" int mgb_solve(mgb_solver *s, double threshold, int max_cycles,
                         double *history, int *cycles)
{
    // test_mg_3d.c:40-66
    double norm = 1e9;
    int c = 0;
    while (norm > threshold && c < max_cycles) {
        double ss;
        if (mgb_vcycle(s, &ss))
            return 1;
        norm = sqrt(ss);
        if (history)
            history[c] = norm;
        c++;
    }
    if (cycles)
        *cycles = c;
    return 0;
}

extern "C" int mgb_timing(mgb_solver *s, int level, int stage, int *calls, double *seconds)
{
    if (!s)
        return fail("null solver");
    if (level < 0 || level >= s->L || stage < 0 || stage >= MGB_NUM_STAGES)
        return fail("level/stage out of range");
    if (calls)
        *calls = s->calls[(size_t)level * MGB_NUM_STAGES + stage];
    if (seconds)
        *seconds = s->secs[(size_t)level * MGB_NUM_STAGES + stage];
    return 0;
}

extern "C" int mgb_timing_reset(mgb_solver *s)
{
    if (!s)
        return fail("null solver");
    std::fill(s->secs.begin(), s->secs.end(), 0.);
    std::fill(s->calls.begin(), s->calls.end(), 0);
    return 0;
}

extern "C" long long mgb_launch_count(const mgb_solver *s) { return s ? s->launches : 0; }

extern "C" int mgb_timer_start(mgb_solver *s)
{
    if (bind(s))
        return 1;
    if (!s->sw_a) {
        CK(cudaEventCreate(&s->sw_a));
        CK(cudaEventCreate(&s->sw_b));
    }
    CK(cudaEventRecord(s->sw_a, s->st));
    return 0;
}

extern "C" int mgb_timer_stop(mgb_solver *s, double *seconds)
{
    if (bind(s))
        return 1;
    if (!s->sw_a)
        return fail("mgb_timer_stop without mgb_timer_start");
    CK(cudaEventRecord(s->sw_b, s->st));
    CK(cudaEventSynchronize(s->sw_b));
    float ms = 0.f;
    CK(cudaEventElapsedTime(&ms, s->sw_a, s->sw_b));
    if (seconds)
        *seconds = 1e-3 * (double)ms;
    return 0;
}

extern "C" void *mgb_stream(mgb_solver *s) { return s ? (void *)s->st : nullptr; }

// ----------------------------------------------------------------------------
// stateless host-array entry points (raw-pointer API of the reference)
// ----------------------------------------------------------------------------
namespace {

struct Tmp {
    Geo g;
    DevArray a;
    ~Tmp() { dev_free(&a); }
};

struct Scratch {
    double *p = nullptr;
    ~Scratch()
    {
        if (p)
            cudaFree(p);
    }
};

int host_ready()
{
    int n = 0;
    if (mgb_device_count(&n))
        return 1;
    if (n < 1)
        return fail("no CUDA device: libmgb has no CPU fallback");
    return 0;
}

int tmp_make(Tmp *t, int ni, int nj, int nk)
{
    if (ni < 3 || nj < 3 || nk < 3)
        return fail("extents must be >= 3");
    t->g = make_geo(ni, nj, nk, ni, 0);
    return dev_alloc(t->g, &t->a);
}

int tmp_upload(Tmp *t, const double *host, double *stage)
{
    const size_t n = (size_t)t->g.ni * t->g.nj * t->g.nk;
    CK(cudaMemcpy(stage, host, sizeof(double) * n, cudaMemcpyHostToDevice));
    launch_pack(t->g, stage, t->a.base, 0);
    CKLAUNCH();
    CK(cudaDeviceSynchronize());
    return 0;
}

int tmp_download(Tmp *t, double *host, double *stage)
{
    const size_t n = (size_t)t->g.ni * t->g.nj * t->g.nk;
    launch_unpack(t->g, t->a.base, stage, 0);
    CKLAUNCH();
    CK(cudaMemcpy(host, stage, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}

}  // namespace

extern "C" int mgb_host_smooth(double *v, const double *d, int ni, int nj, int nk,
                               double h, int iters, int first_red)
{
    if (host_ready())
        return 1;
    Tmp tv, td;
    Scratch st;
    if (tmp_make(&tv, ni, nj, nk) || tmp_make(&td, ni, nj, nk))
        return 1;
    CK(cudaMalloc(&st.p, sizeof(double) * (size_t)ni * nj * nk));
    if (tmp_upload(&tv, v, st.p) || tmp_upload(&td, d, st.p))
        return 1;
    const double hSq = h * h;
    for (int it = 0; it < iters; it++) {
        launch_half_sweep(tv.g, tv.a.base, td.a.base, hSq, first_red ? 1 : 0, 1, ni - 1, 0);
        launch_half_sweep(tv.g, tv.a.base, td.a.base, hSq, first_red ? 0 : 1, 1, ni - 1, 0);
    }
    CKLAUNCH();
    return tmp_download(&tv, v, st.p);
}

// GaussSeidelSmoother(v, d, N, h, iters) on caller-owned host arrays (test_gs_3d.c:56):
// the lexicographic sweeps and, with `edges`, the updateEdgeValues that ends the routine
extern "C" int mgb_host_gs_lex(double *v, const double *d, int ni, int nj, int nk, double h,
                               int iters, int edges)
{
    if (host_ready())
        return 1;
    Tmp tv, td;
    Scratch st, bar;
    if (tmp_make(&tv, ni, nj, nk) || tmp_make(&td, ni, nj, nk))
        return 1;
    CK(cudaMalloc(&st.p, sizeof(double) * (size_t)ni * nj * nk));
    CK(cudaMalloc(&bar.p, sizeof(double)));
    if (tmp_upload(&tv, v, st.p) || tmp_upload(&td, d, st.p))
        return 1;
    if (launch_gs_lex(tv.g, tv.a.base, td.a.base, h * h, iters,
                      reinterpret_cast<unsigned int *>(bar.p), 0))
        return fail("mgb_host_gs_lex: cooperative launch failed: %s",
                    cudaGetErrorString(cudaGetLastError()));
    if (edges)
        launch_edge_values(tv.g, tv.a.base, 0);
    CKLAUNCH();
    return tmp_download(&tv, v, st.p);
}

extern "C" int mgb_host_residual(const double *v, const double *d, int ni, int nj,
                                 int nk, double h, double *res, double *sumsq)
{
    if (host_ready())
        return 1;
    Tmp tv, td, tr;
    Scratch st, sc;
    if (tmp_make(&tv, ni, nj, nk) || tmp_make(&td, ni, nj, nk))
        return 1;
    if (res && tmp_make(&tr, ni, nj, nk))
        return 1;
    CK(cudaMalloc(&st.p, sizeof(double) * (size_t)ni * nj * nk));
    CK(cudaMalloc(&sc.p, sizeof(double) * (kMaxPartials + 1)));
    if (tmp_upload(&tv, v, st.p) || tmp_upload(&td, d, st.p))
        return 1;
    // res keeps its previous boundary values (the reference only writes the
    // interior, mg_3d.h:807-826)
    if (res && tmp_upload(&tr, res, st.p))
        return 1;
    launch_residual(tv.g, tv.a.base, td.a.base, res ? tr.a.base : nullptr, 1. / (h * h), 1,
                    ni - 1, sc.p + 1, sc.p, 0);
    CKLAUNCH();
    double ss = 0.;
    CK(cudaMemcpy(&ss, sc.p, sizeof(double), cudaMemcpyDeviceToHost));
    if (sumsq)
        *sumsq = ss;
    if (res)
        return tmp_download(&tr, res, st.p);
    return 0;
}

extern "C" int mgb_host_restrict(const double *r, int nif, int njf, int nkf, double *dc,
                                 int nic, int njc, int nkc)
{
    if (host_ready())
        return 1;
    if (nif != 2 * nic - 1 || njf != 2 * njc - 1 || nkf != 2 * nkc - 1)
        return fail("fine extents must be 2*coarse-1");
    Tmp tf, tc;
    Scratch st;
    if (tmp_make(&tf, nif, njf, nkf) || tmp_make(&tc, nic, njc, nkc))
        return 1;
    CK(cudaMalloc(&st.p, sizeof(double) * (size_t)nif * njf * nkf));
    if (tmp_upload(&tf, r, st.p))
        return 1;
    launch_restrict(tf.g, tf.a.base, tc.g, tc.a.base, 0, nic, 0);
    CKLAUNCH();
    return tmp_download(&tc, dc, st.p);
}

extern "C" int mgb_host_prolong_correct(const double *ec, int nic, int njc, int nkc,
                                        double *ef, int nif, int njf, int nkf)
{
    if (host_ready())
        return 1;
    if (nif != 2 * nic - 1 || njf != 2 * njc - 1 || nkf != 2 * nkc - 1)
        return fail("fine extents must be 2*coarse-1");
    Tmp tf, tc;
    Scratch st;
    if (tmp_make(&tf, nif, njf, nkf) || tmp_make(&tc, nic, njc, nkc))
        return 1;
    CK(cudaMalloc(&st.p, sizeof(double) * (size_t)nif * njf * nkf));
    if (tmp_upload(&tf, ef, st.p) || tmp_upload(&tc, ec, st.p))
        return 1;
    launch_prolong_correct(tc.g, tc.a.base, tf.g, tf.a.base, 0, nif, 0);
    CKLAUNCH();
    return tmp_download(&tf, ef, st.p);
}

extern "C" int mgb_host_coarse_matrix(double *A, int ni, int nj, int nk, double h)
{
    if (host_ready())
        return 1;
    const size_t n = (size_t)ni * nj * nk;
    Scratch a;
    CK(cudaMalloc(&a.p, sizeof(double) * n * n));
    launch_coarse_matrix(a.p, ni, nj, nk, h, 0);
    CKLAUNCH();
    CK(cudaMemcpy(A, a.p, sizeof(double) * n * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mgb_host_lu_factor(double *a, int n)
{
    if (host_ready())
        return 1;
    Scratch d;
    CK(cudaMalloc(&d.p, sizeof(double) * (size_t)n * n));
    CK(cudaMemcpy(d.p, a, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
    launch_lu_factor_band(d.p, n, lu_bandwidth(d.p, n, 0), 0);
    CKLAUNCH();
    CK(cudaMemcpy(a, d.p, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToHost));
    return 0;
}

extern "C" int mgb_host_lu_solve(const double *lu, int n, const double *b, double *x)
{
    if (host_ready())
        return 1;
    if (n > MGB_MAX_DENSE_N)
        return fail("n=%d exceeds MGB_MAX_DENSE_N", n);
    if (n < 1)
        return fail("n < 1");
    Scratch d, vb, vx;
    CK(cudaMalloc(&d.p, sizeof(double) * (size_t)n * n));
    CK(cudaMalloc(&vb.p, sizeof(double) * n));
    CK(cudaMalloc(&vx.p, sizeof(double) * n));
    CK(cudaMemcpy(d.p, lu, sizeof(double) * (size_t)n * n, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(vb.p, b, sizeof(double) * n, cudaMemcpyHostToDevice));
    // zeros (of either sign) outside the band of the factor contribute nothing
    struct Tiles {
        LuBand B{};
        ~Tiles() { lu_band_free(&B); }
    } tl;
    if (lu_band_alloc(&tl.B, n, lu_bandwidth(d.p, n, 0)))
        return fail("cudaMalloc of the factor's tiles failed");
    launch_lu_extract_band(d.p, tl.B, 0);
    launch_lu_solve_dense(tl.B, vb.p, vx.p, 0);
    CKLAUNCH();
    CK(cudaMemcpy(x, vx.p, sizeof(double) * n, cudaMemcpyDeviceToHost));
    return 0;
}
