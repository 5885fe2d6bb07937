// halo.cuh -- the multi-GPU exchange FUSED into the compute kernels.
//
// One process per GPU, slabs over i (the reference's `omp for schedule(static)`
// over i, mg_3d.h:658,681,729,753,807,962,1006); the barrier between two colour
// half-sweeps (the implicit barrier that ends each `omp for`) is where the
// freshly written boundary planes have to reach the neighbours.  There is no
// exchange kernel and no NCCL call on that path:
//
//   PUSH  the CTAs of a kernel that WRITE a boundary plane store every value
//         twice -- to their own array and, through the peer mapping (CUDA IPC
//         over NVLink), straight into the neighbour's halo plane.  When a CTA
//         that takes part in a direction is done it fences at system scope and
//         bumps that direction's arrival counter; the last one to arrive writes
//         the sequence number of the step into the neighbour's flag with
//         st.release.sys.  The chunks that hold boundary planes are scheduled
//         first, so the transfer runs underneath the interior of the same launch.
//   WAIT  the CTAs of the NEXT kernel whose chunk touches a halo plane spin on
//         that flag (ld.acquire.sys) before their first load.  Everything the
//         neighbour was asked to send so far is awaited (sequence numbers are
//         counted identically on every rank), so the wait needs no knowledge of
//         which kernel sent what.
//
// Sequence numbers = epoch * kHaloEpochStride + (sends so far in this epoch, per
// direction); the offset is a launch-time constant, the epoch a device-resident
// counter bumped by one tiny kernel at the end of every collective operation, so
// a captured V-cycle replays as a CUDA graph.  A wait before the first send of
// an epoch refers to the last send of the previous ones (prev_up / prev_low,
// written by the same closing kernel).
//
// Why no write can overtake a read.  A rank signals step t in a direction only
// after ALL its CTAs that read the halo planes of that side in the same kernel
// are done (they are the CTAs that push); the neighbour's step-t+1 push into
// those planes starts only after it has seen that signal.  Kernels that read
// halos without pushing (residual norm, prolongation) are followed by an
// exchange in both directions before the next write into the planes they read.
#pragma once
#include <cstdint>

namespace mgb {

constexpr unsigned long long kHaloEpochStride = 1ull << 20;  // sends per epoch and direction
constexpr int kMaxRanks = 16;

// layout of the per-rank flag block (unsigned long long slots), IPC-mapped by the peers
enum {
    XF_FROM_LOW = 0,   // written by the lower neighbour
    XF_FROM_UP = 1,    // written by the upper neighbour
    XF_EPOCH = 2,      // local: current epoch
    XF_PREV_UP = 3,    // local: sequence number of the last up-send of earlier epochs
    XF_PREV_LOW = 4,   //        ... of the last down-send
    XF_CNT_UP = 6,     // local arrival counters (32-bit each)
    XF_CNT_LOW = 7,
    XF_SCRATCH = 8,
    XF_READY = 16,     // [rank]: "rank has finished reading the gathered level"
    XF_DATA = 32,      // [rank]: "rank's slab of the gathered level has arrived here"
    XF_NORMFLAG = 48,  // [rank]: "rank's partial sum has arrived here"
    XF_NORMPART = 64,  // [2][kMaxRanks] doubles: partial sums by epoch parity
    XF_GATHER_CNT = 128,  // [rank]: local arrival counters of the gather's copy blocks
    XF_SLOTS = 160
};

struct HaloWait {
    const unsigned long long *flag;  // nullptr: no neighbour on that side
    const unsigned long long *prev;  // value to wait for while off == 0
    unsigned long long off;          // sends of the neighbour so far in this epoch
};

struct HaloPush {
    double *dst[2];                 // peer addresses (meaning depends on the kernel)
    int plane[2];                   // local plane indices the stores mirror (-1: unused)
    unsigned long long *peer_flag;  // nullptr: nothing goes this way
    unsigned int *count;            // local arrival counter
    unsigned int nblocks;           // CTAs taking part (filled in by the launch wrapper)
    unsigned long long off;         // sequence offset of this send within the epoch
};

struct HaloCtl {
    const unsigned long long *epoch;  // nullptr: single GPU, no halo work at all
    HaloWait wait_low, wait_up;
    HaloPush push_up, push_low;
    unsigned long long timeout_ns;    // 0: wait for ever
    unsigned int *err;                // host-mapped; bit 0 / 1: gave up on the lower / upper side
};

#ifdef __CUDACC__
__device__ __forceinline__ unsigned long long halo_global_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ unsigned long long halo_ld_acquire(const unsigned long long *p)
{
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void halo_st_release(unsigned long long *p, unsigned long long v)
{
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// spin until *flag >= target; gives up after timeout_ns (sets the error bit, returns)
static __device__ __noinline__ void halo_spin(const unsigned long long *flag, unsigned long long target,
                                       unsigned long long timeout_ns, unsigned int *err,
                                       unsigned int bit)
{
    const unsigned long long t0 = halo_global_ns();
    unsigned int spins = 0;
    while (halo_ld_acquire(flag) < target) {
        if (timeout_ns && (++spins & 1023u) == 0 && halo_global_ns() - t0 > timeout_ns) {
            if (err) {
                atomicOr_system(err, bit);
                __threadfence_system();
            }
            return;
        }
    }
}

// one thread: everything the neighbour on that side has been asked to send so far
__device__ __forceinline__ void halo_wait_one(const HaloCtl &h, const HaloWait &w, unsigned int bit)
{
    if (!w.flag)
        return;
    const unsigned long long target = w.off ? *h.epoch * kHaloEpochStride + w.off : *w.prev;
    if (halo_ld_acquire(w.flag) >= target)
        return;
    halo_spin(w.flag, target, h.timeout_ns, h.err, bit);
}

// Start of a kernel, called by ALL threads of every CTA before the first halo read:
// the first chunk of the plane range waits for the lower side, the last one for the upper.
// The fence lets the TMA unit (async proxy) see what the generic-proxy acquire has seen.
__device__ __forceinline__ void halo_wait_cta(const HaloCtl &h, bool first_chunk, bool last_chunk)
{
    if (!h.epoch)
        return;
    const bool wl = first_chunk && h.wait_low.flag, wu = last_chunk && h.wait_up.flag;
    if (!wl && !wu)
        return;  // an interior chunk: nothing of this block's input comes from a neighbour
    if (threadIdx.x == 0) {
        if (wl)
            halo_wait_one(h, h.wait_low, 1u);
        if (wu)
            halo_wait_one(h, h.wait_up, 2u);
        asm volatile("fence.proxy.async;" ::: "memory");
    }
    __syncthreads();
}

// End of a CTA that stored into the neighbour's memory, called by ALL its threads.
__device__ __forceinline__ void halo_signal_cta(const HaloCtl &h, const HaloPush &p)
{
    if (!p.peer_flag)
        return;
    // the barrier orders every thread's peer stores before thread 0's fence (fences are
    // cumulative over what the fencing thread has synchronised with: the pattern of a
    // grid-wide barrier), so ONE system-scope fence per CTA makes them all visible ...
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence_system();
        const unsigned int prev = atomicAdd(p.count, 1u);
        if (prev == p.nblocks - 1) {  // ... and the last CTA to get here has seen all of them
            *p.count = 0;
            __threadfence_system();
            halo_st_release(p.peer_flag, *h.epoch * kHaloEpochStride + p.off);
        }
    }
    // Round 2 tried to shorten this (fence.acq_rel.sys instead of the sc fences, one fence and
    // a relaxed flag store in the last CTA): 0.6 % faster on 2 GPUs and WRONG on 8 -- with
    // the consumer's CTAs already resident (programmatic dependent launch) the flag overtook
    // peer stores of the CTA's other warps (tests/dist_check.py, 65^3 over 8 ranks: norm off
    // in the 4th digit after cycle 2; either change alone passed).  The sc fences stay.
}

// a smoother's store of the pair (r0, r1) at in-plane offset `off` of local plane `plane`,
// repeated into the halo of whichever neighbour(s) that plane is a boundary plane for
// (with two owned planes per rank the first plane goes BOTH ways)
__device__ __forceinline__ void halo_mirror_pair(const HaloCtl &h, int plane, long long off,
                                                 bool ok0, bool ok1, double r0, double r1)
{
    double *pu = nullptr, *pl = nullptr;
    if (h.push_up.peer_flag) {
        if (plane == h.push_up.plane[0]) pu = h.push_up.dst[0];
        else if (plane == h.push_up.plane[1]) pu = h.push_up.dst[1];
    }
    if (h.push_low.peer_flag && plane == h.push_low.plane[0])
        pl = h.push_low.dst[0];
#pragma unroll
    for (int side = 0; side < 2; side++) {
        double *pp = side ? pl : pu;
        if (!pp)
            continue;
        pp += off;
        if (ok0 && ok1)
            *reinterpret_cast<double2 *>(pp) = make_double2(r0, r1);
        else if (ok0)
            pp[0] = r0;
        else if (ok1)
            pp[1] = r1;
    }
}

// does [ia, ib) contain one of the direction's planes?
__device__ __forceinline__ bool halo_takes_part(const HaloPush &p, int ia, int ib)
{
    return p.peer_flag && ((p.plane[0] >= ia && p.plane[0] < ib) ||
                           (p.plane[1] >= ia && p.plane[1] < ib));
}
#endif  // __CUDACC__

// host: number of plane chunks [lo + c*chunk, ...) of [lo, hi) that contain one of the planes
inline unsigned int halo_chunks_with(const HaloPush &p, int lo, int hi, int chunk)
{
    if (!p.peer_flag || chunk < 1)
        return 0;
    int c0 = -1, c1 = -1;
    if (p.plane[0] >= lo && p.plane[0] < hi)
        c0 = (p.plane[0] - lo) / chunk;
    if (p.plane[1] >= lo && p.plane[1] < hi)
        c1 = (p.plane[1] - lo) / chunk;
    if (c0 < 0 && c1 < 0)
        return 0;
    return (c0 >= 0 && c1 >= 0 && c0 != c1) ? 2u : 1u;
}

}  // namespace mgb
