// launch.h -- every kernel of the cycle is launched through launch_k(): a plain launch
// plus the "programmatic stream serialization" attribute (programmatic dependent launch).
//
// A V-cycle is ~60 dependent kernels, two thirds of them a few microseconds long; inside a
// CUDA graph each kernel -> kernel edge still costs the drain of the previous grid, the
// launch of the next one and its ramp-up.  With the attribute set, the blocks of kernel
// N+1 may become resident while the last blocks of kernel N are still running; they sit in
// `griddepcontrol.wait` (pdl_enter() below -- the first statement of every kernel
// launched this way, executed by every thread before it touches memory) until kernel N has
// completed and its writes are visible, so the data dependences between stages are exactly
// those of a plain stream.  Transitively safe: kernel N cannot complete before its own
// threads return from their wait, i.e. before kernel N-1 has completed.
// The attribute survives stream capture (programmatic edges in the graph).
// Measured on B200 (round 2): inside the graph-replayed cycle no difference beyond
// run-to-run noise (tools/cycle_case.py --batch 1: 3.695 vs 3.678 ms at 513^3, 0.212 vs
// 0.214 ms at 129^3 -- the small kernels are bound by their own first-load latency and
// drain, which a wait at the top cannot overlap); on back-to-back eager launches of
// mid-sized kernels it hides the launch gap: RB-GS at 257^3 (tools/bench_rbgs.py) 77.95 ->
// 74.66 us per full sweep.  ON by default in single-GPU processes; MGB_PDL=0 gives plain
// launches (pdl_enter() is then a no-op in hardware).  The TMA half-sweep kernel waits only after its
// shared-memory set-up (pdl_trigger() / pdl_wait() apart).
#pragma once
#include <cuda_runtime.h>

#include <cstdlib>
#include <utility>

namespace mgb {

// a partitioned solver exists in this process: its kernels also synchronise with PEER GPUs
// through flags in memory, PDL buys them nothing measurable (2 GPUs: 3.891 vs 3.873 ms per
// cycle) and makes every latent ordering weakness visible (halo.cuh) -- plain launches there
// unless MGB_PDL=1 insists
inline bool &pdl_dist_active()
{
    static bool active = false;
    return active;
}

inline bool pdl_enabled()
{
    static const int mode = getenv("MGB_PDL") ? atoi(getenv("MGB_PDL")) : -1;
    return mode < 0 ? !pdl_dist_active() : mode != 0;
}

// cudaFuncSetAttribute acts on the CURRENT device: call sites remember per device what they
// have set (a process may hold solvers on several GPUs)
inline bool first_on_device(unsigned long long &seen)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64)
        return true;
    if ((seen >> dev) & 1ull)
        return false;
    seen |= 1ull << dev;
    return true;
}
// the same for an attribute that grows with the request (dynamic shared memory)
inline bool grows_on_device(size_t (&allowed)[64], size_t want)
{
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64)
        return want > 48 * 1024;
    if (allowed[dev] == 0)
        allowed[dev] = 48 * 1024;
    if (want <= allowed[dev])
        return false;
    allowed[dev] = want;
    return true;
}

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kern)(P...), dim3 grid, dim3 block, size_t smem,
                            cudaStream_t st, A &&...args)
{
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    if (pdl_enabled()) {
        cfg.attrs = at;
        cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kern, std::forward<A>(args)...);
}

// first statement of every kernel launched through launch_k(): let the next
// kernel's blocks become resident as soon as there is room, then wait until the previous
// kernel has completed and its writes are visible.  No memory access may precede it.
__device__ __forceinline__ void pdl_trigger()
{
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
__device__ __forceinline__ void pdl_wait()
{
    asm volatile("griddepcontrol.wait;" ::: "memory");
}
__device__ __forceinline__ void pdl_enter()
{
    pdl_trigger();
    pdl_wait();
}

}  // namespace mgb
