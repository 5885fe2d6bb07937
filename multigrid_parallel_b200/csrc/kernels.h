// kernels.h -- launch wrappers of the sm_100a kernels (host-callable, C++).
#pragma once
#include <cuda_runtime.h>

#include "geom.h"
#include "halo.cuh"
#include "lu_band.cuh"

namespace mgb {

// number of partial sums a residual / sumsq launch may produce
constexpr int kMaxPartials = 1 << 16;

// natural <-> colour-split (API boundary only)
void launch_pack(const Geo &g, const double *nat, double *split, cudaStream_t st);
void launch_unpack(const Geo &g, const double *split, double *nat, cudaStream_t st);

// the same for a contiguous run [first, first+n) of the natural layout; `nat` holds
// just that run
void launch_pack_range(const Geo &g, const double *nat, double *split, long long first,
                       long long n, cudaStream_t st);
void launch_unpack_range(const Geo &g, const double *split, double *nat, long long first,
                         long long n, cudaStream_t st);

// Dirichlet faces: analytic BCFunc (mg_3d.h:89-90, 1147-1239)
void launch_set_dirichlet(const Geo &g, double *a, double h, cudaStream_t st);

// one colour of the RB-GS smoother over local planes [il_lo, il_hi)
// (mg_3d.h:432-443, 658-702)
void launch_half_sweep(const Geo &g, double *v, const double *d, double hSq,
                       int colour, int il_lo, int il_hi, cudaStream_t st,
                       const HaloCtl *h = nullptr);

// GaussSeidelSmoother's lexicographic sweeps as a hyperplane wavefront (gslex.cu,
// mg_3d.h:546-637); whole levels only; `bar` = one unsigned int of device scratch;
// non-zero: the cooperative launch failed
int launch_gs_lex(const Geo &g, double *v, const double *d, double hSq, int iters,
                  unsigned int *bar, cudaStream_t st);
// kernel selection of launch_gs_lex (MGB_G_GSLEX_TILE, env MGB_GSLEX_TILE)
void gs_lex_set_mode(int mode);

// the same when every neighbour is known to be zero (first sweep of a coarse level)
void launch_first_sweep_zero(const Geo &g, double *v, const double *d, double hSq, int colour,
                             int il_lo, int il_hi, cudaStream_t st, const HaloCtl *h = nullptr);

// residual (mg_3d.h:794-842) over local planes [il_lo, il_hi); r may be
// nullptr; the sum of squares lands in *out_sumsq (device) via `partials`
void launch_residual(const Geo &g, const double *v, const double *d, double *r,
                     double invHsq, int il_lo, int il_hi, double *partials,
                     double *out_sumsq, cudaStream_t st, const HaloCtl *h = nullptr);

// full-weighting restriction r(fine) -> d(coarse) (mg_3d.h:844-998) for local
// coarse planes [Il_lo, Il_hi)
void launch_restrict(const Geo &gf, const double *rf, const Geo &gc, double *dc,
                     int Il_lo, int Il_hi, cudaStream_t st);

// the two above in one pass, the fine residual never stored (17 B/DOF)
void launch_residual_restrict(const Geo &gf, const double *vf, const double *df,
                              double invHsq, const Geo &gc, double *dc, int Il_lo,
                              int Il_hi, cudaStream_t st, const HaloCtl *h = nullptr);

// trilinear prolongation + correction (mg_3d.h:1000-1145) for local fine
// planes [il_lo, il_hi)
// cmask: bit c set = colour c is corrected (3 = the reference's operation; inside the
// V-cycle only the colour the post-smoother does not overwrite first is needed)
void launch_prolong_correct(const Geo &gc, const double *ec, const Geo &gf,
                            double *ef, int il_lo, int il_hi, cudaStream_t st, int cmask = 3,
                            const HaloCtl *h = nullptr);
// a[p] = a[p] + 0. on the face points of colour `colour` in local planes [il_lo, il_hi):
// what the reference's `ef[p] += 0.` does to a boundary value (-0. becomes +0.)
void launch_add_zero_faces(const Geo &g, double *a, int colour, int il_lo, int il_hi,
                           cudaStream_t st);

// updateEdgeValues (mg_3d.h:304-430) on a whole (unpartitioned) level array
void launch_edge_values(const Geo &g, double *a, cudaStream_t st);

// sum of squares of the entries of one or two ranges (pads are zero) -> *out
void launch_sumsq(const double *a0, long long n0, const double *a1, long long n1,
                  double *partials, double *out, cudaStream_t st);
// sum over the points of local planes [il_lo, il_hi) of (u - BCFunc(ih,jh,kh))^2
// (test_mg_3d.c:78-97)
void launch_error_sumsq(const Geo &g, const double *u, double h, int il_lo, int il_hi,
                        double *partials, double *out, cudaStream_t st);

// second stage of every reduction: *out = sum of partials[0..n) in a fixed order
void launch_finish_sum(const double *partials, int n, double *out, cudaStream_t st);

// ---- tile kernels (tile.cu): plane-marching blocks fed by TMA bulk copies ----
// Both return false (nothing launched) when the level does not suit them; the
// caller then uses the plain kernels above.  colour < 0: no fused sweep;
// colour = 0/1: the half-sweep of that colour over the same planes is fused in
// (planes must then be ALL interior planes of the level, single-GPU levels only).
bool tile_enabled();
void tile_set(int on /* <0: keep */, long long min_plane /* <0: keep */);
// h (partitioned levels): the fused halo waits / pushes of halo.cuh
bool launch_tile_residual(const Geo &g, double *v, const double *d, double hSq, double invHsq,
                          int colour, int il_lo, int il_hi, double *partials, double *out_sumsq,
                          cudaStream_t st, const HaloCtl *h = nullptr);
bool launch_tile_residual_restrict(const Geo &gf, double *vf, const double *df, double hSq,
                                   double invHsq, int colour, const Geo &gc, double *dc,
                                   int Il_lo, int Il_hi, cudaStream_t st,
                                   const HaloCtl *h = nullptr);

// ---- multi-GPU plumbing outside the compute kernels (kernels.cu; halo.cuh has the
// fused part) ----
// explicit halo step: copy up to four contiguous runs of doubles per direction into the
// neighbour GPU's memory (IPC-mapped peer pointers), then -- once every block's stores are
// fenced system-wide -- release this direction's sequence number into the neighbour's flag
struct HaloRun {
    const double *src[4] = {nullptr, nullptr, nullptr, nullptr};
    double *dst[4] = {nullptr, nullptr, nullptr, nullptr};
    long long n[4] = {0, 0, 0, 0};
    unsigned long long *peer_flag = nullptr;  // nullptr: nothing goes this way
    unsigned int *count = nullptr;            // local arrival counter
    unsigned long long off = 0;               // sequence offset within the epoch
};
void launch_halo_push(const HaloRun &up, const HaloRun &low, const unsigned long long *epoch,
                      cudaStream_t st);
// explicit wait for everything both neighbours were asked to send so far (h.wait_*)
void launch_halo_wait(const HaloCtl &h, cudaStream_t st);
// end of a collective operation (see halo.cuh)
void launch_epoch_close(unsigned long long *xf, unsigned long long n_up, unsigned long long n_low,
                        cudaStream_t st);
// all-gather of the first replicated level's rhs slabs / sum of a scalar in rank order
struct GatherHost {
    int me, nranks;
    const double *src[2];
    long long n[2];
    double *dst[kMaxRanks][2];
    unsigned long long *peer_xf[kMaxRanks];
    unsigned long long *my_xf;
    unsigned long long off, timeout_ns;
    unsigned int *err;
};
void launch_gather(const GatherHost &g, cudaStream_t st);
struct NormHost {
    int me, nranks;
    double *scalar;
    unsigned long long *peer_xf[kMaxRanks];
    unsigned long long *my_xf;
    unsigned long long off, timeout_ns;
    unsigned int *err;
};
void launch_norm_exchange(const NormHost &g, cudaStream_t st);

// ---- the deep coarse levels as one single-block, shared-memory kernel (tail.cu) ----
struct TailLevel {
    Geo g;
    double *u, *d, *r;
    double hSq, invHsq;
};
struct TailP {
    int top;          // levels top .. 0 .. top are done in the kernel
    int gs;           // smoothing iterations per leg
    int zero_top;     // level `top` starts from a zero guess (it is a coarse level)
    LuBand lu;        // factorised operator of level 0 (tile form)
    int phase;        // 0: whole sub-cycle in one kernel (coarse system of <= 32 unknowns);
                      // 1: down leg, 2: up leg, the stand-alone LU solve launched in between
    int smem;         // working set held in shared memory (set by launch_coarse_tail)
    TailLevel lv[8];
};
void launch_coarse_tail(const TailP &p, cudaStream_t st);

// the half-sweep through the TMA ring (tile.cu); false: level too small, use the marching kernel
bool launch_tile_half_sweep(const Geo &g, double *v, const double *d, double hSq, int colour,
                            int il_lo, int il_hi, cudaStream_t st, const HaloCtl *h = nullptr);

// launch plan of a tile kernel (no launch): see tile.cu
void tile_plan_query(int kind, const Geo &gf, const Geo *gc, int p_lo, int p_hi, long long *out);
// prolongation + correction through the TMA ring (tile.cu); false: use launch_prolong_correct's
// marching kernels
bool launch_tile_prolong(const Geo &gc, const double *ec, const Geo &gf, double *ef, int il_lo,
                         int il_hi, int cmask, cudaStream_t st, const HaloCtl *h = nullptr);

// dense coarse operator + LU (mg_3d.h:147-273, gauss_elim.h:9-60)
void launch_coarse_matrix(double *A, int ni, int nj, int nk, double h,
                          cudaStream_t st);
// band-limited in-place factorisation of the dense n x n array `a` (half
// bandwidth bw; bw >= n-1 = plain dense), two launches whatever n is
void launch_lu_factor_band(double *a, int n, int bw, cudaStream_t st);
// half bandwidth of a dense matrix (max |r-c| over its non-zeros); synchronises
int lu_bandwidth(const double *a, int n, cudaStream_t st);
// storage of the tile form of the factor (lu_band.cuh); returns non-zero on failure
int lu_band_alloc(LuBand *B, int n, int bw);
void lu_band_free(LuBand *B);
// dense factor -> tiles, diagonal and its reciprocal
void launch_lu_extract_band(const double *a, const LuBand &B, cudaStream_t st);
// x = (LU)^-1 b in one launch: plain dense vectors, or level 0's colour-split d / u
void launch_lu_solve_dense(const LuBand &B, const double *b, double *x, cudaStream_t st);
void launch_lu_solve_level(const LuBand &B, const Geo &g, const double *d0, double *u0,
                           cudaStream_t st);
// launches issued through the wrappers above (all threads, all solvers)
long long launches_issued();

}  // namespace mgb
