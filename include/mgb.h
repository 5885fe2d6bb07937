/*
 * mgb.h -- C ABI of libmgb.so: the 3D geometric-multigrid V-cycle hot path
 * (red-black Gauss-Seidel, residual, full-weighting restriction, trilinear
 * prolongation, dense-LU coarsest solve) as hand-written sm_100a CUDA.
 *
 * This is the drop-in boundary for knram06/multigrid_parallel's mg_3d.h: plain
 * pointers and sizes, no C++/torch types.  Every entry point names the
 * reference interface it replaces (file:line into the reference tree).  The
 * host-side mirror of the reference's own API (SolverInitialize & co.) lives in
 * multigrid_parallel_b200/compat/mg_3d.h and is written purely against this
 * header.
 *
 * Conventions
 *   - all grids are fp64; host arrays are in the reference's NATURAL layout
 *     p = (i*nj + j)*nk + k, k contiguous (mg_3d.h:43-44); device storage is
 *     colour-split and private to the library
 *   - level 0 is the coarsest, level L-1 the finest (mg_3d.h:41)
 *   - colour 1 = "red" = (i+j+k) odd, colour 0 = "black" (mg_3d.h:669-693)
 *   - every function returns 0 on success, non-zero on failure;
 *     mgb_last_error() then describes the failure.  There is NO CPU fallback:
 *     without a usable CUDA device every call fails.
 */
#ifndef MGB_H
#define MGB_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgb_solver mgb_solver;

enum { MGB_U = 0, MGB_D = 1, MGB_R = 2 };          /* which level array   */
enum { MGB_BLACK = 0, MGB_RED = 1 };               /* colours             */
enum {                                             /* TimingInfo stages, mg_3d.h:136-137 */
    MGB_ST_SMOOTH1 = 0, MGB_ST_RESID1 = 1, MGB_ST_RESTRICT = 2,
    MGB_ST_RECURSE = 3, MGB_ST_PROLONG = 4, MGB_ST_SMOOTH2 = 5,
    MGB_ST_RESID2 = 6, MGB_NUM_STAGES = 7
};
enum {                                             /* mgb_set_option keys */
    MGB_OPT_GRAPH = 0,      /* 1: replay the V-cycle as one CUDA graph (default 1)        */
    MGB_OPT_PROFILE = 1,    /* 1: per-stage CUDA-event timing, eager launches (default 0)  */
    MGB_OPT_FUSE = 2,       /* 0: every stage its own kernel, r stored; 1 (default): fused
                             * residual+restriction, r not stored; 2: also the last colour
                             * of each smoother leg inside the residual kernel that follows */
    MGB_OPT_GRAPH_LEVELS = 3,/* only levels < this are graphed when PROFILE=1              */
    MGB_OPT_TAIL = 4,       /* 1 (default): the levels of <= ~17^3 points run as ONE kernel */
    MGB_OPT_ZERO_GUESS = 5, /* 1 (default): coarse levels are not zeroed before pre-smoothing,
                             * their first half-sweep takes the zero guess as given       */
    MGB_OPT_PROLONG_MASK = 6/* 1 (default): inside the cycle the prolongation corrects only
                             * the colour the post-smoother does not overwrite first      */
};

/* ---- errors / device ---------------------------------------------------- */
const char *mgb_last_error(void);
int mgb_device_count(int *count);
const char *mgb_version(void);

/* ---- lifecycle ----------------------------------------------------------
 * mgb_create replaces SolverInitialize (mg_3d.h:107-144: level sizes
 * (c-1)*2^l+1, zeroed u/d/r per level, spacing = 1/(n_fine-1)) and the
 * coarse-operator half of SolverGetDetails (mg_3d.h:281-288:
 * constructCoarseMatrixA with the coarse spacing + convertToLU_InPlace,
 * gauss_elim.h:9-29) -- built and factorised on the GPU, once.
 * (ci,cj,ck) = coarsest-grid points per axis, each 2^m+1; the reference only
 * has cubes (ci=cj=ck).  spacing h = 1/(nk_fine-1). */
int mgb_create(mgb_solver **out, int ci, int cj, int ck, int levels,
               int gs_iters, int device);
int mgb_destroy(mgb_solver *s);
int mgb_levels(const mgb_solver *s);
int mgb_dims(const mgb_solver *s, int level, int *ni, int *nj, int *nk);
double mgb_spacing(const mgb_solver *s, int level);
int mgb_set_option(mgb_solver *s, int key, int value);
/* process-wide kernel selection (affects launches captured afterwards; CUDA
 * graphs already built keep what they captured) */
enum {
    MGB_G_TILE = 0,           /* 1 (default): TMA tile kernels for residual / residual+restrict */
    MGB_G_TILE_MIN_PLANE = 1, /* use them on levels with nj*nk >= value (default 40000)         */
    MGB_G_GSLEX_TILE = 2      /* mgb_gs_lex kernel: 1 (default) by level size, 0 global
                                 hyperplanes, 2..5 force the tile wavefront with tiles of
                                 8x16x32, 8x8x32, 8x32x32, 16x16x32 points */
};
int mgb_set_global(int key, long long value);
/* the launch plan the TMA tile kernels would use on an ni x nj x nk level, without
 * launching anything (works without a GPU; used by the CPU-side tests): kind 0
 * residual norm, 1 residual+restrict (into the (n+1)/2 level), 2 half-sweep,
 * 3 prolongation; out12 = {ok, grid.x, grid.y, grid.z, threads, dynamic smem
 * bytes, tile rows, tile quads, rows advance, quads advance, coarse rows per
 * tile, planes per chunk} */
int mgb_tile_plan(int kind, int ni, int nj, int nk, long long *out12);
int mgb_sync(mgb_solver *s);

/* ---- level arrays: the reference hands out raw host pointers u[l], d[l],
 * r[l] (mg_3d.h:26, 278-279); here they cross the boundary explicitly.
 * `host` is natural layout, ni*nj*nk doubles; pinned or pageable. */
int mgb_upload(mgb_solver *s, int level, int which, const double *host);
int mgb_download(mgb_solver *s, int level, int which, double *host);
/* the same for a contiguous run of the natural layout: doubles [first, first+count)
 * of the local planes; `host` points at the first of them */
int mgb_upload_range(mgb_solver *s, int level, int which, long long first, long long count,
                     const double *host);
int mgb_download_range(mgb_solver *s, int level, int which, long long first, long long count,
                       double *host);
int mgb_zero(mgb_solver *s, int level, int which);
/* setupBoundaryConditions (mg_3d.h:1147-1239) on the device array */
int mgb_set_dirichlet(mgb_solver *s, int level, int which);
/* updateEdgeValues (mg_3d.h:304-430; SolverSmoothenEdgeValues, 1422-1423): inner
 * points of the 12 edges = mean of their two inward neighbours, then the 8 corners
 * = mean of their three edge neighbours.  Unpartitioned levels only. */
int mgb_edge_values(mgb_solver *s, int level, int which);
/* GaussSeidelSmoother's sweeps (mg_3d.h:546-634; driver test_gs_3d.c:56): `iters`
 * LEXICOGRAPHIC Gauss-Seidel sweeps over the interior of u[level] with rhs d[level],
 * run as a hyperplane wavefront on the GPU -- bit-identical to the reference's serial
 * (i,j,k) loop.  The reference routine ends with updateEdgeValues: call
 * mgb_edge_values(s, level, MGB_U) after it for the whole routine.  Unpartitioned
 * levels only. */
int mgb_gs_lex(mgb_solver *s, int level, int iters);
/* single-grid sessions (levels == 1, the raw-pointer API preSmoother / postSmoother /
 * calculateResidual(v, d, N, h, ..) of test_rb_gs_3d.c:56-101): the caller's spacing */
int mgb_set_spacing(mgb_solver *s, double h);
/* page-lock a host range the caller owns (cudaHostRegister) so that mgb_upload /
 * mgb_download run at the pinned PCIe rate; the drop-in headers do this for the
 * finest-level grid / rhs arrays they hand out (SolverGetDetails, mg_3d.h:275-293) */
int mgb_pin_host(void *p, unsigned long long bytes);
int mgb_unpin_host(void *p);
/* GetL2NormOfVector (mg_3d.h:783-792): sum of squares over ALL points */
int mgb_sumsq(mgb_solver *s, int level, int which, double *sumsq);
/* test_mg_3d.c:78-97: sum over all points of (u - BCFunc)^2 at the finest level */
int mgb_error_sumsq(mgb_solver *s, double *sumsq);

/* ---- operators, one level at a time -------------------------------------
 * mgb_half_sweep : one colour of smoothenAtIndex over the interior
 *                  (mg_3d.h:432-443, loops 658-702)
 * mgb_smooth     : preSmoother (first_red=1, mg_3d.h:640-709) or
 *                  postSmoother (first_red=0, mg_3d.h:711-781)
 * mgb_residual   : calculateResidual (mg_3d.h:794-842); store_r=0 is the
 *                  res==NULL form; *sumsq = sum of diff^2 (caller takes sqrt)
 * mgb_restrict   : restrictResidual r[level] -> d[level-1] (mg_3d.h:844-998)
 * mgb_residual_restrict : the two above fused, r[level] not stored
 * mgb_prolong_correct   : prolongateAndCorrectError u[level-1] -> u[level]
 *                  (mg_3d.h:1000-1145)
 * mgb_coarse_solve      : solveWithLU on level 0 (gauss_elim.h:31-60) */
int mgb_half_sweep(mgb_solver *s, int level, int colour);
/* tuning aid, not part of the reference's API: the same over the local planes
 * [il_lo, il_hi) only (what one rank of a partitioned solver runs on its slab) */
int mgb_debug_half_sweep_range(mgb_solver *s, int level, int colour, int il_lo, int il_hi);
int mgb_smooth(mgb_solver *s, int level, int iters, int first_red);
int mgb_residual(mgb_solver *s, int level, int store_r, double *sumsq);
int mgb_restrict(mgb_solver *s, int level);
int mgb_residual_restrict(mgb_solver *s, int level);
int mgb_prolong_correct(mgb_solver *s, int level);
/* one colour of the smoother (mg_3d.h:681-702 / 753-773) and the stage that
 * follows it in vcycle (mg_3d.h:1294+1310 resp. 1354) in ONE pass over the
 * level: the new values never make a round trip through HBM before the
 * residual reads them.  Same arithmetic as mgb_half_sweep followed by
 * mgb_residual_restrict / mgb_residual(store_r=0), bit for bit. */
int mgb_sweep_residual_restrict(mgb_solver *s, int level, int colour);
int mgb_sweep_residual(mgb_solver *s, int level, int colour, double *sumsq);
int mgb_coarse_solve(mgb_solver *s);
/* the factorised coarse operator, row-major n x n (n = ci*cj*ck): what
 * convertToLU_InPlace leaves in `A` (gauss_elim.h:9-29), bit for bit */
int mgb_coarse_lu_download(mgb_solver *s, double *host_lu);
/* size of the coarsest system, the half bandwidth (cj*ck) the factorisation and
 * the solves are restricted to (exact: see csrc/lu.cu), and the device time the
 * build + factorisation took at mgb_create (the reference's
 * constructCoarseMatrixA + convertToLU_InPlace, mg_3d.h:281-288) */
int mgb_coarse_info(const mgb_solver *s, int *n, int *half_bandwidth, double *factor_seconds);

/* ---- the V-cycle ---------------------------------------------------------
 * mgb_vcycle = SolverLinSolve -> vcycle (mg_3d.h:1415-1420, 1242-1362): one
 * V(gs,gs) cycle from the finest level; *sumsq = squared 2-norm of the
 * post-smoothing residual (the reference returns per-thread sqrt partials whose
 * squares the driver sums, test_mg_3d.c:45-59).
 * mgb_solve = the driver loop test_mg_3d.c:40-66: cycle while
 * sqrt(sumsq) > threshold; history[c] = norm after cycle c+1. */
int mgb_vcycle(mgb_solver *s, double *sumsq);
/* n cycles enqueued back to back; only the last norm is read back (no per-cycle host
 * round trip: for a fixed cycle count, where the driver loop test_mg_3d.c:40-66 would
 * not look at the intermediate norms anyway) */
int mgb_vcycles(mgb_solver *s, int n, double *sumsq);
/* mgb_fmg_init = SolverFMGInitialize (mg_3d.h:1364-1404; upstream keeps it
 * commented out, the live copy mg_dirichlet_analytic.c:771-806 predates today's
 * vcycle signature): coarsest LU solve, then per level interpolate the coarser
 * solution, impose the boundary values, zero the coarser level, one V-cycle
 * entered at that level.  Exactly those statements in that order -- including
 * their quirks: the LU solve overwrites the boundary values put into u[0], and a
 * V-cycle entered below the finest level zeroes its entry level first
 * (mg_3d.h:1254-1260) -- so every level's u and d equal the reference's bit for
 * bit (tests/golden/fmg.json).  *sumsq = squared residual norm after the last
 * (finest-level) cycle.  Call it where mg_dirichlet_analytic.c:984-989 does:
 * after the boundary values are in place, before the V-cycle loop. */
int mgb_fmg_init(mgb_solver *s, double *sumsq);
int mgb_solve(mgb_solver *s, double threshold, int max_cycles,
              double *history, int *cycles);

/* ---- timing_info.h feed: accumulated seconds and call counts per level and
 * stage (mg_3d.h:1279-1359), measured with CUDA events when MGB_OPT_PROFILE=1 */
int mgb_timing(mgb_solver *s, int level, int stage, int *calls, double *seconds);
int mgb_timing_reset(mgb_solver *s);
/* number of kernel launches (graph nodes included) issued since creation */
long long mgb_launch_count(const mgb_solver *s);
/* CUDA-event stopwatch on the solver's own stream (the stream every kernel of
 * this solver is launched on): start records an event, stop records a second
 * one, waits for it and returns the device time between them */
int mgb_timer_start(mgb_solver *s);
int mgb_timer_stop(mgb_solver *s, double *seconds);
/* the solver's cudaStream_t (for interop, e.g. torch.cuda.ExternalStream) */
void *mgb_stream(mgb_solver *s);

/* ---- multi-GPU: one process per GPU, the i axis cut into slabs -----------
 * The reference partitions every loop over i with `omp for schedule(static)`
 * (mg_3d.h:658,681,729,753,807,962,1006); here rank p of P owns the planes
 * [p*w, (p+1)*w), w = (ni-1)/P, plus halo planes, and after each colour
 * half-sweep the freshly written boundary plane goes to the neighbour's halo
 * over NCCL (NVLink).  Levels coarser than the first partitioned one are
 * agglomerated on rank 0.  Results are bit-identical to the single-GPU path.
 *   rank 0 : mgb_nccl_unique_id(buf[128]); ship buf to all ranks (any transport)
 *   every  : mgb_create_dist(..., rank, nranks, buf, min_planes, min_points)
 * In a partitioned solver mgb_upload/mgb_download/mgb_set_dirichlet address the
 * rank's LOCAL planes [i0, i0+li) of mgb_local_range; norms are summed over
 * the ranks; every compute entry point is collective over the ranks. */
int mgb_nccl_unique_id(void *out128);
int mgb_create_dist(mgb_solver **out, int ci, int cj, int ck, int levels,
                    int gs_iters, int device, int rank, int nranks,
                    const void *nccl_uid128, int min_planes_per_rank,
                    long long min_points_per_rank /* < 0: default 2^22 / nranks */);
int mgb_dist_info(const mgb_solver *s, int *rank, int *nranks, int *first_dist_level);
int mgb_local_range(const mgb_solver *s, int level, int *i0, int *li, int *own_lo,
                    int *own_hi);
/* slab arithmetic, usable without a GPU */
int mgb_plan_slab(int ni, int nranks, int rank, int *own_lo, int *own_hi);
int mgb_plan_first_dist_level(int ci, int cj, int ck, int levels, int nranks,
                              int min_planes, long long min_points);

/* ---- stateless array entry points for the reference's raw-pointer API
 * (test_rb_gs_3d.c:70-81 and test_lu.c:33-42 call these on caller-owned host
 * arrays).  Each call stages host -> device, runs the CUDA kernels, stages
 * back; nothing is computed on the host. */
int mgb_host_smooth(double *v, const double *d, int ni, int nj, int nk,
                    double h, int iters, int first_red);
/* GaussSeidelSmoother(v, d, N, h, iters) (mg_3d.h:546-637): lexicographic sweeps and,
 * with edges != 0, the updateEdgeValues(v, N) that ends the reference routine */
int mgb_host_gs_lex(double *v, const double *d, int ni, int nj, int nk, double h, int iters,
                    int edges);
int mgb_host_residual(const double *v, const double *d, int ni, int nj, int nk,
                      double h, double *res /* may be NULL */, double *sumsq);
int mgb_host_restrict(const double *r, int nif, int njf, int nkf, double *dc,
                      int nic, int njc, int nkc);
int mgb_host_prolong_correct(const double *ec, int nic, int njc, int nkc,
                             double *ef, int nif, int njf, int nkf);
int mgb_host_coarse_matrix(double *A, int ni, int nj, int nk, double h);
int mgb_host_lu_factor(double *a, int n);
int mgb_host_lu_solve(const double *lu, int n, const double *b, double *x);

/* ---- writeOutputData (postprocess.h:5-47) with the text produced on the GPU -----------
 * The reference's ASCII legacy-VTK file of a grid of ni x nj x nk values (k fastest) and
 * spacing h, as a stream of byte chunks: header, one "%10.8e %10.8e %10.8e\n" line per
 * point, the POINT_DATA header, one "%10.8e\n" line per value.  Concatenated, the chunks
 * are byte for byte what the reference's fprintf loop writes (csrc/vtk.cu: exact decimal
 * conversion on the device; the few values whose rounding the device cannot decide with
 * certainty make their chunk fall back to the C library's snprintf).  `values` is a host
 * array that must stay readable until mgb_vtk_close.  Usage:
 *     mgb_vtk_open(&w, grid, N, N, N, h, 0);
 *     while (!mgb_vtk_next(w, &p, &n) && n) fwrite(p, 1, n, file);
 *     mgb_vtk_close(w);
 * A chunk stays valid until the next mgb_vtk_next / mgb_vtk_close on the same writer. */
typedef struct mgb_vtk mgb_vtk;
int mgb_vtk_open(mgb_vtk **w, const double *values, int ni, int nj, int nk, double h,
                 int device);
int mgb_vtk_next(mgb_vtk *w, const char **bytes, long long *n);
int mgb_vtk_host_chunks(const mgb_vtk *w, long long *chunks);
int mgb_vtk_close(mgb_vtk *w);

/* ---- resident single-grid session for the RB-GS microbenchmark
 * (test_rb_gs_3d.c flow: one grid, no hierarchy).  A 1-level solver:
 * mgb_create(.., levels=1, ..) gives u/d/r on the finest grid only and no
 * coarse operator is built when ci*cj*ck exceeds MGB_MAX_DENSE_N. */
#define MGB_MAX_DENSE_N 8192

#ifdef __cplusplus
}
#endif
#endif /* MGB_H */
