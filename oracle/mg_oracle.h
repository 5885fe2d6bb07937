/*
 * mg_oracle.h -- CPU oracle for the 3D multigrid V-cycle hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline/reference
 * legs may load it, and there only as the checker.
 *
 * It is a plain-C, serial restatement of the algorithm in the reference's
 * mg_3d.h / gauss_elim.h, generalised from cubes to (ni,nj,nk) boxes with
 * 64-bit linear indices (needed for the weak-scaling grids).  For cubes every
 * function reproduces the reference bit for bit when both are built with
 * `gcc -O2 -ffp-contract=off` (pinned in tests/test_oracle_vs_ref.py against
 * oracle/_ref/libmg_ref.so, which is the reference compiled from
 * /root/reference, and against tests/golden/).
 *
 * Index convention (reference mg_3d.h:43-44,173): p = (i*nj + j)*nk + k,
 * k contiguous, i slowest.
 */
#ifndef MG_ORACLE_H
#define MG_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

/* mg_3d.h:89-90 */
double orc_bcfunc(double x, double y, double z);

/* mg_3d.h:1147-1239: analytic values on the six faces */
void orc_set_dirichlet(double *v, int ni, int nj, int nk, double h);

/* mg_3d.h:640-709 (first_red=1) and 711-781 (first_red=0) */
void orc_smooth(double *v, const double *d, int ni, int nj, int nk, double h,
                int iters, int first_red);

/* one colour only: colour 1 = red = (i+j+k) odd, 0 = black (mg_3d.h:669-693) */
/* GaussSeidelSmoother (mg_3d.h:546-637): lexicographic sweeps (+ updateEdgeValues) */
void orc_gs_lex(double *v, const double *d, int ni, int nj, int nk, double h, int iters,
                int edges);
/* updateEdgeValues (mg_3d.h:304-430) for a box */
void orc_edge_values(double *v, int ni, int nj, int nk);
void orc_half_sweep(double *v, const double *d, int ni, int nj, int nk,
                    double h, int colour);

/* mg_3d.h:794-842; res may be NULL; returns sqrt(sum of squares), summed in
 * (i,j,k) order like a one-thread run of the reference */
double orc_residual(const double *v, const double *d, int ni, int nj, int nk,
                    double h, double *res);

/* mg_3d.h:844-998 */
void orc_restrict(const double *r, int nif, int njf, int nkf, double *dc,
                  int nic, int njc, int nkc);

/* mg_3d.h:1000-1145 */
void orc_prolong_correct(const double *ec, int nic, int njc, int nkc,
                         double *ef, int nif, int njf, int nkf);

/* mg_3d.h:147-273 */
void orc_coarse_matrix(double *A, int ni, int nj, int nk, double h);

/* gauss_elim.h:9-29 and 31-60 */
void orc_lu_factor(double *a, int n);
void orc_lu_solve(const double *lu, int n, const double *b, double *x);

/* mg_3d.h:783-792 */
double orc_l2norm(const double *d, long n);
/* writeOutputData (postprocess.h:5-47) for a box; non-zero on I/O failure */
int orc_write_vtk(const char *path, const double *grid, int ni, int nj, int nk, double h);

/* ---- multilevel driver (mg_3d.h:107-144, 275-293, 1242-1362) ---- */
typedef struct orc_mg orc_mg;

/* coarse extents (ci,cj,ck), `levels` grids, `gs` smoothing iterations;
 * h_fine = 1/(nk_fine-1) like GRID_LENGTH=1 in the reference drivers */
orc_mg *orc_mg_create(int ci, int cj, int ck, int levels, int gs);
void orc_mg_destroy(orc_mg *m);
void orc_mg_dims(const orc_mg *m, int level, int *ni, int *nj, int *nk);
double *orc_mg_u(orc_mg *m, int level);
double *orc_mg_d(orc_mg *m, int level);
double *orc_mg_r(orc_mg *m, int level);
double orc_mg_h(const orc_mg *m);
/* one V-cycle from the finest level; returns the post-smoothing residual norm */
double orc_mg_vcycle(orc_mg *m);

/* SolverFMGInitialize (mg_3d.h:1364-1404, commented out upstream) restated on
 * the functions above; see mg_oracle.c for what the statement order implies */
void orc_mg_fmg_init(orc_mg *m);
/* test_mg_3d.c:17-29: BCs into the faces of d and u on the finest level; returns ||d|| */
double orc_mg_setup_problem(orc_mg *m);

/* test_mg_3d.c:8-68 flow: BCs into d and u, threshold tol*||d||, cycle until
 * norm <= threshold or max_cycles; history[c] = norm after cycle c+1.
 * Returns the number of cycles run. */
int orc_mg_solve(orc_mg *m, double tol, int max_cycles, double *history,
                 double *init_norm);

#ifdef __cplusplus
}
#endif
#endif
