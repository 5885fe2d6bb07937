/*
 * mg_oracle.c -- CPU oracle (see mg_oracle.h: TEST INFRASTRUCTURE ONLY).
 *
 * Serial restatement of /root/reference/mg_3d.h + gauss_elim.h for boxes.
 * Build with: gcc -O2 -ffp-contract=off (no -march=native, no fast-math):
 * the operation order below IS the contract.
 */
#include "mg_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef long long i64;

#define IDX(i, j, k) (((i64)(i) * nj + (j)) * nk + (k))

double orc_bcfunc(double x, double y, double z)
{
    /* mg_3d.h:89-90  x*x - 2*y*y + z*z, evaluated left to right */
    return x * x - 2 * y * y + z * z;
}

void orc_set_dirichlet(double *v, int ni, int nj, int nk, double h)
{
    /* mg_3d.h:1147-1239.  Edges/corners are written several times with the
     * same value, so the face order is immaterial. */
    for (int i = 0; i < ni; i++)
        for (int j = 0; j < nj; j++)
            for (int k = 0; k < nk; k++) {
                if (i == 0 || i == ni - 1 || j == 0 || j == nj - 1 || k == 0 ||
                    k == nk - 1)
                    v[IDX(i, j, k)] = orc_bcfunc(i * h, j * h, k * h);
            }
}

void orc_half_sweep(double *v, const double *d, int ni, int nj, int nk,
                    double h, int colour)
{
    /* mg_3d.h:432-443 (point update) and 658-702 (colour loops).
     * colour 1 (red): k starts at 1+(i+j)%2, i.e. (i+j+k) odd. */
    const double hSq = h * h;
    const double sixth = 1. / 6;
    const i64 sj = nk, si = (i64)nj * nk;
    for (int i = 1; i < ni - 1; i++)
        for (int j = 1; j < nj - 1; j++) {
            int k0 = 1 + ((i + j + (colour ? 0 : 1)) % 2);
            for (int k = k0; k < nk - 1; k += 2) {
                i64 p = IDX(i, j, k);
                v[p] = sixth * (v[p - si] + v[p + si] + v[p - sj] + v[p + sj] +
                                v[p - 1] + v[p + 1] - hSq * d[p]);
            }
        }
}

void orc_smooth(double *v, const double *d, int ni, int nj, int nk, double h,
                int iters, int first_red)
{
    /* preSmoother: red then black (mg_3d.h:655-703);
     * postSmoother: black then red (mg_3d.h:726-774). */
    for (int s = 0; s < iters; s++) {
        orc_half_sweep(v, d, ni, nj, nk, h, first_red ? 1 : 0);
        orc_half_sweep(v, d, ni, nj, nk, h, first_red ? 0 : 1);
    }
}

void orc_edge_values(double *v, int ni, int nj, int nk)
{
    /* updateEdgeValues (mg_3d.h:304-430) for a box: the inner points of the 12 edges
     * become the mean of their two inward neighbours (311-397: the four edges along j
     * and along k of the X faces, then the four along i -- an edge point reads only
     * face-interior points, so the order among edges is immaterial), then each of the 8
     * corners the mean of its three edge neighbours, added in k, j, i order (399-429). */
    const i64 s[3] = {(i64)nj * nk, nk, 1};
    const int n[3] = {ni, nj, nk};
    for (int fa = 0; fa < 3; fa++) { /* axis along the edge */
        const int b = (fa + 1) % 3, c = (fa + 2) % 3;
        for (int eb = 0; eb < 2; eb++)
            for (int ec = 0; ec < 2; ec++) {
                const i64 inb = eb ? -s[b] : s[b], inc = ec ? -s[c] : s[c];
                for (int t = 1; t < n[fa] - 1; t++) {
                    const i64 p = t * s[fa] + (eb ? n[b] - 1 : 0) * s[b] + (ec ? n[c] - 1 : 0) * s[c];
                    v[p] = 0.5 * (v[p + inb] + v[p + inc]);
                }
            }
    }
    for (int ei = 0; ei < 2; ei++)
        for (int ej = 0; ej < 2; ej++)
            for (int ek = 0; ek < 2; ek++) {
                const i64 p = (ei ? ni - 1 : 0) * s[0] + (ej ? nj - 1 : 0) * s[1] + (ek ? nk - 1 : 0);
                const i64 ini = ei ? -s[0] : s[0], inj = ej ? -s[1] : s[1], ink = ek ? -1 : 1;
                v[p] = (1. / 3) * (v[p + ink] + v[p + inj] + v[p + ini]);
            }
}

void orc_gs_lex(double *v, const double *d, int ni, int nj, int nk, double h, int iters,
                int edges)
{
    /* GaussSeidelSmoother (mg_3d.h:546-637): `iters` lexicographic sweeps (559-634: the
     * same point formula as smoothenAtIndex, points visited in ascending (i,j,k), each
     * update visible to the next), then updateEdgeValues (635). */
    const double hSq = h * h;
    const double sixth = 1. / 6;
    const i64 sj = nk, si = (i64)nj * nk;
    for (int s = 0; s < iters; s++)
        for (int i = 1; i < ni - 1; i++)
            for (int j = 1; j < nj - 1; j++)
                for (int k = 1; k < nk - 1; k++) {
                    i64 p = IDX(i, j, k);
                    v[p] = sixth * (v[p - si] + v[p + si] + v[p - sj] + v[p + sj] +
                                    v[p - 1] + v[p + 1] - hSq * d[p]);
                }
    if (edges)
        orc_edge_values(v, ni, nj, nk);
}

double orc_residual(const double *v, const double *d, int ni, int nj, int nk,
                    double h, double *res)
{
    /* mg_3d.h:794-842 */
    const double invHsq = 1. / (h * h);
    const i64 sj = nk, si = (i64)nj * nk;
    double acc = 0.;
    for (int i = 1; i < ni - 1; i++)
        for (int j = 1; j < nj - 1; j++)
            for (int k = 1; k < nk - 1; k++) {
                i64 p = IDX(i, j, k);
                double diff =
                    d[p] - invHsq * (v[p - si] + v[p + si] + v[p - sj] +
                                     v[p + sj] + v[p - 1] + v[p + 1] - 6 * v[p]);
                if (res)
                    res[p] = diff;
                acc += diff * diff;
            }
    return sqrt(acc);
}

void orc_restrict(const double *r, int nif, int njf, int nkf, double *dc,
                  int nic, int njc, int nkc)
{
    /* mg_3d.h:844-998.  Faces: injection of r(2I,2J,2K) (881-957); interior:
     * 27-point full weighting, accumulated from 0. in (ti,tj,tk) order
     * (961-995).  Weight of an offset = 2^-(3+number of non-centre axes). */
    double w[3][3][3];
    for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++)
            for (int c = 0; c < 3; c++) {
                int off = (a != 1) + (b != 1) + (c != 1);
                w[a][b][c] = ldexp(1.0, -(3 + off));
            }
    for (int I = 0; I < nic; I++)
        for (int J = 0; J < njc; J++)
            for (int K = 0; K < nkc; K++) {
                i64 pc = ((i64)I * njc + J) * nkc + K;
                i64 pf = ((i64)(2 * I) * njf + 2 * J) * nkf + 2 * K;
                if (I == 0 || I == nic - 1 || J == 0 || J == njc - 1 || K == 0 ||
                    K == nkc - 1) {
                    dc[pc] = r[pf];
                    continue;
                }
                double val = 0.;
                for (int a = 0; a < 3; a++)
                    for (int b = 0; b < 3; b++)
                        for (int c = 0; c < 3; c++) {
                            i64 q = pf + (i64)(a - 1) * njf * nkf +
                                    (i64)(b - 1) * nkf + (c - 1);
                            val += r[q] * w[a][b][c];
                        }
                dc[pc] = val;
            }
}

void orc_prolong_correct(const double *ec, int nic, int njc, int nkc,
                         double *ef, int nif, int njf, int nkf)
{
    /* mg_3d.h:1000-1145.  For each fine point the contributing coarse corners
     * are added to 0. one by one in the reference's order, then scaled:
     *   3 odd axes: i-major, then j, then k          (1023-1049)  * 0.125
     *   2 odd axes: i even -> j fastest, then k      (1059-1068)
     *               j even -> i fastest, then k      (1070-1079)
     *               k even -> j fastest, then i      (1080-1089)  * 0.25
     *   1 odd axis: low end, high end                (1101-1134)  * 0.5
     *   0 odd axes: the coarse value itself          (1137-1138)
     * (nif is implied by the loops; kept for symmetry.) */
    (void)nic;
    const i64 sJ = nkc, sI = (i64)njc * nkc;
    for (int i = 0; i < nif; i++)
        for (int j = 0; j < njf; j++)
            for (int k = 0; k < nkf; k++) {
                const int oi = i & 1, oj = j & 1, ok = k & 1;
                const i64 base = (i64)(i >> 1) * sI + (i64)(j >> 1) * sJ + (k >> 1);
                i64 off[8];
                int n = 0;
                double scale = 1.;
                switch (oi + oj + ok) {
                case 3:
                    for (int a = 0; a < 2; a++)
                        for (int b = 0; b < 2; b++)
                            for (int c = 0; c < 2; c++)
                                off[n++] = a * sI + b * sJ + c;
                    scale = 0.125;
                    break;
                case 2:
                    if (!oi) { /* j fastest, then k */
                        off[0] = 0; off[1] = sJ; off[2] = 1; off[3] = sJ + 1;
                    } else if (!oj) { /* i fastest, then k */
                        off[0] = 0; off[1] = sI; off[2] = 1; off[3] = sI + 1;
                    } else { /* j fastest, then i */
                        off[0] = 0; off[1] = sJ; off[2] = sI; off[3] = sI + sJ;
                    }
                    n = 4;
                    scale = 0.25;
                    break;
                case 1:
                    off[0] = 0;
                    off[1] = oi ? sI : (oj ? sJ : 1);
                    n = 2;
                    scale = 0.5;
                    break;
                default:
                    break;
                }
                double add;
                if (n == 0) {
                    add = ec[base];
                } else {
                    add = 0.;
                    for (int t = 0; t < n; t++)
                        add += ec[base + off[t]];
                    add *= scale;
                }
                ef[((i64)i * njf + j) * nkf + k] += add;
            }
}

void orc_coarse_matrix(double *A, int ni, int nj, int nk, double h)
{
    /* mg_3d.h:147-273: identity rows on the boundary, 7-point rows scaled by
     * 1/h^2 in the interior.  A must be zero-filled by the caller (calloc). */
    const double invHsq = 1. / (h * h);
    const double one = 1. * invHsq, six = 6. * invHsq;
    const i64 n = (i64)ni * nj * nk;
    const i64 sj = nk, si = (i64)nj * nk;
    for (int i = 0; i < ni; i++)
        for (int j = 0; j < nj; j++)
            for (int k = 0; k < nk; k++) {
                i64 p = IDX(i, j, k);
                double *row = A + p * n;
                if (i == 0 || i == ni - 1 || j == 0 || j == nj - 1 || k == 0 ||
                    k == nk - 1) {
                    row[p] = 1.;
                } else {
                    row[p - si] = one; row[p + si] = one;
                    row[p - sj] = one; row[p + sj] = one;
                    row[p - 1] = one;  row[p + 1] = one;
                    row[p] = -six;
                }
            }
}

void orc_lu_factor(double *a, int n)
{
    /* gauss_elim.h:9-29: Doolittle, no pivoting, multiplier = a_ki * (1/a_ii) */
    for (int p = 0; p < n - 1; p++) {
        const double *prow = a + (i64)p * n;
        const double pinv = 1. / prow[p];
        for (int r = p + 1; r < n; r++) {
            double *row = a + (i64)r * n;
            const double z = row[p] * pinv;
            row[p] = z;
            for (int c = p + 1; c < n; c++)
                row[c] -= z * prow[c];
        }
    }
}

void orc_lu_solve(const double *lu, int n, const double *b, double *x)
{
    /* gauss_elim.h:31-60: forward sum over ascending j, backward sum over
     * DESCENDING j, each from 0. */
    for (int r = 0; r < n; r++) {
        const double *row = lu + (i64)r * n;
        double s = 0.;
        for (int c = 0; c < r; c++)
            s += row[c] * x[c];
        x[r] = b[r] - s;
    }
    for (int r = n - 1; r >= 0; r--) {
        const double *row = lu + (i64)r * n;
        double s = 0.;
        for (int c = n - 1; c > r; c--)
            s += row[c] * x[c];
        x[r] = (x[r] - s) / row[r];
    }
}

int orc_write_vtk(const char *path, const double *grid, int ni, int nj, int nk, double h)
{
    /* writeOutputData (postprocess.h:5-47) for a box: header (13-19), one line of
     * coordinates h*i, h*j, h*k per point with k fastest (22-33), the POINT_DATA header
     * (36-41), one value per line (42-44); every number through "%10.8e" */
    FILE *f = fopen(path, "w");
    if (!f)
        return 1;
    const long total = (long)ni * nj * nk;
    fprintf(f, "# vtk DataFile Version 2.0\nPotential data\nASCII\nDATASET STRUCTURED_GRID\n"
               "DIMENSIONS %d %d %d\nPOINTS %d float\n", ni, nj, nk, (int)total);
    for (int i = 0; i < ni; i++) {
        double x = h * i;
        for (int j = 0; j < nj; j++) {
            double y = h * j;
            for (int k = 0; k < nk; k++) {
                double z = h * k;
                fprintf(f, "%10.8e %10.8e %10.8e\n", x, y, z);
            }
        }
    }
    fprintf(f, "\nPOINT_DATA %d\nSCALARS data float 1\nLOOKUP_TABLE default\n", (int)total);
    for (long c = 0; c < total; c++)
        fprintf(f, "%10.8e\n", grid[c]);
    return fclose(f) != 0;
}

double orc_l2norm(const double *d, long n)
{
    /* mg_3d.h:783-792 */
    double s = 0.;
    for (long t = 0; t < n; t++)
        s += d[t] * d[t];
    return sqrt(s);
}

/* ------------------------------------------------------------------ */

struct orc_mg {
    int levels, gs;
    int *ni, *nj, *nk;
    double **u, **d, **r;
    double *lu;
    double h; /* finest spacing */
};

orc_mg *orc_mg_create(int ci, int cj, int ck, int levels, int gs)
{
    /* mg_3d.h:107-144 (level l has (c-1)*2^l+1 points per axis, level 0 is
     * the coarsest) and 275-293 (coarse operator built with the COARSE
     * spacing h*2^(L-1), factorised once). */
    orc_mg *m = calloc(1, sizeof *m);
    m->levels = levels;
    m->gs = gs;
    m->ni = malloc(sizeof(int) * levels);
    m->nj = malloc(sizeof(int) * levels);
    m->nk = malloc(sizeof(int) * levels);
    m->u = malloc(sizeof(double *) * levels);
    m->d = malloc(sizeof(double *) * levels);
    m->r = malloc(sizeof(double *) * levels);
    for (int l = 0; l < levels; l++) {
        m->ni[l] = (ci - 1) * (1 << l) + 1;
        m->nj[l] = (cj - 1) * (1 << l) + 1;
        m->nk[l] = (ck - 1) * (1 << l) + 1;
        size_t n = (size_t)m->ni[l] * m->nj[l] * m->nk[l];
        m->u[l] = calloc(n, sizeof(double));
        m->d[l] = calloc(n, sizeof(double));
        m->r[l] = calloc(n, sizeof(double));
    }
    m->h = 1. / (m->nk[levels - 1] - 1);
    size_t nc = (size_t)ci * cj * ck;
    m->lu = calloc(nc * nc, sizeof(double));
    double hc = m->h * (1 << (levels - 1));
    orc_coarse_matrix(m->lu, ci, cj, ck, hc);
    orc_lu_factor(m->lu, (int)nc);
    return m;
}

void orc_mg_destroy(orc_mg *m)
{
    if (!m)
        return;
    for (int l = 0; l < m->levels; l++) {
        free(m->u[l]); free(m->d[l]); free(m->r[l]);
    }
    free(m->u); free(m->d); free(m->r);
    free(m->ni); free(m->nj); free(m->nk);
    free(m->lu);
    free(m);
}

void orc_mg_dims(const orc_mg *m, int level, int *ni, int *nj, int *nk)
{
    *ni = m->ni[level]; *nj = m->nj[level]; *nk = m->nk[level];
}
double *orc_mg_u(orc_mg *m, int level) { return m->u[level]; }
double *orc_mg_d(orc_mg *m, int level) { return m->d[level]; }
double *orc_mg_r(orc_mg *m, int level) { return m->r[level]; }
double orc_mg_h(const orc_mg *m) { return m->h; }

static double cycle_level(orc_mg *m, int q, double h)
{
    /* mg_3d.h:1242-1362 */
    const int ni = m->ni[q], nj = m->nj[q], nk = m->nk[q];
    const size_t n = (size_t)ni * nj * nk;
    if (q < m->levels - 1)
        memset(m->u[q], 0, n * sizeof(double)); /* 1254-1260 */
    if (q == 0) {
        orc_lu_solve(m->lu, (int)n, m->d[0], m->u[0]); /* 1262-1277 */
        return 0.;
    }
    orc_smooth(m->u[q], m->d[q], ni, nj, nk, h, m->gs, 1);       /* 1282 */
    orc_residual(m->u[q], m->d[q], ni, nj, nk, h, m->r[q]);      /* 1294 */
    orc_restrict(m->r[q], ni, nj, nk, m->d[q - 1], m->ni[q - 1],
                 m->nj[q - 1], m->nk[q - 1]);                     /* 1310 */
    cycle_level(m, q - 1, 2 * h);                                 /* 1303,1320 */
    orc_prolong_correct(m->u[q - 1], m->ni[q - 1], m->nj[q - 1], m->nk[q - 1],
                        m->u[q], ni, nj, nk);                     /* 1331 */
    orc_smooth(m->u[q], m->d[q], ni, nj, nk, h, m->gs, 0);       /* 1341 */
    return orc_residual(m->u[q], m->d[q], ni, nj, nk, h, NULL);  /* 1354 */
}

double orc_mg_vcycle(orc_mg *m)
{
    return cycle_level(m, m->levels - 1, m->h);
}

void orc_mg_fmg_init(orc_mg *m)
{
    /* mg_3d.h:1364-1404 (commented there; live against an older vcycle in
     * mg_dirichlet_analytic.c:771-806), statement by statement, for boxes:
     * BCs into u[0] -- overwritten by the LU solve of d[0] -- then per level:
     * interpolate the coarser solution into u[l], BCs onto its faces, zero
     * u[l-1], one V-cycle entered at level l (which zeroes u[l] again unless l
     * is the finest level, 1254-1260). */
    double h = m->h * (1 << (m->levels - 1)); /* = GRID_LENGTH/(coarseGridNum-1) */
    orc_set_dirichlet(m->u[0], m->ni[0], m->nj[0], m->nk[0], h);
    orc_lu_solve(m->lu, m->ni[0] * m->nj[0] * m->nk[0], m->d[0], m->u[0]);
    for (int l = 1; l < m->levels; l++) {
        h = h * 0.5;
        orc_prolong_correct(m->u[l - 1], m->ni[l - 1], m->nj[l - 1], m->nk[l - 1], m->u[l],
                            m->ni[l], m->nj[l], m->nk[l]);
        orc_set_dirichlet(m->u[l], m->ni[l], m->nj[l], m->nk[l], h);
        memset(m->u[l - 1], 0,
               sizeof(double) * (size_t)m->ni[l - 1] * m->nj[l - 1] * m->nk[l - 1]);
        cycle_level(m, l, h);
    }
}

/* test_mg_3d.c:17-29 on the finest level; returns ||d|| */
double orc_mg_setup_problem(orc_mg *m)
{
    const int L = m->levels - 1;
    const int ni = m->ni[L], nj = m->nj[L], nk = m->nk[L];
    orc_set_dirichlet(m->d[L], ni, nj, nk, m->h);
    const double init = orc_l2norm(m->d[L], (long)ni * nj * nk);
    orc_set_dirichlet(m->u[L], ni, nj, nk, m->h);
    return init;
}

int orc_mg_solve(orc_mg *m, double tol, int max_cycles, double *history,
                 double *init_norm)
{
    /* test_mg_3d.c:17-67: BCs into rhs faces, ||rhs|| as the reference norm,
     * BCs into the solution faces, cycle while norm > tol*||rhs||. */
    const int L = m->levels - 1;
    const int ni = m->ni[L], nj = m->nj[L], nk = m->nk[L];
    orc_set_dirichlet(m->d[L], ni, nj, nk, m->h);
    const double init = orc_l2norm(m->d[L], (long)ni * nj * nk);
    orc_set_dirichlet(m->u[L], ni, nj, nk, m->h);
    if (init_norm)
        *init_norm = init;
    const double cmp = init * tol;
    double norm = 1e9;
    int c = 0;
    while (norm > cmp && c < max_cycles) {
        /* the driver squares and re-roots the (one-thread) partial:
         * test_mg_3d.c:52-59 */
        double part = orc_mg_vcycle(m);
        norm = sqrt(0. + part * part);
        if (history)
            history[c] = norm;
        c++;
    }
    return c;
}
