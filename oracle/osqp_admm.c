/* ORACLE — test infrastructure only.  Never linked into the product library.
 *
 * Plain-C restatement of the OSQP ADMM algorithm (same algorithm and constants as
 * oracle/osqp_admm.py, see that file's header for provenance: OSQP is a third-party
 * dependency of the reference that is absent from /root/reference and not
 * installable here; PARITY UNPINNED against an OSQP binary).
 *
 * Reference call sites this stands in for:
 *   /root/reference/Control/MPC/mpc_kinematics.py:205-211   prob.setup(...); prob.solve()
 *   /root/reference/Control/MPC/mpc_dynamics.py:248-252, 398-402
 *   /root/reference/vehicle_lateral_mpc_slack_increment.py:118-121, 237-248
 *
 * It is a general sparse-QP solver (CSC P upper triangle, CSC A), written
 * independently of the numpy version so that the two cross-check each other:
 *   - Ruiz equilibration + cost scaling on the CSC arrays (scaling.c: scale_data)
 *   - rho per constraint type on the scaled bounds (auxil.c: set_rho_vec)
 *   - the KKT solve is done on the reduced system (P + sigma I + A' rho A) x = r,
 *     z~ = A x, which is algebraically the quasi-definite KKT system OSQP factors
 *     with QDLDL (paper eq. (17)); factored ONCE (cached) as a banded Cholesky
 *     under a caller-supplied fill-reducing permutation (OSQP uses AMD)
 *   - x/z/y updates with relaxation alpha, projection onto [l,u]
 *   - unscaled residuals / tolerances / infeasibility tests every
 *     check_termination iterations (auxil.c: check_termination)
 *
 * oracle_solve_batch() runs one independent solve per QP under OpenMP: this is
 * the "reference's per-QP OSQP loop on host cores" that bench.py's cpu_baseline
 * leg and `bench.py --impl reference` time.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OSQP_INFTY 1e30
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define RHO_TOL 1e-4
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4

enum { ST_SOLVED = 1, ST_SOLVED_INACC = 2, ST_PINF_INACC = 3, ST_DINF_INACC = 4,
       ST_MAX_ITER = -2, ST_PINF = -3, ST_DINF = -4, ST_UNSOLVED = -10 };

typedef struct {
    double rho, sigma, alpha, eps_abs, eps_rel, eps_prim_inf, eps_dual_inf;
    int max_iter, scaling, check_termination, warm_start;
} OracleSettings;

typedef struct {
    int iter, status;
    double pri_res, dua_res;
} OracleInfo;

static double limit_scaling(double v) {
    v = v < MIN_SCALING ? 1.0 : v;
    return v > MAX_SCALING ? MAX_SCALING : v;
}
static double norm_inf(const double *v, int n) {
    double r = 0; for (int i = 0; i < n; i++) { double a = fabs(v[i]); if (a > r) r = a; } return r;
}
static double norm_inf_scaled(const double *s, const double *v, int n) {
    double r = 0; for (int i = 0; i < n; i++) { double a = fabs(s[i] * v[i]); if (a > r) r = a; } return r;
}

typedef struct {
    int n, m, bw;
    const int *Pp, *Pi, *Ap, *Ai;
    double *Px, *Ax, *q, *l, *u;           /* scaled copies */
    double *D, *E, *Dinv, *Einv, c, cinv;
    double *rho, *rho_inv;
    int *ctype;
    const int *perm;                        /* perm[j] = position of variable j */
    double *Lb;                             /* banded Cholesky, n x (bw+1), Lb[j*(bw+1)+d] = L[j+d][j] */
    double *x, *z, *y, *xp, *zp, *xt, *zt, *w, *dx, *dy, *Ax_, *Px_, *Aty;
    OracleSettings s;
} Work;

static void A_mul(const Work *k, const double *x, double *out) {           /* out = A x */
    memset(out, 0, sizeof(double) * k->m);
    for (int j = 0; j < k->n; j++)
        for (int p = k->Ap[j]; p < k->Ap[j + 1]; p++) out[k->Ai[p]] += k->Ax[p] * x[j];
}
static void At_mul(const Work *k, const double *y, double *out) {          /* out = A' y */
    for (int j = 0; j < k->n; j++) {
        double s = 0;
        for (int p = k->Ap[j]; p < k->Ap[j + 1]; p++) s += k->Ax[p] * y[k->Ai[p]];
        out[j] = s;
    }
}
static void P_mul(const Work *k, const double *x, double *out) {           /* out = P x, P sym upper */
    memset(out, 0, sizeof(double) * k->n);
    for (int j = 0; j < k->n; j++)
        for (int p = k->Pp[j]; p < k->Pp[j + 1]; p++) {
            int i = k->Pi[p];
            out[i] += k->Px[p] * x[j];
            if (i != j) out[j] += k->Px[p] * x[i];
        }
}

static void scale_data(Work *k) {
    int n = k->n, m = k->m;
    double *Dt = (double *)malloc(sizeof(double) * n), *Et = (double *)malloc(sizeof(double) * m);
    for (int i = 0; i < n; i++) k->D[i] = 1.0;
    for (int i = 0; i < m; i++) k->E[i] = 1.0;
    k->c = 1.0;
    for (int it = 0; it < k->s.scaling; it++) {
        for (int i = 0; i < n; i++) Dt[i] = 0;
        for (int i = 0; i < m; i++) Et[i] = 0;
        for (int j = 0; j < n; j++) {
            for (int p = k->Pp[j]; p < k->Pp[j + 1]; p++) {
                double a = fabs(k->Px[p]); int i = k->Pi[p];
                if (a > Dt[j]) Dt[j] = a;
                if (a > Dt[i]) Dt[i] = a;
            }
            for (int p = k->Ap[j]; p < k->Ap[j + 1]; p++) {
                double a = fabs(k->Ax[p]); int i = k->Ai[p];
                if (a > Dt[j]) Dt[j] = a;
                if (a > Et[i]) Et[i] = a;
            }
        }
        for (int i = 0; i < n; i++) Dt[i] = 1.0 / sqrt(limit_scaling(Dt[i]));
        for (int i = 0; i < m; i++) Et[i] = 1.0 / sqrt(limit_scaling(Et[i]));
        for (int j = 0; j < n; j++) {
            for (int p = k->Pp[j]; p < k->Pp[j + 1]; p++) k->Px[p] *= Dt[k->Pi[p]] * Dt[j];
            for (int p = k->Ap[j]; p < k->Ap[j + 1]; p++) k->Ax[p] *= Et[k->Ai[p]] * Dt[j];
            k->q[j] *= Dt[j];
            k->D[j] *= Dt[j];
        }
        for (int i = 0; i < m; i++) k->E[i] *= Et[i];
        /* cost normalisation */
        for (int i = 0; i < n; i++) Dt[i] = 0;
        for (int j = 0; j < n; j++)
            for (int p = k->Pp[j]; p < k->Pp[j + 1]; p++) {
                double a = fabs(k->Px[p]); int i = k->Pi[p];
                if (a > Dt[j]) Dt[j] = a;
                if (a > Dt[i]) Dt[i] = a;
            }
        double c_temp = 0;
        for (int i = 0; i < n; i++) c_temp += Dt[i];
        c_temp /= n;
        double nq = limit_scaling(norm_inf(k->q, n));
        c_temp = c_temp > nq ? c_temp : nq;
        c_temp = 1.0 / limit_scaling(c_temp);
        for (int p = 0; p < k->Pp[n]; p++) k->Px[p] *= c_temp;
        for (int i = 0; i < n; i++) k->q[i] *= c_temp;
        k->c *= c_temp;
    }
    k->cinv = 1.0 / k->c;
    for (int i = 0; i < n; i++) k->Dinv[i] = 1.0 / k->D[i];
    for (int i = 0; i < m; i++) { k->Einv[i] = 1.0 / k->E[i]; k->l[i] *= k->E[i]; k->u[i] *= k->E[i]; }
    free(Dt); free(Et);
}

static void set_rho_vec(Work *k) {
    for (int i = 0; i < k->m; i++) {
        if (k->l[i] < -OSQP_INFTY * MIN_SCALING && k->u[i] > OSQP_INFTY * MIN_SCALING) {
            k->ctype[i] = -1; k->rho[i] = RHO_MIN;
        } else if (k->u[i] - k->l[i] < RHO_TOL) {
            k->ctype[i] = 1; k->rho[i] = RHO_EQ_OVER_RHO_INEQ * k->s.rho;
        } else {
            k->ctype[i] = 0; k->rho[i] = k->s.rho;
        }
        k->rho_inv[i] = 1.0 / k->rho[i];
    }
}

/* reduced KKT: M = P + sigma I + A' diag(rho) A in permuted band storage, then Cholesky */
static int factor(Work *k) {
    int n = k->n, m = k->m;
    /* CSR of A (row -> list of (col, val)) */
    int *rp = (int *)calloc(m + 1, sizeof(int));
    int nnz = k->Ap[n];
    int *rc = (int *)malloc(sizeof(int) * (nnz ? nnz : 1));
    double *rv = (double *)malloc(sizeof(double) * (nnz ? nnz : 1));
    for (int p = 0; p < nnz; p++) rp[k->Ai[p] + 1]++;
    for (int i = 0; i < m; i++) rp[i + 1] += rp[i];
    int *fill = (int *)malloc(sizeof(int) * (m ? m : 1));
    memcpy(fill, rp, sizeof(int) * m);
    for (int j = 0; j < n; j++)
        for (int p = k->Ap[j]; p < k->Ap[j + 1]; p++) { int r = k->Ai[p]; rc[fill[r]] = j; rv[fill[r]] = k->Ax[p]; fill[r]++; }
    int bw = 0;
    for (int r = 0; r < m; r++)
        for (int a = rp[r]; a < rp[r + 1]; a++)
            for (int b = rp[r]; b < rp[r + 1]; b++) {
                int d = k->perm[rc[a]] - k->perm[rc[b]]; if (d > bw) bw = d;
            }
    for (int j = 0; j < n; j++)
        for (int p = k->Pp[j]; p < k->Pp[j + 1]; p++) {
            int d = abs(k->perm[k->Pi[p]] - k->perm[j]); if (d > bw) bw = d;
        }
    k->bw = bw;
    int ld = bw + 1;
    free(k->Lb);
    k->Lb = (double *)calloc((size_t)n * ld, sizeof(double));
    double *L = k->Lb;
    for (int j = 0; j < n; j++) {
        L[(size_t)k->perm[j] * ld] += k->s.sigma;
        for (int p = k->Pp[j]; p < k->Pp[j + 1]; p++) {
            int a = k->perm[k->Pi[p]], b = k->perm[j];
            int lo = a < b ? a : b, hi = a < b ? b : a;
            L[(size_t)lo * ld + (hi - lo)] += k->Px[p];
        }
    }
    for (int r = 0; r < m; r++)
        for (int a = rp[r]; a < rp[r + 1]; a++)
            for (int b = rp[r]; b < rp[r + 1]; b++) {
                int pa = k->perm[rc[a]], pb = k->perm[rc[b]];
                if (pa >= pb) L[(size_t)pb * ld + (pa - pb)] += k->rho[r] * rv[a] * rv[b];
            }
    /* band Cholesky (right-looking) */
    for (int j = 0; j < n; j++) {
        double d = L[(size_t)j * ld];
        if (!(d > 0)) { free(rp); free(rc); free(rv); free(fill); return -1; }
        d = sqrt(d);
        L[(size_t)j * ld] = d;
        int lim = (n - 1 - j) < bw ? (n - 1 - j) : bw;
        for (int i = 1; i <= lim; i++) L[(size_t)j * ld + i] /= d;
        for (int c = 1; c <= lim; c++) {
            double lc = L[(size_t)j * ld + c];
            if (lc == 0.0) continue;
            for (int i = c; i <= lim; i++) L[(size_t)(j + c) * ld + (i - c)] -= L[(size_t)j * ld + i] * lc;
        }
    }
    free(rp); free(rc); free(rv); free(fill);
    return 0;
}

static void chol_solve(const Work *k, double *b) {        /* b in permuted order, in place */
    int n = k->n, bw = k->bw, ld = bw + 1;
    const double *L = k->Lb;
    for (int j = 0; j < n; j++) {
        b[j] /= L[(size_t)j * ld];
        int lim = (n - 1 - j) < bw ? (n - 1 - j) : bw;
        double bj = b[j];
        for (int i = 1; i <= lim; i++) b[j + i] -= L[(size_t)j * ld + i] * bj;
    }
    for (int j = n - 1; j >= 0; j--) {
        int lim = (n - 1 - j) < bw ? (n - 1 - j) : bw;
        double s = b[j];
        for (int i = 1; i <= lim; i++) s -= L[(size_t)j * ld + i] * b[j + i];
        b[j] = s / L[(size_t)j * ld];
    }
}

static void iterate(Work *k) {
    int n = k->n, m = k->m;
    double alpha = k->s.alpha, sigma = k->s.sigma;
    memcpy(k->xp, k->x, sizeof(double) * n);
    memcpy(k->zp, k->z, sizeof(double) * m);
    for (int i = 0; i < m; i++) k->zt[i] = k->rho[i] * k->zp[i] - k->y[i];
    At_mul(k, k->zt, k->xt);
    for (int j = 0; j < n; j++) k->w[k->perm[j]] = sigma * k->xp[j] - k->q[j] + k->xt[j];
    chol_solve(k, k->w);
    for (int j = 0; j < n; j++) k->xt[j] = k->w[k->perm[j]];
    A_mul(k, k->xt, k->zt);
    for (int j = 0; j < n; j++) { k->x[j] = alpha * k->xt[j] + (1.0 - alpha) * k->xp[j]; k->dx[j] = k->x[j] - k->xp[j]; }
    for (int i = 0; i < m; i++) {
        double zr = alpha * k->zt[i] + (1.0 - alpha) * k->zp[i];
        double v = zr + k->rho_inv[i] * k->y[i];
        v = v < k->l[i] ? k->l[i] : v; v = v > k->u[i] ? k->u[i] : v;
        k->z[i] = v;
        k->dy[i] = k->rho[i] * (zr - v);
        k->y[i] += k->dy[i];
    }
}

static int is_primal_infeasible(Work *k, double eps) {
    int m = k->m, n = k->n;
    double *dy = k->zt;                      /* scratch */
    for (int i = 0; i < m; i++) {
        double v = k->dy[i];
        int up_inf = k->u[i] > OSQP_INFTY * MIN_SCALING, lo_inf = k->l[i] < -OSQP_INFTY * MIN_SCALING;
        if (up_inf && lo_inf) v = 0; else if (up_inf) v = v < 0 ? v : 0; else if (lo_inf) v = v > 0 ? v : 0;
        dy[i] = v;
    }
    double nd = norm_inf_scaled(k->E, dy, m);
    if (nd > eps) {
        double lhs = 0;
        for (int i = 0; i < m; i++) lhs += k->u[i] * (dy[i] > 0 ? dy[i] : 0) + k->l[i] * (dy[i] < 0 ? dy[i] : 0);
        if (lhs < -eps * nd) {
            At_mul(k, dy, k->w);
            return norm_inf_scaled(k->Dinv, k->w, n) < eps * nd;
        }
    }
    return 0;
}
static int is_dual_infeasible(Work *k, double eps) {
    int m = k->m, n = k->n;
    double nd = norm_inf_scaled(k->D, k->dx, n);
    if (nd > eps) {
        double qdx = 0; for (int j = 0; j < n; j++) qdx += k->q[j] * k->dx[j];
        if (qdx < -k->c * eps * nd) {
            P_mul(k, k->dx, k->w);
            if (norm_inf_scaled(k->Dinv, k->w, n) < k->c * eps * nd) {
                A_mul(k, k->dx, k->zt);
                for (int i = 0; i < m; i++) {
                    double a = k->Einv[i] * k->zt[i];
                    if ((k->u[i] < OSQP_INFTY * MIN_SCALING && a > eps * nd) ||
                        (k->l[i] > -OSQP_INFTY * MIN_SCALING && a < -eps * nd)) return 0;
                }
                return 1;
            }
        }
    }
    return 0;
}

static int check_termination(Work *k, int approximate, OracleInfo *info) {
    int n = k->n, m = k->m;
    double ea = k->s.eps_abs, er = k->s.eps_rel, ep = k->s.eps_prim_inf, ed = k->s.eps_dual_inf;
    if (approximate) { ea *= 10; er *= 10; ep *= 10; ed *= 10; }
    A_mul(k, k->x, k->Ax_); P_mul(k, k->x, k->Px_); At_mul(k, k->y, k->Aty);
    double pri = 0, dua = 0;
    for (int i = 0; i < m; i++) { double a = fabs(k->Einv[i] * (k->Ax_[i] - k->z[i])); if (a > pri) pri = a; }
    for (int j = 0; j < n; j++) { double a = fabs(k->Dinv[j] * (k->q[j] + k->Aty[j] + k->Px_[j])); if (a > dua) dua = a; }
    dua *= k->cinv;
    info->pri_res = pri; info->dua_res = dua;
    int prim_ok = 1, dual_ok;
    if (m) {
        double a = norm_inf_scaled(k->Einv, k->z, m), b = norm_inf_scaled(k->Einv, k->Ax_, m);
        double eps_prim = ea + er * (a > b ? a : b);
        prim_ok = pri < eps_prim;
        if (!prim_ok && is_primal_infeasible(k, ep)) return approximate ? ST_PINF_INACC : ST_PINF;
    }
    double a = norm_inf_scaled(k->Dinv, k->q, n), b = norm_inf_scaled(k->Dinv, k->Aty, n), c = norm_inf_scaled(k->Dinv, k->Px_, n);
    double mx = a > b ? a : b; mx = mx > c ? mx : c;
    double eps_dual = ea + er * k->cinv * mx;
    dual_ok = dua < eps_dual;
    if (!dual_ok && is_dual_infeasible(k, ed)) return approximate ? ST_DINF_INACC : ST_DINF;
    if (prim_ok && dual_ok) return approximate ? ST_SOLVED_INACC : ST_SOLVED;
    return 0;
}

static double *dalloc(int n) { return (double *)calloc(n > 0 ? n : 1, sizeof(double)); }

/* prob.setup(): copy + scale the data, rho per constraint type, cached factorisation; x = z = y = 0.
 * Returns 0, or -1 if the reduced KKT matrix is not positive definite. */
static int work_setup(Work *kp, int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                      const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                      const int *perm, const OracleSettings *s, int **ident_out) {
    Work k; memset(&k, 0, sizeof(k));
    k.n = n; k.m = m; k.Pp = Pp; k.Pi = Pi; k.Ap = Ap; k.Ai = Ai; k.s = *s;
    k.s.rho = k.s.rho < RHO_MIN ? RHO_MIN : (k.s.rho > RHO_MAX ? RHO_MAX : k.s.rho);
    *ident_out = NULL;
    if (!perm) { int *ident = (int *)malloc(sizeof(int) * n); for (int i = 0; i < n; i++) ident[i] = i; perm = ident; *ident_out = ident; }
    k.perm = perm;
    int pnz = Pp[n], anz = Ap[n];
    k.Px = dalloc(pnz); memcpy(k.Px, Px, sizeof(double) * pnz);
    k.Ax = dalloc(anz); memcpy(k.Ax, Ax, sizeof(double) * anz);
    k.q = dalloc(n); memcpy(k.q, q, sizeof(double) * n);
    k.l = dalloc(m); k.u = dalloc(m);
    for (int i = 0; i < m; i++) { k.l[i] = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i]; k.u[i] = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i]; }
    k.D = dalloc(n); k.Dinv = dalloc(n); k.E = dalloc(m); k.Einv = dalloc(m);
    k.rho = dalloc(m); k.rho_inv = dalloc(m); k.ctype = (int *)calloc(m > 0 ? m : 1, sizeof(int));
    k.x = dalloc(n); k.xp = dalloc(n); k.xt = dalloc(n); k.w = dalloc(n); k.dx = dalloc(n); k.Px_ = dalloc(n); k.Aty = dalloc(n);
    k.z = dalloc(m); k.zp = dalloc(m); k.zt = dalloc(m); k.y = dalloc(m); k.dy = dalloc(m); k.Ax_ = dalloc(m);
    scale_data(&k);
    set_rho_vec(&k);
    int rc = factor(&k);
    *kp = k;
    return rc;
}
static void work_free(Work *k, int *ident) {
    free(k->Px); free(k->Ax); free(k->q); free(k->l); free(k->u); free(k->D); free(k->Dinv); free(k->E); free(k->Einv);
    free(k->rho); free(k->rho_inv); free(k->ctype); free(k->x); free(k->xp); free(k->xt); free(k->w); free(k->dx);
    free(k->Px_); free(k->Aty); free(k->z); free(k->zp); free(k->zt); free(k->y); free(k->dy); free(k->Ax_); free(k->Lb);
    free(ident);
}
/* prob.solve(): the ADMM loop from the current (x, z, y) — zeros first unless warm_start (osqp.c: osqp_solve) */
static void work_solve(Work *k, OracleInfo *info) {
    if (!k->s.warm_start) {
        memset(k->x, 0, sizeof(double) * k->n); memset(k->z, 0, sizeof(double) * k->m); memset(k->y, 0, sizeof(double) * k->m);
    }
    int it, checked = 0, st = 0;
    for (it = 1; it <= k->s.max_iter; it++) {
        iterate(k);
        checked = 0;
        if (k->s.check_termination && it % k->s.check_termination == 0) {
            checked = 1;
            st = check_termination(k, 0, info);
            if (st) break;
        }
    }
    if (it > k->s.max_iter) it = k->s.max_iter;
    if (!st && !checked) st = check_termination(k, 0, info);
    if (!st) { st = check_termination(k, 1, info); if (!st) st = ST_MAX_ITER; }
    info->iter = it; info->status = st;
}
/* prob.update(l=, u=) (osqp.c: osqp_update_bounds): scale with the existing E, re-evaluate the rho vector and refactor
 * when a constraint changed type (auxil.c: update_rho_vec).  Returns 1 if it refactored. */
static int work_update_bounds(Work *k, const double *l, const double *u) {
    int m = k->m, changed = 0;
    int *old = (int *)malloc(sizeof(int) * (m ? m : 1));
    memcpy(old, k->ctype, sizeof(int) * m);
    for (int i = 0; i < m; i++) {
        double lo = l[i] < -OSQP_INFTY ? -OSQP_INFTY : l[i], hi = u[i] > OSQP_INFTY ? OSQP_INFTY : u[i];
        k->l[i] = k->E[i] * lo; k->u[i] = k->E[i] * hi;
    }
    set_rho_vec(k);
    for (int i = 0; i < m; i++) if (old[i] != k->ctype[i]) changed = 1;
    free(old);
    if (changed) factor(k);
    return changed;
}

/* One QP: setup (scale, rho, factor) + solve.  x0/y0 (unscaled) are an optional warm start.
 * Returns 0, or -1 if the reduced KKT matrix is not positive definite. */
int oracle_osqp_solve(int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                      const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                      const int *perm, const OracleSettings *s, const double *x0, const double *y0,
                      double *x_out, double *y_out, OracleInfo *info) {
    Work k; int *ident;
    int rc = work_setup(&k, n, m, Pp, Pi, Px, q, Ap, Ai, Ax, l, u, perm, s, &ident);
    info->iter = 0; info->status = ST_UNSOLVED; info->pri_res = info->dua_res = NAN;
    if (rc == 0) {
        k.s.warm_start = 1;                      /* (x, z, y) are zeros or the caller's warm start) */
        if (x0) { for (int j = 0; j < n; j++) k.x[j] = k.Dinv[j] * x0[j]; A_mul(&k, k.x, k.z); }
        if (y0) for (int i = 0; i < m; i++) k.y[i] = k.Einv[i] * y0[i] * k.c;
        work_solve(&k, info);
        for (int j = 0; j < n; j++) x_out[j] = k.D[j] * k.x[j];
        for (int i = 0; i < m; i++) y_out[i] = k.cinv * k.E[i] * k.y[i];
    }
    work_free(&k, ident);
    return rc;
}

/* B QPs sharing one sparsity pattern; per-QP values are contiguous rows:
 * Px[B][pnz], q[B][n], Ax[B][anz], l[B][m], u[B][m]; outputs x[B][n], y[B][m], iter[B], status[B].
 * nthreads <= 0 -> all cores.  Returns the number of threads used. */
int oracle_solve_batch(int B, int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                       const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                       const int *perm, const OracleSettings *s, double *x_out, double *y_out,
                       int *iter_out, int *status_out, int nthreads) {
    int pnz = Pp[n], anz = Ap[n], used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 4)
#endif
    for (int b = 0; b < B; b++) {
        OracleInfo info;
        oracle_osqp_solve(n, m, Pp, Pi, Px + (size_t)b * pnz, q + (size_t)b * n, Ap, Ai, Ax + (size_t)b * anz,
                          l + (size_t)b * m, u + (size_t)b * m, perm, s, NULL, NULL,
                          x_out + (size_t)b * n, y_out + (size_t)b * m, &info);
        iter_out[b] = info.iter; status_out[b] = info.status;
    }
    return used;
}

/* The closed loop of vehicle_lateral_mpc_slack_increment.py:123-269 for B independent scenarios (OpenMP over scenarios):
 * prob.setup() once, then `steps` times { prob.update(l, u) with the initial-state rows -x0 (rows 0..nx-1 of l, u; the
 * inequality bounds of step k come from the optional schedule l_sched/u_sched [nsched][m], switched in at sched_step[]);
 * warm-started prob.solve(); apply the first input: x0 <- Apl x0 + Bpl u0 }.
 * Apl [B][nx*nx], Bpl [B][nx*nu] row-major plant matrices, u_index = position of u_0 in the QP's variable vector.
 * Outputs: iters [B][steps], status [B][steps], u_applied [B][steps][nu], traj [B][steps+1][nx].
 * Returns the number of threads used. */
int oracle_closed_loop_batch(int B, int n, int m, const int *Pp, const int *Pi, const double *Px, const double *q,
                             const int *Ap, const int *Ai, const double *Ax, const double *l, const double *u,
                             const int *perm, const OracleSettings *s, int steps, int nx, int nu, int u_index,
                             const double *Apl, const double *Bpl, const double *x0,
                             int nsched, const int *sched_step, const double *l_sched, const double *u_sched,
                             int *iters, int *status, double *u_applied, double *traj, int nthreads) {
    int pnz = Pp[n], anz = Ap[n], used = 1;
#ifdef _OPENMP
    if (nthreads > 0) omp_set_num_threads(nthreads);
    used = omp_get_max_threads();
#pragma omp parallel for schedule(dynamic, 1)
#endif
    for (int b = 0; b < B; b++) {
        Work k; int *ident;
        double *lb = (double *)malloc(sizeof(double) * m), *ub = (double *)malloc(sizeof(double) * m);
        double *x = (double *)malloc(sizeof(double) * nx), *xn = (double *)malloc(sizeof(double) * nx);
        memcpy(lb, l + (size_t)b * m, sizeof(double) * m); memcpy(ub, u + (size_t)b * m, sizeof(double) * m);
        memcpy(x, x0 + (size_t)b * nx, sizeof(double) * nx);
        int rc = work_setup(&k, n, m, Pp, Pi, Px + (size_t)b * pnz, q + (size_t)b * n, Ap, Ai, Ax + (size_t)b * anz, lb, ub,
                            perm, s, &ident);
        if (traj) memcpy(traj + (size_t)b * (steps + 1) * nx, x, sizeof(double) * nx);
        for (int t = 0; t < steps && rc == 0; t++) {
            int upd = t > 0;
            for (int j = 0; j < nsched; j++)
                if (sched_step[j] == t) {
                    memcpy(lb + nx, l_sched + (size_t)j * m + nx, sizeof(double) * (m - nx));
                    memcpy(ub + nx, u_sched + (size_t)j * m + nx, sizeof(double) * (m - nx));
                    upd = 1;
                }
            for (int i = 0; i < nx; i++) { lb[i] = -x[i]; ub[i] = -x[i]; }
            if (upd) work_update_bounds(&k, lb, ub);
            OracleInfo info;
            work_solve(&k, &info);
            iters[(size_t)b * steps + t] = info.iter; status[(size_t)b * steps + t] = info.status;
            const double *Ab = Apl + (size_t)b * nx * nx, *Bb = Bpl + (size_t)b * nx * nu;
            for (int i = 0; i < nx; i++) {
                double v = 0;
                for (int j = 0; j < nx; j++) v += Ab[i * nx + j] * x[j];
                for (int j = 0; j < nu; j++) v += Bb[i * nu + j] * (k.D[u_index + j] * k.x[u_index + j]);
                xn[i] = v;
            }
            for (int j = 0; j < nu; j++) if (u_applied) u_applied[((size_t)b * steps + t) * nu + j] = k.D[u_index + j] * k.x[u_index + j];
            memcpy(x, xn, sizeof(double) * nx);
            if (traj) memcpy(traj + ((size_t)b * (steps + 1) + t + 1) * nx, x, sizeof(double) * nx);
        }
        work_free(&k, ident);
        free(lb); free(ub); free(x); free(xn);
    }
    return used;
}
