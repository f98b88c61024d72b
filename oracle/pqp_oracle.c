/*
 * pqp_oracle.c — CPU oracle for the path-QP solve path.  TEST INFRASTRUCTURE ONLY
 * (see pqp_oracle.h for who may load it and for the "parity unpinned" statement).
 *
 * Two parts:
 *  (1) Assembly of the QP exactly as the reference builds it
 *        BaseSolver::BaseSolver      /root/reference/src/solver/base_solver.cpp:15-39
 *        BaseSolver::setCost         base_solver.cpp:119-148
 *        BaseSolver::setConstraints  base_solver.cpp:150-261
 *        BaseSolver::getSoftBounds   base_solver.cpp:290-296
 *      directly in CSC (the reference's dense m x nv temporaries are reproduced only as an
 *      optional cost model, assemble_dense_style()).
 *  (2) A restatement of the OSQP algorithm (OSQP 0.6.x behaviour):
 *      Ruiz equilibration, rho vector, quasi-definite KKT + sparse LDL' (up-looking,
 *      elimination-tree based — the published QDLDL/LDL algorithm), ADMM iteration,
 *      unscaled termination test, infeasibility certificates, adaptive rho, and the
 *      update_bounds / update_A warm re-solve path used by
 *      updateProblemFormulationAndSolve (base_solver.cpp:97-117).
 *
 * Stated deviations from a real OSQP run:
 *   - adaptive_rho_interval is a fixed number of iterations (default 25); OSQP's default
 *     derives it from wall-clock time (0.4 x setup time) and is not reproducible.
 *   - the fill-reducing ordering is a structure-derived elimination order instead of AMD
 *     (any symmetric permutation of a quasi-definite matrix gives the same solve up to
 *     rounding).
 *   - A keeps its full structural pattern (17n-5 entries) even when a coefficient is
 *     exactly 0, so the second solve always takes osqp-eigen's "same pattern" branch.
 */
#define _GNU_SOURCE
#include "pqp_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define OSQP_INFTY 1e30
#define MIN_SCALING 1e-4
#define MAX_SCALING 1e4
#define RHO_MIN 1e-6
#define RHO_MAX 1e6
#define RHO_EQ_OVER_RHO_INEQ 1e3
#define RHO_TOL 1e-4
#define UNKNOWN (-1)

static double dmax(double a, double b) { return a > b ? a : b; }
static double dmin(double a, double b) { return a < b ? a : b; }

struct pqo_ws {
    pqp_params prm;
    int n, p, nv, m, nnzA;
    /* inputs kept for re-linearisation */
    double *s, *kref, *lin_l, *lin_psi, *lin_k, *b0lb, *b0ub, *b1lb, *b1ub;
    double inst[PQP_NINST];
    /* unscaled data */
    int *Ap, *Ai;
    double *Ax0, *l0, *u0, *Pd0;
    /* scaled data */
    double *Ax, *l, *u, *Pd;
    double *D, *E, *Dinv, *Einv, c, cinv;
    double *D_temp, *D_temp_A, *E_temp;
    /* CSR view of A (indices into Ax) */
    int *Rp, *Rj, *Re;
    /* rho */
    double rho, *rho_vec, *rho_inv_vec;
    int *constr_type;
    /* KKT (permuted, upper CSC) */
    int nk, nnzK, *perm, *pinv, *Kp, *Ki, *Kkind, *Kidx;
    double *Kx;
    /* LDL */
    int *etree, *Lnz, *Lp, *Li, nnzL;
    double *Lx, *Dg, *Dginv;
    int *iwork;
    unsigned char *bwork;
    double *fwork, *bp;
    /* iterates (scaled) */
    double *x, *z, *y, *x_prev, *z_prev, *xz_tilde, *delta_x, *delta_y;
    double *Axv, *Px, *Aty, *Atdy, *Adx;
    /* info */
    int iter, status, rho_updates;
    double pri_res, dua_res, obj_val;
};

/* ------------------------------------------------------------------ index maps (App. A.1/A.3) */
static int var_state(int i, int k) { return 3 * i + k; }
static int var_ctrl(int n, int i) { return 3 * n + i; }
static int var_slack(int n, int p, int i, int j) { /* j: 0 front/center, 1 rear */
    if (i < p) return 4 * n - 1 + 2 * i + j;
    return 4 * n - 1 + 2 * p + (i - p);
}
static int row_dyn(int i, int k) { return 3 * i + k; }
static int row_kappa(int n, int i) { return 3 * n + i; }
static int row_coll(int n, int p, int i, int j) {
    if (i < p) return 4 * n + 2 * i + j;
    return 4 * n + 2 * p + (i - p);
}

/* base_solver.cpp:290-296 */
static void soft_bounds(double lb, double ub, double safety_margin, double *olb, double *oub) {
    const double clearance = ub - lb;
    const double min_clearance = 0.1;
    double remain = dmax(min_clearance, clearance - 2 * safety_margin);
    double shrink = dmax(0.0, (clearance - remain) / 2.0);
    *olb = lb + shrink;
    *oub = ub - shrink;
}

typedef struct { int r, c; double v; } trip;

/* Emit the entries of A and the bounds for the current linearisation point
 * (base_solver.cpp:150-261). trips must hold 17n-5 (p==n) entries at most 17n. */
static int emit_constraints(const pqo_ws *w, trip *t, double *lo, double *up) {
    const int n = w->n, p = w->p, m = w->m;
    const pqp_params *q = &w->prm;
    int cnt = 0;
    memset(lo, 0, sizeof(double) * m);
    memset(up, 0, sizeof(double) * m);
    /* transition part: -I on every state (:161-163) */
    for (int j = 0; j < 3 * n; ++j) { t[cnt].r = j; t[cnt].c = j; t[cnt].v = -1.0; ++cnt; }
    /* x0 (:216-220) */
    lo[0] = up[0] = -w->inst[PQP_I_L0];
    lo[1] = up[1] = -w->inst[PQP_I_PSI0];
    lo[2] = up[2] = -w->inst[PQP_I_K0];
    for (int i = 0; i + 1 < n; ++i) {
        const double xl = w->lin_l[i], xp = w->lin_psi[i], xk = w->lin_k[i];
        const double xk_next = w->lin_k[i + 1];
        const double ds = w->s[i + 1] - w->s[i];
        const double cp = cos(xp), tp = tan(xp);
        /* df_x (:169-171) */
        const double f00 = -xk * tp, f01 = (1 - xk * xl) / pow(cp, 2);
        const double f10 = -xk * xk / cp, f11 = (1 - xk * xl) * xk * tp / cp,
                     f12 = (1 - xk * xl) / cp;
        /* A = ds*df_x + I (:175); structural entries (0,0),(0,1),(1,0),(1,1),(1,2),(2,2) */
        const int r0 = row_dyn(i + 1, 0);
        t[cnt++] = (trip){r0, var_state(i, 0), ds * f00 + 1.0};
        t[cnt++] = (trip){r0, var_state(i, 1), ds * f01};
        t[cnt++] = (trip){r0 + 1, var_state(i, 0), ds * f10};
        t[cnt++] = (trip){r0 + 1, var_state(i, 1), ds * f11 + 1.0};
        t[cnt++] = (trip){r0 + 1, var_state(i, 2), ds * f12};
        t[cnt++] = (trip){r0 + 2, var_state(i, 2), 1.0};
        /* B = ds*(0,0,1)' (:176-178) */
        t[cnt++] = (trip){r0 + 2, var_ctrl(n, i), ds};
        /* c = ds*(f - df_x*x - df_u*u) (:179-186); bounds = -c (:221-224) */
        const double u_in = (xk_next - xk) / ds;
        const double g0 = (1 - xk * xl) * tp;
        const double g1 = (1 - xk * xl) * xk / cp - w->kref[i];
        const double g2 = u_in;
        const double c0 = ds * (g0 - (f00 * xl + f01 * xp));
        const double c1 = ds * (g1 - (f10 * xl + f11 * xp + f12 * xk));
        const double c2 = ds * (g2 - u_in);
        lo[r0] = up[r0] = -c0;
        lo[r0 + 1] = up[r0 + 1] = -c1;
        lo[r0 + 2] = up[r0 + 2] = -c2;
    }
    /* kappa rows (:189-191, :226-231) */
    const double kappa_limit = tan(q->max_steering_angle) / q->wheel_base;
    for (int i = 0; i < n; ++i) {
        const int r = row_kappa(n, i);
        t[cnt++] = (trip){r, var_state(i, 2), 1.0};
        lo[r] = -kappa_limit;
        up[r] = kappa_limit;
    }
    /* collision rows (:193-206, :233-248) */
    for (int i = 0; i < n; ++i) {
        if (i < p) {
            const int rf = row_coll(n, p, i, 0), rr = rf + 1;
            t[cnt++] = (trip){rf, var_state(i, 0), 1.0};
            t[cnt++] = (trip){rf, var_state(i, 1), q->front_length};
            t[cnt++] = (trip){rr, var_state(i, 0), 1.0};
            t[cnt++] = (trip){rr, var_state(i, 1), q->rear_length};
            t[cnt++] = (trip){rf, var_slack(n, p, i, 0), 1.0};
            t[cnt++] = (trip){rr, var_slack(n, p, i, 1), 1.0};
            soft_bounds(w->b0lb[i], w->b0ub[i], q->expected_safety_margin, &lo[rf], &up[rf]);
            soft_bounds(w->b1lb[i], w->b1ub[i], q->expected_safety_margin, &lo[rr], &up[rr]);
        } else {
            const int rc = row_coll(n, p, i, 0);
            t[cnt++] = (trip){rc, var_state(i, 0), 1.0};
            t[cnt++] = (trip){rc, var_slack(n, p, i, 0), 1.0};
            soft_bounds(w->b0lb[i], w->b0ub[i], q->expected_safety_margin, &lo[rc], &up[rc]);
        }
    }
    /* end state (:207-209, :250-259) */
    t[cnt++] = (trip){m - 2, 3 * n - 3, 1.0};
    t[cnt++] = (trip){m - 1, 3 * n - 2, 1.0};
    lo[m - 2] = q->end_l_lb;
    up[m - 2] = q->end_l_ub;
    lo[m - 1] = w->inst[PQP_I_EPSI_LO];
    up[m - 1] = w->inst[PQP_I_EPSI_HI];
    return cnt;
}

static int trip_cmp(const void *a, const void *b) {
    const trip *x = (const trip *)a, *y = (const trip *)b;
    if (x->c != y->c) return x->c - y->c;
    return x->r - y->r;
}

/* base_solver.cpp:119-148: diagonal Hessian */
static void emit_cost(const pqo_ws *w, double *Pd) {
    const int n = w->n, p = w->p;
    const pqp_params *q = &w->prm;
    memset(Pd, 0, sizeof(double) * w->nv);
    for (int i = 0; i < n; ++i) {
        Pd[var_state(i, 0)] += q->weight_l;
        Pd[var_state(i, 2)] += q->weight_kappa;
        if (i < p) {
            Pd[var_slack(n, p, i, 0)] += q->weight_slack;
            Pd[var_slack(n, p, i, 1)] += q->weight_slack;
        } else {
            Pd[var_slack(n, p, i, 0)] += q->weight_slack;
        }
        if (i != n - 1) Pd[var_ctrl(n, i)] += q->weight_dkappa;
    }
}

/* Cost model of the reference's dense temporaries (base_solver.cpp:122,145,159,210):
 * zero-fill an m x nv and an nv x nv dense matrix, write the entries, scan for nonzeros. */
static double assemble_dense_style(const pqo_ws *w) {
    const size_t m = w->m, nv = w->nv;
    double *cons = (double *)calloc(m * nv, sizeof(double));
    double *hess = (double *)calloc(nv * nv, sizeof(double));
    if (!cons || !hess) { free(cons); free(hess); return 0; }
    for (int j = 0; j < w->nv; ++j)
        for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e) cons[(size_t)j * m + w->Ai[e]] = w->Ax0[e];
    for (int j = 0; j < w->nv; ++j) hess[(size_t)j * nv + j] = w->Pd0[j];
    size_t nnz = 0;
    double acc = 0;
    for (size_t e = 0; e < m * nv; ++e) if (cons[e] != 0.0) { ++nnz; acc += cons[e]; }
    for (size_t e = 0; e < nv * nv; ++e) if (hess[e] != 0.0) { ++nnz; acc += hess[e]; }
    free(cons);
    free(hess);
    return acc + (double)nnz;
}

/* ------------------------------------------------------------------ OSQP: scaling */
static void limit_scaling(double *v, int n) {
    for (int i = 0; i < n; ++i) {
        v[i] = v[i] < MIN_SCALING ? 1.0 : v[i];
        v[i] = v[i] > MAX_SCALING ? MAX_SCALING : v[i];
    }
}

static void scale_data(pqo_ws *w) {
    const int nv = w->nv, m = w->m;
    memcpy(w->Ax, w->Ax0, sizeof(double) * w->nnzA);
    memcpy(w->Pd, w->Pd0, sizeof(double) * nv);
    memcpy(w->l, w->l0, sizeof(double) * m);
    memcpy(w->u, w->u0, sizeof(double) * m);
    w->c = 1.0;
    for (int j = 0; j < nv; ++j) w->D[j] = 1.0;
    for (int i = 0; i < m; ++i) w->E[i] = 1.0;
    for (int pass = 0; pass < w->prm.scaling; ++pass) {
        /* inf-norms of the columns of [P; A] and of the rows of A */
        for (int i = 0; i < m; ++i) w->E_temp[i] = 0.0;
        for (int j = 0; j < nv; ++j) {
            double pn = fabs(w->Pd[j]); /* P is diagonal: column norm of sym-triu P */
            double an = 0.0;
            for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e) {
                const double a = fabs(w->Ax[e]);
                an = dmax(an, a);
                w->E_temp[w->Ai[e]] = dmax(w->E_temp[w->Ai[e]], a);
            }
            w->D_temp[j] = dmax(pn, an);
        }
        limit_scaling(w->D_temp, nv);
        limit_scaling(w->E_temp, m);
        for (int j = 0; j < nv; ++j) w->D_temp[j] = 1.0 / sqrt(w->D_temp[j]);
        for (int i = 0; i < m; ++i) w->E_temp[i] = 1.0 / sqrt(w->E_temp[i]);
        /* P <- DPD, A <- EAD (q = 0) */
        for (int j = 0; j < nv; ++j) {
            w->Pd[j] *= w->D_temp[j] * w->D_temp[j];
            for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e)
                w->Ax[e] *= w->E_temp[w->Ai[e]] * w->D_temp[j];
            w->D[j] *= w->D_temp[j];
        }
        for (int i = 0; i < m; ++i) w->E[i] *= w->E_temp[i];
        /* cost normalisation: mean column norm of P vs ||q||_inf (q = 0 -> treated as 1) */
        double mean = 0.0;
        for (int j = 0; j < nv; ++j) mean += fabs(w->Pd[j]);
        mean /= nv;
        double inf_norm_q = 0.0;
        limit_scaling(&inf_norm_q, 1);
        double c_temp = dmax(mean, inf_norm_q);
        limit_scaling(&c_temp, 1);
        c_temp = 1.0 / c_temp;
        for (int j = 0; j < nv; ++j) w->Pd[j] *= c_temp;
        w->c *= c_temp;
    }
    w->cinv = 1.0 / w->c;
    for (int j = 0; j < nv; ++j) w->Dinv[j] = 1.0 / w->D[j];
    for (int i = 0; i < m; ++i) w->Einv[i] = 1.0 / w->E[i];
    for (int i = 0; i < m; ++i) { w->l[i] *= w->E[i]; w->u[i] *= w->E[i]; }
}

/* ------------------------------------------------------------------ OSQP: rho vector */
static void set_rho_vec(pqo_ws *w) {
    w->rho = dmin(dmax(w->rho, RHO_MIN), RHO_MAX);
    for (int i = 0; i < w->m; ++i) {
        if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) {
            w->constr_type[i] = -1;
            w->rho_vec[i] = RHO_MIN;
        } else if (w->u[i] - w->l[i] < RHO_TOL) {
            w->constr_type[i] = 1;
            w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->rho;
        } else {
            w->constr_type[i] = 0;
            w->rho_vec[i] = w->rho;
        }
        w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
    }
}

/* returns 1 if any constraint type changed */
static int update_rho_vec(pqo_ws *w) {
    int changed = 0;
    for (int i = 0; i < w->m; ++i) {
        int type;
        double r;
        if (w->l[i] < -OSQP_INFTY * MIN_SCALING && w->u[i] > OSQP_INFTY * MIN_SCALING) {
            type = -1; r = RHO_MIN;
        } else if (w->u[i] - w->l[i] < RHO_TOL) {
            type = 1; r = RHO_EQ_OVER_RHO_INEQ * w->rho;
        } else {
            type = 0; r = w->rho;
        }
        if (w->constr_type[i] != type) {
            w->constr_type[i] = type;
            w->rho_vec[i] = r;
            w->rho_inv_vec[i] = 1.0 / r;
            changed = 1;
        }
    }
    return changed;
}

/* ------------------------------------------------------------------ KKT + sparse LDL' */
/* Elimination order: leaves of the KKT graph first (slacks, kappa rows, controls, x0/end
 * rows), then the stage chain. Stands in for AMD. */
static void build_order(pqo_ws *w) {
    const int n = w->n, p = w->p, nv = w->nv, m = w->m;
    int k = 0;
    int *perm = w->perm;
    for (int r = 0; r < 3; ++r) perm[k++] = nv + row_dyn(0, r);
    for (int i = 0; i < n; ++i) {
        perm[k++] = var_slack(n, p, i, 0);
        if (i < p) perm[k++] = var_slack(n, p, i, 1);
        perm[k++] = nv + row_kappa(n, i);
        if (i < n - 1) perm[k++] = var_ctrl(n, i);
        if (i == n - 1) { perm[k++] = nv + m - 2; perm[k++] = nv + m - 1; }
        perm[k++] = nv + row_coll(n, p, i, 0);
        if (i < p) perm[k++] = nv + row_coll(n, p, i, 1);
        for (int c = 0; c < 3; ++c) perm[k++] = var_state(i, c);
        if (i < n - 1) for (int r = 0; r < 3; ++r) perm[k++] = nv + row_dyn(i + 1, r);
    }
    for (int i = 0; i < w->nk; ++i) w->pinv[perm[i]] = i;
}

typedef struct { int r, c, kind, idx; } kent;
static int kent_cmp(const void *a, const void *b) {
    const kent *x = (const kent *)a, *y = (const kent *)b;
    if (x->c != y->c) return x->c - y->c;
    return x->r - y->r;
}

/* pattern of the permuted upper-triangular KKT [[P+sigma I, A'],[A, -diag(1/rho)]] */
static int build_kkt_pattern(pqo_ws *w) {
    const int nv = w->nv, m = w->m;
    const int cnt_max = nv + w->nnzA + m;
    kent *e = (kent *)malloc(sizeof(kent) * cnt_max);
    if (!e) return -1;
    int cnt = 0;
    for (int j = 0; j < nv; ++j) { int pj = w->pinv[j]; e[cnt++] = (kent){pj, pj, 0, j}; }
    for (int j = 0; j < nv; ++j)
        for (int a = w->Ap[j]; a < w->Ap[j + 1]; ++a) {
            int pr = w->pinv[nv + w->Ai[a]], pc = w->pinv[j];
            if (pr > pc) { int t = pr; pr = pc; pc = t; }
            e[cnt++] = (kent){pr, pc, 1, a};
        }
    for (int r = 0; r < m; ++r) { int pr = w->pinv[nv + r]; e[cnt++] = (kent){pr, pr, 2, r}; }
    qsort(e, cnt, sizeof(kent), kent_cmp);
    w->nnzK = cnt;
    w->Kp = (int *)calloc(w->nk + 1, sizeof(int));
    w->Ki = (int *)malloc(sizeof(int) * cnt);
    w->Kx = (double *)malloc(sizeof(double) * cnt);
    w->Kkind = (int *)malloc(sizeof(int) * cnt);
    w->Kidx = (int *)malloc(sizeof(int) * cnt);
    for (int i = 0; i < cnt; ++i) {
        w->Kp[e[i].c + 1]++;
        w->Ki[i] = e[i].r;
        w->Kkind[i] = e[i].kind;
        w->Kidx[i] = e[i].idx;
    }
    for (int j = 0; j < w->nk; ++j) w->Kp[j + 1] += w->Kp[j];
    free(e);
    return 0;
}

static void fill_kkt_values(pqo_ws *w) {
    for (int i = 0; i < w->nnzK; ++i) {
        switch (w->Kkind[i]) {
            case 0: w->Kx[i] = w->Pd[w->Kidx[i]] + w->prm.sigma; break;
            case 1: w->Kx[i] = w->Ax[w->Kidx[i]]; break;
            default: w->Kx[i] = -w->rho_inv_vec[w->Kidx[i]]; break;
        }
    }
}

/* elimination tree and column counts of L for an upper-triangular CSC matrix */
static int ldl_symbolic(pqo_ws *w) {
    const int nk = w->nk;
    int *work = w->iwork;
    for (int i = 0; i < nk; ++i) { work[i] = 0; w->Lnz[i] = 0; w->etree[i] = UNKNOWN; }
    for (int j = 0; j < nk; ++j) {
        work[j] = j;
        for (int e = w->Kp[j]; e < w->Kp[j + 1]; ++e) {
            int i = w->Ki[e];
            if (i > j) return -1;
            while (work[i] != j) {
                if (w->etree[i] == UNKNOWN) w->etree[i] = j;
                w->Lnz[i]++;
                work[i] = j;
                i = w->etree[i];
            }
        }
    }
    w->Lp[0] = 0;
    for (int i = 0; i < nk; ++i) w->Lp[i + 1] = w->Lp[i] + w->Lnz[i];
    w->nnzL = w->Lp[nk];
    return 0;
}

/* up-looking numeric LDL': row k of L from a sparse triangular solve whose pattern is the
 * union of elimination-tree paths of the entries of column k of the upper triangle */
static int ldl_numeric(pqo_ws *w) {
    const int nk = w->nk;
    int *yIdx = w->iwork, *elim = w->iwork + nk, *next = w->iwork + 2 * nk;
    unsigned char *mark = w->bwork;
    double *yv = w->fwork;
    for (int i = 0; i < nk; ++i) { mark[i] = 0; yv[i] = 0.0; next[i] = w->Lp[i]; w->Dg[i] = 0.0; }
    for (int k = 0; k < nk; ++k) {
        int nnzY = 0;
        for (int e = w->Kp[k]; e < w->Kp[k + 1]; ++e) {
            int b = w->Ki[e];
            if (b == k) { w->Dg[k] = w->Kx[e]; continue; }
            yv[b] = w->Kx[e];
            if (!mark[b]) {
                int ne = 0, nx = b;
                while (nx != UNKNOWN && nx < k && !mark[nx]) {
                    mark[nx] = 1;
                    elim[ne++] = nx;
                    nx = w->etree[nx];
                }
                while (ne) yIdx[nnzY++] = elim[--ne];
            }
        }
        for (int i = nnzY - 1; i >= 0; --i) {
            const int c = yIdx[i];
            const int pos = next[c];
            const double yc = yv[c];
            for (int j = w->Lp[c]; j < pos; ++j) yv[w->Li[j]] -= w->Lx[j] * yc;
            w->Li[pos] = k;
            w->Lx[pos] = yc * w->Dginv[c];
            w->Dg[k] -= yc * w->Lx[pos];
            next[c]++;
            yv[c] = 0.0;
            mark[c] = 0;
        }
        if (w->Dg[k] == 0.0) return -1;
        w->Dginv[k] = 1.0 / w->Dg[k];
    }
    return 0;
}

static void ldl_solve(const pqo_ws *w, double *b) {
    const int nk = w->nk;
    double *bp = w->bp;
    for (int i = 0; i < nk; ++i) bp[i] = b[w->perm[i]];
    for (int i = 0; i < nk; ++i) {
        const double v = bp[i];
        for (int j = w->Lp[i]; j < w->Lp[i + 1]; ++j) bp[w->Li[j]] -= w->Lx[j] * v;
    }
    for (int i = 0; i < nk; ++i) bp[i] *= w->Dginv[i];
    for (int i = nk - 1; i >= 0; --i) {
        double v = bp[i];
        for (int j = w->Lp[i]; j < w->Lp[i + 1]; ++j) v -= w->Lx[j] * bp[w->Li[j]];
        bp[i] = v;
    }
    for (int i = 0; i < nk; ++i) b[w->perm[i]] = bp[i];
}

static int refactor(pqo_ws *w) {
    fill_kkt_values(w);
    return ldl_numeric(w);
}

/* ------------------------------------------------------------------ small linear algebra */
static void mat_vec_A(const pqo_ws *w, const double *x, double *out) { /* out = A x */
    for (int i = 0; i < w->m; ++i) out[i] = 0.0;
    for (int j = 0; j < w->nv; ++j) {
        const double xj = x[j];
        for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e) out[w->Ai[e]] += w->Ax[e] * xj;
    }
}
static void mat_tvec_A(const pqo_ws *w, const double *y, double *out) { /* out = A' y */
    for (int j = 0; j < w->nv; ++j) {
        double acc = 0.0;
        for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e) acc += w->Ax[e] * y[w->Ai[e]];
        out[j] = acc;
    }
}
static double norm_inf(const double *v, int n) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = dmax(r, fabs(v[i]));
    return r;
}
static double scaled_norm_inf(const double *s, const double *v, int n) {
    double r = 0.0;
    for (int i = 0; i < n; ++i) r = dmax(r, fabs(s[i] * v[i]));
    return r;
}

/* ------------------------------------------------------------------ setup */
static void *xcalloc(size_t n, size_t sz) { return calloc(n ? n : 1, sz); }

static int assemble(pqo_ws *w, int first) {
    const int ntr_max = 17 * w->n + 8;
    trip *t = (trip *)malloc(sizeof(trip) * ntr_max);
    if (!t) return -1;
    int cnt = emit_constraints(w, t, w->l0, w->u0);
    qsort(t, cnt, sizeof(trip), trip_cmp);
    if (first) {
        w->nnzA = cnt;
        w->Ap = (int *)xcalloc(w->nv + 1, sizeof(int));
        w->Ai = (int *)xcalloc(cnt, sizeof(int));
        w->Ax0 = (double *)xcalloc(cnt, sizeof(double));
        w->Ax = (double *)xcalloc(cnt, sizeof(double));
        for (int i = 0; i < cnt; ++i) { w->Ap[t[i].c + 1]++; w->Ai[i] = t[i].r; }
        for (int j = 0; j < w->nv; ++j) w->Ap[j + 1] += w->Ap[j];
    } else if (cnt != w->nnzA) {
        free(t);
        return -1;
    }
    for (int i = 0; i < cnt; ++i) w->Ax0[i] = t[i].v;
    free(t);
    return 0;
}

pqo_ws *pqo_setup(const pqp_params *prm, int n, int p, const double *knots, int stride,
                  const double *inst) {
    if (!prm || !knots || !inst || n < 2 || p < 0 || p > n) return NULL;
    pqo_ws *w = (pqo_ws *)calloc(1, sizeof(pqo_ws));
    if (!w) return NULL;
    w->prm = *prm;
    w->n = n;
    w->p = p;
    /* sizes (base_solver.cpp:22-37) */
    w->nv = 3 * n + (n - 1) + (p + n);
    w->m = 4 * n + p + n + 2;
    const int nv = w->nv, m = w->m;
    double **fields[9] = {&w->s, &w->kref, &w->lin_l, &w->lin_psi, &w->lin_k,
                          &w->b0lb, &w->b0ub, &w->b1lb, &w->b1ub};
    for (int f = 0; f < PQP_NFIELDS; ++f) {
        *fields[f] = (double *)xcalloc(n, sizeof(double));
        memcpy(*fields[f], knots + (size_t)f * stride, sizeof(double) * n);
    }
    memcpy(w->inst, inst, sizeof(double) * PQP_NINST);
    w->l0 = (double *)xcalloc(m, sizeof(double));
    w->u0 = (double *)xcalloc(m, sizeof(double));
    w->l = (double *)xcalloc(m, sizeof(double));
    w->u = (double *)xcalloc(m, sizeof(double));
    w->Pd0 = (double *)xcalloc(nv, sizeof(double));
    w->Pd = (double *)xcalloc(nv, sizeof(double));
    w->D = (double *)xcalloc(nv, sizeof(double));
    w->Dinv = (double *)xcalloc(nv, sizeof(double));
    w->E = (double *)xcalloc(m, sizeof(double));
    w->Einv = (double *)xcalloc(m, sizeof(double));
    w->D_temp = (double *)xcalloc(nv, sizeof(double));
    w->D_temp_A = (double *)xcalloc(nv, sizeof(double));
    w->E_temp = (double *)xcalloc(m, sizeof(double));
    w->rho_vec = (double *)xcalloc(m, sizeof(double));
    w->rho_inv_vec = (double *)xcalloc(m, sizeof(double));
    w->constr_type = (int *)xcalloc(m, sizeof(int));
    emit_cost(w, w->Pd0);
    if (assemble(w, 1)) { pqo_free(w); return NULL; }
    /* osqp_setup: scale, rho vector, KKT, factor */
    for (int i = 0; i < m; ++i) { w->E[i] = 1.0; w->Einv[i] = 1.0; }
    for (int j = 0; j < nv; ++j) { w->D[j] = 1.0; w->Dinv[j] = 1.0; }
    w->c = w->cinv = 1.0;
    if (w->prm.scaling) scale_data(w);
    else {
        memcpy(w->Ax, w->Ax0, sizeof(double) * w->nnzA);
        memcpy(w->Pd, w->Pd0, sizeof(double) * nv);
        memcpy(w->l, w->l0, sizeof(double) * m);
        memcpy(w->u, w->u0, sizeof(double) * m);
    }
    w->rho = w->prm.rho;
    set_rho_vec(w);
    w->nk = nv + m;
    const int nk = w->nk;
    w->perm = (int *)xcalloc(nk, sizeof(int));
    w->pinv = (int *)xcalloc(nk, sizeof(int));
    build_order(w);
    if (build_kkt_pattern(w)) { pqo_free(w); return NULL; }
    w->etree = (int *)xcalloc(nk, sizeof(int));
    w->Lnz = (int *)xcalloc(nk, sizeof(int));
    w->Lp = (int *)xcalloc(nk + 1, sizeof(int));
    w->iwork = (int *)xcalloc(3 * (size_t)nk, sizeof(int));
    w->bwork = (unsigned char *)xcalloc(nk, 1);
    w->fwork = (double *)xcalloc(nk, sizeof(double));
    w->bp = (double *)xcalloc(nk, sizeof(double));
    w->Dg = (double *)xcalloc(nk, sizeof(double));
    w->Dginv = (double *)xcalloc(nk, sizeof(double));
    if (ldl_symbolic(w)) { pqo_free(w); return NULL; }
    w->Li = (int *)xcalloc(w->nnzL, sizeof(int));
    w->Lx = (double *)xcalloc(w->nnzL, sizeof(double));
    if (refactor(w)) { pqo_free(w); return NULL; }
    w->x = (double *)xcalloc(nv, sizeof(double));
    w->x_prev = (double *)xcalloc(nv, sizeof(double));
    w->delta_x = (double *)xcalloc(nv, sizeof(double));
    w->Px = (double *)xcalloc(nv, sizeof(double));
    w->Aty = (double *)xcalloc(nv, sizeof(double));
    w->Atdy = (double *)xcalloc(nv, sizeof(double));
    w->z = (double *)xcalloc(m, sizeof(double));
    w->z_prev = (double *)xcalloc(m, sizeof(double));
    w->y = (double *)xcalloc(m, sizeof(double));
    w->delta_y = (double *)xcalloc(m, sizeof(double));
    w->Axv = (double *)xcalloc(m, sizeof(double));
    w->Adx = (double *)xcalloc(m, sizeof(double));
    w->xz_tilde = (double *)xcalloc(nk, sizeof(double));
    w->status = PQP_UNSOLVED;
    return w;
}

void pqo_free(pqo_ws *w) {
    if (!w) return;
    free(w->s); free(w->kref); free(w->lin_l); free(w->lin_psi); free(w->lin_k);
    free(w->b0lb); free(w->b0ub); free(w->b1lb); free(w->b1ub);
    free(w->Ap); free(w->Ai); free(w->Ax0); free(w->l0); free(w->u0); free(w->Pd0);
    free(w->Ax); free(w->l); free(w->u); free(w->Pd);
    free(w->D); free(w->E); free(w->Dinv); free(w->Einv);
    free(w->D_temp); free(w->D_temp_A); free(w->E_temp);
    free(w->Rp); free(w->Rj); free(w->Re);
    free(w->rho_vec); free(w->rho_inv_vec); free(w->constr_type);
    free(w->perm); free(w->pinv); free(w->Kp); free(w->Ki); free(w->Kkind); free(w->Kidx); free(w->Kx);
    free(w->etree); free(w->Lnz); free(w->Lp); free(w->Li); free(w->Lx); free(w->Dg); free(w->Dginv);
    free(w->iwork); free(w->bwork); free(w->fwork); free(w->bp);
    free(w->x); free(w->z); free(w->y); free(w->x_prev); free(w->z_prev); free(w->xz_tilde);
    free(w->delta_x); free(w->delta_y); free(w->Axv); free(w->Px); free(w->Aty); free(w->Atdy); free(w->Adx);
    free(w);
}

/* ------------------------------------------------------------------ residuals / termination */
static double compute_pri_res(pqo_ws *w) {
    mat_vec_A(w, w->x, w->Axv);
    for (int i = 0; i < w->m; ++i) w->z_prev[i] = w->Axv[i] - w->z[i];
    if (w->prm.scaling) return scaled_norm_inf(w->Einv, w->z_prev, w->m);
    return norm_inf(w->z_prev, w->m);
}
static double compute_pri_tol(const pqo_ws *w, double eps_abs, double eps_rel) {
    double mx;
    if (w->prm.scaling)
        mx = dmax(scaled_norm_inf(w->Einv, w->z, w->m), scaled_norm_inf(w->Einv, w->Axv, w->m));
    else
        mx = dmax(norm_inf(w->z, w->m), norm_inf(w->Axv, w->m));
    return eps_abs + eps_rel * mx;
}
static double compute_dua_res(pqo_ws *w) {
    /* q = 0 */
    for (int j = 0; j < w->nv; ++j) w->Px[j] = w->Pd[j] * w->x[j];
    mat_tvec_A(w, w->y, w->Aty);
    for (int j = 0; j < w->nv; ++j) w->x_prev[j] = w->Px[j] + w->Aty[j];
    if (w->prm.scaling) return w->cinv * scaled_norm_inf(w->Dinv, w->x_prev, w->nv);
    return norm_inf(w->x_prev, w->nv);
}
static double compute_dua_tol(const pqo_ws *w, double eps_abs, double eps_rel) {
    double mx;
    if (w->prm.scaling) {
        mx = 0.0; /* ||Dinv q|| = 0 */
        mx = dmax(mx, scaled_norm_inf(w->Dinv, w->Aty, w->nv));
        mx = dmax(mx, scaled_norm_inf(w->Dinv, w->Px, w->nv));
        return eps_abs + eps_rel * w->cinv * mx;
    }
    mx = dmax(norm_inf(w->Aty, w->nv), norm_inf(w->Px, w->nv));
    return eps_abs + eps_rel * mx;
}

static int is_primal_infeasible(pqo_ws *w, double eps) {
    double norm_dy, ineq_lhs = 0.0;
    for (int i = 0; i < w->m; ++i) {
        if (w->u[i] > OSQP_INFTY * MIN_SCALING) {
            if (w->l[i] < -OSQP_INFTY * MIN_SCALING) w->delta_y[i] = 0.0;
            else w->delta_y[i] = dmin(w->delta_y[i], 0.0);
        } else if (w->l[i] < -OSQP_INFTY * MIN_SCALING) {
            w->delta_y[i] = dmax(w->delta_y[i], 0.0);
        }
    }
    if (w->prm.scaling) norm_dy = scaled_norm_inf(w->E, w->delta_y, w->m);
    else norm_dy = norm_inf(w->delta_y, w->m);
    if (norm_dy > eps) {
        for (int i = 0; i < w->m; ++i)
            ineq_lhs += w->u[i] * dmax(w->delta_y[i], 0.0) + w->l[i] * dmin(w->delta_y[i], 0.0);
        if (ineq_lhs < -eps * norm_dy) {
            mat_tvec_A(w, w->delta_y, w->Atdy);
            if (w->prm.scaling) for (int j = 0; j < w->nv; ++j) w->Atdy[j] *= w->Dinv[j];
            return norm_inf(w->Atdy, w->nv) < eps * norm_dy;
        }
    }
    return 0;
}

static int is_dual_infeasible(pqo_ws *w, double eps) {
    double norm_dx, cost_scaling;
    if (w->prm.scaling) { norm_dx = scaled_norm_inf(w->D, w->delta_x, w->nv); cost_scaling = w->c; }
    else { norm_dx = norm_inf(w->delta_x, w->nv); cost_scaling = 1.0; }
    if (norm_dx > eps) {
        /* q' delta_x = 0 with q = 0 */
        if (0.0 < -cost_scaling * eps * norm_dx) {
            return 0; /* unreachable: kept to mirror the three-condition structure */
        }
    }
    return 0;
}

static void update_info(pqo_ws *w, int iter) {
    w->iter = iter;
    double obj = 0.0;
    for (int j = 0; j < w->nv; ++j) obj += 0.5 * w->Pd[j] * w->x[j] * w->x[j];
    w->obj_val = w->prm.scaling ? obj * w->cinv : obj;
    w->pri_res = w->m ? compute_pri_res(w) : 0.0;
    w->dua_res = compute_dua_res(w);
}

static int check_termination(pqo_ws *w, int approximate) {
    double eps_abs = w->prm.eps_abs, eps_rel = w->prm.eps_rel;
    double eps_pinf = w->prm.eps_prim_inf, eps_dinf = w->prm.eps_dual_inf;
    int prim_ok = 0, dual_ok = 0, prim_inf = 0, dual_inf = 0;
    if (w->pri_res > OSQP_INFTY || w->dua_res > OSQP_INFTY || w->pri_res != w->pri_res ||
        w->dua_res != w->dua_res) {
        w->status = PQP_NUMERICAL_ERROR;
        return 1;
    }
    if (approximate) { eps_abs *= 10; eps_rel *= 10; eps_pinf *= 10; eps_dinf *= 10; }
    const double eps_prim = compute_pri_tol(w, eps_abs, eps_rel);
    if (w->pri_res < eps_prim) prim_ok = 1;
    else prim_inf = is_primal_infeasible(w, eps_pinf);
    const double eps_dual = compute_dua_tol(w, eps_abs, eps_rel);
    if (w->dua_res < eps_dual) dual_ok = 1;
    else dual_inf = is_dual_infeasible(w, eps_dinf);
    if (prim_ok && dual_ok) {
        w->status = approximate ? PQP_SOLVED_INACCURATE : PQP_SOLVED;
        return 1;
    } else if (prim_inf) {
        w->status = approximate ? PQP_PRIMAL_INFEASIBLE_INACCURATE : PQP_PRIMAL_INFEASIBLE;
        return 1;
    } else if (dual_inf) {
        w->status = approximate ? PQP_DUAL_INFEASIBLE_INACCURATE : PQP_DUAL_INFEASIBLE;
        return 1;
    }
    return 0;
}

static double compute_rho_estimate(const pqo_ws *w) {
    double pri = norm_inf(w->z_prev, w->m);  /* scaled Ax - z */
    double dua = norm_inf(w->x_prev, w->nv); /* scaled Px + q + A'y */
    double pn = dmax(norm_inf(w->z, w->m), norm_inf(w->Axv, w->m));
    pri /= (pn + 1e-10);
    double dn = dmax(0.0, norm_inf(w->Aty, w->nv));
    dn = dmax(dn, norm_inf(w->Px, w->nv));
    dua /= (dn + 1e-10);
    double est = w->rho * sqrt(pri / (dua + 1e-10));
    return dmin(dmax(est, RHO_MIN), RHO_MAX);
}

static int update_rho(pqo_ws *w, double rho_new) {
    w->rho = dmin(dmax(rho_new, RHO_MIN), RHO_MAX);
    for (int i = 0; i < w->m; ++i) {
        if (w->constr_type[i] == 0) {
            w->rho_vec[i] = w->rho;
            w->rho_inv_vec[i] = 1.0 / w->rho;
        } else if (w->constr_type[i] == 1) {
            w->rho_vec[i] = RHO_EQ_OVER_RHO_INEQ * w->rho;
            w->rho_inv_vec[i] = 1.0 / w->rho_vec[i];
        }
    }
    return refactor(w);
}

/* ------------------------------------------------------------------ osqp_solve */
int pqo_solve(pqo_ws *w) {
    const int nv = w->nv, m = w->m;
    const double alpha = w->prm.alpha, sigma = w->prm.sigma;
    int iter, can_check = 0;
    w->status = PQP_UNSOLVED;
    w->rho_updates = 0;
    w->iter = 0;
    for (iter = 1; iter <= w->prm.max_iter; ++iter) {
        double *t;
        t = w->x; w->x = w->x_prev; w->x_prev = t;
        t = w->z; w->z = w->z_prev; w->z_prev = t;
        /* update_xz_tilde: rhs then KKT solve */
        for (int j = 0; j < nv; ++j) w->xz_tilde[j] = sigma * w->x_prev[j]; /* - q, q = 0 */
        for (int i = 0; i < m; ++i) w->xz_tilde[nv + i] = w->z_prev[i] - w->rho_inv_vec[i] * w->y[i];
        {
            /* solve K [x~; nu] = rhs, then z~ = rhs_z + rho^-1 nu */
            double *sol = w->fwork; /* fwork is free outside ldl_numeric */
            memcpy(sol, w->xz_tilde, sizeof(double) * w->nk);
            ldl_solve(w, sol);
            for (int j = 0; j < nv; ++j) w->xz_tilde[j] = sol[j];
            for (int i = 0; i < m; ++i) w->xz_tilde[nv + i] += w->rho_inv_vec[i] * sol[nv + i];
        }
        /* update_x */
        for (int j = 0; j < nv; ++j) {
            w->x[j] = alpha * w->xz_tilde[j] + (1.0 - alpha) * w->x_prev[j];
            w->delta_x[j] = w->x[j] - w->x_prev[j];
        }
        /* update_z */
        for (int i = 0; i < m; ++i) {
            double v = alpha * w->xz_tilde[nv + i] + (1.0 - alpha) * w->z_prev[i] +
                       w->rho_inv_vec[i] * w->y[i];
            w->z[i] = dmin(dmax(v, w->l[i]), w->u[i]);
        }
        /* update_y */
        for (int i = 0; i < m; ++i) {
            w->delta_y[i] = w->rho_vec[i] * (alpha * w->xz_tilde[nv + i] +
                                              (1.0 - alpha) * w->z_prev[i] - w->z[i]);
            w->y[i] += w->delta_y[i];
        }
        can_check = w->prm.check_termination && (iter % w->prm.check_termination == 0);
        if (can_check) {
            update_info(w, iter);
            if (check_termination(w, 0)) break;
        }
        if (w->prm.adaptive_rho && w->prm.adaptive_rho_interval &&
            (iter % w->prm.adaptive_rho_interval == 0)) {
            if (!can_check) update_info(w, iter);
            double rho_new = compute_rho_estimate(w);
            if (rho_new > w->rho * w->prm.adaptive_rho_tolerance ||
                rho_new < w->rho / w->prm.adaptive_rho_tolerance) {
                if (update_rho(w, rho_new)) { w->status = PQP_NUMERICAL_ERROR; return w->status; }
                w->rho_updates++;
            }
        }
    }
    if (!can_check) {
        update_info(w, iter - 1);
        check_termination(w, 0);
    }
    if (w->status == PQP_UNSOLVED) {
        if (!check_termination(w, 1)) w->status = PQP_MAX_ITER_REACHED;
    }
    return w->status;
}

/* ------------------------------------------------------------------ warm update */
int pqo_update(pqo_ws *w, const double *l, const double *psi, const double *k) {
    if (l != w->lin_l) memcpy(w->lin_l, l, sizeof(double) * w->n);
    if (psi != w->lin_psi) memcpy(w->lin_psi, psi, sizeof(double) * w->n);
    if (k != w->lin_k) memcpy(w->lin_k, k, sizeof(double) * w->n);
    if (assemble(w, 0)) return -1;
    /* osqp_update_bounds: scale the new bounds with the CURRENT E, refresh constraint types */
    for (int i = 0; i < w->m; ++i) { w->l[i] = w->l0[i] * w->E[i]; w->u[i] = w->u0[i] * w->E[i]; }
    if (update_rho_vec(w)) { if (refactor(w)) return -1; }
    /* osqp_update_A: unscale, overwrite, scale from scratch (new D, E, c; the scaled
     * iterates x, z, y are NOT touched), refactor */
    if (w->prm.scaling) scale_data(w);
    else {
        memcpy(w->Ax, w->Ax0, sizeof(double) * w->nnzA);
        memcpy(w->l, w->l0, sizeof(double) * w->m);
        memcpy(w->u, w->u0, sizeof(double) * w->m);
    }
    if (refactor(w)) return -1;
    w->status = PQP_UNSOLVED;
    return 0;
}

/* Receding-horizon update: every input of the instance changes (window shifted along the
 * reference, new clearance bounds, new x0), the sparsity pattern does not. Same OSQP calls as
 * pqo_update: osqp_update_bounds, then osqp_update_A. */
int pqo_update_full(pqo_ws *w, const double *knots, int stride, const double *inst) {
    double **fields[9] = {&w->s, &w->kref, &w->lin_l, &w->lin_psi, &w->lin_k,
                          &w->b0lb, &w->b0ub, &w->b1lb, &w->b1ub};
    for (int f = 0; f < PQP_NFIELDS; ++f) memcpy(*fields[f], knots + (size_t)f * stride, sizeof(double) * w->n);
    memcpy(w->inst, inst, sizeof(double) * PQP_NINST);
    return pqo_update(w, w->lin_l, w->lin_psi, w->lin_k);
}

/* ------------------------------------------------------------------ getters */
int pqo_nv(const pqo_ws *w) { return w->nv; }
int pqo_m(const pqo_ws *w) { return w->m; }
int pqo_iters(const pqo_ws *w) { return w->iter; }
int pqo_status(const pqo_ws *w) { return w->status; }
int pqo_rho_updates(const pqo_ws *w) { return w->rho_updates; }
double pqo_rho(const pqo_ws *w) { return w->rho; }
double pqo_cost(const pqo_ws *w) { return w->obj_val; }
double pqo_pri_res(const pqo_ws *w) { return w->pri_res; }
double pqo_dua_res(const pqo_ws *w) { return w->dua_res; }
int pqo_nnz_L(const pqo_ws *w) { return w->nnzL; }
int pqo_nnz_A(const pqo_ws *w) { return w->nnzA; }
void pqo_get_x(const pqo_ws *w, double *x) { for (int j = 0; j < w->nv; ++j) x[j] = w->D[j] * w->x[j]; }
void pqo_get_y(const pqo_ws *w, double *y) { for (int i = 0; i < w->m; ++i) y[i] = w->cinv * w->E[i] * w->y[i]; }
void pqo_get_z(const pqo_ws *w, double *z) { for (int i = 0; i < w->m; ++i) z[i] = w->Einv[i] * w->z[i]; }
void pqo_get_scaled_iterates(const pqo_ws *w, double *x, double *z, double *y) {
    memcpy(x, w->x, sizeof(double) * w->nv);
    memcpy(z, w->z, sizeof(double) * w->m);
    memcpy(y, w->y, sizeof(double) * w->m);
}
void pqo_get_scaling(const pqo_ws *w, double *D, double *E, double *c) {
    memcpy(D, w->D, sizeof(double) * w->nv);
    memcpy(E, w->E, sizeof(double) * w->m);
    *c = w->c;
}
void pqo_get_sol(const pqo_ws *w, double *sol, int stride) {
    const int n = w->n;
    for (int i = 0; i < n; ++i) {
        sol[0 * stride + i] = w->D[3 * i] * w->x[3 * i];
        sol[1 * stride + i] = w->D[3 * i + 1] * w->x[3 * i + 1];
        sol[2 * stride + i] = w->D[3 * i + 2] * w->x[3 * i + 2];
        sol[3 * stride + i] = (i < n - 1) ? w->D[3 * n + i] * w->x[3 * n + i] : 0.0;
    }
}
void pqo_get_problem(const pqo_ws *w, int *Ap, int *Ai, double *Ax, double *l, double *u,
                     double *Pdiag) {
    memcpy(Ap, w->Ap, sizeof(int) * (w->nv + 1));
    memcpy(Ai, w->Ai, sizeof(int) * w->nnzA);
    memcpy(Ax, w->Ax0, sizeof(double) * w->nnzA);
    memcpy(l, w->l0, sizeof(double) * w->m);
    memcpy(u, w->u0, sizeof(double) * w->m);
    memcpy(Pdiag, w->Pd0, sizeof(double) * w->nv);
}

/* include/tools/tools.hpp:24-35 */
static double constrain_angle(double a) {
    while (a > M_PI) a -= 2 * M_PI;
    while (a < -M_PI) a += 2 * M_PI;
    return a;
}

/* base_solver.cpp:263-288 */
void pqo_frenet_to_cartesian(int n, const double *ref_x, const double *ref_y,
                             const double *ref_heading, const double *l, const double *psi,
                             double *out_x, double *out_y, double *out_heading) {
    for (int i = 0; i < n; ++i) {
        const double angle = ref_heading[i];
        out_heading[i] = constrain_angle(angle + psi[i]);
        const double new_angle = constrain_angle(angle + M_PI_2);
        out_x[i] = ref_x[i] + l[i] * cos(new_angle);
        out_y[i] = ref_y[i] + l[i] * sin(new_angle);
    }
}

/* ------------------------------------------------------------------ batch driver */
static double now_s(void) {
    struct timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int pqo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

double pqo_solve_batch(const pqp_params *prm, const pqp_batch_in *in, const pqp_batch_out *out,
                       int nthreads, int mode, int dense_assembly) {
    if (!prm || !in || !out || !in->knots || !in->inst || !in->n || !out->sol) return -1.0;
    const int B = in->batch, nmax = in->n_max;
    const int nvmax = 6 * nmax - 1, mmax = 6 * nmax + 2;
    volatile double sink = 0.0;
    int err = 0;
    if (nthreads < 1) nthreads = 1;
    const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int b = 0; b < B; ++b) {
        const int n = in->n[b];
        const int p = in->p ? in->p[b] : n;
        const double *kn = in->knots + (size_t)b * PQP_NFIELDS * nmax;
#ifdef _OPENMP
        const int dbg = getenv("PQO_DEBUG") != NULL;
        const double tb0 = now_s();
#endif
        pqo_ws *w = pqo_setup(prm, n, p, kn, nmax, in->inst + (size_t)b * PQP_NINST);
        if (!w) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
            continue;
        }
        const double tb1 = now_s();
        if (dense_assembly) sink += assemble_dense_style(w);
        int st = pqo_solve(w);
        int iters = w->iter;
        const double tb2 = now_s();
        if (dbg) fprintf(stderr, "inst %d setup %.2f ms solve %.2f ms iters %d\n", b, 1e3 * (tb1 - tb0), 1e3 * (tb2 - tb1), iters);
        if (mode == 1 && (st == PQP_SOLVED)) {
            double *sol = (double *)malloc(sizeof(double) * 4 * n);
            pqo_get_sol(w, sol, n);
            pqo_update(w, sol, sol + n, sol + 2 * n);
            if (dense_assembly) sink += assemble_dense_style(w);
            st = pqo_solve(w);
            iters += w->iter;
            free(sol);
        }
        double *so = out->sol + (size_t)b * 4 * nmax;
        pqo_get_sol(w, so, nmax);
        if (out->cost) out->cost[b] = w->obj_val;
        if (out->status) out->status[b] = st;
        if (out->iters) out->iters[b] = iters;
        if (out->x_full) pqo_get_x(w, out->x_full + (size_t)b * nvmax);
        if (out->y_full) pqo_get_y(w, out->y_full + (size_t)b * mmax);
        if (out->z_full) pqo_get_z(w, out->z_full + (size_t)b * mmax);
        if (out->info) {
            double *inf = out->info + (size_t)b * PQP_NINFO;
            inf[PQP_INFO_PRI_RES] = w->pri_res;
            inf[PQP_INFO_DUA_RES] = w->dua_res;
            inf[PQP_INFO_RHO] = w->rho;
            inf[PQP_INFO_RHO_UPDATES] = w->rho_updates;
        }
        pqo_free(w);
#ifdef _OPENMP
        if (dbg) fprintf(stderr, "inst %d thread %d start %.4f end %.4f\n", b, omp_get_thread_num(), tb0 - t0, now_s() - t0);
#endif
    }
    const double t1 = now_s();
    (void)sink;
    return err ? -1.0 : (t1 - t0);
}

/* number of threads an OpenMP region with `nthreads` requested actually gets (diagnostic) */
int pqo_omp_probe(int nthreads) {
    int cnt = 0;
#ifdef _OPENMP
#pragma omp parallel num_threads(nthreads)
    {
#pragma omp atomic
        cnt++;
    }
#else
    cnt = 1;
#endif
    return cnt;
}

/* ------------------------------------------------------------------ batch checkers (tests only)
 * OSQP's own unscaled termination test of a GIVEN point (x, y, z) - e.g. the
 * CUDA library's x_full / y_full / z_full outputs, reference index order - evaluated in FP64 on the
 * (P, A, l, u) this oracle assembles from the same inputs (linearisation fields included), for every
 * instance of a batch. rep[b*6] = {pri_res, eps_pri, dua_res, eps_dua, worst violation of l <= z <= u
 * on finite bounds, 0.5 x'Px}. Returns 0, or -1 if an instance could not be set up. */
int pqo_termination_batch(const pqp_params *prm, const pqp_batch_in *in, const double *x_full,
                          const double *y_full, const double *z_full, double *rep, int nthreads) {
    if (!prm || !in || !x_full || !y_full || !z_full || !rep) return -1;
    const int B = in->batch, nmax = in->n_max;
    const int nvmax = 6 * nmax - 1, mmax = 6 * nmax + 2;
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int b = 0; b < B; ++b) {
        const int n = in->n[b];
        const int p = in->p ? in->p[b] : n;
        pqo_ws *w = pqo_setup(prm, n, p, in->knots + (size_t)b * PQP_NFIELDS * nmax, nmax,
                              in->inst + (size_t)b * PQP_NINST);
        if (!w) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
            continue;
        }
        const double *x = x_full + (size_t)b * nvmax, *y = y_full + (size_t)b * mmax, *z = z_full + (size_t)b * mmax;
        double *Ax = (double *)calloc((size_t)w->m, sizeof(double));
        double nAx = 0.0, nz = 0.0, pri = 0.0, viol = 0.0, dua = 0.0, nPx = 0.0, nAty = 0.0, cost = 0.0;
        for (int j = 0; j < w->nv; ++j)
            for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e) Ax[w->Ai[e]] += w->Ax0[e] * x[j];
        for (int i = 0; i < w->m; ++i) {
            pri = dmax(pri, fabs(Ax[i] - z[i]));
            nAx = dmax(nAx, fabs(Ax[i]));
            nz = dmax(nz, fabs(z[i]));
            if (w->l0[i] > -1e29) viol = dmax(viol, w->l0[i] - z[i]);
            if (w->u0[i] < 1e29) viol = dmax(viol, z[i] - w->u0[i]);
        }
        for (int j = 0; j < w->nv; ++j) {
            double aty = 0.0;
            for (int e = w->Ap[j]; e < w->Ap[j + 1]; ++e) aty += w->Ax0[e] * y[w->Ai[e]];
            const double px = w->Pd0[j] * x[j];
            dua = dmax(dua, fabs(px + aty));
            nPx = dmax(nPx, fabs(px));
            nAty = dmax(nAty, fabs(aty));
            cost += 0.5 * px * x[j];
        }
        double *r = rep + (size_t)b * 6;
        r[0] = pri;
        r[1] = prm->eps_abs + prm->eps_rel * dmax(nAx, nz);
        r[2] = dua;
        r[3] = prm->eps_abs + prm->eps_rel * dmax(nPx, nAty);  /* q = 0 */
        r[4] = viol;
        r[5] = cost;
        free(Ax);
        pqo_free(w);
    }
    return err ? -1 : 0;
}

/* A batch of persistent workspaces: the reference's solver object per path, kept across the warm
 * re-solves of a receding-horizon run (BASELINE configs[4]). */
struct pqo_batch {
    int B, nmax;
    pqo_ws **w;
};

void pqo_batch_free(pqo_batch *pb) {
    if (!pb) return;
    for (int b = 0; b < pb->B; ++b) pqo_free(pb->w[b]);
    free(pb->w);
    free(pb);
}
pqo_batch *pqo_batch_setup(const pqp_params *prm, const pqp_batch_in *in, int nthreads) {
    if (!prm || !in) return NULL;
    pqo_batch *pb = (pqo_batch *)calloc(1, sizeof(pqo_batch));
    pb->B = in->batch;
    pb->nmax = in->n_max;
    pb->w = (pqo_ws **)calloc((size_t)in->batch, sizeof(pqo_ws *));
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int b = 0; b < in->batch; ++b) {
        pb->w[b] = pqo_setup(prm, in->n[b], in->p ? in->p[b] : in->n[b], in->knots + (size_t)b * PQP_NFIELDS * in->n_max,
                             in->n_max, in->inst + (size_t)b * PQP_NINST);
        if (!pb->w[b]) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
        }
    }
    if (err) { pqo_batch_free(pb); return NULL; }
    return pb;
}
/* osqp_update_bounds + osqp_update_A of every instance from a new knot/inst block (pqo_update_full) */
int pqo_batch_update_full(pqo_batch *pb, const pqp_batch_in *in, int nthreads) {
    if (!pb || !in || in->batch != pb->B || in->n_max != pb->nmax) return -1;
    int err = 0;
    if (nthreads < 1) nthreads = 1;
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 4) num_threads(nthreads)
#endif
    for (int b = 0; b < pb->B; ++b)
        if (pqo_update_full(pb->w[b], in->knots + (size_t)b * PQP_NFIELDS * pb->nmax, pb->nmax, in->inst + (size_t)b * PQP_NINST)) {
#ifdef _OPENMP
#pragma omp atomic write
#endif
            err = 1;
        }
    return err ? -1 : 0;
}
/* solve every instance (warm after the first call) and write the batch outputs */
double pqo_batch_solve(pqo_batch *pb, const pqp_batch_out *out, int nthreads) {
    if (!pb || !out || !out->sol) return -1.0;
    const int nmax = pb->nmax, nvmax = 6 * nmax - 1, mmax = 6 * nmax + 2;
    if (nthreads < 1) nthreads = 1;
    const double t0 = now_s();
#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 1) num_threads(nthreads)
#endif
    for (int b = 0; b < pb->B; ++b) {
        pqo_ws *w = pb->w[b];
        const int st = pqo_solve(w);
        pqo_get_sol(w, out->sol + (size_t)b * 4 * nmax, nmax);
        if (out->cost) out->cost[b] = w->obj_val;
        if (out->status) out->status[b] = st;
        if (out->iters) out->iters[b] = w->iter;
        if (out->x_full) pqo_get_x(w, out->x_full + (size_t)b * nvmax);
        if (out->y_full) pqo_get_y(w, out->y_full + (size_t)b * mmax);
        if (out->z_full) pqo_get_z(w, out->z_full + (size_t)b * mmax);
        if (out->info) {
            double *inf = out->info + (size_t)b * PQP_NINFO;
            inf[PQP_INFO_PRI_RES] = w->pri_res;
            inf[PQP_INFO_DUA_RES] = w->dua_res;
            inf[PQP_INFO_RHO] = w->rho;
            inf[PQP_INFO_RHO_UPDATES] = w->rho_updates;
        }
    }
    return now_s() - t0;
}
