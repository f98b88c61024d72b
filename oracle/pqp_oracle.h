/*
 * pqp_oracle.h — CPU oracle for the path-QP solve path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this. The product (libpqp_b200.so) never links or calls it.
 *
 * PARITY UNPINNED: the arithmetic of this path lives in OSQP + osqp-eigen, third-party
 * libraries that are neither vendored nor version-pinned by the reference
 * (script/install_deps.sh:102,116 clone HEAD; the API used implies the 0.6.x line) and are
 * not present in this environment; the reference ships no test, golden vector or fixture
 * for this path. The restatement below follows
 *   - assembly:   /root/reference/src/solver/base_solver.cpp:15-39,119-261,290-296
 *   - epilogue:   base_solver.cpp:263-288, include/tools/tools.hpp:24-35
 *   - iteration:  the published OSQP algorithm (Stellato et al., Math. Prog. Comp. 2020)
 *                 with OSQP 0.6.x's scaling / rho / termination rules
 * and is validated by an independent KKT-optimality check (tests/test_oracle.py), not
 * by reference-produced numbers.
 */
#ifndef PQP_ORACLE_H_
#define PQP_ORACLE_H_

#include "../include/pqp.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pqo_ws pqo_ws;

/* BaseSolver ctor + solve() up to and including initSolver (base_solver.cpp:15-39,56-87):
 * assemble P, A, l, u for one instance and run osqp_setup (scaling, rho vector, KKT
 * factorisation). knots: PQP_NFIELDS fields, field f at knots[f*stride + i]. */
pqo_ws *pqo_setup(const pqp_params *prm, int n, int p, const double *knots, int stride,
                  const double *inst);
void pqo_free(pqo_ws *ws);

/* solver_.solve() (base_solver.cpp:88 / :110). Returns the PQP_* status. */
int pqo_solve(pqo_ws *ws);

/* updateProblemFormulationAndSolve up to the solve (base_solver.cpp:100-107): re-linearise
 * about (l, psi, k), osqp_update_bounds, osqp_update_A (same pattern). */
int pqo_update(pqo_ws *ws, const double *l, const double *psi, const double *k);

/* same, but the whole instance changes (receding-horizon window shift): new knots block and
 * per-instance scalars; the pattern (n, p) must be unchanged. */
int pqo_update_full(pqo_ws *ws, const double *knots, int stride, const double *inst);

int pqo_nv(const pqo_ws *ws);
int pqo_m(const pqo_ws *ws);
int pqo_iters(const pqo_ws *ws);
int pqo_status(const pqo_ws *ws);
int pqo_rho_updates(const pqo_ws *ws);
double pqo_rho(const pqo_ws *ws);
double pqo_cost(const pqo_ws *ws);
double pqo_pri_res(const pqo_ws *ws);
double pqo_dua_res(const pqo_ws *ws);
int pqo_nnz_L(const pqo_ws *ws);
/* unscaled primal x (nv), dual y (m), z (m) in the reference's index order */
void pqo_get_x(const pqo_ws *ws, double *x);
void pqo_get_y(const pqo_ws *ws, double *y);
void pqo_get_z(const pqo_ws *ws, double *z);
/* scaled internal iterates (what persists between the two solves) and the scaling */
void pqo_get_scaled_iterates(const pqo_ws *ws, double *x, double *z, double *y);
void pqo_get_scaling(const pqo_ws *ws, double *D, double *E, double *c);
/* l, psi, kappa, u per knot: sol[f*stride + i] */
void pqo_get_sol(const pqo_ws *ws, double *sol, int stride);
/* unscaled problem data: A in CSC (Ap nv+1, Ai/Ax nnz), l, u (m), diag(P) (nv) */
int pqo_nnz_A(const pqo_ws *ws);
void pqo_get_problem(const pqo_ws *ws, int *Ap, int *Ai, double *Ax, double *l, double *u,
                     double *Pdiag);

/* getOptimizedPath (base_solver.cpp:263-288): x, y, heading per knot from (l, psi). */
void pqo_frenet_to_cartesian(int n, const double *ref_x, const double *ref_y,
                             const double *ref_heading, const double *l, const double *psi,
                             double *out_x, double *out_y, double *out_heading);

/* Whole batch on `nthreads` host threads (one instance per thread at a time), same
 * buffers as the C ABI (pqp_batch_in / pqp_batch_out, host pointers). mode 0: cold solve;
 * mode 1: cold solve + re-linearise about the result + warm re-solve (the reference's
 * two-solve protocol, path_optimizer.cpp:138-153; outputs are those of the second solve
 * and iters is the sum). dense_assembly != 0 additionally performs the reference's dense
 * m x nv zero-fill + scan (base_solver.cpp:122,145,159,210) so its cost is included.
 * Returns wall seconds spent (steady clock), < 0 on error. */
double pqo_solve_batch(const pqp_params *prm, const pqp_batch_in *in, const pqp_batch_out *out,
                       int nthreads, int mode, int dense_assembly);
int pqo_max_threads(void);
int pqo_omp_probe(int nthreads);

/* OSQP's unscaled termination test of a given (x, y, z) per instance, in FP64, on the (P, A, l, u)
 * assembled here; rep[b*6] = {pri_res, eps_pri, dua_res, eps_dua, worst bound violation of z, 0.5 x'Px}. */
int pqo_termination_batch(const pqp_params *prm, const pqp_batch_in *in, const double *x_full,
                          const double *y_full, const double *z_full, double *rep, int nthreads);

/* Persistent per-path workspaces for warm re-solve sequences (receding horizon, configs[4]). */
typedef struct pqo_batch pqo_batch;
pqo_batch *pqo_batch_setup(const pqp_params *prm, const pqp_batch_in *in, int nthreads);
int pqo_batch_update_full(pqo_batch *pb, const pqp_batch_in *in, int nthreads);
double pqo_batch_solve(pqo_batch *pb, const pqp_batch_out *out, int nthreads);
void pqo_batch_free(pqo_batch *pb);

#ifdef __cplusplus
}
#endif
#endif
