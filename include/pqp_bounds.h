/* pqp_bounds.h — C ABI of the clearance-bounds front end ("next").
 *
 * Replaces, for a batch of reference paths over ONE obstacle map:
 *   ReferencePathImpl::updateBoundsImproved        src/data_struct/reference_path_impl.cpp:177-230
 *   ReferencePathImpl::getClearanceWithDirectionStrict                             ...:232-312
 *   getDirectionalProjectionByNewton               src/tools/tools.cpp:156-189
 *   tk::spline::operator() / deriv                 src/tools/spline.cpp:252-330
 *   Map::getObstacleDistance                       src/tools/Map.cpp:16-22 (grid_map bilinear lookup)
 *
 * One CUDA thread per (path, knot, anchor in {front, rear, centre}); FP64 arithmetic like the
 * reference; the float distance layer of the map stays resident in HBM/L2 for the life of the
 * handle. Output feeds pqp_solve directly (bounds can be written into the solver's knot block).
 * Plain pointers and sizes only; every function returns a PQP_* code from pqp.h and never throws.
 * There is no CPU fallback.
 */
#ifndef PQP_BOUNDS_H
#define PQP_BOUNDS_H

#include <stdint.h>

#include "pqp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* The map's "distance" layer as grid_map holds it (demo.cpp:98-113): float, rows x cols,
 * row-major [i * cols + j]; cell (i, j) is centred at
 *   (center_x + length_x / 2 - (i + 1/2) resolution, center_y + length_y / 2 - (j + 1/2) resolution)
 * with length_x = rows * resolution, length_y = cols * resolution. Host pointer; copied at create. */
typedef struct pqp_bounds_map {
    int32_t rows, cols;
    double resolution;
    double center_x, center_y;
    const float *distance;
} pqp_bounds_map;

/* The gflags the reference's bounds code reads (src/config/planning_flags.cpp). The search
 * constants hard-coded in reference_path_impl.cpp:237-277 (0.5 m radius, 0.3 m / 0.05 m steps,
 * 6 m range, 0.2 m minimum space) and tools.cpp:166 (20 Newton steps, 1e-5) are hard-coded here too. */
typedef struct pqp_bounds_params {
    double front_length;   /* 3.9  */
    double rear_length;    /* -1.0 */
    double car_width;      /* 2.0  */
    double safety_margin;  /* 0.3  */
    double epsilon;        /* 1e-6: isEqual() tolerance of the "path is blocked" test */
} pqp_bounds_params;

#define PQP_SPLINE_ROWS 9 /* abscissa; a, b, c, y of x(s); a, b, c, y of y(s)  (tk::spline m_x, m_a, m_b, m_c, m_y) */
#define PQP_STATE_ROWS 4  /* s, x, y, heading of a reference state */
#define PQP_BOUND_ROWS 6  /* front lb, front ub, rear lb, rear ub, centre lb, centre ub */

typedef struct pqp_bounds_in {
    int32_t batch, n_max, k_max;
    const double *states;   /* [b][PQP_STATE_ROWS][n_max]  reference_states_ */
    const int32_t *n;       /* [b] number of reference states */
    const double *spline;   /* [b][PQP_SPLINE_ROWS][k_max]  x_s_, y_s_ (same abscissae) */
    const int32_t *k;       /* [b] number of spline points (>= 3) */
} pqp_bounds_in;

typedef struct pqp_bounds_out {
    double *bounds;      /* [b][PQP_BOUND_ROWS][n_max]; entries at i >= n_valid[b] are unspecified */
    int32_t *n_valid;    /* [b] index of the first blocked state (= n[b] when none): the size
                            reference_states_ is cut to (reference_path_impl.cpp:222-229) */
    double *knots;       /* optional: solver knot block [b][PQP_NFIELDS][n_max]; receives the
                            front bounds in PQP_F_B0_LB/UB and the rear bounds in PQP_F_B1_LB/UB */
} pqp_bounds_out;

/* ReferencePathImpl::buildReferenceFromSpline (reference_path_impl.cpp:314-338):
 * walk each spline from s = 0 to max_s with the curvature-dependent step (delta_s_larger where
 * |k| < 0.08, delta_s_smaller where |k| > 0.2, linear in between; fixed delta_s_larger when
 * dynamic_segmentation = 0) and emit x, y, heading (tools.cpp:32-36), curvature (:38-44), s. */
typedef struct pqp_states_in {
    int32_t batch, n_max, k_max;
    const double *spline;   /* [b][PQP_SPLINE_ROWS][k_max] */
    const int32_t *k;       /* [b] */
    const double *max_s;    /* [b] max_s_ of the reference path */
    double delta_s_smaller; /* 0.15 (path_optimizer.cpp: buildReferenceFromSpline(0.15, 0.3)) */
    double delta_s_larger;  /* 0.3 */
    int32_t dynamic_segmentation; /* FLAGS_enable_dynamic_segmentation (default true) */
} pqp_states_in;

typedef struct pqp_states_out {
    double *states;     /* [b][PQP_STATE_ROWS][n_max]: s, x, y, heading */
    double *curvature;  /* [b][n_max] */
    int32_t *n;         /* [b] states written = min(total, n_max) */
    int32_t *total;     /* optional [b]: states the reference would emit (no n_max cap) */
    double *knots;      /* optional solver knot block [b][PQP_NFIELDS][n_max]: receives PQP_F_S,
                           PQP_F_KREF and the first linearisation point of path_optimizer.cpp:128-137
                           (PQP_F_L = PQP_F_PSI = 0, PQP_F_K = curvature) */
} pqp_states_out;

typedef struct pqp_bounds_handle pqp_bounds_handle;

void pqp_bounds_default_params(pqp_bounds_params *p);
int pqp_bounds_create(const pqp_bounds_map *map, const pqp_bounds_params *params, int32_t device,
                      pqp_bounds_handle **out);
void pqp_bounds_destroy(pqp_bounds_handle *h);
/* Host buffers in, host buffers out (H2D, kernel, D2H; synchronous). */
int pqp_bounds_compute(pqp_bounds_handle *h, const pqp_bounds_in *in, const pqp_bounds_out *out);
/* Device buffers; asynchronous on `stream` (a cudaStream_t passed as void*). */
int pqp_bounds_compute_device(pqp_bounds_handle *h, const pqp_bounds_in *in, const pqp_bounds_out *out,
                              void *stream);
/* buildReferenceFromSpline for a batch: host buffers (synchronous) / device buffers (async). */
int pqp_bounds_build_states(pqp_bounds_handle *h, const pqp_states_in *in, const pqp_states_out *out);
int pqp_bounds_build_states_device(pqp_bounds_handle *h, const pqp_states_in *in, const pqp_states_out *out,
                                   void *stream);
int pqp_bounds_last_kernel_ms(pqp_bounds_handle *h, float *ms);
const char *pqp_bounds_last_error(pqp_bounds_handle *h);

#ifdef __cplusplus
}
#endif
#endif
