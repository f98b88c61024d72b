// bounds_driver.cpp — TEST ONLY. Compiles the clearance-bounds kernel's per-thread source
// (path_optimizer_2_b200/csrc/pqp_bounds_core.cuh) for the host and loops over the tasks the
// CUDA grid would run, so the kernel arithmetic can be checked against the oracle on a
// GPU-less box. The product has no CPU path; nothing in the package references this file.
#include <cmath>
#include <cstdint>

#include "../../include/pqp_bounds.h"
#include "../../path_optimizer_2_b200/csrc/pqp_bounds_core.cuh"

extern "C" int pqb_emu_compute(const pqp_bounds_map *map, const pqp_bounds_params *p, const pqp_bounds_in *in,
                               const pqp_bounds_out *out) {
    pqb::MapView m;
    m.dist = map->distance;
    m.rows = map->rows;
    m.cols = map->cols;
    m.res = map->resolution;
    m.inv_res = 1.0 / map->resolution;
    m.half_lx = 0.5 * map->rows * map->resolution;
    m.half_ly = 0.5 * map->cols * map->resolution;
    m.cx = map->center_x;
    m.cy = map->center_y;
    pqb::Params P{p->front_length, p->rear_length, p->car_width, p->safety_margin, p->epsilon};
    const int n_max = in->n_max, k_max = in->k_max;
    for (int b = 0; b < in->batch; ++b) {
        const int nb = in->n[b] > n_max ? n_max : in->n[b];
        out->n_valid[b] = nb;
        const double *st = in->states + (size_t)b * PQP_STATE_ROWS * n_max;
        const double *sb = in->spline + (size_t)b * PQP_SPLINE_ROWS * k_max;
        pqb::SplineView sp{sb, sb + k_max, sb + 2 * k_max, sb + 3 * k_max, sb + 4 * k_max,
                           sb + 5 * k_max, sb + 6 * k_max, sb + 7 * k_max, sb + 8 * k_max, in->k[b]};
        double *ob = out->bounds + (size_t)b * PQP_BOUND_ROWS * n_max;
        for (int i = 0; i < nb; ++i)
            for (int anchor = 0; anchor < 3; ++anchor) {
                double lb, ub;
                pqb::anchor_bounds(m, P, sp, st[i], st[n_max + i], st[2 * n_max + i], st[3 * n_max + i], anchor, lb, ub);
                ob[(size_t)(2 * anchor) * n_max + i] = lb;
                ob[(size_t)(2 * anchor + 1) * n_max + i] = ub;
                if (anchor < 2) {
                    if (out->knots) {
                        double *kb = out->knots + (size_t)b * PQP_NFIELDS * n_max;
                        const int f = anchor == 0 ? PQP_F_B0_LB : PQP_F_B1_LB;
                        kb[(size_t)f * n_max + i] = lb;
                        kb[(size_t)(f + 1) * n_max + i] = ub;
                    }
                    if (std::fabs(ub - lb) < P.epsilon && i < out->n_valid[b]) out->n_valid[b] = i;
                }
            }
    }
    return 0;
}

extern "C" int pqb_emu_build_states(const pqp_states_in *in, const pqp_states_out *out) {
    const int n_max = in->n_max, k_max = in->k_max;
    for (int b = 0; b < in->batch; ++b) {
        const double *sb = in->spline + (size_t)b * PQP_SPLINE_ROWS * k_max;
        pqb::SplineView sp{sb, sb + k_max, sb + 2 * k_max, sb + 3 * k_max, sb + 4 * k_max,
                           sb + 5 * k_max, sb + 6 * k_max, sb + 7 * k_max, sb + 8 * k_max, in->k[b]};
        const int total = pqb::build_states(sp, in->max_s[b], in->delta_s_smaller, in->delta_s_larger,
                                            in->dynamic_segmentation != 0, n_max, out->total != nullptr,
                                            out->states + (size_t)b * PQP_STATE_ROWS * n_max, out->curvature + (size_t)b * n_max);
        out->n[b] = total < n_max ? total : n_max;
        if (out->total) out->total[b] = total;
    }
    return 0;
}
