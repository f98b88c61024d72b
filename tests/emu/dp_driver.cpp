// dp_driver.cpp — runs the lattice-DP routine of the CUDA library (pqp_dp_core.cuh) on the CPU, one
// thread per path (SerialCtx). TEST ONLY: lets `-m "not gpu"` tests check the kernel's source against
// oracle/dp_oracle.py on the GPU-less build box. Not part of the product.
#include <cstdint>
#include <cstring>
#include <vector>

#include "../../path_optimizer_2_b200/csrc/pqp_dp_core.cuh"

extern "C" int dp_emu_search(const float *dist, int rows, int cols, double res, double lateral_range, double lateral_spacing,
                             double lon_spacing, double car_width, int batch, int k_max, int layers_max, const double *spline,
                             const int32_t *k, const double *length, const double *start, int32_t *status, int32_t *n_layers,
                             int32_t *n_out, double *layer_s, double *lower, double *upper, int32_t *chosen, double *vehicle_l,
                             double *target_s, double *cost, int8_t *parent, uint8_t *feasible) {
    pqb::MapView map;
    map.dist = dist;
    map.rows = rows;
    map.cols = cols;
    map.res = res;
    map.inv_res = 1.0 / res;
    map.half_lx = 0.5 * rows * res;
    map.half_ly = 0.5 * cols * res;
    map.cx = map.cy = 0.0;
    pqdp::Params prm = {lateral_range, lateral_spacing, lon_spacing, car_width};
    const int J = pqdp::lateral_count(prm);
    std::vector<double> nx((size_t)layers_max * J), ny(nx.size()), dis(nx.size()), ref((size_t)layers_max * 4);
    for (int b = 0; b < batch; ++b) {
        pqdp::PathIO io;
        const double *row = spline + (size_t)b * 9 * k_max;
        io.sp.sx = row;
        io.sp.xa = row + k_max; io.sp.xb = row + 2 * (size_t)k_max; io.sp.xc = row + 3 * (size_t)k_max; io.sp.xy = row + 4 * (size_t)k_max;
        io.sp.ya = row + 5 * (size_t)k_max; io.sp.yb = row + 6 * (size_t)k_max; io.sp.yc = row + 7 * (size_t)k_max; io.sp.yy = row + 8 * (size_t)k_max;
        io.sp.k = k[b];
        io.length = length[b];
        io.sx = start[3 * b]; io.sy = start[3 * b + 1]; io.sh = start[3 * b + 2];
        io.layers_max = layers_max;
        const size_t lo = (size_t)b * layers_max, t = lo * J;
        io.ok = status + b; io.n_layers = n_layers + b; io.n_out = n_out + b; io.chosen = chosen + lo;
        io.layer_s = layer_s + lo; io.lower = lower + lo; io.upper = upper + lo; io.vehicle_l = vehicle_l + b; io.target_s = target_s + b;
        io.cost = cost + t; io.parent = parent + t; io.feasible = feasible + t;
        io.nx = nx.data(); io.ny = ny.data(); io.dis = dis.data(); io.ref = ref.data();
        double shd[6 * pqdp::kJMax + 8];
        int shi[8];
        pqdp::dp_search_path(pqdp::SerialCtx(), map, prm, io, shd, shi);
    }
    return J;
}
