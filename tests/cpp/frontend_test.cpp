// Drives pqp::dropin::ReferenceFrontEndT the way ReferencePathImpl does: buildReferenceFromSpline,
// then updateBoundsImproved. Input file: rows cols res; <rows*cols floats>; k; 9 rows of k doubles;
// max_s. Output: "states <ok> <n>", "bounds <ok> <n_after> <blocked 0|1>", then per state
// s x y heading k f_lb f_ub r_lb r_ub c_lb c_ub front_x front_y.
#include <cstdio>
#include <memory>
#include <vector>

#include "../../include/pqp_reference_path.hpp"
#include "ref_stub.hpp"

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    FILE *f = std::fopen(argv[1], "r");
    if (!f) return 2;
    int rows, cols, k;
    double res, max_s;
    if (std::fscanf(f, "%d %d %lf", &rows, &cols, &res) != 3) return 2;
    std::vector<float> dist(static_cast<size_t>(rows) * cols);
    for (auto &v : dist)
        if (std::fscanf(f, "%f", &v) != 1) return 2;
    if (std::fscanf(f, "%d", &k) != 1) return 2;
    pqp::dropin::SplineRows xs, ys;
    std::vector<double> *r[9] = {&xs.x, &xs.a, &xs.b, &xs.c, &xs.y, &ys.a, &ys.b, &ys.c, &ys.y};
    for (auto *row : r) {
        row->resize(k);
        for (auto &v : *row)
            if (std::fscanf(f, "%lf", &v) != 1) return 2;
    }
    ys.x = xs.x;
    if (std::fscanf(f, "%lf", &max_s) != 1) return 2;
    std::fclose(f);
    pqp::dropin::ReferenceFrontEndT<stub::State, stub::VehicleStateBound> fe(dist.data(), rows, cols, res, 0.0, 0.0);
    std::vector<stub::State> states;
    const bool ok1 = fe.buildReferenceFromSpline(xs, ys, max_s, 0.15, 0.3, true, &states);
    std::printf("states %d %zu %s\n", ok1 ? 1 : 0, states.size(), ok1 ? "" : fe.lastError().c_str());
    const std::vector<stub::State> all = states;
    std::vector<stub::VehicleStateBound> bounds;
    std::shared_ptr<stub::VehicleStateBound> blocked;
    const bool ok2 = fe.updateBoundsImproved(xs, ys, &states, &bounds, &blocked);
    std::printf("bounds %d %zu %d %s\n", ok2 ? 1 : 0, states.size(), blocked ? 1 : 0, ok2 ? "" : fe.lastError().c_str());
    for (size_t i = 0; i < bounds.size(); ++i) {
        const stub::State &s = states[i];
        const stub::VehicleStateBound &b = bounds[i];
        std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", s.s, s.x, s.y, s.heading,
                    s.k, b.front.lb, b.front.ub, b.rear.lb, b.rear.ub, b.center.lb, b.center.ub, b.front.x, b.front.y);
    }
    return 0;
}
