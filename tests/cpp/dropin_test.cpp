// Drives include/pqp_base_solver.hpp exactly the way PathOptimizer::optimizePath drives the
// reference's BaseSolver (path_optimizer.cpp:128-157): construct on the stack, solve(), then
// updateProblemFormulationAndSolve(*final_path, final_path) with input and output aliased.
// Input: a text file written by tests/test_dropin.py; output: one line per knot.
#include <cstdio>
#include <fstream>
#include <iostream>

#include "../../include/pqp_base_solver.hpp"
#include "ref_stub.hpp"

using Solver = pqp::dropin::BaseSolverT<stub::ReferencePath, stub::VehicleState, stub::SlState>;

int main(int argc, char **argv) {
    if (argc < 2) return 2;
    std::ifstream f(argv[1]);
    size_t n;
    f >> n;
    stub::ReferencePath ref;
    stub::VehicleState veh;
    ref.states_.resize(n);
    ref.bounds_.resize(n);
    for (size_t i = 0; i < n; ++i) {
        auto &s = ref.states_[i];
        auto &b = ref.bounds_[i];
        f >> s.s >> s.k >> s.x >> s.y >> s.heading >> b.front.lb >> b.front.ub >> b.rear.lb >> b.rear.ub;
        b.center.lb = b.front.lb;
        b.center.ub = b.front.ub;
    }
    f >> veh.offset_ >> veh.heading_error_ >> veh.start_.k >> veh.target_.heading;
    int constraint_end_heading;
    f >> constraint_end_heading;
    std::vector<stub::SlState> input_path;
    for (const auto &rs : ref.states_) {  // path_optimizer.cpp:128-137
        stub::SlState st;
        st.x = rs.x; st.y = rs.y; st.heading = rs.heading; st.s = rs.s; st.k = rs.k;
        input_path.push_back(st);
    }
    pqp::dropin::SolverFlags flags;
    flags.constraint_end_heading = constraint_end_heading != 0;
    Solver solver(ref, veh, input_path, flags);
    std::vector<stub::SlState> final_path;
    const bool ok1 = solver.solve(&final_path);
    std::printf("solve %d status %d iters %d cost %.17g err '%s'\n", ok1 ? 1 : 0, solver.lastStatus(),
                solver.lastIterations(), solver.lastCost(), solver.lastError().c_str());
    if (!ok1) return 0;
    const bool ok2 = solver.updateProblemFormulationAndSolve(final_path, &final_path);
    std::printf("resolve %d status %d iters %d cost %.17g\n", ok2 ? 1 : 0, solver.lastStatus(),
                solver.lastIterations(), solver.lastCost());
    if (!ok2) return 0;
    for (const auto &p : final_path)
        std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", p.x, p.y, p.heading, p.k, p.d_k, p.l, p.d_heading);
    return 0;
}
