// pqp_base_solver.hpp — drop-in replacement for the reference's BaseSolver on top of the C ABI.
//
// Mirrors /root/reference/include/solver/base_solver.hpp:20-71 and
// /root/reference/src/solver/base_solver.cpp:
//   BaseSolver(const ReferencePath&, const VehicleState&, const std::vector<SlState>&)   :15-39
//   virtual bool solve(std::vector<SlState>*)                                            :56-95
//   virtual bool updateProblemFormulationAndSolve(const std::vector<SlState>&,
//                                                 std::vector<SlState>*)                 :97-117
// Same constructor arguments, same argument meaning, same bool error behaviour (false unless
// the solver reports SOLVED; nothing throws), same output fields filled
// (x, y, heading, k, d_k, l, d_heading; s/v/a left at 0 — base_solver.cpp:268-285), and it is
// safe under the caller's in-place aliasing of the second call (path_optimizer.cpp:153).
//
// The class is a template over the reference's own types so this header does not need (or
// copy) the reference's headers. Inside the reference tree one line binds it:
//
//   namespace PathOptimizationNS {
//   using BaseSolver = pqp::dropin::BaseSolverT<ReferencePath, VehicleState, SlState>;
//   }
//
// (see INTEGRATION.md). The host side only packs plain arrays; assembly of P/A/l/u, the OSQP
// iteration and the warm state live on the GPU behind include/pqp.h. There is no CPU path:
// without a B200 the constructor's handle creation fails and both calls return false.
#ifndef PQP_BASE_SOLVER_HPP_
#define PQP_BASE_SOLVER_HPP_

#include <algorithm>
#include <cmath>
#include <cstddef>
#include <string>
#include <vector>

#include "pqp.h"

namespace pqp {
namespace dropin {

// The gflags the reference's solver reads. Defaults = planning_flags.cpp.
struct SolverFlags {
    bool rough_constraints_far_away = false;  // planning_flags.cpp:112
    double precise_planning_length = 30.0;    // :114
    bool constraint_end_heading = true;       // :98
    int device = 0;                           // CUDA device ordinal
};

inline double constrainAngle(double angle) {  // include/tools/tools.hpp:24-35
    while (angle > M_PI) angle -= 2 * M_PI;
    while (angle < -M_PI) angle += 2 * M_PI;
    return angle;
}

template <class ReferencePathT, class VehicleStateT, class SlStateT>
class BaseSolverT {
 public:
    BaseSolverT() = delete;
    BaseSolverT(const ReferencePathT &reference_path, const VehicleStateT &vehicle_state,
                const std::vector<SlStateT> &input_path, const SolverFlags &flags = SolverFlags(),
                const pqp_params *params = nullptr)
        : n_(input_path.size()),
          reference_path_(reference_path),
          vehicle_state_(vehicle_state),
          input_path_(input_path),
          flags_(flags) {
        // base_solver.cpp:22-37
        precise_planning_size_ = n_;
        if (flags_.rough_constraints_far_away) {
            auto it = std::lower_bound(input_path.begin(), input_path.end(), flags_.precise_planning_length,
                                       [](const SlStateT &state, double s) { return state.s < s; });
            precise_planning_size_ = static_cast<size_t>(std::distance(input_path.begin(), it));
        }
        if (params) params_ = *params;
        else pqp_default_params(&params_);
        if (n_ >= 2 && n_ <= 255) {
            if (pqp_create(&params_, static_cast<int32_t>(n_), 1, flags_.device, &handle_) != PQP_OK) {
                error_ = pqp_last_error(nullptr);
                handle_ = nullptr;
            }
        } else {
            error_ = "number of knots must be in [2, 255]";  // the reference underflows for n < 2 (:23)
        }
    }
    BaseSolverT(const BaseSolverT &) = delete;
    BaseSolverT &operator=(const BaseSolverT &) = delete;
    virtual ~BaseSolverT() {
        if (handle_) pqp_destroy(handle_);
    }

    // base_solver.cpp:56-95
    virtual bool solve(std::vector<SlStateT> *optimized_path) {
        if (!handle_ || !optimized_path) return false;
        pack(input_path_);
        return run(false, optimized_path);
    }

    // base_solver.cpp:97-117. input_path and *optimized_path may be the same vector.
    virtual bool updateProblemFormulationAndSolve(const std::vector<SlStateT> &input_path,
                                                  std::vector<SlStateT> *optimized_path) {
        if (!handle_ || !optimized_path || input_path.size() != n_) return false;
        input_path_ = input_path;  // copy first: the output may alias the input (:100 vs :266)
        pack(input_path_);
        return run(true, optimized_path);
    }

    int lastStatus() const { return status_; }
    int lastIterations() const { return iters_; }
    double lastCost() const { return cost_; }
    const std::string &lastError() const { return error_; }

 private:
    // flat POD pack of everything setConstraints reads (base_solver.cpp:150-261)
    void pack(const std::vector<SlStateT> &lin) {
        const auto &ref_states = reference_path_.getReferenceStates();
        const auto &bounds = reference_path_.getBounds();
        const size_t n = n_, p = precise_planning_size_;
        knots_.assign(PQP_NFIELDS * n, 0.0);
        for (size_t i = 0; i < n; ++i) {
            knots_[PQP_F_S * n + i] = ref_states[i].s;
            knots_[PQP_F_KREF * n + i] = ref_states[i].k;
            knots_[PQP_F_L * n + i] = lin[i].l;
            knots_[PQP_F_PSI * n + i] = lin[i].d_heading;
            knots_[PQP_F_K * n + i] = lin[i].k;
            if (i < p) {
                knots_[PQP_F_B0_LB * n + i] = bounds[i].front.lb;
                knots_[PQP_F_B0_UB * n + i] = bounds[i].front.ub;
                knots_[PQP_F_B1_LB * n + i] = bounds[i].rear.lb;
                knots_[PQP_F_B1_UB * n + i] = bounds[i].rear.ub;
            } else {
                knots_[PQP_F_B0_LB * n + i] = bounds[i].center.lb;
                knots_[PQP_F_B0_UB * n + i] = bounds[i].center.ub;
            }
        }
        const auto init_error = vehicle_state_.getInitError();
        inst_[PQP_I_L0] = init_error[0];
        inst_[PQP_I_PSI0] = init_error[1];
        inst_[PQP_I_K0] = vehicle_state_.getStartState().k;
        inst_[PQP_I_EPSI_LO] = -1e30;  // OsqpEigen::INFTY (:252-253)
        inst_[PQP_I_EPSI_HI] = 1e30;
        if (flags_.constraint_end_heading && reference_path_.isBlocked() == nullptr) {
            const double end_psi = constrainAngle(vehicle_state_.getTargetState().heading - ref_states.back().heading);
            if (end_psi < 70 * M_PI / 180) {  // signed test, as in the reference (:256)
                inst_[PQP_I_EPSI_LO] = end_psi - 0.087;
                inst_[PQP_I_EPSI_HI] = end_psi + 0.087;
            }
        }
    }

    bool run(bool warm, std::vector<SlStateT> *optimized_path) {
        const int32_t n = static_cast<int32_t>(n_), p = static_cast<int32_t>(precise_planning_size_);
        sol_.assign(4 * n_, 0.0);
        pqp_batch_in in;
        in.batch = 1;
        in.n_max = n;
        in.knots = knots_.data();
        in.inst = inst_;
        in.n = &n;
        in.p = &p;
        pqp_batch_out out = {};
        out.sol = sol_.data();
        out.cost = &cost_;
        out.status = &status_;
        out.iters = &iters_;
        const int rc = warm ? pqp_resolve(handle_, &in, &out) : pqp_solve(handle_, &in, &out);
        if (rc != PQP_OK) {
            error_ = pqp_last_error(handle_);
            return false;
        }
        if (status_ != PQP_SOLVED) return false;  // osqp-eigen: solve() fails unless OSQP_SOLVED
        get_optimized_path(optimized_path);
        return true;
    }

    // base_solver.cpp:263-288
    void get_optimized_path(std::vector<SlStateT> *optimized_path) const {
        optimized_path->clear();
        const auto &ref_states = reference_path_.getReferenceStates();
        const size_t n = n_;
        for (size_t i = 0; i != n; ++i) {
            SlStateT pt;
            const double angle = ref_states[i].heading;
            const double l = sol_[i], psi = sol_[n + i];
            pt.heading = constrainAngle(angle + psi);
            pt.d_heading = psi;
            pt.l = l;
            const double new_angle = constrainAngle(angle + M_PI_2);
            pt.x = ref_states[i].x + l * std::cos(new_angle);
            pt.y = ref_states[i].y + l * std::sin(new_angle);
            pt.k = sol_[2 * n + i];
            if (i < n - 1) pt.d_k = sol_[3 * n + i];
            optimized_path->push_back(pt);
        }
    }

 protected:
    const size_t n_;
    size_t precise_planning_size_{};
    const ReferencePathT &reference_path_;
    const VehicleStateT &vehicle_state_;
    std::vector<SlStateT> input_path_;
    SolverFlags flags_;
    pqp_params params_{};
    pqp_handle *handle_ = nullptr;
    std::vector<double> knots_, sol_;
    double inst_[PQP_NINST] = {0, 0, 0, 0, 0};
    double cost_ = 0.0;
    int32_t status_ = PQP_UNSOLVED, iters_ = 0;
    std::string error_;
};

}  // namespace dropin
}  // namespace pqp
#endif  // PQP_BASE_SOLVER_HPP_
