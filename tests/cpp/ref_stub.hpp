// Minimal stand-ins for the reference's boundary types, written for the tests of
// include/pqp_base_solver.hpp (field and accessor NAMES follow
// /root/reference/include/data_struct/data_struct.hpp:14-32,74-93, reference_path.hpp:21-46 and
// vehicle_state_frenet.hpp:13-37 so that the drop-in template binds to either).
#pragma once
#include <memory>
#include <vector>

namespace stub {
struct State {
    double x{}, y{}, heading{}, k{}, d_k{}, s{}, v{}, a{};
};
struct SlState : public State {
    double l{}, d_heading{};
};
struct VehicleStateBound {
    struct SingleBound {
        double ub{}, lb{}, x{}, y{}, heading{};
    } front, rear, center;
};
class ReferencePath {
 public:
    const std::vector<State> &getReferenceStates() const { return states_; }
    const std::vector<VehicleStateBound> &getBounds() const { return bounds_; }
    std::shared_ptr<VehicleStateBound> isBlocked() const { return blocked_; }
    double getLength() const { return states_.empty() ? 0.0 : states_.back().s; }
    std::vector<State> states_;
    std::vector<VehicleStateBound> bounds_;
    std::shared_ptr<VehicleStateBound> blocked_;
};
class VehicleState {
 public:
    VehicleState() = default;
    VehicleState(const State &start_state, const State &end_state, double offset = 0.0, double heading_error = 0.0)
        : start_(start_state), target_(end_state), offset_(offset), heading_error_(heading_error) {}
    const State &getStartState() const { return start_; }
    const State &getTargetState() const { return target_; }
    std::vector<double> getInitError() const { return {offset_, heading_error_}; }
    State start_, target_;
    double offset_{}, heading_error_{};
};
}  // namespace stub
