// smoother_driver.cpp — runs the smoother-QP routine of the CUDA library (pqp_smoother_core.cuh) on the CPU
// with one lane. TEST ONLY: lets `-m "not gpu"` tests check the kernel's source against the generic OSQP
// restatement (oracle/smoother_oracle.py) on the GPU-less build box. Not part of the product.
#include <cstring>
#include <vector>

#include "../../path_optimizer_2_b200/csrc/pqp_smoother_core.cuh"

namespace {
pqs::Settings settings(const double *s) {
    pqs::Settings st;
    st.rho = s[0]; st.sigma = s[1]; st.alpha = s[2]; st.eps_abs = s[3]; st.eps_rel = s[4]; st.eps_prim_inf = s[5];
    st.eps_dual_inf = s[6]; st.adaptive_rho_tolerance = s[7];
    st.max_iter = (int)s[8]; st.check_termination = (int)s[9]; st.scaling = (int)s[10]; st.adaptive_rho = (int)s[11];
    st.adaptive_rho_interval = (int)s[12];
    return st;
}
}  // namespace

extern "C" {
// info[6] = status, iters, rho_updates, pri_res, dua_res, obj
int smoother_emu_tension(int p, const double *xl, const double *yl, const double *al, const double *kl, const double *sl,
                         const double *weights, const double *set, double *rx, double *ry, double *rs, double *x_full,
                         double *info) {
    std::vector<unsigned char> scratch(pqs::scratch_bytes(p));
    std::vector<double> band(pqs::band_doubles(p));
    pqs::Work W;
    pqs::carve(W, scratch.data(), p, band.data());
    pqs::TensionWeights tw = {weights[0], weights[1], weights[2]};
    const pqs::Result R = pqs::tension_smooth(pqs::SerialLane(), W, settings(set), tw, p, xl, yl, al, kl, sl, rx, ry, rs, x_full);
    info[0] = R.status; info[1] = R.iters; info[2] = R.rho_updates; info[3] = R.pri_res; info[4] = R.dua_res; info[5] = R.obj;
    return 0;
}
int smoother_emu_post(int p, const double *layer_s, const double *lower, const double *upper, double vehicle_l,
                      const double *weights, const double *set, double *offsets, double *x_full, double *info) {
    std::vector<unsigned char> scratch(pqs::scratch_bytes(p));
    std::vector<double> band(pqs::band_doubles(p));
    pqs::Work W;
    pqs::carve(W, scratch.data(), p, band.data());
    pqs::PostWeights pw = {weights[0], weights[1], weights[2]};
    const pqs::Result R = pqs::post_smooth(pqs::SerialLane(), W, settings(set), pw, p, layer_s, lower, upper, vehicle_l, offsets, x_full);
    info[0] = R.status; info[1] = R.iters; info[2] = R.rho_updates; info[3] = R.pri_res; info[4] = R.dua_res; info[5] = R.obj;
    return 0;
}
}
