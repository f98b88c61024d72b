// emu_driver.cpp — runs the kernel source (pqp_kernel.cuh) under the fiber warp emulator.
// TEST ONLY: lets `-m "not gpu"` tests exercise the kernel's arithmetic on the GPU-less
// build container. Not part of the product; the product library has no CPU path.
#define PQP_EMU 1
#include <cstdio>
#include <cstring>
#include <vector>

#include "../../path_optimizer_2_b200/csrc/pqp_host_common.h"

namespace {

struct EmuHandle {
    pqp_params prm;
    int n_max, batch_max, chunk;
    bool fp64 = false;  // params.reserved bit 1: run the FP64 instantiation
    bool incr = false;  // increment form of the ADMM step (bit 32; bit 64 forces the textbook form; default C >= 4)
    std::vector<double> warm, scal, dy, rho;   // sized for the wider scalar type
    std::vector<double> smem;
    bool solved = false;
};

template <int C, typename real, bool Incr>
void run_form(const pqp::KernelArgs &ka, int qp, real *smem) {
    warp_emu::Warp warp;
    warp_emu::current() = &warp;
    const double *src = ka.knots + (size_t)qp * 9 * ka.n_max;
    const int stride = ka.n_max;
    warp.run([&](int lane) {
        pqp::QpWarp<C, real, pqp::SmemStore<C, real>, Incr> w(ka, pqp::SmemStore<C, real>(smem, lane), lane, qp);
        w.run(src, stride, (size_t)(qp + ka.qp0));
    });
    warp_emu::current() = nullptr;
}
// both forms of the ADMM step are instantiated for both scalar types (the library only builds the
// increment form in FP32); `incr` as the library decides it: reserved bit 32 / 64, default C >= 4
template <int C, typename real>
void run_one(const pqp::KernelArgs &ka, int qp, real *smem, bool incr) {
    if (incr) run_form<C, real, true>(ka, qp, smem);
    else run_form<C, real, false>(ka, qp, smem);
}

int run_batch(EmuHandle *h, const pqp_batch_in *in, const pqp_batch_out *out, int mode) {
    if (!h || !in || !out || !in->knots || !in->inst || !in->n || !out->sol) return PQP_E_INVALID;
    if (in->batch > h->batch_max || in->n_max != h->n_max) return PQP_E_INVALID;
    if (mode == 1 && !h->solved) return PQP_E_STATE;
    for (int b = 0; b < in->batch; ++b)
        if (in->n[b] < 2 || in->n[b] > h->n_max) return PQP_E_INVALID;
    pqp::KernelArgs ka;
    ka.prm = pqp::make_dev_params(h->prm);
    ka.batch = in->batch;
    ka.n_max = in->n_max;
    ka.mode = mode;
    ka.use_tma = 0;
    ka.qp0 = 0;
    ka.knots = in->knots;
    ka.inst = in->inst;
    ka.n = in->n;
    ka.p = in->p;
    ka.sol = out->sol;
    ka.cost = out->cost;
    ka.status = out->status;
    ka.iters = out->iters;
    ka.flags = nullptr;
    ka.work_counter = nullptr;
    ka.ready = nullptr;
    ka.done = nullptr;
    ka.host_done = nullptr;
    ka.n_chunks = 0;
    ka.x_full = out->x_full;
    ka.y_full = out->y_full;
    ka.z_full = out->z_full;
    ka.info = out->info;
    ka.warm = h->warm.data();
    ka.scal = h->scal.data();
    ka.dy = h->dy.data();
    ka.rho_state = h->rho.data();
    for (int b = 0; b < in->batch; ++b) {
        std::fill(h->smem.begin(), h->smem.end(), 0.0);
        if (h->fp64) {
            double *sm = h->smem.data();
            switch (h->chunk) {
                case 1: run_one<1, double>(ka, b, sm, h->incr); break;
                case 2: run_one<2, double>(ka, b, sm, h->incr); break;
                case 4: run_one<4, double>(ka, b, sm, h->incr); break;
                case 8: run_one<8, double>(ka, b, sm, h->incr); break;
                default: run_one<16, double>(ka, b, sm, h->incr); break;
            }
        } else {
            float *sm = reinterpret_cast<float *>(h->smem.data());
            switch (h->chunk) {
                case 1: run_one<1, float>(ka, b, sm, h->incr); break;
                case 2: run_one<2, float>(ka, b, sm, h->incr); break;
                case 4: run_one<4, float>(ka, b, sm, h->incr); break;
                case 8: run_one<8, float>(ka, b, sm, h->incr); break;
                default: run_one<16, float>(ka, b, sm, h->incr); break;
            }
        }
    }
    h->solved = true;
    return PQP_OK;
}

}  // namespace

extern "C" {

void *emu_create(const pqp_params *prm, int n_max, int batch_max) {
    if (!prm || n_max < 2 || n_max > 32 * pqp::kMaxChunk - 1 || batch_max < 1 || !pqp::params_valid(*prm)) return nullptr;
    EmuHandle *h = new EmuHandle;
    h->prm = *prm;
    h->n_max = n_max;
    h->batch_max = batch_max;
    h->chunk = pqp::chunk_for(n_max);
    const int c = h->chunk;
    h->fp64 = (prm->reserved & 2) != 0;
    h->incr = (prm->reserved & 32) != 0 || ((prm->reserved & 64) == 0 && !h->fp64 && h->chunk >= 4);
    h->warm.assign((size_t)batch_max * pqp::warm_floats(c), 0.0);
    h->scal.assign((size_t)batch_max * pqp::scal_floats(c), 0.0);
    h->dy.assign((size_t)batch_max * pqp::dy_floats(c), 0.0);
    h->rho.assign(batch_max, 0.0);
    h->smem.assign(pqp::smem_floats(c), 0.0);
    return h;
}
void emu_destroy(void *h) { delete static_cast<EmuHandle *>(h); }
int emu_solve(void *h, const pqp_batch_in *in, const pqp_batch_out *out) {
    return run_batch(static_cast<EmuHandle *>(h), in, out, 0);
}
int emu_resolve(void *h, const pqp_batch_in *in, const pqp_batch_out *out) {
    return run_batch(static_cast<EmuHandle *>(h), in, out, 1);
}
int emu_chunk(void *h) { return static_cast<EmuHandle *>(h)->chunk; }
void emu_default_params(pqp_params *p) { pqp::default_params(p); }
}
