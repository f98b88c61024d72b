// pqp_bounds.cu — clearance-bounds kernel and its C ABI (include/pqp_bounds.h).
//
// One thread per (path, reference state, anchor); a warp = one anchor type of 32 consecutive
// states, so its ray-marches walk neighbouring map cells in lock step. The
// float distance layer (2 MB for the demo map) is read through the read-only path and stays in
// the 126 MB L2; per state the kernel reads 32 B of state + its share of the spline and writes
// 48 B of bounds - it is gather-latency bound, not HBM bound.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstring>
#include <new>
#include <string>

#include "../../include/pqp_bounds.h"
#include "pqp_bounds_core.cuh"
#include "pqp_bounds_internal.h"
#include "pqp_device_guard.h"

namespace {

struct BoundsArgs {
    pqb::MapView map;
    pqb::Params prm;
    int batch, n_max, k_max;
    const double *states;
    const int32_t *n;
    const double *spline;
    const int32_t *k;
    double *bounds;
    int32_t *n_valid;
    double *knots;
};

__global__ void init_n_valid_kernel(const int32_t *__restrict__ n, int32_t *__restrict__ n_valid, int batch, int n_max) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) {
        int v = n[b];
        n_valid[b] = v < 0 ? 0 : (v > n_max ? n_max : v);
    }
}

// grid.x covers (path, state) pairs, grid.y the anchor: a warp handles one anchor type of 32
// consecutive states (0.15-0.3 m apart), so its lanes run the same Newton / march phases over
// neighbouring map cells.
__global__ void __launch_bounds__(128) clearance_bounds_kernel(const BoundsArgs a) {
    const long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const long long total = (long long)a.batch * a.n_max;
    if (t >= total) return;
    const int anchor = (int)blockIdx.y;
    const int i = (int)(t % a.n_max);
    const int b = (int)(t / a.n_max);
    int nb = a.n[b];
    nb = nb > a.n_max ? a.n_max : nb;
    if (i >= nb) return;
    const double *st = a.states + (size_t)b * PQP_STATE_ROWS * a.n_max;
    const double s = st[i], x = st[a.n_max + i], y = st[2 * a.n_max + i], heading = st[3 * a.n_max + i];
    const double *sb = a.spline + (size_t)b * PQP_SPLINE_ROWS * a.k_max;
    pqb::SplineView sp;
    sp.sx = sb;
    sp.xa = sb + 1 * (size_t)a.k_max; sp.xb = sb + 2 * (size_t)a.k_max; sp.xc = sb + 3 * (size_t)a.k_max; sp.xy = sb + 4 * (size_t)a.k_max;
    sp.ya = sb + 5 * (size_t)a.k_max; sp.yb = sb + 6 * (size_t)a.k_max; sp.yc = sb + 7 * (size_t)a.k_max; sp.yy = sb + 8 * (size_t)a.k_max;
    sp.k = a.k[b] > a.k_max ? a.k_max : a.k[b];
    double lb, ub;
    pqb::anchor_bounds(a.map, a.prm, sp, s, x, y, heading, anchor, lb, ub);
    double *ob = a.bounds + (size_t)b * PQP_BOUND_ROWS * a.n_max;
    ob[(size_t)(2 * anchor) * a.n_max + i] = lb;
    ob[(size_t)(2 * anchor + 1) * a.n_max + i] = ub;
    if (anchor < 2) {
        if (a.knots) {
            double *kb = a.knots + (size_t)b * PQP_NFIELDS * a.n_max;
            const int f = anchor == 0 ? PQP_F_B0_LB : PQP_F_B1_LB;
            kb[(size_t)f * a.n_max + i] = lb;
            kb[(size_t)(f + 1) * a.n_max + i] = ub;
        }
        // isEqual(bound[0], bound[1]) -> "Path is blocked" (reference_path_impl.cpp:216-220)
        if (fabs(ub - lb) < a.prm.epsilon) atomicMin(a.n_valid + b, i);
    }
}

struct StatesArgs {
    int batch, n_max, k_max;
    const double *spline;
    const int32_t *k;
    const double *max_s;
    double ds_small, ds_large;
    int dynamic;
    double *states, *curvature, *knots;
    int32_t *n, *total;
};

// one thread per path: the walk is sequential in s (each step depends on the curvature at the last)
__global__ void __launch_bounds__(64) reference_states_kernel(const StatesArgs a) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= a.batch) return;
    const double *sb = a.spline + (size_t)b * PQP_SPLINE_ROWS * a.k_max;
    pqb::SplineView sp;
    sp.sx = sb;
    sp.xa = sb + 1 * (size_t)a.k_max; sp.xb = sb + 2 * (size_t)a.k_max; sp.xc = sb + 3 * (size_t)a.k_max; sp.xy = sb + 4 * (size_t)a.k_max;
    sp.ya = sb + 5 * (size_t)a.k_max; sp.yb = sb + 6 * (size_t)a.k_max; sp.yc = sb + 7 * (size_t)a.k_max; sp.yy = sb + 8 * (size_t)a.k_max;
    sp.k = a.k[b] > a.k_max ? a.k_max : a.k[b];
    double *st = a.states + (size_t)b * PQP_STATE_ROWS * a.n_max;
    double *cv = a.curvature + (size_t)b * a.n_max;
    const int total = pqb::build_states(sp, a.max_s[b], a.ds_small, a.ds_large, a.dynamic != 0, a.n_max, a.total != nullptr, st, cv);
    const int nb = total < a.n_max ? total : a.n_max;
    a.n[b] = nb;
    if (a.total) a.total[b] = total;
    if (a.knots) {
        double *kb = a.knots + (size_t)b * PQP_NFIELDS * a.n_max;
        for (int i = 0; i < nb; ++i) {
            kb[(size_t)PQP_F_S * a.n_max + i] = st[i];
            kb[(size_t)PQP_F_KREF * a.n_max + i] = cv[i];
            kb[(size_t)PQP_F_L * a.n_max + i] = 0.0;
            kb[(size_t)PQP_F_PSI * a.n_max + i] = 0.0;
            kb[(size_t)PQP_F_K * a.n_max + i] = cv[i];
        }
    }
}

}  // namespace

struct pqp_bounds_handle {
    int device = 0;
    pqb::MapView map{};
    pqb::Params prm{};
    float *d_dist = nullptr;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // staging for the host-pointer call (grown on demand)
    double *d_states = nullptr, *d_spline = nullptr, *d_bounds = nullptr, *d_knots = nullptr;
    int32_t *d_n = nullptr, *d_k = nullptr, *d_nvalid = nullptr;
    double *d_maxs = nullptr, *d_curv = nullptr;
    int32_t *d_total = nullptr;
    size_t cap_maxs = 0, cap_curv = 0, cap_total = 0;
    size_t cap_states = 0, cap_bounds = 0, cap_spline = 0, cap_knots = 0, cap_n = 0, cap_k = 0, cap_nvalid = 0;
    bool timed = false;
    std::string err;
};

namespace pqb {
const MapView *handle_map(const pqp_bounds_handle *h) { return h ? &h->map : nullptr; }
int handle_device(const pqp_bounds_handle *h) { return h ? h->device : -1; }
}  // namespace pqb

namespace {

thread_local std::string g_bounds_create_error;

#define PQB_CUDA(h, call)                                                          \
    do {                                                                           \
        cudaError_t e_ = (call);                                                   \
        if (e_ != cudaSuccess) {                                                   \
            (h)->err = std::string(#call) + ": " + cudaGetErrorString(e_);         \
            return PQP_E_CUDA;                                                     \
        }                                                                          \
    } while (0)

int validate(pqp_bounds_handle *h, const pqp_bounds_in *in, const pqp_bounds_out *out) {
    if (!h) return PQP_E_INVALID;
    if (!in || !out || in->batch <= 0 || in->n_max <= 0 || in->k_max < 3 || !in->states || !in->n || !in->spline ||
        !in->k || !out->bounds || !out->n_valid) {
        h->err = "pqp_bounds: null pointer or non-positive size (batch, n_max > 0, k_max >= 3)";
        return PQP_E_INVALID;
    }
    return PQP_OK;
}

int launch(pqp_bounds_handle *h, const pqp_bounds_in *in, const pqp_bounds_out *out, cudaStream_t s) {
    BoundsArgs a;
    a.map = h->map;
    a.prm = h->prm;
    a.batch = in->batch;
    a.n_max = in->n_max;
    a.k_max = in->k_max;
    a.states = in->states;
    a.n = in->n;
    a.spline = in->spline;
    a.k = in->k;
    a.bounds = out->bounds;
    a.n_valid = out->n_valid;
    a.knots = out->knots;
    PQB_CUDA(h, cudaEventRecord(h->ev0, s));
    init_n_valid_kernel<<<(in->batch + 255) / 256, 256, 0, s>>>(in->n, out->n_valid, in->batch, in->n_max);
    const long long total = (long long)in->batch * in->n_max;
    const long long blocks = (total + 127) / 128;
    if (blocks > 0x7fffffffLL) {
        h->err = "pqp_bounds: batch * n_max too large for one launch";
        return PQP_E_INVALID;
    }
    clearance_bounds_kernel<<<dim3((unsigned)blocks, 3), 128, 0, s>>>(a);
    PQB_CUDA(h, cudaGetLastError());
    PQB_CUDA(h, cudaEventRecord(h->ev1, s));
    h->timed = true;
    return PQP_OK;
}

template <typename T>
cudaError_t grow(T **p, size_t *cap, size_t need) {
    if (need <= *cap) return cudaSuccess;
    if (*p) cudaFree(*p);
    *p = nullptr;
    *cap = 0;
    cudaError_t e = cudaMalloc(reinterpret_cast<void **>(p), need * sizeof(T));
    if (e == cudaSuccess) *cap = need;
    return e;
}

}  // namespace

extern "C" {

void pqp_bounds_default_params(pqp_bounds_params *p) {
    if (!p) return;
    p->front_length = 3.9;
    p->rear_length = -1.0;
    p->car_width = 2.0;
    p->safety_margin = 0.3;
    p->epsilon = 1e-6;
}

int pqp_bounds_create(const pqp_bounds_map *map, const pqp_bounds_params *params, int32_t device,
                      pqp_bounds_handle **out) {
    if (!out) return PQP_E_INVALID;
    *out = nullptr;
    if (!map || !map->distance || map->rows < 2 || map->cols < 2 || !(map->resolution > 0.0)) {
        g_bounds_create_error = "pqp_bounds_create: map needs rows, cols >= 2, resolution > 0 and a distance layer";
        return PQP_E_INVALID;
    }
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) {
        g_bounds_create_error = "pqp_bounds_create: no CUDA device (there is no CPU fallback)";
        return PQP_E_NO_DEVICE;
    }
    pqp_bounds_handle *h = new (std::nothrow) pqp_bounds_handle();
    if (!h) return PQP_E_CUDA;
    h->device = device;
    pqp_bounds_params dflt;
    pqp_bounds_default_params(&dflt);
    const pqp_bounds_params &p = params ? *params : dflt;
    h->prm.front_length = p.front_length;
    h->prm.rear_length = p.rear_length;
    h->prm.car_width = p.car_width;
    h->prm.safety_margin = p.safety_margin;
    h->prm.epsilon = p.epsilon;
    const size_t cells = (size_t)map->rows * map->cols;
    pqp::DeviceGuard guard_(device);
    cudaError_t e = cudaSuccess;
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&h->d_dist), cells * sizeof(float));
    if (e == cudaSuccess) e = cudaMemcpy(h->d_dist, map->distance, cells * sizeof(float), cudaMemcpyHostToDevice);
    if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev0);
    if (e == cudaSuccess) e = cudaEventCreate(&h->ev1);
    if (e != cudaSuccess) {
        g_bounds_create_error = std::string("pqp_bounds_create: ") + cudaGetErrorString(e);
        pqp_bounds_destroy(h);
        return PQP_E_CUDA;
    }
    h->map.dist = h->d_dist;
    h->map.rows = map->rows;
    h->map.cols = map->cols;
    h->map.res = map->resolution;
    h->map.inv_res = 1.0 / map->resolution;
    h->map.half_lx = 0.5 * map->rows * map->resolution;
    h->map.half_ly = 0.5 * map->cols * map->resolution;
    h->map.cx = map->center_x;
    h->map.cy = map->center_y;
    *out = h;
    return PQP_OK;
}

void pqp_bounds_destroy(pqp_bounds_handle *h) {
    if (!h) return;
    pqp::DeviceGuard guard_(h->device);
    cudaFree(h->d_dist);
    cudaFree(h->d_states);
    cudaFree(h->d_spline);
    cudaFree(h->d_bounds);
    cudaFree(h->d_knots);
    cudaFree(h->d_n);
    cudaFree(h->d_k);
    cudaFree(h->d_nvalid);
    cudaFree(h->d_maxs);
    cudaFree(h->d_curv);
    cudaFree(h->d_total);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    if (h->stream) cudaStreamDestroy(h->stream);
    delete h;
}

int pqp_bounds_compute_device(pqp_bounds_handle *h, const pqp_bounds_in *in, const pqp_bounds_out *out, void *stream) {
    int rc = validate(h, in, out);
    if (rc != PQP_OK) return rc;
    pqp::DeviceGuard guard_(h->device);
    return launch(h, in, out, static_cast<cudaStream_t>(stream));
}

int pqp_bounds_compute(pqp_bounds_handle *h, const pqp_bounds_in *in, const pqp_bounds_out *out) {
    int rc = validate(h, in, out);
    if (rc != PQP_OK) return rc;
    pqp::DeviceGuard guard_(h->device);
    const size_t B = in->batch, ns = B * PQP_STATE_ROWS * in->n_max, nsp = B * PQP_SPLINE_ROWS * in->k_max;
    const size_t nb = B * PQP_BOUND_ROWS * in->n_max, nk = out->knots ? B * PQP_NFIELDS * in->n_max : 0;
    PQB_CUDA(h, grow(&h->d_states, &h->cap_states, ns));
    PQB_CUDA(h, grow(&h->d_bounds, &h->cap_bounds, nb));
    PQB_CUDA(h, grow(&h->d_spline, &h->cap_spline, nsp));
    if (nk) PQB_CUDA(h, grow(&h->d_knots, &h->cap_knots, nk));
    PQB_CUDA(h, grow(&h->d_n, &h->cap_n, B));
    PQB_CUDA(h, grow(&h->d_k, &h->cap_k, B));
    PQB_CUDA(h, grow(&h->d_nvalid, &h->cap_nvalid, B));
    cudaStream_t s = h->stream;
    PQB_CUDA(h, cudaMemcpyAsync(h->d_states, in->states, ns * sizeof(double), cudaMemcpyHostToDevice, s));
    PQB_CUDA(h, cudaMemcpyAsync(h->d_spline, in->spline, nsp * sizeof(double), cudaMemcpyHostToDevice, s));
    PQB_CUDA(h, cudaMemcpyAsync(h->d_n, in->n, B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    PQB_CUDA(h, cudaMemcpyAsync(h->d_k, in->k, B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    if (nk) PQB_CUDA(h, cudaMemcpyAsync(h->d_knots, out->knots, nk * sizeof(double), cudaMemcpyHostToDevice, s));
    pqp_bounds_in din = *in;
    din.states = h->d_states;
    din.spline = h->d_spline;
    din.n = h->d_n;
    din.k = h->d_k;
    pqp_bounds_out dout;
    dout.bounds = h->d_bounds;
    dout.n_valid = h->d_nvalid;
    dout.knots = nk ? h->d_knots : nullptr;
    rc = launch(h, &din, &dout, s);
    if (rc != PQP_OK) return rc;
    PQB_CUDA(h, cudaMemcpyAsync(out->bounds, h->d_bounds, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
    PQB_CUDA(h, cudaMemcpyAsync(out->n_valid, h->d_nvalid, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (nk) PQB_CUDA(h, cudaMemcpyAsync(out->knots, h->d_knots, nk * sizeof(double), cudaMemcpyDeviceToHost, s));
    PQB_CUDA(h, cudaStreamSynchronize(s));
    return PQP_OK;
}

static int validate_states(pqp_bounds_handle *h, const pqp_states_in *in, const pqp_states_out *out) {
    if (!h) return PQP_E_INVALID;
    if (!in || !out || in->batch <= 0 || in->n_max <= 0 || in->k_max < 3 || !in->spline || !in->k || !in->max_s ||
        !out->states || !out->curvature || !out->n || !(in->delta_s_larger > 0.0) || !(in->delta_s_smaller > 0.0) ||
        in->delta_s_smaller > in->delta_s_larger) {
        h->err = "pqp_bounds_build_states: null pointer, non-positive size, or 0 < delta_s_smaller <= delta_s_larger violated";
        return PQP_E_INVALID;
    }
    return PQP_OK;
}

static int launch_states(pqp_bounds_handle *h, const pqp_states_in *in, const pqp_states_out *out, cudaStream_t s) {
    StatesArgs a;
    a.batch = in->batch;
    a.n_max = in->n_max;
    a.k_max = in->k_max;
    a.spline = in->spline;
    a.k = in->k;
    a.max_s = in->max_s;
    a.ds_small = in->delta_s_smaller;
    a.ds_large = in->delta_s_larger;
    a.dynamic = in->dynamic_segmentation;
    a.states = out->states;
    a.curvature = out->curvature;
    a.knots = out->knots;
    a.n = out->n;
    a.total = out->total;
    PQB_CUDA(h, cudaEventRecord(h->ev0, s));
    reference_states_kernel<<<(in->batch + 63) / 64, 64, 0, s>>>(a);
    PQB_CUDA(h, cudaGetLastError());
    PQB_CUDA(h, cudaEventRecord(h->ev1, s));
    h->timed = true;
    return PQP_OK;
}

int pqp_bounds_build_states_device(pqp_bounds_handle *h, const pqp_states_in *in, const pqp_states_out *out, void *stream) {
    int rc = validate_states(h, in, out);
    if (rc != PQP_OK) return rc;
    pqp::DeviceGuard guard_(h->device);
    return launch_states(h, in, out, static_cast<cudaStream_t>(stream));
}

int pqp_bounds_build_states(pqp_bounds_handle *h, const pqp_states_in *in, const pqp_states_out *out) {
    int rc = validate_states(h, in, out);
    if (rc != PQP_OK) return rc;
    pqp::DeviceGuard guard_(h->device);
    const size_t B = in->batch, ns = B * PQP_STATE_ROWS * in->n_max, nsp = B * PQP_SPLINE_ROWS * in->k_max;
    const size_t nc = B * in->n_max, nk = out->knots ? B * PQP_NFIELDS * in->n_max : 0;
    PQB_CUDA(h, grow(&h->d_states, &h->cap_states, ns));
    PQB_CUDA(h, grow(&h->d_curv, &h->cap_curv, nc));
    PQB_CUDA(h, grow(&h->d_spline, &h->cap_spline, nsp));
    if (nk) PQB_CUDA(h, grow(&h->d_knots, &h->cap_knots, nk));
    PQB_CUDA(h, grow(&h->d_n, &h->cap_n, B));
    PQB_CUDA(h, grow(&h->d_k, &h->cap_k, B));
    PQB_CUDA(h, grow(&h->d_total, &h->cap_total, B));
    PQB_CUDA(h, grow(&h->d_maxs, &h->cap_maxs, B));
    cudaStream_t s = h->stream;
    PQB_CUDA(h, cudaMemcpyAsync(h->d_spline, in->spline, nsp * sizeof(double), cudaMemcpyHostToDevice, s));
    PQB_CUDA(h, cudaMemcpyAsync(h->d_k, in->k, B * sizeof(int32_t), cudaMemcpyHostToDevice, s));
    PQB_CUDA(h, cudaMemcpyAsync(h->d_maxs, in->max_s, B * sizeof(double), cudaMemcpyHostToDevice, s));
    PQB_CUDA(h, cudaMemsetAsync(h->d_states, 0, ns * sizeof(double), s));
    PQB_CUDA(h, cudaMemsetAsync(h->d_curv, 0, nc * sizeof(double), s));
    if (nk) PQB_CUDA(h, cudaMemcpyAsync(h->d_knots, out->knots, nk * sizeof(double), cudaMemcpyHostToDevice, s));
    pqp_states_in din = *in;
    din.spline = h->d_spline;
    din.k = h->d_k;
    din.max_s = h->d_maxs;
    pqp_states_out dout;
    dout.states = h->d_states;
    dout.curvature = h->d_curv;
    dout.n = h->d_n;
    dout.total = h->d_total;
    dout.knots = nk ? h->d_knots : nullptr;
    rc = launch_states(h, &din, &dout, s);
    if (rc != PQP_OK) return rc;
    PQB_CUDA(h, cudaMemcpyAsync(out->states, h->d_states, ns * sizeof(double), cudaMemcpyDeviceToHost, s));
    PQB_CUDA(h, cudaMemcpyAsync(out->curvature, h->d_curv, nc * sizeof(double), cudaMemcpyDeviceToHost, s));
    PQB_CUDA(h, cudaMemcpyAsync(out->n, h->d_n, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (out->total) PQB_CUDA(h, cudaMemcpyAsync(out->total, h->d_total, B * sizeof(int32_t), cudaMemcpyDeviceToHost, s));
    if (nk) PQB_CUDA(h, cudaMemcpyAsync(out->knots, h->d_knots, nk * sizeof(double), cudaMemcpyDeviceToHost, s));
    PQB_CUDA(h, cudaStreamSynchronize(s));
    return PQP_OK;
}

int pqp_bounds_last_kernel_ms(pqp_bounds_handle *h, float *ms) {
    if (!h || !ms) return PQP_E_INVALID;
    if (!h->timed) {
        *ms = 0.0f;
        return PQP_OK;
    }
    PQB_CUDA(h, cudaEventSynchronize(h->ev1));
    PQB_CUDA(h, cudaEventElapsedTime(ms, h->ev0, h->ev1));
    return PQP_OK;
}

const char *pqp_bounds_last_error(pqp_bounds_handle *h) { return h ? h->err.c_str() : g_bounds_create_error.c_str(); }

}  // extern "C"
