// pqp_multi.cu — include/pqp_multi.h: one process, the GPUs of one box, contiguous shards, one
// ncclAllGather of 16 B per instance. Built on the public single-device ABI (pqp.h) only.
//
// NCCL is bound at run time (dlopen): no link-time dependency, and a process that already carries a
// libnccl.so.2 (torch's) shares it.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "../../include/pqp_multi.h"
#include "pqp_device_guard.h"

namespace {

thread_local std::string g_multi_create_error;

// the handful of NCCL entry points this file uses (signatures of nccl.h 2.x; ncclInt8 == 0)
typedef struct ncclComm *ncclComm_t;
struct Nccl {
    void *lib = nullptr;
    int (*CommInitAll)(ncclComm_t *, int, const int *) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllGather)(const void *, void *, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char *(*GetErrorString)(int) = nullptr;
    bool load(std::string *err) {
        if (lib) return true;
        for (const char *name : {"libnccl.so.2", "libnccl.so"}) {
            lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
            if (lib) break;
        }
        if (!lib) {
            *err = std::string("cannot load NCCL: ") + dlerror();
            return false;
        }
        CommInitAll = reinterpret_cast<decltype(CommInitAll)>(dlsym(lib, "ncclCommInitAll"));
        CommDestroy = reinterpret_cast<decltype(CommDestroy)>(dlsym(lib, "ncclCommDestroy"));
        AllGather = reinterpret_cast<decltype(AllGather)>(dlsym(lib, "ncclAllGather"));
        GroupStart = reinterpret_cast<decltype(GroupStart)>(dlsym(lib, "ncclGroupStart"));
        GroupEnd = reinterpret_cast<decltype(GroupEnd)>(dlsym(lib, "ncclGroupEnd"));
        GetErrorString = reinterpret_cast<decltype(GetErrorString)>(dlsym(lib, "ncclGetErrorString"));
        if (!CommInitAll || !CommDestroy || !AllGather || !GroupStart || !GroupEnd || !GetErrorString) {
            *err = "libnccl.so.2 lacks an expected symbol";
            return false;
        }
        return true;
    }
};
Nccl &nccl() {
    static Nccl n;
    return n;
}

}  // namespace

struct pqp_multi {
    int n_dev = 0, n_max = 0, batch_max = 0, per_max = 0;
    std::vector<int> dev;
    std::vector<pqp_handle *> h;
    std::vector<cudaStream_t> stream;
    std::vector<cudaEvent_t> ev0, ev1;
    std::vector<pqp_result_rec *> d_send, d_table;  // per device: its packed shard / the gathered table
    std::vector<ncclComm_t> comm;
    int last_batch = 0, last_per = 0;
    float gather_ms = 0.0f;
    std::string err;
};

namespace {

int mfail(pqp_multi *m, int code, const std::string &msg) {
    if (m) m->err = msg;
    else g_multi_create_error = msg;
    return code;
}

int run(pqp_multi *m, const pqp_batch_in *in, const pqp_batch_out *out, bool warm) {
    if (!m) return PQP_E_INVALID;
    if (!out || !out->sol) return mfail(m, PQP_E_INVALID, "null output");
    int B;
    if (in) {
        if (!in->knots || !in->inst || !in->n) return mfail(m, PQP_E_INVALID, "null buffer");
        if (in->batch < 1 || in->batch > m->batch_max) return mfail(m, PQP_E_INVALID, "batch out of range");
        if (in->n_max != m->n_max) return mfail(m, PQP_E_INVALID, "n_max differs from the handle's");
        B = in->batch;
    } else {
        if (!warm) return mfail(m, PQP_E_INVALID, "null batch");
        B = m->last_batch;
        if (B < 1) return mfail(m, PQP_E_STATE, "resolve(NULL) needs a previous solve");
    }
    if (warm && B != m->last_batch) return mfail(m, PQP_E_STATE, "resolve needs a previous solve of the same batch");
    const int G = m->n_dev, nmax = m->n_max;
    const int per = (B + G - 1) / G;
    const size_t nvm = 6 * (size_t)nmax - 1, mm = 6 * (size_t)nmax + 2;
    std::vector<int> rc(G, PQP_OK);
    std::vector<std::thread> workers;
    // one host thread per device: each runs that device's own (pipelined, synchronous) host call
    auto work = [&](int d) {
        const int lo = std::min(B, d * per), hi = std::min(B, lo + per);
        if (lo >= hi) return;
        pqp_batch_in din;
        pqp_batch_out dout = {};
        dout.sol = out->sol + (size_t)lo * 4 * nmax;
        dout.cost = out->cost ? out->cost + lo : nullptr;
        dout.status = out->status ? out->status + lo : nullptr;
        dout.iters = out->iters ? out->iters + lo : nullptr;
        dout.x_full = out->x_full ? out->x_full + (size_t)lo * nvm : nullptr;
        dout.y_full = out->y_full ? out->y_full + (size_t)lo * mm : nullptr;
        dout.z_full = out->z_full ? out->z_full + (size_t)lo * mm : nullptr;
        dout.info = out->info ? out->info + (size_t)lo * PQP_NINFO : nullptr;
        if (in) {
            din.batch = hi - lo;
            din.n_max = nmax;
            din.knots = in->knots + (size_t)lo * PQP_NFIELDS * nmax;
            din.inst = in->inst + (size_t)lo * PQP_NINST;
            din.n = in->n + lo;
            din.p = in->p ? in->p + lo : nullptr;
        }
        rc[d] = warm ? pqp_resolve(m->h[d], in ? &din : nullptr, &dout) : pqp_solve(m->h[d], &din, &dout);
    };
    for (int d = 1; d < G; ++d) workers.emplace_back(work, d);
    work(0);
    for (auto &t : workers) t.join();
    for (int d = 0; d < G; ++d)
        if (rc[d] != PQP_OK) return mfail(m, rc[d], std::string("device ") + std::to_string(m->dev[d]) + ": " + pqp_last_error(m->h[d]));
    m->last_batch = B;
    m->last_per = per;
    // pack on every device, then one all-gather of per * 16 bytes per device
    for (int d = 0; d < G; ++d) {
        pqp::DeviceGuard g(m->dev[d]);
        const int lo = std::min(B, d * per), hi = std::min(B, lo + per);
        cudaEventRecord(m->ev0[d], m->stream[d]);
        cudaMemsetAsync(m->d_send[d], 0, (size_t)per * sizeof(pqp_result_rec), m->stream[d]);
        if (hi > lo) {
            const double *cost;
            const int32_t *status, *iters;
            int r = pqp_resident_results(m->h[d], nullptr, &cost, &status, &iters);
            if (r == PQP_OK) r = pqp_pack_results_device(m->h[d], hi - lo, cost, status, iters, m->d_send[d], m->stream[d]);
            if (r != PQP_OK) return mfail(m, r, pqp_last_error(m->h[d]));
        }
    }
    if (G == 1) {
        pqp::DeviceGuard g(m->dev[0]);
        cudaMemcpyAsync(m->d_table[0], m->d_send[0], (size_t)per * sizeof(pqp_result_rec), cudaMemcpyDeviceToDevice, m->stream[0]);
    } else {
        Nccl &N = nccl();
        int e = N.GroupStart();
        for (int d = 0; d < G && e == 0; ++d)
            e = N.AllGather(m->d_send[d], m->d_table[d], (size_t)per * sizeof(pqp_result_rec), /*ncclInt8*/ 0, m->comm[d], m->stream[d]);
        const int e2 = N.GroupEnd();
        if (e == 0) e = e2;
        if (e != 0) return mfail(m, PQP_E_CUDA, std::string("ncclAllGather: ") + N.GetErrorString(e));
    }
    float worst = 0.0f;
    for (int d = 0; d < G; ++d) {
        pqp::DeviceGuard g(m->dev[d]);
        cudaEventRecord(m->ev1[d], m->stream[d]);
        const cudaError_t ce = cudaStreamSynchronize(m->stream[d]);
        if (ce != cudaSuccess) return mfail(m, PQP_E_CUDA, std::string("gather: ") + cudaGetErrorString(ce));
        float ms = 0.0f;
        cudaEventElapsedTime(&ms, m->ev0[d], m->ev1[d]);
        worst = std::max(worst, ms);
    }
    m->gather_ms = worst;
    return PQP_OK;
}

}  // namespace

extern "C" {

const char *pqp_multi_last_error(const pqp_multi *m) { return m ? m->err.c_str() : g_multi_create_error.c_str(); }

int pqp_multi_create(const pqp_params *params, int32_t n_max, int32_t batch_max, int32_t n_devices,
                     const int32_t *devices, pqp_multi **out) {
    if (!out) return mfail(nullptr, PQP_E_INVALID, "out is null");
    *out = nullptr;
    if (!params || n_devices < 1 || batch_max < 1) return mfail(nullptr, PQP_E_INVALID, "invalid arguments");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev < n_devices)
        return mfail(nullptr, PQP_E_NO_DEVICE, "fewer CUDA devices than n_devices (there is no CPU fallback)");
    if (n_devices > 1 && !nccl().load(&g_multi_create_error)) return PQP_E_NO_DEVICE;
    pqp_multi *m = new (std::nothrow) pqp_multi;
    if (!m) return mfail(nullptr, PQP_E_INVALID, "out of host memory");
    m->n_dev = n_devices;
    m->n_max = n_max;
    m->batch_max = batch_max;
    m->per_max = (batch_max + n_devices - 1) / n_devices;
    const size_t G = n_devices;
    m->dev.resize(G);
    m->h.assign(G, nullptr);
    m->stream.assign(G, nullptr);
    m->ev0.assign(G, nullptr);
    m->ev1.assign(G, nullptr);
    m->d_send.assign(G, nullptr);
    m->d_table.assign(G, nullptr);
    m->comm.assign(G, nullptr);
    for (int d = 0; d < n_devices; ++d) m->dev[d] = devices ? devices[d] : d;
    for (int d = 0; d < n_devices; ++d) {
        int rc = pqp_create(params, n_max, m->per_max, m->dev[d], &m->h[d]);
        if (rc != PQP_OK) {
            g_multi_create_error = pqp_last_error(nullptr);
            pqp_multi_destroy(m);
            return rc;
        }
        pqp::DeviceGuard g(m->dev[d]);
        cudaError_t e = cudaStreamCreateWithFlags(&m->stream[d], cudaStreamNonBlocking);
        if (e == cudaSuccess) e = cudaEventCreate(&m->ev0[d]);
        if (e == cudaSuccess) e = cudaEventCreate(&m->ev1[d]);
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_send[d]), (size_t)m->per_max * sizeof(pqp_result_rec));
        if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void **>(&m->d_table[d]), G * m->per_max * sizeof(pqp_result_rec));
        if (e != cudaSuccess) {
            g_multi_create_error = std::string("pqp_multi_create: ") + cudaGetErrorString(e);
            pqp_multi_destroy(m);
            return PQP_E_CUDA;
        }
    }
    if (n_devices > 1) {
        const int e = nccl().CommInitAll(m->comm.data(), n_devices, m->dev.data());
        if (e != 0) {
            g_multi_create_error = std::string("ncclCommInitAll: ") + nccl().GetErrorString(e);
            for (auto &c : m->comm) c = nullptr;
            pqp_multi_destroy(m);
            return PQP_E_CUDA;
        }
    }
    *out = m;
    return PQP_OK;
}

int pqp_multi_destroy(pqp_multi *m) {
    if (!m) return PQP_OK;
    for (int d = 0; d < m->n_dev; ++d) {
        pqp::DeviceGuard g(m->dev[d]);
        if (m->stream[d]) cudaStreamSynchronize(m->stream[d]);
        if (m->comm[d]) nccl().CommDestroy(m->comm[d]);
        if (m->h[d]) pqp_destroy(m->h[d]);
        cudaFree(m->d_send[d]);
        cudaFree(m->d_table[d]);
        if (m->ev0[d]) cudaEventDestroy(m->ev0[d]);
        if (m->ev1[d]) cudaEventDestroy(m->ev1[d]);
        if (m->stream[d]) cudaStreamDestroy(m->stream[d]);
    }
    delete m;
    return PQP_OK;
}

int pqp_multi_solve(pqp_multi *m, const pqp_batch_in *in, const pqp_batch_out *out) {
    if (!m) return PQP_E_INVALID;
    if (!in) return mfail(m, PQP_E_INVALID, "null batch");
    return run(m, in, out, false);
}

int pqp_multi_resolve(pqp_multi *m, const pqp_batch_in *in, const pqp_batch_out *out) { return run(m, in, out, true); }

int pqp_multi_gathered(pqp_multi *m, int32_t i, const pqp_result_rec **table, int32_t *per) {
    if (!m || i < 0 || i >= m->n_dev) return PQP_E_INVALID;
    if (m->last_batch < 1) return mfail(m, PQP_E_STATE, "nothing solved yet");
    if (table) *table = m->d_table[i];
    if (per) *per = m->last_per;
    return PQP_OK;
}

int pqp_multi_shard(pqp_multi *m, int32_t i, int32_t *first, int32_t *count) {
    if (!m || i < 0 || i >= m->n_dev) return PQP_E_INVALID;
    const int lo = std::min(m->last_batch, i * m->last_per), hi = std::min(m->last_batch, lo + m->last_per);
    if (first) *first = lo;
    if (count) *count = hi - lo;
    return PQP_OK;
}

int pqp_multi_handle(pqp_multi *m, int32_t i, pqp_handle **h) {
    if (!m || !h || i < 0 || i >= m->n_dev) return PQP_E_INVALID;
    *h = m->h[i];
    return PQP_OK;
}

int pqp_multi_last_gather_ms(pqp_multi *m, float *ms) {
    if (!m || !ms) return PQP_E_INVALID;
    *ms = m->gather_ms;
    return PQP_OK;
}

}  // extern "C"
