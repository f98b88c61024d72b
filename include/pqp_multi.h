/*
 * pqp_multi.h — the batched path-QP solver over the GPUs of one box, behind the C ABI.
 *
 * the C++ host "stays as is" and is a single process, so the multi-GPU path is a
 * single-process one: one pqp_handle and one stream per device, the batch sharded in contiguous
 * blocks (every instance is independent - there is no data-path collective), and ONE exchange
 * afterwards: an ncclAllGather over NVLink of {cost f64, status i32, iters i32} = 16 B per instance, so
 * that every device holds the outcome of the whole batch (what a device-resident consumer - e.g. a
 * planner picking the cheapest of the candidate paths - needs). BASELINE configs[3]/[4].
 *
 * NCCL is opened at run time (dlopen("libnccl.so.2")): single-GPU users of libpqp_b200.so do not
 * need it, and a process that already loaded NCCL (torch) shares that copy. With n_devices == 1 no
 * NCCL call is made at all.
 *
 * reference interface replaced: the same BaseSolver::solve / updateProblemFormulationAndSolve pair
 * (base_solver.cpp:56-117) as pqp.h, for a batch that spans devices.
 */
#ifndef PQP_MULTI_H_
#define PQP_MULTI_H_

#include "pqp.h"

#ifdef __cplusplus
extern "C" {
#endif

/* one gathered record per instance (the layout of the all-gathered table) */
typedef struct pqp_result_rec {
    double cost;
    int32_t status;
    int32_t iters;
} pqp_result_rec;

typedef struct pqp_multi pqp_multi;

/* n_devices GPUs (devices[i], or 0..n_devices-1 when devices == NULL) that together hold up to
 * batch_max instances of up to n_max knots. Fails with PQP_E_NO_DEVICE when fewer devices exist or
 * (n_devices > 1) NCCL cannot be loaded. */
int pqp_multi_create(const pqp_params *params, int32_t n_max, int32_t batch_max, int32_t n_devices,
                     const int32_t *devices, pqp_multi **out);
int pqp_multi_destroy(pqp_multi *m);

/* pqp_solve / pqp_resolve for a batch sharded over the devices (HOST pointers, same structs): device
 * i takes instances [i * per, min(batch, (i + 1) * per)), per = ceil(batch / n_devices); the shards
 * run concurrently (one host thread per device around that device's own pipelined host call),
 * results land in `out` exactly as from a single device, then the records are packed on each device
 * (one kernel) and all-gathered. pqp_multi_resolve(in == NULL) re-linearises about the resident
 * solutions like pqp_resolve. */
int pqp_multi_solve(pqp_multi *m, const pqp_batch_in *in, const pqp_batch_out *out);
int pqp_multi_resolve(pqp_multi *m, const pqp_batch_in *in, const pqp_batch_out *out);

/* The gathered table on device index i after the last solve: *table is a DEVICE pointer (on that
 * device) to n_devices * per records, instance b at [ (b / per) * per + b % per ] = [b]; *per as
 * above. Valid until the next call on m. */
int pqp_multi_gathered(pqp_multi *m, int32_t device_index, const pqp_result_rec **table, int32_t *per);
/* shard of device index i in the last batch */
int pqp_multi_shard(pqp_multi *m, int32_t device_index, int32_t *first, int32_t *count);
/* the single-device handle behind device index i (e.g. for pqp_last_kernel_ms) */
int pqp_multi_handle(pqp_multi *m, int32_t device_index, pqp_handle **h);
/* device time (ms) of the last gather: pack kernels + ncclAllGather, max over devices */
int pqp_multi_last_gather_ms(pqp_multi *m, float *ms);
const char *pqp_multi_last_error(const pqp_multi *m); /* m may be NULL: last create error */

/* Building blocks, also used by one-process-per-GPU callers (bench.py under torchrun packs with this
 * kernel and gathers with torch.distributed): */
/* {cost, status, iters} of `batch` instances -> packed[batch] records, one launch, device pointers */
int pqp_pack_results_device(pqp_handle *h, int32_t batch, const double *cost, const int32_t *status,
                            const int32_t *iters, pqp_result_rec *packed, void *stream);
/* device pointers of the results the last host-pointer call left resident in the handle */
int pqp_resident_results(pqp_handle *h, const double **sol, const double **cost, const int32_t **status,
                         const int32_t **iters);

#ifdef __cplusplus
}
#endif
#endif /* PQP_MULTI_H_ */
