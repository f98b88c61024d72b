/* multi_test.c — plain C against include/pqp_multi.h: shards one batch over G devices of the box
 * in a single process (the way the reference's C++ host would), and checks that
 *   (1) every output equals the single-device pqp_solve / pqp_resolve of the same batch, bit for bit;
 *   (2) the all-gathered {cost, status, iters} table on EVERY device equals those outputs.
 * Input: a binary file written by tests/test_multi.py: int32 B, int32 n_max, then knots, inst, n.
 * Usage: multi_test <file> <n_devices>.  Prints "ok ..." or "FAIL ...". */
#include <cuda_runtime_api.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/pqp_multi.h"

static void *xmalloc(size_t n) {
    void *p = calloc(n ? n : 1, 1);
    if (!p) { fprintf(stderr, "out of memory\n"); exit(2); }
    return p;
}

int main(int argc, char **argv) {
    if (argc < 3) return 2;
    FILE *f = fopen(argv[1], "rb");
    if (!f) return 2;
    const int G = atoi(argv[2]);
    int32_t B, nmax;
    if (fread(&B, 4, 1, f) != 1 || fread(&nmax, 4, 1, f) != 1) return 2;
    const size_t kn = (size_t)B * PQP_NFIELDS * nmax, in_ = (size_t)B * PQP_NINST, so = (size_t)B * 4 * nmax;
    double *knots = xmalloc(kn * 8), *inst = xmalloc(in_ * 8);
    int32_t *n = xmalloc((size_t)B * 4);
    if (fread(knots, 8, kn, f) != kn || fread(inst, 8, in_, f) != in_ || fread(n, 4, (size_t)B, f) != (size_t)B) return 2;
    fclose(f);
    pqp_params prm;
    pqp_default_params(&prm);
    pqp_batch_in bin = {B, nmax, knots, inst, n, NULL};
    double *sol1 = xmalloc(so * 8), *cost1 = xmalloc((size_t)B * 8), *solm = xmalloc(so * 8), *costm = xmalloc((size_t)B * 8);
    int32_t *st1 = xmalloc((size_t)B * 4), *it1 = xmalloc((size_t)B * 4), *stm = xmalloc((size_t)B * 4), *itm = xmalloc((size_t)B * 4);
    pqp_batch_out o1 = {sol1, cost1, st1, it1, NULL, NULL, NULL, NULL}, om = {solm, costm, stm, itm, NULL, NULL, NULL, NULL};

    pqp_multi *m = NULL;
    int rc = pqp_multi_create(&prm, nmax, B, G, NULL, &m);
    if (rc) { printf("FAIL create rc %d: %s\n", rc, pqp_multi_last_error(NULL)); return 0; }
    pqp_handle *h = NULL;
    rc = pqp_create(&prm, nmax, B, 0, &h);
    if (rc) { printf("FAIL single create rc %d: %s\n", rc, pqp_last_error(NULL)); return 0; }

    for (int pass = 0; pass < 2; ++pass) { /* 0: cold solve, 1: warm re-solve about the device-resident solutions */
        rc = pass ? pqp_resolve(h, NULL, &o1) : pqp_solve(h, &bin, &o1);
        if (rc) { printf("FAIL single pass %d rc %d: %s\n", pass, rc, pqp_last_error(h)); return 0; }
        rc = pass ? pqp_multi_resolve(m, NULL, &om) : pqp_multi_solve(m, &bin, &om);
        if (rc) { printf("FAIL multi pass %d rc %d: %s\n", pass, rc, pqp_multi_last_error(m)); return 0; }
        if (memcmp(st1, stm, (size_t)B * 4) || memcmp(it1, itm, (size_t)B * 4) || memcmp(cost1, costm, (size_t)B * 8)) {
            printf("FAIL pass %d: status / iters / cost differ from the single-device run\n", pass);
            return 0;
        }
        for (int b = 0; b < B; ++b)
            for (int fi = 0; fi < 4; ++fi)
                if (memcmp(sol1 + ((size_t)b * 4 + fi) * nmax, solm + ((size_t)b * 4 + fi) * nmax, (size_t)n[b] * 8)) {
                    printf("FAIL pass %d: sol differs at instance %d field %d\n", pass, b, fi);
                    return 0;
                }
        int covered = 0;
        for (int d = 0; d < G; ++d) {
            const pqp_result_rec *table;
            int32_t per, first, count;
            pqp_multi_gathered(m, d, &table, &per);
            pqp_multi_shard(m, d, &first, &count);
            covered += count;
            pqp_result_rec *host = xmalloc((size_t)G * per * sizeof(pqp_result_rec));
            pqp_handle *hd;
            pqp_multi_handle(m, d, &hd);
            cudaSetDevice(d);
            if (cudaMemcpy(host, table, (size_t)G * per * sizeof(pqp_result_rec), cudaMemcpyDeviceToHost) != cudaSuccess) {
                printf("FAIL table copy from device %d\n", d);
                return 0;
            }
            for (int b = 0; b < B; ++b)
                if (host[b].status != stm[b] || host[b].iters != itm[b] || memcmp(&host[b].cost, &costm[b], 8)) {
                    printf("FAIL pass %d: gathered table on device %d differs at instance %d\n", pass, d, b);
                    return 0;
                }
            free(host);
        }
        if (covered != B) { printf("FAIL shards cover %d of %d\n", covered, B); return 0; }
    }
    float gms = 0.0f;
    pqp_multi_last_gather_ms(m, &gms);
    int solved = 0;
    for (int b = 0; b < B; ++b) solved += stm[b] == PQP_SOLVED;
    printf("ok devices %d batch %d solved %d gather_ms %.3f\n", G, B, solved, gms);
    pqp_multi_destroy(m);
    pqp_destroy(h);
    return 0;
}
