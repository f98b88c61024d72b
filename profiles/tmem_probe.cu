// tmem_probe.cu — can Blackwell tensor memory serve as per-lane scratch for the ADMM state?
// Each warp of a 4-warp CTA owns the 32 TMEM lanes of its sub-partition; tcgen05.st/ld with the
// 32x32b shape moves N consecutive 32-bit columns of "its" lane per thread, which is exactly
// the [field][stage][lane] layout the solver keeps in shared memory today.
// Measures: round-trip correctness, dependent-load latency, streaming read throughput, and the
// same for LDS.128 as the yard-stick. Build: nvcc -gencode arch=compute_100a,code=sm_100a.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tmem_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
                 : "memory");
}
__device__ __forceinline__ void tmem_ld4(uint32_t taddr, uint32_t &a, uint32_t &b, uint32_t &c, uint32_t &d) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

struct Result {
    unsigned long long errors, lat_tmem_ld4, lat_tmem_ld16, lat_lds128, thr_tmem_ld16_cycles, thr_lds128_cycles, st_ld_roundtrip;
    unsigned int base;
};

// grid = nblocks, block = 128 threads (4 warps), dynamic smem for the LDS comparison
__global__ void probe(Result *out, int iters) {
    extern __shared__ __align__(16) uint32_t sm[];
    __shared__ uint32_t tbase;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t ncols = 512;
    if (warp == 0) tmem_alloc(&tbase, ncols);
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tbase + ((uint32_t)(32 * warp) << 16);
    unsigned long long errors = 0;
    // 1. fill all 512 columns of this warp's 32 lanes, read back, verify
    for (uint32_t c = 0; c < ncols; c += 4) {
        const uint32_t v = (blockIdx.x << 24) ^ (warp << 20) ^ (lane << 12) ^ c;
        tmem_st4(base + c, v, v + 1, v + 2, v + 3);
    }
    tmem_wait_st();
    for (uint32_t c = 0; c < ncols; c += 4) {
        uint32_t a, b, cc, d;
        tmem_ld4(base + c, a, b, cc, d);
        tmem_wait_ld();
        const uint32_t v = (blockIdx.x << 24) ^ (warp << 20) ^ (lane << 12) ^ c;
        errors += (a != v) + (b != v + 1) + (cc != v + 2) + (d != v + 3);
    }
    // 2. dependent-load latency: the loaded value feeds the next address (column offsets 0/4 stored)
    for (uint32_t c = 0; c < ncols; c += 4) tmem_st4(base + c, (c + 4) % ncols, 0, 0, 0);
    tmem_wait_st();
    uint32_t col = 0, x1, x2, x3;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        tmem_ld4(base + col, col, x1, x2, x3);
        tmem_wait_ld();
    }
    long long t1 = clock64();
    unsigned long long lat4 = (unsigned long long)(t1 - t0) / iters;
    uint32_t v16[16];
    col = 0;
    for (uint32_t c = 0; c < ncols; c += 16) {
        tmem_st4(base + c, (c + 16) % ncols, 0, 0, 0);
    }
    tmem_wait_st();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        tmem_ld16(base + col, v16);
        tmem_wait_ld();
        col = v16[0];
    }
    t1 = clock64();
    unsigned long long lat16 = (unsigned long long)(t1 - t0) / iters;
    // 3. same chain through shared memory (LDS.128)
    uint4 *s4 = reinterpret_cast<uint4 *>(sm) + warp * 64 * 32;
    for (int c = 0; c < 64; ++c) s4[c * 32 + lane] = make_uint4((c + 1) % 64, 0, 0, 0);
    __syncwarp();
    uint32_t idx = 0;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) idx = s4[idx * 32 + lane].x;
    t1 = clock64();
    unsigned long long lats = (unsigned long long)(t1 - t0) / iters;
    // 4. streaming throughput: independent loads, all 4 warps at once
    __syncthreads();
    uint32_t acc = idx + col;
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (uint32_t c = 0; c < 512; c += 16) {
            tmem_ld16(base + c, v16);
            tmem_wait_ld();
            acc += v16[0] + v16[15];
        }
    }
    t1 = clock64();
    unsigned long long thr_t = (unsigned long long)(t1 - t0) / iters;  // cycles per 512 columns (64 KB per warp)
    __syncthreads();
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll 8
        for (int c = 0; c < 64; ++c) {
            uint4 q = s4[c * 32 + lane];
            acc += q.x + q.w;
        }
    }
    t1 = clock64();
    unsigned long long thr_s = (unsigned long long)(t1 - t0) / iters;  // cycles per 64 x LDS.128 (32 KB per warp)
    // 5. store -> wait -> load round trip
    t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        tmem_st4(base + 8, acc, acc, acc, acc);
        tmem_wait_st();
        tmem_ld4(base + 8, acc, x1, x2, x3);
        tmem_wait_ld();
    }
    t1 = clock64();
    unsigned long long rt = (unsigned long long)(t1 - t0) / iters;
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase, ncols);
    if (lane == 0) {
        Result &r = out[blockIdx.x * 4 + warp];
        r.errors = errors + (acc == 0xdeadbeef);
        r.lat_tmem_ld4 = lat4;
        r.lat_tmem_ld16 = lat16;
        r.lat_lds128 = lats;
        r.thr_tmem_ld16_cycles = thr_t;
        r.thr_lds128_cycles = thr_s;
        r.st_ld_roundtrip = rt;
        r.base = tbase;
    }
}

int main() {
    const int nblocks = 148, iters = 200;
    Result *d;
    cudaMalloc(&d, sizeof(Result) * nblocks * 4);
    const size_t smem_bytes = 4 * 64 * 32 * sizeof(uint4);  // 32 KB per warp
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes);
    probe<<<nblocks, 128, smem_bytes>>>(d, iters);
    cudaError_t e = cudaDeviceSynchronize();
    printf("kernel: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<Result> h(nblocks * 4);
    cudaMemcpy(h.data(), d, sizeof(Result) * h.size(), cudaMemcpyDeviceToHost);
    unsigned long long err = 0;
    for (auto &r : h) err += r.errors;
    printf("round-trip errors: %llu (of %d warps x 512 columns x 32 lanes)\n", err, nblocks * 4);
    for (int w = 0; w < 4; ++w) {
        const Result &r = h[w];
        printf("warp %d: tmem base 0x%08x | dependent latency: tcgen05.ld.x4 %llu cyc, .x16 %llu cyc, LDS.128 %llu cyc | "
               "stream: tmem 64 KB/warp (x16) %llu cyc, smem 32 KB/warp (LDS.128) %llu cyc | st+ld round trip %llu cyc\n",
               w, r.base, r.lat_tmem_ld4, r.lat_tmem_ld16, r.lat_lds128, r.thr_tmem_ld16_cycles, r.thr_lds128_cycles,
               r.st_ld_roundtrip);
    }
    return 0;
}
