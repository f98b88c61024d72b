// pqp_api.cu — C ABI (include/pqp.h) of the batched path-QP solver for B200 (sm_100a).
//
// Host side: handle with device buffers + per-instance warm state, H2D/D2H staging for the
// host-pointer entry points, launches. Device side: the kernel entry (one warp = one CTA =
// one QP instance; the instance's knot block is staged into shared memory with one TMA
// bulk copy + mbarrier) around the QpWarp<C> solver of pqp_kernel.cuh, and the FP64
// Frenet->Cartesian epilogue kernel (base_solver.cpp:263-288).
//
// There is no CPU fallback in this library: every entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <time.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <algorithm>
#include <string>
#include <vector>

#include "../../include/pqp_multi.h"
#include "pqp_device_guard.h"
#include "pqp_host_common.h"

namespace {

thread_local std::string g_create_error;

// ------------------------------------------------------------------ device: TMA staging
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// 1-D bulk copy global -> shared, completion counted on the mbarrier (SASS: UBLKCP)
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

// ------------------------------------------------------------------ device: kernel entry
// One CTA = one warp = one QP instance. Dynamic shared memory: NFIELD*C*32 floats of
// solver state (the factor region doubles as the FP64 input staging buffer) + 1 mbarrier.
template <int C, typename real, bool Incr>
__global__ void __launch_bounds__(32) pqp_admm_kernel(const __grid_constant__ pqp::KernelArgs ka) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    real *smem = reinterpret_cast<real *>(smem_raw);
    const int lane = threadIdx.x;
    const int qp = blockIdx.x;
    if (qp >= ka.batch) return;
    const double *src = ka.knots + (size_t)qp * PQP_NFIELDS * ka.n_max;
    if (ka.use_tma) {
        uint64_t *bar = reinterpret_cast<uint64_t *>(smem + pqp::NFIELD * C * 32);
        double *stage = reinterpret_cast<double *>(smem + pqp::FDI * C * 32);
        const uint32_t bytes = (uint32_t)(PQP_NFIELDS * ka.n_max * sizeof(double));
        if (lane == 0) {
            mbar_init(bar, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        if (lane == 0) {
            mbar_expect_tx(bar, bytes);
            tma_load_1d(stage, src, bytes, bar);
        }
        mbar_wait(bar, 0);
        src = stage;
    }
    pqp::QpWarp<C, real, pqp::SmemStore<C, real>, Incr> w(ka, pqp::SmemStore<C, real>(smem, lane), lane, qp);
    w.run(src, ka.n_max, (size_t)(qp + ka.qp0));
}


// FP64 instantiation at C = 16 (256 <= n_max <= 511): its state (72 x 16 x 32 doubles = 288 KB per QP)
// exceeds shared memory, so the same storage policy is pointed at a per-CTA block of global memory
// (it lives in L1/L2; only the owning lane ever touches an element). This is the parity / escalation
// path, not the throughput path: a few resident CTAs loop over the instances.
template <int C, typename real, bool Incr>
__global__ void __launch_bounds__(32) pqp_admm_kernel_gmem(const __grid_constant__ pqp::KernelArgs ka, real *state) {
    const int lane = threadIdx.x;
    real *st = state + (size_t)blockIdx.x * pqp::NFIELD * C * 32;
    for (int qp = blockIdx.x; qp < ka.batch; qp += gridDim.x) {
        const double *src = ka.knots + (size_t)qp * PQP_NFIELDS * ka.n_max;
        pqp::QpWarp<C, real, pqp::SmemStore<C, real>, Incr> w(ka, pqp::SmemStore<C, real>(st, lane), lane, qp);
        w.run(src, ka.n_max, (size_t)(qp + ka.qp0));
        __syncwarp();
    }
}

// ------------------------------------------------------------------ device: tensor-memory storage policy
// The solver keeps a QP's per-stage state as [group][stage][lane] float4. Tensor memory is 128
// lanes x 512 32-bit columns per SM and a warp reaches the 32 lanes of its own sub-partition
// with tcgen05.ld/st (32x32b shapes: each thread gets N consecutive columns of "its" lane), so
// the layout maps 1:1 with column = (k * kGroups + slot(g)) * 4 + component. A CTA of four
// warps (four QPs) owns all 512 columns; at C = 8 sixteen of the eighteen groups fit and the two
// least-frequently read ones (clearance weights / proximal weights) stay in shared memory.
// Measured on B200 (profiles/r1/tmem_probe_result.txt): dependent tcgen05.ld latency <= LDS.128,
// streaming bandwidth ~6x shared memory.
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "r"(ncols));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ float tmem_ld1(uint32_t a) {
    uint32_t r;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(r) : "r"(a) : "memory");
    tmem_wait_ld();
    return __uint_as_float(r);
}
__device__ __forceinline__ void tmem_st1(uint32_t a, float v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(a), "r"(__float_as_uint(v)) : "memory");
}
__device__ __forceinline__ void tmem_ld4_nowait(uint32_t a, float (&o)[4]) {
    uint32_t r0, r1, r2, r3;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(r0), "=r"(r1), "=r"(r2), "=r"(r3) : "r"(a) : "memory");
    o[0] = __uint_as_float(r0); o[1] = __uint_as_float(r1); o[2] = __uint_as_float(r2); o[3] = __uint_as_float(r3);
}
__device__ __forceinline__ void tmem_ldx8_nowait(uint32_t a, float *o) {
    uint32_t r[8];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ldx16_nowait(uint32_t a, float *o) {
    uint32_t r[16];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 16; ++i) o[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ldx32_nowait(uint32_t a, float *o) {
    uint32_t r[32];
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31]) : "r"(a) : "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) o[i] = __uint_as_float(r[i]);
}
// G consecutive float4 groups (G = 1, 2, 4, 8) starting at column address a, in ONE tcgen05.ld
template <int G> __device__ __forceinline__ void tmem_ldg_nowait(uint32_t a, float *o) {
    if (G == 8) tmem_ldx32_nowait(a, o);
    else if (G == 4) tmem_ldx16_nowait(a, o);
    else if (G == 2) tmem_ldx8_nowait(a, o);
    else {
        float t[4];
        tmem_ld4_nowait(a, t);
        o[0] = t[0]; o[1] = t[1]; o[2] = t[2]; o[3] = t[3];
    }
}
__device__ __forceinline__ void tmem_st4(uint32_t a, const float4 &v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(__float_as_uint(v.x)),
                 "r"(__float_as_uint(v.y)), "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w))
                 : "memory");
}

// where the cyclic-reduction levels are unrolled (see cr_forward_level); overridable for A/B builds
#ifndef PQP_UNROLL_CR_COND
#define PQP_UNROLL_CR_COND true
#endif
template <int C, int COLS = 512>
struct TmemStore {
    typedef float4 Vec4;
    static constexpr int kAll = pqp::NFIELD / 4;                       // 18 groups
    static constexpr int kFit = COLS / (4 * C);                        // groups that fit in this warp's columns
    static constexpr int kGroups = kFit < kAll ? kFit : kAll;          // groups kept in TMEM
    static constexpr bool kUnrollCr = PQP_UNROLL_CR_COND;              // see cr_forward_level
    static constexpr int kSpill = kAll - kGroups;                      // 2 at C = 8 (512 columns) and at C = 4 with 256 columns, 10 at C = 16, else 0
    // Which groups leave tensor memory first when the warp's columns do not hold all 18.
    // C <= 8 (two groups leave): the clearance / proximal weights GR5, GS6 (least frequently read), then the other
    // read-only groups (4, 3, 2, 1, 0), then the factor from its end (17 .. 12), the read-write groups last.
    // C = 16 (ten groups leave): the factor first (17 .. 12), then GR5, GS6, then 4, 3, .. (measured +2.6 % at
    // n = 400). Sending the factor to shared memory at C = 8 as well - tensor memory is read at 64 B per cycle,
    // shared memory at 128 B, and the sweeps' factor reads are half of all loads - measured -1.7 %
    // (profiles/r2/README.md): four warps share the one shared-memory port.
    static constexpr bool kFactorFirst = C >= 16;
    uint32_t tb;  // TMEM address of this warp's lane block (lane base in bits 31:16)
    uint32_t cur;  // stage cursor: tb + k * (columns per stage) of the stage a hot loop is at (see seek)
    float *sm;    // spill area (shared memory) of this warp: [spill slot][k][lane] float4
    int lane;
    __device__ TmemStore(uint32_t t, float *s, int l) : tb(t), cur(t), sm(s), lane(l) {}
    __host__ __device__ static constexpr int spill_rank(int g) {
        return kFactorFirst
                   ? (g >= pqp::GF0 ? 17 - g : g == pqp::GR5 ? 6 : g == pqp::GS6 ? 7 : g <= 4 ? 8 + (4 - g) : 13 + (11 - g))
                   : (g == pqp::GR5 ? 0 : g == pqp::GS6 ? 1 : g <= 4 ? 6 - g : g >= pqp::GF0 ? 7 + (17 - g) : 13 + (11 - g));
    }
    __host__ __device__ static constexpr bool spilled(int g) { return spill_rank(g) < kSpill; }
    // number of spilled groups with an index below g, in closed form (no loops: g reaches these helpers as a
    // function argument and must fold to a constant after inlining - a loop here grew the kernel by 40 %)
    __host__ __device__ static constexpr int imax(int a, int b) { return a > b ? a : b; }
    __host__ __device__ static constexpr int imin(int a, int b) { return a < b ? a : b; }
    __host__ __device__ static constexpr int spilled_below(int g) {
        return kFactorFirst
                   ? ((kSpill > 6 && g > pqp::GR5) ? 1 : 0) + ((kSpill > 7 && g > pqp::GS6) ? 1 : 0) +
                         imax(0, imin(g, 5) - imax(0, 13 - kSpill)) +    // read-only groups 4, 3, .. leave from the top
                         imax(0, imin(g, 12) - imax(7, 25 - kSpill)) +   // read-write groups 11 .. 7 (last to leave)
                         imax(0, imin(g, 18) - imax(12, 18 - kSpill))    // factor groups 17 .. 12 (first to leave)
                   : ((kSpill > 0 && g > pqp::GR5) ? 1 : 0) + ((kSpill > 1 && g > pqp::GS6) ? 1 : 0) +
                         imax(0, imin(g, 5) - imax(0, 7 - kSpill)) +     // read-only groups 4, 3, .. leave from the top
                         imax(0, imin(g, 12) - imax(7, 25 - kSpill)) +   // read-write groups 11 .. 7 (last to leave)
                         imax(0, imin(g, 18) - imax(12, 25 - kSpill));   // factor groups 17 .. 12
    }
    __host__ __device__ static constexpr int slot(int g) { return g - spilled_below(g); }
    __host__ __device__ static constexpr int spslot(int g) { return spilled_below(g); }
    __device__ uint32_t col(int g, int k) const { return tb + (uint32_t)((k * kGroups + slot(g)) * 4); }
    __device__ float4 *sp(int g, int k) const { return reinterpret_cast<float4 *>(sm) + ((spslot(g) * C + k) * 32 + lane); }
    __device__ float ld(int f, int k) const {
        const int g = f >> 2;
        if (spilled(g)) return reinterpret_cast<const float *>(sp(g, k))[f & 3];
        return tmem_ld1(col(g, k) + (f & 3));
    }
    __device__ void st(int f, int k, float v) {
        const int g = f >> 2;
        if (spilled(g)) { reinterpret_cast<float *>(sp(g, k))[f & 3] = v; return; }
        tmem_st1(col(g, k) + (f & 3), v);
        tmem_wait_st();
    }
    // Groups g0 .. g0+N-1 of one stage: groups kept in tensor memory sit in consecutive columns (slot(g) counts
    // the unspilled groups below g), so a run of 2 / 4 / 8 of them is ONE tcgen05.ld.x8 / x16 / x32 instead of one
    // .x4 per group. run_at(g, end): length of the longest such run starting at g.
    __host__ __device__ static constexpr int run_at(int g, int end) {
        int n = 0;
        while (g + n < end && !spilled(g + n)) ++n;
        return n >= 8 ? 8 : n >= 4 ? 4 : n >= 2 ? 2 : n;
    }
    template <int G0, int N, int I> struct Loader {
        static constexpr int kRun = I < N ? (spilled(G0 + I) ? 0 : run_at(G0 + I, G0 + N)) : 0;
        static constexpr int kStep = kRun > 0 ? kRun : 1;
        // base: column address of the stage block; k: stage index for the spilled groups
        __device__ static void go(const TmemStore &st, uint32_t base, int k, float *out) {
            if (I >= N) return;
            if (kRun == 0) {
                const float4 v = *st.sp(G0 + I, k);
                out[4 * I] = v.x; out[4 * I + 1] = v.y; out[4 * I + 2] = v.z; out[4 * I + 3] = v.w;
            } else {
                tmem_ldg_nowait<(kRun > 0 ? kRun : 1)>(base + (uint32_t)(slot(G0 + I) * 4), out + 4 * I);
            }
            Loader<G0, N, (I + kStep < N ? I + kStep : N)>::go(st, base, k, out);
        }
    };
    __device__ float4 ld4(int g, int k) const {
        if (spilled(g)) return *sp(g, k);
        float o[4];
        tmem_ld4_nowait(col(g, k), o);
        tmem_wait_ld();
        return make_float4(o[0], o[1], o[2], o[3]);
    }
    __device__ void st4(int g, int k, const float4 &v) {
        if (spilled(g)) { *sp(g, k) = v; return; }
        tmem_st4(col(g, k), v);
    }
    // split form: issue the loads of N consecutive groups, wait later (wait_ld) together with other batches
    template <int N> __device__ void ld4n_nowait(int g0, int k, float (&out)[4 * N]) const {
#pragma unroll
        for (int g = 0; g < N; ++g) {
            float o[4];
            if (spilled(g0 + g)) {
                const float4 v = *sp(g0 + g, k);
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            } else {
                tmem_ld4_nowait(col(g0 + g, k), o);
            }
            out[4 * g] = o[0]; out[4 * g + 1] = o[1]; out[4 * g + 2] = o[2]; out[4 * g + 3] = o[3];
        }
    }
    __device__ void wait_ld() const { tmem_wait_ld(); }
    // compile-time group range: runs of consecutive tensor-memory groups become one wide load (see Loader)
    template <int G0, int N> __device__ void ld_run_nowait(int k, float (&out)[4 * N]) const {
        Loader<G0, N, 0>::go(*this, tb + (uint32_t)(k * kGroups * 4), k, out);
    }
    template <int G0, int N> __device__ void ld_run_nowait_cur(int k, float (&out)[4 * N]) const {
        Loader<G0, N, 0>::go(*this, cur, k, out);
    }
    template <int G0, int N> __device__ void ld_run_nowait_ahead(int k, float (&out)[4 * N]) const {
        Loader<G0, N, 0>::go(*this, ahead(), k < C - 1 ? k + 1 : k, out);
    }
    // Stage cursor. tcgen05.ld/st take their address from a UNIFORM register; an address computed from a loop
    // counter that ptxas keeps in a vector register costs one R2UR per access (5.7 % of the kernel's executed
    // instructions before this, profiles/r2/README.md). seek(k) produces the stage's column base in a uniform
    // register once per stage; the `_cur` accessors address "the cursor's stage" (the caller passes the same
    // stage as k, used by the spilled groups).
    static constexpr uint32_t kStride = (uint32_t)(kGroups * 4);
    __device__ void seek(int k) {
        // (redux.sync delivers its result in a uniform register; a loop-carried cursor would be moved to a vector
        // register by ptxas and converted back at every access)
        cur = C == 1 ? tb : __reduce_or_sync(0xffffffffu, tb + (uint32_t)k * kStride);
    }
    __device__ uint32_t ccol(int g) const { return cur + (uint32_t)(slot(g) * 4); }
    __device__ void st4_cur(int g, int k, const float4 &v) {
        if (spilled(g)) { *sp(g, k) = v; return; }
        tmem_st4(ccol(g), v);
    }
    template <int N> __device__ void ld4n_nowait_cur(int g0, int k, float (&out)[4 * N]) const {
#pragma unroll
        for (int g = 0; g < N; ++g) {
            float o[4];
            if (spilled(g0 + g)) {
                const float4 v = *sp(g0 + g, k);
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            } else {
                tmem_ld4_nowait(ccol(g0 + g), o);
            }
            out[4 * g] = o[0]; out[4 * g + 1] = o[1]; out[4 * g + 2] = o[2]; out[4 * g + 3] = o[3];
        }
    }
    // the same for the stage AFTER the cursor's (prefetch; the cursor's own stage again when that is the last one)
    __device__ uint32_t ahead() const {
        const uint32_t last = tb + (uint32_t)((C - 1) * kGroups * 4);
        return (cur == last) ? cur : cur + kStride;
    }
    template <int N> __device__ void ld4n_nowait_ahead(int g0, int k, float (&out)[4 * N]) const {
        const uint32_t a = ahead();
        const int ka = k < C - 1 ? k + 1 : k;
#pragma unroll
        for (int g = 0; g < N; ++g) {
            float o[4];
            if (spilled(g0 + g)) {
                const float4 v = *sp(g0 + g, ka);
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            } else {
                tmem_ld4_nowait(a + (uint32_t)(slot(g0 + g) * 4), o);
            }
            out[4 * g] = o[0]; out[4 * g + 1] = o[1]; out[4 * g + 2] = o[2]; out[4 * g + 3] = o[3];
        }
    }
    // group g two stages after the cursor's (clamped to the last stage)
    __device__ void ld4_nowait_ahead_next(int g, int k, float (&out)[4]) const {
        if (spilled(g)) {
            const float4 v = *sp(g, k < C - 2 ? k + 2 : C - 1);
            out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        } else {
            const uint32_t last = tb + (uint32_t)((C - 1) * kGroups * 4);
            const uint32_t a1 = ahead();
            const uint32_t a2 = (a1 == last) ? a1 : a1 + kStride;
            tmem_ld4_nowait(a2 + (uint32_t)(slot(g) * 4), out);
        }
    }
    // group g of the stage AFTER the cursor's (of the cursor's own stage when that is the last one)
    __device__ void ld4_nowait_next(int g, int k, float (&out)[4]) const {
        if (spilled(g)) {
            const float4 v = *sp(g, k < C - 1 ? k + 1 : k);
            out[0] = v.x; out[1] = v.y; out[2] = v.z; out[3] = v.w;
        } else {
            const uint32_t last = tb + (uint32_t)((C - 1) * kGroups * 4);
            const uint32_t a = (cur == last) ? cur : cur + kStride;
            tmem_ld4_nowait(a + (uint32_t)(slot(g) * 4), out);
        }
    }
    template <int N> __device__ void ld4n(int g0, int k, float (&out)[4 * N]) const {
#pragma unroll
        for (int g = 0; g < N; ++g) {
            float o[4];
            if (spilled(g0 + g)) {
                const float4 v = *sp(g0 + g, k);
                o[0] = v.x; o[1] = v.y; o[2] = v.z; o[3] = v.w;
            } else {
                tmem_ld4_nowait(col(g0 + g, k), o);
            }
            out[4 * g] = o[0]; out[4 * g + 1] = o[1]; out[4 * g + 2] = o[2]; out[4 * g + 3] = o[3];
        }
        tmem_wait_ld();
    }
    __device__ void fence() { tmem_wait_st(); }
};


// compile-time check of the closed-form spill map against its definition (rank order) for every store in use
template <class Store> constexpr bool spill_map_consistent() {
    int seen = 0;
    for (int g = 0; g <= Store::kAll; ++g) {
        if (Store::spilled_below(g) != seen) return false;
        if (g < Store::kAll && Store::spilled(g)) ++seen;
    }
    if (seen != Store::kSpill) return false;
    for (int a = 0; a < Store::kAll; ++a)  // the ranks are a permutation of 0 .. 17
        for (int b = a + 1; b < Store::kAll; ++b)
            if (Store::spill_rank(a) == Store::spill_rank(b)) return false;
    return true;
}
static_assert(spill_map_consistent<TmemStore<1, 512>>() && spill_map_consistent<TmemStore<2, 512>>() &&
                  spill_map_consistent<TmemStore<4, 256>>() && spill_map_consistent<TmemStore<8, 512>>() &&
                  spill_map_consistent<TmemStore<16, 512>>(),
              "TmemStore::spilled_below does not match spill_rank");

// Tensor-memory persistent kernel (FP32): one CTA per SM with WT warps (4, or 8 when two warps
// can share a TMEM sub-partition: 256 columns each); every warp solves one QP at a time and takes
// the next unsolved instance from a global work counter (iteration counts vary 3x between
// instances; a static assignment would idle).
// A hybrid variant that added shared-memory-backed warps to the same CTA was measured slower
// (instruction-fetch bound either as two instantiations or as one with run-time backend
// selection): profiles/r1/README.md.
template <int C, int WT, bool Incr>
__global__ void __launch_bounds__(32 * WT, 1) pqp_admm_kernel_tmem(const __grid_constant__ pqp::KernelArgs ka) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ uint32_t tbase_s;
    constexpr int kShare = WT > 4 ? (WT + 3) / 4 : 1;  // warps per tensor-memory sub-partition
    constexpr int kCols = 512 / kShare;
    typedef TmemStore<C, kCols> Store;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    constexpr uint32_t need = (uint32_t)(Store::kGroups * 4 * C) * kShare;
    constexpr uint32_t ncols = need <= 32 ? 32 : need <= 64 ? 64 : need <= 128 ? 128 : need <= 256 ? 256 : 512;
    constexpr size_t kSpillFloats = (size_t)(Store::kSpill > 0 ? Store::kSpill : 0) * C * 32 * 4;
    if (warp == 0) tmem_alloc(&tbase_s, ncols);
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    // lanes of sub-partition (warp & 3); the second warp of a sub-partition takes the upper columns
    const uint32_t tb_v = tbase_s + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)((warp >> 2) * Store::kGroups * 4 * C);
    const uint32_t tb = __reduce_or_sync(0xffffffffu, tb_v);
    float *spill = reinterpret_cast<float *>(smem_raw) + (size_t)warp * kSpillFloats;
    for (;;) {
        int qp = 0;
        if (lane == 0) qp = atomicAdd(ka.work_counter, 1);
        qp = __shfl_sync(0xffffffffu, qp, 0);
        if (qp >= ka.batch) break;
        int chunk = 0;
        if (ka.ready) {
            // streamed launch: this instance's inputs may still be in flight
            while (chunk + 1 < ka.n_chunks && qp >= ka.chunk_lo[chunk + 1]) ++chunk;
            bool landed = true;
            if (lane == 0) {
                const volatile int *flag = ka.ready + chunk;
                const long long t0 = clock64();
                while (*flag == 0) {
                    __nanosleep(200);
                    if (clock64() - t0 > 20000000000LL) {  // ~10 s: give up, never hang
                        landed = false;
                        atomicExch(ka.work_counter, 1 << 30);  // and stop everybody else from taking work
                        break;
                    }
                }
                __threadfence();
            }
            landed = __shfl_sync(0xffffffffu, landed ? 1 : 0, 0) != 0;
            if (!landed) {
                if (lane == 0 && ka.status) ka.status[qp] = pqp::kUnsolved;
                continue;
            }
        }
        const double *src = ka.knots + (size_t)qp * PQP_NFIELDS * ka.n_max;
        pqp::QpWarp<C, float, Store, Incr> w(ka, Store(tb, spill, lane), lane, qp);
        w.run(src, ka.n_max, (size_t)(blockIdx.x * WT + warp));
        if (ka.done) {
            __syncwarp();
            if (lane == 0) {
                __threadfence();  // this instance's outputs before the count
                const int len = ka.chunk_lo[chunk + 1] - ka.chunk_lo[chunk];
                if (atomicAdd(ka.done + chunk, 1) + 1 == len) {
                    __threadfence_system();
                    *reinterpret_cast<volatile int *>(ka.host_done + chunk) = 1;
                }
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 0) tmem_dealloc(tbase_s, ncols);
}

// BaseSolver::getOptimizedPath (base_solver.cpp:263-288) in FP64, one thread per knot.
__device__ __forceinline__ double constrain_angle(double a) {
    while (a > M_PI) a -= 2 * M_PI;
    while (a < -M_PI) a += 2 * M_PI;
    return a;
}
__global__ void frenet_to_cartesian_kernel(int batch, int n_max, const int *__restrict__ n,
                                           const double *__restrict__ ref, const double *__restrict__ sol,
                                           double *__restrict__ out) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * n_max) return;
    const int b = idx / n_max, i = idx - b * n_max;
    if (i >= n[b]) return;
    const size_t r = (size_t)b * 3 * n_max, s = (size_t)b * 4 * n_max;
    const double angle = ref[r + 2 * n_max + i];
    const double l = sol[s + i], psi = sol[s + n_max + i];
    const double new_angle = constrain_angle(angle + M_PI_2);
    out[r + i] = ref[r + i] + l * cos(new_angle);
    out[r + n_max + i] = ref[r + n_max + i] + l * sin(new_angle);
    out[r + 2 * n_max + i] = constrain_angle(angle + psi);
}

// previous solution -> linearisation fields of the device knot buffer
__global__ void relinearise_kernel(int batch, int n_max, const double *__restrict__ sol,
                                   double *__restrict__ knots) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * n_max) return;
    const int b = idx / n_max, i = idx - b * n_max;
    const size_t s = (size_t)b * 4 * n_max, k = (size_t)b * PQP_NFIELDS * n_max;
    knots[k + (size_t)PQP_F_L * n_max + i] = sol[s + i];
    knots[k + (size_t)PQP_F_PSI * n_max + i] = sol[s + n_max + i];
    knots[k + (size_t)PQP_F_K * n_max + i] = sol[s + 2 * (size_t)n_max + i];
}

// Receding-horizon tick (BASELINE configs[4]): the planning window advances by `tick` knots along
// an extended reference ext[b][9][ext_len]; the new linearisation point is the previous solution
// shifted by one knot (last knot repeated) and the new x0 is the previous solution at knot 1.
__global__ void advance_window_kernel(int batch, int n_max, int ext_len, int tick, const double *__restrict__ ext,
                                      const double *__restrict__ sol, double *__restrict__ knots,
                                      double *__restrict__ inst) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= batch * n_max) return;
    const int b = idx / n_max, i = idx - b * n_max;
    const double *e = ext + (size_t)b * PQP_NFIELDS * ext_len + tick + i;
    const double *sb = sol + (size_t)b * 4 * n_max;
    double *k = knots + (size_t)b * PQP_NFIELDS * n_max + i;
    const int src = i + 1 < n_max ? i + 1 : n_max - 1;
#pragma unroll
    for (int f = 0; f < PQP_NFIELDS; ++f) {
        double v = e[(size_t)f * ext_len];
        if (f == PQP_F_L) v = sb[src];
        if (f == PQP_F_PSI) v = sb[n_max + src];
        if (f == PQP_F_K) v = sb[2 * (size_t)n_max + src];
        k[(size_t)f * n_max] = v;
    }
    if (i == 0) {
        double *ib = inst + (size_t)b * PQP_NINST;
        const int one = n_max > 1 ? 1 : 0;
        ib[PQP_I_L0] = sb[one];
        ib[PQP_I_PSI0] = sb[n_max + one];
        ib[PQP_I_K0] = sb[2 * (size_t)n_max + one];
    }
}

// After an FP64 re-solve: the instance's warm slot (FP32) takes over the FP64 run's scaled iterates and rho, so
// that a later pqp_resolve warm-starts from the run whose result was returned, not from the abandoned FP32 one.
__global__ void adopt_warm_kernel(int count, size_t warm_len, const int *__restrict__ slots, const double *__restrict__ e_warm,
                                  const double *__restrict__ e_rho, float *__restrict__ d_warm, float *__restrict__ d_rho) {
    const int j = blockIdx.y;
    if (j >= count) return;
    const size_t b = (size_t)slots[j];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < warm_len; i += (size_t)gridDim.x * blockDim.x)
        d_warm[b * warm_len + i] = (float)e_warm[(size_t)j * warm_len + i];
    if (blockIdx.x == 0 && threadIdx.x == 0) d_rho[b] = (float)e_rho[j];
}

// {cost, status, iters} -> one 16-byte record per instance (the payload of the multi-GPU all-gather)
__global__ void pack_results_kernel(int batch, const double *__restrict__ cost, const int *__restrict__ status,
                                    const int *__restrict__ iters, pqp_result_rec *__restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= batch) return;
    pqp_result_rec r;
    r.cost = cost[b];
    r.status = status[b];
    r.iters = iters[b];
    out[b] = r;
}

}  // namespace

// ------------------------------------------------------------------ handle
struct pqp_handle {
    pqp_params prm;
    int n_max = 0, batch_max = 0, device = 0, chunk = 0;
    int sm_count = 0, warps_per_sm = 0;
    size_t smem_bytes = 0;
    static const int kStreams = 8;
    cudaStream_t stream = nullptr;            // streams[0]
    cudaStream_t streams[kStreams] = {};      // one per pipeline chunk
    cudaEvent_t evs[kStreams] = {};
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // device buffers owned by the handle
    double *d_knots = nullptr, *d_inst = nullptr, *d_sol = nullptr, *d_cost = nullptr, *d_info = nullptr;
    double *d_xf = nullptr, *d_yf = nullptr, *d_zf = nullptr, *d_ref = nullptr, *d_xy = nullptr;
    int *d_n = nullptr, *d_p = nullptr, *d_status = nullptr, *d_iters = nullptr;
    void *d_warm = nullptr, *d_scal = nullptr, *d_dy = nullptr, *d_rho = nullptr;
    double *d_state64 = nullptr;   // per-CTA state of the FP64 kernel at C = 16 (global memory)
    int *d_ready = nullptr, *d_done = nullptr;      // per-chunk flags of the streamed launch
    int *h_done = nullptr, *d_hdone = nullptr;     // mapped pinned host memory + its device alias
    static const int kCounters = 64;
    int *d_counters = nullptr;
    unsigned counter_next = 0;
    bool use_tmem = false;         // params.reserved bit 3: state in tensor memory (FP32 kernel)
    bool incr = false;             // ADMM step in increment form (FP32 kernels)
    bool fp64 = false;             // params.reserved bit 1: iterate in FP64
    bool escalate = true;          // params.reserved bit 2 clears it
    bool cold_only = false;        // params.reserved bit 128: no warm state is kept (pqp_resolve* is refused)
    float stage_ms[PQP_NSTAGES] = {0, 0, 0, 0, 0};
    cudaEvent_t ev_st[4] = {};     // stage boundaries of the single-stream host call
    size_t smem_bytes64 = 0;
    bool prepared64 = false;
    int *d_flags = nullptr;
    int *h_status = nullptr, *h_flags = nullptr;  // pinned host mirrors
    // FP64 escalation of suspected-infeasible instances (small side batch)
    int esc_cap = 0;
    double *e_knots = nullptr, *e_inst = nullptr, *e_sol = nullptr, *e_cost = nullptr, *e_info = nullptr;
    double *e_xf = nullptr, *e_yf = nullptr, *e_zf = nullptr;
    int *e_n = nullptr, *e_p = nullptr, *e_status = nullptr, *e_iters = nullptr, *e_slots = nullptr;
    void *e_warm = nullptr, *e_scal = nullptr, *e_dy = nullptr, *e_rho = nullptr;
    long long escalated = 0;
    double *d_sol2 = nullptr;
    int *d_n2 = nullptr;
    bool solved = false, host_inputs_resident = false, d_p_valid = false;
    int last_batch = 0;
    float last_ms = 0.0f;
    long long launches = 0;
    std::string err;
};

namespace {

using pqp::DeviceGuard;

int fail(pqp_handle *h, int code, const std::string &msg) {
    if (h) h->err = msg;
    else g_create_error = msg;
    return code;
}
int cuda_fail(pqp_handle *h, cudaError_t e, const char *what) {
    return fail(h, PQP_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define PQP_CUDA(h, call)                                         \
    do {                                                          \
        cudaError_t e_ = (call);                                  \
        if (e_ != cudaSuccess) return cuda_fail((h), e_, #call);  \
    } while (0)

// the increment form is only instantiated for FP32 (the FP64 kernel is exact in either form)
template <int C, typename real, bool Incr>
cudaError_t launch(const pqp::KernelArgs &ka, size_t smem, cudaStream_t s) {
    pqp_admm_kernel<C, real, Incr><<<ka.batch, 32, smem, s>>>(ka);
    return cudaGetLastError();
}
template <int C, typename real>
cudaError_t launch_form(const pqp::KernelArgs &ka, size_t smem, cudaStream_t s, bool incr) {
    if (incr && sizeof(real) == 4) return launch<C, float, true>(ka, smem, s);
    return launch<C, real, false>(ka, smem, s);
}
// C = 16 exists for FP32 only under the shared-memory policy (144 KB per QP); FP64 uses the
// global-memory kernel above
template <typename real> struct Has16 { static constexpr bool value = sizeof(real) == 4; };
template <typename real>
cudaError_t launch_chunk(int chunk, const pqp::KernelArgs &ka, size_t smem, cudaStream_t s, bool incr = false) {
    switch (chunk) {
        case 1: return launch_form<1, real>(ka, smem, s, incr);
        case 2: return launch_form<2, real>(ka, smem, s, incr);
        case 4: return launch_form<4, real>(ka, smem, s, incr);
        case 8: return launch_form<8, real>(ka, smem, s, incr);
        default:
            if constexpr (Has16<real>::value) return launch_form<16, real>(ka, smem, s, incr);
            else return cudaErrorInvalidConfiguration;
    }
}
template <int C, typename real, bool Incr>
cudaError_t prepare(size_t smem, int *blocks_per_sm) {
    cudaError_t e = cudaFuncSetAttribute(pqp_admm_kernel<C, real, Incr>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, pqp_admm_kernel<C, real, Incr>, 32, smem);
}
template <int C, typename real>
cudaError_t prepare_form(size_t smem, int *bps, bool incr) {
    if (incr && sizeof(real) == 4) return prepare<C, float, true>(smem, bps);
    return prepare<C, real, false>(smem, bps);
}
template <typename real>
cudaError_t prepare_chunk(int chunk, size_t smem, int *bps, bool incr = false) {
    switch (chunk) {
        case 1: return prepare_form<1, real>(smem, bps, incr);
        case 2: return prepare_form<2, real>(smem, bps, incr);
        case 4: return prepare_form<4, real>(smem, bps, incr);
        case 8: return prepare_form<8, real>(smem, bps, incr);
        default:
            if constexpr (Has16<real>::value) return prepare_form<16, real>(smem, bps, incr);
            else { *bps = 2; return cudaSuccess; }  // global-memory kernel: nothing to prepare
    }
}

// tensor-memory persistent variant: one CTA per SM (the dynamic shared-memory request is padded
// so that a second CTA cannot be resident and spin on tcgen05.alloc)
// Warps (= QPs) per CTA: 8 at C = 4 (two warps share a sub-partition, 256 columns each), 4 at
// C <= 8, 2 at C = 16 (n <= 511: 8 groups in the warp's 512 columns + 10 groups = 80 KB of shared
// memory per warp; a shared-memory-only layout would hold one QP per SM).
template <int C> struct TmemCfg {
    static constexpr int WT = (C == 4) ? 8 : (C == 16 ? 2 : 4);
    typedef TmemStore<C, (WT > 4 ? 512 / (WT / 4) : 512)> Store;
    static constexpr size_t kSpillBytes = (size_t)WT * Store::kSpill * C * 32 * 4 * sizeof(float);
    // at least 120 KB so that a second CTA cannot become resident
    static constexpr size_t kSmem = kSpillBytes > 120 * 1024 ? kSpillBytes : 120 * 1024;
};
template <int C, bool Incr>
cudaError_t launch_tmem(const pqp::KernelArgs &ka, cudaStream_t s, int sm_count) {
    constexpr int W = TmemCfg<C>::WT;
    int ctas = (ka.batch + W - 1) / W;
    if (ctas > sm_count) ctas = sm_count;
    pqp_admm_kernel_tmem<C, W, Incr><<<ctas, 32 * W, TmemCfg<C>::kSmem, s>>>(ka);
    return cudaGetLastError();
}
template <int C>
cudaError_t launch_tmem_form(const pqp::KernelArgs &ka, cudaStream_t s, int sm_count, bool incr) {
    return incr ? launch_tmem<C, true>(ka, s, sm_count) : launch_tmem<C, false>(ka, s, sm_count);
}
cudaError_t launch_tmem_chunk(int chunk, const pqp::KernelArgs &ka, cudaStream_t s, int sm_count, bool incr) {
    switch (chunk) {
        case 1: return launch_tmem_form<1>(ka, s, sm_count, incr);
        case 2: return launch_tmem_form<2>(ka, s, sm_count, incr);
        case 4: return launch_tmem_form<4>(ka, s, sm_count, incr);
        case 8: return launch_tmem_form<8>(ka, s, sm_count, incr);
        default: return launch_tmem_form<16>(ka, s, sm_count, incr);
    }
}
template <int C>
cudaError_t prepare_tmem() {
    cudaError_t e = cudaFuncSetAttribute(pqp_admm_kernel_tmem<C, TmemCfg<C>::WT, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmemCfg<C>::kSmem);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(pqp_admm_kernel_tmem<C, TmemCfg<C>::WT, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TmemCfg<C>::kSmem);
}
cudaError_t prepare_tmem_chunk(int chunk) {
    switch (chunk) {
        case 1: return prepare_tmem<1>();
        case 2: return prepare_tmem<2>();
        case 4: return prepare_tmem<4>();
        case 8: return prepare_tmem<8>();
        default: return prepare_tmem<16>();
    }
}
int tmem_warps(int chunk) { return chunk == 4 ? 8 : (chunk == 16 ? 2 : 4); }
size_t tmem_spill_bytes(int chunk) {  // per warp
    switch (chunk) {
        case 1: return TmemCfg<1>::kSpillBytes / TmemCfg<1>::WT;
        case 2: return TmemCfg<2>::kSpillBytes / TmemCfg<2>::WT;
        case 4: return TmemCfg<4>::kSpillBytes / TmemCfg<4>::WT;
        case 8: return TmemCfg<8>::kSpillBytes / TmemCfg<8>::WT;
        default: return TmemCfg<16>::kSpillBytes / TmemCfg<16>::WT;
    }
}

int validate_batch(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out) {
    if (!h) return PQP_E_INVALID;
    if (!in || !out) return fail(h, PQP_E_INVALID, "null batch");
    if (!in->knots || !in->inst || !in->n || !out->sol) return fail(h, PQP_E_INVALID, "null buffer");
    if (in->batch < 1 || in->batch > h->batch_max) return fail(h, PQP_E_INVALID, "batch out of range");
    if (in->n_max != h->n_max) return fail(h, PQP_E_INVALID, "n_max differs from the handle's");
    return PQP_OK;
}

template <typename T>
cudaError_t dmalloc(T **p, size_t count) {
    return cudaMalloc(reinterpret_cast<void **>(p), count * sizeof(T));
}

// device pointers in `in`/`out`; asynchronous
struct StreamedLaunch {
    const int *ready = nullptr;
    int *done = nullptr, *host_done = nullptr;
    int n_chunks = 0;
    int chunk_lo[9] = {};
};

int run_device(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out, int mode,
               cudaStream_t s, bool timed, int qp0 = 0, bool esc = false, int *flags = nullptr,
               const StreamedLaunch *sl = nullptr) {
    pqp::KernelArgs ka;
    ka.prm = pqp::make_dev_params(h->prm);
    ka.batch = in->batch;
    ka.n_max = in->n_max;
    ka.mode = mode | ((h->cold_only && !esc) ? 2 : 0);
    ka.qp0 = qp0;
    const size_t block_bytes = (size_t)PQP_NFIELDS * in->n_max * sizeof(double);
    ka.use_tma = ((block_bytes % 16) == 0 && (reinterpret_cast<uintptr_t>(in->knots) % 16) == 0) ? 1 : 0;
    ka.knots = in->knots;
    ka.inst = in->inst;
    ka.n = in->n;
    ka.p = in->p;
    ka.sol = out->sol;
    ka.cost = out->cost;
    ka.status = out->status;
    ka.iters = out->iters;
    ka.x_full = out->x_full;
    ka.y_full = out->y_full;
    ka.z_full = out->z_full;
    ka.info = out->info;
    ka.flags = flags;
    ka.work_counter = nullptr;
    ka.ready = sl ? sl->ready : nullptr;
    ka.done = sl ? sl->done : nullptr;
    ka.host_done = sl ? sl->host_done : nullptr;
    ka.n_chunks = sl ? sl->n_chunks : 0;
    for (int c = 0; c < 9; ++c) ka.chunk_lo[c] = sl ? sl->chunk_lo[c] : 0;
    ka.warm = esc ? h->e_warm : h->d_warm;
    ka.scal = esc ? h->e_scal : h->d_scal;
    ka.dy = esc ? h->e_dy : h->d_dy;
    ka.rho_state = esc ? h->e_rho : h->d_rho;
    if (timed) PQP_CUDA(h, cudaEventRecord(h->ev0, s));
    if (esc || h->fp64) {
        if (!h->prepared64) {
            int bps = 0;
            PQP_CUDA(h, prepare_chunk<double>(h->chunk, h->smem_bytes64, &bps));
            h->prepared64 = true;
        }
        if (h->chunk == 16) {
            const int ctas = std::min(ka.batch, 2 * h->sm_count);
            if (!h->d_state64) PQP_CUDA(h, dmalloc(&h->d_state64, (size_t)2 * h->sm_count * pqp::smem_floats(16)));
            pqp_admm_kernel_gmem<16, double, false><<<ctas, 32, 0, s>>>(ka, h->d_state64);
            PQP_CUDA(h, cudaGetLastError());
        } else {
            PQP_CUDA(h, launch_chunk<double>(h->chunk, ka, h->smem_bytes64, s));
        }
    } else if (h->use_tmem) {
        // one work counter per concurrently running launch (pipeline chunk); counters rotate
        int *ctr = h->d_counters + (h->counter_next++ % pqp_handle::kCounters);
        PQP_CUDA(h, cudaMemsetAsync(ctr, 0, sizeof(int), s));
        ka.work_counter = ctr;
        PQP_CUDA(h, launch_tmem_chunk(h->chunk, ka, s, h->sm_count, h->incr));
    } else {
        PQP_CUDA(h, launch_chunk<float>(h->chunk, ka, h->smem_bytes, s, h->incr));
    }
    if (timed) PQP_CUDA(h, cudaEventRecord(h->ev1, s));
    h->launches++;
    if (!esc) {
        h->solved = true;
        if (qp0 == 0) h->last_batch = in->batch;
    }
    return PQP_OK;
}

// Instances the FP32 kernel could neither solve nor certify (iteration cap reached while the
// first two conditions of OSQP's primal-infeasibility certificate held) are re-solved cold by
// the FP64 instantiation of the same kernel: FP32 iterates resolve |A'dy| < 1e-4 |dy| only to
// ~5e-4, FP64 ones reproduce the reference's PRIMAL_INFEASIBLE status. Rare path,
// plain synchronous copies.
int escalate_fp64(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out, int B) {
    std::vector<int> idx;
    for (int b = 0; b < B; ++b)
        // also an "inaccurate" certificate at the iteration cap: in FP32 the 10x-tolerance test can
        // pass on rounding noise alone, the FP64 run decides (normally: PRIMAL_INFEASIBLE proper)
        if ((h->h_status[b] == PQP_MAX_ITER_REACHED && (h->h_flags[b] & 1)) ||
            h->h_status[b] == PQP_PRIMAL_INFEASIBLE_INACCURATE)
            idx.push_back(b);
    if (idx.empty()) return PQP_OK;
    const int nmax = h->n_max;
    const size_t c = h->chunk;
    const size_t nvm = 6 * (size_t)nmax - 1, mm = 6 * (size_t)nmax + 2;
    if (!h->e_knots) {
        h->esc_cap = h->batch_max < 256 ? h->batch_max : 256;
        const size_t E = h->esc_cap;
        PQP_CUDA(h, dmalloc(&h->e_knots, E * PQP_NFIELDS * nmax));
        PQP_CUDA(h, dmalloc(&h->e_inst, E * PQP_NINST));
        PQP_CUDA(h, dmalloc(&h->e_n, E));
        PQP_CUDA(h, dmalloc(&h->e_p, E));
        PQP_CUDA(h, dmalloc(&h->e_sol, E * 4 * nmax));
        PQP_CUDA(h, dmalloc(&h->e_cost, E));
        PQP_CUDA(h, dmalloc(&h->e_status, E));
        PQP_CUDA(h, dmalloc(&h->e_iters, E));
        PQP_CUDA(h, dmalloc(&h->e_info, E * PQP_NINFO));
        PQP_CUDA(h, dmalloc(&h->e_xf, E * nvm));
        PQP_CUDA(h, dmalloc(&h->e_yf, E * mm));
        PQP_CUDA(h, dmalloc(&h->e_zf, E * mm));
        PQP_CUDA(h, cudaMalloc(&h->e_warm, E * pqp::warm_floats(c) * sizeof(double)));
        PQP_CUDA(h, cudaMalloc(&h->e_scal, E * pqp::scal_floats(c) * sizeof(double)));
        PQP_CUDA(h, cudaMalloc(&h->e_dy, E * pqp::dy_floats(c) * sizeof(double)));
        PQP_CUDA(h, cudaMalloc(&h->e_rho, E * sizeof(double)));
        PQP_CUDA(h, dmalloc(&h->e_slots, E));
    }
    cudaStream_t s = h->stream;
    const size_t kb = (size_t)PQP_NFIELDS * nmax;
    for (size_t start = 0; start < idx.size(); start += h->esc_cap) {
        const int E = (int)std::min(idx.size() - start, (size_t)h->esc_cap);
        for (int j = 0; j < E; ++j) {
            const int b = idx[start + j];
            // inputs come from the handle's device copy (valid for both host-batch and NULL-batch calls)
            PQP_CUDA(h, cudaMemcpyAsync(h->e_knots + j * kb, h->d_knots + b * kb, kb * sizeof(double), cudaMemcpyDeviceToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->e_inst + (size_t)j * PQP_NINST, h->d_inst + (size_t)b * PQP_NINST, PQP_NINST * sizeof(double), cudaMemcpyDeviceToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->e_n + j, h->d_n + b, sizeof(int), cudaMemcpyDeviceToDevice, s));
            if (h->d_p_valid) PQP_CUDA(h, cudaMemcpyAsync(h->e_p + j, h->d_p + b, sizeof(int), cudaMemcpyDeviceToDevice, s));
        }
        PQP_CUDA(h, cudaMemsetAsync(h->e_xf, 0, (size_t)E * nvm * sizeof(double), s));
        PQP_CUDA(h, cudaMemsetAsync(h->e_yf, 0, (size_t)E * mm * sizeof(double), s));
        PQP_CUDA(h, cudaMemsetAsync(h->e_zf, 0, (size_t)E * mm * sizeof(double), s));
        pqp_batch_in din = {E, nmax, h->e_knots, h->e_inst, h->e_n, h->d_p_valid ? h->e_p : nullptr};
        pqp_batch_out dout = {h->e_sol, h->e_cost, h->e_status, h->e_iters, h->e_xf, h->e_yf, h->e_zf, h->e_info};
        int rc = run_device(h, &din, &dout, 0, s, false, 0, true, nullptr);
        if (rc) return rc;
        for (int j = 0; j < E; ++j) {
            const int b = idx[start + j];
            PQP_CUDA(h, cudaMemcpyAsync(out->sol + (size_t)b * 4 * nmax, h->e_sol + (size_t)j * 4 * nmax, 4 * (size_t)nmax * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->cost) PQP_CUDA(h, cudaMemcpyAsync(out->cost + b, h->e_cost + j, sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->status) PQP_CUDA(h, cudaMemcpyAsync(out->status + b, h->e_status + j, sizeof(int), cudaMemcpyDeviceToHost, s));
            if (out->iters) PQP_CUDA(h, cudaMemcpyAsync(out->iters + b, h->e_iters + j, sizeof(int), cudaMemcpyDeviceToHost, s));
            if (out->info) PQP_CUDA(h, cudaMemcpyAsync(out->info + (size_t)b * PQP_NINFO, h->e_info + (size_t)j * PQP_NINFO, PQP_NINFO * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->x_full) PQP_CUDA(h, cudaMemcpyAsync(out->x_full + b * nvm, h->e_xf + j * nvm, nvm * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->y_full) PQP_CUDA(h, cudaMemcpyAsync(out->y_full + b * mm, h->e_yf + j * mm, mm * sizeof(double), cudaMemcpyDeviceToHost, s));
            if (out->z_full) PQP_CUDA(h, cudaMemcpyAsync(out->z_full + b * mm, h->e_zf + j * mm, mm * sizeof(double), cudaMemcpyDeviceToHost, s));
            // keep the device-resident results coherent too (resolve(NULL) linearises about d_sol)
            PQP_CUDA(h, cudaMemcpyAsync(h->d_sol + (size_t)b * 4 * nmax, h->e_sol + (size_t)j * 4 * nmax, 4 * (size_t)nmax * sizeof(double), cudaMemcpyDeviceToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->d_status + b, h->e_status + j, sizeof(int), cudaMemcpyDeviceToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->d_cost + b, h->e_cost + j, sizeof(double), cudaMemcpyDeviceToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->d_iters + b, h->e_iters + j, sizeof(int), cudaMemcpyDeviceToDevice, s));
        }
        // the warm slots of these instances adopt the FP64 run's scaled iterates and rho
        if (!h->cold_only) {
            PQP_CUDA(h, cudaMemcpyAsync(h->e_slots, idx.data() + start, (size_t)E * sizeof(int), cudaMemcpyHostToDevice, s));
            const size_t wf = pqp::warm_floats(c);
            adopt_warm_kernel<<<dim3((unsigned)((wf + 255) / 256), (unsigned)E), 256, 0, s>>>(
                E, wf, h->e_slots, static_cast<const double *>(h->e_warm), static_cast<const double *>(h->e_rho),
                static_cast<float *>(h->d_warm), static_cast<float *>(h->d_rho));
            PQP_CUDA(h, cudaGetLastError());
        }
        PQP_CUDA(h, cudaStreamSynchronize(s));
        h->escalated += E;
    }
    return PQP_OK;
}



// Host-buffer solve on the persistent tensor-memory kernel as ONE launch over the whole batch:
// the kernel starts immediately and each warp waits (ready[chunk]) until the H2D copy of the
// chunk its next instance belongs to has landed; finished chunks are announced through mapped
// host memory (host_done[chunk]) so their D2H copies overlap the rest of the solve. Compared
// with one launch per chunk this removes the per-launch tail of a persistent kernel.
int run_host_streamed(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out, int mode, int B) {
    const int nmax = h->n_max;
    const size_t nvm = 6 * (size_t)nmax - 1, mm = 6 * (size_t)nmax + 2;
    // up to kStreams chunks of >= 256 instances: small batches (configs[1]: 1024) still overlap
    // their copies with the solve, large ones (8192) move 1024 instances per chunk
    const int min_chunk = 256;  // profiles/chunk_sweep.sh
    int nchunks = (B + min_chunk - 1) / min_chunk;
    if (nchunks > pqp_handle::kStreams) nchunks = pqp_handle::kStreams;
    const int per = (B + nchunks - 1) / nchunks;
    nchunks = (B + per - 1) / per;
    // chunk boundaries. Large batches use a ramp instead of equal chunks: nothing can start before the FIRST chunk
    // has landed and nothing is returned before the LAST chunk has finished, so those two are small (1/32 and 2/32 of
    // the batch) and the middle ones carry the volume.
    int lo_of[pqp_handle::kStreams + 1];
    if (nchunks == pqp_handle::kStreams && B >= 32 * min_chunk) {
        static const int w[pqp_handle::kStreams] = {1, 3, 6, 6, 6, 5, 3, 2};  // / 32
        int acc = 0;
        for (int c = 0; c < nchunks; ++c) {
            lo_of[c] = (int)((long long)B * acc / 32);
            acc += w[c];
        }
        lo_of[nchunks] = B;
    } else {
        for (int c = 0; c <= nchunks; ++c) lo_of[c] = (c * per < B) ? c * per : B;
    }
    cudaStream_t sk = h->streams[0], sc = h->streams[1], sd = h->streams[2];
    for (int c = 0; c < pqp_handle::kStreams; ++c) h->h_done[c] = 0;
    PQP_CUDA(h, cudaMemsetAsync(h->d_ready, 0, pqp_handle::kStreams * sizeof(int), sk));
    PQP_CUDA(h, cudaMemsetAsync(h->d_done, 0, pqp_handle::kStreams * sizeof(int), sk));
    if (out->x_full) PQP_CUDA(h, cudaMemsetAsync(h->d_xf, 0, (size_t)B * nvm * sizeof(double), sk));
    if (out->y_full) PQP_CUDA(h, cudaMemsetAsync(h->d_yf, 0, (size_t)B * mm * sizeof(double), sk));
    if (out->z_full) PQP_CUDA(h, cudaMemsetAsync(h->d_zf, 0, (size_t)B * mm * sizeof(double), sk));
    PQP_CUDA(h, cudaEventRecord(h->ev0, sk));
    PQP_CUDA(h, cudaStreamWaitEvent(sc, h->ev0, 0));
    PQP_CUDA(h, cudaStreamWaitEvent(sd, h->ev0, 0));
    pqp_batch_in din = {B, nmax, h->d_knots, h->d_inst, h->d_n, (h->d_p_valid = in->p != nullptr) ? h->d_p : nullptr};
    pqp_batch_out dout = {h->d_sol, h->d_cost, h->d_status, h->d_iters, out->x_full ? h->d_xf : nullptr,
                          out->y_full ? h->d_yf : nullptr, out->z_full ? h->d_zf : nullptr, h->d_info};
    StreamedLaunch sl;
    sl.ready = h->d_ready;
    sl.done = h->d_done;
    sl.host_done = h->d_hdone;
    sl.n_chunks = nchunks;
    for (int c = 0; c <= pqp_handle::kStreams; ++c) sl.chunk_lo[c] = c <= nchunks ? lo_of[c] : B;
    // The chunk copies are queued BEFORE the launch: they are asynchronous (pinned buffers), so the
    // kernel still overlaps them, and a tool that makes launches blocking (ncu serialises kernels)
    // cannot leave the kernel waiting for copies the host has not issued yet.
    for (int c = 0; c < nchunks; ++c) {
        const int lo = lo_of[c], hi = lo_of[c + 1];
        const size_t nb = (size_t)(hi - lo);
        PQP_CUDA(h, cudaMemcpyAsync(h->d_knots + (size_t)lo * PQP_NFIELDS * nmax, in->knots + (size_t)lo * PQP_NFIELDS * nmax,
                                    nb * PQP_NFIELDS * nmax * sizeof(double), cudaMemcpyHostToDevice, sc));
        PQP_CUDA(h, cudaMemcpyAsync(h->d_inst + (size_t)lo * PQP_NINST, in->inst + (size_t)lo * PQP_NINST,
                                    nb * PQP_NINST * sizeof(double), cudaMemcpyHostToDevice, sc));
        PQP_CUDA(h, cudaMemcpyAsync(h->d_n + lo, in->n + lo, nb * sizeof(int), cudaMemcpyHostToDevice, sc));
        if (in->p) PQP_CUDA(h, cudaMemcpyAsync(h->d_p + lo, in->p + lo, nb * sizeof(int), cudaMemcpyHostToDevice, sc));
        PQP_CUDA(h, cudaMemsetAsync(h->d_ready + c, 1, sizeof(int), sc));  // "chunk c has landed"
    }
    int rc = run_device(h, &din, &dout, mode, sk, false, 0, false, h->d_flags, &sl);
    if (rc) return rc;
    h->host_inputs_resident = true;
    for (int c = 0; c < nchunks; ++c) {
        const int lo = lo_of[c], hi = lo_of[c + 1];
        const size_t nb = (size_t)(hi - lo);
        // wait until the kernel announces the chunk (or has ended, e.g. after an error)
        // (short spin, then 20 us naps: a chunk takes ~ a millisecond, and a caller that drives several handles
        // from several threads must not burn a core per handle)
        volatile int *flag = h->h_done + c;
        for (unsigned spin = 0; *flag == 0; ++spin) {
            if (spin < 4096) continue;
            if ((spin & 0x3f) == 0 && cudaStreamQuery(sk) != cudaErrorNotReady) break;
            const struct timespec nap = {0, 20000};
            nanosleep(&nap, nullptr);
        }
        PQP_CUDA(h, cudaMemcpyAsync(h->h_status + lo, h->d_status + lo, nb * sizeof(int), cudaMemcpyDeviceToHost, sd));
        PQP_CUDA(h, cudaMemcpyAsync(h->h_flags + lo, h->d_flags + lo, nb * sizeof(int), cudaMemcpyDeviceToHost, sd));
        PQP_CUDA(h, cudaMemcpyAsync(out->sol + (size_t)lo * 4 * nmax, h->d_sol + (size_t)lo * 4 * nmax, nb * 4 * nmax * sizeof(double), cudaMemcpyDeviceToHost, sd));
        if (out->cost) PQP_CUDA(h, cudaMemcpyAsync(out->cost + lo, h->d_cost + lo, nb * sizeof(double), cudaMemcpyDeviceToHost, sd));
        if (out->status) PQP_CUDA(h, cudaMemcpyAsync(out->status + lo, h->d_status + lo, nb * sizeof(int), cudaMemcpyDeviceToHost, sd));
        if (out->iters) PQP_CUDA(h, cudaMemcpyAsync(out->iters + lo, h->d_iters + lo, nb * sizeof(int), cudaMemcpyDeviceToHost, sd));
        if (out->info) PQP_CUDA(h, cudaMemcpyAsync(out->info + (size_t)lo * PQP_NINFO, h->d_info + (size_t)lo * PQP_NINFO, nb * PQP_NINFO * sizeof(double), cudaMemcpyDeviceToHost, sd));
        if (out->x_full) PQP_CUDA(h, cudaMemcpyAsync(out->x_full + lo * nvm, h->d_xf + lo * nvm, nb * nvm * sizeof(double), cudaMemcpyDeviceToHost, sd));
        if (out->y_full) PQP_CUDA(h, cudaMemcpyAsync(out->y_full + lo * mm, h->d_yf + lo * mm, nb * mm * sizeof(double), cudaMemcpyDeviceToHost, sd));
        if (out->z_full) PQP_CUDA(h, cudaMemcpyAsync(out->z_full + lo * mm, h->d_zf + lo * mm, nb * mm * sizeof(double), cudaMemcpyDeviceToHost, sd));
    }
    h->last_batch = B;
    PQP_CUDA(h, cudaStreamSynchronize(sk));
    PQP_CUDA(h, cudaStreamSynchronize(sc));
    PQP_CUDA(h, cudaEventRecord(h->ev1, sd));
    PQP_CUDA(h, cudaStreamSynchronize(sd));
    PQP_CUDA(h, cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    for (int i = 0; i < PQP_NSTAGES; ++i) h->stage_ms[i] = -1.0f;  // copies and kernel overlap: only the total is defined
    h->stage_ms[PQP_STAGE_TOTAL] = h->last_ms;
    if (!h->fp64 && h->escalate && mode == 0) return escalate_fp64(h, in, out, B);
    return PQP_OK;
}

int run_host(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out, int mode) {
    int rc;
    const int nmax = h->n_max;
    cudaStream_t s = h->stream;
    DeviceGuard guard_(h->device);
    int B;
    if (in) {
        rc = validate_batch(h, in, out);
        if (rc) return rc;
        B = in->batch;
        for (int b = 0; b < B; ++b) {
            if (in->n[b] < 2 || in->n[b] > nmax) return fail(h, PQP_E_INVALID, "n[b] must be in [2, n_max]");
            if (in->p && (in->p[b] < 0 || in->p[b] > in->n[b])) return fail(h, PQP_E_INVALID, "p[b] must be in [0, n[b]]");
        }
        if (mode == 1 && h->cold_only) return fail(h, PQP_E_STATE, "cold-only handle (option bit 128) keeps no warm state");
        if (mode == 1 && (!h->solved || B != h->last_batch)) return fail(h, PQP_E_STATE, "resolve needs a previous solve of the same batch");
        h->host_inputs_resident = true;
        h->d_p_valid = in->p != nullptr;
        if (h->use_tmem && B >= 512) {
            const size_t nvm0 = 6 * (size_t)nmax - 1, mm0 = 6 * (size_t)nmax + 2;
            if (out->x_full && !h->d_xf) PQP_CUDA(h, dmalloc(&h->d_xf, (size_t)h->batch_max * nvm0));
            if (out->y_full && !h->d_yf) PQP_CUDA(h, dmalloc(&h->d_yf, (size_t)h->batch_max * mm0));
            if (out->z_full && !h->d_zf) PQP_CUDA(h, dmalloc(&h->d_zf, (size_t)h->batch_max * mm0));
            return run_host_streamed(h, in, out, mode, B);
        }
    } else {
        if (mode != 1) return fail(h, PQP_E_INVALID, "null batch");
        if (h->cold_only) return fail(h, PQP_E_STATE, "cold-only handle (option bit 128) keeps no warm state");
        if (!out || !out->sol) return fail(h, PQP_E_INVALID, "null output");
        if (!h->solved || !h->host_inputs_resident) return fail(h, PQP_E_STATE, "resolve(NULL) needs a previous host-API solve");
        B = h->last_batch;
        const int threads = 256, blocks = (B * nmax + threads - 1) / threads;
        relinearise_kernel<<<blocks, threads, 0, s>>>(B, nmax, h->d_sol, h->d_knots);
        PQP_CUDA(h, cudaGetLastError());
        h->launches++;
    }
    // optional full outputs are allocated on first use
    const size_t nvm = 6 * (size_t)nmax - 1, mm = 6 * (size_t)nmax + 2;
    if (out->x_full && !h->d_xf) PQP_CUDA(h, dmalloc(&h->d_xf, (size_t)h->batch_max * nvm));
    if (out->y_full && !h->d_yf) PQP_CUDA(h, dmalloc(&h->d_yf, (size_t)h->batch_max * mm));
    if (out->z_full && !h->d_zf) PQP_CUDA(h, dmalloc(&h->d_zf, (size_t)h->batch_max * mm));
    // Chunked pipeline, one stream per chunk: chunk c+1's H2D copy and chunk c-1's D2H copy
    // overlap chunk c's kernel (fully asynchronous when the caller's buffers are pinned), and
    // because every chunk's kernel sits in its own stream the block scheduler back-fills SMs
    // from the next chunk as soon as its inputs have landed (no partial-wave bubble between
    // chunks). Scratch slots are addressed by the instance's index in the whole batch (qp0),
    // so chunking is invisible to the warm state.
    const int kMinChunk = 1024;
    int nchunks = (B + kMinChunk - 1) / kMinChunk;
    if (nchunks > pqp_handle::kStreams) nchunks = pqp_handle::kStreams;
    if (nchunks < 1) nchunks = 1;
    const int per = (B + nchunks - 1) / nchunks;
    cudaStream_t *streams = h->streams;
    // every chunk stream starts after whatever the handle's main stream has queued so far
    PQP_CUDA(h, cudaEventRecord(h->ev0, streams[0]));
    for (int c = 1; c < nchunks; ++c) PQP_CUDA(h, cudaStreamWaitEvent(streams[c], h->ev0, 0));
    for (int c = 0; c < nchunks; ++c) {
        const int lo = c * per, hi = (lo + per < B) ? lo + per : B;
        if (lo >= hi) break;
        const size_t nb = (size_t)(hi - lo);
        cudaStream_t s = streams[c];
        if (in) {
            PQP_CUDA(h, cudaMemcpyAsync(h->d_knots + (size_t)lo * PQP_NFIELDS * nmax, in->knots + (size_t)lo * PQP_NFIELDS * nmax,
                                        nb * PQP_NFIELDS * nmax * sizeof(double), cudaMemcpyHostToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->d_inst + (size_t)lo * PQP_NINST, in->inst + (size_t)lo * PQP_NINST,
                                        nb * PQP_NINST * sizeof(double), cudaMemcpyHostToDevice, s));
            PQP_CUDA(h, cudaMemcpyAsync(h->d_n + lo, in->n + lo, nb * sizeof(int), cudaMemcpyHostToDevice, s));
            if (in->p) PQP_CUDA(h, cudaMemcpyAsync(h->d_p + lo, in->p + lo, nb * sizeof(int), cudaMemcpyHostToDevice, s));
        }
        if (out->x_full) PQP_CUDA(h, cudaMemsetAsync(h->d_xf + lo * nvm, 0, nb * nvm * sizeof(double), s));
        if (out->y_full) PQP_CUDA(h, cudaMemsetAsync(h->d_yf + lo * mm, 0, nb * mm * sizeof(double), s));
        if (out->z_full) PQP_CUDA(h, cudaMemsetAsync(h->d_zf + lo * mm, 0, nb * mm * sizeof(double), s));
        if (nchunks == 1) PQP_CUDA(h, cudaEventRecord(h->ev_st[0], s));
        pqp_batch_in din;
        din.batch = (int32_t)nb;
        din.n_max = nmax;
        din.knots = h->d_knots + (size_t)lo * PQP_NFIELDS * nmax;
        din.inst = h->d_inst + (size_t)lo * PQP_NINST;
        din.n = h->d_n + lo;
        din.p = h->d_p_valid ? h->d_p + lo : nullptr;
        pqp_batch_out dout;
        dout.sol = h->d_sol + (size_t)lo * 4 * nmax;
        dout.cost = h->d_cost + lo;
        dout.status = h->d_status + lo;
        dout.iters = h->d_iters + lo;
        dout.x_full = out->x_full ? h->d_xf + lo * nvm : nullptr;
        dout.y_full = out->y_full ? h->d_yf + lo * mm : nullptr;
        dout.z_full = out->z_full ? h->d_zf + lo * mm : nullptr;
        dout.info = h->d_info + (size_t)lo * PQP_NINFO;
        rc = run_device(h, &din, &dout, mode, s, false, lo, false, h->d_flags + lo);
        if (rc) return rc;
        if (nchunks == 1) PQP_CUDA(h, cudaEventRecord(h->ev_st[1], s));
        PQP_CUDA(h, cudaMemcpyAsync(h->h_status + lo, dout.status, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
        PQP_CUDA(h, cudaMemcpyAsync(h->h_flags + lo, h->d_flags + lo, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
        PQP_CUDA(h, cudaMemcpyAsync(out->sol + (size_t)lo * 4 * nmax, dout.sol, nb * 4 * nmax * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (out->cost) PQP_CUDA(h, cudaMemcpyAsync(out->cost + lo, dout.cost, nb * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (out->status) PQP_CUDA(h, cudaMemcpyAsync(out->status + lo, dout.status, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (out->iters) PQP_CUDA(h, cudaMemcpyAsync(out->iters + lo, dout.iters, nb * sizeof(int), cudaMemcpyDeviceToHost, s));
        if (out->info) PQP_CUDA(h, cudaMemcpyAsync(out->info + (size_t)lo * PQP_NINFO, dout.info, nb * PQP_NINFO * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (out->x_full) PQP_CUDA(h, cudaMemcpyAsync(out->x_full + lo * nvm, dout.x_full, nb * nvm * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (out->y_full) PQP_CUDA(h, cudaMemcpyAsync(out->y_full + lo * mm, dout.y_full, nb * mm * sizeof(double), cudaMemcpyDeviceToHost, s));
        if (out->z_full) PQP_CUDA(h, cudaMemcpyAsync(out->z_full + lo * mm, dout.z_full, nb * mm * sizeof(double), cudaMemcpyDeviceToHost, s));
    }
    h->last_batch = B;
    // join: stream[0] waits for every chunk stream; ev1 marks the end of the whole pipeline
    for (int c = 1; c < nchunks; ++c) {
        PQP_CUDA(h, cudaEventRecord(h->evs[c], streams[c]));
        PQP_CUDA(h, cudaStreamWaitEvent(streams[0], h->evs[c], 0));
    }
    PQP_CUDA(h, cudaEventRecord(h->ev1, streams[0]));
    PQP_CUDA(h, cudaStreamSynchronize(streams[0]));
    PQP_CUDA(h, cudaEventElapsedTime(&h->last_ms, h->ev0, h->ev1));
    // per-stage device times of a single-stream call (the reference's TimeRecorder stages of
    // BaseSolver::solve, base_solver.cpp:57-93; here cost / constraints / solver set-up / ADMM are one kernel)
    for (int i = 0; i < PQP_NSTAGES; ++i) h->stage_ms[i] = -1.0f;
    if (nchunks == 1) {
        PQP_CUDA(h, cudaEventElapsedTime(&h->stage_ms[PQP_STAGE_H2D], h->ev0, h->ev_st[0]));
        PQP_CUDA(h, cudaEventElapsedTime(&h->stage_ms[PQP_STAGE_KERNEL], h->ev_st[0], h->ev_st[1]));
        PQP_CUDA(h, cudaEventElapsedTime(&h->stage_ms[PQP_STAGE_D2H], h->ev_st[1], h->ev1));
    }
    h->stage_ms[PQP_STAGE_TOTAL] = h->last_ms;
    h->stage_ms[PQP_STAGE_ESCALATION] = 0.0f;
    // cold solves only: the FP64 re-solve starts cold, which is not what a warm re-solve under an iteration
    // cap (receding horizon) is; the increment-form kernel certifies infeasibility by itself there
    if (!h->fp64 && h->escalate && mode == 0) {
        PQP_CUDA(h, cudaEventRecord(h->ev_st[2], streams[0]));
        rc = escalate_fp64(h, in, out, B);
        if (rc) return rc;
        PQP_CUDA(h, cudaEventRecord(h->ev_st[3], streams[0]));
        PQP_CUDA(h, cudaEventSynchronize(h->ev_st[3]));
        PQP_CUDA(h, cudaEventElapsedTime(&h->stage_ms[PQP_STAGE_ESCALATION], h->ev_st[2], h->ev_st[3]));
    }
    return PQP_OK;
}

}  // namespace

// ------------------------------------------------------------------ C ABI
extern "C" {

int pqp_version(void) { return PQP_VERSION; }

int pqp_default_params(pqp_params *params) {
    if (!params) return PQP_E_INVALID;
    pqp::default_params(params);
    return PQP_OK;
}

const char *pqp_last_error(const pqp_handle *h) { return h ? h->err.c_str() : g_create_error.c_str(); }

int pqp_create(const pqp_params *params, int32_t n_max, int32_t batch_max, int32_t device,
               pqp_handle **out) {
    if (!out) return fail(nullptr, PQP_E_INVALID, "out is null");
    *out = nullptr;
    if (!params || !pqp::params_valid(*params)) return fail(nullptr, PQP_E_INVALID, "invalid params");
    if (n_max < 2 || n_max > 32 * pqp::kMaxChunk - 1) return fail(nullptr, PQP_E_INVALID, "n_max must be in [2, 511]");
    if (batch_max < 1) return fail(nullptr, PQP_E_INVALID, "batch_max must be >= 1");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(nullptr, PQP_E_NO_DEVICE, std::string("no CUDA device: ") + cudaGetErrorString(e));
    if (device < 0 || device >= ndev) return fail(nullptr, PQP_E_NO_DEVICE, "device index out of range");
    cudaDeviceProp prop;
    if ((e = cudaGetDeviceProperties(&prop, device)) != cudaSuccess) return cuda_fail(nullptr, e, "cudaGetDeviceProperties");
    if (prop.major != 10)
        return fail(nullptr, PQP_E_NO_DEVICE, "this library is built for sm_100a (B200) only; found sm_" +
                                                  std::to_string(prop.major) + std::to_string(prop.minor));
    pqp_handle *h = new (std::nothrow) pqp_handle;
    if (!h) return fail(nullptr, PQP_E_INVALID, "out of host memory");
    h->prm = *params;
    h->n_max = n_max;
    h->batch_max = batch_max;
    h->device = device;
    h->chunk = pqp::chunk_for(n_max);
    h->sm_count = prop.multiProcessorCount;
    h->smem_bytes = pqp::smem_floats(h->chunk) * sizeof(float) + 16;
#define PQP_CREATE_CUDA(call)                                                     \
    do {                                                                          \
        cudaError_t e_ = (call);                                                  \
        if (e_ != cudaSuccess) {                                                  \
            g_create_error = std::string(#call) + ": " + cudaGetErrorString(e_);  \
            pqp_destroy(h);                                                       \
            return PQP_E_CUDA;                                                    \
        }                                                                         \
    } while (0)
    DeviceGuard guard_(device);
    h->fp64 = (params->reserved & 2) != 0;
    h->escalate = (params->reserved & 4) == 0;
    h->cold_only = (params->reserved & 128) != 0;
    // storage policy of the FP32 kernel (measured, profiles/r2/policy_ab.log): tensor memory for n_max >= 64
    // (chunks of 4, 8, 16 stages per lane: 8 / 4 / 2 QPs per SM; at n = 240 shared memory would hold 3 per SM, at
    // n = 120 it holds 6 but runs 3.48 ms against 3.21 ms per 8192 instances, at 1024 instances 0.92 vs 0.76 ms),
    // shared memory below (n = 60: 1.59 vs 1.82 ms). Bits 8 / 16 force one or the other.
    const bool auto_tmem = h->chunk >= 4;
    h->use_tmem = !h->fp64 && ((params->reserved & 8) != 0 || (auto_tmem && (params->reserved & 16) == 0));
    // form of the ADMM step (FP32 kernels): the increment form (dx solve, row values A x carried and
    // advanced by alpha A dx) follows the FP64 oracle's rho schedule - same iteration count in ~99-100 %
    // of instances, ~6 % fewer iterations - for ~6 % more work per iteration. Measured against the
    // textbook form: n = 240 equal (892 k solves/s), n = 120 +5 %, the 1024 x 120 shared-map batch
    // +15 % (one wave: its slowest instances need fewer iterations), n = 60 -3.5 %. Default: increment
    // form for n_max >= 64, textbook form below; bits 32 / 64 force one or the other.
    h->incr = !h->fp64 && ((params->reserved & 32) != 0 || ((params->reserved & 64) == 0 && h->chunk >= 4));
    if (h->use_tmem) PQP_CREATE_CUDA(prepare_tmem_chunk(h->chunk));
    h->smem_bytes64 = pqp::smem_floats(h->chunk) * sizeof(double) + 16;
    int bps = 0;
    if (h->fp64) {
        PQP_CREATE_CUDA(prepare_chunk<double>(h->chunk, h->smem_bytes64, &bps));
        h->prepared64 = true;
    } else {
        PQP_CREATE_CUDA(prepare_chunk<float>(h->chunk, h->smem_bytes, &bps, h->incr));
    }
    h->warps_per_sm = h->use_tmem ? tmem_warps(h->chunk) : bps;
    for (int i = 0; i < pqp_handle::kStreams; ++i) {
        PQP_CREATE_CUDA(cudaStreamCreateWithFlags(&h->streams[i], cudaStreamNonBlocking));
        PQP_CREATE_CUDA(cudaEventCreateWithFlags(&h->evs[i], cudaEventDisableTiming));
    }
    h->stream = h->streams[0];
    PQP_CREATE_CUDA(cudaEventCreate(&h->ev0));
    PQP_CREATE_CUDA(cudaEventCreate(&h->ev1));
    const size_t B = batch_max, c = h->chunk;
    PQP_CREATE_CUDA(dmalloc(&h->d_knots, B * PQP_NFIELDS * n_max));
    PQP_CREATE_CUDA(dmalloc(&h->d_inst, B * PQP_NINST));
    PQP_CREATE_CUDA(dmalloc(&h->d_n, B));
    PQP_CREATE_CUDA(dmalloc(&h->d_p, B));
    PQP_CREATE_CUDA(dmalloc(&h->d_sol, B * 4 * n_max));
    PQP_CREATE_CUDA(dmalloc(&h->d_cost, B));
    PQP_CREATE_CUDA(dmalloc(&h->d_status, B));
    PQP_CREATE_CUDA(dmalloc(&h->d_iters, B));
    PQP_CREATE_CUDA(dmalloc(&h->d_info, B * PQP_NINFO));
    const size_t esz = h->fp64 ? sizeof(double) : sizeof(float);
    // Warm state (scaled x, z, y and rho: the OSQP workspace that persists between solve and
    // updateProblemFormulationAndSolve) is per instance; a cold-only handle keeps none. The Ruiz
    // factors and delta_y are scratch of one solve: per instance under the one-CTA-per-instance
    // launch, per resident warp under the persistent launch (592 blocks that live in L2).
    if (!h->cold_only) {
        PQP_CREATE_CUDA(cudaMalloc(&h->d_warm, B * pqp::warm_floats(c) * esz));
        PQP_CREATE_CUDA(cudaMemset(h->d_warm, 0, B * pqp::warm_floats(c) * esz));
        PQP_CREATE_CUDA(cudaMalloc(&h->d_rho, B * esz));
    }
    const size_t slots = h->use_tmem ? (size_t)h->sm_count * tmem_warps(h->chunk) : B;
    PQP_CREATE_CUDA(cudaMalloc(&h->d_scal, slots * pqp::scal_floats(c) * esz));
    PQP_CREATE_CUDA(cudaMalloc(&h->d_dy, slots * pqp::dy_floats(c) * esz));
    for (int i = 0; i < 4; ++i) PQP_CREATE_CUDA(cudaEventCreate(&h->ev_st[i]));
    PQP_CREATE_CUDA(dmalloc(&h->d_counters, (size_t)pqp_handle::kCounters));
    PQP_CREATE_CUDA(dmalloc(&h->d_ready, (size_t)pqp_handle::kStreams));
    PQP_CREATE_CUDA(dmalloc(&h->d_done, (size_t)pqp_handle::kStreams));
    PQP_CREATE_CUDA(cudaHostAlloc(reinterpret_cast<void **>(&h->h_done), pqp_handle::kStreams * sizeof(int), cudaHostAllocMapped));
    PQP_CREATE_CUDA(cudaHostGetDevicePointer(reinterpret_cast<void **>(&h->d_hdone), h->h_done, 0));
    PQP_CREATE_CUDA(dmalloc(&h->d_flags, B));
    PQP_CREATE_CUDA(cudaMemset(h->d_flags, 0, B * sizeof(int)));
    PQP_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h->h_status), B * sizeof(int)));
    PQP_CREATE_CUDA(cudaMallocHost(reinterpret_cast<void **>(&h->h_flags), B * sizeof(int)));
    PQP_CREATE_CUDA(cudaMemset(h->d_sol, 0, B * 4 * n_max * sizeof(double)));
#undef PQP_CREATE_CUDA
    *out = h;
    return PQP_OK;
}

int pqp_destroy(pqp_handle *h) {
    if (!h) return PQP_OK;
    DeviceGuard guard_(h->device);
    for (int i = 0; i < pqp_handle::kStreams; ++i) if (h->streams[i]) cudaStreamSynchronize(h->streams[i]);
    cudaFree(h->d_knots); cudaFree(h->d_inst); cudaFree(h->d_n); cudaFree(h->d_p);
    cudaFree(h->d_sol); cudaFree(h->d_cost); cudaFree(h->d_status); cudaFree(h->d_iters);
    cudaFree(h->d_info); cudaFree(h->d_xf); cudaFree(h->d_yf); cudaFree(h->d_zf);
    cudaFree(h->d_ref); cudaFree(h->d_xy); cudaFree(h->d_sol2); cudaFree(h->d_n2);
    cudaFree(h->d_warm); cudaFree(h->d_scal); cudaFree(h->d_dy); cudaFree(h->d_rho);
    cudaFree(h->d_flags);
    cudaFree(h->d_state64);
    cudaFree(h->d_counters);
    cudaFree(h->d_ready);
    cudaFree(h->d_done);
    if (h->h_done) cudaFreeHost(h->h_done);
    if (h->h_status) cudaFreeHost(h->h_status);
    if (h->h_flags) cudaFreeHost(h->h_flags);
    cudaFree(h->e_knots); cudaFree(h->e_inst); cudaFree(h->e_sol); cudaFree(h->e_cost); cudaFree(h->e_info);
    cudaFree(h->e_xf); cudaFree(h->e_yf); cudaFree(h->e_zf); cudaFree(h->e_n); cudaFree(h->e_p);
    cudaFree(h->e_slots); cudaFree(h->e_status); cudaFree(h->e_iters); cudaFree(h->e_warm); cudaFree(h->e_scal); cudaFree(h->e_dy); cudaFree(h->e_rho);
    if (h->ev0) cudaEventDestroy(h->ev0);
    if (h->ev1) cudaEventDestroy(h->ev1);
    for (int i = 0; i < 4; ++i) if (h->ev_st[i]) cudaEventDestroy(h->ev_st[i]);
    for (int i = 0; i < pqp_handle::kStreams; ++i) {
        if (h->streams[i]) cudaStreamDestroy(h->streams[i]);
        if (h->evs[i]) cudaEventDestroy(h->evs[i]);
    }
    delete h;
    return PQP_OK;
}

int pqp_solve(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out) {
    if (!h) return PQP_E_INVALID;
    if (!in) return fail(h, PQP_E_INVALID, "null batch");
    return run_host(h, in, out, 0);
}

int pqp_resolve(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out) {
    if (!h) return PQP_E_INVALID;
    return run_host(h, in, out, 1);
}

int pqp_solve_device(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out, void *stream) {
    int rc = validate_batch(h, in, out);
    if (rc) return rc;
    DeviceGuard guard_(h->device);
    h->host_inputs_resident = false;
    return run_device(h, in, out, 0, static_cast<cudaStream_t>(stream), false);
}

int pqp_resolve_device(pqp_handle *h, const pqp_batch_in *in, const pqp_batch_out *out, void *stream) {
    int rc = validate_batch(h, in, out);
    if (rc) return rc;
    if (h->cold_only) return fail(h, PQP_E_STATE, "cold-only handle (option bit 128) keeps no warm state");
    if (!h->solved || in->batch != h->last_batch) return fail(h, PQP_E_STATE, "resolve needs a previous solve of the same batch");
    DeviceGuard guard_(h->device);
    return run_device(h, in, out, 1, static_cast<cudaStream_t>(stream), false);
}

int pqp_frenet_to_cartesian_device(pqp_handle *h, int32_t batch, const int32_t *n, const double *ref_xyh,
                                   const double *sol, double *out_xyh, void *stream) {
    if (!h) return PQP_E_INVALID;
    if (!n || !ref_xyh || !sol || !out_xyh || batch < 1) return fail(h, PQP_E_INVALID, "null buffer");
    DeviceGuard guard_(h->device);
    const int threads = 256, blocks = (batch * h->n_max + threads - 1) / threads;
    frenet_to_cartesian_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(batch, h->n_max, n, ref_xyh, sol, out_xyh);
    PQP_CUDA(h, cudaGetLastError());
    h->launches++;
    return PQP_OK;
}

int pqp_relinearise_device(pqp_handle *h, int32_t batch, const double *sol, double *knots, void *stream) {
    if (!h) return PQP_E_INVALID;
    if (!sol || !knots || batch < 1) return fail(h, PQP_E_INVALID, "null buffer");
    DeviceGuard guard_(h->device);
    const int threads = 256, blocks = (batch * h->n_max + threads - 1) / threads;
    relinearise_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(batch, h->n_max, sol, knots);
    PQP_CUDA(h, cudaGetLastError());
    h->launches++;
    return PQP_OK;
}

int pqp_advance_window_device(pqp_handle *h, int32_t batch, int32_t ext_len, int32_t tick, const double *ext_knots,
                              const double *sol, double *knots, double *inst, void *stream) {
    if (!h) return PQP_E_INVALID;
    if (!ext_knots || !sol || !knots || !inst || batch < 1) return fail(h, PQP_E_INVALID, "null buffer");
    if (tick < 0 || tick + h->n_max > ext_len) return fail(h, PQP_E_INVALID, "window [tick, tick + n_max) leaves the extended reference");
    DeviceGuard guard_(h->device);
    const int threads = 256, blocks = (batch * h->n_max + threads - 1) / threads;
    advance_window_kernel<<<blocks, threads, 0, static_cast<cudaStream_t>(stream)>>>(batch, h->n_max, ext_len, tick, ext_knots, sol,
                                                                                   knots, inst);
    PQP_CUDA(h, cudaGetLastError());
    h->launches++;
    return PQP_OK;
}

int pqp_frenet_to_cartesian(pqp_handle *h, int32_t batch, const int32_t *n, const double *ref_xyh,
                            const double *sol, double *out_xyh) {
    if (!h) return PQP_E_INVALID;
    if (!n || !ref_xyh || !sol || !out_xyh || batch < 1 || batch > h->batch_max) return fail(h, PQP_E_INVALID, "bad arguments");
    DeviceGuard guard_(h->device);
    const size_t B = batch, nm = h->n_max;
    if (!h->d_ref) PQP_CUDA(h, dmalloc(&h->d_ref, (size_t)h->batch_max * 3 * nm));
    if (!h->d_xy) PQP_CUDA(h, dmalloc(&h->d_xy, (size_t)h->batch_max * 3 * nm));
    if (!h->d_sol2) PQP_CUDA(h, dmalloc(&h->d_sol2, (size_t)h->batch_max * 4 * nm));
    if (!h->d_n2) PQP_CUDA(h, dmalloc(&h->d_n2, (size_t)h->batch_max));
    cudaStream_t s = h->stream;
    PQP_CUDA(h, cudaMemcpyAsync(h->d_ref, ref_xyh, B * 3 * nm * sizeof(double), cudaMemcpyHostToDevice, s));
    PQP_CUDA(h, cudaMemcpyAsync(h->d_sol2, sol, B * 4 * nm * sizeof(double), cudaMemcpyHostToDevice, s));
    PQP_CUDA(h, cudaMemcpyAsync(h->d_n2, n, B * sizeof(int), cudaMemcpyHostToDevice, s));
    PQP_CUDA(h, cudaMemsetAsync(h->d_xy, 0, B * 3 * nm * sizeof(double), s));
    int rc = pqp_frenet_to_cartesian_device(h, batch, h->d_n2, h->d_ref, h->d_sol2, h->d_xy, s);
    if (rc) return rc;
    PQP_CUDA(h, cudaMemcpyAsync(out_xyh, h->d_xy, B * 3 * nm * sizeof(double), cudaMemcpyDeviceToHost, s));
    PQP_CUDA(h, cudaStreamSynchronize(s));
    return PQP_OK;
}

int pqp_last_kernel_ms(pqp_handle *h, float *ms) {
    if (!h || !ms) return PQP_E_INVALID;
    *ms = h->last_ms;
    return PQP_OK;
}

int pqp_pack_results_device(pqp_handle *h, int32_t batch, const double *cost, const int32_t *status,
                            const int32_t *iters, pqp_result_rec *packed, void *stream) {
    if (!h) return PQP_E_INVALID;
    if (!cost || !status || !iters || !packed || batch < 1) return fail(h, PQP_E_INVALID, "null buffer");
    DeviceGuard guard_(h->device);
    pack_results_kernel<<<(batch + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(batch, cost, status, iters, packed);
    PQP_CUDA(h, cudaGetLastError());
    h->launches++;
    return PQP_OK;
}

int pqp_resident_results(pqp_handle *h, const double **sol, const double **cost, const int32_t **status,
                         const int32_t **iters) {
    if (!h) return PQP_E_INVALID;
    if (!h->solved || !h->host_inputs_resident) return fail(h, PQP_E_STATE, "no host-pointer solve has run on this handle");
    if (sol) *sol = h->d_sol;
    if (cost) *cost = h->d_cost;
    if (status) *status = h->d_status;
    if (iters) *iters = h->d_iters;
    return PQP_OK;
}

int pqp_last_stage_ms(pqp_handle *h, float *ms) {
    if (!h || !ms) return PQP_E_INVALID;
    for (int i = 0; i < PQP_NSTAGES; ++i) ms[i] = h->stage_ms[i];
    return PQP_OK;
}

int pqp_launch_count(pqp_handle *h, int64_t *count) {
    if (!h || !count) return PQP_E_INVALID;
    *count = h->launches;
    return PQP_OK;
}

int pqp_kernel_info(pqp_handle *h, int32_t *sm_count, int32_t *warps_per_sm, int32_t *smem_per_warp) {
    if (!h) return PQP_E_INVALID;
    if (sm_count) *sm_count = h->sm_count;
    if (warps_per_sm) *warps_per_sm = h->warps_per_sm;
    // tensor-memory policy: only the groups that do not fit the warp's columns live in shared memory
    if (smem_per_warp)
        *smem_per_warp = h->use_tmem ? (int32_t)tmem_spill_bytes(h->chunk) : (int32_t)h->smem_bytes;
    return PQP_OK;
}

}  // extern "C"
