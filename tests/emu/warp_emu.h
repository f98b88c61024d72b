// warp_emu.h — lock-step warp emulator for the CPU-only build container.  TEST ONLY.
//
// The build container has no GPU, so `-m "not gpu"` tests compile the *same* kernel
// source (path_optimizer_2_b200/csrc/pqp_kernel.cuh) for the host with PQP_EMU defined
// and run each 32-lane warp as 32 cooperative fibers (ucontext). Warp shuffles are an
// exchange through a slot array with a fiber barrier. This file is never part of
// libpqp_b200.so; the product has no CPU path.
#pragma once
#include <ucontext.h>

#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

namespace warp_emu {

constexpr int kWarp = 32;
constexpr size_t kStack = 1024 * 1024;

struct Warp {
    ucontext_t main_ctx;
    ucontext_t ctx[kWarp];
    std::vector<char> stacks;
    bool finished[kWarp];
    unsigned nshfl[kWarp];
    int cur = 0;
    int active = 0;
    int arrived = 0;
    unsigned gen = 0;
    uint64_t slots[2][kWarp];
    std::function<void(int)> body;

    Warp() : stacks(kStack * kWarp) {}

    static void trampoline(unsigned lo, unsigned hi) {
        Warp *w = reinterpret_cast<Warp *>((uintptr_t(hi) << 32) | uintptr_t(lo));
        int lane = w->cur;
        w->body(lane);
        w->finished[lane] = true;
        w->active--;
        if (w->active > 0 && w->arrived == w->active) { w->arrived = 0; w->gen++; }
        w->switch_next(lane);
    }

    void switch_next(int from) {
        for (int i = 1; i <= kWarp; ++i) {
            int nxt = (from + i) % kWarp;
            if (!finished[nxt]) {
                cur = nxt;
                if (finished[from]) setcontext(&ctx[nxt]);
                else swapcontext(&ctx[from], &ctx[nxt]);
                return;
            }
        }
        setcontext(&main_ctx);  // all lanes done
    }

    void run(std::function<void(int)> f) {
        body = std::move(f);
        active = kWarp;
        arrived = 0;
        for (int l = 0; l < kWarp; ++l) {
            finished[l] = false;
            nshfl[l] = 0;
            getcontext(&ctx[l]);
            ctx[l].uc_stack.ss_sp = stacks.data() + kStack * l;
            ctx[l].uc_stack.ss_size = kStack;
            ctx[l].uc_link = nullptr;
            uintptr_t p = reinterpret_cast<uintptr_t>(this);
            makecontext(&ctx[l], (void (*)())trampoline, 2, unsigned(p & 0xffffffffu),
                        unsigned(p >> 32));
        }
        cur = 0;
        swapcontext(&main_ctx, &ctx[0]);
    }

    // all unfinished lanes must call this the same number of times
    void barrier(int lane) {
        unsigned my = gen;
        arrived++;
        if (arrived == active) { arrived = 0; gen++; return; }
        while (gen == my) switch_next(lane);
    }
};

inline Warp *&current() {
    static thread_local Warp *w = nullptr;
    return w;
}

template <typename T>
inline T exchange(T v, int lane, int src_lane) {
    static_assert(sizeof(T) <= 8, "shuffle payload too large");
    Warp *w = current();
    unsigned par = (w->nshfl[lane]++) & 1u;
    uint64_t raw = 0;
    std::memcpy(&raw, &v, sizeof(T));
    w->slots[par][lane] = raw;
    w->barrier(lane);
    T out;
    if (src_lane < 0 || src_lane >= kWarp) src_lane = lane;
    raw = w->slots[par][src_lane];
    std::memcpy(&out, &raw, sizeof(T));
    return out;
}

}  // namespace warp_emu
