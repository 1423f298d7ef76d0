// Test-only SIMT emulation: lets the warp- and CTA-synchronous DEVICE code of the product headers
// (csrc/blob_device.cuh, csrc/match_device.cuh) run unchanged on the host, one std::thread per CUDA thread,
// so that the algorithms the kernels execute can be checked against the oracle on a machine without a GPU.
// Barriers and warp collectives are rendezvous points of the participating threads; shared-memory atomics
// are host atomics on plain memory.  Only what those headers use is provided.  NOT part of libmocap_b200.so.
#pragma once
#include <cuda_runtime.h>          // vector types, attribute macros (the CUDA intrinsics are not declared for g++)
#include <stdint.h>
#include <string.h>
#include <atomic>
#include <barrier>
#include <chrono>
#include <stdio.h>
#include <stdlib.h>
#include <functional>
#include <memory>
#include <thread>
#include <vector>

namespace simt {

struct Block {
    int nt;
    std::barrier<> cta;
    std::vector<std::unique_ptr<std::barrier<>>> warp;
    std::vector<unsigned long long> xchg;            // one slot per thread for the warp collectives
    explicit Block(int n) : nt(n), cta(n), xchg(n) {
        for (int w = 0; w < (n + 31) / 32; ++w) {
            const int members = (w + 1) * 32 <= n ? 32 : n - w * 32;
            warp.emplace_back(new std::barrier<>(members));
        }
    }
};
struct Tid { unsigned x, y, z; };
inline thread_local Tid tid{0, 0, 0};
// SIMT_TRACE=1 in the environment: a watchdog prints, for a launch that has not finished after 20 s, which
// rendezvous every thread last entered (a deadlock in emulated code is a divergent barrier in the device code)
inline const char* volatile last_op[4096];
inline const char* volatile last_where[4096];
inline thread_local Block* blk = nullptr;
// grids of several CTAs (persistent kernels with a grid-wide barrier, csrc/ba_device.cuh)
inline thread_local Tid bid{0, 0, 0};
inline Tid gdim{1, 1, 1};
inline Tid bdim{1, 1, 1};
inline std::barrier<>* grid_barrier = nullptr;

inline unsigned slot() { return (bid.x * bdim.x + tid.x) & 4095u; }        // trace slot of this thread, unique across the CTAs of a grid
inline void warp_sync(const char* what = "warp_sync") { last_op[slot()] = what; blk->warp[tid.x >> 5]->arrive_and_wait(); last_op[slot()] = "running"; }
inline unsigned long long exchange(unsigned long long v, int src_lane) {      // value of lane src_lane of my warp
    blk->xchg[tid.x] = v;
    warp_sync("shuffle");
    const unsigned long long r = blk->xchg[(tid.x & ~31u) + (unsigned)src_lane];
    warp_sync("shuffle (2)");
    return r;
}

// run fn on nt threads that form one CTA (threadIdx.x = 0 .. nt-1)
inline void launch(int nt, const std::function<void()>& fn) {
    Block b(nt);
    std::vector<std::thread> th;
    std::atomic<int> running{nt};
    for (int t = 0; t < nt; ++t)
        th.emplace_back([&, t] { tid = Tid{(unsigned)t, 0, 0}; blk = &b; last_op[t] = "running"; fn(); last_op[t] = "exited"; --running; });
    std::thread watchdog;
    if (getenv("SIMT_TRACE"))
        watchdog = std::thread([&] {
            for (int s = 0; s < 200 && running.load() > 0; ++s) std::this_thread::sleep_for(std::chrono::milliseconds(100));
            if (running.load() > 0)
                for (int t = 0; t < nt; ++t) fprintf(stderr, "simt: thread %d (warp %d lane %d): %s at %s\n", t, t >> 5, t & 31, last_op[t], last_where[t] ? last_where[t] : "?");
        });
    for (auto& t : th) t.join();
    if (watchdog.joinable()) watchdog.join();
}

// run fn on nb CTAs of nt threads each (blockIdx.x = 0 .. nb-1), all co-resident; simt_grid_sync() is a rendezvous
// of every thread of the grid
inline void launch_grid(int nb, int nt, const std::function<void()>& fn) {
    std::vector<std::unique_ptr<Block>> blocks;
    for (int b = 0; b < nb; ++b) blocks.emplace_back(new Block(nt));
    std::barrier<> gb(nb * nt);
    grid_barrier = &gb;
    gdim = Tid{(unsigned)nb, 1, 1};
    bdim = Tid{(unsigned)nt, 1, 1};
    std::vector<std::thread> th;
    for (int b = 0; b < nb; ++b)
        for (int t = 0; t < nt; ++t)
            th.emplace_back([&, b, t] { tid = Tid{(unsigned)t, 0, 0}; bid = Tid{(unsigned)b, 0, 0}; blk = blocks[b].get(); fn(); });
    for (auto& t : th) t.join();
    grid_barrier = nullptr;
    gdim = Tid{1, 1, 1};
}

}  // namespace simt

#define blockIdx (simt::bid)
#define gridDim (simt::gdim)
#define blockDim (simt::bdim)
inline void simt_grid_sync() { simt::grid_barrier->arrive_and_wait(); }

#ifndef __launch_bounds__
#define __launch_bounds__(...)
#endif
#ifndef __noinline__
#define __noinline__ __attribute__((noinline))
#endif

// ---- what the device headers call ------------------------------------------------------------------------
#define threadIdx (simt::tid)
inline void simt_syncthreads(const char* where) { simt::last_where[simt::slot()] = where; simt::last_op[simt::slot()] = "__syncthreads"; simt::blk->cta.arrive_and_wait(); simt::last_op[simt::slot()] = "running"; }
inline void simt_syncwarp(const char* where) { simt::last_where[simt::slot()] = where; simt::warp_sync("__syncwarp"); }
inline void __threadfence() { std::atomic_thread_fence(std::memory_order_seq_cst); }

template <typename T> inline T simt_shfl(T v, int src, const char* where) {
    simt::last_where[simt::slot()] = where;
    static_assert(sizeof(T) <= 8, "shuffle width");
    unsigned long long raw = 0; memcpy(&raw, &v, sizeof(T));
    raw = simt::exchange(raw, src & 31);
    T r; memcpy(&r, &raw, sizeof(T)); return r;
}
template <typename T> inline T simt_shfl_up(T v, unsigned d, const char* where) {
    const int lane = (int)(simt::tid.x & 31);
    return simt_shfl(v, lane >= (int)d ? lane - (int)d : lane, where);
}
template <typename T> inline T simt_shfl_down(T v, unsigned d, const char* where) {
    const int lane = (int)(simt::tid.x & 31);
    return simt_shfl(v, lane + (int)d < 32 ? lane + (int)d : lane, where);
}
template <typename T> inline T simt_shfl_xor(T v, int x, const char* where) { return simt_shfl(v, (int)(simt::tid.x & 31) ^ x, where); }
inline unsigned simt_ballot(int pred, const char* where) {
    simt::last_where[simt::slot()] = where;
    simt::blk->xchg[simt::tid.x] = pred ? 1ull : 0ull;
    simt::warp_sync("ballot");
    unsigned r = 0;
    const unsigned base = simt::tid.x & ~31u;
    for (unsigned l = 0; l < 32 && base + l < (unsigned)simt::blk->nt; ++l) r |= (unsigned)simt::blk->xchg[base + l] << l;
    simt::warp_sync();
    return r;
}
inline unsigned simt_match_any(unsigned v, const char* where) {
    simt::last_where[simt::slot()] = where;
    simt::blk->xchg[simt::tid.x] = v;
    simt::warp_sync("match_any");
    unsigned r = 0;
    const unsigned base = simt::tid.x & ~31u;
    for (unsigned l = 0; l < 32 && base + l < (unsigned)simt::blk->nt; ++l) r |= (simt::blk->xchg[base + l] == v ? 1u : 0u) << l;
    simt::warp_sync();
    return r;
}
#define SIMT_STR2(x) #x
#define SIMT_STR(x) SIMT_STR2(x)
#define SIMT_HERE __FILE__ ":" SIMT_STR(__LINE__)
#define __syncthreads() simt_syncthreads(SIMT_HERE)
#define __syncwarp(...) simt_syncwarp(SIMT_HERE)
#define __shfl_sync(m, v, s) simt_shfl((v), (s), SIMT_HERE)
#define __shfl_up_sync(m, v, d) simt_shfl_up((v), (d), SIMT_HERE)
#define __shfl_down_sync(m, v, d) simt_shfl_down((v), (d), SIMT_HERE)
#define __shfl_xor_sync(m, v, x) simt_shfl_xor((v), (x), SIMT_HERE)
#define __ballot_sync(m, p) simt_ballot((p), SIMT_HERE)
#define __match_any_sync(m, v) simt_match_any((v), SIMT_HERE)

inline long long __double_as_longlong(double d) { long long r; memcpy(&r, &d, 8); return r; }
inline double __longlong_as_double(long long v) { double r; memcpy(&r, &v, 8); return r; }
inline int __float_as_int(float f) { int r; memcpy(&r, &f, 4); return r; }
inline float __int_as_float(int v) { float r; memcpy(&r, &v, 4); return r; }
inline int __popc(unsigned v) { return __builtin_popcount(v); }
inline int __ffs(int v) { return __builtin_ffs(v); }
inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
template <typename T> inline T __ldcg(const T* p) { return *reinterpret_cast<const volatile T*>(p); }
template <typename T> inline T __ldg(const T* p) { return *p; }

inline unsigned atomicAdd(unsigned* p, unsigned v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline int atomicAdd(int* p, int v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicAdd(unsigned long long* p, unsigned long long v) { return __atomic_fetch_add(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicExch(unsigned* p, unsigned v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned long long atomicExch(unsigned long long* p, unsigned long long v) { return __atomic_exchange_n(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicOr(unsigned* p, unsigned v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline unsigned atomicMin(unsigned* p, unsigned v) {
    unsigned old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMin(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old > v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline int atomicMax(int* p, int v) {
    int old = __atomic_load_n(p, __ATOMIC_SEQ_CST);
    while (old < v && !__atomic_compare_exchange_n(p, &old, v, false, __ATOMIC_SEQ_CST, __ATOMIC_SEQ_CST)) {}
    return old;
}
inline unsigned long long atomicOr(unsigned long long* p, unsigned long long v) { return __atomic_fetch_or(p, v, __ATOMIC_SEQ_CST); }
inline int __popcll(unsigned long long v) { return __builtin_popcountll(v); }
inline int __ffsll(long long v) { return __builtin_ffsll(v); }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
