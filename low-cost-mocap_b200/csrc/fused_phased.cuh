// Phase-synchronous variant of the single-pass pipeline kernel (MOCAP_PIPELINE=phased, opt-in).
//
// k_pipeline_fused lets the warp that finishes an image reduce it at once and the warp that finishes a
// frame-set match it at once.  With heavy frame-sets (8 cameras x 16 markers) that puts three large code
// regions -- stream loop, blob reduce, matcher + DLT, 130 KB of SASS against a 32 KB instruction cache -- on
// every SM at every moment, one warp each, and 35 % of all warp samples wait for instructions
// (profiles/ncu_full_r01_c8m16.csv).  Here the follow-up work is QUEUED per CTA and the CTA changes phase as a
// whole: all warps stream until 8 finished images (or 8 finished frame-sets) are waiting, then all warps
// reduce images, then all warps match frame-sets, so the warps of a CTA fetch the same code at the same time.
//
//   phase A  every warp claims units and streams them exactly as k_pipeline_fused does; the warp that completes
//            the last slice of an image appends the image to the CTA's image queue
//   phase B  (after a CTA barrier) warp w reduces queued images w, w + 8, ...; an image that completes its
//            frame-set appends the set to the CTA's set queue (or to the global worklist if an image of the set
//            was deferred)
//   phase C  (after a CTA barrier, once 8 sets wait or nothing is left to stream) warp w matches sets w, w + 8, ...
// Every image is queued by exactly one warp of exactly one CTA and every CTA drains its own queues before it
// exits, so nothing is lost; the global counters, worklists and self-resetting scratch are those of the fused
// kernel (FusedParams), and so are the results.
//
// Status: logic checked on the host (tests/test_device_code_on_host.py runs this kernel under the SIMT
// emulation, incl. ThreadSanitizer); not yet measured on a GPU, hence not the default.
#pragma once
#include "fused_device.cuh"

#define PHASED_QCAP 32        // queue capacity: trigger (8) + one push per warp (8) + carry-over, with room
#define PHASED_TRIGGER 8

struct PhasedQueues {
    uint32_t img[PHASED_QCAP];
    uint32_t set[PHASED_QCAP];
    uint32_t n_img, n_set;
    uint32_t exhausted;       // some warp of this CTA found the unit counter past the end
};

// a 32-bit word other warps update with atomics, polled without ordering (a volatile load on the device; the
// host-run checks map it to a relaxed atomic load, which is what it means)
__device__ __forceinline__ uint32_t phased_peek(const uint32_t* p) {
#if defined(__CUDACC__)
    return *reinterpret_cast<const volatile uint32_t*>(p);
#else
    return __atomic_load_n(p, __ATOMIC_RELAXED);
#endif
}

// phase B for one image: segment list -> blobs; returns the frame-set to match if this was its last image
// (and none of its images was deferred), else -1.  Mirrors the first half of finish_image.
template <bool WIDE>
__device__ __forceinline__ int phased_reduce_image(const FusedParams& P, unsigned char* slab, int img, int lane) {
    __threadfence();                                   // acquire: every warp's list entries of this image
    const unsigned cnt = __ldcg(&P.seg_count[img]);
    bool deferred = false;
    if (cnt == 0) {
        if (lane == 0) { P.blob_n[img] = 0; if (P.img_flags) P.img_flags[img] = 0; }
    } else {
        bool ok = cnt <= BLOB_WE;
        if (ok) {
            WarpSlab& sl = *reinterpret_cast<WarpSlab*>(slab);
            BlobSmem sm;
            sm.seg = sl.seg; sm.parent = sl.parent; sm.base = sl.base; sm.node_seg = sl.node_seg;
            sm.node_bits = sl.node_bits; sm.rank = sl.rank; sm.acc = sl.acc; sm.wsum = nullptr; sm.hs = nullptr;
            sm.rowfirst = BLOB_ROWFIRST(sl, WIDE); sm.row_cap = WIDE ? 0 : BLOB_ROWS;
            const uint32_t* lst = P.seg_list + (size_t)img * P.E;
            for (int i = lane; i < (int)cnt; i += 32) sm.seg[i] = __ldcg(lst + i);
            __syncwarp();
            ok = blob_reduce<32, true, WIDE>(sm, (int)cnt, BLOB_WE, BLOB_WACC, P.W, P.H, P.MB,
                                             P.blob_xy + (size_t)img * P.MB * 2, P.blob_n + img, nullptr,
                                             P.img_flags ? P.img_flags + img : nullptr, 0);
            __syncwarp();
        }
        if (lane == 0) {
            if (ok) P.seg_count[img] = 0;             // self-cleaning
            else P.img_worklist[atomicAdd(P.img_work_count, 1u)] = (uint32_t)img;
        }
        deferred = !ok;
    }
    if (lane == 0) P.img_done[img] = 0;
    const int set = img / P.C;
    if (deferred && lane == 0) atomicOr(&P.set_defer[set], 1u);
    __threadfence();                                   // release: blob list of this image
    __syncwarp();
    unsigned sd = 0;
    if (lane == 0) {
        __threadfence();                               // publishing lane: fence after the warp barrier
        sd = atomicAdd(&P.set_done[set], 1u);
    }
    sd = __shfl_sync(0xffffffffu, sd, 0);
    if (sd != (unsigned)P.C - 1) return -1;
    __threadfence();                                   // acquire: blob lists of the other cameras
    unsigned defer = 0;
    if (lane == 0) {
        defer = __ldcg(&P.set_defer[set]);
        P.set_done[set] = 0; P.set_defer[set] = 0;
        if (defer) P.set_worklist[atomicAdd(P.set_work_count, 1u)] = (uint32_t)set;
    }
    defer = __shfl_sync(0xffffffffu, defer, 0);
    return defer ? -1 : set;
}

template <bool WIDE, bool USE_AND>
__global__ void __launch_bounds__(FUSED_WARPS * 32, 4)
k_pipeline_phased(const FusedParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    PhasedQueues& q = *reinterpret_cast<PhasedQueues*>(smem_raw + P.slab_bytes * FUSED_WARPS);     // behind the warps' slabs
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem_raw + P.slab_bytes * warp;
    if (threadIdx.x == 0) { q.n_img = 0; q.n_set = 0; q.exhausted = 0; }
    __syncthreads();

    while (true) {
        // ---- phase A: stream until enough follow-up work waits (or nothing is left to stream) --------------
        while (true) {
            unsigned stop = 0;
            if (lane == 0)
                stop = (phased_peek(&q.n_img) >= PHASED_TRIGGER || phased_peek(&q.n_set) >= PHASED_TRIGGER || phased_peek(&q.exhausted)) ? 1u : 0u;
            stop = __shfl_sync(0xffffffffu, stop, 0);
            if (stop) break;
            unsigned long long u = 0;
            if (lane == 0) u = atomicAdd(P.unit_counter, 1ull);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (u >= (unsigned long long)P.total_units) {
                if (lane == 0) atomicExch(&q.exhausted, 1u);
                break;
            }
            const int img = (int)(u / P.units_per_image);
            const int unit = (int)(u - (unsigned long long)img * P.units_per_image);
            const int s_begin = unit * P.iters_per_unit * FUSED_SEGS_PER_ITER;
            const int s_end = min(P.seg_per_image, s_begin + P.iters_per_unit * FUSED_SEGS_PER_ITER);
            const uint4* src = P.frames + (size_t)img * P.seg_per_image;
            const int n_full = (s_end - s_begin) / FUSED_SEGS_PER_ITER;
            if (n_full > 0) {                            // rolling window of FUSED_UNROLL loads per lane (see k_pipeline_fused)
                const uint4* sp = src + s_begin + lane;
                uint4 v[FUSED_UNROLL];
#pragma unroll
                for (int k = 0; k < FUSED_UNROLL; ++k) v[k] = ldg_stream(sp + k * 32);
                for (int it = 0; it < n_full; ++it) {
                    const bool more = it + 1 < n_full;
                    const uint4* nx = sp + (it + 1) * FUSED_SEGS_PER_ITER;
#pragma unroll
                    for (int k = 0; k < FUSED_UNROLL; ++k) {
                        if (any_above<USE_AND>(v[k], P.tc)) {
                            const uint32_t h0 = swar_gt(v[k].x, P.tc), h1 = swar_gt(v[k].y, P.tc);
                            const uint32_t h2 = swar_gt(v[k].z, P.tc), h3 = swar_gt(v[k].w, P.tc);
                            const int si = s_begin + it * FUSED_SEGS_PER_ITER + k * 32 + lane;
                            const uint32_t m = nibble_of(h0) | (nibble_of(h1) << 4) | (nibble_of(h2) << 8) | (nibble_of(h3) << 12);
                            const uint32_t slot = atomicAdd(&P.seg_count[img], 1u);
                            if (slot < (uint32_t)P.E) P.seg_list[(size_t)img * P.E + slot] = ((uint32_t)si << 16) | m;
                        }
                        if (more) v[k] = ldg_stream(nx + k * 32);
                    }
                }
            }
            for (int si = s_begin + n_full * FUSED_SEGS_PER_ITER + lane; si < s_end; si += 32) {      // ragged end
                const uint4 x = ldg_stream(src + si);
                if (!any_above<USE_AND>(x, P.tc)) continue;
                const uint32_t h0 = swar_gt(x.x, P.tc), h1 = swar_gt(x.y, P.tc);
                const uint32_t h2 = swar_gt(x.z, P.tc), h3 = swar_gt(x.w, P.tc);
                const uint32_t m = nibble_of(h0) | (nibble_of(h1) << 4) | (nibble_of(h2) << 8) | (nibble_of(h3) << 12);
                const uint32_t slot = atomicAdd(&P.seg_count[img], 1u);
                if (slot < (uint32_t)P.E) P.seg_list[(size_t)img * P.E + slot] = ((uint32_t)si << 16) | m;
            }
            __threadfence();                               // release: this warp's list entries
            __syncwarp();
            if (lane == 0) {
                __threadfence();                           // publishing lane: fence after the warp barrier
                const unsigned done = atomicAdd(&P.img_done[img], 1u);
                if (done == (unsigned)P.units_per_image - 1) q.img[atomicAdd(&q.n_img, 1u)] = (uint32_t)img;
            }
        }
        __syncthreads();                                   // queues are complete, nobody streams

        // ---- phase B: all warps reduce the queued images --------------------------------------------------
        const unsigned ni = q.n_img;
        for (unsigned k = warp; k < ni; k += FUSED_WARPS) {
            const int set = phased_reduce_image<WIDE>(P, slab, (int)q.img[k], lane);
            if (set >= 0 && lane == 0) q.set[atomicAdd(&q.n_set, 1u)] = (uint32_t)set;
            __syncwarp();
        }
        __syncthreads();                                   // the set queue is complete
        const unsigned ns = q.n_set;
        const bool exhausted = q.exhausted != 0;
        const bool match_now = ns >= PHASED_TRIGGER || exhausted;

        // ---- phase C: all warps match the queued frame-sets ------------------------------------------------
        if (match_now) {
            for (unsigned k = warp; k < ns; k += FUSED_WARPS) {
                const int set = (int)q.set[k];
                WarpState ws = carve_warp_state(slab, P.RMAX, P.C, P.KC, P.MB);
                match_triangulate_warp(P.tb, ws, P.blob_xy + (size_t)set * P.C * P.MB * 2, P.blob_n + (size_t)set * P.C, set, lane,
                                       P.C, P.MB, P.RMAX, P.KC, P.GMAX, P.obj, P.err, P.n_obj, P.set_flags, nullptr, P.track_xy,
                                       P.img_flags ? P.img_flags + (size_t)set * P.C : nullptr);
                __syncwarp();
            }
        }
        __syncthreads();                                   // everybody has read the queues
        if (threadIdx.x == 0) { q.n_img = 0; if (match_now) q.n_set = 0; }
        __syncthreads();
        if (exhausted) break;                              // nothing left to stream and both queues were drained
    }
}
