// The single-pass pipeline kernel itself (launch code: fused_kernel.cu).  In a header of its own so that the
// host-run check of the device code (tests/hostcheck/fused_emu_host.cpp) compiles the very same kernel.
#pragma once
#include "fused_common.cuh"

#define FUSED_WARPS 8
#define FUSED_UNROLL 6
#define FUSED_SEGS_PER_ITER (32 * FUSED_UNROLL)      // 192 segments = 3 KB per warp iteration

// A streamed 128-bit word holds a byte above the threshold (rare).  1 channel: the word IS a 16-pixel segment -- append
// its mask.  3 channels (H x W x 3 interleaved, the layout _find_dot receives, helpers.py:143-145): a 16-pixel segment is
// three consecutive words, grey = (c0*9798 + c1*19235 + c2*3735 + 16384) >> 15 <= max(c0, c1, c2), so a segment can
// only hold a pixel above the threshold if one of its three words passes the packed byte test; the lane whose word
// is the FIRST of its segment to pass re-reads the segment (cache resident) and does the per-pixel arithmetic, the
// lanes of later passing words of the same segment stand down -- every segment is appended once.
template <bool USE_AND, bool CH3>
__device__ __forceinline__ void fused_hit(const FusedParams& P, int img, const uint4* __restrict__ src, int si, const uint4& v) {
    uint32_t m, seg;
    if (!CH3) {
        const uint32_t h0 = swar_gt(v.x, P.tc), h1 = swar_gt(v.y, P.tc);
        const uint32_t h2 = swar_gt(v.z, P.tc), h3 = swar_gt(v.w, P.tc);
        m = nibble_of(h0) | (nibble_of(h1) << 4) | (nibble_of(h2) << 8) | (nibble_of(h3) << 12);
        seg = (uint32_t)si;
    } else {
        seg = (uint32_t)si / 3u;
        const int r = si - 3 * (int)seg;
        uint32_t w[12];
        bool earlier = false;
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            const uint4 x = __ldg(src + 3 * seg + q);
            if (q < r && any_above<USE_AND>(x, P.tc)) earlier = true;
            w[4 * q + 0] = x.x; w[4 * q + 1] = x.y; w[4 * q + 2] = x.z; w[4 * q + 3] = x.w;
        }
        if (earlier) return;
        m = 0;
#pragma unroll
        for (int p = 0; p < 16; ++p) {
            const int o = 3 * p;
            const uint32_t c0 = (w[o >> 2] >> ((o & 3) * 8)) & 0xffu;
            const uint32_t c1 = (w[(o + 1) >> 2] >> (((o + 1) & 3) * 8)) & 0xffu;
            const uint32_t c2 = (w[(o + 2) >> 2] >> (((o + 2) & 3) * 8)) & 0xffu;
            const int grey = (int)((c0 * 9798u + c1 * 19235u + c2 * 3735u + 16384u) >> 15);      // cv2's 8-bit RGB2GRAY
            m |= (grey > P.threshold ? 1u : 0u) << p;
        }
        if (!m) return;
    }
    const uint32_t slot = atomicAdd(&P.seg_count[img], 1u);
    if (slot < (uint32_t)P.E) P.seg_list[(size_t)img * P.E + slot] = (seg << 16) | m;
}

template <bool WIDE, bool USE_AND, bool CH3 = false>
__global__ void __launch_bounds__(FUSED_WARPS * 32, 4)
k_pipeline_fused(const FusedParams P) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = smem_raw + P.slab_bytes * warp;

    while (true) {
        unsigned long long u = 0;
        if (lane == 0) u = atomicAdd(P.unit_counter, 1ull);
        u = __shfl_sync(0xffffffffu, u, 0);
        if (u >= (unsigned long long)P.total_units) break;
        const int img = (int)(u / P.units_per_image);
        const int unit = (int)(u - (unsigned long long)img * P.units_per_image);

        // ---- stream this slice of the image -----------------------------------------------------
        const int s_begin = unit * P.iters_per_unit * FUSED_SEGS_PER_ITER;
        const int s_end = min(P.u4_per_image, s_begin + P.iters_per_unit * FUSED_SEGS_PER_ITER);
        const uint4* src = P.frames + (size_t)img * P.u4_per_image;
        // rolling window: every lane keeps FUSED_UNROLL 128-bit loads in flight at all times -- an
        // element is consumed and its register immediately re-armed with the load of the next
        // iteration, so the warp never drains its memory pipeline between iterations.  Whole
        // iterations run without bounds checks; a ragged end (seg_per_image % 256 != 0) is handled after.
        const int n_full = (s_end - s_begin) / FUSED_SEGS_PER_ITER;
        if (n_full > 0) {
            const uint4* sp = src + s_begin + lane;
            uint4 v[FUSED_UNROLL];
#pragma unroll
            for (int q = 0; q < FUSED_UNROLL; ++q) v[q] = ldg_stream(sp + q * 32);
            for (int it = 0; it < n_full; ++it) {
                const bool more = it + 1 < n_full;               // warp-uniform
                const uint4* nx = sp + (it + 1) * FUSED_SEGS_PER_ITER;
#pragma unroll
                for (int q = 0; q < FUSED_UNROLL; ++q) {
                    if (any_above<USE_AND>(v[q], P.tc))           // rare: a marker crosses these 16 bytes
                        fused_hit<USE_AND, CH3>(P, img, src, s_begin + it * FUSED_SEGS_PER_ITER + q * 32 + lane, v[q]);
                    if (more) v[q] = ldg_stream(nx + q * 32);     // re-arm this slot at once
                }
            }
        }
        for (int si = s_begin + n_full * FUSED_SEGS_PER_ITER + lane; si < s_end; si += 32) {      // ragged end
            const uint4 x = ldg_stream(src + si);
            if (any_above<USE_AND>(x, P.tc)) fused_hit<USE_AND, CH3>(P, img, src, si, x);
        }
        __threadfence();                                   // release: this warp's list entries
        __syncwarp();
        unsigned done = 0;
        if (lane == 0) {
            __threadfence();                               // the publishing lane's fence comes AFTER the warp barrier: the other
            done = atomicAdd(&P.img_done[img], 1u);        // lanes' writes are then ordered before the counter (PTX memory model)
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done != (unsigned)P.units_per_image - 1) continue;

        finish_image<WIDE>(P, slab, img, lane);
    }
}

