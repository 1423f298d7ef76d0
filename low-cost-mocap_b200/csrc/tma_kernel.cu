// S1 + S2 + S3 in one persistent kernel, frames streamed by the bulk-copy engine (TMA, 1-D
// cp.async.bulk global -> shared with mbarrier completion) instead of by per-thread loads.
//
// Why: in k_pipeline_fused a warp that goes off to reduce an image or match a frame-set has no loads in
// flight while it does so; when the sparse stages are heavy (8 cameras x 16 markers) too few warps are
// streaming at any time to keep HBM busy.  Here the bytes in flight belong to the CTA, not to a warp:
//   * a ring of TMA_NS stages of TMA_CH bytes per CTA, each filled by ONE cp.async.bulk that signals the
//     stage's mbarrier (complete_tx); the data in flight costs no registers and no issue slots;
//   * any warp that has nothing else to do takes the next ticket, waits on that stage's mbarrier, thresholds
//     the 192 segments from shared memory (conflict-free 128-bit loads), and at once re-arms the stage with
//     the CTA's next chunk -- the consumer of a stage is the producer of its next generation, so there is no
//     dedicated producer warp and no empty-barrier;
//   * a warp that completes an image / a frame-set does the follow-up work (finish_image) while the ring
//     keeps turning under the other warps.
// Measured (profiles/README.md): correct (bit-identical to k_pipeline_fused, tests/test_parity_gpu.py) but
// 2.8x slower on the 4-camera workload (5.4 ms vs 1.94 ms per 10 000 frame-sets): the per-chunk bookkeeping
// (ticket, hand-out lock, completion atomic with its round trip to L2) costs more than the loads it
// replaces, and the plain-load kernel already runs at 0.985 of the measured HBM peak.  Kept as the
// MOCAP_PIPELINE=tma variant, not the default.
// Hazards: ticket t maps to stage t % NS, generation t / NS.  A consumer first spins on gen[stage] == its
// generation (published by the refiller after arming the barrier), so it can never look at a barrier more
// than one phase ahead (the mbarrier parity test only distinguishes adjacent phases).
#include "fused_common.cuh"

#define TMA_WARPS 8
#define TMA_NS 12
#define TMA_CH_SEGS 128                         // 16-pixel segments per chunk
#define TMA_CH (TMA_CH_SEGS * 16)                // 2 KB
#define TMA_CHUNKS_PER_UNIT 32                   // chunks a CTA claims from the global counter at a time

struct TmaRing {
    unsigned long long full[TMA_NS];             // mbarriers
    unsigned gen[TMA_NS];                        // generation the stage currently holds (or is being filled for)
    int chunk_img[TMA_NS];                       // image of the chunk in the stage, -1: no more work
    int chunk_off[TMA_NS];                       // first segment of the chunk within its image
    int chunk_len[TMA_NS];                       // segments in the chunk
    unsigned ticket;
    unsigned lock;
    long long cur, end;                          // CTA-local range of global chunk ids still to hand out
    int exhausted;
    int valid_in_ring;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" :: "r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(unsigned long long* bar, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                 : "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :: "r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// hand out the CTA's next chunk (called by ONE lane).  false: the whole batch has been handed out.
__device__ __forceinline__ bool next_chunk(const FusedParams& P, TmaRing* R, long long total_chunks, int chunks_per_image,
                                           int& img, int& off, int& len) {
    while (atomicCAS(&R->lock, 0u, 1u) != 0u) { }
    __threadfence_block();
    bool ok = false;
    long long g = 0;
    if (R->cur >= R->end && !R->exhausted) {
        const unsigned long long u = atomicAdd(P.unit_counter, 1ull);
        const long long b = (long long)u * TMA_CHUNKS_PER_UNIT;
        if (b < total_chunks) { R->cur = b; R->end = min(b + TMA_CHUNKS_PER_UNIT, total_chunks); }
        else R->exhausted = 1;
    }
    if (R->cur < R->end) { g = R->cur; R->cur = g + 1; atomicAdd(&R->valid_in_ring, 1); ok = true; }   // atomic: consumers decrement outside the lock
    __threadfence_block();
    atomicExch(&R->lock, 0u);
    if (ok) {
        img = (int)(g / chunks_per_image);
        const int idx = (int)(g - (long long)img * chunks_per_image);
        off = idx * TMA_CH_SEGS;
        len = min(TMA_CH_SEGS, P.seg_per_image - off);
    }
    return ok;
}

// (re)fill stage s for generation g: arm its barrier and start the bulk copy, or mark it empty
__device__ __forceinline__ void fill_stage(const FusedParams& P, TmaRing* R, unsigned char* ring, int s, unsigned g,
                                           long long total_chunks, int chunks_per_image) {
    int img = -1, off = 0, len = 0;
    const bool ok = next_chunk(P, R, total_chunks, chunks_per_image, img, off, len);
    R->chunk_img[s] = ok ? img : -1;
    R->chunk_off[s] = off;
    R->chunk_len[s] = len;
    if (ok) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // earlier generic reads of the stage vs the async write
        mbar_expect_tx(&R->full[s], (unsigned)len * 16u);
        bulk_g2s(ring + (size_t)s * TMA_CH, P.frames + (size_t)img * P.seg_per_image + off, (unsigned)len * 16u, &R->full[s]);
    } else {
        mbar_arrive(&R->full[s]);                                           // completes the phase without data
    }
    __threadfence_block();
    *reinterpret_cast<volatile unsigned*>(&R->gen[s]) = g;                  // publish: stage now belongs to generation g
}

template <bool WIDE, bool USE_AND>
__global__ void __launch_bounds__(TMA_WARPS * 32, 3)
k_pipeline_tma(const FusedParams P, long long total_chunks, int chunks_per_image) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                         // TMA_NS * TMA_CH, 128-byte aligned
    TmaRing* R = reinterpret_cast<TmaRing*>(smem_raw + (size_t)TMA_NS * TMA_CH);
    unsigned char* slabs = smem_raw + (size_t)TMA_NS * TMA_CH + ((sizeof(TmaRing) + 127) & ~(size_t)127);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    unsigned char* slab = slabs + P.slab_bytes * warp;

    if (threadIdx.x == 0) {
        for (int s = 0; s < TMA_NS; ++s) { mbar_init(&R->full[s], 1); R->gen[s] = 0xffffffffu; }
        R->ticket = 0; R->lock = 0; R->cur = 0; R->end = 0; R->exhausted = 0; R->valid_in_ring = 0;
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (int s = 0; s < TMA_NS; ++s) fill_stage(P, R, ring, s, 0u, total_chunks, chunks_per_image);
    __syncthreads();

    while (true) {
        unsigned t = 0;
        if (lane == 0) t = atomicAdd(&R->ticket, 1u);
        t = __shfl_sync(0xffffffffu, t, 0);
        const int s = (int)(t % TMA_NS);
        const unsigned g = t / TMA_NS;
        // the stage must have been re-armed for my generation before its barrier may be examined
        while (*reinterpret_cast<volatile unsigned*>(&R->gen[s]) != g) { }
        while (!mbar_try_wait(&R->full[s], g & 1u)) { }
        const int img = *reinterpret_cast<volatile int*>(&R->chunk_img[s]);
        if (img < 0) {                                                      // an empty stage: pass it on, maybe leave
            __syncwarp();
            int leave = 0;
            if (lane == 0) {
                fill_stage(P, R, ring, s, g + 1u, total_chunks, chunks_per_image);
                leave = (*reinterpret_cast<volatile int*>(&R->exhausted) != 0 &&
                         *reinterpret_cast<volatile int*>(&R->valid_in_ring) == 0) ? 1 : 0;
            }
            leave = __shfl_sync(0xffffffffu, leave, 0);
            if (leave) break;
            continue;
        }
        const int off = *reinterpret_cast<volatile int*>(&R->chunk_off[s]);
        const int len = *reinterpret_cast<volatile int*>(&R->chunk_len[s]);
        const uint4* st = reinterpret_cast<const uint4*>(ring + (size_t)s * TMA_CH);
#pragma unroll
        for (int q = 0; q < TMA_CH_SEGS / 32; ++q) {
            const int k = q * 32 + lane;
            if (k < len) {
                const uint4 x = st[k];
                if (any_above<USE_AND>(x, P.tc)) {
                    const uint32_t h0 = swar_gt(x.x, P.tc), h1 = swar_gt(x.y, P.tc);
                    const uint32_t h2 = swar_gt(x.z, P.tc), h3 = swar_gt(x.w, P.tc);
                    const uint32_t m = nibble_of(h0) | (nibble_of(h1) << 4) | (nibble_of(h2) << 8) | (nibble_of(h3) << 12);
                    const uint32_t slot = atomicAdd(&P.seg_count[img], 1u);
                    if (slot < (uint32_t)P.E) P.seg_list[(size_t)img * P.E + slot] = ((uint32_t)(off + k) << 16) | m;
                }
            }
        }
        __threadfence();                                                    // release: this warp's list entries
        __syncwarp();                                                       // every lane is done reading the stage
        unsigned done = 0;
        if (lane == 0) {
            atomicSub(&R->valid_in_ring, 1);
            fill_stage(P, R, ring, s, g + 1u, total_chunks, chunks_per_image);    // the ring keeps turning
            done = atomicAdd(&P.img_done[img], 1u);
        }
        done = __shfl_sync(0xffffffffu, done, 0);
        if (done != (unsigned)chunks_per_image - 1) continue;
        finish_image<WIDE>(P, slab, img, lane);
    }
}

size_t fused_slab_bytes(const mocap_config& c);

static size_t tma_smem_bytes(const mocap_config& c) {
    return (size_t)TMA_NS * TMA_CH + ((sizeof(TmaRing) + 127) & ~(size_t)127) + fused_slab_bytes(c) * TMA_WARPS;
}

int launch_pipeline_tma(mocap_ctx* ctx, const uint8_t* frames, int n_sets, int threshold,
                        double* obj, double* err, int32_t* n_obj, int32_t* set_flags) {
    if (n_sets <= 0) return MOCAP_OK;
    const mocap_config& c = ctx->cfg;
    FusedParams P;
    memset(&P, 0, sizeof(P));
    P.frames = reinterpret_cast<const uint4*>(frames);
    P.n_sets = n_sets; P.C = c.n_cam; P.W = c.width; P.H = c.height;
    P.seg_per_image = c.width * c.height / MOCAP_SEG_PX;
    const int chunks_per_image = (P.seg_per_image + TMA_CH_SEGS - 1) / TMA_CH_SEGS;
    const long long total_chunks = (long long)n_sets * c.n_cam * chunks_per_image;
    if (threshold < 0) { P.tc.addc = 0x80808080u; P.tc.use_and = 0; }
    else if (threshold >= 255) { P.tc.addc = 0; P.tc.use_and = 1; }
    else {
        const uint32_t T1 = (uint32_t)threshold + 1u;
        P.tc.use_and = T1 > 128 ? 1u : 0u;
        P.tc.addc = (T1 > 128 ? 256u - T1 : 128u - T1) * 0x01010101u;
    }
    P.E = c.max_segments;
    P.seg_count = ctx->d_seg_count; P.seg_list = ctx->d_seg_list;
    P.img_done = ctx->d_img_done; P.set_done = ctx->d_set_done; P.set_defer = ctx->d_set_done + ctx->cap_images;
    P.unit_counter = ctx->d_unit_counter;
    P.blob_xy = ctx->d_blob_xy; P.blob_n = ctx->d_blob_n; P.img_flags = ctx->d_img_flags;
    P.img_worklist = ctx->d_worklist; P.img_work_count = ctx->d_work_count;
    P.set_worklist = ctx->d_set_worklist; P.set_work_count = ctx->d_work_count + 2;
    P.tb = ctx->d_tables;
    P.MB = c.max_blobs; P.RMAX = c.max_roots; P.KC = c.max_cands; P.GMAX = (uint32_t)c.max_groups;
    P.obj = obj; P.err = err; P.n_obj = n_obj; P.set_flags = set_flags;
    P.track_xy = ctx->track_xy_cur;
    P.slab_bytes = fused_slab_bytes(c);
    const size_t smem = tma_smem_bytes(c);
    const long long mx = c.width > c.height ? c.width : c.height;
    const bool wide = 6ll * mx * c.width * c.height >= (1ll << 32);

    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_unit_counter, 0, sizeof(unsigned long long), ctx->stream));
    if (ctx->timing_on) {
        if (ctx->tim_used == 64) { const int st = timing_flush(ctx); if (st) return st; }
        CUDA_TRY(ctx, cudaEventRecord(ctx->tim_ev[2 * ctx->tim_used], ctx->stream));
    }
    const int grid = ctx->num_sms * ctx->tma_ctas_per_sm;
    if (wide) {
        if (P.tc.use_and) k_pipeline_tma<true, true><<<grid, TMA_WARPS * 32, smem, ctx->stream>>>(P, total_chunks, chunks_per_image);
        else k_pipeline_tma<true, false><<<grid, TMA_WARPS * 32, smem, ctx->stream>>>(P, total_chunks, chunks_per_image);
    } else {
        if (P.tc.use_and) k_pipeline_tma<false, true><<<grid, TMA_WARPS * 32, smem, ctx->stream>>>(P, total_chunks, chunks_per_image);
        else k_pipeline_tma<false, false><<<grid, TMA_WARPS * 32, smem, ctx->stream>>>(P, total_chunks, chunks_per_image);
    }
    CUDA_TRY(ctx, cudaGetLastError());
    if (ctx->timing_on) {
        CUDA_TRY(ctx, cudaEventRecord(ctx->tim_ev[2 * ctx->tim_used + 1], ctx->stream));
        ctx->tim_used += 1;
    }
    ctx->launches += 1;
    int st = launch_blob_fallback(ctx, ctx->d_blob_xy, ctx->d_blob_n, nullptr, ctx->d_img_flags, n_sets * ctx->cfg.n_cam);
    if (st) return st;
    return launch_match_list(ctx, ctx->d_blob_xy, ctx->d_blob_n, ctx->d_set_worklist, ctx->d_work_count + 2, n_sets,
                             obj, err, n_obj, set_flags);
}

int tma_kernel_init(mocap_ctx* ctx) {
    const size_t smem = tma_smem_bytes(ctx->cfg);
    ctx->tma_ctas_per_sm = 0;
    if (smem > 220 * 1024) return MOCAP_OK;
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pipeline_tma<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pipeline_tma<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pipeline_tma<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    CUDA_TRY(ctx, cudaFuncSetAttribute(k_pipeline_tma<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int per_sm = 0;
    CUDA_TRY(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pipeline_tma<false, false>, TMA_WARPS * 32, smem));
    ctx->tma_ctas_per_sm = per_sm;
    return MOCAP_OK;
}
