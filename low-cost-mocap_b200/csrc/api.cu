// C ABI of libmocap_b200.so (see include/mocap_b200.h for the contract and the
// reference functions each entry point replaces).
#include <stdarg.h>
#include <stdlib.h>
#include <math.h>
#include <new>
#include "common.cuh"
#include "camera_tables.h"

int blob_kernels_init(mocap_ctx* ctx);
int match_kernels_init(mocap_ctx* ctx);
int fused_kernel_init(mocap_ctx* ctx);
int tma_kernel_init(mocap_ctx* ctx);
int ba_dev_init(mocap_ctx* ctx);

int mocap_fail(mocap_ctx* ctx, int code, const char* fmt, ...) {
    if (ctx) {
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(ctx->err, sizeof(ctx->err), fmt, ap);
        va_end(ap);
    }
    return code;
}

int ensure_scratch(mocap_ctx* ctx, size_t bytes) {
    if (bytes <= ctx->scratch_bytes) return MOCAP_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_scratch);
    ctx->d_scratch = nullptr; ctx->scratch_bytes = 0;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_scratch, bytes));
    ctx->scratch_bytes = bytes;
    return MOCAP_OK;
}


extern "C" {

const char* mocap_status_string(int status) {
    switch (status) {
        case MOCAP_OK: return "ok";
        case MOCAP_EINVAL: return "invalid argument";
        case MOCAP_ENODEV: return "no usable CUDA device (libmocap_b200 needs an sm_100 GPU; there is no CPU fallback)";
        case MOCAP_ECUDA: return "CUDA runtime error";
        case MOCAP_ENOMEM: return "out of memory";
        case MOCAP_ESTATE: return "call order error (cameras not set?)";
        default: return "unknown status";
    }
}

void mocap_default_config(mocap_config* cfg, int n_cam, int width, int height) {
    cfg->device = 0;
    cfg->n_cam = n_cam;
    cfg->width = width;
    cfg->height = height;
    cfg->max_blobs = 32;
    cfg->max_segments = 1024;
    cfg->max_roots = 64;
    cfg->max_cands = 8;
    cfg->max_groups = 4096;
}

const char* mocap_last_error(const mocap_ctx* ctx) { return ctx ? ctx->err : "no context"; }

static bool is_pow2(int v) { return v > 0 && (v & (v - 1)) == 0; }

int mocap_create(mocap_ctx** out, const mocap_config* cfg) {
    if (!out || !cfg) return MOCAP_EINVAL;
    *out = nullptr;
    if (cfg->n_cam < 1 || cfg->n_cam > MOCAP_MAX_CAM || cfg->width < 16 || cfg->width % MOCAP_SEG_PX != 0 ||
        cfg->height < 1 || cfg->max_blobs < 1 || cfg->max_blobs > MOCAP_MAX_BLOBS ||
        !is_pow2(cfg->max_segments) || cfg->max_segments < 64 || cfg->max_segments > 4096 ||
        cfg->max_roots < 1 || cfg->max_roots > MOCAP_MAX_ROOTS || cfg->max_cands < 1 || cfg->max_cands > MOCAP_MAX_CANDS ||
        cfg->max_groups < 1 || (long long)cfg->width * cfg->height / MOCAP_SEG_PX > 65535)
        return MOCAP_EINVAL;
    int n_dev = 0;
    if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev <= 0 || cfg->device < 0 || cfg->device >= n_dev) {
        cudaGetLastError();
        return MOCAP_ENODEV;
    }
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess) return MOCAP_ENODEV;
    if (prop.major != 10) return MOCAP_ENODEV;        // the fatbin holds sm_100a code only
    if (cudaSetDevice(cfg->device) != cudaSuccess) return MOCAP_ENODEV;

    mocap_ctx* ctx = new (std::nothrow) mocap_ctx();
    if (!ctx) return MOCAP_ENOMEM;
    memset(ctx, 0, sizeof(*ctx));
    ctx->cfg = *cfg;
    ctx->num_sms = prop.multiProcessorCount;
    ctx->h_tables.n_cam = cfg->n_cam;
    int st = MOCAP_OK;
    do {
        if (cudaMalloc(&ctx->d_tables, sizeof(CameraTables)) != cudaSuccess) { st = MOCAP_ENOMEM; break; }
        if (cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking) != cudaSuccess) { st = MOCAP_ECUDA; break; }
        if (cudaStreamCreateWithFlags(&ctx->copy_stream2, cudaStreamNonBlocking) != cudaSuccess) { st = MOCAP_ECUDA; break; }
        for (int k = 0; k < 2 * 64; ++k)
            if (cudaEventCreate(&ctx->tim_ev[k]) != cudaSuccess) { st = MOCAP_ECUDA; break; }
        if (st) break;
        for (int k = 0; k < 2; ++k) {
            if (cudaEventCreateWithFlags(&ctx->stage_free[k], cudaEventDisableTiming) != cudaSuccess) { st = MOCAP_ECUDA; break; }
            if (cudaEventCreateWithFlags(&ctx->copied[k], cudaEventDisableTiming) != cudaSuccess) { st = MOCAP_ECUDA; break; }
        }
        if (st) break;
        if ((st = blob_kernels_init(ctx)) != MOCAP_OK) break;
        if ((st = match_kernels_init(ctx)) != MOCAP_OK) break;
        {
            // MOCAP_PIPELINE = split | fused | phased | tma pins the pipeline (A/B measurements); unset: single-pass kernel,
            // except for batches of heavy frame-sets (see pick_fused)
            const char* mode = getenv("MOCAP_PIPELINE");
            ctx->use_fused = (mode && strcmp(mode, "split") == 0) ? 0 : 1;
            ctx->use_tma = (mode && strcmp(mode, "tma") == 0) ? 1 : 0;
            ctx->use_phased = (mode && strcmp(mode, "phased") == 0) ? 1 : 0;
            ctx->pipeline_auto = (mode && mode[0]) ? 0 : 1;
        }
        {
            void* h = nullptr;
            if (cudaHostAlloc(&h, 2 * sizeof(unsigned long long), cudaHostAllocMapped) != cudaSuccess) { st = MOCAP_ENOMEM; break; }
            memset(h, 0, 2 * sizeof(unsigned long long));
            ctx->h_stat = static_cast<volatile unsigned long long*>(h);
            void* d = nullptr;
            if (cudaHostGetDevicePointer(&d, h, 0) != cudaSuccess) { st = MOCAP_ECUDA; break; }
            ctx->d_stat_host = static_cast<unsigned long long*>(d);
            if (cudaMalloc(&ctx->d_stat_acc, sizeof(unsigned long long)) != cudaSuccess) { st = MOCAP_ENOMEM; break; }
            if (cudaMemset(ctx->d_stat_acc, 0, sizeof(unsigned long long)) != cudaSuccess) { st = MOCAP_ECUDA; break; }
        }
        if ((st = fused_kernel_init(ctx)) != MOCAP_OK) break;
        if ((st = tma_kernel_init(ctx)) != MOCAP_OK) break;
        if ((st = ba_dev_init(ctx)) != MOCAP_OK) break;
        if (ctx->tma_ctas_per_sm < 1) ctx->use_tma = 0;
    } while (0);
    if (st != MOCAP_OK) { mocap_destroy(ctx); return st; }
    *out = ctx;
    return MOCAP_OK;
}

void mocap_destroy(mocap_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->cfg.device);
    cudaDeviceSynchronize();
    cudaFree(ctx->d_tables);
    cudaFree(ctx->d_seg_count); cudaFree(ctx->d_seg_list); cudaFree(ctx->d_worklist); cudaFree(ctx->d_work_count);
    cudaFree(ctx->d_set_worklist); cudaFree(ctx->d_img_done); cudaFree(ctx->d_set_done); cudaFree(ctx->d_unit_counter);
    cudaFree(ctx->d_blob_xy); cudaFree(ctx->d_blob_n); cudaFree(ctx->d_img_flags);
    cudaFree(ctx->d_stage[0]); cudaFree(ctx->d_stage[1]);
    cudaFree(ctx->d_obj); cudaFree(ctx->d_err); cudaFree(ctx->d_nobj); cudaFree(ctx->d_setflags);
    cudaFree(ctx->d_scratch);
    cudaFree(ctx->d_ba_ws);
    cudaFree(ctx->d_match_counter);
    cudaFree(ctx->d_match_items); cudaFree(ctx->d_match_partial); cudaFree(ctx->d_match_range); cudaFree(ctx->d_match_arrive);
    cudaFree(ctx->d_pp_m1); cudaFree(ctx->d_pp_m2); cudaFree(ctx->d_pp_rot);
    cudaFree(ctx->d_stat_acc);
    if (ctx->h_stat) cudaFreeHost(const_cast<unsigned long long*>(ctx->h_stat));
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->copy_stream2) cudaStreamDestroy(ctx->copy_stream2);
    for (int k = 0; k < 2 * 64; ++k) if (ctx->tim_ev[k]) cudaEventDestroy(ctx->tim_ev[k]);
    for (int k = 0; k < 2; ++k) if (ctx->stage_free[k]) cudaEventDestroy(ctx->stage_free[k]);
    for (int k = 0; k < 2; ++k) if (ctx->copied[k]) cudaEventDestroy(ctx->copied[k]);
    delete ctx;
}

int mocap_set_stream(mocap_ctx* ctx, void* cuda_stream) {
    if (!ctx) return MOCAP_EINVAL;
    ctx->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return MOCAP_OK;
}

// ---- host-side camera tables: camera_tables.h ---------------------------------------------
int mocap_set_cameras(mocap_ctx* ctx, const double* K, const double* R, const double* t) {
    if (!ctx || !K || !R || !t) return MOCAP_EINVAL;
    const int C = ctx->cfg.n_cam;
    CameraTables& T = ctx->h_tables;
    build_camera_tables(T, C, K, R, t);
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_tables, &T, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));      // T lives in pageable memory of the ctx
    ctx->cameras_set = true;
    return MOCAP_OK;
}

int mocap_set_world_transform(mocap_ctx* ctx, const double* M) {
    if (!ctx) return MOCAP_EINVAL;
    CameraTables& T = ctx->h_tables;
    T.use_world = M ? 1 : 0;
    if (M) memcpy(T.world, M, 16 * sizeof(double));
    CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_tables, &T, sizeof(T), cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MOCAP_OK;
}

// Single-pass kernel or three-kernel pipeline for this batch of 1-channel frames?  The single-pass kernel wins
// while the sparse stages are light (config 2: 1.93 ms against 2.2 ms per 10 000 frame-sets); with many blobs per
// frame-set the matcher's code is large and, interleaved with the stream loop on every SM, lives on instruction-
// cache misses (config 3 shape: 3.73 ms against 3.10 ms per 4000 frame-sets), so batches that follow a heavy
// batch take the three-kernel pipeline.  The blob count of the previous batch arrives through mapped host memory
// (written by the last CTA of the blob fallback kernel): no synchronisation, a stale or torn value only steers
// this heuristic, both pipelines give the same results.
static bool pick_fused(const mocap_ctx* ctx, int channels) {
    if (!ctx->use_fused) return false;
    if (channels != 1 && (ctx->use_tma || ctx->use_phased)) return false;     // the experimental variants stream 1-channel frames only
    if (!ctx->pipeline_auto) return true;
    const unsigned long long blobs = ctx->h_stat[0], images = ctx->h_stat[1];
    if (images == 0) return true;
    return (double)blobs * ctx->cfg.n_cam <= (double)MOCAP_HEAVY_BLOBS_PER_SET * (double)images;
}

// ---- scratch management ------------------------------------------------------------------
static int ensure_images(mocap_ctx* ctx, int n_images) {
    if (n_images <= ctx->cap_images) return MOCAP_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_seg_count); cudaFree(ctx->d_seg_list); cudaFree(ctx->d_worklist); cudaFree(ctx->d_work_count); cudaFree(ctx->d_set_worklist); cudaFree(ctx->d_img_done); cudaFree(ctx->d_set_done); cudaFree(ctx->d_unit_counter); cudaFree(ctx->d_blob_xy); cudaFree(ctx->d_blob_n); cudaFree(ctx->d_img_flags);
    ctx->d_seg_count = nullptr; ctx->d_seg_list = nullptr; ctx->d_worklist = nullptr; ctx->d_work_count = nullptr; ctx->d_set_worklist = nullptr; ctx->d_img_done = nullptr; ctx->d_set_done = nullptr; ctx->d_unit_counter = nullptr; ctx->d_blob_xy = nullptr; ctx->d_blob_n = nullptr; ctx->d_img_flags = nullptr;
    ctx->cap_images = 0;
    const size_t n = (size_t)n_images;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_seg_count, n * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_seg_list, n * ctx->cfg.max_segments * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_worklist, n * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_work_count, 4 * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_work_count, 0, 4 * sizeof(uint32_t), ctx->stream));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_set_worklist, n * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_img_done, n * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_set_done, 2 * n * sizeof(uint32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_unit_counter, sizeof(unsigned long long)));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_img_done, 0, n * sizeof(uint32_t), ctx->stream));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_set_done, 0, 2 * n * sizeof(uint32_t), ctx->stream));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_blob_xy, n * ctx->cfg.max_blobs * 2 * sizeof(int32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_blob_n, n * sizeof(int32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_img_flags, n * sizeof(int32_t)));
    CUDA_TRY(ctx, cudaMemsetAsync(ctx->d_seg_count, 0, n * sizeof(uint32_t), ctx->stream));   // kept zero by k_blob_reduce afterwards
    ctx->cap_images = n_images;
    return MOCAP_OK;
}

static int ensure_sets(mocap_ctx* ctx, int n_sets) {
    if (n_sets <= ctx->cap_sets) return MOCAP_OK;
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_obj); cudaFree(ctx->d_err); cudaFree(ctx->d_nobj); cudaFree(ctx->d_setflags);
    ctx->d_obj = nullptr; ctx->d_err = nullptr; ctx->d_nobj = nullptr; ctx->d_setflags = nullptr; ctx->cap_sets = 0;
    const size_t n = (size_t)n_sets, R = ctx->cfg.max_roots;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_obj, n * R * 3 * sizeof(double)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_err, n * R * sizeof(double)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_nobj, n * sizeof(int32_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_setflags, n * sizeof(int32_t)));
    ctx->cap_sets = n_sets;
    return MOCAP_OK;
}

// ---- S1 ------------------------------------------------------------------------------------
int mocap_detect_dev(mocap_ctx* ctx, const uint8_t* frames, int n_images, int channels, int threshold,
                     int32_t* blob_xy, int32_t* blob_n, int64_t* blob_mom, int32_t* img_flags) {
    if (!ctx) return MOCAP_EINVAL;
    if (!frames || !blob_xy || !blob_n || n_images < 0 || (channels != 1 && channels != 3))
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_detect_dev: bad argument");
    if ((reinterpret_cast<uintptr_t>(frames) & 15u) != 0)
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_detect_dev: frames must be 16-byte aligned");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    int st = ensure_images(ctx, n_images);
    if (st) return st;
    return launch_detect(ctx, frames, n_images, channels, threshold, blob_xy, blob_n, blob_mom, img_flags);
}

// ---- S2+S3 ---------------------------------------------------------------------------------
int mocap_match_triangulate_dev(mocap_ctx* ctx, const int32_t* blob_xy, const int32_t* blob_n, int n_frame_sets,
                                double* obj, double* err, int32_t* n_obj, int32_t* set_flags, int32_t* chosen) {
    if (!ctx) return MOCAP_EINVAL;
    if (!blob_xy || !blob_n || !obj || !err || !n_obj || n_frame_sets < 0)
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_match_triangulate_dev: bad argument");
    if (!ctx->cameras_set) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_cameras has not been called");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    return launch_match(ctx, blob_xy, blob_n, n_frame_sets, obj, err, n_obj, set_flags, chosen);
}

// ---- S1+S2+S3 ------------------------------------------------------------------------------
int mocap_pipeline_dev(mocap_ctx* ctx, const uint8_t* frames, int n_frame_sets, int channels, int threshold,
                       double* obj, double* err, int32_t* n_obj, int32_t* set_flags) {
    return mocap_pipeline_tracks_dev(ctx, frames, n_frame_sets, channels, threshold, obj, err, n_obj, set_flags, nullptr);
}

int mocap_pipeline_tracks_dev(mocap_ctx* ctx, const uint8_t* frames, int n_frame_sets, int channels, int threshold,
                              double* obj, double* err, int32_t* n_obj, int32_t* set_flags, int32_t* track_xy) {
    if (!ctx) return MOCAP_EINVAL;
    if (!frames || !obj || !err || !n_obj || n_frame_sets < 0 || (channels != 1 && channels != 3))
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_pipeline_dev: bad argument");
    if (!ctx->cameras_set) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_cameras has not been called");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    const int C = ctx->cfg.n_cam;
    const size_t set_bytes = (size_t)C * ctx->cfg.width * ctx->cfg.height * channels;
    const bool fused = pick_fused(ctx, channels);
    // frame-sets per launch group: bounds the segment-list scratch (max_segments * 4 B per image)
    const int chunk = fused ? (65536 / C > 0 ? 65536 / C : 1) : 4096;
    int st = ensure_images(ctx, (n_frame_sets < chunk ? n_frame_sets : chunk) * C);
    if (st) return st;
    for (int s0 = 0; s0 < n_frame_sets; s0 += chunk) {
        const int ns = (n_frame_sets - s0 < chunk) ? n_frame_sets - s0 : chunk;
        // the launchers hand this to the matcher (frame-set indices inside a launch group start at 0)
        ctx->track_xy_cur = track_xy ? track_xy + (size_t)s0 * ctx->cfg.max_roots * C * 2 : nullptr;
        if (fused) {
            if (ctx->use_tma)
                st = launch_pipeline_tma(ctx, frames + (size_t)s0 * set_bytes, ns, threshold, obj + (size_t)s0 * ctx->cfg.max_roots * 3,
                                         err + (size_t)s0 * ctx->cfg.max_roots, n_obj + s0, set_flags ? set_flags + s0 : nullptr);
            else
                st = launch_pipeline_fused(ctx, frames + (size_t)s0 * set_bytes, ns, threshold, obj + (size_t)s0 * ctx->cfg.max_roots * 3,
                                           err + (size_t)s0 * ctx->cfg.max_roots, n_obj + s0, set_flags ? set_flags + s0 : nullptr, channels);
            if (st) break;
            continue;
        }
        st = launch_detect(ctx, frames + (size_t)s0 * set_bytes, ns * C, channels, threshold,
                           ctx->d_blob_xy, ctx->d_blob_n, nullptr, ctx->d_img_flags);
        if (st) break;
        ctx->img_flags_cur = ctx->d_img_flags;
        st = launch_match(ctx, ctx->d_blob_xy, ctx->d_blob_n, ns, obj + (size_t)s0 * ctx->cfg.max_roots * 3,
                          err + (size_t)s0 * ctx->cfg.max_roots, n_obj + s0, set_flags ? set_flags + s0 : nullptr, nullptr);
        ctx->img_flags_cur = nullptr;
        if (st) break;
    }
    ctx->track_xy_cur = nullptr;
    return st;
}

int mocap_pipeline_host(mocap_ctx* ctx, const uint8_t* frames, int n_frame_sets, int channels, int threshold,
                        double* obj, double* err, int32_t* n_obj, int32_t* set_flags) {
    if (!ctx) return MOCAP_EINVAL;
    if (!frames || !obj || !err || !n_obj || n_frame_sets < 0 || (channels != 1 && channels != 3))
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_pipeline_host: bad argument");
    if (!ctx->cameras_set) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_cameras has not been called");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    const int C = ctx->cfg.n_cam, RM = ctx->cfg.max_roots;
    const size_t set_bytes = (size_t)C * ctx->cfg.width * ctx->cfg.height * channels;
    int chunk = (int)((size_t)(256u << 20) / set_bytes);      // ~256 MB per staging buffer
    if (chunk < 1) chunk = 1;
    if (chunk > n_frame_sets) chunk = n_frame_sets > 0 ? n_frame_sets : 1;
    const size_t need = (size_t)chunk * set_bytes;
    if (need > ctx->stage_bytes) {
        CUDA_TRY(ctx, cudaDeviceSynchronize());
        for (int k = 0; k < 2; ++k) { cudaFree(ctx->d_stage[k]); ctx->d_stage[k] = nullptr; }
        ctx->stage_bytes = 0;
        for (int k = 0; k < 2; ++k) CUDA_TRY(ctx, cudaMalloc(&ctx->d_stage[k], need));
        ctx->stage_bytes = need;
    }
    int st = ensure_sets(ctx, n_frame_sets);
    if (st) return st;
    st = ensure_images(ctx, chunk * C);
    if (st) return st;
    cudaEvent_t* copied = ctx->copied;
    // the copies must not start before earlier work on the caller's stream has finished with the staging buffers
    CUDA_TRY(ctx, cudaEventRecord(copied[0], ctx->stream));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream, copied[0], 0));
    CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->copy_stream2, copied[0], 0));
    const bool fused = pick_fused(ctx, channels);
    int k = 0, n_chunks = 0;
    for (int s0 = 0; s0 < n_frame_sets; s0 += chunk, k ^= 1, ++n_chunks) {
        const int ns = (n_frame_sets - s0 < chunk) ? n_frame_sets - s0 : chunk;
        cudaStream_t cs = k ? ctx->copy_stream2 : ctx->copy_stream;
        if (n_chunks >= 2) CUDA_TRY(ctx, cudaStreamWaitEvent(cs, ctx->stage_free[k], 0));
        CUDA_TRY(ctx, cudaMemcpyAsync(ctx->d_stage[k], frames + (size_t)s0 * set_bytes, (size_t)ns * set_bytes,
                                      cudaMemcpyHostToDevice, cs));
        CUDA_TRY(ctx, cudaEventRecord(copied[k], cs));
        CUDA_TRY(ctx, cudaStreamWaitEvent(ctx->stream, copied[k], 0));
        if (fused) {
            if (ctx->use_tma)
                st = launch_pipeline_tma(ctx, ctx->d_stage[k], ns, threshold, ctx->d_obj + (size_t)s0 * RM * 3, ctx->d_err + (size_t)s0 * RM,
                                         ctx->d_nobj + s0, ctx->d_setflags + s0);
            else
                st = launch_pipeline_fused(ctx, ctx->d_stage[k], ns, threshold, ctx->d_obj + (size_t)s0 * RM * 3, ctx->d_err + (size_t)s0 * RM,
                                           ctx->d_nobj + s0, ctx->d_setflags + s0, channels);
            if (st) break;
            CUDA_TRY(ctx, cudaEventRecord(ctx->stage_free[k], ctx->stream));
            continue;
        }
        st = launch_detect(ctx, ctx->d_stage[k], ns * C, channels, threshold, ctx->d_blob_xy, ctx->d_blob_n, nullptr, ctx->d_img_flags);
        if (st) break;
        CUDA_TRY(ctx, cudaEventRecord(ctx->stage_free[k], ctx->stream));
        ctx->img_flags_cur = ctx->d_img_flags;
        st = launch_match(ctx, ctx->d_blob_xy, ctx->d_blob_n, ns, ctx->d_obj + (size_t)s0 * RM * 3, ctx->d_err + (size_t)s0 * RM,
                          ctx->d_nobj + s0, ctx->d_setflags + s0, nullptr);
        ctx->img_flags_cur = nullptr;
        if (st) break;
    }
    if (st) return st;
    const size_t n = (size_t)n_frame_sets;
    CUDA_TRY(ctx, cudaMemcpyAsync(obj, ctx->d_obj, n * RM * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(err, ctx->d_err, n * RM * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(n_obj, ctx->d_nobj, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (set_flags) CUDA_TRY(ctx, cudaMemcpyAsync(set_flags, ctx->d_setflags, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MOCAP_OK;
}

int mocap_locate_objects_dev(mocap_ctx* ctx, const double* obj, const double* err, const int32_t* n_obj, int n_frame_sets,
                             int max_objects, double* objects, int32_t* drone_index, int32_t* n_objects) {
    if (!ctx) return MOCAP_EINVAL;
    if (!obj || !err || !n_obj || !objects || !drone_index || !n_objects || n_frame_sets < 0 || max_objects < 1)
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_locate_objects_dev: bad argument");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    return launch_locate(ctx, obj, err, n_obj, n_frame_sets, max_objects, objects, drone_index, n_objects);
}

// ---- S3 on explicit correspondences ----------------------------------------------------------
int mocap_triangulate_dev(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                          double* X, double* err, uint8_t* valid) {
    if (!ctx) return MOCAP_EINVAL;
    if (!obs || !mask || !X || n_points < 0) return mocap_fail(ctx, MOCAP_EINVAL, "mocap_triangulate_dev: bad argument");
    if (!ctx->cameras_set) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_cameras has not been called");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    return launch_triangulate(ctx, obs, mask, n_points, nullptr, X, err, valid);
}

static int tri_host_common(mocap_ctx* ctx, const double* obs, const uint8_t* mask, const double* X_in, int n_points,
                           double* X, double* err, uint8_t* valid) {
    if (!ctx) return MOCAP_EINVAL;
    if (!obs || !mask || n_points < 0) return mocap_fail(ctx, MOCAP_EINVAL, "triangulate: bad argument");
    if (!ctx->cameras_set) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_cameras has not been called");
    if (n_points == 0) return MOCAP_OK;
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    const int C = ctx->cfg.n_cam;
    const size_t n = (size_t)n_points;
    const size_t b_obs = n * C * 2 * sizeof(double), b_X = n * 3 * sizeof(double), b_err = n * sizeof(double);
    const size_t b_mask = (n * C + 15) & ~(size_t)15, b_valid = (n + 15) & ~(size_t)15;
    int st = ensure_scratch(ctx, b_obs + 2 * b_X + b_err + b_mask + b_valid);
    if (st) return st;
    unsigned char* p = static_cast<unsigned char*>(ctx->d_scratch);
    double* d_obs = reinterpret_cast<double*>(p); p += b_obs;
    double* d_X = reinterpret_cast<double*>(p); p += b_X;
    double* d_Xin = reinterpret_cast<double*>(p); p += b_X;
    double* d_err = reinterpret_cast<double*>(p); p += b_err;
    uint8_t* d_mask = p; p += b_mask;
    uint8_t* d_valid = p;
    CUDA_TRY(ctx, cudaMemcpyAsync(d_obs, obs, b_obs, cudaMemcpyHostToDevice, ctx->stream));
    CUDA_TRY(ctx, cudaMemcpyAsync(d_mask, mask, n * C, cudaMemcpyHostToDevice, ctx->stream));
    if (X_in) CUDA_TRY(ctx, cudaMemcpyAsync(d_Xin, X_in, b_X, cudaMemcpyHostToDevice, ctx->stream));
    st = launch_triangulate(ctx, d_obs, d_mask, n_points, X_in ? d_Xin : nullptr, d_X, d_err, d_valid);
    if (st) return st;
    if (X) CUDA_TRY(ctx, cudaMemcpyAsync(X, d_X, b_X, cudaMemcpyDeviceToHost, ctx->stream));
    if (err) CUDA_TRY(ctx, cudaMemcpyAsync(err, d_err, b_err, cudaMemcpyDeviceToHost, ctx->stream));
    if (valid) CUDA_TRY(ctx, cudaMemcpyAsync(valid, d_valid, n, cudaMemcpyDeviceToHost, ctx->stream));
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return MOCAP_OK;
}

int mocap_triangulate_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                           double* X, double* err, uint8_t* valid) {
    if (ctx && !X) return mocap_fail(ctx, MOCAP_EINVAL, "mocap_triangulate_host: X is NULL");
    return tri_host_common(ctx, obs, mask, nullptr, n_points, X, err, valid);
}

int mocap_reprojection_errors_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask, const double* X,
                                   int n_points, double* err, uint8_t* valid) {
    if (ctx && (!X || !err)) return mocap_fail(ctx, MOCAP_EINVAL, "mocap_reprojection_errors_host: NULL argument");
    return tri_host_common(ctx, obs, mask, X, n_points, nullptr, err, valid);
}

// ---- misc ------------------------------------------------------------------------------------
int mocap_host_alloc(void** out, uint64_t bytes) {
    if (!out) return MOCAP_EINVAL;
    *out = nullptr;
    if (cudaHostAlloc(out, (size_t)bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); return MOCAP_ENOMEM; }
    return MOCAP_OK;
}
void mocap_host_free(void* p) { if (p) cudaFreeHost(p); }

uint64_t mocap_launch_count(const mocap_ctx* ctx) { return ctx ? ctx->launches : 0; }

int mocap_enable_kernel_timing(mocap_ctx* ctx, int on) {
    if (!ctx) return MOCAP_EINVAL;
    ctx->timing_on = on ? 1 : 0;
    return MOCAP_OK;
}
int mocap_detect_kernel_ms(mocap_ctx* ctx, int reset, double* avg_ms, int* n_launches) {
    if (!ctx) return MOCAP_EINVAL;
    const int st = timing_flush(ctx);
    if (st) return st;
    if (avg_ms) *avg_ms = ctx->detect_ms_n ? ctx->detect_ms_sum / ctx->detect_ms_n : 0.0;
    if (n_launches) *n_launches = ctx->detect_ms_n;
    if (reset) { ctx->detect_ms_sum = 0.0; ctx->detect_ms_n = 0; }
    return MOCAP_OK;
}

}  // extern "C"
