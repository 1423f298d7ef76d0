// Shared declarations of libmocap_b200.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include "../../include/mocap_b200.h"

#define MOCAP_MAX_CAM        16
#define MOCAP_MAX_BLOBS      64
#define MOCAP_MAX_ROOTS      128
#define MOCAP_MAX_CANDS      16
#define MOCAP_ACC_CAP        256     // blobs accumulated per image before max_blobs truncation
#define MOCAP_SEG_PX         16      // pixels per threshold segment (one 128-bit load of 1-channel data)

// Camera tables, device resident (built on the host by mocap_set_cameras).
struct CameraTables {
    double Pkc[MOCAP_MAX_CAM][MOCAP_MAX_CAM][12];  // Pkc[k][c] = K_k [R_c|t_c]  (helpers.py:305-308 uses K of the k-th PRESENT view)
    double F[MOCAP_MAX_CAM][MOCAP_MAX_CAM][9];     // F[r][c]: line in camera c of a point in camera r (helpers.py:362)
    double R[MOCAP_MAX_CAM][9];
    double t[MOCAP_MAX_CAM][3];
    double fx[MOCAP_MAX_CAM], fy[MOCAP_MAX_CAM], cx[MOCAP_MAX_CAM], cy[MOCAP_MAX_CAM];  // cv.projectPoints reads these four of K
    double Kmat[MOCAP_MAX_CAM][9];                 // full intrinsics (S4 rebuilds K_k [R|t] for trial poses)
    double world[16];                              // to_world_coords_matrix (helpers.py:99)
    int    use_world;
    int    n_cam;
};

struct mocap_ctx {
    mocap_config cfg;
    cudaStream_t stream;
    cudaStream_t copy_stream;
    cudaStream_t copy_stream2;   // staging buffers alternate between two copy streams (two copy engines in flight)
    char         err[512];
    bool         cameras_set;
    CameraTables* d_tables;
    CameraTables  h_tables;
    // detection scratch, sized for cap_images
    int       cap_images;
    uint32_t* d_seg_count;
    uint32_t* d_seg_list;
    uint32_t* d_worklist;     // images deferred by the warp-level blob kernel
    uint32_t* d_work_count;   // [4]: image worklist count + finished-CTA counter, set worklist count + finished-CTA counter
    uint32_t* d_set_worklist; // frame-sets deferred by the fused kernel
    uint32_t* d_img_done;     // fused kernel: finished units per image (self-resetting)
    uint32_t* d_set_done;     // fused kernel: [2*cap_images] finished images per set, then deferred marks
    unsigned long long* d_unit_counter;
    int       fused_ctas_per_sm;
    int       tma_ctas_per_sm;   // 0: bulk-copy kernel unavailable for this configuration
    int       use_tma;           // MOCAP_PIPELINE=tma: stream through the bulk-copy ring kernel
    int       use_phased;        // MOCAP_PIPELINE=phased: phase-synchronous variant of the single-pass kernel (opt-in, fused_phased.cuh)
    int       pipeline_auto;     // MOCAP_PIPELINE unset: heavy batches (many blobs per frame-set) take the three-kernel pipeline
    unsigned long long* d_stat_acc;                  // device accumulator of the blob statistic
    volatile unsigned long long* h_stat;             // pinned, mapped: {blobs, images} of the last batch, written by the GPU
    unsigned long long* d_stat_host;                 // device alias of h_stat
    int       use_fused;      // 1: single fused pipeline kernel for 1-channel frames (default)
    int32_t*  d_blob_xy;
    int32_t*  d_blob_n;
    int32_t*  d_img_flags;
    // host-path staging
    uint8_t*  d_stage[2];
    size_t    stage_bytes;
    cudaEvent_t stage_free[2]; cudaEvent_t copied[2];   // per staging buffer: its kernel has finished / its copy has landed
    double*   d_obj; double* d_err; int32_t* d_nobj; int32_t* d_setflags;
    int       cap_sets;
    // generic scratch for the *_host triangulation / BA entry points
    void*     d_scratch; size_t scratch_bytes;
    // S4 on the device (ba_dev.cu): workspace of k_ba_solve, launch shape
    void*     d_ba_ws; size_t ba_ws_bytes; int ba_threads; int ba_grid; size_t ba_smem;
    unsigned* d_match_counter; int match_ctas_per_sm;   // k_match_triangulate: claim counters [4], resident CTAs per SM
    // chunked matcher (match_device.cuh MatchSplit): items, their partial results, the roots each touches, arrivals per frame-set
    void* d_match_items; unsigned long long* d_match_partial; int* d_match_range; unsigned* d_match_arrive;
    int match_chunk, match_item_cap, match_cap_sets;
    const int32_t* img_flags_cur;   // set by the pipelines: the matcher folds the images' S1 flags into the frame-set's
    int32_t*  track_xy_cur;   // set for the duration of mocap_pipeline_tracks_dev: where the matcher leaves the winners' pixels
    // capture-side preprocessing
    int16_t*  d_pp_m1; uint16_t* d_pp_m2; int* d_pp_rot; int pp_in_w, pp_in_h;
    // accounting
    uint64_t  launches;
    int       timing_on;
    cudaEvent_t tim_ev[2 * 64];   // pairs bracketing k_threshold_segments, resolved lazily
    int       tim_used;
    double    detect_ms_sum;
    int       detect_ms_n;
    int       num_sms;
};

int mocap_fail(mocap_ctx* ctx, int code, const char* fmt, ...);
#define CUDA_TRY(ctx, call)                                                              \
    do {                                                                                 \
        cudaError_t e__ = (call);                                                        \
        if (e__ != cudaSuccess)                                                          \
            return mocap_fail((ctx), MOCAP_ECUDA, "%s failed: %s (%s:%d)", #call,        \
                              cudaGetErrorString(e__), __FILE__, __LINE__);              \
    } while (0)

// kernel launchers (each returns a MOCAP_* status; all enqueue on ctx->stream)
int launch_detect(mocap_ctx* ctx, const uint8_t* frames, int n_images, int channels, int threshold,
                  int32_t* blob_xy, int32_t* blob_n, int64_t* blob_mom, int32_t* img_flags);
int launch_match(mocap_ctx* ctx, const int32_t* blob_xy, const int32_t* blob_n, int n_sets,
                 double* obj, double* err, int32_t* n_obj, int32_t* set_flags, int32_t* chosen);
int launch_triangulate(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                       const double* X_in, double* X, double* err, uint8_t* valid);
int ensure_scratch(mocap_ctx* ctx, size_t bytes);
int launch_locate(mocap_ctx* ctx, const double* obj, const double* err, const int32_t* n_obj, int n_sets,
                  int max_objects, double* out, int32_t* drone_index, int32_t* n_out);
int launch_blob_fallback(mocap_ctx* ctx, int32_t* blob_xy, int32_t* blob_n, int64_t* blob_mom, int32_t* img_flags, int n_images);
// frame-sets with more blobs than this (average of the previous batch) are cheaper through the three-kernel
// pipeline: the matcher's large code then runs in a kernel of its own instead of evicting the stream loop
#define MOCAP_HEAVY_BLOBS_PER_SET 48
// candidate groups per work item of the chunked matcher: frame-sets with more are evaluated by several warps
#define MOCAP_MATCH_CHUNK 512
int launch_match_list(mocap_ctx* ctx, const int32_t* blob_xy, const int32_t* blob_n, const uint32_t* set_list, uint32_t* set_count,
                      int n_sets_max, double* obj, double* err, int32_t* n_obj, int32_t* set_flags);
int launch_pipeline_tma(mocap_ctx* ctx, const uint8_t* frames, int n_sets, int threshold,
                        double* obj, double* err, int32_t* n_obj, int32_t* set_flags);
int launch_pipeline_fused(mocap_ctx* ctx, const uint8_t* frames, int n_sets, int threshold,
                          double* obj, double* err, int32_t* n_obj, int32_t* set_flags, int channels = 1);
int timing_flush(mocap_ctx* ctx);
