/*
 * mocap_b200.h -- C ABI of libmocap_b200.so: the B200 (sm_100a) multi-view
 * marker-tracking core that stands in for the per-frame geometry path of
 * jyjblrd/Low-Cost-Mocap, computer_code/api/helpers.py.
 *
 * The reference has no FFI of its own: its hot path is five module-level Python
 * functions.  Each entry point below names the reference function (file:line,
 * relative to the reference repository) it replaces; the Python mirror that
 * re-creates the reference signatures on top of this ABI lives in
 * low-cost-mocap_b200/api.py and the binding a maintainer adds is shown in
 * INTEGRATION.md.
 *
 * Conventions
 *   - plain C, plain pointers and sizes; no torch / numpy types.
 *   - every function returns MOCAP_OK (0) or a negative MOCAP_E* code;
 *     mocap_last_error(ctx) gives the text of the last failure of that ctx.
 *   - "_dev" entry points take DEVICE pointers and only enqueue work on the
 *     context's stream (mocap_set_stream); they never synchronise.
 *     "_host" entry points take HOST pointers, copy in/out and return when the
 *     results are in the host buffers.
 *   - one context per host thread (the reference's Cameras singleton is
 *     unsynchronised, Singleton.py:3); contexts are independent.  A context owns scratch
 *     (segment lists, counters, blob lists) shared by all of its calls: use it from ONE
 *     stream at a time -- before pointing it at another stream with mocap_set_stream, the
 *     work already enqueued through it must have finished or be ordered before the new
 *     stream's work (event).  The bundle-adjustment entry points use a workspace of their
 *     own and may run on a second stream next to the pipeline entry points.
 *   - there is no CPU fallback: without a CUDA device every call fails with
 *     MOCAP_ENODEV.
 *
 * Layouts (row-major, innermost last)
 *   frames      uint8  [n_frame_sets][n_cam][H][W]        (channels == 1)
 *               uint8  [n_frame_sets][n_cam][H][W][3]     (channels == 3, the
 *               layout Cameras._find_dot receives, helpers.py:143)
 *   blob_xy     int32  [n_images][max_blobs][2]   (x, y) = int(m10/m00), int(m01/m00)
 *   blob_n      int32  [n_images]                 n_images = n_frame_sets * n_cam
 *   blob_mom    int64  [n_images][max_blobs][4]   {2*m00, 6*m10, 6*m01, pixel count}
 *   img_flags   int32  [n_images]                 MOCAP_F_* bits
 *   obj         double [n_frame_sets][max_roots][3]
 *   err         double [n_frame_sets][max_roots]  mean squared reprojection error, px^2
 *   n_obj       int32  [n_frame_sets]
 *   set_flags   int32  [n_frame_sets]             MOCAP_F_* bits
 *   chosen      int32  [n_frame_sets][max_roots][n_cam]  blob index per camera of the
 *                                                 winning correspondence, -1 = no view
 */
#ifndef MOCAP_B200_H
#define MOCAP_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MOCAP_OK          0
#define MOCAP_EINVAL     -1   /* bad argument                                  */
#define MOCAP_ENODEV     -2   /* no usable CUDA device / wrong architecture    */
#define MOCAP_ECUDA      -3   /* a CUDA runtime call failed                    */
#define MOCAP_ENOMEM     -4
#define MOCAP_ESTATE     -5   /* e.g. cameras not set                          */

/* per-image / per-frame-set overflow flags (results for that unit are truncated) */
#define MOCAP_F_SEGMENTS  1   /* more above-threshold 16-px segments than max_segments */
#define MOCAP_F_BLOBS     2   /* more blobs than max_blobs                               */
#define MOCAP_F_ROOTS     4   /* more roots than max_roots                               */
#define MOCAP_F_CANDS     8   /* more than max_cands candidates on one epipolar line     */
#define MOCAP_F_GROUPS   16   /* more than max_groups candidate groups for one root      */
#define MOCAP_F_HOLES    32   /* per image, not an overflow.  Blobs with holes are reproduced as cv.findContours(RETR_TREE) +
                                 cv.moments treat them (helpers.py:147-158): one more contour -- one more point -- per hole, the
                                 outer contour's moments over the FILLED blob, cv2's hierarchy order.  The bit is left only
                                 when that slow path could not run: a holed blob wider or taller than 62 pixels, or more
                                 than 64 holes in the image; the image then carries one centre per blob from its set pixels. */

#if defined(__GNUC__)
#define MOCAP_API __attribute__((visibility("default")))
#else
#define MOCAP_API
#endif

typedef struct mocap_ctx mocap_ctx;

typedef struct mocap_config {
    int device;         /* CUDA device ordinal                                            */
    int n_cam;          /* cameras per frame-set (1..16)                                   */
    int width;          /* image width, multiple of 16 (640)                              */
    int height;         /* image height (480)                                             */
    int max_blobs;      /* blobs kept per image (<= 64)                                   */
    int max_segments;   /* above-threshold 16-px segments handled per image (power of two, 64..4096) */
    int max_roots;      /* roots per frame-set in the matcher (<= 128)                    */
    int max_cands;      /* candidates kept per (root, camera) (<= 16)                     */
    int max_groups;     /* candidate groups evaluated per root                            */
} mocap_config;

/* Fills *cfg with defaults for n_cam cameras of width x height. */
MOCAP_API void mocap_default_config(mocap_config* cfg, int n_cam, int width, int height);

MOCAP_API int  mocap_create(mocap_ctx** out, const mocap_config* cfg);
MOCAP_API void mocap_destroy(mocap_ctx* ctx);
MOCAP_API const char* mocap_last_error(const mocap_ctx* ctx);
/* Text for a status when no context exists (mocap_create failed). */
MOCAP_API const char* mocap_status_string(int status);

/* cudaStream_t on which all *_dev work is enqueued (NULL = the legacy default stream). */
MOCAP_API int mocap_set_stream(mocap_ctx* ctx, void* cuda_stream);

/* Camera model of the session.  HOST pointers: K[n_cam][9], R[n_cam][9], t[n_cam][3],
 * row-major doubles.  Replaces the state the reference keeps in
 * Cameras.camera_params (helpers.py:19-22,188-193) and in the camera_poses list
 * every hot-path function receives (helpers.py:293,339; set at helpers.py:171-175).
 * Builds on the host, once: P_kc = K_k [R_c|t_c] (helpers.py:305-308,351-355) and the
 * fundamental matrices F_rc of every ordered camera pair
 * (cv.sfm.fundamentalFromProjections, helpers.py:362). */
MOCAP_API int mocap_set_cameras(mocap_ctx* ctx, const double* K, const double* R, const double* t);

/* Optional epilogue of the matcher: object points leave in world coordinates
 * (helpers.py:96-103: flip x,y; 4x4 to_world_coords_matrix; dehomogenise; swap y,z).
 * M = NULL switches it off (default).  HOST pointer, 16 doubles row-major. */
MOCAP_API int mocap_set_world_transform(mocap_ctx* ctx, const double* M);

/* S1 -- replaces Cameras._find_dot (helpers.py:143-163) for n_images images at once:
 * gray (cvtColor RGB2GRAY when channels == 3), pix > threshold, 8-connected blobs,
 * contour-polygon moments, centre = int(m10/m00), int(m01/m00); zero-area blobs are
 * dropped; blobs leave in cv.findContours order (descending raster position of the
 * blob's first pixel).  blob_mom and img_flags may be NULL. */
MOCAP_API int mocap_detect_dev(mocap_ctx* ctx, const uint8_t* frames, int n_images, int channels,
                     int threshold, int32_t* blob_xy, int32_t* blob_n,
                     int64_t* blob_mom, int32_t* img_flags);

/* S2+S3 -- replaces find_point_correspondance_and_object_points (helpers.py:339-421),
 * incl. the triangulate_points / calculate_reprojection_errors calls inside it
 * (helpers.py:408-419), for n_frame_sets frame-sets at once.  Input is the output of
 * mocap_detect_dev.  chosen and set_flags may be NULL. */
MOCAP_API int mocap_match_triangulate_dev(mocap_ctx* ctx, const int32_t* blob_xy, const int32_t* blob_n,
                                int n_frame_sets, double* obj, double* err, int32_t* n_obj,
                                int32_t* set_flags, int32_t* chosen);

/* S1+S2+S3 back to back on the context's stream, intermediate blob lists kept in
 * context-owned device memory: the per-frame body of Cameras._camera_read
 * (helpers.py:84-103) without capture/preprocessing. */
MOCAP_API int mocap_pipeline_dev(mocap_ctx* ctx, const uint8_t* frames, int n_frame_sets, int channels,
                       int threshold, double* obj, double* err, int32_t* n_obj, int32_t* set_flags);

/* mocap_pipeline_dev that also leaves, for every emitted point, the pixel of the winning correspondence in each
 * camera: track_xy int32 [n_frame_sets][max_roots][n_cam][2], (-1, -1) where the camera has no view (NULL: not
 * written).  This is the (F, C, 2) image_points array with None entries that bundle_adjustment receives
 * (helpers.py:244, index.py:272), per triangulated point. */
MOCAP_API int mocap_pipeline_tracks_dev(mocap_ctx* ctx, const uint8_t* frames, int n_frame_sets, int channels,
                              int threshold, double* obj, double* err, int32_t* n_obj, int32_t* set_flags,
                              int32_t* track_xy);

/* Same with HOST buffers: frames are copied to the device in chunks overlapped with
 * compute, results copied back; returns after the results are in the host buffers.
 * For best speed pass page-locked memory (mocap_host_alloc). */
MOCAP_API int mocap_pipeline_host(mocap_ctx* ctx, const uint8_t* frames, int n_frame_sets, int channels,
                        int threshold, double* obj, double* err, int32_t* n_obj, int32_t* set_flags);

/* Capture-side preprocessing -- replaces the per-camera body of Cameras._camera_read before S1
 * (helpers.py:70-82): rot90, make_square (zero pad + 8-row feather, helpers.py:507-523), cv.undistort,
 * cv.GaussianBlur 9x9, cv.filter2D with the 5x5 sharpening kernel, cvtColor RGB2BGR -- in one kernel,
 * bit-exact with cv2's 8-bit arithmetic.  The context must be square: width == height == in_width (the
 * reference's make_square only handles landscape frames padded top and bottom).  HOST pointers:
 * rotation int [n_cam] (0 or 2, camera-params.json "rotation"), K double [n_cam][9], dist double
 * [n_cam][5] = k1 k2 p1 p2 k3 (camera-params.json "distortion_coef"). */
MOCAP_API int mocap_set_preprocess(mocap_ctx* ctx, int in_width, int in_height, const int* rotation,
                         const double* K, const double* dist);
/* raw uint8 [n_images][in_height][in_width][3] -> out uint8 [n_images][S][S][3].  DEVICE pointers.
 * n_images = n_frame_sets * n_cam (camera index = image index mod n_cam). */
MOCAP_API int mocap_preprocess_dev(mocap_ctx* ctx, const uint8_t* raw_frames, int n_images, uint8_t* out_frames);
/* The whole per-frame body of Cameras._camera_read (helpers.py:70-103) for n_frame_sets frame-sets of RAW
 * camera frames: preprocessing (as mocap_preprocess_dev) -> S1 -> S2+S3.  The preprocessing kernel also
 * emits the grey plane _find_dot's cvtColor(RGB2GRAY) would derive from the processed frame
 * (helpers.py:144), so S1-S3 read 1 byte per pixel; results are those of S1 on the 3-channel frames.
 * processed (uint8 [n_images][S][S][3], the frames the reference goes on to JPEG-encode) may be NULL when
 * the caller does not display them (then they are never written).  DEVICE pointers. */
MOCAP_API int mocap_pipeline_raw_dev(mocap_ctx* ctx, const uint8_t* raw_frames, int n_frame_sets, int threshold,
                           uint8_t* processed, double* obj, double* err, int32_t* n_obj, int32_t* set_flags);
/* The fixed-point undistortion map of one camera as built by mocap_set_preprocess (the tables
 * cv.initUndistortRectifyMap(..., CV_16SC2) returns): m1 int16 [S][S][2], m2 uint16 [S][S].  HOST pointers. */
MOCAP_API int mocap_get_undistort_map(mocap_ctx* ctx, int cam, int16_t* m1, uint16_t* m2);

/* Tracking hand-off -- replaces locate_objects (helpers.py:424-480) for n_frame_sets frame-sets at
 * once: marker triplets (two markers 0.15 apart, a third 0.095 from both, tolerance 0.025) -> object
 * records.  Input is the matcher's output (obj/err/n_obj, with the world transform set if the caller
 * wants world coordinates as the reference does).  objects double [n_frame_sets][max_objects][5] =
 * {x, y, z, heading, error}; drone_index int32 [n_frame_sets][max_objects]; n_objects int32
 * [n_frame_sets].  DEVICE pointers. */
MOCAP_API int mocap_locate_objects_dev(mocap_ctx* ctx, const double* obj, const double* err, const int32_t* n_obj,
                             int n_frame_sets, int max_objects, double* objects, int32_t* drone_index,
                             int32_t* n_objects);

/* S3 -- replaces triangulate_points (helpers.py:330-336) and
 * calculate_reprojection_errors (helpers.py:203-211) on explicit correspondences.
 * obs double [n_points][n_cam][2], mask uint8 [n_points][n_cam] (0 = [None, None]).
 * X double [n_points][3], err double [n_points], valid uint8 [n_points]
 * (0 where the reference returns [None]*3 / None, i.e. fewer than two views).
 * err may be NULL.  DEVICE pointers. */
MOCAP_API int mocap_triangulate_dev(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                          double* X, double* err, uint8_t* valid);
/* HOST-pointer convenience form of the above. */
MOCAP_API int mocap_triangulate_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                           double* X, double* err, uint8_t* valid);
/* calculate_reprojection_errors for GIVEN points (helpers.py:203-241). HOST pointers. */
MOCAP_API int mocap_reprojection_errors_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask,
                                   const double* X, int n_points, double* err, uint8_t* valid);

/* Cold-start extrinsics -- replaces the body of calculate_camera_pose up to its bundle_adjustment call
 * (index.py:229-270): per adjacent camera pair a fundamental matrix from the common observations,
 * E = K1^T F K0 (cv.sfm.essentialFromFundamental with the intrinsics of cameras 0 and 1), the four
 * motions of cv.sfm.motionFromEssential, the reference's cheirality vote and the pose chain.  The
 * reference's F comes from a randomised cv.findFundamentalMat(FM_RANSAC); here F is a deterministic
 * normalised 8-point estimate re-fitted twice on its 1 px Sampson inliers, or -- F_given != NULL --
 * supplied by the caller (double [n_cam-1][9], x2^T F x1 = 0).  HOST pointers: obs/mask as for
 * mocap_bundle_adjust_host; R [n_cam][9], t [n_cam][3] out; F_used [n_cam-1][9] and votes
 * [n_cam-1][4] (points in front of the cameras per candidate) may be NULL. */
MOCAP_API int mocap_calibrate_init_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                              const double* F_given, double* R, double* t, double* F_used, int* votes);

/* S4 -- replaces bundle_adjustment (helpers.py:244-290): robust (Cauchy) trust-region
 * least squares over the poses of cameras 1..C-1 (rotation vector + translation;
 * camera 0 pinned at (I,0); the reference's focal parameters are dead, helpers.py:267-270),
 * residual_j = float32(mean squared reprojection error of point j after DLT
 * re-triangulation with the trial poses).  HOST pointers; R,t are in/out. */
typedef struct mocap_ba_options {
    double ftol;        /* 1e-2  (helpers.py:288)                */
    double xtol;        /* 1e-8  (scipy default)                 */
    double gtol;        /* 1e-8  (scipy default)                 */
    int    max_nfev;    /* 0 -> 100 * n_params (scipy default)   */
    int    jacobian;    /* 0: 2-point finite differences of the float32 residuals -- exactly what scipy
                              differentiates for the reference (percent-level quantisation noise,
                              SURVEY.md section 7); 1 (default): the same differences taken before the
                              float32 cast */
    int    prefit;      /* 1 (default): first run a classic Levenberg-Marquardt bundle adjustment over
                              poses AND points (analytic Jacobians, per-point 3x3 blocks eliminated by
                              Schur complement, dense reduced camera system) on the plain squared
                              reprojection error, then polish on the reference objective; 0: reference
                              iteration only */
    int    prefit_max_iter;  /* 50 */
    int    engine;      /* 0 (default): the whole solve in ONE persistent, grid-synchronous kernel (k_ba_solve): no host round
                              trips; 1: host-stepped (optimiser control on the host, one launch per phase, a stream
                              synchronisation per step) -- the same algorithm, kept as a cross-check */
} mocap_ba_options;

typedef struct mocap_ba_report {
    double cost_initial;   /* reference objective 0.5 * sum log1p(r^2) at the start   */
    double cost_final;     /* ... at the returned poses                                */
    double optimality;     /* ||J^T f||_inf at the end                                 */
    int    n_iterations;   /* trust-region iterations on the reference objective       */
    int    n_fev;          /* residual-vector evaluations (each = n_points DLTs)       */
    int    status;         /* scipy-style: 0 max_nfev, 1 gtol, 2 ftol, 3 xtol, 4 both  */
    int    n_residuals;
    double prefit_cost_initial;  /* 0.5 * sum of squared pixel residuals before / after the prefit */
    double prefit_cost_final;
    int    prefit_iterations;
    int    n_launches;     /* kernels launched by this call                            */
    int    n_tr_solves;    /* device-resident solve: trust-region sub-problems solved, and ... */
    int    n_tr_newton;    /* ... Newton iterations on the damping alpha they took in total (<= 10 each)  */
    float  phase_ms[8];    /* device-resident solve only: wall time per phase, CTA 0's clock --
                              0 set-up and control, 1 prefit: reduced camera system (Schur complement), 2 prefit: dense solve,
                              3 prefit: back-substitution + trial cost, 4 polish: finite-difference Jacobian + normal equations,
                              5 polish: tridiagonalisation, 6 polish: trust-region sub-problems, 7 polish: trial evaluations */
} mocap_ba_report;

MOCAP_API void mocap_ba_default_options(mocap_ba_options* opt);
MOCAP_API int  mocap_bundle_adjust_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                              double* R, double* t, const mocap_ba_options* opt,
                              mocap_ba_report* report);
/* The same with DEVICE buffers, stream-ordered, never synchronises: ONE cooperative launch of k_ba_solve runs the
 * whole bundle adjustment (residuals, Jacobians, Schur complement, the <= 90 x 90 dense solves, step control).
 * obs double [n_points_max][n_cam][2], mask uint8 [n_points_max][n_cam]; n_points (device int32, may be NULL =
 * n_points_max) is read by the kernel, so the count can come from mocap_tracks_to_observations_dev without a
 * host round trip; R [n_cam][9], t [n_cam][3] device doubles in/out; opt is a HOST pointer (NULL = defaults; the
 * engine field is ignored); report is a DEVICE pointer (may be NULL; status -3: no point with two views). */
MOCAP_API int  mocap_bundle_adjust_dev(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points_max,
                             const int32_t* n_points, double* R, double* t, const mocap_ba_options* opt,
                             mocap_ba_report* report);
/* CTAs of k_ba_solve for this context (1 .. number of SMs; 0 = the default, one per SM).  One solve is a cooperative
 * grid whose serial sections (dense solves, tridiagonalisation) leave most CTAs waiting at the grid barrier, so
 * INDEPENDENT solves -- the reference runs one bundle_adjustment per recorded batch (index.py:249-276), a session with
 * several batches has several -- finish sooner side by side: K contexts on K streams with SMs / K CTAs each
 * (measured, 8 cameras x 18 800 points: 4 solves 5.24 ms in turn on 148 CTAs, 3.87 ms as 2 x 74, 3.2 ms as 4 x 37).
 * The result does not depend on the number of CTAs. */
MOCAP_API int  mocap_set_ba_grid(mocap_ctx* ctx, int n_ctas);
/* Matcher output of a batch -> the explicit correspondences S4 consumes (BASELINE config 3: S1-S3, then one
 * bundle adjustment per batch), on the device: track_xy int32 [n_frame_sets][max_roots][n_cam][2] as written by
 * mocap_pipeline_tracks_dev ((-1, -1) = no view), n_obj / err the matcher's outputs; tracks whose reprojection
 * error exceeds max_err are left out (max_err <= 0: keep all; err may then be NULL).  One row per kept track, in
 * frame order then root order: obs double [capacity][n_cam][2], mask uint8 [capacity][n_cam]; *n_points (device)
 * = number of rows written (<= capacity).  DEVICE pointers, stream-ordered. */
MOCAP_API int  mocap_tracks_to_observations_dev(mocap_ctx* ctx, const int32_t* track_xy, const int32_t* n_obj,
                                      const double* err, int n_frame_sets, double max_err, double* obs,
                                      uint8_t* mask, int32_t* n_points, int capacity);
/* residual vector of S4 at explicit poses (helpers.py:264-276); r float [n_points],
 * valid uint8 [n_points]; returns the number of valid residuals in *n_valid. HOST pointers. */
MOCAP_API int  mocap_ba_residuals_host(mocap_ctx* ctx, const double* obs, const uint8_t* mask, int n_points,
                             const double* R, const double* t, float* r, uint8_t* valid, int* n_valid);

/* page-locked host memory for the *_host entry points */
MOCAP_API int  mocap_host_alloc(void** out, uint64_t bytes);
MOCAP_API void mocap_host_free(void* p);

/* Number of kernels this context has launched so far (bench.py's gpu_launches). */
MOCAP_API uint64_t mocap_launch_count(const mocap_ctx* ctx);
/* Average device time in ms of the blob-detection kernel over the launches since the
 * last call with reset != 0, measured with CUDA events on the context's stream when
 * enabled by mocap_enable_kernel_timing(ctx, 1).  Synchronises the stream. */
MOCAP_API int  mocap_enable_kernel_timing(mocap_ctx* ctx, int on);
MOCAP_API int  mocap_detect_kernel_ms(mocap_ctx* ctx, int reset, double* avg_ms, int* n_launches);

#ifdef __cplusplus
}
#endif
#endif /* MOCAP_B200_H */
