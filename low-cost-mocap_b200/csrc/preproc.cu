// SURVEY.md section 8(f) "next" #2: capture-side preprocessing, the step right before S1.
//
// Replaces the per-camera body of Cameras._camera_read (reference computer_code/api/helpers.py:70-82):
//   rot90 -> make_square (zero pad to a square + 8-row feather, helpers.py:507-523) -> cv.undistort ->
//   cv.GaussianBlur 9x9 (sigma 0) -> cv.filter2D with the 5x5 sharpening kernel -> cvtColor RGB2BGR
// as ONE kernel per batch of frames: every CTA produces a 32x32 output tile and keeps all
// intermediates in shared memory, so a raw frame is read once and the processed frame written once
// (the reference makes five full passes over every frame on the CPU).
//
// Every stage follows OpenCV's 8-bit arithmetic exactly (each verified bit-for-bit against cv2 4.13):
//   undistort   initUndistortRectifyMap in double -> fixed-point map (1/32 px) -> bilinear remap with
//               integer weights (32-fx)(32-fy)*32 ... summing to 2^15, (acc + 2^14) >> 15,
//               BORDER_CONSTANT 0.  The map is built once per session on the host.
//   GaussianBlur 9x9, sigma 0 -> sigma 1.7 -> the fixed-point kernel [4 13 30 51 60 51 30 13 4]/256,
//               separable, 16.16 accumulation, (acc + 2^15) >> 16, BORDER_REFLECT_101
//   filter2D    integer correlation with the 5x5 kernel, BORDER_REFLECT_101, saturate to [0, 255]
#include <vector>
#include "common.cuh"

#define PP_T 32                     // output tile
#define PP_G (PP_T + 4)             // blurred region needed by the 5x5 filter
#define PP_U (PP_T + 12)            // undistorted region needed by the 9x9 blur of that

// undistortion map per camera: m1 int16 [S][S][2] integer source coordinates (x, y);
//                            m2 uint16 [S][S]   (fy << 5) | fx, the 1/32 px fractions

__device__ __forceinline__ int reflect101(int i, int n) {
    if (i < 0) i = -i;
    if (i >= n) i = 2 * (n - 1) - i;
    return i;
}

// pixel of make_square(rot90(raw, k)) at (y, x), channel c; 0 outside (helpers.py:507-523)
__device__ __forceinline__ int squared_pixel(const uint8_t* __restrict__ raw, int in_w, int in_h, int rot, int S, int ay,
                                             int y, int x, int c) {
    if (x < 0 || x >= S || y < 0 || y >= S) return 0;
    int ry = y - ay;                       // row in the rotated frame
    int scale8 = 8;                        // feather: value * (1 - (i+1)/8), truncated
    if (ry < 0) {                          // rows above the frame: copies of row 0, fading out
        const int i = -ry - 1;
        if (i >= 8) return 0;
        scale8 = 7 - i; ry = 0;
    } else if (ry >= in_h) {
        const int i = ry - in_h;
        if (i >= 8) return 0;
        scale8 = 7 - i; ry = in_h - 1;
    }
    int sx = x, sy = ry;
    if (rot == 2) { sx = in_w - 1 - x; sy = in_h - 1 - ry; }      // np.rot90(k=2)
    const int v = raw[((size_t)sy * in_w + sx) * 3 + c];
    return (v * scale8) >> 3;              // exact: (1 - alpha) is a multiple of 1/8
}

__global__ void __launch_bounds__(256)
k_preprocess(const uint8_t* __restrict__ raw_frames, int n_images, int C, int in_w, int in_h, int S,
             const int* __restrict__ rotation, const int16_t* __restrict__ m1, const uint16_t* __restrict__ m2,
             uint8_t* __restrict__ out) {
    __shared__ uint8_t U[PP_U][PP_U][3];
    __shared__ uint16_t Gh[PP_U][PP_G][3];          // horizontal pass, Q8.8
    __shared__ uint8_t G[PP_G][PP_G][3];
    const int img = blockIdx.z, cam = img % C;
    const int x0 = blockIdx.x * PP_T, y0 = blockIdx.y * PP_T;
    const uint8_t* raw = raw_frames + (size_t)img * in_w * in_h * 3;
    const int rot = rotation[cam];
    const int ay = (S - in_h) / 2;
    // absolute coordinate ranges held by the tile buffers (clipped to the image: reads go through reflect101)
    const int gx_lo = max(0, x0 - 2), gx_hi = min(S, x0 + PP_T + 2), gy_lo = max(0, y0 - 2), gy_hi = min(S, y0 + PP_T + 2);
    const int ux_lo = max(0, gx_lo - 4), ux_hi = min(S, gx_hi + 4), uy_lo = max(0, gy_lo - 4), uy_hi = min(S, gy_hi + 4);
    const int uw = ux_hi - ux_lo, uh = uy_hi - uy_lo, gw = gx_hi - gx_lo, gh = gy_hi - gy_lo;

    // ---- 1. undistorted pixels of the region: fixed-point bilinear remap of the squared frame
    const int16_t* m1c = m1 + (size_t)cam * S * S * 2;
    const uint16_t* m2c = m2 + (size_t)cam * S * S;
    for (int i = threadIdx.x; i < uw * uh; i += blockDim.x) {
        const int ty = i / uw, tx = i - ty * uw;
        const int y = uy_lo + ty, x = ux_lo + tx;
        const int sx = m1c[((size_t)y * S + x) * 2], sy = m1c[((size_t)y * S + x) * 2 + 1];
        const int f = m2c[(size_t)y * S + x];
        const int fx = f & 31, fy = (f >> 5) & 31;
        const int w00 = (32 - fx) * (32 - fy) * 32, w01 = fx * (32 - fy) * 32, w10 = (32 - fx) * fy * 32, w11 = fx * fy * 32;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const int acc = squared_pixel(raw, in_w, in_h, rot, S, ay, sy, sx, c) * w00 +
                            squared_pixel(raw, in_w, in_h, rot, S, ay, sy, sx + 1, c) * w01 +
                            squared_pixel(raw, in_w, in_h, rot, S, ay, sy + 1, sx, c) * w10 +
                            squared_pixel(raw, in_w, in_h, rot, S, ay, sy + 1, sx + 1, c) * w11;
            U[ty][tx][c] = (uint8_t)((acc + (1 << 14)) >> 15);
        }
    }
    __syncthreads();

    // ---- 2. Gaussian 9x9, horizontal then vertical, fixed point
    const int kg[9] = {4, 13, 30, 51, 60, 51, 30, 13, 4};
    for (int i = threadIdx.x; i < uh * gw; i += blockDim.x) {
        const int ty = i / gw, tx = i - ty * gw;
        const int x = gx_lo + tx;
        int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int ux = reflect101(x + k - 4, S) - ux_lo;
            a0 += kg[k] * U[ty][ux][0]; a1 += kg[k] * U[ty][ux][1]; a2 += kg[k] * U[ty][ux][2];
        }
        Gh[ty][tx][0] = (uint16_t)a0; Gh[ty][tx][1] = (uint16_t)a1; Gh[ty][tx][2] = (uint16_t)a2;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < gh * gw; i += blockDim.x) {
        const int ty = i / gw, tx = i - ty * gw;
        const int y = gy_lo + ty;
        int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int k = 0; k < 9; ++k) {
            const int uy = reflect101(y + k - 4, S) - uy_lo;
            a0 += kg[k] * Gh[uy][tx][0]; a1 += kg[k] * Gh[uy][tx][1]; a2 += kg[k] * Gh[uy][tx][2];
        }
        G[ty][tx][0] = (uint8_t)((a0 + (1 << 15)) >> 16);
        G[ty][tx][1] = (uint8_t)((a1 + (1 << 15)) >> 16);
        G[ty][tx][2] = (uint8_t)((a2 + (1 << 15)) >> 16);
    }
    __syncthreads();

    // ---- 3. 5x5 sharpening filter (helpers.py:75-80), saturate, RGB -> BGR, store
    const int kf[5][5] = {{-2, -1, -1, -1, -2}, {-1, 1, 3, 1, -1}, {-1, 3, 4, 3, -1}, {-1, 1, 3, 1, -1}, {-2, -1, -1, -1, -2}};
    uint8_t* dst = out + (size_t)img * S * S * 3;
    for (int i = threadIdx.x; i < PP_T * PP_T; i += blockDim.x) {
        const int ty = i / PP_T, tx = i - ty * PP_T;
        const int y = y0 + ty, x = x0 + tx;
        if (y >= S || x >= S) continue;
        int a0 = 0, a1 = 0, a2 = 0;
#pragma unroll
        for (int dy = 0; dy < 5; ++dy) {
            const int gy = reflect101(y + dy - 2, S) - gy_lo;
#pragma unroll
            for (int dx = 0; dx < 5; ++dx) {
                const int gx = reflect101(x + dx - 2, S) - gx_lo;
                a0 += kf[dy][dx] * G[gy][gx][0]; a1 += kf[dy][dx] * G[gy][gx][1]; a2 += kf[dy][dx] * G[gy][gx][2];
            }
        }
        uint8_t* o = dst + ((size_t)y * S + x) * 3;
        o[0] = (uint8_t)min(max(a2, 0), 255);          // cvtColor RGB2BGR (helpers.py:82)
        o[1] = (uint8_t)min(max(a1, 0), 255);
        o[2] = (uint8_t)min(max(a0, 0), 255);
    }
}

// cv.initUndistortRectifyMap(K, dist, I, K, (S, S), CV_16SC2): per output pixel the source position
// in 1/32 px (double arithmetic, round half to even).  dist = k1 k2 p1 p2 k3.
static void build_undistort_map(const double* K, const double* dist, int S, int16_t* m1, uint16_t* m2) {
    const double fx = K[0], fy = K[4], u0 = K[2], v0 = K[5];
    const double k1 = dist[0], k2 = dist[1], p1 = dist[2], p2 = dist[3], k3 = dist[4];
    // inverse of the (upper triangular) new camera matrix
    const double ir0 = 1.0 / fx, ir1 = -K[1] / (fx * fy), ir2 = (K[1] * v0 - u0 * fy) / (fx * fy);
    const double ir4 = 1.0 / fy, ir5 = -v0 / fy;
    for (int i = 0; i < S; ++i) {
        const double xr = i * ir1 + ir2, yr = i * ir4 + ir5;
        for (int j = 0; j < S; ++j) {
            const double x = xr + j * ir0, y = yr;
            const double x2 = x * x, y2 = y * y, r2 = x2 + y2, _2xy = 2 * x * y;
            const double kr = (1 + ((k3 * r2 + k2) * r2 + k1) * r2) / 1.0;
            const double xd = x * kr + p1 * _2xy + p2 * (r2 + 2 * x2);
            const double yd = y * kr + p1 * (r2 + 2 * y2) + p2 * _2xy;
            const double u = fx * xd + u0, v = fy * yd + v0;
            const long iu = lrint(u * 32.0), iv = lrint(v * 32.0);
            long qx = iu >> 5, qy = iv >> 5;
            if (qx < -32768) qx = -32768; if (qx > 32767) qx = 32767;
            if (qy < -32768) qy = -32768; if (qy > 32767) qy = 32767;
            m1[((size_t)i * S + j) * 2] = (int16_t)qx;
            m1[((size_t)i * S + j) * 2 + 1] = (int16_t)qy;
            m2[(size_t)i * S + j] = (uint16_t)(((iv & 31) << 5) | (iu & 31));
        }
    }
}

extern "C" {

int mocap_set_preprocess(mocap_ctx* ctx, int in_width, int in_height, const int* rotation, const double* K, const double* dist) {
    if (!ctx) return MOCAP_EINVAL;
    const int C = ctx->cfg.n_cam, S = ctx->cfg.width;
    if (!rotation || !K || !dist || ctx->cfg.height != S || in_width != S || in_height > S - 16 || in_height < 1)
        return mocap_fail(ctx, MOCAP_EINVAL, "mocap_set_preprocess: the context must be square (width == height == raw width) and the raw frame "
                                             "landscape with at least 8 pad rows above and below (the reference's make_square only works for that)");
    for (int c = 0; c < C; ++c)
        if (rotation[c] != 0 && rotation[c] != 2)
            return mocap_fail(ctx, MOCAP_EINVAL, "mocap_set_preprocess: rotation must be 0 or 2 (a quarter turn makes the frame portrait, which make_square cannot feather)");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    std::vector<int16_t> m1((size_t)C * S * S * 2);
    std::vector<uint16_t> m2((size_t)C * S * S);
    for (int c = 0; c < C; ++c) build_undistort_map(K + 9 * c, dist + 5 * c, S, m1.data() + (size_t)c * S * S * 2, m2.data() + (size_t)c * S * S);
    CUDA_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->d_pp_m1); cudaFree(ctx->d_pp_m2); cudaFree(ctx->d_pp_rot);
    ctx->d_pp_m1 = nullptr; ctx->d_pp_m2 = nullptr; ctx->d_pp_rot = nullptr;
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_pp_m1, m1.size() * sizeof(int16_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_pp_m2, m2.size() * sizeof(uint16_t)));
    CUDA_TRY(ctx, cudaMalloc(&ctx->d_pp_rot, C * sizeof(int)));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_pp_m1, m1.data(), m1.size() * sizeof(int16_t), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_pp_m2, m2.data(), m2.size() * sizeof(uint16_t), cudaMemcpyHostToDevice));
    CUDA_TRY(ctx, cudaMemcpy(ctx->d_pp_rot, rotation, C * sizeof(int), cudaMemcpyHostToDevice));
    ctx->pp_in_w = in_width; ctx->pp_in_h = in_height;
    return MOCAP_OK;
}

int mocap_get_undistort_map(mocap_ctx* ctx, int cam, int16_t* m1, uint16_t* m2) {
    if (!ctx) return MOCAP_EINVAL;
    const int S = ctx->cfg.width;
    if (!ctx->d_pp_m1 || cam < 0 || cam >= ctx->cfg.n_cam || !m1 || !m2) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_get_undistort_map: no preprocessing set / bad argument");
    CUDA_TRY(ctx, cudaMemcpy(m1, ctx->d_pp_m1 + (size_t)cam * S * S * 2, (size_t)S * S * 2 * sizeof(int16_t), cudaMemcpyDeviceToHost));
    CUDA_TRY(ctx, cudaMemcpy(m2, ctx->d_pp_m2 + (size_t)cam * S * S, (size_t)S * S * sizeof(uint16_t), cudaMemcpyDeviceToHost));
    return MOCAP_OK;
}

int mocap_preprocess_dev(mocap_ctx* ctx, const uint8_t* raw_frames, int n_images, uint8_t* out_frames) {
    if (!ctx) return MOCAP_EINVAL;
    if (!raw_frames || !out_frames || n_images < 0) return mocap_fail(ctx, MOCAP_EINVAL, "mocap_preprocess_dev: bad argument");
    if (!ctx->d_pp_m1) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_preprocess has not been called");
    if (n_images == 0) return MOCAP_OK;
    if (n_images > 65535) return mocap_fail(ctx, MOCAP_EINVAL, "mocap_preprocess_dev: at most 65535 images per call");
    CUDA_TRY(ctx, cudaSetDevice(ctx->cfg.device));
    const int S = ctx->cfg.width, tiles = (S + PP_T - 1) / PP_T;
    k_preprocess<<<dim3(tiles, tiles, n_images), 256, 0, ctx->stream>>>(raw_frames, n_images, ctx->cfg.n_cam, ctx->pp_in_w, ctx->pp_in_h, S,
                                                                          ctx->d_pp_rot, ctx->d_pp_m1, ctx->d_pp_m2, out_frames);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 1;
    return MOCAP_OK;
}

int mocap_pipeline_raw_dev(mocap_ctx* ctx, const uint8_t* raw_frames, int n_frame_sets, int threshold, uint8_t* processed,
                           double* obj, double* err, int32_t* n_obj, int32_t* set_flags) {
    if (!ctx) return MOCAP_EINVAL;
    if (!raw_frames || !obj || !err || !n_obj || n_frame_sets < 0) return mocap_fail(ctx, MOCAP_EINVAL, "mocap_pipeline_raw_dev: bad argument");
    if (!ctx->d_pp_m1) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_preprocess has not been called");
    if (!ctx->cameras_set) return mocap_fail(ctx, MOCAP_ESTATE, "mocap_set_cameras has not been called");
    if (n_frame_sets == 0) return MOCAP_OK;
    const int C = ctx->cfg.n_cam, S = ctx->cfg.width;
    const size_t img_bytes = (size_t)S * S * 3;
    const int chunk = 4096 / C > 0 ? 4096 / C : 1;             // frame-sets per launch group (bounded scratch)
    uint8_t* work = processed;
    if (!work) {                                               // caller does not want the frames: recycle one chunk of scratch
        int st = ensure_scratch(ctx, (size_t)chunk * C * img_bytes);
        if (st) return st;
        work = static_cast<uint8_t*>(ctx->d_scratch);
    }
    for (int s0 = 0; s0 < n_frame_sets; s0 += chunk) {
        const int ns = n_frame_sets - s0 < chunk ? n_frame_sets - s0 : chunk;
        uint8_t* dst = processed ? processed + (size_t)s0 * C * img_bytes : work;
        int st = mocap_preprocess_dev(ctx, raw_frames + (size_t)s0 * C * ctx->pp_in_w * ctx->pp_in_h * 3, ns * C, dst);
        if (st) return st;
        st = mocap_pipeline_dev(ctx, dst, ns, 3, threshold, obj + (size_t)s0 * ctx->cfg.max_roots * 3, err + (size_t)s0 * ctx->cfg.max_roots,
                                n_obj + s0, set_flags ? set_flags + s0 : nullptr);
        if (st) return st;
    }
    return MOCAP_OK;
}

}  // extern "C"
