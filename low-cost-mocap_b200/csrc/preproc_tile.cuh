// One output tile of the capture-side preprocessing (preproc.cu), written as four per-thread stage
// functions over byte arrays so that the SAME arithmetic can be stepped through on the host
// (tests/hostcheck/preproc_host.cpp runs the stages tid by tid) and checked against the reference's
// cv2 chain on a machine without a GPU.
//
// Reference: computer_code/api/helpers.py:70-82 (Cameras._camera_read) and :507-523 (make_square).
//
// Tile = PP_TX x PP_TY output pixels.  All intermediates are PLANAR per channel in shared memory so
// the filters run on packed bytes:
//   U    [3][PP_UH][PP_UW]      u8   undistorted pixels of the tile + 6 px apron (BORDER_REFLECT_101 is
//                                    applied here, by undistorting the mirrored coordinate, so no later
//                                    stage needs border logic: the blur of a mirrored image is the mirror
//                                    of the blur because the kernel is symmetric)
//   GhT  [PP_GHT_CH][PP_GW][PP_GHT_STRIDE]   u32  horizontal 9-tap pass, Q8.8 in u16, TRANSPOSED and packed as row
//                                    pairs so the vertical pass is a 2-way 16x8 dot product per word
//   G    [3][PP_GH][PP_GW]      u8   blurred pixels of the tile + 2 px apron (in U's storage: plane c of G ends
//                                    before plane c of U does, so writing G channel by channel, each after its
//                                    horizontal pass has consumed U, never touches a U plane still to be read)
// OpenCV's 8-bit GaussianBlur accumulates exactly (kernel sums to 256, Q8.8 after one pass, Q8.16 after
// two, one rounding at the end), so the result is (sum_yx ky kx U + 2^15) >> 16 in any order.
#pragma once
#include <stdint.h>
#include <stddef.h>

#if defined(__CUDACC__)
#define PP_HD __host__ __device__ __forceinline__
#define PP_UNROLL _Pragma("unroll")
#else
#define PP_HD static inline
#define PP_UNROLL
#endif

#ifndef PP_F
#define PP_F 2                             // frames of ONE camera per CTA: a pixel's map decode is shared by them
#endif
#define PP_TX 64
#ifndef PP_TY
#define PP_TY 64                           // multiple of 4 (64x64: 48 KB of shared memory, 1.41x apron overhead; 64x32: 28 KB, 1.63x)
#endif
#define PP_UW (PP_TX + 12)                 // 76 bytes = 19 words per row
#define PP_UH (PP_TY + 12)                 // 76 rows = 38 row pairs
#define PP_GW (PP_TX + 4)                  // 68 = 17 groups of 4
#define PP_GH (PP_TY + 4)                  // 68 = 17 groups of 4
#define PP_GHT_STRIDE (PP_UH / 2 + 1)      // 39 words per column: odd, so column-parallel access is conflict free
#define PP_U_BYTES (3 * PP_UH * PP_UW)
#ifndef PP_GHT_CH
#define PP_GHT_CH 1                        // channel planes GhT holds (1 or 3): the two blur passes run per group of
#endif                                     // PP_GHT_CH channels; 1 keeps the tile at 28 KB so that 5 CTAs are resident
#define PP_GHT_BYTES (PP_GHT_CH * PP_GW * PP_GHT_STRIDE * 4)
#define PP_G_BYTES (3 * PP_GH * PP_GW)
// one U per frame; G reuses its frame's U: U is dead once the horizontal pass has run, G is born in the vertical pass
#define PP_SMEM_BYTES (PP_F * PP_U_BYTES + PP_GHT_BYTES)
static_assert(PP_G_BYTES <= PP_U_BYTES && PP_U_BYTES % 4 == 0 && PP_TY % 4 == 0 && PP_TX % 4 == 0, "tile layout");

struct PPFrame {
    const uint8_t* raw[PP_F];  // [in_h][in_w][3] raw frames of ONE camera (consecutive frame-sets)
    uint8_t* out[PP_F];        // [S][S][3] processed frames, or null
    uint8_t* gray[PP_F];       // [S][S] what _find_dot's cvtColor(RGB2GRAY) makes of the processed frame, or null
    int n_frames;              // 1 .. PP_F entries of the arrays above are in use
    const int32_t* m1;         // [n_cam][S][S] (sy << 16) | (sx & 0xffff): integer source coordinates of cv2's fixed-point map
    const uint16_t* m2;        // [n_cam][S][S] (fy << 5) | fx, the 1/32 px fractions
    int in_w, in_h, S, rot, ay;
    int map_offset;            // cam * S * S: where this camera's map starts in m1 / m2
    int word_stores;           // out / gray rows are 4-byte aligned (S % 4 == 0 and aligned bases)
};

// ---- packed dot products (native on the device, spelled out on the host) ---------------------------
PP_HD uint32_t pp_dp4a_uu(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp4a(a, b, c);
#else
    for (int i = 0; i < 4; ++i) c += ((a >> (8 * i)) & 255u) * ((b >> (8 * i)) & 255u);
    return c;
#endif
}
// a: four unsigned bytes, b: four SIGNED bytes
PP_HD int pp_dp4a_us(uint32_t a, uint32_t b, int c) {
#if defined(__CUDA_ARCH__)
    int d;
    asm("dp4a.u32.s32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    for (int i = 0; i < 4; ++i) c += (int)((a >> (8 * i)) & 255u) * (int)(int8_t)((b >> (8 * i)) & 255u);
    return c;
#endif
}
// a: two u16, b: two u8 in its low half
PP_HD uint32_t pp_dp2a_lo(uint32_t a, uint32_t b, uint32_t c) {
#if defined(__CUDA_ARCH__)
    return __dp2a_lo(a, b, c);
#else
    return c + (a & 0xffffu) * (b & 255u) + (a >> 16) * ((b >> 8) & 255u);
#endif
}

// ---- filter taps ---------------------------------------------------------------------------------------
// GaussianBlur 9x9 with sigma 0 -> sigma 1.7 -> OpenCV's fixed-point kernel, /256
PP_HD constexpr uint32_t pp_gauss_tap(int t) {
    return t == 0 || t == 8 ? 4u : t == 1 || t == 7 ? 13u : t == 2 || t == 6 ? 30u : t == 3 || t == 5 ? 51u : t == 4 ? 60u : 0u;
}
// weights for output j of a group of 4 against word w of the 12-byte window: byte b meets tap 4w + b - j
PP_HD constexpr uint32_t pp_gauss_word4(int j, int w) {
    return pp_gauss_tap(4 * w + 0 - j) | (pp_gauss_tap(4 * w + 1 - j) << 8) | (pp_gauss_tap(4 * w + 2 - j) << 16) |
           (pp_gauss_tap(4 * w + 3 - j) << 24);
}
// weights for output j of a group of 4 rows against row-pair word m: halves meet taps 2m - j and 2m + 1 - j
PP_HD constexpr uint32_t pp_gauss_word2(int j, int m) { return pp_gauss_tap(2 * m - j) | (pp_gauss_tap(2 * m + 1 - j) << 8); }

// the 5x5 sharpening kernel of helpers.py:75-80
PP_HD constexpr int pp_sharpen_tap(int dy, int dx) {
    return (dx < 0 || dx > 4) ? 0
         : (dy == 0 || dy == 4) ? ((dx == 0 || dx == 4) ? -2 : -1)
         : (dx == 0 || dx == 4) ? -1
         : (dy == 2) ? (dx == 2 ? 4 : 3)
         : (dx == 2 ? 3 : 1);
}
PP_HD constexpr uint32_t pp_sharpen_word4(int dy, int j, int w) {
    return ((uint32_t)(pp_sharpen_tap(dy, 4 * w + 0 - j) & 255)) | ((uint32_t)(pp_sharpen_tap(dy, 4 * w + 1 - j) & 255) << 8) |
           ((uint32_t)(pp_sharpen_tap(dy, 4 * w + 2 - j) & 255) << 16) | ((uint32_t)(pp_sharpen_tap(dy, 4 * w + 3 - j) & 255) << 24);
}

// BORDER_REFLECT_101 for coordinates within one image width of the image: |i| mirrored about n - 1
// (mirror = 2 (n - 1)).  The clamp only ever touches apron pixels of partial tiles that no valid output reads.
PP_HD int pp_reflect101(int i, int n, int mirror) {
    const int a = i < 0 ? -i : i;
    const int b = mirror - a;
    const int r = a < b ? a : b;
    return r < 0 ? 0 : r;
}

// Where make_square(rot90(raw, k)) takes its pixel (y, x) from (helpers.py:507-523): the frame is centred
// vertically in the square, the 8 rows above and below it are copies of the edge row fading out, everything
// else is 0.  Returns the feather factor in eighths (0: the pixel is 0) and the byte offset into raw.
PP_HD int pp_squared_tap(const PPFrame& f, int y, int x, int& offset) {
    offset = 0;
    if (x < 0 || x >= f.S || y < 0 || y >= f.S) return 0;
    int ry = y - f.ay;
    int scale8 = 8;                        // feather: value * (1 - (i+1)/8), truncated
    if (ry < 0) {
        if (ry < -8) return 0;
        scale8 = 8 + ry; ry = 0;
    } else if (ry >= f.in_h) {
        if (ry >= f.in_h + 8) return 0;
        scale8 = 7 - (ry - f.in_h); ry = f.in_h - 1;
    }
    int sx = x, sy = ry;
    if (f.rot == 2) { sx = f.in_w - 1 - x; sy = f.in_h - 1 - ry; }      // np.rot90(k=2)
    offset = (sy * f.in_w + sx) * 3;
    return scale8;
}

// ---- stage 1: cv.undistort = fixed-point bilinear remap (weights sum to 2^15 in OpenCV; the common factor
// 32 is dropped here: (32 s + 2^14) >> 15 == (s + 2^9) >> 10), BORDER_CONSTANT 0
PP_HD int pp_map_index(const PPFrame& f, int x0, int y0, int i, int mirror) {
    const int ty = (int)((unsigned)i / (unsigned)PP_UW), tx = i - ty * PP_UW;
    const int y = pp_reflect101(y0 - 6 + ty, f.S, mirror), x = pp_reflect101(x0 - 6 + tx, f.S, mirror);
    return f.map_offset + y * f.S + x;                     // 32-bit index: n_cam * S * S < 2^31 is checked at set-up
}

PP_HD void pp_stage_undistort(const PPFrame& f, uint8_t* U, int x0, int y0, int tid, int nt) {
    const int S = f.S, mirror = 2 * (S - 1), N = PP_UH * PP_UW;
    // the map entry of the NEXT pixel is fetched before the current pixel's taps are, so the two dependent
    // global round trips of a pixel (map -> taps) overlap with its neighbour's
    int32_t m_next = 0;
    int fr_next = 0;
    if (tid < N) { const int mi = pp_map_index(f, x0, y0, tid, mirror); m_next = f.m1[mi]; fr_next = f.m2[mi]; }
    for (int i = tid; i < N; i += nt) {
        const int32_t m = m_next;
        const int fr = fr_next;
        if (i + nt < N) { const int mi = pp_map_index(f, x0, y0, i + nt, mirror); m_next = f.m1[mi]; fr_next = f.m2[mi]; }
        const int sx = (int16_t)(m & 0xffff), sy = m >> 16;
        const int fx = fr & 31, fy = (fr >> 5) & 31;
        const int w00 = (32 - fx) * (32 - fy), w01 = fx * (32 - fy), w10 = (32 - fx) * fy, w11 = fx * fy;
        const int ry = sy - f.ay;
        // everything above is per PIXEL OF THE CAMERA; the frames of this camera only differ in the bytes read
        if ((unsigned)sx < (unsigned)(S - 1) && (unsigned)ry < (unsigned)(f.in_h - 1)) {
            // all four taps inside the camera frame proper: two row pointers, constant byte offsets
            const bool flip = f.rot == 2;
            const int o = flip ? ((f.in_h - 1 - ry) * f.in_w + (f.in_w - 1 - sx)) * 3 : (ry * f.in_w + sx) * 3;
            PP_UNROLL
            for (int k = 0; k < PP_F; ++k) {
                if (k >= f.n_frames) break;
                const uint8_t* p = f.raw[k] + o;
                int v0, v1, v2;
                if (flip) {
                    const uint8_t* q = p - 3 * f.in_w;
                    v0 = p[0] * w00 + p[-3] * w01 + q[0] * w10 + q[-3] * w11;
                    v1 = p[1] * w00 + p[-2] * w01 + q[1] * w10 + q[-2] * w11;
                    v2 = p[2] * w00 + p[-1] * w01 + q[2] * w10 + q[-1] * w11;
                } else {
                    const uint8_t* q = p + 3 * f.in_w;
                    v0 = p[0] * w00 + p[3] * w01 + q[0] * w10 + q[3] * w11;
                    v1 = p[1] * w00 + p[4] * w01 + q[1] * w10 + q[4] * w11;
                    v2 = p[2] * w00 + p[5] * w01 + q[2] * w10 + q[5] * w11;
                }
                uint8_t* u = U + k * PP_U_BYTES + i;       // i == ty * PP_UW + tx
                u[0] = (uint8_t)((v0 + 512) >> 10);
                u[PP_UH * PP_UW] = (uint8_t)((v1 + 512) >> 10);
                u[2 * PP_UH * PP_UW] = (uint8_t)((v2 + 512) >> 10);
            }
        } else {
            // rows -9 .. in_h + 7: the feathered rows, the frame edge or the side border, tap by tap (rare, kept
            // compact); anything further out is the zero padding of make_square and stays 0
            const bool some = (unsigned)(ry + 9) < (unsigned)(f.in_h + 17);
            PP_UNROLL
            for (int k = 0; k < PP_F; ++k) {                // unrolled so that f.raw[k] stays in registers
                if (k >= f.n_frames) break;
                int v0 = 512, v1 = 512, v2 = 512;
                if (some) {
#if defined(__CUDACC__)
#pragma unroll 1
#endif
                    for (int t = 0; t < 4; ++t) {
                        int o;
                        const int s8 = pp_squared_tap(f, sy + (t >> 1), sx + (t & 1), o);
                        if (s8 == 0) continue;
                        const int w = t == 0 ? w00 : t == 1 ? w01 : t == 2 ? w10 : w11;
                        const uint8_t* p = f.raw[k] + o;
                        v0 += ((p[0] * s8) >> 3) * w;      // exact: the feather factor is a multiple of 1/8
                        v1 += ((p[1] * s8) >> 3) * w;
                        v2 += ((p[2] * s8) >> 3) * w;
                    }
                }
                uint8_t* u = U + k * PP_U_BYTES + i;
                u[0] = (uint8_t)(v0 >> 10);
                u[PP_UH * PP_UW] = (uint8_t)(v1 >> 10);
                u[2 * PP_UH * PP_UW] = (uint8_t)(v2 >> 10);
            }
        }
    }
}

// ---- stage 2a: horizontal 9 taps.  One item = 4 outputs x 2 rows of one channel: 3 words per row in,
// 3 four-way dot products per output, 4 row-pair words out.
// Channels c0 .. c0 + PP_GHT_CH - 1 of U -> GhT (which holds PP_GHT_CH channel planes).
PP_HD void pp_stage_blur_h(const uint8_t* U, uint32_t* GhT, int c0, int tid, int nt) {
    const uint32_t* Uw = reinterpret_cast<const uint32_t*>(U);
    const int RP = PP_UH / 2, NG = PP_GW / 4;
    for (int i = tid; i < PP_GHT_CH * RP * NG; i += nt) {
        const unsigned ui = (unsigned)i, q1 = ui / (unsigned)RP, q2 = q1 / (unsigned)NG;
        const int rp = (int)(ui - q1 * RP), g = (int)(q1 - q2 * NG), c = (int)q2;
        uint32_t o[2][4];
        PP_UNROLL
        for (int r = 0; r < 2; ++r) {
            const uint32_t* row = Uw + ((c0 + c) * PP_UH + 2 * rp + r) * (PP_UW / 4) + g;
            const uint32_t w0 = row[0], w1 = row[1], w2 = row[2];
            PP_UNROLL
            for (int j = 0; j < 4; ++j)
                o[r][j] = pp_dp4a_uu(w0, pp_gauss_word4(j, 0), pp_dp4a_uu(w1, pp_gauss_word4(j, 1), pp_dp4a_uu(w2, pp_gauss_word4(j, 2), 0u)));
        }
        PP_UNROLL
        for (int j = 0; j < 4; ++j) GhT[(c * PP_GW + 4 * g + j) * PP_GHT_STRIDE + rp] = o[0][j] | (o[1][j] << 16);
    }
}

// ---- stage 2b: vertical 9 taps + the single rounding.  One item = 4 output rows of one column of one
// channel: 6 row-pair words in, 5 two-way dot products per output.
PP_HD void pp_stage_blur_v(const uint32_t* GhT, uint8_t* G, int c0, int tid, int nt) {
    const int NQ = PP_GH / 4;
    for (int i = tid; i < PP_GHT_CH * NQ * PP_GW; i += nt) {
        const unsigned ui = (unsigned)i, q1 = ui / (unsigned)PP_GW, q2 = q1 / (unsigned)NQ;
        const int x = (int)(ui - q1 * PP_GW), yq = (int)(q1 - q2 * NQ), c = (int)q2;
        const uint32_t* col = GhT + (c * PP_GW + x) * PP_GHT_STRIDE + 2 * yq;
        uint32_t w[6];
        PP_UNROLL
        for (int m = 0; m < 6; ++m) w[m] = col[m];
        PP_UNROLL
        for (int j = 0; j < 4; ++j) {
            uint32_t acc = 1u << 15;
            PP_UNROLL
            for (int m = 0; m < 6; ++m)
                if (pp_gauss_word2(j, m) != 0u) acc = pp_dp2a_lo(w[m], pp_gauss_word2(j, m), acc);
            G[((c0 + c) * PP_GH + 4 * yq + j) * PP_GW + x] = (uint8_t)(acc >> 16);
        }
    }
}

// ---- stage 3: cv.filter2D with the 5x5 kernel (integer correlation, saturate), cvtColor RGB2BGR, store.
// One item = 4 neighbouring output pixels, all 3 channels = 12 contiguous output bytes.
PP_HD void pp_stage_sharpen_store(const PPFrame& f, int k, const uint8_t* G, int x0, int y0, int tid, int nt) {
    const uint32_t* Gw = reinterpret_cast<const uint32_t*>(G);
    const int NG = PP_TX / 4;
    for (int i = tid; i < NG * PP_TY; i += nt) {
        const int g = i % NG, ty = i / NG;
        const int y = y0 + ty, x = x0 + 4 * g;
        if (y >= f.S || x >= f.S) continue;
        int a[3][4];
        PP_UNROLL
        for (int c = 0; c < 3; ++c) {
            PP_UNROLL
            for (int j = 0; j < 4; ++j) a[c][j] = 0;
            PP_UNROLL
            for (int dy = 0; dy < 5; ++dy) {
                const uint32_t* row = Gw + (c * PP_GH + ty + dy) * (PP_GW / 4) + g;
                const uint32_t w0 = row[0], w1 = row[1];
                PP_UNROLL
                for (int j = 0; j < 4; ++j)
                    a[c][j] = pp_dp4a_us(w0, pp_sharpen_word4(dy, j, 0), pp_dp4a_us(w1, pp_sharpen_word4(dy, j, 1), a[c][j]));
            }
            PP_UNROLL
            for (int j = 0; j < 4; ++j) a[c][j] = a[c][j] < 0 ? 0 : (a[c][j] > 255 ? 255 : a[c][j]);
        }
        // RGB -> BGR (helpers.py:82): byte order per pixel is channel 2, 1, 0
        if (f.out[k]) {
            uint8_t* o = f.out[k] + ((size_t)y * f.S + x) * 3;
            if (f.word_stores) {
                uint32_t* ow = reinterpret_cast<uint32_t*>(o);
                ow[0] = (uint32_t)a[2][0] | ((uint32_t)a[1][0] << 8) | ((uint32_t)a[0][0] << 16) | ((uint32_t)a[2][1] << 24);
                ow[1] = (uint32_t)a[1][1] | ((uint32_t)a[0][1] << 8) | ((uint32_t)a[2][2] << 16) | ((uint32_t)a[1][2] << 24);
                ow[2] = (uint32_t)a[0][2] | ((uint32_t)a[2][3] << 8) | ((uint32_t)a[1][3] << 16) | ((uint32_t)a[0][3] << 24);
            } else {
                for (int j = 0; j < 4 && x + j < f.S; ++j) {
                    o[3 * j + 0] = (uint8_t)a[2][j]; o[3 * j + 1] = (uint8_t)a[1][j]; o[3 * j + 2] = (uint8_t)a[0][j];
                }
            }
        }
        // S1 reads the processed frame through cv2's RGB2GRAY (helpers.py:144), 15-bit fixed point with byte 0
        // as "R": emitting that plane here lets the marker pipeline read 1 byte per pixel instead of 3
        if (f.gray[k]) {
            uint32_t gq[4];
            PP_UNROLL
            for (int j = 0; j < 4; ++j) gq[j] = (uint32_t)(a[2][j] * 9798 + a[1][j] * 19235 + a[0][j] * 3735 + 16384) >> 15;
            uint8_t* o = f.gray[k] + (size_t)y * f.S + x;
            if (f.word_stores) *reinterpret_cast<uint32_t*>(o) = gq[0] | (gq[1] << 8) | (gq[2] << 16) | (gq[3] << 24);
            else for (int j = 0; j < 4 && x + j < f.S; ++j) o[j] = (uint8_t)gq[j];
        }
    }
}
