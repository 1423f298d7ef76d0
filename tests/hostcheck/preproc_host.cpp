// Test-only host build of the preprocessing tile stages in low-cost-mocap_b200/csrc/preproc_tile.cuh:
// the stages of every tile are stepped through thread by thread in the kernel's order (barriers fall between
// the stages, so running each stage for all threads in turn is what the CTA computes), so that the
// packed-byte arithmetic can be checked against the reference's cv2 chain on a machine without a GPU.
// NOT part of libmocap_b200.so and never used by the product path.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../low-cost-mocap_b200/csrc/preproc_tile.cuh"

extern "C" int hc_frames_per_group() { return PP_F; }

// n_frames frames of one camera: raw [n][in_h][in_w][3] -> out [n][S][S][3] (or null), gray [n][S][S] (or null)
extern "C" void hc_preprocess(const uint8_t* raw, int n_frames, int in_w, int in_h, int S, int rot, const int16_t* m1, const uint16_t* m2,
                              uint8_t* out, uint8_t* gray, int word_stores, int n_threads) {
    std::vector<uint32_t> smem(PP_SMEM_BYTES / 4 + 4);
    uint8_t* base = reinterpret_cast<uint8_t*>(smem.data());
    uint8_t* U = base;
    uint32_t* GhT = reinterpret_cast<uint32_t*>(base + PP_F * PP_U_BYTES);
    for (int g0 = 0; g0 < n_frames; g0 += PP_F) {
        PPFrame f;
        f.n_frames = 0;
        for (int k = 0; k < PP_F; ++k) {
            const bool in = g0 + k < n_frames;
            f.raw[k] = in ? raw + (size_t)(g0 + k) * in_w * in_h * 3 : nullptr;
            f.out[k] = (in && out) ? out + (size_t)(g0 + k) * S * S * 3 : nullptr;
            f.gray[k] = (in && gray) ? gray + (size_t)(g0 + k) * S * S : nullptr;
            f.n_frames += in ? 1 : 0;
        }
        f.m1 = reinterpret_cast<const int32_t*>(m1); f.m2 = m2; f.map_offset = 0;
        f.in_w = in_w; f.in_h = in_h; f.S = S; f.rot = rot; f.ay = (S - in_h) / 2; f.word_stores = word_stores;
        for (int y0 = 0; y0 < S; y0 += PP_TY)
            for (int x0 = 0; x0 < S; x0 += PP_TX) {
                memset(base, 0xA5, PP_SMEM_BYTES);                 // stale shared memory must not matter
                for (int t = 0; t < n_threads; ++t) pp_stage_undistort(f, U, x0, y0, t, n_threads);
                for (int k = 0; k < f.n_frames; ++k) {
                    uint8_t* Uk = U + k * PP_U_BYTES;
                    for (int c0 = 0; c0 < 3; c0 += PP_GHT_CH) {
                        for (int t = 0; t < n_threads; ++t) pp_stage_blur_h(Uk, GhT, c0, t, n_threads);
                        for (int t = 0; t < n_threads; ++t) pp_stage_blur_v(GhT, Uk, c0, t, n_threads);
                    }
                    for (int t = 0; t < n_threads; ++t) pp_stage_sharpen_store(f, k, Uk, x0, y0, t, n_threads);
                }
            }
    }
}
