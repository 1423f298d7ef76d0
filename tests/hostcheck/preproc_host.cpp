// Test-only host build of the preprocessing tile stages in low-cost-mocap_b200/csrc/preproc_tile.cuh:
// the four stages of every tile are stepped through thread by thread (barriers fall between the stages,
// so running each stage for all threads in turn is what the CTA computes), so that the packed-byte
// arithmetic can be checked against the reference's cv2 chain on a machine without a GPU.
// NOT part of libmocap_b200.so and never used by the product path.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../low-cost-mocap_b200/csrc/preproc_tile.cuh"

extern "C" void hc_preprocess(const uint8_t* raw, int in_w, int in_h, int S, int rot, const int16_t* m1, const uint16_t* m2,
                              uint8_t* out, uint8_t* gray, int word_stores, int n_threads) {
    PPFrame f;
    f.raw = raw; f.m1 = reinterpret_cast<const int32_t*>(m1); f.m2 = m2; f.out = out; f.gray = gray;
    f.in_w = in_w; f.in_h = in_h; f.S = S; f.rot = rot; f.ay = (S - in_h) / 2; f.word_stores = word_stores; f.map_offset = 0;
    std::vector<uint32_t> smem(PP_SMEM_BYTES / 4 + 4);
    uint8_t* base = reinterpret_cast<uint8_t*>(smem.data());
    uint8_t* U = base;
    uint32_t* GhT = reinterpret_cast<uint32_t*>(base + PP_U_BYTES);
    uint8_t* G = base;                                  // over U, as in the kernel
    for (int y0 = 0; y0 < S; y0 += PP_TY)
        for (int x0 = 0; x0 < S; x0 += PP_TX) {
            memset(base, 0xA5, PP_SMEM_BYTES);                 // stale shared memory must not matter
            for (int t = 0; t < n_threads; ++t) pp_stage_undistort(f, U, x0, y0, t, n_threads);
            for (int c0 = 0; c0 < 3; c0 += PP_GHT_CH) {
                for (int t = 0; t < n_threads; ++t) pp_stage_blur_h(U, GhT, c0, t, n_threads);
                for (int t = 0; t < n_threads; ++t) pp_stage_blur_v(GhT, G, c0, t, n_threads);
            }
            for (int t = 0; t < n_threads; ++t) pp_stage_sharpen_store(f, G, x0, y0, t, n_threads);
        }
}
