// Test-only: the S2 + S3 device code of low-cost-mocap_b200/csrc/match_device.cuh (epipolar candidate search,
// candidate groups, DLT + reprojection error per group, segmented warp argmin) run UNCHANGED on the host through
// the SIMT emulation in simt_emu.h, 32 std::threads = the one warp that k_match_triangulate gives a frame-set,
// on camera tables built by the same host code mocap_set_cameras uses (csrc/camera_tables.h).
// NOT part of libmocap_b200.so and never used by the product path.
#include "simt_emu.h"
#include "../../low-cost-mocap_b200/csrc/match_device.cuh"
#include "../../low-cost-mocap_b200/csrc/camera_tables.h"

// K [C][9], R [C][9], t [C][3]; blob_xy int32 [n_sets][C][MB][2], blob_n int32 [n_sets][C]
// -> obj [n_sets][RMAX][3], err [n_sets][RMAX], n_obj [n_sets], flags [n_sets]   (as mocap_match_triangulate_dev)
extern "C" int hc_match_triangulate(const double* K, const double* R, const double* t, int C, const int32_t* blob_xy, const int32_t* blob_n,
                                    int n_sets, int MB, int RMAX, int KC, unsigned GMAX, double* obj, double* err, int32_t* n_obj, int32_t* flags) {
    static CameraTables T;                                   // 70 KB: not on the stack
    memset(&T, 0, sizeof(T));
    build_camera_tables(T, C, K, R, t);
    std::vector<unsigned long long> raw(warp_state_bytes(RMAX, C, KC, MB) / 8 + 2);
    for (int set = 0; set < n_sets; ++set) {
        simt::launch(32, [&] {
            WarpState ws = carve_warp_state(reinterpret_cast<unsigned char*>(raw.data()), RMAX, C, KC, MB);
            match_triangulate_warp(&T, ws, blob_xy + (size_t)set * C * MB * 2, blob_n + (size_t)set * C, set, (int)(threadIdx.x & 31),
                                   C, MB, RMAX, KC, GMAX, obj, err, n_obj, flags, nullptr);
        });
    }
    return 0;
}
