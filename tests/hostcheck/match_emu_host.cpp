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

// The chunked matcher (match_sets_body + match_chunks_body = k_match_triangulate + k_match_chunks) on a grid of n_ctas CTAs of
// 4 warps: frame-sets with more than `chunk` candidate groups are cut into items of `chunk` groups, item_cap bounds the list.
// stats[0] = items allocated, stats[1] = items claimed, stats[2] = arrival counters left non-zero (must be 0).
extern "C" int hc_match_triangulate_chunked(const double* K, const double* R, const double* t, int C, const int32_t* blob_xy, const int32_t* blob_n,
                                            int n_sets, int MB, int RMAX, int KC, unsigned GMAX, unsigned chunk, unsigned item_cap, int n_ctas,
                                            double* obj, double* err, int32_t* n_obj, int32_t* flags, int32_t* track_xy, long long* stats) {
    static CameraTables T;
    memset(&T, 0, sizeof(T));
    build_camera_tables(T, C, K, R, t);
    const int warps = 4;
    const size_t per_warp = warp_state_bytes(RMAX, C, KC, MB);
    std::vector<std::vector<unsigned long long>> smem(n_ctas, std::vector<unsigned long long>(per_warp * warps / 8 + 2));
    std::vector<unsigned> counters(4, 0u), arrive(n_sets, 0u);
    std::vector<MatchItem> items(item_cap ? item_cap : 1);
    std::vector<unsigned long long> partial((size_t)(item_cap ? item_cap : 1) * RMAX * MATCH_PARTIAL_WORDS, 0xDEADBEEFDEADBEEFull);
    std::vector<int> range((size_t)(item_cap ? item_cap : 1) * 2, -7);
    MatchSplit sp;
    memset(&sp, 0, sizeof(sp));
    sp.counters = counters.data();
    if (chunk) {
        sp.items = items.data(); sp.partial = partial.data(); sp.range = range.data(); sp.arrive = arrive.data();
        sp.chunk = chunk; sp.item_cap = item_cap;
    }
    simt::launch_grid(n_ctas, warps * 32, [&] {
        const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
        WarpState ws = carve_warp_state(reinterpret_cast<unsigned char*>(smem[blockIdx.x].data()) + per_warp * wid, RMAX, C, KC, MB);
        match_sets_body(&T, ws, lane, blob_xy, blob_n, n_sets, C, MB, RMAX, KC, GMAX, sp, obj, err, n_obj, flags, nullptr, track_xy, nullptr);
    });
    if (chunk)
        simt::launch_grid(n_ctas, warps * 32, [&] {
            const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
            WarpState ws = carve_warp_state(reinterpret_cast<unsigned char*>(smem[blockIdx.x].data()) + per_warp * wid, RMAX, C, KC, MB);
            match_chunks_body(&T, ws, lane, blob_xy, blob_n, C, MB, RMAX, KC, GMAX, sp, obj, err, n_obj, flags, nullptr, track_xy, nullptr);
        });
    stats[0] = counters[1]; stats[1] = counters[2]; stats[2] = 0;
    for (unsigned a : arrive) stats[2] += a != 0;
    return 0;
}
