/* Plain C use of libmocap_b200 (what a cgo / JNI / N-API binding would call).
 *   gcc -I include examples/pipeline_host.c -L low-cost-mocap_b200 -lmocap_b200 -o pipeline_host
 * Processes n synthetic frame-sets (one bright square per camera) through S1+S2+S3 from host memory. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "mocap_b200.h"

int main(void) {
    enum { C = 2, W = 640, H = 480, N = 4, R = 64 };
    mocap_config cfg;
    mocap_ctx* ctx = NULL;
    mocap_default_config(&cfg, C, W, H);
    int rc = mocap_create(&ctx, &cfg);
    if (rc != MOCAP_OK) { fprintf(stderr, "mocap_create: %s\n", mocap_status_string(rc)); return 1; }
    const double K[C][9] = {{600, 0, 320, 0, 600, 240, 0, 0, 1}, {600, 0, 320, 0, 600, 240, 0, 0, 1}};
    const double Rm[C][9] = {{1, 0, 0, 0, 1, 0, 0, 0, 1}, {1, 0, 0, 0, 1, 0, 0, 0, 1}};
    const double t[C][3] = {{0, 0, 0}, {-0.5, 0, 0}};
    rc = mocap_set_cameras(ctx, &K[0][0], &Rm[0][0], &t[0][0]);
    if (rc != MOCAP_OK) { fprintf(stderr, "%s\n", mocap_last_error(ctx)); return 1; }
    unsigned char* frames = NULL;
    double *obj = NULL, *err = NULL;
    int32_t *n_obj = NULL, *flags = NULL;
    mocap_host_alloc((void**)&frames, (uint64_t)N * C * W * H);
    mocap_host_alloc((void**)&obj, sizeof(double) * N * R * 3);
    mocap_host_alloc((void**)&err, sizeof(double) * N * R);
    mocap_host_alloc((void**)&n_obj, sizeof(int32_t) * N);
    mocap_host_alloc((void**)&flags, sizeof(int32_t) * N);
    memset(frames, 0, (size_t)N * C * W * H);
    for (int s = 0; s < N; ++s)
        for (int c = 0; c < C; ++c)          /* a point at (0.1, 0.05, 3): u = 340 - 100 c, v = 250 */
            for (int y = 248; y < 253; ++y)
                for (int x = 338 - 100 * c; x < 343 - 100 * c; ++x) frames[((size_t)(s * C + c) * H + y) * W + x] = 255;
    rc = mocap_pipeline_host(ctx, frames, N, 1, 51, obj, err, n_obj, flags);
    if (rc != MOCAP_OK) { fprintf(stderr, "%s\n", mocap_last_error(ctx)); return 1; }
    for (int s = 0; s < N; ++s)
        for (int k = 0; k < n_obj[s]; ++k)
            printf("frame-set %d point %d: %.4f %.4f %.4f  err %.4f px^2\n", s, k, obj[(s * R + k) * 3], obj[(s * R + k) * 3 + 1],
                   obj[(s * R + k) * 3 + 2], err[s * R + k]);
    mocap_host_free(frames); mocap_host_free(obj); mocap_host_free(err); mocap_host_free(n_obj); mocap_host_free(flags);
    mocap_destroy(ctx);
    return 0;
}
