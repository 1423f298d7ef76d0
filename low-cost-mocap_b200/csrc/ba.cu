// S4 placeholder (filled in by the bundle-adjustment milestone).
#include "common.cuh"
extern "C" {
void mocap_ba_default_options(mocap_ba_options* opt) {
    opt->ftol = 1e-2; opt->xtol = 1e-8; opt->gtol = 1e-8; opt->max_nfev = 0; opt->jacobian = 0;
}
int mocap_bundle_adjust_host(mocap_ctx* ctx, const double*, const uint8_t*, int, double*, double*,
                             const mocap_ba_options*, mocap_ba_report*) {
    return mocap_fail(ctx, MOCAP_ESTATE, "bundle adjustment not built yet");
}
int mocap_ba_residuals_host(mocap_ctx* ctx, const double*, const uint8_t*, int, const double*, const double*,
                            float*, uint8_t*, int*) {
    return mocap_fail(ctx, MOCAP_ESTATE, "bundle adjustment not built yet");
}
}
