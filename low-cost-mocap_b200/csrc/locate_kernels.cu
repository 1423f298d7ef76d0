// SURVEY.md section 8(f) "next" #3: marker triplets -> drone pose, per frame-set.
//
// Replaces locate_objects (reference computer_code/api/helpers.py:424-480): pairwise distance matrix of
// the frame's 3D points; a point i with >= 2 neighbours at 0.095 +- 0.025 looks, in the reference's
// cartesian-product order, for the first ordered pair (a, b) of those neighbours that is 0.15 +- 0.025
// apart; the object sits midway between a and b, heads along a - b (folded into [-pi/2, pi/2], sign
// flipped), its error is the mean of the three reprojection errors and its droneIndex comes from the
// side of the axis point i lies on.  i, a and b all go on the reference's already_matched_points list, but that list
// only screens the OUTER index: a matched point is skipped as a later i, yet stays available as a / b of another
// i (so up to one object per point can come out; the mirror sizes max_objects accordingly).
// The greedy scan is order dependent, so one thread owns one frame-set and walks it sequentially;
// frame-sets are independent.
#include "common.cuh"
#include "geom.cuh"

#define LOC_D1 0.095
#define LOC_D2 0.15
#define LOC_TOL 0.025

__device__ __forceinline__ double pt_dist(const double* __restrict__ P, int a, int b) {
    const double dx = DSUB(P[3 * a], P[3 * b]), dy = DSUB(P[3 * a + 1], P[3 * b + 1]), dz = DSUB(P[3 * a + 2], P[3 * b + 2]);
    return sqrt(DADD(DADD(DMUL(dx, dx), DMUL(dy, dy)), DMUL(dz, dz)));      // np.sqrt(np.sum((a-b)**2)), helpers.py:434,446
}

__global__ void __launch_bounds__(128)
k_locate_objects(const double* __restrict__ obj, const double* __restrict__ err, const int32_t* __restrict__ n_obj,
                 int n_sets, int RMAX, int max_objects, double* __restrict__ out /*[n_sets][max_objects][5]*/,
                 int32_t* __restrict__ drone_index, int32_t* __restrict__ n_out) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_sets) return;
    const double* P = obj + (size_t)s * RMAX * 3;
    const double* E = err + (size_t)s * RMAX;
    const int K = min(n_obj[s], RMAX);
    // bit i of used: point i already produced an object (helpers.py:436-437)
    unsigned long long used_lo = 0ull, used_hi = 0ull;
    int found = 0;
    for (int i = 0; i < K; ++i) {
        if ((i < 64 ? (used_lo >> i) : (used_hi >> (i - 64))) & 1ull) continue;
        int cnt = 0;
        for (int j = 0; j < K; ++j) cnt += (fabs(DSUB(pt_dist(P, i, j), LOC_D1)) < LOC_TOL) ? 1 : 0;
        if (cnt < 2) continue;
        bool done = false;
        for (int a = 0; a < K && !done; ++a) {                  // cartesian_product(matches, matches), helpers.py:443
            if (!(fabs(DSUB(pt_dist(P, i, a), LOC_D1)) < LOC_TOL)) continue;
            for (int b = 0; b < K; ++b) {
                if (!(fabs(DSUB(pt_dist(P, i, b), LOC_D1)) < LOC_TOL)) continue;
                if (fabs(DSUB(pt_dist(P, a, b), LOC_D2)) > LOC_TOL) continue;
                // helpers.py:453-455: i, a, b all go on the list that only ever screens i
                if (i < 64) used_lo |= 1ull << i; else used_hi |= 1ull << (i - 64);
                if (a < 64) used_lo |= 1ull << a; else used_hi |= 1ull << (a - 64);
                if (b < 64) used_lo |= 1ull << b; else used_hi |= 1ull << (b - 64);
                const double lx = DADD(P[3 * a], P[3 * b]) / 2.0, ly = DADD(P[3 * a + 1], P[3 * b + 1]) / 2.0, lz = DADD(P[3 * a + 2], P[3 * b + 2]) / 2.0;
                const double e = DADD(DADD(E[i], E[a]), E[b]) / 3.0;
                double hx = DSUB(P[3 * a], P[3 * b]), hy = DSUB(P[3 * a + 1], P[3 * b + 1]), hz = DSUB(P[3 * a + 2], P[3 * b + 2]);
                const double nrm = sqrt(DADD(DADD(DMUL(hx, hx), DMUL(hy, hy)), DMUL(hz, hz)));
                hx /= nrm; hy /= nrm;
                double heading = atan2(hy, hx);
                const double pi = 3.141592653589793;
                if (heading > pi / 2) heading = heading - pi;
                if (heading < -pi / 2) heading = heading + pi;
                if (found < max_objects) {
                    double* o = out + ((size_t)s * max_objects + found) * 5;
                    o[0] = lx; o[1] = ly; o[2] = lz; o[3] = -heading; o[4] = e;
                    drone_index[(size_t)s * max_objects + found] = (DSUB(P[3 * i + 1], ly) > 0.0) ? 0 : 1;
                }
                ++found;
                done = true;
                break;
            }
        }
    }
    n_out[s] = min(found, max_objects);
}

int launch_locate(mocap_ctx* ctx, const double* obj, const double* err, const int32_t* n_obj, int n_sets,
                  int max_objects, double* out, int32_t* drone_index, int32_t* n_out) {
    if (n_sets <= 0) return MOCAP_OK;
    k_locate_objects<<<(n_sets + 127) / 128, 128, 0, ctx->stream>>>(obj, err, n_obj, n_sets, ctx->cfg.max_roots, max_objects,
                                                                      out, drone_index, n_out);
    CUDA_TRY(ctx, cudaGetLastError());
    ctx->launches += 1;
    return MOCAP_OK;
}
