// Test-only host build of the HD geometry functions in low-cost-mocap_b200/csrc/geom.cuh, so
// that the DLT / Jacobi / projection arithmetic can be checked against the golden vectors on a
// machine without a GPU.  NOT part of libmocap_b200.so and never used by the product path.
#include <stdint.h>
#include "../../low-cost-mocap_b200/csrc/geom.cuh"

extern "C" {
// obs [n][C][2], mask [n][C], P [C][C][12] (Pkc), R [C][9], t [C][3], K4 [C][4] = fx fy cx cy
void hc_triangulate(const double* obs, const uint8_t* mask, int n, int C, const double* Pkc,
                    const double* R, const double* t, const double* K4, double* X, double* err, uint8_t* valid) {
    for (int f = 0; f < n; ++f) {
        const double* o = obs + (size_t)f * C * 2;
        const uint8_t* m = mask + (size_t)f * C;
        int nv = 0;
        for (int c = 0; c < C; ++c) nv += m[c] ? 1 : 0;
        valid[f] = nv > 1;
        if (nv <= 1) continue;
        Sym4 B;
        sym4_zero(B);
        int k = 0;
        for (int c = 0; c < C; ++c)
            if (m[c]) { dlt_add_view(B, Pkc + ((size_t)k * C + c) * 12, o[2 * c], o[2 * c + 1]); ++k; }
        double Xf[3];
        dlt_solve(B, Xf);
        X[3 * f] = Xf[0]; X[3 * f + 1] = Xf[1]; X[3 * f + 2] = Xf[2];
        double sq[64];
        k = 0;
        for (int c = 0; c < C; ++c)
            if (m[c]) {
                float u, v;
                project_like_cv(R + 9 * c, t + 3 * c, K4[4 * k], K4[4 * k + 1], K4[4 * k + 2], K4[4 * k + 3], Xf, u, v);
                const double dx = o[2 * c] - (double)u, dy = o[2 * c + 1] - (double)v;
                sq[2 * k] = dx * dx; sq[2 * k + 1] = dy * dy;
                ++k;
            }
        err[f] = mean_like_numpy(sq, 2 * nv, false);
    }
}
}

// how often the fast null-vector path settles, and how far it is from the Jacobi result
extern "C" void hc_null_vector_compare(const double* Bs, int n, double* max_rel_diff, int* n_fallback) {
    double worst = 0.0; int fb = 0;
    for (int i = 0; i < n; ++i) {
        Sym4 B; for (int k = 0; k < 10; ++k) B.v[k] = Bs[10 * i + k];
        double a[4], b[4];
        sym4_null_vector(B, a);
        if (!sym4_null_vector_invit(B, b)) { ++fb; continue; }
        // compare dehomogenised
        for (int k = 0; k < 3; ++k) {
            const double xa = a[k] / a[3], xb = b[k] / b[3];
            const double d = fabs(xa - xb) / fmax(1.0, fabs(xa));
            if (d > worst) worst = d;
        }
    }
    *max_rel_diff = worst; *n_fallback = fb;
}
