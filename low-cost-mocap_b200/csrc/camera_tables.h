// Host-side construction of the camera tables the matcher and the triangulation kernels read
// (mocap_set_cameras).  Plain host code in a header of its own so that the host-run checks of the device code
// (tests/hostcheck) build the very same tables.
#pragma once
#include <math.h>
#include <string.h>
#include "common.cuh"

static double det3(const double a[3][3]) {
    return a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1]) - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0]) +
           a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]);
}
static double det4(const double m[4][4]) {       // cofactor expansion along row 0, same order as oracle/sfm_shim.py
    double d = 0.0;
    for (int j = 0; j < 4; ++j) {
        double minor[3][3];
        for (int r = 1; r < 4; ++r) {
            int cc = 0;
            for (int c = 0; c < 4; ++c) if (c != j) minor[r - 1][cc++] = m[r][c];
        }
        const double term = m[0][j] * det3(minor);
        d = (j % 2 == 0) ? d + term : d - term;
    }
    return d;
}
// libmv FundamentalFromProjections (cv.sfm.fundamentalFromProjections, helpers.py:362)
static void fundamental_from_projections(const double* P1, const double* P2, double* F) {
    static const int pair[3][2] = {{1, 2}, {2, 0}, {0, 1}};
    for (int i = 0; i < 3; ++i)
        for (int j = 0; j < 3; ++j) {
            double m[4][4];
            for (int c = 0; c < 4; ++c) {
                m[0][c] = P1[pair[j][0] * 4 + c]; m[1][c] = P1[pair[j][1] * 4 + c];
                m[2][c] = P2[pair[i][0] * 4 + c]; m[3][c] = P2[pair[i][1] * 4 + c];
            }
            F[i * 3 + j] = det4(m);
        }
}

// K [C][9], R [C][9], t [C][3] (row-major) -> Pkc, F, R, t, the four intrinsics cv.projectPoints reads, Kmat
static void build_camera_tables(CameraTables& T, int C, const double* K, const double* R, const double* t) {
    for (int c = 0; c < C; ++c) {
        memcpy(T.R[c], R + 9 * c, 9 * sizeof(double));
        memcpy(T.t[c], t + 3 * c, 3 * sizeof(double));
        memcpy(T.Kmat[c], K + 9 * c, 9 * sizeof(double));
        T.fx[c] = K[9 * c + 0]; T.fy[c] = K[9 * c + 4]; T.cx[c] = K[9 * c + 2]; T.cy[c] = K[9 * c + 5];
    }
    for (int k = 0; k < C; ++k)
        for (int c = 0; c < C; ++c)
            for (int i = 0; i < 3; ++i)
                for (int j = 0; j < 4; ++j) {
                    // K_k @ [R_c | t_c] the way a BLAS dgemm micro-kernel sums it (fused multiply-adds over k)
                    double acc = K[9 * k + 3 * i + 0] * (j < 3 ? R[9 * c + 0 * 3 + j] : t[3 * c + 0]);
                    acc = fma(K[9 * k + 3 * i + 1], (j < 3 ? R[9 * c + 1 * 3 + j] : t[3 * c + 1]), acc);
                    acc = fma(K[9 * k + 3 * i + 2], (j < 3 ? R[9 * c + 2 * 3 + j] : t[3 * c + 2]), acc);
                    T.Pkc[k][c][4 * i + j] = acc;
                }
    for (int r = 0; r < C; ++r)
        for (int c = 0; c < C; ++c) fundamental_from_projections(T.Pkc[r][r], T.Pkc[c][c], T.F[r][c]);
    T.n_cam = C;
}
