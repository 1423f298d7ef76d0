"""Restatement of the three ``cv2.sfm`` (opencv_contrib / libmv) functions the
reference calls (helpers.py:362, index.py:247-248).  opencv_contrib is not
vendored by the reference, its version is unpinned (README.md:17 "compile
OpenCV from source") and it is not installed here, so parity at this boundary
is UNPINNED: this file restates the published libmv algorithm
(libmv/multiview/fundamental.cc ``FundamentalFromProjections``,
``EssentialFromFundamental``, ``MotionFromEssential``).
"""
import types

import numpy as np


def _det4(rows):
    m = np.asarray(rows, dtype=np.float64)
    # cofactor expansion along the first row (the CUDA host side uses the same
    # expression order, so the two agree to the last bit on the same inputs)
    def det3(a):
        return (a[0][0] * (a[1][1] * a[2][2] - a[1][2] * a[2][1])
                - a[0][1] * (a[1][0] * a[2][2] - a[1][2] * a[2][0])
                + a[0][2] * (a[1][0] * a[2][1] - a[1][1] * a[2][0]))
    d = 0.0
    for j in range(4):
        minor = [[m[r][c] for c in range(4) if c != j] for r in range(1, 4)]
        term = m[0][j] * det3(minor)
        d = d + term if j % 2 == 0 else d - term
    return d


def fundamentalFromProjections(P1, P2):
    """F with x2ᵀ F x1 = 0:  F[i][j] = det([X_j ; Y_i])  (libmv)."""
    P1 = np.asarray(P1, dtype=np.float64)
    P2 = np.asarray(P2, dtype=np.float64)
    X = [P1[[1, 2]], P1[[2, 0]], P1[[0, 1]]]
    Y = [P2[[1, 2]], P2[[2, 0]], P2[[0, 1]]]
    F = np.zeros((3, 3))
    for i in range(3):
        for j in range(3):
            F[i, j] = _det4(np.vstack([X[j], Y[i]]))
    return F


def essentialFromFundamental(F, K1, K2):
    return np.asarray(K2, dtype=np.float64).T @ np.asarray(F, dtype=np.float64) @ np.asarray(K1, dtype=np.float64)


def motionFromEssential(E):
    U, s, Vt = np.linalg.svd(np.asarray(E, dtype=np.float64))
    if np.linalg.det(U) < 0:
        U[:, 2] *= -1
    if np.linalg.det(Vt) < 0:
        Vt[2, :] *= -1
    W = np.array([[0.0, -1.0, 0.0], [1.0, 0.0, 0.0], [0.0, 0.0, 1.0]])
    UWVt = U @ W @ Vt
    UWtVt = U @ W.T @ Vt
    u3 = U[:, 2].reshape(3, 1)
    return [UWVt, UWVt, UWtVt, UWtVt], [u3, -u3, u3, -u3]


def namespace():
    return types.SimpleNamespace(
        fundamentalFromProjections=fundamentalFromProjections,
        essentialFromFundamental=essentialFromFundamental,
        motionFromEssential=motionFromEssential,
    )
