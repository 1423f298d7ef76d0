// Host-side optimiser control of S4: the trust-region-reflective iteration that
// scipy.optimize.least_squares runs for the reference's bundle_adjustment
// (computer_code/api/helpers.py:287-289 -> method 'trf', loss 'cauchy', tr_solver 'exact',
// x_scale 1, no bounds; scipy/optimize/_lsq/trf.py trf_no_bounds, common.py
// solve_lsq_trust_region / update_tr_radius / check_termination -- SURVEY.md Appendix A.3).
//
// Everything that touches the observations (residuals, finite-difference Jacobian, robust
// scaling, the reduction to normal equations) runs on the GPU (ba.cu).  What is left here is
// the O(n^3) dense step on the n <= 90 live parameters and the accept/reject logic:
//   scipy takes the SVD J = U S V^T and works with s and suf = s * (U^T f).  Because
//   suf = V^T (J^T f) = V^T g and s^2 = eig(J^T J), the same quantities come from the
//   symmetric eigen-decomposition of A = J^T J, which is all the device has to hand over.
// The reference's parameter vector carries C dead focal entries (helpers.py:267-270); they make
// scipy's Jacobian rank deficient, so scipy never takes its Gauss-Newton branch -- mirrored here
// by always solving the regularised problem (full_rank == false).
#pragma once
#include <math.h>
#include <stddef.h>
#include <vector>
#ifdef TRF_TRACE
#include <stdio.h>
#endif

namespace trf {

struct Options { double ftol, xtol, gtol; int max_nfev; double initial_radius; };   // initial_radius 0: scipy's ||x0||
struct Report {
    double cost_initial, cost_final, optimality;
    int n_iterations, n_fev, n_jev, status;
};

// The problem as the GPU presents it.
struct Problem {
    int n_full;                 // length of x (reference layout: 1 + 7 (C-1))
    int n_live;                 // live parameters (6 (C-1))
    const int* live;            // live[j] = index in x of live parameter j
    // at an ACCEPTED point: robust-scaled normal equations A (n_live x n_live, row-major),
    // g (n_live), cost = 0.5 sum rho(f^2).  Returns 0 on success; *finite = 0 if a residual is not finite.
    virtual int linearize(const double* x, double* A, double* g, double* cost, int* finite) = 0;
    // at a TRIAL point: cost only.
    virtual int trial_cost(const double* x, double* cost, int* finite) = 0;
    virtual ~Problem() {}
};

// ---- symmetric eigen-decomposition: Householder tridiagonalisation + implicit QL (EISPACK
//      tred2/tql2 scheme).  a: n x n row-major symmetric in, eigenvectors (columns) out; d: eigenvalues.
inline bool sym_eig(int n, std::vector<double>& a, std::vector<double>& d) {
    std::vector<double> e(n, 0.0);
    d.assign(n, 0.0);
    auto A = [&](int i, int j) -> double& { return a[(size_t)i * n + j]; };
    for (int i = n - 1; i > 0; --i) {
        const int l = i - 1;
        double h = 0.0, scale = 0.0;
        if (l > 0) {
            for (int k = 0; k <= l; ++k) scale += fabs(A(i, k));
            if (scale == 0.0) e[i] = A(i, l);
            else {
                for (int k = 0; k <= l; ++k) { A(i, k) /= scale; h += A(i, k) * A(i, k); }
                double f = A(i, l);
                double g = f >= 0.0 ? -sqrt(h) : sqrt(h);
                e[i] = scale * g;
                h -= f * g;
                A(i, l) = f - g;
                f = 0.0;
                for (int j = 0; j <= l; ++j) {
                    A(j, i) = A(i, j) / h;
                    g = 0.0;
                    for (int k = 0; k <= j; ++k) g += A(j, k) * A(i, k);
                    for (int k = j + 1; k <= l; ++k) g += A(k, j) * A(i, k);
                    e[j] = g / h;
                    f += e[j] * A(i, j);
                }
                const double hh = f / (h + h);
                for (int j = 0; j <= l; ++j) {
                    f = A(i, j);
                    e[j] = g = e[j] - hh * f;
                    for (int k = 0; k <= j; ++k) A(j, k) -= f * e[k] + g * A(i, k);
                }
            }
        } else e[i] = A(i, l);
        d[i] = h;
    }
    d[0] = 0.0; e[0] = 0.0;
    for (int i = 0; i < n; ++i) {
        const int l = i - 1;
        if (d[i] != 0.0) {
            for (int j = 0; j <= l; ++j) {
                double g = 0.0;
                for (int k = 0; k <= l; ++k) g += A(i, k) * A(k, j);
                for (int k = 0; k <= l; ++k) A(k, j) -= g * A(k, i);
            }
        }
        d[i] = A(i, i);
        A(i, i) = 1.0;
        for (int j = 0; j <= l; ++j) A(j, i) = A(i, j) = 0.0;
    }
    for (int i = 1; i < n; ++i) e[i - 1] = e[i];
    e[n - 1] = 0.0;
    for (int l = 0; l < n; ++l) {
        int iter = 0, m;
        do {
            for (m = l; m < n - 1; ++m) {
                const double dd = fabs(d[m]) + fabs(d[m + 1]);
                if (fabs(e[m]) + dd == dd) break;
            }
            if (m != l) {
                if (iter++ == 60) return false;
                double g = (d[l + 1] - d[l]) / (2.0 * e[l]);
                double r = hypot(g, 1.0);
                g = d[m] - d[l] + e[l] / (g + (g >= 0.0 ? fabs(r) : -fabs(r)));
                double s = 1.0, c = 1.0, p = 0.0;
                int i;
                for (i = m - 1; i >= l; --i) {
                    double f = s * e[i];
                    const double b = c * e[i];
                    e[i + 1] = (r = hypot(f, g));
                    if (r == 0.0) { d[i + 1] -= p; e[m] = 0.0; break; }
                    s = f / r; c = g / r;
                    g = d[i + 1] - p;
                    r = (d[i] - g) * s + 2.0 * c * b;
                    d[i + 1] = g + (p = s * r);
                    g = c * r - b;
                    for (int k = 0; k < n; ++k) {
                        f = A(k, i + 1);
                        A(k, i + 1) = s * A(k, i) + c * f;
                        A(k, i) = c * A(k, i) - s * f;
                    }
                }
                if (r == 0.0 && i >= l) continue;
                d[l] -= p; e[l] = g; e[m] = 0.0;
            }
        } while (m != l);
    }
    return true;
}

inline double norm2(const double* v, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) s += v[i] * v[i];
    return sqrt(s);
}

// common.py solve_lsq_trust_region with full_rank == false.  lam = s^2, suf = V^T g.
inline void solve_tr(int n, const std::vector<double>& lam, const std::vector<double>& suf,
                     const std::vector<double>& V, double Delta, double& alpha, std::vector<double>& p) {
    auto phi_and_derivative = [&](double al, double& phi, double& phi_prime) {
        double pn2 = 0.0, acc = 0.0;
        for (int i = 0; i < n; ++i) {
            const double den = lam[i] + al;
            const double q = suf[i] / den;
            pn2 += q * q;
            acc += suf[i] * suf[i] / (den * den * den);
        }
        const double p_norm = sqrt(pn2);
        phi = p_norm - Delta;
        phi_prime = -acc / p_norm;
    };
    double alpha_upper = norm2(suf.data(), n) / Delta;
    double alpha_lower = 0.0;
    if (alpha == 0.0) alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
    for (int it = 0; it < 10; ++it) {
        if (alpha < alpha_lower || alpha > alpha_upper)
            alpha = fmax(0.001 * alpha_upper, sqrt(alpha_lower * alpha_upper));
        double phi, phi_prime;
        phi_and_derivative(alpha, phi, phi_prime);
        if (phi < 0) alpha_upper = alpha;
        const double ratio = phi / phi_prime;
        alpha_lower = fmax(alpha_lower, alpha - ratio);
        alpha -= (phi + Delta) * ratio / Delta;
        if (fabs(phi) < 0.01 * Delta) break;
    }
    p.assign(n, 0.0);
    for (int k = 0; k < n; ++k) {
        const double c = -suf[k] / (lam[k] + alpha);
        for (int i = 0; i < n; ++i) p[i] += V[(size_t)i * n + k] * c;
    }
    const double pn = norm2(p.data(), n);
    if (pn > 0.0) for (int i = 0; i < n; ++i) p[i] *= Delta / pn;
}

// trf.py trf_no_bounds.  x: in/out (n_full).  Returns 0 or the first non-zero status of a callback.
inline int minimize(Problem& prob, double* x, const Options& opt, Report& rep) {
    const int n = prob.n_live, nf = prob.n_full;
    std::vector<double> A((size_t)n * n), g(n), V, lam, suf(n), p, x_new(nf), A_keep;
    double cost = 0.0;
    int finite = 1;
    int st = prob.linearize(x, A.data(), g.data(), &cost, &finite);
    if (st) return st;
    rep.cost_initial = cost;
    int nfev = 1, njev = 1;
    if (!finite) { rep.status = -1; rep.cost_final = cost; rep.n_fev = nfev; rep.n_jev = njev; rep.n_iterations = 0; rep.optimality = 0; return 0; }
    double Delta = opt.initial_radius > 0.0 ? opt.initial_radius : norm2(x, nf);
    if (Delta == 0.0) Delta = 1.0;
    const int max_nfev = opt.max_nfev > 0 ? opt.max_nfev : nf * 100;
    double alpha = 0.0;
    int termination = -99, iteration = 0;
    double g_norm = 0.0;
    while (true) {
        g_norm = 0.0;
        for (int i = 0; i < n; ++i) g_norm = fmax(g_norm, fabs(g[i]));
        if (g_norm < opt.gtol) termination = 1;
        if (termination != -99 || nfev == max_nfev) break;

        V = A;                                            // eigen-decomposition of J^T J
        if (!sym_eig(n, V, lam)) { termination = -2; break; }
        for (int k = 0; k < n; ++k) {
            if (lam[k] < 0.0) lam[k] = 0.0;              // rounding: J^T J is positive semi-definite
            double s = 0.0;
            for (int i = 0; i < n; ++i) s += V[(size_t)i * n + k] * g[i];
            suf[k] = s;
        }
        double actual_reduction = -1.0, cost_new = cost;
        while (actual_reduction <= 0 && nfev < max_nfev) {
            solve_tr(n, lam, suf, V, Delta, alpha, p);
            double q = 0.0, l = 0.0;                      // predicted = -(0.5 p^T A p + g^T p)
            for (int i = 0; i < n; ++i) {
                double r = 0.0;
                for (int j = 0; j < n; ++j) r += A[(size_t)i * n + j] * p[j];
                q += p[i] * r;
                l += p[i] * g[i];
            }
            const double predicted_reduction = -(0.5 * q + l);
            for (int i = 0; i < nf; ++i) x_new[i] = x[i];
            for (int j = 0; j < n; ++j) x_new[prob.live[j]] = x[prob.live[j]] + p[j];
            st = prob.trial_cost(x_new.data(), &cost_new, &finite);
            if (st) return st;
            ++nfev;
            const double step_h_norm = norm2(p.data(), n);
            if (!finite) { Delta = 0.25 * step_h_norm; continue; }
            actual_reduction = cost - cost_new;
            // update_tr_radius
            double ratio;
            if (predicted_reduction > 0) ratio = actual_reduction / predicted_reduction;
            else if (predicted_reduction == 0 && actual_reduction == 0) ratio = 1;
            else ratio = 0;
            double Delta_new = Delta;
            if (ratio < 0.25) Delta_new = 0.25 * step_h_norm;
            else if (ratio > 0.75 && step_h_norm > 0.95 * Delta) Delta_new = Delta * 2.0;
            // check_termination
            const double step_norm = step_h_norm, x_norm = norm2(x, nf);
            const bool ftol_ok = actual_reduction < opt.ftol * cost && ratio > 0.25;
            const bool xtol_ok = step_norm < opt.xtol * (opt.xtol + x_norm);
            if (ftol_ok && xtol_ok) termination = 4;
            else if (ftol_ok) termination = 2;
            else if (xtol_ok) termination = 3;
#ifdef TRF_TRACE
            fprintf(stderr, "  trial nfev=%d Delta=%.4e step=%.4e pred=%.4e act=%.4e ratio=%.3f cost_new=%.6e alpha=%.3e\n", nfev, Delta, step_h_norm, predicted_reduction, actual_reduction, ratio, cost_new, alpha);
#endif
            if (termination != -99) break;
            alpha *= Delta / Delta_new;
            Delta = Delta_new;
        }
        if (actual_reduction > 0) {
            for (int i = 0; i < nf; ++i) x[i] = x_new[i];
            st = prob.linearize(x, A.data(), g.data(), &cost, &finite);   // cost == cost_new (same residuals)
            if (st) return st;
            ++njev;
        }
        ++iteration;
#ifdef TRF_TRACE
        fprintf(stderr, "iter %d nfev=%d cost=%.6e\n", iteration, nfev, cost);
#endif
    }
    if (termination == -99) termination = 0;
    rep.cost_final = cost;
    rep.optimality = g_norm;
    rep.n_iterations = iteration;
    rep.n_fev = nfev;
    rep.n_jev = njev;
    rep.status = termination;
    return 0;
}

// ---- parameterisation helpers (scipy.spatial.transform.Rotation, helpers.py:247-262, 278-285) ----
// rotation vector -> matrix through the unit quaternion, as Rotation.from_rotvec(...).as_matrix()
inline void rotvec_to_matrix(const double rv[3], double R[9]) {
    const double angle = sqrt(rv[0] * rv[0] + rv[1] * rv[1] + rv[2] * rv[2]);
    double scale;
    if (angle <= 1e-3) {
        const double a2 = angle * angle;
        scale = 0.5 - a2 / 48.0 + a2 * a2 / 3840.0;
    } else scale = sin(angle / 2.0) / angle;
    double x = scale * rv[0], y = scale * rv[1], z = scale * rv[2], w = cos(angle / 2.0);
    const double nq = sqrt(x * x + y * y + z * z + w * w);
    x /= nq; y /= nq; z /= nq; w /= nq;
    const double x2 = x * x, y2 = y * y, z2 = z * z, w2 = w * w;
    const double xy = x * y, zw = z * w, xz = x * z, yw = y * w, yz = y * z, xw = x * w;
    R[0] = x2 - y2 - z2 + w2; R[3] = 2 * (xy + zw);       R[6] = 2 * (xz - yw);
    R[1] = 2 * (xy - zw);     R[4] = -x2 + y2 - z2 + w2;  R[7] = 2 * (yz + xw);
    R[2] = 2 * (xz + yw);     R[5] = 2 * (yz - xw);       R[8] = -x2 - y2 + z2 + w2;
}

// matrix -> rotation vector, as Rotation.from_matrix(R).as_rotvec()
inline void matrix_to_rotvec(const double R[9], double rv[3]) {
    const double m00 = R[0], m11 = R[4], m22 = R[8], tr = m00 + m11 + m22;
    double dec[4] = {m00, m11, m22, tr};
    int choice = 0;
    for (int i = 1; i < 4; ++i) if (dec[i] > dec[choice]) choice = i;
    double q[4];   // x y z w
    if (choice != 3) {
        const int i = choice, j = (i + 1) % 3, k = (j + 1) % 3;
        q[i] = 1 - dec[3] + 2 * R[i * 3 + i];
        q[j] = R[j * 3 + i] + R[i * 3 + j];
        q[k] = R[k * 3 + i] + R[i * 3 + k];
        q[3] = R[k * 3 + j] - R[j * 3 + k];
    } else {
        q[0] = R[2 * 3 + 1] - R[1 * 3 + 2];
        q[1] = R[0 * 3 + 2] - R[2 * 3 + 0];
        q[2] = R[1 * 3 + 0] - R[0 * 3 + 1];
        q[3] = 1 + dec[3];
    }
    const double nq = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2] + q[3] * q[3]);
    for (int i = 0; i < 4; ++i) q[i] /= nq;
    if (q[3] < 0) for (int i = 0; i < 4; ++i) q[i] = -q[i];
    const double sn = sqrt(q[0] * q[0] + q[1] * q[1] + q[2] * q[2]);
    const double angle = 2.0 * atan2(sn, q[3]);
    double scale;
    if (angle <= 1e-3) {
        const double a2 = angle * angle;
        scale = 2.0 + a2 / 12.0 + 7.0 * a2 * a2 / 2880.0;
    } else scale = angle / sin(angle / 2.0);
    rv[0] = scale * q[0]; rv[1] = scale * q[1]; rv[2] = scale * q[2];
}

}  // namespace trf
