// Test-only CPU stand-in for the GPU side of S4: evaluates residuals / finite-difference
// Jacobian / robust-scaled normal equations with the same HD geometry code (geom.cuh) and feeds
// the product's optimiser control (trf_core.h), so the trust-region logic can be compared with
// scipy on a machine without a GPU.  NOT part of libmocap_b200.so.
#include <stdint.h>
#include <string.h>
#include <vector>
#include "../../low-cost-mocap_b200/csrc/geom.cuh"
#include "../../low-cost-mocap_b200/csrc/trf_core.h"

namespace {
struct CpuBA : trf::Problem {
    const double* obs; const uint8_t* mask; int m, C; const double* K;   // K [C][9]
    std::vector<int> live_idx;
    std::vector<int> valid;     // indices of points with >= 2 views
    int jac_mode = 0;           // 0: FD on float32 residuals (reference), 1: FD on float64 residuals

    void poses(const double* x, std::vector<double>& Rt) const {
        Rt.assign((size_t)C * 12, 0.0);
        Rt[0] = Rt[5] = Rt[10] = 1.0;
        for (int c = 1; c < C; ++c) {
            const double* q = x + 1 + 7 * (c - 1);
            double R[9];
            trf::rotvec_to_matrix(q + 1, R);
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) Rt[c * 12 + 4 * i + j] = R[3 * i + j]; Rt[c * 12 + 4 * i + 3] = q[4 + i]; }
        }
    }
    double residual(int f, const std::vector<double>& Rt) const {
        const double* o = obs + (size_t)f * C * 2; const uint8_t* mk = mask + (size_t)f * C;
        Sym4 B; sym4_zero(B);
        int k = 0;
        double P[12];
        for (int c = 0; c < C; ++c) if (mk[c]) {
            for (int i = 0; i < 3; ++i) for (int j = 0; j < 4; ++j) {
                double acc = K[9 * k + 3 * i] * Rt[c * 12 + j];
                acc = fma(K[9 * k + 3 * i + 1], Rt[c * 12 + 4 + j], acc);
                acc = fma(K[9 * k + 3 * i + 2], Rt[c * 12 + 8 + j], acc);
                P[4 * i + j] = acc;
            }
            dlt_add_view(B, P, o[2 * c], o[2 * c + 1]); ++k;
        }
        double X[3]; dlt_solve(B, X);
        double sq[64]; k = 0;
        for (int c = 0; c < C; ++c) if (mk[c]) {
            double R[9], t[3];
            for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) R[3 * i + j] = Rt[c * 12 + 4 * i + j]; t[i] = Rt[c * 12 + 4 * i + 3]; }
            float u, v;
            project_like_cv(R, t, K[9 * k], K[9 * k + 4], K[9 * k + 2], K[9 * k + 5], X, u, v);
            const double dx = o[2 * c] - (double)u, dy = o[2 * c + 1] - (double)v;
            sq[2 * k] = dx * dx; sq[2 * k + 1] = dy * dy; ++k;
        }
        return mean_like_numpy(sq, 2 * k, false);
    }
    void residuals(const double* x, std::vector<float>& f, std::vector<double>* fd = nullptr) const {
        std::vector<double> Rt; poses(x, Rt);
        f.resize(valid.size());
        if (fd) fd->resize(valid.size());
        for (size_t i = 0; i < valid.size(); ++i) { const double r = residual(valid[i], Rt); f[i] = (float)r; if (fd) (*fd)[i] = r; }
    }
    static double cost_of(const std::vector<float>& f, int* finite) {
        double c = 0.0; *finite = 1;
        for (float v : f) { if (!isfinite(v)) *finite = 0; const float z = v * v; c += (double)log1pf(z); }
        return 0.5 * c;
    }
    int trial_cost(const double* x, double* cost, int* finite) override {
        std::vector<float> f; residuals(x, f); *cost = cost_of(f, finite); return 0;
    }
    int linearize(const double* x, double* A, double* g, double* cost, int* finite) override {
        std::vector<float> f0; std::vector<double> f0d; residuals(x, f0, &f0d);
        *cost = cost_of(f0, finite);
        const int n = n_live, mm = (int)f0.size();
        std::vector<double> J((size_t)mm * n);
        std::vector<double> xp(x, x + n_full);
        for (int j = 0; j < n; ++j) {
            const int idx = live[j];
            const double x0 = x[idx];
            const double h = 1.4901161193847656e-08 * (x0 >= 0 ? 1.0 : -1.0) * fmax(1.0, fabs(x0));
            xp[idx] = x0 + h;
            const double dx = xp[idx] - x0;
            std::vector<float> f1; std::vector<double> f1d; residuals(xp.data(), f1, &f1d);
            for (int i = 0; i < mm; ++i)
                J[(size_t)i * n + j] = jac_mode == 0 ? (double)(float)(f1[i] - f0[i]) / dx : (f1d[i] - f0d[i]) / dx;
            xp[idx] = x0;
        }
        std::vector<double> fs(mm);
        for (int i = 0; i < mm; ++i) {
            const float fv = f0[i], z = fv * fv, t = 1.0f + z;
            const float rho1 = 1.0f / t, rho2 = -1.0f / (t * t);
            double js = (double)rho1 + 2.0 * (double)rho2 * (double)z;
            if (js < 2.220446049250313e-16) js = 2.220446049250313e-16;
            js = sqrt(js);
            fs[i] = (double)(float)((double)fv * ((double)rho1 / js));
            for (int j = 0; j < n; ++j) J[(size_t)i * n + j] *= js;
        }
        for (int a = 0; a < n; ++a) {
            for (int b = 0; b < n; ++b) { double s = 0; for (int i = 0; i < mm; ++i) s += J[(size_t)i * n + a] * J[(size_t)i * n + b]; A[(size_t)a * n + b] = s; }
            double s = 0; for (int i = 0; i < mm; ++i) s += J[(size_t)i * n + a] * fs[i]; g[a] = s;
        }
        return 0;
    }
};
}  // namespace

extern "C" int hc_bundle_adjust(const double* obs, const uint8_t* mask, int m, int C, const double* K,
                                double* R, double* t, double ftol, int max_nfev, double* report /*[7]*/, int jac_mode) {
    CpuBA p;
    p.jac_mode = jac_mode;
    p.obs = obs; p.mask = mask; p.m = m; p.C = C; p.K = K;
    for (int f = 0; f < m; ++f) { int nv = 0; for (int c = 0; c < C; ++c) nv += mask[(size_t)f * C + c] ? 1 : 0; if (nv > 1) p.valid.push_back(f); }
    p.n_full = 1 + 7 * (C - 1);
    for (int c = 1; c < C; ++c) for (int q = 1; q < 7; ++q) p.live_idx.push_back(1 + 7 * (c - 1) + q);
    p.n_live = (int)p.live_idx.size();
    p.live = p.live_idx.data();
    std::vector<double> x(p.n_full);
    x[0] = K[0];
    for (int c = 1; c < C; ++c) {
        double* q = x.data() + 1 + 7 * (c - 1);
        q[0] = K[9 * (c - 1)];                         // helpers.py:281-282 reads K[i] with i enumerating poses[1:]
        trf::matrix_to_rotvec(R + 9 * c, q + 1);
        q[4] = t[3 * c]; q[5] = t[3 * c + 1]; q[6] = t[3 * c + 2];
    }
    trf::Options opt{ftol, 1e-8, 1e-8, max_nfev, 0.0};
    trf::Report rep{};
    int st = trf::minimize(p, x.data(), opt, rep);
    for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0;
    t[0] = t[1] = t[2] = 0.0;
    for (int c = 1; c < C; ++c) {
        const double* q = x.data() + 1 + 7 * (c - 1);
        trf::rotvec_to_matrix(q + 1, R + 9 * c);
        t[3 * c] = q[4]; t[3 * c + 1] = q[5]; t[3 * c + 2] = q[6];
    }
    report[0] = rep.cost_initial; report[1] = rep.cost_final; report[2] = rep.optimality;
    report[3] = rep.n_iterations; report[4] = rep.n_fev; report[5] = rep.status; report[6] = rep.n_jev;
    return st;
}

extern "C" void hc_rotvec_roundtrip(const double* R, double* rv, double* R2) {
    trf::matrix_to_rotvec(R, rv);
    trf::rotvec_to_matrix(rv, R2);
}

// ---- generic check of the trust-region control against scipy: an exponential fit with a dead
//      parameter (rank-deficient Jacobian, like the reference's dead focal entries), Cauchy loss
namespace {
struct ExpFit : trf::Problem {
    std::vector<double> tt, yy;
    std::vector<int> live_idx;
    void resid(const double* x, std::vector<double>& f) const {
        f.resize(tt.size());
        for (size_t i = 0; i < tt.size(); ++i) f[i] = x[0] * exp(-x[1] * tt[i]) + x[2] - yy[i];     // x[3] is dead
    }
    static double cost_of(const std::vector<double>& f) { double c = 0; for (double v : f) c += log1p(v * v); return 0.5 * c; }
    int trial_cost(const double* x, double* cost, int* finite) override {
        std::vector<double> f; resid(x, f); *cost = cost_of(f); *finite = 1; return 0;
    }
    int linearize(const double* x, double* A, double* g, double* cost, int* finite) override {
        std::vector<double> f0; resid(x, f0); *cost = cost_of(f0); *finite = 1;
        const int n = n_live, m = (int)f0.size();
        std::vector<double> J((size_t)m * n), xp(x, x + n_full), fs(m);
        for (int j = 0; j < n; ++j) {
            const int idx = live[j]; const double x0 = x[idx];
            const double h = 1.4901161193847656e-08 * (x0 >= 0 ? 1.0 : -1.0) * fmax(1.0, fabs(x0));
            xp[idx] = x0 + h; const double dx = xp[idx] - x0;
            std::vector<double> f1; resid(xp.data(), f1);
            for (int i = 0; i < m; ++i) J[(size_t)i * n + j] = (f1[i] - f0[i]) / dx;
            xp[idx] = x0;
        }
        for (int i = 0; i < m; ++i) {
            const double z = f0[i] * f0[i], t = 1 + z, rho1 = 1 / t, rho2 = -1 / (t * t);
            double js = rho1 + 2 * rho2 * z; if (js < 2.220446049250313e-16) js = 2.220446049250313e-16; js = sqrt(js);
            fs[i] = f0[i] * rho1 / js;
            for (int j = 0; j < n; ++j) J[(size_t)i * n + j] *= js;
        }
        for (int a = 0; a < n; ++a) {
            for (int b = 0; b < n; ++b) { double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + a] * J[(size_t)i * n + b]; A[(size_t)a * n + b] = s; }
            double s = 0; for (int i = 0; i < m; ++i) s += J[(size_t)i * n + a] * fs[i]; g[a] = s;
        }
        return 0;
    }
};
}  // namespace

extern "C" int hc_trf_expfit(const double* t, const double* y, int m, double* x /*[4] in/out*/, double ftol, double* report /*[7]*/) {
    ExpFit p;
    p.tt.assign(t, t + m); p.yy.assign(y, y + m);
    p.n_full = 4; p.live_idx = {0, 1, 2}; p.n_live = 3; p.live = p.live_idx.data();
    trf::Options opt{ftol, 1e-8, 1e-8, 0, 0.0};
    trf::Report rep{};
    const int st = trf::minimize(p, x, opt, rep);
    report[0] = rep.cost_initial; report[1] = rep.cost_final; report[2] = rep.optimality;
    report[3] = rep.n_iterations; report[4] = rep.n_fev; report[5] = rep.status; report[6] = rep.n_jev;
    return st;
}
