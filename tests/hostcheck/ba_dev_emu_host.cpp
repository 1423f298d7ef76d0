// Test-only: the device-resident bundle adjustment k_ba_solve (csrc/ba_device.cuh: persistent grid, grid-wide
// barriers, tile-wise accumulation of the normal equations, Cholesky-based trust-region sub-problem, accept /
// reject logic) run UNCHANGED on the host through the SIMT emulation in simt_emu.h -- several CTAs of real
// threads -- so that the whole S4 solve can be compared with the host-stepped model (ba_host.cpp, trf_core.h) and
// with scipy on a machine without a GPU.  NOT part of libmocap_b200.so and never used by the product path.
#include "simt_emu.h"
#include "../../low-cost-mocap_b200/csrc/ba_device.cuh"
#include "../../low-cost-mocap_b200/csrc/camera_tables.h"

// obs [m][C][2], mask [m][C], K [C][9], R [C][9] / t [C][3] in-out; report [11] = the mocap_ba_report fields
extern "C" int hc_ba_solve_dev(const double* obs, const uint8_t* mask, int m, int C, const double* K, double* R, double* t,
                               double ftol, int max_nfev, int jac_mode, int prefit, int prefit_max_iter, int n_ctas, int n_threads,
                               double* report) {
    static CameraTables T;
    memset(&T, 0, sizeof(T));
    build_camera_tables(T, C, K, R, t);
    const int n = 6 * (C - 1), npair = n * (n + 1) / 2, pstride = npair + 2 * n + 8;
    std::vector<double> X((size_t)m * 3), Xnew((size_t)m * 3), part((size_t)n_ctas * pstride), fin(pstride), cpart((size_t)2 * n_ctas * 4);
    std::vector<uint8_t> valid(m);
    unsigned bar[2] = {0, 0};
    mocap_ba_report rep;
    memset(&rep, 0, sizeof(rep));
    BAParams P;
    memset(&P, 0, sizeof(P));
    P.tb = &T; P.obs = obs; P.mask = mask; P.m_dev = nullptr; P.m_max = m; P.C = C; P.R = R; P.t = t;
    P.ftol = ftol; P.xtol = 1e-8; P.gtol = 1e-8; P.max_nfev = max_nfev; P.jac_mode = jac_mode; P.prefit = prefit;
    P.prefit_max_iter = prefit_max_iter;
    P.X = X.data(); P.Xnew = Xnew.data(); P.valid = valid.data(); P.part = part.data(); P.pstride = pstride; P.fin = fin.data();
    P.cpart = cpart.data(); P.bar = bar; P.report = &rep;
    const size_t smem = ba_smem_bytes(C, n_threads);
    std::vector<std::vector<unsigned char>> sm(n_ctas, std::vector<unsigned char>(smem + 16));
    simt::launch_grid(n_ctas, n_threads, [&] {
        unsigned char* base = sm[blockIdx.x].data();
        base += (16 - ((uintptr_t)base & 15)) & 15;
        ba_solve_body(P, base);
    });
    report[0] = rep.cost_initial; report[1] = rep.cost_final; report[2] = rep.optimality; report[3] = rep.n_iterations;
    report[4] = rep.n_fev; report[5] = rep.status; report[6] = rep.n_residuals; report[7] = rep.prefit_cost_initial;
    report[8] = rep.prefit_cost_final; report[9] = rep.prefit_iterations; report[10] = (double)smem;
    report[11] = rep.n_tr_solves; report[12] = rep.n_tr_newton;
    return 0;
}
