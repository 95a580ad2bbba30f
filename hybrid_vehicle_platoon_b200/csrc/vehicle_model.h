// vehicle_model.h -- host-side constants of the vehicle / PWA model and controller parameters.
// Values restate models.py:13-28,57-92,276-282,397-492 and misc/common_controller_params.py:14-23
// of the reference; tests/test_host_logic.py::test_model_tables_match_reference checks them against tables captured from the
// reference itself (tests/golden/rollout_golden.npz).
#pragma once
#include <math.h>
#include "miqp_core.cuh"

namespace hvp {

struct VehicleModel {
    // true traction curve (models.py:13-28); t[3][0] = 100/415 is the reference's own typo (Q3)
    double tr_t[6][3] = {{253.54, 4056.7, 3042},  {184, 2944.75, 2208.55}, {132.22, 2115.6, 1586.7},
                         {100.0 / 415, 1605, 1205}, {72.88, 1166, 874.7},   {52.4, 838, 628.3}};
    double tr_v[6][4] = {{2.0706, 4.12158, 9.29, 12.38},     {2.85, 5.675, 12.7956, 17.06},
                         {3.9705, 7.90316, 17.8105, 23.7474}, {5.228, 10.42, 23.454, 31.2704},
                         {7.203, 14.335, 32.31, 43.0802},     {10.027, 19.956, 44.978, 59.9715}};
    double c_fric = 0.5, mu = 0.01, grav = 9.8;
    double bgear[6] = {4057, 2945, 2116, 1607, 1166, 838};
    double vl[6] = {3.94, 5.43, 7.56, 9.96, 13.70, 19.10};
    double vh[6] = {9.46, 13.04, 18.15, 23.90, 32.93, 45.84};
    double v_min = 3.94, v_max = 45.84, p_min = 0.0, p_max = 10000.0, u_min = -1.0, u_max = 1.0;
    // gear-switch velocities (models.py:401-403)
    void gear_limits(double lim[5]) const {
        for (int i = 1; i < 6; ++i) lim[i - 1] = (vh[i] - vl[i]) / 2 + vl[i];
    }
};

// Controller parameters (Params, misc/common_controller_params.py:14-23) + PWA-gear region data.
inline void fill_local_params(LocalParams& P, int N, double d0, double t0, double tight, int max_nodes,
                              double mip_gap = 0.0, double time_limit_ms = 0.0) {
    VehicleModel M;
    P.N = N; P.hint = nullptr; P.max_nodes = max_nodes; P.mip_gap = mip_gap; P.time_limit_ns = (long long)(time_limit_ms * 1e6); P.hull = 0; P.dive = 1; P.node_batch = 27; P.sibling_bound = 1; P.warm = 1; P.d0 = d0; P.t0 = t0; P.tight = tight;
    P.qxp = 1.0; P.qxv = 0.1; P.qu = 1.0; P.w = 1e4;
    P.a_acc = 2.5; P.a_dec = -2.0; P.d_safe = 25.0;
    P.vmin = M.v_min; P.vmax = M.v_max; P.pmin = M.p_min; P.pmax = M.p_max;
    P.umin = M.u_min; P.umax = M.u_max;
    double beta = (3 * M.c_fric * M.v_max * M.v_max) / 16;   // models.py:276-282
    double alpha = M.v_max / 2;
    P.c1 = beta / alpha;
    P.c2 = (M.c_fric * M.v_max * M.v_max - beta) / (M.v_max - alpha);
    P.dfr = beta - alpha * ((M.c_fric * M.v_max * M.v_max - beta) / (M.v_max - alpha));
    P.mug = M.mu * M.grav;
    double lim[5];
    M.gear_limits(lim);
    const double e[NREG + 1] = {-HUGE_VAL, lim[0], lim[1], lim[2], alpha, lim[3], lim[4], HUGE_VAL};
    for (int i = 0; i <= NREG; ++i) P.edge[i] = e[i];
    const int gear_of_region[NREG] = {0, 1, 2, 3, 3, 4, 5};  // models.py:458-466
    for (int r = 0; r < NREG; ++r) P.bgear[r] = M.bgear[gear_of_region[r]];
    // inverse tracking Hessians (flat_core.cuh setup): H[i][j] = hw1 (N-1-max(i,j)) + (i == j ? hd : hw2)
    for (int var = 0; var < 3; ++var) {
        const double nterm = var == 2 ? 2.0 : 1.0, tf = var >= 1 ? 1.0 : 0.0;
        const double hw1 = 2.0 * P.qxp * nterm, hw2 = 2.0 * P.qxp * (tf * t0);
        const double hd = 2.0 * P.qxp * (tf * t0 * t0) + 2.0 * P.qxv * nterm;
        double A[81];
        for (int e = 0; e < 81; ++e) P.h0inv[var][e] = 0.0;
        if (N < 1 || N > 9) continue;
        for (int i = 0; i < N; ++i)
            for (int j = 0; j < N; ++j) A[i * N + j] = hw1 * (double)(N - 1 - (i > j ? i : j)) + (i == j ? hd : hw2);
        for (int k = 0; k < N; ++k) {                      // in-place Gauss-Jordan (SPD: no pivoting needed)
            const double pinv = 1.0 / A[k * N + k];
            for (int j = 0; j < N; ++j) A[k * N + j] *= pinv;
            A[k * N + k] = pinv;
            for (int i = 0; i < N; ++i) {
                if (i == k) continue;
                const double f = A[i * N + k];
                for (int j = 0; j < N; ++j) A[i * N + j] = (j == k) ? -f * pinv : A[i * N + j] - f * A[k * N + j];
            }
        }
        for (int i = 0; i < N; ++i)                        // symmetrise the round-off
            for (int j = 0; j < N; ++j) P.h0inv[var][i * N + j] = 0.5 * (A[i * N + j] + A[j * N + i]);
    }
}

}  // namespace hvp
