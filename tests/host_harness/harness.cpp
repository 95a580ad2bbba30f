// Host build of the product's per-thread MIQP core (csrc/miqp_core.cuh) -- TEST HARNESS ONLY.
// Lets the CPU test-suite check the branch-and-bound / QP algorithm against the oracle without
// a GPU.  The product package never loads this library; its only execution path is CUDA.
#include <stdint.h>
#include "../../hybrid_vehicle_platoon_b200/csrc/vehicle_model.h"
#include "../../hybrid_vehicle_platoon_b200/csrc/coop_core.cuh"
#include "../../hybrid_vehicle_platoon_b200/csrc/flat_core.cuh"

extern "C" void hvh_local_miqp_batch(int batch, int N, const int32_t* flags, double d0, double t0,
                                     double tight, int max_nodes, const double* mass, const double* x0,
                                     const double* xf, const double* xb, const double* xl, double* u,
                                     double* x, int32_t* modes, double* obj, int32_t* status,
                                     int32_t* nodes, int32_t* qp_iters) {
    hvp::LocalParams P;
    hvp::fill_local_params(P, N, d0, t0, tight, max_nodes);
    size_t S = 2 * (size_t)(N + 1);
    for (int i = 0; i < batch; ++i) {
        hvp::LocalSolver<12, 1> sol;
        double W[hvp::LocalLayout<12>::SIZE];
        sol.setup(W, &P, flags[i], mass[i], x0 + 2 * (size_t)i, xf ? xf + S * i : nullptr,
                  xb ? xb + S * i : nullptr, xl ? xl + S * i : nullptr);
        hvp::LocalResult R = sol.solve(u + (size_t)N * i, x + S * i, modes + (size_t)N * i);
        obj[i] = R.obj; status[i] = R.status; nodes[i] = R.nodes; qp_iters[i] = R.qp_iters;
    }
}

// Host emulation of the cooperative (G lanes per problem) solver, coop_core.cuh.
template <int G>
static void coop_batch(int batch, int N, const int32_t* flags, double d0, double t0, double tight, int max_nodes,
                       const double* mass, const double* x0, const double* xf, const double* xb, const double* xl,
                       double* u, double* x, int32_t* modes, double* obj, int32_t* status, int32_t* nodes,
                       int32_t* qp_iters) {
    hvp::LocalParams P;
    hvp::fill_local_params(P, N, d0, t0, tight, max_nodes);
    size_t S = 2 * (size_t)(N + 1);
    for (int i = 0; i < batch; ++i) {
        hvp::CoopSolver<hvp::HostBK<G>> sol;
        sol.setup(&P, flags[i], mass[i], x0 + 2 * (size_t)i, xf ? xf + S * i : nullptr, xb ? xb + S * i : nullptr,
                  xl ? xl + S * i : nullptr);
        hvp::LocalResult R = sol.solve(u + (size_t)N * i, x + S * i, modes + (size_t)N * i);
        obj[i] = R.obj; status[i] = R.status; nodes[i] = R.nodes; qp_iters[i] = R.qp_iters;
    }
}

extern "C" void hvh_coop_miqp_batch(int batch, int N, const int32_t* flags, double d0, double t0, double tight,
                                    int max_nodes, const double* mass, const double* x0, const double* xf,
                                    const double* xb, const double* xl, double* u, double* x, int32_t* modes,
                                    double* obj, int32_t* status, int32_t* nodes, int32_t* qp_iters) {
    if (N <= 8) coop_batch<8>(batch, N, flags, d0, t0, tight, max_nodes, mass, x0, xf, xb, xl, u, x, modes, obj, status, nodes, qp_iters);
    else coop_batch<16>(batch, N, flags, d0, t0, tight, max_nodes, mass, x0, xf, xb, xl, u, x, modes, obj, status, nodes, qp_iters);
}

// Host build of the flat state-machine solver (flat_core.cuh), one problem at a time.
template <int N>
static void flat_batch(int batch, const int32_t* flags, double d0, double t0, double tight, int max_nodes,
                       const double* mass, const double* x0, const double* xf, const double* xb, const double* xl,
                       double* u, double* x, int32_t* modes, double* obj, int32_t* status, int32_t* nodes,
                       int32_t* qp_iters) {
    hvp::LocalParams P;
    hvp::fill_local_params(P, N, d0, t0, tight, max_nodes);
    size_t S = 2 * (size_t)(N + 1);
    for (int i = 0; i < batch; ++i) {
        hvp::FlatSolver<N, 1> sol;
        double W[hvp::FlatLayout<N>::SIZE];
        hvp::FlatCold<N> cold;
        sol.setup(W, &P, flags[i], mass[i], x0 + 2 * (size_t)i, xf ? xf + S * i : nullptr,
                  xb ? xb + S * i : nullptr, xl ? xl + S * i : nullptr, x + S * i + (N + 1) + 1, &cold);
        while (sol.state != hvp::FlatSolver<N, 1>::S_DONE) sol.trip();
        hvp::LocalResult R = sol.finish(u + (size_t)N * i, x + S * i, modes + (size_t)N * i);
        obj[i] = R.obj; status[i] = R.status; nodes[i] = R.nodes; qp_iters[i] = R.qp_iters;
    }
}

extern "C" int hvh_flat_miqp_batch(int batch, int N, const int32_t* flags, double d0, double t0, double tight,
                                   int max_nodes, const double* mass, const double* x0, const double* xf,
                                   const double* xb, const double* xl, double* u, double* x, int32_t* modes,
                                   double* obj, int32_t* status, int32_t* nodes, int32_t* qp_iters) {
#define HVH_CASE(NN) case NN: flat_batch<NN>(batch, flags, d0, t0, tight, max_nodes, mass, x0, xf, xb, xl, u, x, modes, obj, status, nodes, qp_iters); return 0;
    switch (N) { HVH_CASE(2) HVH_CASE(3) HVH_CASE(4) HVH_CASE(5) HVH_CASE(6) HVH_CASE(7) HVH_CASE(8) HVH_CASE(9) default: return -1; }
#undef HVH_CASE
}
