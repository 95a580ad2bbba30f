// Host build of the product's per-thread MIQP core (csrc/miqp_core.cuh) -- TEST HARNESS ONLY.
// Lets the CPU test-suite check the branch-and-bound / QP algorithm against the oracle without
// a GPU.  The product package never loads this library; its only execution path is CUDA.
#include <stdint.h>
#include "../../hybrid_vehicle_platoon_b200/csrc/vehicle_model.h"
#include "../../hybrid_vehicle_platoon_b200/csrc/coop_core.cuh"
#include "../../hybrid_vehicle_platoon_b200/csrc/flat_core.cuh"
#include "scalar_solver.h"

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

// Host emulation of the split search of the cooperative solver (CoopSolver::sub_M, local_miqp.cu coop_split_kernel): the
// M workers of a tree run one after the other WITHOUT exchanging their incumbents (the worst case for pruning); the minimum
// over the workers must be the optimum of the plain search, and every sub-tree must have exactly one owner.
extern "C" void hvh_coop_split_batch(int batch, int N, int M, const int32_t* flags, double d0, double t0, double tight,
                                     const double* mass, const double* x0, const double* xf, const double* xb,
                                     const double* xl, double* obj, int32_t* nodes_sum, int32_t* nodes_max) {
    hvp::LocalParams P;
    hvp::fill_local_params(P, N, d0, t0, tight, 0);
    size_t S = 2 * (size_t)(N + 1);
    double u[16], x[40];
    int32_t modes[16];
    for (int i = 0; i < batch; ++i) {
        double best = HUGE_VAL;
        int ns = 0, nm = 0;
        for (int w = 0; w < M; ++w) {
            hvp::CoopSolver<hvp::HostBK<8>> sol;
            sol.setup(&P, flags[i], mass[i], x0 + 2 * (size_t)i, xf ? xf + S * i : nullptr, xb ? xb + S * i : nullptr,
                      xl ? xl + S * i : nullptr);
            sol.sub_M = M; sol.sub_w = w;
            hvp::LocalResult R = sol.solve(u, x, modes);
            if (R.obj < best) best = R.obj;
            ns += R.nodes; if (R.nodes > nm) nm = R.nodes;
        }
        obj[i] = best; nodes_sum[i] = ns; nodes_max[i] = nm;
    }
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

// Host check of the sub-tree adoption of the flat kernel's tail (flat_core.cuh: open_level / adopt_prefix): the owner
// runs `after_nodes` nodes, then hands EVERY open region of its shallowest open level but (if nothing else is open)
// one to fresh solvers, each of which sets the same problem up, follows the owner's prefix and searches its branch from
// the owner's incumbent; the owner finishes the rest.  Returns the best objective over owner and adopters and the
// number of adopters; the caller compares it with the plain solve.
template <int N>
static double flat_adopt_one(int flags, double d0, double t0, double tight, double mass, const double* x0, const double* xf,
                             const double* xb, const double* xl, int after_nodes, int* n_adopters, int* total_nodes) {
    hvp::LocalParams P;
    hvp::fill_local_params(P, N, d0, t0, tight, 0);
    using Sol = hvp::FlatSolver<N, 1>;
    Sol own;
    double W[hvp::FlatLayout<N>::SIZE], best[N];
    hvp::FlatCold<N> cold;
    own.setup(W, &P, flags, mass, x0, xf, xb, xl, best, &cold);
    *n_adopters = 0;
    double result = HUGE_VAL;
    int nodes = 0;
    bool donated = false;
    while (own.state != Sol::S_DONE) {
        if (!donated && own.state == Sol::S_NEXT && own.nodes >= after_nodes) {
            int l, cb, total;
            if (own.open_level(l, cb, total) && total >= 2) {
                int give = cb, keep = 0;
                if (total == __builtin_popcount((unsigned)cb)) { keep = cb & -cb; give = cb & ~keep; }
                for (int c = 0; c < 7; ++c) {
                    if (!((give >> c) & 1)) continue;
                    Sol th;
                    double W2[hvp::FlatLayout<N>::SIZE], best2[N];
                    hvp::FlatCold<N> cold2;
                    th.setup(W2, &P, flags, mass, x0, xf, xb, xl, best2, &cold2);
                    th.adopt_prefix(own.modes_pk, l, c, own.inc);
                    while (th.state != Sol::S_DONE) th.trip();
                    if (th.inc < result) result = th.inc;
                    nodes += th.nodes;
                    ++*n_adopters;
                }
                own.set_cand(l, cb & ~give);
                donated = true;
            }
        }
        own.trip();
    }
    if (own.inc < result) result = own.inc;
    *total_nodes = nodes + own.nodes;
    return result;
}

extern "C" int hvh_flat_adopt_batch(int batch, int N, const int32_t* flags, double d0, double t0, double tight,
                                    const double* mass, const double* x0, const double* xf, const double* xb,
                                    const double* xl, int after_nodes, double* obj, int32_t* adopters, int32_t* nodes) {
    const size_t S = 2 * (size_t)(N + 1);
    for (int i = 0; i < batch; ++i) {
        int na = 0, nn = 0;
        double r;
#define HVH_CASE(NN) case NN: r = flat_adopt_one<NN>(flags[i], d0, t0, tight, mass[i], x0 + 2 * (size_t)i, xf + S * i, xb + S * i, xl + S * i, after_nodes, &na, &nn); break;
        switch (N) { HVH_CASE(3) HVH_CASE(4) HVH_CASE(5) HVH_CASE(6) HVH_CASE(7) HVH_CASE(8) default: return -1; }
#undef HVH_CASE
        obj[i] = r; adopters[i] = na; nodes[i] = nn;
    }
    return 0;
}
