// hvp_cpu_bnb.cpp -- the STRONGEST CPU implementation of the per-vehicle local MIQP this repo has, for
// bench.py's cpu_baseline / `--impl reference` legs ONLY (test infrastructure; the product never loads it).
//
// What it is: the product's own branch-and-bound (depth-first over the PWA region sequence, relaxed
// unfixed stages, bounded-multiplier dual active-set QP, dual-bound pruning; the algorithm of
// hybrid_vehicle_platoon_b200/csrc/flat_core.cuh, which is written to compile for the host as well) WITH the
// two refinements that pay on a CPU (sibling bounds from the parent's dual, warm-started first children:
// FlatSolver<N, 1, ALG2 = true>; a third less work per solve, measured slower on the GPU and therefore not
// in the CUDA kernels), compiled with -O3 -march=native and run under OpenMP over every host core.
// Problem definition: fleet_decent_mld.py:61-208 on dmpcpwa MpcMld.
//
// Why it exists: the independent checker (hvp_oracle.c) enumerates every reachable mode sequence
// (85 leaf QPs per solve at N = 6) -- a fair CHECKER but an unfairly slow BASELINE.  The reference's
// solver (Gurobi behind dmpcpwa) is not installable here, so the honest CPU figure to quote a GPU speed-up
// against is the same 10-node branch-and-bound on the host cores.  It is itself checked against the
// enumeration oracle in tests/test_oracle_miqp.py::test_cpu_bnb_port_matches_enumeration.
#include <stdint.h>
#include <stddef.h>
#include <stdlib.h>
#ifdef _OPENMP
#include <omp.h>
#endif
#include "../hybrid_vehicle_platoon_b200/csrc/vehicle_model.h"
#include "../hybrid_vehicle_platoon_b200/csrc/flat_core.cuh"

template <int N>
static void bnb_batch(int batch, const int32_t* flags, double d0, double t0, double tight, const double* mass,
                      const double* x0, const double* xf, const double* xb, const double* xl, double* u, double* x,
                      int32_t* modes, double* obj, int32_t* status, int32_t* nodes, int32_t* qp_iters) {
    hvp::LocalParams P;
    hvp::fill_local_params(P, N, d0, t0, tight, 0);
    P.dive = 0; P.sibling_bound = 1; P.warm = 1;        // the variant that is fastest on a CPU (flat_core.cuh: ALG2)
    if (getenv("HVC_DIVE")) P.dive = atoi(getenv("HVC_DIVE"));                    // A/B switches for experiments
    if (getenv("HVC_SIBLING")) P.sibling_bound = atoi(getenv("HVC_SIBLING"));
    if (getenv("HVC_WARM")) P.warm = atoi(getenv("HVC_WARM"));
    const size_t S = 2 * (size_t)(N + 1);
#pragma omp parallel for schedule(dynamic, 64)
    for (int i = 0; i < batch; ++i) {
        hvp::FlatSolver<N, 1, true> sol;
        double W[hvp::FlatLayout<N>::SIZE];
        hvp::FlatCold<N> cold;
        sol.setup(W, &P, flags[i], mass[i], x0 + 2 * (size_t)i, xf ? xf + S * i : nullptr, xb ? xb + S * i : nullptr,
                  xl ? xl + S * i : nullptr, x + S * i + (N + 1) + 1, &cold);
        while (sol.state != hvp::FlatSolver<N, 1, true>::S_DONE) sol.trip();
        hvp::LocalResult R = sol.finish(u + (size_t)N * i, x + S * i, modes + (size_t)N * i);
        obj[i] = R.obj; status[i] = R.status; nodes[i] = R.nodes; qp_iters[i] = R.qp_iters;
    }
}

extern "C" int hvc_local_miqp_bnb(int batch, int N, const int32_t* flags, double d0, double t0, double tight,
                                  const double* mass, const double* x0, const double* xf, const double* xb,
                                  const double* xl, double* u, double* x, int32_t* modes, double* obj, int32_t* status,
                                  int32_t* nodes, int32_t* qp_iters) {
#define HVC_CASE(NN) case NN: bnb_batch<NN>(batch, flags, d0, t0, tight, mass, x0, xf, xb, xl, u, x, modes, obj, status, nodes, qp_iters); return 0;
    switch (N) { HVC_CASE(3) HVC_CASE(4) HVC_CASE(5) HVC_CASE(6) HVC_CASE(7) HVC_CASE(8) HVC_CASE(9) default: return -1; }
#undef HVC_CASE
}

extern "C" void hvc_set_threads(int n) {
#ifdef _OPENMP
    if (n >= 1) omp_set_num_threads(n);
#endif
}

extern "C" int hvc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
