// miqp_core.cuh -- per-vehicle local hybrid-MPC MIQP, solved by one GPU thread.
//
// Problem: LocalMpcMld of the reference (fleet_decent_mld.py:61-208, fleet_seq_mld.py:63-219) on
// the PWA-gear model (models.py:397-492) in MLD form (dmpcpwa MpcMld).
//
// B200-first formulation (NOT the reference's big-M model):
//   * variables are the VELOCITIES x_j = v_{j+1}, j = 0..N-1 (positions are prefix sums, inputs
//     are the affine map u_k = (v_{k+1} - a_k v_k - c_k)/b_k of the active region), so every
//     row is one of three structured kinds: a single entry (bounds), a two-entry difference
//     (accel / input limits) or a prefix sum (position box, safe distance).  Rows are never
//     materialised: products with them are O(1)..O(N) closed forms;
//   * the slack variables are eliminated exactly: w*s with s >= max(0, g(v)) is an L1 penalty,
//     i.e. a row whose multiplier is bounded by w (bounded-dual Goldfarb-Idnani);
//   * branch-and-bound over the region sequence, depth first, child nearest to the relaxed
//     velocity first; a node fixes the regions of stages 0..L-1 and relaxes stages L..N-1
//     (no input cost, no input bounds, only accel + state box) -- a valid lower bound because
//     every dropped term is non-negative and every dropped row only enlarges the feasible set;
//   * per-thread state is two packed symmetric matrices (H^-1 and the inverse of the active
//     normal matrix N'H^-1N, maintained by bordering updates) plus a few N-vectors: 129 doubles
//     at N = 6.  All of it lives in ONE strided work array W (element e of the calling thread is
//     W[e*ST]); on the GPU that array is shared memory with ST = 32, so a warp's access to
//     element e touches 32 consecutive doubles (conflict-free even when lanes use different e).
//
// The file is host+device so that tests can compile it with g++ (ST = 1) and check the algorithm
// against the oracle without a GPU (tests/host_harness); the product only ever runs it inside
// the CUDA kernels of local_miqp.cu.
#pragma once
#include <math.h>
#include <stdint.h>
#include <string.h>

#if defined(__CUDACC__)
#define HVP_HD __host__ __device__ __forceinline__
#define HVP_HDN __host__ __device__ __noinline__
#else
#define HVP_HD inline
#define HVP_HDN
#endif

namespace hvp {

// device clock in ns (%globaltimer); the host builds of the solvers (tests, CPU port) have no time limit
HVP_HD long long hvp_now_ns() {
#if defined(__CUDA_ARCH__)
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return (long long)t;
#else
    return 0;
#endif
}
// order-preserving 64-bit key of a double (atomicMin on keys = min of the doubles)
HVP_HD unsigned long long hvp_d2key(double d) {
    unsigned long long u;
    memcpy(&u, &d, 8);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
HVP_HD double hvp_key2d(unsigned long long k) {
    const unsigned long long u = (k >> 63) ? (k & 0x7fffffffffffffffull) : ~k;
    double d;
    memcpy(&d, &u, 8);
    return d;
}
// an objective must be below this to replace / beat the incumbent
HVP_HD double hvp_cut(double inc, double gap) {
    const double a = fabs(inc);
    return inc - fmax(1e-9 * fmax(1.0, a), gap * a);
}

constexpr int NREG = 7;

// Constants shared by every problem of a batch (filled on the host, see vehicle_model.h).
struct LocalParams {
    int N;
    int max_nodes;
    int hull;                            // flat kernel: interval-hull tightening of the unfixed stages
    int dive;                            // flat kernel: first descent solves only the leaf
    int sibling_bound;                   // flat kernel: prune children by the dual bound of their solved parent
    int warm;                            // flat kernel: warm-start a first child from its parent's active set
    int node_batch;                      // flat kernel: lanes that must wait for node set-up before a warp runs it
                                         // (27 of 32: measured optimum 26-28 after the r01 profile pass; 32 before it)
    const int32_t* hint;                 // [batch][N] device pointer or NULL: region sequence to try FIRST (hvp.h modes_hint)
    double mip_gap;                      // relative pruning gap (0: proven optimal)
    long long time_limit_ns;             // per-problem budget on the device clock (0: none)
    double d0, t0, tight;
    double qxp, qxv, qu, w;              // Params.Q_x, Q_u, w (common_controller_params.py:14-23)
    double a_acc, a_dec, d_safe;
    double vmin, vmax, pmin, pmax;       // D,E rows (models.py:474-475)
    double umin, umax;                   // F,G rows (models.py:476-479)
    double edge[NREG + 1];               // region r is edge[r] <= v <= edge[r+1] (models.py:414-444)
    double bgear[NREG];                  // traction gain of region r before /m (models.py:458-466)
    double c1, c2, dfr, mug;             // PWA friction (models.py:276-282), mu*g
    // flat kernel: inverse of the pure tracking Hessian (no stage fixed) for the three cost structures a vehicle
    // can have -- [0] one reference term without headway (leader, or back neighbour only), [1] front neighbour
    // only, [2] front and back neighbour -- row-major N x N; filled on the host (vehicle_model.h)
    double h0inv[3][81];
};

struct LocalResult {
    double obj;
    int status;
    int nodes;
    int qp_iters;
};

#define HVP_ST_OPTIMAL 2
#define HVP_ST_INFEASIBLE 3
#define HVP_ST_NODE_LIMIT 8
#define HVP_ST_TIME_LIMIT 9
#define HVP_ST_NUMERIC 12

// constraint types of the velocity-space node QP (all written as  n'x <= rhs)
enum : int { T_UB = 0, T_LB, T_ACC, T_DEC, T_UHI, T_ULO, T_PHI, T_PLO, T_SF, T_SB, T_COUNT };

}  // namespace hvp
