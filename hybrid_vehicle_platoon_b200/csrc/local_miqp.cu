// local_miqp.cu -- batched per-vehicle hybrid-MPC MIQPs on sm_100a.
//
// Replaces the Gurobi solve behind LocalMpcMld.solve_mpc (fleet_decent_mld.py:21-223 /
// fleet_seq_mld.py:21-234 via dmpcpwa MpcMld.solve_mpc).
//
//   coop_miqp_kernel<G>  (default): a group of G = 8 (N <= 8) or 16 lanes solves one MIQP
//       cooperatively, all state in registers, warp shuffles for the reductions (coop_core.cuh).
//       The node QPs have 4..12 variables -- far too small for a CTA -- so a warp carries 4 (or 2)
//       independent branch-and-bound trees and the batch (scenario x vehicle x ADMM round)
//       fills the grid.
//   local_miqp_kernel<NMAX> (HVP_LOCAL_KERNEL=scalar): the first design, one thread per MIQP
//       with its state in shared memory (miqp_core.cuh); kept for A/B measurements -- ncu showed
//       5.8 of 32 lanes active per instruction because every lane runs its own tree.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "coop_core.cuh"
#include "hvp_internal.h"
#include "miqp_core.cuh"

namespace hvp {

template <int NMAX>
__global__ void __launch_bounds__(LOCAL_BLOCK, 6)
local_miqp_kernel(const __grid_constant__ LocalParams P, int64_t batch, const int32_t* __restrict__ flags,
                  const double* __restrict__ mass, const double* __restrict__ x0,
                  const double* __restrict__ xf, const double* __restrict__ xb,
                  const double* __restrict__ xl, double* __restrict__ u, double* __restrict__ x,
                  int32_t* __restrict__ modes, double* __restrict__ obj, int32_t* __restrict__ status,
                  int32_t* __restrict__ nodes, int32_t* __restrict__ qp_iters) {
    extern __shared__ double smem[];
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= batch) return;
    const int N = P.N;
    const size_t S = 2 * (size_t)(N + 1);
    // per-warp slab of LocalLayout::SIZE x 32 doubles; element e of lane l is slab[e*32 + l]
    double* W = smem + (size_t)(threadIdx.x >> 5) * (LocalLayout<NMAX>::SIZE * 32) + (threadIdx.x & 31);
    LocalSolver<NMAX, 32> sol;
    sol.setup(W, &P, flags[i], mass[i], x0 + 2 * i, xf ? xf + S * i : nullptr, xb ? xb + S * i : nullptr,
              xl ? xl + S * i : nullptr);
    LocalResult R = sol.solve(u + (size_t)N * i, x + S * i, modes + (size_t)N * i);
    obj[i] = R.obj;
    status[i] = R.status;
    nodes[i] = R.nodes;
    if (qp_iters) qp_iters[i] = R.qp_iters;
}

template <int G>
__global__ void __launch_bounds__(COOP_BLOCK)
coop_miqp_kernel(const __grid_constant__ LocalParams P, int64_t batch, const int32_t* __restrict__ flags,
                 const double* __restrict__ mass, const double* __restrict__ x0,
                 const double* __restrict__ xf, const double* __restrict__ xb,
                 const double* __restrict__ xl, double* __restrict__ u, double* __restrict__ x,
                 int32_t* __restrict__ modes, double* __restrict__ obj, int32_t* __restrict__ status,
                 int32_t* __restrict__ nodes, int32_t* __restrict__ qp_iters) {
    const int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G;   // problem of this group
    if (i >= batch) return;                                                   // whole group leaves
    const int N = P.N;
    const size_t S = 2 * (size_t)(N + 1);
    CoopSolver<DevBK<G>> sol;
    sol.setup(&P, flags[i], mass[i], x0 + 2 * i, xf ? xf + S * i : nullptr, xb ? xb + S * i : nullptr,
              xl ? xl + S * i : nullptr);
    const LocalResult R = sol.solve(u + (size_t)N * i, x + S * i, modes + (size_t)N * i);
    if (sol.bk.ln == 0) {
        obj[i] = R.obj;
        status[i] = R.status;
        nodes[i] = R.nodes;
        if (qp_iters) qp_iters[i] = R.qp_iters;
    }
}

static bool use_scalar_kernel() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("HVP_LOCAL_KERNEL");
        v = (e && strcmp(e, "scalar") == 0) ? 1 : 0;
    }
    return v == 1;
}

cudaError_t launch_local_miqp(const LocalParams& P, int64_t batch, const int32_t* flags, const double* mass,
                              const double* x0, const double* xf, const double* xb, const double* xl,
                              double* u, double* x, int32_t* modes, double* obj, int32_t* status,
                              int32_t* nodes, int32_t* qp_iters, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
    if (!use_scalar_kernel()) {
        if (P.N <= 8) {
            const int per_block = COOP_BLOCK / 8;
            const unsigned g = (unsigned)((batch + per_block - 1) / per_block);
            coop_miqp_kernel<8><<<g, COOP_BLOCK, 0, stream>>>(P, batch, flags, mass, x0, xf, xb, xl, u, x, modes, obj,
                                                             status, nodes, qp_iters);
        } else {
            const int per_block = COOP_BLOCK / 16;
            const unsigned g = (unsigned)((batch + per_block - 1) / per_block);
            coop_miqp_kernel<16><<<g, COOP_BLOCK, 0, stream>>>(P, batch, flags, mass, x0, xf, xb, xl, u, x, modes, obj,
                                                              status, nodes, qp_iters);
        }
        return cudaGetLastError();
    }
    const unsigned grid = (unsigned)((batch + LOCAL_BLOCK - 1) / LOCAL_BLOCK);
#define HVP_LAUNCH(NM)                                                                                   \
    {                                                                                                    \
        const size_t smem = (size_t)(LOCAL_BLOCK / 32) * LocalLayout<NM>::SIZE * 32 * sizeof(double);    \
        static bool configured = false;                                                                  \
        if (!configured) {                                                                               \
            cudaError_t e = cudaFuncSetAttribute(local_miqp_kernel<NM>,                                  \
                                                 cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (e != cudaSuccess) return e;                                                              \
            e = cudaFuncSetAttribute(local_miqp_kernel<NM>, cudaFuncAttributePreferredSharedMemoryCarveout, \
                                     cudaSharedmemCarveoutMaxShared);                                    \
            if (e != cudaSuccess) return e;                                                              \
            configured = true;                                                                           \
        }                                                                                                \
        local_miqp_kernel<NM><<<grid, LOCAL_BLOCK, smem, stream>>>(P, batch, flags, mass, x0, xf, xb, xl, u, \
                                                                   x, modes, obj, status, nodes, qp_iters); \
    }
    if (P.N <= 6) HVP_LAUNCH(6)
    else if (P.N <= 8) HVP_LAUNCH(8)
    else HVP_LAUNCH(12)
#undef HVP_LAUNCH
    return cudaGetLastError();
}

}  // namespace hvp
