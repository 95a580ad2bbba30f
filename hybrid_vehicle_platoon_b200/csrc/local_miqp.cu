// local_miqp.cu -- batched per-vehicle hybrid-MPC MIQPs on sm_100a: one thread per MIQP.
//
// Replaces the Gurobi solve behind LocalMpcMld.solve_mpc (fleet_decent_mld.py:21-223 /
// fleet_seq_mld.py:21-234 via dmpcpwa MpcMld.solve_mpc).  Each thread runs the whole
// branch-and-bound of miqp_core.cuh for one problem: the node QPs are 4..12-variable dense
// problems, far too small to spread over a CTA, so the parallel axis is the batch
// (scenario x vehicle x ADMM round) and a warp holds 32 independent trees.
#include <cuda_runtime.h>
#include <stdint.h>

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

cudaError_t launch_local_miqp(const LocalParams& P, int64_t batch, const int32_t* flags, const double* mass,
                              const double* x0, const double* xf, const double* xb, const double* xl,
                              double* u, double* x, int32_t* modes, double* obj, int32_t* status,
                              int32_t* nodes, int32_t* qp_iters, cudaStream_t stream) {
    if (batch <= 0) return cudaSuccess;
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
